"""Step 0 of process_pipeline (apply_color_profile, frame_enhancer.py:56-99).
CPU part: the oracle against golden vectors of the unmodified reference and against live cv2.
GPU part (marked): the kernel through the C ABI and the drop-in class against the same vectors."""
import hashlib
import importlib
import json
import os
import sys

import numpy as np
import pytest

from chessboard_vision_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
FRAMES = {"board_90x121": lambda: synth.board_frame(90, 121, 21), "noise_64x96": lambda: synth.noise_frame(64, 96, 22)}


@pytest.fixture(scope="module")
def golden():
    z = np.load(os.path.join(G, "color_profile.npz"))
    return z, json.loads(str(z["profiles"])), json.loads(str(z["kat"]))


def test_oracle_vs_reference(oracle, golden):
    z, profiles, kat = golden
    for pname, prof in profiles.items():
        for fname, gen in FRAMES.items():
            assert np.array_equal(oracle.apply_color_profile(gen(), prof), z[pname + "/" + fname]), (pname, fname)
        assert sha(oracle.apply_color_profile(synth.noise_frame(1080, 1920, 0), prof)) == kat[pname]["apply_1920x1080"]
        # whole chain with the profile on: only the bilateral's <= 1 LSB (x9 through the sharpen kernel) may differ
        full = oracle.process_pipeline(oracle.apply_color_profile(synth.board_frame(96, 128, 1), prof), True)
        assert np.abs(full.astype(int) - z[pname + "/pipeline_board_96x128"]).max() <= 9
    assert oracle.apply_color_profile(synth.board_frame(8, 8, 0), {}) is not None


def test_oracle_components_vs_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    a = np.arange(1 << 24, dtype=np.uint32)
    allc = np.stack([a & 255, (a >> 8) & 255, (a >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(oracle.bgr2hsv(allc), cv2.cvtColor(allc, cv2.COLOR_BGR2HSV))
    rng = np.random.default_rng(1)
    for W in (64, 100, 97, 643, 33, 31, 1920):        # vector body (32-pixel blocks) and scalar row tails
        hsv = np.stack([rng.integers(0, 180, (23, W)), rng.integers(0, 256, (23, W)), rng.integers(0, 256, (23, W))], -1).astype(np.uint8)
        assert np.array_equal(oracle.hsv2bgr(hsv, 32), cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)), W
    h, s, v = np.meshgrid(np.arange(180), np.arange(256), np.arange(256), indexing="ij")
    hsv = np.stack([h, s, v], -1).reshape(180 * 256, 256, 3).astype(np.uint8)      # every (h, s, v), vector body
    assert np.array_equal(oracle.hsv2bgr(hsv, 32), cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR))
    src = np.resize(np.arange(256, dtype=np.uint8), (3, 1000))
    for alpha, beta in [(1.48, -30), (0.5, 10), (2.46, 0), (-1.3, 40), (0.01, 0.5), (3.7, -200.25)]:
        assert np.array_equal(oracle.convert_scale_abs(src, alpha, beta), cv2.convertScaleAbs(src, alpha=alpha, beta=beta))


def test_dropin_host_logic_with_profile(oracle, golden, monkeypatch, tmp_path):
    """ImageEnhancer picks color_profile.json up from the working directory (frame_enhancer.py:46-54)."""
    import chessboard_vision_b200.engine as engine_mod
    import chessboard_vision_b200.dropin as dropin
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_engine import FakeEngine
    z, profiles, _ = golden
    monkeypatch.setitem(engine_mod._default, 0, FakeEngine())
    monkeypatch.syspath_prepend(dropin.PATH)
    monkeypatch.chdir(tmp_path)
    json.dump(profiles["repo"], open("color_profile.json", "w"))
    sys.modules.pop("frame_enhancer", None)
    fe = importlib.import_module("frame_enhancer")
    try:
        e = fe.ImageEnhancer()
        assert e.profile == profiles["repo"]
        assert np.array_equal(e.apply_color_profile(FRAMES["board_90x121"]()), z["repo/board_90x121"])
        full = e.process_pipeline(synth.board_frame(96, 128, 1))
        assert np.abs(full.astype(int) - z["repo/pipeline_board_96x128"]).max() <= 9
    finally:
        sys.modules.pop("frame_enhancer", None)


@pytest.mark.gpu
def test_gpu_vs_reference_and_oracle(engine, oracle, golden):
    z, profiles, kat = golden
    for pname, prof in profiles.items():
        for fname, gen in FRAMES.items():
            assert np.array_equal(engine.apply_color_profile(gen(), prof), z[pname + "/" + fname]), (pname, fname)
        assert sha(engine.apply_color_profile(synth.noise_frame(1080, 1920, 0), prof)) == kat[pname]["apply_1920x1080"]
        img = synth.board_frame(96, 128, 1)
        full = engine.process_pipeline(img, engine.enhance_params(profile=prof))
        assert np.array_equal(full, oracle.process_pipeline(oracle.apply_color_profile(img, prof), True))
        assert np.abs(full.astype(int) - z[pname + "/pipeline_board_96x128"]).max() <= 9
    for shape in ((135, 241), (31, 33), (64, 64)):                         # row tails and batches
        batch = synth.frame_batch(2, *shape, "noise", 5)
        got = engine.apply_color_profile(batch, profiles["radical"])
        for i in range(2):
            assert np.array_equal(got[i], oracle.apply_color_profile(batch[i], profiles["radical"]))
