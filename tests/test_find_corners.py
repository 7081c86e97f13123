"""board_detection.find_chessboard_corners (board_detection.py:4-27, calibration time): gray ->
GaussianBlur(7x7, 1) -> Canny(30, 100) -> dilate(5x5, 3) on the device, contour logic on the host.
CPU: oracle vs the committed cv2 / reference results; GPU: kernels vs those and vs the oracle."""
import hashlib
import importlib
import json
import os
import sys

import numpy as np
import pytest

from chessboard_vision_b200 import synth
import chessboard_vision_b200.dropin as dropin

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
GOLD = json.load(open(os.path.join(G, "find_corners.json")))
SCENES = {"1080x1920_1": (1080, 1920, 1), "720x1280_2": (720, 1280, 2), "1080x1920_3": (1080, 1920, 3)}


def _small():
    return synth.noise_frame(97, 133, 4)[:, :, 1].copy()


def test_oracle_blur_dilate_and_mask_match_cv2_digests(oracle):
    s = _small()
    assert sha(oracle.gaussian(s, 7, 1)) == GOLD["blur_7_1_97x133"]
    assert sha(oracle.gaussian(s, 9, 2.5)) == GOLD["blur_9_2.5_97x133"]
    assert sha(oracle.dilate(s, 5, 5, 3)) == GOLD["dilate_5x5x3_97x133"]
    assert sha(oracle.dilate(s, 3, 7, 2)) == GOLD["dilate_7x3x2_97x133"]
    assert list(oracle.gaussian_kernel_q8_sigma(7, 1.0)) == [1, 14, 62, 102, 62, 14, 1]
    assert list(oracle.gaussian_kernel_q8_sigma(7, 0.0)) == list(oracle.gaussian_kernel_q8(7))
    key = "720x1280_2"
    im = synth.table_scene(*SCENES[key])
    m = oracle.contour_mask(im)
    assert sha(m) == GOLD[key]["mask_sha"] and int(np.count_nonzero(m)) == GOLD[key]["mask_px"]
    with pytest.raises(ValueError):
        oracle.dilate(s, 4, 5, 1)


def _load_dropin(engine_factory):
    import chessboard_vision_b200.engine as engine_mod
    saved = engine_mod._default.get(0)
    engine_mod._default[0] = engine_factory
    sys.path.insert(0, dropin.PATH)
    sys.modules.pop("board_detection", None)
    mod = importlib.import_module("board_detection")
    return mod, engine_mod, saved


def _unload(engine_mod, saved):
    sys.path.remove(dropin.PATH)
    sys.modules.pop("board_detection", None)
    if saved is None:
        engine_mod._default.pop(0, None)
    else:
        engine_mod._default[0] = saved


def test_dropin_corners_with_oracle_engine():
    pytest.importorskip("cv2")
    import fake_engine
    mod, em, saved = _load_dropin(fake_engine.FakeEngine())
    try:
        key = "720x1280_2"
        got = mod.find_chessboard_corners(synth.table_scene(*SCENES[key]))
        assert got.shape == (4, 1, 2) and got.reshape(-1, 2).tolist() == GOLD[key]["corners"]
        assert mod.find_chessboard_corners(synth.board_frame(270, 480, 3)).reshape(-1, 2).tolist() == GOLD["no_board_270x480"]["corners"]
        assert mod.find_chessboard_corners(np.full((64, 64, 3), 90, np.uint8)).size == 0        # nothing to find
        with pytest.raises(ValueError):
            mod.find_chessboard_corners(np.zeros((10, 10), np.uint8))
    finally:
        _unload(em, saved)


@pytest.mark.gpu
def test_gpu_blur_sigma_dilate_vs_cv2_digests_and_oracle(engine, oracle):
    s = _small()
    assert sha(engine.gaussian(s, 7, 1)) == GOLD["blur_7_1_97x133"]
    assert sha(engine.gaussian(s, 9, 2.5)) == GOLD["blur_9_2.5_97x133"]
    assert sha(engine.dilate(s, 5, 5, 3)) == GOLD["dilate_5x5x3_97x133"]
    assert sha(engine.dilate(s, 3, 7, 2)) == GOLD["dilate_7x3x2_97x133"]
    batch = synth.frame_batch(3, 61, 203, "noise", 8)[..., 0].copy()
    got = engine.dilate(batch, 9, 1, 2)
    for i in range(3):
        assert np.array_equal(got[i], oracle.dilate(batch[i], 9, 1, 2))
    for k, sg in ((3, 0.8), (11, 1.3), (5, 0.0), (31, 6.0)):
        assert np.array_equal(engine.gaussian(batch[0], k, sg), oracle.gaussian(batch[0], k, sg))
    with pytest.raises(ValueError):
        engine.dilate(s, 4, 4, 1)
    with pytest.raises(ValueError):
        engine.dilate(s, 31, 31, 3)           # radius 45 > 24


@pytest.mark.gpu
def test_gpu_contour_mask_and_corners_match_reference(engine, oracle):
    pytest.importorskip("cv2")
    mod, em, saved = _load_dropin(engine)
    try:
        for key, (H, W, seed) in SCENES.items():
            im = synth.table_scene(H, W, seed)
            m = engine.contour_mask(im)
            assert sha(m) == GOLD[key]["mask_sha"] and int(np.count_nonzero(m)) == GOLD[key]["mask_px"], key
            got = mod.find_chessboard_corners(im)
            assert got.reshape(-1, 2).tolist() == GOLD[key]["corners"], key
        odd = synth.board_frame(133, 251, 5)
        assert np.array_equal(engine.contour_mask(odd), oracle.contour_mask(odd))
        assert mod.find_chessboard_corners(synth.board_frame(270, 480, 3)).reshape(-1, 2).tolist() == GOLD["no_board_270x480"]["corners"]
        assert mod.find_chessboard_corners(np.full((64, 64, 3), 90, np.uint8)).size == 0
    finally:
        _unload(em, saved)
