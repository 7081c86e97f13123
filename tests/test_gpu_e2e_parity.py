"""GPU: end-to-end parity of the CUDA path with the reference on whole frames (the stage-isolated tests live in
test_gpu_enhance.py / test_gpu_golden.py):
  * the drop-in classes on the real engine against golden/e2e_reference.npz (UNMODIFIED reference, 1080p / 720p /
    480p pairs): masks pixel by pixel, `_has_changed`, LEVE / PARCIAL / TOTAL classes, pct_changed, is_circular;
  * the checker bench.py uses (oracle/parity.py: live cv2 call sequence vs the C ABI) on 1080p pairs;
  * the randomised sweep of tools/fuzz_parity.py against the oracle."""
import importlib
import os
import sys

import numpy as np
import pytest

from chessboard_vision_b200 import synth
import chessboard_vision_b200.dropin as dropin

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import e2e_common

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods(engine):
    import chessboard_vision_b200.engine as engine_mod
    engine_mod._default[0] = engine
    sys.path.insert(0, dropin.PATH)
    out = {}
    for name in ("grid_extractor", "board_detection", "piece_detector", "change_detector", "frame_enhancer"):
        sys.modules.pop(name, None)
        out[name] = importlib.import_module(name)
    yield out
    sys.path.remove(dropin.PATH)
    for name in out:
        sys.modules.pop(name, None)


@pytest.mark.parametrize("idx", range(5))
def test_chain_vs_unmodified_reference(mods, idx):
    z, cases = e2e_common.load_cases()
    e2e_common.assert_case(cases[idx], e2e_common.run_case(mods, z, cases[idx]))


def test_known_answer_mask_1080p(engine):
    """kat.json (SURVEY.md 8c probe 16): Otsu T = 120 and the reference's mask of the 1080p noise frame, per pixel."""
    z, cases = e2e_common.load_cases()
    rec = cases[0]
    enh, gray, binary, T = engine.enhance(synth.noise_frame(1080, 1920, 0))
    ref = np.unpackbits(z["noise_1920x1080_0/prev_mask_bits"])[:1080 * 1920].reshape(1080, 1920)
    assert T == rec["prev"]["otsu_t"] == 120
    assert int(np.count_nonzero((binary > 0) != (ref > 0))) <= 1
    assert abs(int(np.count_nonzero(binary)) - rec["prev"]["white_px"]) <= 1


@pytest.mark.parametrize("kind", ["board", "noise"])
def test_live_checker_1080p(engine, kind):
    pytest.importorskip("cv2")
    from oracle import parity
    H, W = 1080, 1920
    r = parity.run(engine, list(synth.frame_batch(2, H, W, kind, 0)), H, W)
    assert r["otsu_t_equal"] and r["flags_equal"], r
    assert r["mask_px_diff_max"] <= (0 if kind == "board" else 4), r
    assert r["enhanced_max_abs_diff"] <= 9 and r["cd_changed_px_diff_max"] <= 2, r


def test_fuzz_against_the_oracle(engine):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_parity
    assert fuzz_parity.sweep(engine, cases=24, seed=7) == 0
