"""CPU: the end-to-end chain through the drop-in classes on the oracle-backed fake engine against what the
UNMODIFIED reference produced (golden/e2e_reference.npz), and the checker of oracle/parity.py on the same engine.
The CUDA path is bit-identical to the oracle (GPU tests), so these are the numbers the GPU reports."""
import importlib
import os
import sys

import numpy as np
import pytest

from chessboard_vision_b200 import synth
import chessboard_vision_b200.engine as engine_mod
import chessboard_vision_b200.dropin as dropin

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import e2e_common


@pytest.fixture()
def mods(monkeypatch, oracle):
    from fake_engine import FakeEngine
    monkeypatch.setitem(engine_mod._default, 0, FakeEngine())
    monkeypatch.syspath_prepend(dropin.PATH)
    out = {}
    for name in ("grid_extractor", "board_detection", "piece_detector", "change_detector", "frame_enhancer"):
        sys.modules.pop(name, None)
        out[name] = importlib.import_module(name)
    yield out
    for name in out:
        sys.modules.pop(name, None)


@pytest.mark.parametrize("idx", [3, 4])          # 640x480 noise, 1280x720 board (the 1080p cases run on the GPU)
def test_chain_vs_unmodified_reference(mods, idx):
    z, cases = e2e_common.load_cases()
    out = e2e_common.run_case(mods, z, cases[idx])
    e2e_common.assert_case(cases[idx], out)


def test_parity_checker_on_the_oracle(oracle):
    pytest.importorskip("cv2")
    from fake_engine import FakeEngine
    from oracle import parity
    H, W = 270, 480
    r = parity.run(FakeEngine(), [synth.board_frame(H, W, 2), synth.noise_frame(H, W, 4)], H, W)
    assert r["otsu_t_equal"] and r["flags_equal"] and r["mask_px_diff_max"] <= 2 and r["enhanced_max_abs_diff"] <= 9, r
    assert r["cd_classes_ref"][0]["TOTAL"] >= 1 and sum(r["squares_changed_ref"]) >= 2      # the pairs do exercise the flags
