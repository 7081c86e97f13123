"""Canny + SmartGridExtractor.refine_grid (grid_extractor.py:66-121).
CPU: oracle vs golden vectors of the unmodified reference and vs live cv2.  GPU (marked): kernels and drop-in."""
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
RG = json.load(open(os.path.join(G, "refine_grid.json")))
CASES = [(seed, tag) for seed in (7, 8, 9) for tag in ("plain", "pieces")]


def _img(seed, tag):
    b0, b1 = synth.board_with_pieces(seed, seed, 620)
    return b0 if tag == "plain" else b1


@pytest.mark.parametrize("seed,tag", CASES)
def test_oracle_vs_reference(oracle, seed, tag):
    want = RG["%d_%s" % (seed, tag)]
    gx, gy, edges = oracle.refine_grid(_img(seed, tag), return_edges=True)
    assert sha(edges) == want["edges_sha"] and int(np.count_nonzero(edges)) == want["edge_px"]
    assert gx == want["grid_x"] and gy == want["grid_y"]


def test_oracle_canny_vs_cv2(oracle):
    cv2 = pytest.importorskip("cv2")
    for img in (synth.noise_frame(200, 300, 1), synth.board_frame(480, 640, 2), synth.board_frame(97, 133, 3),
                synth.noise_frame(3, 5, 3), synth.noise_frame(1, 50, 3)):
        g = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
        for lo, hi in ((50, 150), (30.7, 100.2), (200, 100), (0, 0)):
            assert np.array_equal(oracle.canny(g, lo, hi), cv2.Canny(g, lo, hi)), (img.shape, lo, hi)
    assert sha(oracle.canny(oracle.gray(synth.board_frame(97, 133, 3)), 30, 100)) == RG["canny_97x133_30_100"]


def test_dropin_refine_grid_host_logic(oracle, monkeypatch):
    import chessboard_vision_b200.engine as engine_mod
    import chessboard_vision_b200.dropin as dropin
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_engine import FakeEngine
    monkeypatch.setitem(engine_mod._default, 0, FakeEngine())
    monkeypatch.syspath_prepend(dropin.PATH)
    sys.modules.pop("grid_extractor", None)
    ge = importlib.import_module("grid_extractor")
    try:
        sg = ge.SmartGridExtractor()
        gx, gy = sg.refine_grid(_img(8, "pieces"))
        assert list(gx) == RG["8_pieces"]["grid_x"] and list(gy) == RG["8_pieces"]["grid_y"]
        assert sg.grid_lines_x is gx and len(sg.split_board(_img(8, "pieces"))) == 64
    finally:
        sys.modules.pop("grid_extractor", None)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,tag", CASES)
def test_gpu_refine_grid(engine, seed, tag):
    want = RG["%d_%s" % (seed, tag)]
    img = _img(seed, tag)
    edges = engine.canny(engine.gray(img), 50, 150)
    assert sha(edges) == want["edges_sha"]
    rows, cols = engine.projections(edges)
    assert np.array_equal(rows, edges.sum(1, dtype=np.uint64)) and np.array_equal(cols, edges.sum(0, dtype=np.uint64))
    gx, gy = engine.refine_grid(img)
    assert [int(v) for v in gx] == want["grid_x"] and [int(v) for v in gy] == want["grid_y"]


@pytest.mark.gpu
def test_gpu_canny_shapes_thresholds_batch(engine, oracle):
    for img in (synth.noise_frame(200, 300, 1), synth.board_frame(97, 133, 3), synth.noise_frame(3, 5, 3),
                synth.noise_frame(1, 50, 3), synth.board_frame(1080, 1920, 0)):
        g = oracle.gray(img)
        for lo, hi in ((50, 150), (30.7, 100.2), (200, 100)):
            assert np.array_equal(engine.canny(g, lo, hi), oracle.canny(g, lo, hi)), (img.shape, lo, hi)
    batch = np.stack([oracle.gray(synth.board_frame(120, 160, s)) for s in range(3)])
    got = engine.canny(batch, 50, 150)
    for i in range(3):
        assert np.array_equal(got[i], oracle.canny(batch[i], 50, 150))
    # a long snake of weak pixels hanging off one strong pixel: the flood has to cross many tiles
    g = np.zeros((64, 640), np.uint8); g[30:34, :] = 40; g[30:34, :4] = 255
    assert np.array_equal(engine.canny(g, 50, 150), oracle.canny(g, 50, 150))
