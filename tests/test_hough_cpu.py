"""CPU: the Hough-circle oracle against cv2.HoughCircles' committed outputs, the library's host-side
geometry (OpenCV's argument handling), and the drop-in PieceDetector's circle path against the
reference's own results (tests/golden/hough.npz, hough_reference.json -- tools/make_golden.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle as O
from chessboard_vision_b200 import _lib, synth
from chessboard_vision_b200.engine import HOUGH_SQUARE_DTYPE

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _geometry(rects, **kw):
    from fake_engine import FakeEngine
    e = FakeEngine()
    p = e.hough_params(**kw)
    return p, e.hough_geometry(rects, p)


def test_oracle_matches_cv2_golden_circles():
    z = np.load(os.path.join(GOLD, "hough.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    plane, rects = z["plane"], [tuple(r) for r in meta["rects"]]
    p2, r2 = synth.shape_atlas(3, 20, 8, 100, 512)
    assert np.array_equal(plane, p2) and rects == r2      # the generator is deterministic
    n_circles = 0
    for ci, case in enumerate(meta["cases"]):
        p, geo = None, None
        try:
            p, geo = _geometry(rects, **case["params"])
        except ValueError:
            pass
        for i, (x, y, w, h) in enumerate(rects):
            want = z["c%d_s%d" % (ci, i)]
            if geo is None:      # a square with minDist 0 in the list: resolve one by one
                try:
                    _, g1 = _geometry([rects[i]], **case["params"])
                except ValueError:
                    assert len(want) == 0
                    continue
                g = g1[0]
            else:
                g = geo[i]
            got = O.hough_circles(plane[y:y + h, x:x + w], dp=max(1.0, case["params"]["dp"]), min_dist=float(g["min_dist"]),
                                  param1=case["params"]["param1"], param2=case["params"]["param2"],
                                  min_radius=int(g["min_radius"]), max_radius=int(g["max_radius"]), max_out=4096)
            if len(want) == 0:
                assert got is None, (ci, i)
            else:
                assert got is not None and got.shape == want.shape, (ci, i)
                assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (ci, i)
                n_circles += len(want)
    assert n_circles > 100


def test_geometry_follows_reference_expressions_and_opencv_fixups():
    rects = [(0, 0, 77, 77), (0, 0, 78, 77), (5, 5, 30, 64), (0, 0, 128, 128)]
    p, geo = _geometry(rects)
    for (x, y, w, h), g in zip(rects, geo):
        md = min(w, h)
        assert g["min_radius"] == int(md * 0.20) and g["max_radius"] == int(md * 0.55) and g["min_dist"] == md // 3
        assert g["acc_rows"] == int(np.ceil(np.float32(h) * (np.float32(1) / np.float32(1.2))))
        assert g["n_bins"] == int(np.rint(np.float32(g["max_radius"] - g["min_radius"]) / np.float32(1.2) * np.float32(10)))
    # maxRadius <= 0 -> max(h, w); maxRadius <= minRadius -> minRadius + 2; dp < 1 -> 1; minRadius < 0 -> 0
    _, g = _geometry([(0, 0, 40, 60)], dp=0.5, min_radius=-3, max_radius=0, min_dist=2.5)
    assert (g[0]["min_radius"], g[0]["max_radius"], g[0]["acc_rows"], g[0]["acc_cols"]) == (0, 60, 60, 40)
    _, g = _geometry([(0, 0, 40, 60)], min_radius=9, max_radius=4, min_dist=2.5)
    assert (g[0]["min_radius"], g[0]["max_radius"]) == (9, 11) and g[0]["min_dist"] == np.float32(2.5)
    with pytest.raises(ValueError):
        _geometry([(0, 0, 2, 2)])                 # minDist = 2 // 3 = 0
    with pytest.raises(ValueError):
        _geometry([(0, 0, 20, 20)], param1=0)


def test_oracle_edge_cases():
    flat = np.full((40, 40), 90, np.uint8)
    assert O.hough_circles(flat, min_dist=13, min_radius=8, max_radius=22) is None
    one = np.zeros((1, 1), np.uint8)
    assert O.hough_circles(one, min_dist=1, min_radius=0, max_radius=0) is None
    with pytest.raises(ValueError):
        O.hough_circles(flat, min_dist=0)
    # a clean disc: one circle at its centre, radius within a pixel
    yy, xx = np.ogrid[:77, :77]
    disc = np.where((xx - 38) ** 2 + (yy - 40) ** 2 <= 24 ** 2, 200, 60).astype(np.uint8)
    c = O.hough_circles(O.gaussian(disc, 5), min_dist=25, min_radius=15, max_radius=42)
    assert c is not None and len(c) == 1
    assert abs(c[0, 0] - 38.5) <= 1.2 and abs(c[0, 1] - 40.5) <= 1.2 and abs(c[0, 2] - 24) <= 1.5


def test_dropin_circle_path_matches_reference(monkeypatch):
    """PieceDetector._detect_circle_unified / detect_piece through the drop-in (oracle-backed engine)
    against the unmodified reference on the 64 squares of three boards."""
    import fake_engine
    from chessboard_vision_b200.dropin import piece_detector as pdm, grid_extractor as ge
    monkeypatch.setattr(pdm, "default_engine", lambda device=0: fake_engine.FakeEngine())
    ref = json.load(open(os.path.join(GOLD, "hough_reference.json")))
    found = 0
    for seed, rows in ref.items():
        _, board = synth.board_with_pieces(int(seed), 7, 620)
        det = pdm.PieceDetector()
        squares = ge.GridExtractor().split_board(board)
        for pos, sq in squares.items():
            want = rows["%d_%d" % pos]
            f, center, radius, kind = det._detect_circle_unified(det._preprocess_square(sq))
            assert bool(f) == want["found"], (seed, pos)
            if f:
                assert list(center) == want["center"] and radius == want["radius"] and kind == want["kind"]
                found += 1
        # the batched path: calibrate_reference runs one Hough launch for all squares
        det.calibrate_reference(squares)
        for pos in squares:
            want = rows["%d_%d" % pos]
            got = det.cached_results[pos]
            assert got["has_piece"] == want["has_piece"] and got["method"] == want["method"], (seed, pos)
            assert got["confidence"] == pytest.approx(want["confidence"], abs=0, rel=0)
            if want["method"] in ("hough", "tower_top"):
                assert list(got["center"]) == want["center"] and got["radius"] == want["radius"]
        res, _ = pdm.PieceDetector().detect_all_pieces(squares, use_smoothing=False)
        for pos in squares:
            assert res[pos]["method"] == rows["%d_%d" % pos]["method"]
    assert found >= 20


def test_ref_cv2_sequence_and_oracle_agree_with_reference_results():
    """oracle/ref_cv2.detect_circle_unified (the CPU arm bench.py times) reproduces the reference's answers."""
    cv2 = pytest.importorskip("cv2")
    from oracle import ref_cv2
    ref = json.load(open(os.path.join(GOLD, "hough_reference.json")))
    rows = ref["11"]
    _, board = synth.board_with_pieces(11, 7, 620)
    for r in range(8):
        for c in range(8):
            sq = board[r * 77:(r + 1) * 77, c * 77:(c + 1) * 77]
            g = O.square_preprocess(sq, 5)
            (found, center, radius, kind), circles = ref_cv2.detect_circle_unified(g)
            want = rows["%d_%d" % (c, 7 - r)]
            assert found == want["found"] and kind == want["kind"]
            mine = O.hough_circles(g, min_dist=25, min_radius=15, max_radius=42)
            if circles is None:
                assert mine is None
            else:
                assert np.array_equal(circles[0], mine)
