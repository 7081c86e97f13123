"""CPU: host logic of the drop-in modules (chessboard_vision_b200/dropin) with the kernels replaced
by an oracle-backed fake engine (tests/fake_engine.py).  Checks the dict contract of split_board,
square packing, the device-state windows (means / variances / reference_squares) and the gating
logic against golden vectors recorded from the unmodified reference classes."""
import importlib
import json
import os
import sys

import numpy as np
import pytest

from chessboard_vision_b200 import synth, hostapi
import chessboard_vision_b200.engine as engine_mod
import chessboard_vision_b200.dropin as dropin

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def mods(monkeypatch, oracle):
    from fake_engine import FakeEngine
    fake = FakeEngine()
    monkeypatch.setitem(engine_mod._default, 0, fake)
    monkeypatch.syspath_prepend(dropin.PATH)
    out = {}
    for name in ("grid_extractor", "board_detection", "piece_detector", "change_detector", "frame_enhancer"):
        sys.modules.pop(name, None)
        out[name] = importlib.import_module(name)
        assert out[name].__file__.startswith(dropin.PATH)
    yield out
    for name in out:
        sys.modules.pop(name, None)


def test_split_board_contract(mods):
    grid = json.load(open(os.path.join(G, "grid.json")))
    board = synth.board_frame(620, 620, 7)
    ge = mods["grid_extractor"]
    for kind, ext in (("linear", ge.GridExtractor()), ("smart", ge.SmartGridExtractor())):
        if kind == "smart":
            assert list(ext.split_board(board).keys()) == [tuple(k) for k, _ in grid["linear"]]   # unset lines -> linear
            ext.grid_lines_x, ext.grid_lines_y = list(synth.CALIB_GRID_X), list(synth.CALIB_GRID_Y)
        sq = ext.split_board(board)
        assert [list(k) for k in sq.keys()] == [k for k, _ in grid[kind]]        # insertion order, rank 8 first
        for (k, rect) in grid[kind]:
            v = sq[tuple(k)]
            x, y, w, h = rect
            assert v.base is board and v.shape == (h, w, 3)                      # views, not copies
            assert np.shares_memory(v, board[y:y + h, x:x + w]) and np.array_equal(v, board[y:y + h, x:x + w])
        b2, rects, keys = hostapi.pack_squares(sq)
        assert b2 is board and rects == [tuple(r) for _, r in grid[kind]]       # zero-copy fast path
    degenerate = ge.SmartGridExtractor()
    degenerate.grid_lines_x = [0, 10, 10, 30, 40, 50, 60, 70, 80]; degenerate.grid_lines_y = list(range(0, 90, 10))
    assert len(degenerate.split_board(board)) == 56                              # zero-width column skipped


def test_pack_squares_atlas():
    sq = {(c, r): np.full((50 + r, 40 + c), 10 * r + c, np.uint8) for r in range(3) for c in range(8)}
    atlas, rects, keys = hostapi.pack_squares(sq)
    assert keys == list(sq.keys())
    for k, (x, y, w, h) in zip(keys, rects):
        assert np.array_equal(atlas[y:y + h, x:x + w], sq[k])
    with pytest.raises(ValueError):
        hostapi.pack_squares({(0, 0): np.zeros((4, 4), np.float32)})
    assert hostapi.pack_squares({}) == (None, [], [])


def test_reorder_and_warp(mods, oracle):
    bd = mods["board_detection"]
    pts = np.array([[[1560, 108]], [[550, 1005]], [[556, 112]], [[1562, 1024]]])
    assert bd.reorder(pts).reshape(4, 2).tolist() == [[556, 112], [1560, 108], [550, 1005], [1562, 1024]]
    z = np.load(os.path.join(G, "warp_small.npz"))
    w, M, S = bd.warp_image(synth.noise_frame(270, 480, 9), z["points"], display_size=(200, 180), margin=20)
    assert S == int(z["board_size"]) and np.array_equal(M, z["matrix"]) and np.array_equal(w, z["warped"])
    with pytest.raises(ValueError):
        bd.warp_image(np.zeros((10, 10), np.uint8), z["points"])


def test_enhancer_surface(mods):
    fe = mods["frame_enhancer"]
    z = np.load(os.path.join(G, "enhancer_small.npz"))
    e = fe.ImageEnhancer()
    assert e.profile == {} and e.sharpen_kernel.tolist() == [[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]]
    assert e.clahe.getClipLimit() == 3.0 and e.clahe.getTilesGridSize() == (8, 8)
    img = z["board_96x128/input"]
    assert np.array_equal(e.correct_lighting(img), z["board_96x128/correct_lighting"])
    assert np.array_equal(e.sharpen(z["board_96x128/reduce_noise"]), z["board_96x128/sharpen"])
    assert np.array_equal(e.normalize_intensity(z["board_96x128/sharpen"]), z["board_96x128/normalize"])
    g, b = e.prepare_analysis(z["board_96x128/normalize"])
    assert np.array_equal(g, z["board_96x128/gray"]) and np.array_equal(b, z["board_96x128/binary"])
    assert np.array_equal(e.clahe.apply(np.ascontiguousarray(z["board_96x128/lab"][..., 0])), z["board_96x128/clahe_l"])
    assert np.array_equal(e.apply_color_profile(img), img)
    assert np.abs(e.process_pipeline(img).astype(int) - z["board_96x128/normalize"]).max() <= 9   # 1 LSB bilateral x sharpen gain
    with pytest.raises(ValueError):
        e.correct_lighting(np.zeros((4, 4), np.uint8))


def _golden_squares(mods):
    board, pieces = synth.board_with_pieces(11, 7, 620)
    ge = mods["grid_extractor"].GridExtractor()
    return ge.split_board(board), ge.split_board(pieces)


def test_change_detector_matches_reference(mods):
    cdj = json.load(open(os.path.join(G, "change_detector.json")))
    sq_ref, sq_cur = _golden_squares(mods)
    cd = mods["change_detector"].ChangeDetector()
    assert not cd.is_calibrated and cd.detect_changes_detailed(sq_cur) == {} and cd.get_focus_count() == 64
    cd.calibrate(sq_ref)
    assert cd.is_calibrated and len(cd.means) == 64 and (3, 3) in cd.variances
    assert cd.means[(0, 7)].dtype == np.float32 and cd.means[(0, 7)].shape == (77, 77)
    assert float(cd.variances[(5, 2)].min()) == 100.0
    detailed = cd.detect_changes_detailed(sq_cur)
    assert {"%d_%d" % k for k in detailed} == set(cdj["detailed"])
    for k, v in detailed.items():
        want = cdj["detailed"]["%d_%d" % k]
        assert v["pct_changed"] == want["pct_changed"] and v["z_score"] == want["z_score"]
        assert v["intensity"] == want["intensity"] and v["center_ratio"] == 1.0
        assert v["is_circular"] == want["is_circular"]
    assert {("%d_%d" % k): v for k, v in cd.detect_changes(sq_cur).items()} == cdj["changes"]
    import hashlib
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
    cd.update_all_references(sq_cur)
    for k in sq_cur:
        assert sha(cd.means[k]) == cdj["means_after_update_sha"]["%d_%d" % k]
        assert sha(cd.variances[k]) == cdj["vars_after_update_sha"]["%d_%d" % k]
    # tuned parameters + focus squares (sensitivity_settings.json values)
    cd2 = mods["change_detector"].ChangeDetector()
    cd2.blur_kernel, cd2.z_threshold, cd2.alpha, cd2.initial_variance = 13, 2.55, 0.13, 600
    cd2.calibrate(sq_ref)
    cd2.set_focus_squares([(0, 7), (3, 3), (4, 4), (7, 0)])
    assert cd2.get_focus_count() == 4
    cd2.update_all_references(sq_cur)
    det2 = cd2.detect_changes_detailed(sq_cur)
    t = cdj["tuned"]
    for k in sq_cur:
        assert sha(cd2.means[k]) == t["means_sha"]["%d_%d" % k] and sha(cd2.variances[k]) == t["vars_sha"]["%d_%d" % k]
    assert {"%d_%d" % k for k in det2} == set(t["detailed"])
    for k, v in det2.items():
        w = t["detailed"]["%d_%d" % k]
        assert (v["pct_changed"], v["z_score"], v["intensity"]) == (w["pct_changed"], w["z_score"], w["intensity"])
    cd2.clear_focus()
    assert cd2.focus_squares == set()


def test_reference_regression_test_ported(mods):
    """The reference's own test (test_change_detector_regression.py:19-54), line for line in behaviour."""
    ChangeDetector = mods["change_detector"].ChangeDetector
    det = ChangeDetector()
    rng = np.random.default_rng(0)
    det.calibrate({(c, r): rng.integers(0, 255, (50, 50), dtype=np.uint8) for r in range(8) for c in range(8)})
    assert det.is_calibrated
    det = ChangeDetector()
    squares = {(c, r): np.zeros((50, 50), np.uint8) for r in range(8) for c in range(8)}
    det.calibrate(squares)
    squares[(3, 3)] = np.full((50, 50), 255, np.uint8)
    changes = det.detect_changes(squares)
    assert (3, 3) in changes and changes[(3, 3)] > 50.0
    detailed = det.detect_changes_detailed(squares)
    assert detailed[(3, 3)]["intensity"] == "TOTAL" and len(detailed) == 1
    want = json.load(open(os.path.join(G, "change_detector.json")))["regression_3_3"]
    assert detailed[(3, 3)]["pct_changed"] == want["pct_changed"] and detailed[(3, 3)]["z_score"] == want["z_score"]
    pat = det.classify_hand_pattern(detailed)
    assert pat == {"is_hand": False, "is_move": False, "move_candidates": {(3, 3)}}
    many = {(i, 0): {"intensity": "TOTAL"} for i in range(2)}
    assert det.classify_hand_pattern(many)["is_hand"]


def test_state_windows_are_dict_like(mods):
    sq_ref, sq_cur = _golden_squares(mods)
    cd = mods["change_detector"].ChangeDetector()
    cd.calibrate(sq_ref)
    m = cd.means[(2, 2)]
    cd.means[(2, 2)] = m + 50                       # host assignment, as a caller of the reference could do
    assert np.array_equal(cd.means[(2, 2)], m + 50)
    det = cd.detect_changes_detailed(sq_ref)         # staged value reaches the device before the launch
    assert (2, 2) in det and det[(2, 2)]["pct_changed"] == 100.0
    del cd.means[(2, 2)]
    assert (2, 2) not in cd.means and len(cd.means) == 63
    assert (2, 2) not in cd.detect_changes_detailed(sq_ref)
    assert list(cd.means.keys())[0] == (0, 7)
    assert cd.means.get((9, 9)) is None


def test_piece_detector_matches_reference(mods):
    z = np.load(os.path.join(G, "piece_detector.npz"))
    sq_ref, sq_cur = _golden_squares(mods)
    pd = mods["piece_detector"].PieceDetector()
    assert (pd.min_radius_ratio, pd.max_radius_ratio, pd.change_threshold, pd.history_size) == (0.20, 0.55, 25, 5)
    pd.update_references(sq_ref)
    assert len(pd.reference_squares) == 64 and pd.cached_results == {}
    for row in z["stats"]:
        pos = (int(row[0]), int(row[1]))
        g = pd._preprocess_square(sq_cur[pos])
        assert pd._has_changed(pos, g) == bool(row[3])
        diff, cm, bm = pd._detect_center_vs_border(g)
        assert (diff, cm, bm) == (row[5], row[6], row[7])
        assert pd._analyze_radial_symmetry(g) == row[8]
    assert pd._has_changed((9, 9), np.zeros((77, 77), np.uint8)) is True          # no reference -> process


def test_detect_all_pieces_flow(mods):
    from fake_engine import FakeEngine
    pytest.importorskip("cv2")
    sq_ref, sq_cur = _golden_squares(mods)
    z = np.load(os.path.join(G, "piece_detector.npz"))
    changed = {(int(r[0]), int(r[1])) for r in z["stats"] if r[3]}
    pd = mods["piece_detector"].PieceDetector()
    pd.calibrate_reference(sq_ref)
    assert len(pd.cached_results) == 64
    n0 = FakeEngine.launches
    res, vis = pd.detect_all_pieces(sq_cur)
    assert FakeEngine.launches - n0 <= 3                   # statistics + Hough (+ one reference update)
    assert vis == changed and list(res.keys()) == list(sq_cur.keys())
    for pos, r in res.items():
        assert set(r) >= {"has_piece", "confidence", "center", "radius", "method", "center_border_diff"}
        assert len(pd.detection_history[pos]) == 1
    # unchanged squares come from the cache; forced squares are re-detected
    res2, vis2 = pd.detect_all_pieces(sq_cur, squares_to_check={(0, 0)})
    assert vis2 <= changed
    for _ in range(4):
        pd.detect_all_pieces(sq_cur)
    assert all(len(h) == 5 for h in pd.detection_history.values())
    occ = pd.get_occupied_squares(sq_cur)
    assert isinstance(occ, set)
    single = pd.detect_piece(sq_cur[(0, 0)], (0, 0))
    assert single.keys() == res[(0, 0)].keys()
