"""GPU: the drop-in modules (reference class / function names) with the real engine, against golden
vectors from the unmodified reference classes, plus the reference's own regression test."""
import hashlib
import importlib
import json
import os
import sys

import numpy as np
import pytest

from chessboard_vision_b200 import synth
import chessboard_vision_b200.dropin as dropin

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


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


def _squares(mods):
    board, pieces = synth.board_with_pieces(11, 7, 620)
    ge = mods["grid_extractor"].GridExtractor()
    return ge.split_board(board), ge.split_board(pieces)


def test_enhancer_class(mods):
    z = np.load(os.path.join(G, "enhancer_small.npz"))
    e = mods["frame_enhancer"].ImageEnhancer()
    for name in ("board_96x128", "noise_90x121"):
        g = lambda k: z[name + "/" + k]
        assert np.array_equal(e.correct_lighting(g("input")), g("correct_lighting"))
        assert np.abs(e.reduce_noise(g("correct_lighting")).astype(int) - g("reduce_noise")).max() <= 1
        assert np.array_equal(e.sharpen(g("reduce_noise")), g("sharpen"))
        assert np.array_equal(e.normalize_intensity(g("sharpen")), g("normalize"))
        gray, binary = e.prepare_analysis(g("normalize"))
        assert np.array_equal(gray, g("gray")) and np.array_equal(binary, g("binary"))
        full = e.process_pipeline(g("input"))
        assert full.shape == g("input").shape and np.abs(full.astype(int) - g("normalize")).max() <= 9
        enh, gray2, binary2 = e.process_and_analyze(g("input"))
        assert np.array_equal(enh, full)
    assert np.array_equal(e.clahe.apply(np.ascontiguousarray(z["board_96x128/lab"][..., 0])), z["board_96x128/clahe_l"])


def test_warp_image_function(mods):
    k = json.load(open(os.path.join(G, "kat.json")))["kat"]["1920x1080"]
    img = synth.noise_frame(1080, 1920, 0)
    pts = mods["board_detection"].reorder(np.array(synth.CALIB_CORNERS_1080P).reshape(4, 1, 2))
    warped, M, S = mods["board_detection"].warp_image(img, pts)
    assert S == 620 and M.ravel().tolist() == k["warp_matrix"] and sha(warped) == k["warp"]


def test_change_detector_class(mods):
    cdj = json.load(open(os.path.join(G, "change_detector.json")))
    sq_ref, sq_cur = _squares(mods)
    cd = mods["change_detector"].ChangeDetector()
    cd.calibrate(sq_ref)
    for key, want in cdj["means_sha"].items():
        c, r = map(int, key.split("_"))
        assert sha(cd.means[(c, r)]) == want
    detailed = cd.detect_changes_detailed(sq_cur)
    assert {"%d_%d" % k for k in detailed} == set(cdj["detailed"])
    for k, v in detailed.items():
        want = cdj["detailed"]["%d_%d" % k]
        assert v["pct_changed"] == want["pct_changed"] and v["z_score"] == want["z_score"]
        assert v["intensity"] == want["intensity"] and v["is_circular"] == want["is_circular"]
    assert {("%d_%d" % k): v for k, v in cd.detect_changes(sq_cur).items()} == cdj["changes"]
    cd.update_all_references(sq_cur)
    for k in sq_cur:
        assert sha(cd.means[k]) == cdj["means_after_update_sha"]["%d_%d" % k]
        assert sha(cd.variances[k]) == cdj["vars_after_update_sha"]["%d_%d" % k]
    cd2 = mods["change_detector"].ChangeDetector()
    cd2.blur_kernel, cd2.z_threshold, cd2.alpha, cd2.initial_variance = 13, 2.55, 0.13, 600
    cd2.calibrate(sq_ref)
    cd2.set_focus_squares([(0, 7), (3, 3), (4, 4), (7, 0)])
    cd2.update_all_references(sq_cur)
    det2 = cd2.detect_changes_detailed(sq_cur)
    t = cdj["tuned"]
    for k in sq_cur:
        assert sha(cd2.means[k]) == t["means_sha"]["%d_%d" % k] and sha(cd2.variances[k]) == t["vars_sha"]["%d_%d" % k]
    assert {"%d_%d" % k for k in det2} == set(t["detailed"])
    for k, v in det2.items():
        w = t["detailed"]["%d_%d" % k]
        assert (v["pct_changed"], v["z_score"], v["intensity"]) == (w["pct_changed"], w["z_score"], w["intensity"])


def test_reference_regression_test(mods):
    """test_change_detector_regression.py:19-54 against the B200 ChangeDetector."""
    ChangeDetector = mods["change_detector"].ChangeDetector
    det = ChangeDetector()
    det.calibrate({(c, r): np.random.randint(0, 255, (50, 50), dtype=np.uint8) for r in range(8) for c in range(8)})
    assert det.is_calibrated
    det = ChangeDetector()
    squares = {(c, r): np.zeros((50, 50), np.uint8) for r in range(8) for c in range(8)}
    det.calibrate(squares)
    squares[(3, 3)] = np.full((50, 50), 255, np.uint8)
    changes = det.detect_changes(squares)
    assert (3, 3) in changes and changes[(3, 3)] > 50.0
    detailed = det.detect_changes_detailed(squares)
    assert (3, 3) in detailed and detailed[(3, 3)]["intensity"] == "TOTAL"
    want = json.load(open(os.path.join(G, "change_detector.json")))["regression_3_3"]
    assert detailed[(3, 3)]["pct_changed"] == want["pct_changed"] and detailed[(3, 3)]["z_score"] == want["z_score"]


def test_piece_detector_class(mods):
    z = np.load(os.path.join(G, "piece_detector.npz"))
    gray_sha = json.loads(str(z["gray_sha"]))
    sq_ref, sq_cur = _squares(mods)
    pd = mods["piece_detector"].PieceDetector()
    pd.update_references(sq_ref)
    for row in z["stats"]:
        pos = (int(row[0]), int(row[1]))
        g = pd._preprocess_square(sq_cur[pos])
        assert sha(g) == gray_sha["%d_%d" % pos]
        assert pd._has_changed(pos, g) == bool(row[3])
        assert pd._detect_center_vs_border(g) == (row[5], row[6], row[7])
        assert pd._analyze_radial_symmetry(g) == row[8]
    changed = {(int(r[0]), int(r[1])) for r in z["stats"] if r[3]}
    pd.calibrate_reference(sq_ref)
    res, vis = pd.detect_all_pieces(sq_cur)
    assert vis == changed and list(res.keys()) == list(sq_cur.keys())
    # the references of squares that were processed and stable now hold the current frame
    res2, vis2 = pd.detect_all_pieces(sq_cur)
    assert vis2 <= vis


def test_piece_detector_circles_match_reference(mods):
    """_detect_circle_unified / calibrate_reference / detect_all_pieces with the Hough kernel against
    the unmodified reference (cv2.HoughCircles) on the 64 squares of three boards."""
    ref = json.load(open(os.path.join(G, "hough_reference.json")))
    found = 0
    for seed, rows in ref.items():
        _, board = synth.board_with_pieces(int(seed), 7, 620)
        det = mods["piece_detector"].PieceDetector()
        squares = mods["grid_extractor"].GridExtractor().split_board(board)
        for pos, sq in squares.items():
            want = rows["%d_%d" % pos]
            f, center, radius, kind = det._detect_circle_unified(det._preprocess_square(sq))
            assert bool(f) == want["found"], (seed, pos)
            if f:
                assert list(center) == want["center"] and radius == want["radius"] and kind == want["kind"]
                found += 1
            full = det.detect_piece(sq, pos)
            assert full["has_piece"] == want["has_piece"] and full["method"] == want["method"]
            assert full["confidence"] == want["confidence"]
        det.calibrate_reference(squares)
        res, _ = mods["piece_detector"].PieceDetector().detect_all_pieces(squares, use_smoothing=False)
        for pos in squares:
            want = rows["%d_%d" % pos]
            for got in (det.cached_results[pos], res[pos]):
                assert got["has_piece"] == want["has_piece"] and got["method"] == want["method"], (seed, pos)
                assert got["confidence"] == want["confidence"]
                if want["method"] in ("hough", "tower_top"):
                    assert list(got["center"]) == want["center"] and got["radius"] == want["radius"]
    assert found >= 20


def test_ragged_and_empty_square_dicts(mods, oracle):
    """Squares of different sizes that are not views of one board (atlas path), empty dicts, a missing key."""
    rng = np.random.default_rng(5)
    shapes = {(c, r): (40 + 3 * r + c, 37 + 2 * c + r) for r in range(3) for c in range(4)}
    ref = {k: rng.integers(0, 256, s + (3,), dtype=np.uint8) for k, s in shapes.items()}
    cur = {k: (v if (k[0] + k[1]) % 3 else 255 - v) for k, v in ref.items()}
    cd = mods["change_detector"].ChangeDetector()
    assert cd.detect_changes_detailed({}) == {} and cd.detect_changes({}) == {}
    cd.calibrate(ref)
    det = cd.detect_changes_detailed(cur)
    for k in ref:
        g0 = oracle.square_preprocess(np.ascontiguousarray(ref[k]), 5)
        g1 = oracle.square_preprocess(np.ascontiguousarray(cur[k]), 5)
        m, v = oracle.cd_calibrate(g0, 100.0)
        assert np.array_equal(cd.means[k], m)
        cnt, zmax = oracle.cd_detect(g1, m, v, 2.5)
        pct = cnt / g1.size * 100
        if pct < 5.0:
            assert k not in det
        else:
            assert det[k]["pct_changed"] == pct and det[k]["z_score"] == float(np.float32(zmax))
    sub = {k: cur[k] for k in list(cur)[:5]}                 # a subset of the calibrated squares
    assert set(cd.detect_changes_detailed(sub)) <= set(sub)
    cd.update_all_references({})                              # nothing to update: no error
    pd = mods["piece_detector"].PieceDetector()
    assert pd.detect_all_pieces({}) == ({}, set())
    pd.update_references(ref)
    res, vis = pd.detect_all_pieces(cur)
    want = {k for k in ref if float(np.mean(np.abs(oracle.square_preprocess(np.ascontiguousarray(cur[k]), 5).astype(int) -
                                                      oracle.square_preprocess(np.ascontiguousarray(ref[k]), 5)))) > 25}
    assert vis == want


def test_error_paths(engine):
    import pytest as _pt
    with _pt.raises(ValueError):
        engine.bgr2lab(np.zeros((4, 4, 3), np.float32))                      # wrong dtype
    with _pt.raises(ValueError):
        engine.bgr2lab(np.zeros((4, 4), np.uint8))                           # wrong rank
    with _pt.raises(ValueError):
        engine.clahe(np.zeros((64, 64), np.uint8), 3.0, (32, 8))             # tile grid beyond 16
    big = np.zeros((400, 400), np.uint8)
    with _pt.raises(ValueError):
        engine.squares(big, [(0, 0, 400, 400)], engine.square_params())      # square too large for shared memory
    st = engine.new_state(1, 50, 50)
    with _pt.raises(Exception):
        engine.squares(np.zeros((60, 60), np.uint8), [(0, 0, 10, 10)], engine.square_params(), st)   # state / board mismatch
    with _pt.raises(Exception):
        engine.squares(np.zeros((50, 50), np.uint8), [(0, 0, 10, 10)], engine.square_params(), st, stream0=3)
    st.free()
