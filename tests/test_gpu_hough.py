"""GPU parity: cv2.HoughCircles per square (PieceDetector._detect_circle_unified,
piece_detector.py:210-270) through the C ABI vs the CPU oracle -- bit-exact circle
lists (order, f32 centre and radius), supports, edge and centre counts."""
import json
import os

import numpy as np
import pytest

from chessboard_vision_b200 import synth, _lib
from chessboard_vision_b200.engine import grid_rects, SQ_PD_STATS, HOUGH_DTYPE

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _oracle_square(oracle, plane, rect, geo):
    x, y, w, h = rect
    g = np.ascontiguousarray(plane[y:y + h, x:x + w])
    return oracle.hough_circles(g, dp=geo["dp"], min_dist=float(geo["min_dist"]), param1=geo["param1"], param2=geo["param2"],
                                min_radius=int(geo["min_radius"]), max_radius=int(geo["max_radius"]), max_out=4096,
                                return_info=True)


def _check(engine, oracle, plane, rects, params, dp, select=None, frame_planes=None):
    planes = plane[None] if frame_planes is None else frame_planes
    got = engine.hough(planes, rects, params, select=select)
    geo = engine.hough_geometry(rects, params)
    n_circles = 0
    for f in range(planes.shape[0]):
        for i, r in enumerate(rects):
            rec = got[f, i]
            if select is not None and not np.asarray(select).reshape(planes.shape[0], -1)[f, i]:
                assert rec["status"] == _lib.HOUGH_SKIPPED
                assert rec["count"] == 0
                continue
            g = dict(dp=dp, param1=params.param1, param2=params.param2, min_dist=geo[i]["min_dist"],
                     min_radius=geo[i]["min_radius"], max_radius=geo[i]["max_radius"])
            circles, support, n_edges, n_centers = _oracle_square(oracle, planes[f], r, g)
            want = 0 if circles is None else len(circles)
            assert rec["status"] == 0
            assert rec["n_edges"] == n_edges and rec["n_centers"] == n_centers, (f, i, r)
            assert rec["count"] == want, (f, i, r, rec["count"], want)
            k = min(want, 16)
            if k:
                assert np.array_equal(rec["xyr"][:k].view(np.uint32), circles[:k].view(np.uint32)), (f, i, r)
                assert np.array_equal(rec["support"][:k], support[:k])
            assert not rec["xyr"][k:].any() and not rec["support"][k:].any()
            n_circles += want
    return n_circles


def test_hough_board_squares_reference_parameters(engine, oracle):
    """64 squares of a warped board with pieces, gray + blur 5 as PieceDetector preprocesses them,
    with the reference's arguments (dp 1.2, minDist side//3, 100 / 25, radii 20-55 %)."""
    rects, _ = grid_rects(620)
    total = 0
    planes = []
    for seed in (11, 12, 13):
        _, pieces = synth.board_with_pieces(seed, 7, 620)
        plane = np.zeros((620, 620), np.uint8)
        for (x, y, w, h) in rects:
            plane[y:y + h, x:x + w] = oracle.square_preprocess(pieces[y:y + h, x:x + w], 5)
        planes.append(plane)
    planes = np.stack(planes)
    total = _check(engine, oracle, None, rects, engine.hough_params(), 1.2, frame_planes=planes)
    assert total >= 20


def test_hough_state_plane_after_square_statistics(engine, oracle):
    """hough_state reads the gray+blur squares k_squares left in the state (plane pd_cur)."""
    rects, _ = grid_rects(620, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
    boards = np.stack([synth.board_with_pieces(s, 7, 620)[1] for s in (21, 22)])
    st = engine.new_state(3, 620, 620)
    engine.squares(boards, rects, engine.square_params(ops=SQ_PD_STATS), state=st, stream0=1)
    sel = np.ones((2, 64), np.uint8); sel[0, ::3] = 0; sel[1, 5:9] = 0
    got = engine.hough_state(st, rects, stream0=1, n=2, select=sel)
    assert got.dtype == HOUGH_DTYPE and got.shape == (2, 64)
    geo = engine.hough_geometry(rects, engine.hough_params())
    found = 0
    for f in range(2):
        cur = st.get(1 + f, _lib.PLANE_PD_CUR)
        for i, r in enumerate(rects):
            if not sel[f, i]:
                assert got[f, i]["status"] == 1 and got[f, i]["count"] == 0
                continue
            g = dict(dp=1.2, param1=100, param2=25, min_dist=geo[i]["min_dist"], min_radius=geo[i]["min_radius"],
                     max_radius=geo[i]["max_radius"])
            circles, support, n_edges, n_centers = _oracle_square(oracle, cur, r, g)
            k = 0 if circles is None else len(circles)
            assert got[f, i]["count"] == k and got[f, i]["n_edges"] == n_edges
            if k:
                assert np.array_equal(got[f, i]["xyr"][:k], circles)
                found += 1
    assert found >= 8
    st.free()


@pytest.mark.parametrize("seed", range(6))
def test_hough_shape_atlas_wild_parameters(engine, oracle, seed):
    """Rectangles of 8..128 pixels a side with discs, rings, bars and noise; dp, thresholds, radii
    and minDist far from the reference's values (including OpenCV's argument fix-ups)."""
    rng = np.random.default_rng(100 + seed)
    plane, rects = synth.shape_atlas(seed, 40)
    combos = [
        dict(dp=1.2, param1=100, param2=25),
        dict(dp=1.0, param1=50, param2=10, min_radius=0, max_radius=0, min_dist=5.5),
        dict(dp=2.0, param1=30, param2=5, min_radius=3, max_radius=2, min_dist=1.0),
        dict(dp=1.7, param1=3, param2=1, min_radius_ratio=0.12, max_radius_ratio=1.0, min_dist=12.0),
        dict(dp=0.5, param1=200, param2=30, min_radius=200, max_radius=10),
        dict(dp=1.5, param1=1, param2=25, min_radius=3, max_radius=0, min_dist=6.0),
    ]
    total = 0
    for c in combos:
        p = engine.hough_params(**c)
        total += _check(engine, oracle, plane, rects, p, max(1.0, c["dp"]))
    assert total > 50


def test_hough_golden_cv2(engine):
    """The committed outputs of cv2.HoughCircles itself (tools/make_golden.py)."""
    z = np.load(os.path.join(GOLD, "hough.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    plane = z["plane"]
    for ci, case in enumerate(meta["cases"]):
        p = engine.hough_params(**case["params"])
        got = engine.hough(plane, [tuple(r) for r in meta["rects"]], p)
        for i in range(len(meta["rects"])):
            want = z["c%d_s%d" % (ci, i)]
            rec = got[0, i]
            assert rec["count"] == len(want), (ci, i)
            k = min(len(want), 16)
            assert np.array_equal(rec["xyr"][:k].view(np.uint32), want[:k].view(np.uint32)), (ci, i)


@pytest.mark.parametrize("seed", range(3))
def test_hough_large_squares_global_workspace(engine, oracle, seed):
    """Squares of 129..254 pixels a side (a 4K camera's 1600-pixel board gives 200-pixel squares) do not fit shared
    memory: the same kernel on a global-memory workspace, same circles as the oracle; mixed with small squares and
    several frames per call."""
    plane, rects = synth.shape_atlas(50 + seed, 7, 129, 254, 1024)
    small_plane, small = synth.shape_atlas(60 + seed, 3, 20, 100, 1024)
    H = max(plane.shape[0], small_plane.shape[0])
    both = np.zeros((H + small_plane.shape[0], 1024), np.uint8)
    both[:plane.shape[0]] = plane
    both[H:H + small_plane.shape[0]] = small_plane
    rects = rects + [(x, y + H, w, h) for (x, y, w, h) in small]
    total = 0
    for c in (dict(dp=1.2, param1=100, param2=25), dict(dp=1.05, param1=60, param2=12, min_radius=5, max_radius=90, min_dist=9.0),
              dict(dp=2.0, param1=40, param2=8, min_radius_ratio=0.1, max_radius_ratio=0.5)):
        total += _check(engine, oracle, both, rects, engine.hough_params(**c), max(1.0, c["dp"]))
    assert total > 5
    frames = np.stack([both, both[::-1].copy()])                 # two frames in one call
    got = engine.hough(frames, rects)
    one = engine.hough(both[::-1].copy(), rects)
    assert got[1].tobytes() == one[0].tobytes()


def test_hough_rejects_oversized_squares_and_bad_parameters(engine):
    plane = np.zeros((300, 300), np.uint8)
    with pytest.raises(ValueError):
        engine.hough(plane, [(0, 0, 255, 64)])
    with pytest.raises(ValueError):
        engine.hough(plane, [(0, 0, 254, 254)], engine.hough_params(dp=1.0))     # 254 + 2 accumulator cells a side
    plane = np.zeros((200, 200), np.uint8)
    with pytest.raises(ValueError):
        engine.hough(plane, [(150, 150, 64, 64)])
    with pytest.raises(ValueError):
        engine.hough(plane, [(0, 0, 64, 64)], engine.hough_params(param2=0))
    with pytest.raises(ValueError):
        engine.hough(plane, [(0, 0, 2, 2)])          # minDist = 2 // 3 = 0
    rec = engine.hough(plane, [(0, 0, 64, 64)])[0, 0]
    assert rec["count"] == 0 and rec["n_edges"] == 0 and rec["status"] == 0
