"""GPU parity: warp_image, 64-square split and the per-square PieceDetector /
ChangeDetector statistics through the C ABI vs the CPU oracle (all bit-exact:
integer sums, IEEE f32 sqrt/div/mul/add without contraction)."""
import numpy as np
import pytest

from chessboard_vision_b200 import synth
from chessboard_vision_b200.engine import (grid_rects, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT,
                                           SQ_CD_UPDATE)
from chessboard_vision_b200 import _lib

pytestmark = pytest.mark.gpu


def test_perspective_matrix_matches_oracle(engine, oracle):
    rng = np.random.default_rng(1)
    for i in range(200):
        pts = np.float32([[0, 0], [1920, 0], [0, 1080], [1920, 1080]]) + rng.uniform(-300, 300, (4, 2)).astype(np.float32)
        dst = np.float32([[0, 0], [620, 0], [0, 620], [620, 620]])
        assert np.array_equal(engine.get_perspective_transform(pts, dst), oracle.get_perspective(pts, dst))


@pytest.mark.parametrize("shape", [(480, 640), (1080, 1920), (2160, 3840)])
def test_warp_calibration_corners(engine, oracle, shape):
    H, W = shape
    img = synth.noise_frame(H, W, 0)
    pts = synth.calib_points(H, W)
    M = engine.get_perspective_transform(pts, [[0, 0], [620, 0], [0, 620], [620, 620]])
    assert np.array_equal(engine.warp(img, M, 620), oracle.warp(img, M, 620))


def test_warp_out_of_frame_and_batch(engine, oracle):
    img = synth.frame_batch(2, 240, 320, "noise", 3)
    pts = np.float32([[-50, -30], [300, 10], [20, 260], [350, 250]])
    M = oracle.get_perspective(pts, [[0, 0], [100, 0], [0, 100], [100, 100]])
    got = engine.warp(img, M, 100)
    for i in range(2):
        assert np.array_equal(got[i], oracle.warp(img[i], M, 100))
    # one matrix per frame, rectangular output
    M2 = np.stack([M, oracle.get_perspective(pts + 7, [[0, 0], [100, 0], [0, 100], [100, 100]])])
    got = engine.warp(img, M2, (90, 70))
    for i in range(2):
        assert np.array_equal(got[i], oracle.warp(img[i], M2[i], (90, 70)))


def _oracle_square_stats(oracle, board, rects, ref_plane=None, k=5):
    out = []
    for (x, y, w, h) in rects:
        sq = board[y:y + h, x:x + w]
        g = oracle.square_preprocess(sq, k)
        ref = None if ref_plane is None else np.ascontiguousarray(ref_plane[y:y + h, x:x + w])
        out.append((g, oracle.pd_square_stats(g, ref)))
    return out


def _check_pd(st, o, has_ref):
    assert st["n"] == o["n"] and st["sum"] == o["sum"] and st["sumsq"] == o["sumsq"]
    assert st["center_sum"] == o["center_sum"] and st["center_cnt"] == o["center_cnt"]
    assert st["border_sum"] == o["border_sum"] and st["border_cnt"] == o["border_cnt"]
    assert list(st["ring_sum"]) == o["ring_sum"] and list(st["ring_cnt"]) == o["ring_cnt"]
    assert st["has_ref"] == int(has_ref)
    if has_ref:
        assert st["sad"] == o["sad"]


@pytest.mark.parametrize("grid", ["linear", "smart"])
def test_pd_stats_and_reference(engine, oracle, grid):
    board = synth.board_frame(620, 620, 2)
    board2 = synth.board_frame(620, 620, 3)
    if grid == "linear":
        rects, keys = grid_rects(620)
    else:
        rects, keys = grid_rects(620, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
    assert len(rects) == 64 and keys[0] == (0, 7) and keys[-1] == (7, 0)
    st = engine.new_state(1, 620, 620)
    p = engine.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF)
    s1 = engine.squares(board, rects, p, st)[0]
    o1 = _oracle_square_stats(oracle, board, rects)
    for i in range(64):
        _check_pd(s1[i], o1[i][1], False)
    # reference plane now holds the preprocessed squares
    ref_plane = st.get(0, _lib.PLANE_PD_REF)
    for i, (x, y, w, h) in enumerate(rects):
        assert np.array_equal(ref_plane[y:y + h, x:x + w], o1[i][0])
    # second frame: SAD against the stored reference, no update
    p2 = engine.square_params(ops=SQ_PD_STATS)
    s2 = engine.squares(board2, rects, p2, st)[0]
    o2 = _oracle_square_stats(oracle, board2, rects, ref_plane)
    for i in range(64):
        _check_pd(s2[i], o2[i][1], True)
    # selective reference update
    sel = np.zeros(64, np.uint8); sel[[3, 17]] = 1
    engine.squares(board2, rects, engine.square_params(ops=SQ_PD_SET_REF), st, select=sel, want_stats=False)
    ref2 = st.get(0, _lib.PLANE_PD_REF)
    for i, (x, y, w, h) in enumerate(rects):
        want = o2[i][0] if sel[i] else o1[i][0]
        assert np.array_equal(ref2[y:y + h, x:x + w], want)
    st.free()


@pytest.mark.parametrize("cd_blur", [5, 13, 1])
def test_cd_calibrate_detect_update(engine, oracle, cd_blur):
    rects, _ = grid_rects(620, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
    b0 = synth.board_frame(620, 620, 4)
    b1 = b0.copy()
    b1[100:300, 150:420] = 255          # a "hand"
    b1[400:440, 10:60] //= 2
    st = engine.new_state(2, 620, 620)   # use slot 1 to exercise stream0
    pc = engine.square_params(ops=SQ_CD_CALIBRATE, cd_blur=cd_blur, initial_variance=100)
    engine.squares(b0, rects, pc, st, stream0=1, want_stats=False)
    mean = st.get(1, _lib.PLANE_CD_MEAN); var = st.get(1, _lib.PLANE_CD_VAR)
    om, ov, og1 = {}, {}, {}
    for i, (x, y, w, h) in enumerate(rects):
        g0 = oracle.square_preprocess(b0[y:y + h, x:x + w], cd_blur)
        m, v = oracle.cd_calibrate(g0, 100.0)
        assert np.array_equal(mean[y:y + h, x:x + w], m) and np.array_equal(var[y:y + h, x:x + w], v)
        om[i], ov[i] = m, v
        og1[i] = oracle.square_preprocess(b1[y:y + h, x:x + w], cd_blur)
    # detect + update in one call
    pd = engine.square_params(ops=SQ_CD_DETECT | SQ_CD_UPDATE, cd_blur=cd_blur, z_threshold=2.5, alpha=0.1)
    s = engine.squares(b1, rects, pd, st, stream0=1)[0]
    mean = st.get(1, _lib.PLANE_CD_MEAN); var = st.get(1, _lib.PLANE_CD_VAR)
    for i, (x, y, w, h) in enumerate(rects):
        cnt, zmax = oracle.cd_detect(og1[i], om[i], ov[i], 2.5)
        assert s[i]["cd_valid"] == 1
        assert s[i]["cd_changed"] == cnt
        assert s[i]["cd_zmax"] == np.float32(zmax)
        oracle.cd_update(og1[i], om[i], ov[i], 0.1)
        assert np.array_equal(mean[y:y + h, x:x + w], om[i]) and np.array_equal(var[y:y + h, x:x + w], ov[i])
    # slot 0 never calibrated: CD not valid there
    s0 = engine.squares(b1, rects, engine.square_params(ops=SQ_CD_DETECT, cd_blur=cd_blur), st, stream0=0)[0]
    assert not s0["cd_valid"].any()
    st.free()


def test_cd_focus_squares_and_gray_atlas(engine, oracle):
    """test_change_detector_regression.py:31-54 in kernel form: 64 gray 50x50 squares, one goes 0 -> 255."""
    rects = [(c * 50, r * 50, 50, 50) for r in range(8) for c in range(8)]
    atlas0 = np.zeros((400, 400), np.uint8)
    atlas1 = atlas0.copy(); atlas1[3 * 50:4 * 50, 3 * 50:4 * 50] = 255
    st = engine.new_state(1, 400, 400)
    engine.squares(atlas0, rects, engine.square_params(ops=SQ_CD_CALIBRATE), st, want_stats=False)
    s = engine.squares(atlas1, rects, engine.square_params(ops=SQ_CD_DETECT), st)[0]
    idx = 3 * 8 + 3
    pct = s["cd_changed"] / s["n"] * 100
    assert pct[idx] > 75 and (np.delete(pct, idx) == 0).all()
    # focus: only selected squares are evaluated
    sel = np.zeros(64, np.uint8); sel[5] = 1
    s = engine.squares(atlas1, rects, engine.square_params(ops=SQ_CD_DETECT), st, select=sel)[0]
    assert s["cd_valid"][5] == 1 and s["cd_valid"].sum() == 1
    st.free()


def test_squares_errors(engine):
    board = np.zeros((100, 100, 3), np.uint8)
    with pytest.raises(ValueError):
        engine.squares(board, [(90, 90, 20, 20)], engine.square_params())
    with pytest.raises(Exception):
        engine.squares(board, [(0, 0, 20, 20)], engine.square_params(ops=SQ_CD_DETECT))   # needs state


def test_full_pipeline_matches_composition(engine, oracle):
    """cvb_pipeline (host buffers) == oracle composition enhance -> warp -> squares, on 2 small frames."""
    H, W = 270, 480
    frames = synth.frame_batch(2, H, W, "board", 20)
    pts = synth.calib_points(H, W)
    S = 160
    M = oracle.get_perspective(pts, [[0, 0], [S, 0], [0, S], [S, S]])
    rects, _ = grid_rects(S)
    st = engine.new_state(2, S, S)
    pp = engine.pipeline_params(squares=engine.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE | SQ_CD_DETECT),
                                board_size=S)
    T, stats = engine.pipeline(frames, M, rects, pp, st)
    for i in range(2):
        enh = oracle.process_pipeline(frames[i], True)
        _, _, rT, _ = oracle.prepare_analysis(enh, True)
        assert T[i] == rT
        board = oracle.warp(enh, M, S)
        o = _oracle_square_stats(oracle, board, rects)
        for j in range(64):
            _check_pd(stats[i, j], o[j][1], False)
            assert stats[i, j]["cd_valid"] == 1 and stats[i, j]["cd_changed"] == 0
    st.free()


@pytest.mark.parametrize("grid", ["linear", "smart", "odd"])
def test_cd_special_values_and_partial_groups(engine, oracle, grid):
    """The 4-pixels-per-thread square kernel on squares that do not start on its 4-pixel grid, with state planes
    holding the values its shortcuts must respect: mean == gray (zero numerator) over variances 0, -0, negative,
    NaN, +inf, denormal and ordinary, and NaN / inf means.  Counts, z-max (NaN propagates as in np.max) and the
    updated planes equal the oracle bit for bit; pixels of neighbouring squares stay untouched."""
    S = 620 if grid != "odd" else 124
    board = synth.board_frame(S, S, 21)
    if grid == "linear":
        rects, _ = grid_rects(S)
    elif grid == "smart":
        rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
    else:
        rects = [(1, 2, 13, 9), (17, 3, 5, 3), (23, 1, 3, 17), (30, 30, 41, 37), (75, 5, 46, 50), (2, 60, 7, 60)]
    st = engine.new_state(1, S, S)
    engine.squares(board, rects, engine.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), st)
    mean, var = st.get(0, _lib.PLANE_CD_MEAN), st.get(0, _lib.PLANE_CD_VAR)
    specials = np.array([0.0, -0.0, -3.0, np.nan, np.inf, 1e-40, 7.5, 100.0], np.float32)
    rng = np.random.default_rng(5)
    var[...] = specials[rng.integers(0, len(specials), var.shape)]
    weird = rng.random(mean.shape) < 0.02
    mean[weird] = np.array([np.nan, np.inf, -np.inf], np.float32)[rng.integers(0, 3, int(weird.sum()))]
    st.set(0, _lib.PLANE_CD_MEAN, mean); st.set(0, _lib.PLANE_CD_VAR, var)
    before_cur = st.get(0, _lib.PLANE_PD_CUR)
    stats = engine.squares(board, rects, engine.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), st)
    m2, v2 = st.get(0, _lib.PLANE_CD_MEAN), st.get(0, _lib.PLANE_CD_VAR)
    inside = np.zeros((S, S), bool)
    for j, (x, y, w, h) in enumerate(rects):
        g = oracle.square_preprocess(board[y:y + h, x:x + w], 5)
        m, v = np.ascontiguousarray(mean[y:y + h, x:x + w]), np.ascontiguousarray(var[y:y + h, x:x + w])
        cnt, zmax = oracle.cd_detect(g, m, v, 2.5)
        assert int(stats[0, j]["cd_changed"]) == cnt, (grid, j)
        assert np.float32(stats[0, j]["cd_zmax"]).tobytes() == np.float32(zmax).tobytes() or (np.isnan(zmax) and np.isnan(stats[0, j]["cd_zmax"]))
        oracle.cd_update(g, m, v, 0.1)          # in place
        # bit-equal, except that a NaN may carry another payload (x86 keeps the operand's, the GPU returns the canonical one)
        same = lambda a, b: np.array_equal(a.view(np.uint32)[~np.isnan(a)], b.view(np.uint32)[~np.isnan(b)]) and \
            np.array_equal(np.isnan(a), np.isnan(b))
        assert same(m, np.ascontiguousarray(m2[y:y + h, x:x + w])), (grid, j)
        assert same(v, np.ascontiguousarray(v2[y:y + h, x:x + w])), (grid, j)
        assert np.array_equal(st.get(0, _lib.PLANE_PD_CUR)[y:y + h, x:x + w], g)
        inside[y:y + h, x:x + w] = True
    # nothing outside the squares was written
    assert m2[~inside].tobytes() == mean[~inside].tobytes() and v2[~inside].tobytes() == var[~inside].tobytes()
    assert np.array_equal(st.get(0, _lib.PLANE_PD_CUR)[~inside], before_cur[~inside])
    st.free()


def test_warp_extreme_coordinates(engine, oracle):
    """Homographies whose horizon crosses the destination: W passes through 0, source coordinates overflow int32 or
    become infinite; the saturating conversion of the kernel equals imgwarp.cpp's clamp + cvRound (the oracle)."""
    img = synth.noise_frame(120, 160, 4)
    rng = np.random.default_rng(9)
    mats = []
    for i in range(12):
        M = np.eye(3)
        M[2, 0], M[2, 1] = rng.uniform(-0.05, 0.05, 2)          # W = 0 somewhere inside the 100 x 100 output
        M[2, 2] = rng.uniform(-1.0, 1.0)
        M[0, 2], M[1, 2] = rng.uniform(-1e6, 1e6, 2)
        M[0, 0] *= 10.0 ** rng.integers(0, 9)
        mats.append(M)
    mats.append(np.array([[1, 0, 0], [0, 1, 0], [0.01, 0, -0.5]], np.float64))      # exact zero of W at x = 50
    mats.append(np.array([[1e300, 0, 0], [0, 1e300, 0], [0, 0, 1e-300]], np.float64))  # products overflow to inf
    for M in mats:
        assert np.array_equal(engine.warp(img, M, 100), oracle.warp(img, M, 100)), M
