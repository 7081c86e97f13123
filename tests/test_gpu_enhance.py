"""GPU parity: frame_enhancer stages through the C ABI vs the CPU oracle.

Bit-exact classes (assert array_equal): LAB both ways, CLAHE histograms / LUTs /
output, sharpen, min-max normalize, gray, Gaussian, Otsu threshold, mask.
Bilateral: the CUDA kernel follows the oracle's tap order and rounding, so it is
compared bit-exactly with the ORACLE; the oracle itself is within 1 LSB of cv2
(tests/test_oracle_vs_cv2.py, golden vectors), which is the tolerance
BASELINE.json states (max abs error <= 1 LSB per uint8 channel).
"""
import numpy as np
import pytest

from chessboard_vision_b200 import synth

pytestmark = pytest.mark.gpu

SHAPES = [(48, 64), (120, 160), (477, 643), (480, 640), (97, 33), (30, 60), (31, 61)]


def frames(kind, H, W, seed=3):
    return synth.board_frame(H, W, seed) if kind == "board" else synth.noise_frame(H, W, seed)


def test_lab_exhaustive(engine, oracle):
    a = np.arange(1 << 24, dtype=np.uint32)
    allc = np.stack([a & 255, (a >> 8) & 255, (a >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(engine.bgr2lab(allc), oracle.bgr2lab(allc))
    assert np.array_equal(engine.lab2bgr(allc), oracle.lab2bgr(allc))


@pytest.mark.parametrize("kind", ["board", "noise"])
@pytest.mark.parametrize("shape", SHAPES)
def test_clahe_plane(engine, oracle, kind, shape):
    L = np.ascontiguousarray(frames(kind, *shape)[..., 1])
    out, hist, lut = engine.clahe(L, 3.0, (8, 8), return_tables=True)
    ref, rhist, rlut = oracle.clahe(L, 3.0, (8, 8), return_tables=True)
    assert np.array_equal(hist, rhist)
    assert np.array_equal(lut, rlut)
    assert np.array_equal(out, ref)


@pytest.mark.parametrize("clip,tiles", [(2.0, (4, 4)), (40.0, (8, 8)), (1.0, (16, 8)), (0.0, (8, 8))])
def test_clahe_params(engine, oracle, clip, tiles):
    L = np.ascontiguousarray(synth.board_frame(240, 322, 5)[..., 0])
    assert np.array_equal(engine.clahe(L, clip, tiles), oracle.clahe(L, clip, tiles))


def test_clahe_constant_and_batch(engine, oracle):
    for v in (0, 128, 255):
        L = np.full((96, 128), v, np.uint8)
        assert np.array_equal(engine.clahe(L), oracle.clahe(L))
    batch = np.stack([np.ascontiguousarray(synth.noise_frame(96, 128, s)[..., 0]) for s in range(3)])
    out = engine.clahe(batch)
    for i in range(3):
        assert np.array_equal(out[i], oracle.clahe(batch[i]))


@pytest.mark.parametrize("kind", ["board", "noise"])
@pytest.mark.parametrize("shape", SHAPES)
def test_correct_lighting(engine, oracle, kind, shape):
    img = frames(kind, *shape)
    assert np.array_equal(engine.correct_lighting(img), oracle.correct_lighting(img))


@pytest.mark.parametrize("kind", ["board", "noise"])
@pytest.mark.parametrize("shape", SHAPES + [(9, 9), (5, 200), (200, 5)])
def test_bilateral(engine, oracle, kind, shape):
    img = frames(kind, *shape)
    got = engine.bilateral(img)
    ref = oracle.bilateral(img, use_fma=True)
    assert np.array_equal(got, ref)


def test_bilateral_sigmas_and_bad_d(engine, oracle):
    img = synth.board_frame(64, 80, 1)
    assert np.array_equal(engine.bilateral(img, 9, 30.0, 10.0), oracle.bilateral(img, 9, 30.0, 10.0, True))
    assert np.array_equal(engine.bilateral(img, 9, 75.0, 75.0), oracle.bilateral(img, 9, 75.0, 75.0, True))
    with pytest.raises(ValueError):
        engine.bilateral(img, 5, 75.0, 75.0)


@pytest.mark.parametrize("shape", SHAPES + [(3, 3), (2, 5), (1, 7)])
def test_sharpen(engine, oracle, shape):
    img = synth.noise_frame(*shape, seed=4)
    assert np.array_equal(engine.sharpen(img), oracle.sharpen(img))


def test_normalize(engine, oracle):
    rng = np.random.default_rng(0)
    for _ in range(12):
        lo = int(rng.integers(0, 200)); hi = int(rng.integers(lo, 256))
        img = rng.integers(lo, hi + 1, (37, 53, 3), dtype=np.uint8)
        got, mm = engine.normalize(img, return_minmax=True)
        ref, mn, mx = oracle.normalize(img, True, return_minmax=True)
        assert (mm[0, 0], mm[0, 1]) == (mn, mx)
        assert np.array_equal(got, ref)
    const = np.full((16, 16, 3), 77, np.uint8)          # max == min edge case
    assert np.array_equal(engine.normalize(const), oracle.normalize(const))


def test_normalize_all_ranges(engine, oracle):
    # every (min,max) pair on one batch: frame i holds all values of its range
    pairs = [(lo, hi) for lo in range(0, 256, 5) for hi in range(lo, 256, 7)]
    batch = np.empty((len(pairs), 16, 16, 3), np.uint8)
    for i, (lo, hi) in enumerate(pairs):
        vals = np.arange(lo, hi + 1, dtype=np.uint8)
        batch[i] = np.resize(vals, 16 * 16 * 3).reshape(16, 16, 3)
        batch[i, 0, 0, 0] = lo; batch[i, -1, -1, -1] = hi
    got = engine.normalize(batch)
    for i, (lo, hi) in enumerate(pairs):
        assert np.array_equal(got[i], oracle.normalize_lut(lo, hi, True)[batch[i]])


@pytest.mark.parametrize("shape", SHAPES)
def test_gray(engine, oracle, shape):
    img = synth.noise_frame(*shape, seed=6)
    assert np.array_equal(engine.gray(img), oracle.gray(img))


@pytest.mark.parametrize("k", [1, 3, 5, 7, 9, 11, 13, 15, 21, 31])
@pytest.mark.parametrize("shape", [(120, 160), (77, 77), (50, 50), (5, 9), (3, 3), (1, 1)])
def test_gaussian(engine, oracle, k, shape):
    g = np.ascontiguousarray(synth.noise_frame(*shape, seed=7)[..., 0])
    assert np.array_equal(engine.gaussian(g, k), oracle.gaussian(g, k))


def test_gaussian_bad_k(engine):
    g = np.zeros((8, 8), np.uint8)
    for k in (0, 2, 33):
        with pytest.raises(ValueError):
            engine.gaussian(g, k)


@pytest.mark.parametrize("kind", ["board", "noise"])
@pytest.mark.parametrize("shape", SHAPES)
def test_prepare_analysis(engine, oracle, kind, shape):
    img = frames(kind, *shape)
    g, b, T, bl, hist = engine.prepare_analysis(img, return_all=True)
    rg, rb, rT, rbl = oracle.prepare_analysis(img, return_all=True)
    assert np.array_equal(g, rg)
    assert np.array_equal(bl, rbl)
    assert np.array_equal(hist, oracle.hist256(rbl))
    assert T == rT
    assert np.array_equal(b, rb)


def test_otsu_degenerate(engine, oracle):
    cases = [np.zeros((10, 10, 3), np.uint8), np.full((10, 10, 3), 255, np.uint8),
             np.repeat(np.array([[0, 255] * 8] * 4, np.uint8)[..., None], 3, -1),
             np.repeat(np.arange(256, dtype=np.uint8)[None, :].repeat(4, 0)[..., None], 3, -1)]
    for img in cases:
        img = np.ascontiguousarray(img)
        g, b, T, bl, hist = engine.prepare_analysis(img, return_all=True)
        rg, rb, rT, rbl = oracle.prepare_analysis(img, return_all=True)
        assert T == rT and np.array_equal(b, rb)


@pytest.mark.parametrize("kind", ["board", "noise"])
@pytest.mark.parametrize("shape", [(48, 64), (120, 160), (477, 643), (480, 640)])
def test_process_pipeline_and_enhance(engine, oracle, kind, shape):
    img = frames(kind, *shape)
    ref = oracle.process_pipeline(img, use_fma=True)
    got = engine.process_pipeline(img)
    assert np.array_equal(got, ref)
    enh, g, b, T = engine.enhance(img)          # host-buffer entry point
    rg, rb, rT, _ = oracle.prepare_analysis(ref, return_all=True)
    assert np.array_equal(enh, ref)
    assert np.array_equal(g, rg) and np.array_equal(b, rb) and T == rT


def test_enhance_batch_matches_single(engine):
    batch = synth.frame_batch(3, 90, 120, "board", 10)
    enh, g, b, T = engine.enhance(batch)
    for i in range(3):
        e1, g1, b1, T1 = engine.enhance(batch[i])
        assert np.array_equal(enh[i], e1) and np.array_equal(g[i], g1) and np.array_equal(b[i], b1) and T[i] == T1


def test_enhance_1080p_properties(engine, oracle):
    """Full-size checks through size-independent properties + one oracle frame."""
    img = synth.board_frame(1080, 1920, 0)
    enh, g, b, T = engine.enhance(img)
    ref = oracle.process_pipeline(img, use_fma=True)
    assert np.array_equal(enh, ref)
    rg, rb, rT, _ = oracle.prepare_analysis(ref, return_all=True)
    assert T == rT and np.array_equal(b, rb) and np.array_equal(g, rg)
    assert set(np.unique(b)) <= {0, 255}
    # idempotence of normalize on an already full-range image
    assert enh.min() == 0 and enh.max() == 255
    assert np.array_equal(engine.normalize(enh), enh)


def test_enhance_4k_vs_oracle_and_stream_state(engine, oracle):
    """BASELINE.json configs[4] shape: 3840x2160 frames, per-stream state resident on the GPU."""
    from chessboard_vision_b200.engine import grid_rects, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE
    H, W, S = 2160, 3840, 620
    f0 = synth.board_frame(H, W, 40)
    f1 = f0.copy(); f1[600:1500, 1400:2400] = 255
    enh, g, b, T = engine.enhance(f0)
    ref = oracle.process_pipeline(f0, use_fma=True)
    assert np.array_equal(enh, ref)
    rg, rb, rT, _ = oracle.prepare_analysis(ref, return_all=True)
    assert T == rT and np.array_equal(b, rb) and np.array_equal(g, rg)
    # two consecutive frames of one 4K stream through the whole path (state slot 3 of 4)
    M = engine.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
    rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
    st = engine.new_state(4, S, S)
    cal = engine.pipeline_params(squares=engine.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), board_size=S)
    run = engine.pipeline_params(squares=engine.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), board_size=S)
    engine.pipeline(f0[None], M, rects, cal, st, stream0=3)
    T1, s1 = engine.pipeline(f1[None], M, rects, run, st, stream0=3)
    board0 = oracle.warp(ref, M, S)
    board1 = oracle.warp(oracle.process_pipeline(f1, True), M, S)
    for j, (x, y, w, h) in enumerate(rects):
        g0 = oracle.square_preprocess(board0[y:y + h, x:x + w], 5)
        g1 = oracle.square_preprocess(board1[y:y + h, x:x + w], 5)
        o = oracle.pd_square_stats(g1, g0)
        m, v = oracle.cd_calibrate(g0, 100.0)
        cnt, zmax = oracle.cd_detect(g1, m, v, 2.5)
        r = s1[0, j]
        assert (r["sum"], r["sumsq"], r["sad"], r["has_ref"]) == (o["sum"], o["sumsq"], o["sad"], 1)
        assert r["cd_valid"] == 1 and r["cd_changed"] == cnt and r["cd_zmax"] == np.float32(zmax)
    st.free()


def test_two_devices(engine):
    """Engines on two GPUs in one process: every entry point runs on its handle's own device whatever device the thread
    had current (only cvb_create / cvb_malloc used to select the device)."""
    from chessboard_vision_b200 import _lib
    from chessboard_vision_b200.engine import Engine
    if _lib.load().cvb_device_count() < 2:
        pytest.skip("needs two GPUs")
    import oracle as O
    other = Engine(1)                      # leaves device 1 current in this thread
    try:
        f = synth.board_frame(135, 240, 3)
        ref = O.process_pipeline(f, True)
        assert np.array_equal(engine.process_pipeline(f), ref)        # handle of device 0, first call after Engine(1)
        assert np.array_equal(other.process_pipeline(f), ref)
        st0, st1 = engine.new_state(1, 64, 64), other.new_state(1, 64, 64)
        a = engine.enhance(f); b = other.enhance(f)
        assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3])) and a[3] == b[3]
        st0.free(); st1.free()
    finally:
        other.close()
