"""CPU: the C oracle against live OpenCV (skipped when cv2 is not importable).  These are the
survey probes kept as permanent tests: exhaustive LAB both ways, CLAHE, sharpen, normalize for
every (min,max), gray, Gaussian for k=1..31, Otsu, warp, perspective matrices, bilateral <= 1 LSB."""
import numpy as np
import pytest

from chessboard_vision_b200 import synth

cv2 = pytest.importorskip("cv2")


def test_lab_exhaustive_both_ways(oracle):
    a = np.arange(1 << 24, dtype=np.uint32)
    allc = np.stack([a & 255, (a >> 8) & 255, (a >> 16) & 255], -1).astype(np.uint8).reshape(4096, 4096, 3)
    assert np.array_equal(oracle.bgr2lab(allc), cv2.cvtColor(allc, cv2.COLOR_BGR2LAB))
    assert np.array_equal(oracle.lab2bgr(allc), cv2.cvtColor(allc, cv2.COLOR_LAB2BGR))


@pytest.mark.parametrize("shape", [(480, 640), (477, 643), (100, 100), (33, 70), (480, 643), (477, 640)])
@pytest.mark.parametrize("kind", ["board", "noise"])
def test_clahe(oracle, shape, kind):
    img = synth.board_frame(*shape, 1) if kind == "board" else synth.noise_frame(*shape, 1)
    L = np.ascontiguousarray(cv2.cvtColor(img, cv2.COLOR_BGR2LAB)[..., 0])
    assert np.array_equal(oracle.clahe(L), cv2.createCLAHE(3.0, (8, 8)).apply(L))


@pytest.mark.parametrize("clip,tiles", [(2.0, (4, 4)), (40.0, (8, 8)), (1.0, (16, 8))])
def test_clahe_params(oracle, clip, tiles):
    L = np.ascontiguousarray(synth.board_frame(480, 640, 5)[..., 1])
    assert np.array_equal(oracle.clahe(L, clip, tiles), cv2.createCLAHE(clip, tiles).apply(L))


def test_sharpen_gray(oracle):
    k = np.array([[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]])
    for shape in [(480, 640), (477, 643), (3, 3), (2, 5), (1, 7)]:
        img = synth.noise_frame(*shape, 4)
        assert np.array_equal(oracle.sharpen(img), cv2.filter2D(img, -1, k))
    img = synth.noise_frame(1080, 1920, 5)
    assert np.array_equal(oracle.gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))


def test_normalize_every_range(oracle):
    bad = 0
    for lo in range(0, 256, 3):
        for hi in range(lo, 256, 2):
            vals = np.arange(lo, hi + 1, dtype=np.uint8)
            arr = np.resize(vals, len(vals) * 3 + 67)
            arr[0], arr[-1] = lo, hi
            ref = cv2.normalize(arr.reshape(1, -1), None, alpha=0, beta=255, norm_type=cv2.NORM_MINMAX).ravel()
            bad += np.count_nonzero(oracle.normalize_lut(lo, hi, True)[arr] != ref)
    assert bad == 0


@pytest.mark.parametrize("k", [1, 3, 5, 7, 9, 11, 13, 15, 17, 21, 31])
def test_gaussian(oracle, k):
    g = cv2.cvtColor(synth.noise_frame(200, 300, 5), cv2.COLOR_BGR2GRAY)
    for (H, W) in [(200, 300), (77, 77), (50, 50), (5, 9), (3, 3)]:
        gg = np.ascontiguousarray(g[:H, :W])
        assert np.array_equal(oracle.gaussian(gg, k), cv2.GaussianBlur(gg, (k, k), 0))
    v = g[100:177, 200:277]          # a non-contiguous view, as split_board hands out
    assert np.array_equal(oracle.gaussian(v, k), cv2.GaussianBlur(v, (k, k), 0))


def test_otsu(oracle):
    for img in (synth.board_frame(480, 640, 7), synth.noise_frame(480, 640, 7)):
        g, b, T, bl = oracle.prepare_analysis(img, True)
        rb = cv2.GaussianBlur(cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), (5, 5), 0)
        rT, rbin = cv2.threshold(rb, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert T == rT and np.array_equal(b, rbin)
    for arr in [np.zeros((10, 10), np.uint8), np.full((10, 10), 255, np.uint8), np.array([[0, 255] * 8] * 4, np.uint8),
                np.array([[3] * 15 + [200]], np.uint8), np.arange(256, dtype=np.uint8)[None, :].repeat(4, 0)]:
        rT, _ = cv2.threshold(arr, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        assert oracle.otsu_from_hist(oracle.hist256(arr), arr.size) == rT


def test_perspective_and_warp(oracle):
    rng = np.random.default_rng(0)
    dst = np.float32([[0, 0], [620, 0], [0, 620], [620, 620]])
    for i in range(300):
        pts = np.float32([[0, 0], [1920, 0], [0, 1080], [1920, 1080]]) + rng.uniform(-300, 300, (4, 2)).astype(np.float32)
        M = cv2.getPerspectiveTransform(pts, dst)
        assert np.array_equal(oracle.get_perspective(pts, dst), M)
        assert np.array_equal(oracle.invert3(M), cv2.invert(M)[1])
    img = synth.noise_frame(1080, 1920, 0)
    for pts in (synth.calib_points(1080, 1920), np.float32([[-50, -30], [1000, 10], [20, 1100], [1950, 1090]])):
        M = cv2.getPerspectiveTransform(pts, dst)
        assert np.array_equal(oracle.warp(img, M, 620), cv2.warpPerspective(img, M, (620, 620)))


def test_bilateral_within_one_lsb(oracle):
    for img in (synth.board_frame(300, 301, 11), synth.noise_frame(300, 301, 11)):
        ref = cv2.bilateralFilter(img, 9, 75, 75)
        d = np.abs(oracle.bilateral(img, use_fma=True).astype(int) - ref)
        assert d.max() <= 1 and np.count_nonzero(d) <= 20


def test_ref_cv2_sequence_matches_oracle(oracle):
    """oracle/ref_cv2.py (what the CPU baseline times) against the C restatement."""
    from oracle import ref_cv2
    img = synth.board_frame(240, 320, 3)
    lit = ref_cv2.correct_lighting(img)
    assert np.array_equal(lit, oracle.correct_lighting(img))
    g, b, t = ref_cv2.prepare_analysis(lit)
    og, ob, oT, _ = oracle.prepare_analysis(lit, True)
    assert np.array_equal(g, og) and np.array_equal(b, ob) and t == oT
    sq = np.ascontiguousarray(img[:77, :77])
    gq = ref_cv2.preprocess_square(sq, 5)
    assert np.array_equal(gq, oracle.square_preprocess(sq, 5))
    from chessboard_vision_b200 import hostapi
    st = oracle.pd_square_stats(gq, np.ascontiguousarray(gq[::-1]))
    r = ref_cv2.pd_statistics(gq, np.ascontiguousarray(gq[::-1]))
    assert r["mean_diff"] == hostapi.mean_abs_diff(st)
    assert (r["center_border_diff"], r["center_mean"], r["border_mean"]) == hostapi.center_vs_border(st)
    assert r["symmetry"] == hostapi.radial_symmetry(st)
    m, v = oracle.cd_calibrate(gq, 100.0)
    g2 = oracle.square_preprocess(np.ascontiguousarray(img[77:154, :77]), 5)
    changed, pct, zmax = ref_cv2.cd_detect(g2, m, v)
    cnt, oz = oracle.cd_detect(g2, m, v, 2.5)
    assert changed == cnt and np.float32(zmax) == np.float32(oz)
    nm, nv = ref_cv2.cd_update(g2, m, v)
    oracle.cd_update(g2, m, v, 0.1)
    assert np.array_equal(nm, m) and np.array_equal(nv, v)
