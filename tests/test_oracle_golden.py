"""CPU: the C oracle against golden vectors produced by the UNMODIFIED reference modules
(tools/make_golden.py; tests/golden/).  Bit-exact for every integer / LUT stage; the bilateral
filter within 1 LSB (BASELINE.json tolerance) with a bounded mismatch count."""
import hashlib
import json
import os

import numpy as np
import pytest

from chessboard_vision_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(G, "enhancer_small.npz"))


@pytest.fixture(scope="module")
def kat():
    return json.load(open(os.path.join(G, "kat.json")))["kat"]


NAMES = ["board_96x128", "noise_90x121", "board_120x160"]


@pytest.mark.parametrize("name", NAMES)
def test_stage_isolated(oracle, small, name):
    g = lambda k: small[name + "/" + k]
    img = g("input")
    assert np.array_equal(oracle.bgr2lab(img), g("lab"))
    assert np.array_equal(oracle.clahe(np.ascontiguousarray(g("lab")[..., 0])), g("clahe_l"))
    assert np.array_equal(oracle.correct_lighting(img), g("correct_lighting"))
    bil = oracle.bilateral(g("correct_lighting"), use_fma=True)
    d = np.abs(bil.astype(int) - g("reduce_noise"))
    assert d.max() <= 1 and np.count_nonzero(d) <= max(4, d.size // 50000)   # <= 1 LSB, a handful of values
    # later stages fed the reference's own previous output: bit-exact
    assert np.array_equal(oracle.sharpen(g("reduce_noise")), g("sharpen"))
    assert np.array_equal(oracle.normalize(g("sharpen"), True), g("normalize"))
    gray, binary, T, blur = oracle.prepare_analysis(g("normalize"), return_all=True)
    assert np.array_equal(gray, g("gray")) and np.array_equal(blur, g("blur"))
    assert T == int(g("otsu_t")) and np.array_equal(binary, g("binary"))


@pytest.mark.parametrize("size", ["640x480", "1920x1080"])
def test_known_answers_full_size(oracle, kat, size):
    W, H = map(int, size.split("x"))
    k = kat[size]
    img = synth.noise_frame(H, W, 0)
    assert sha(img) == k["input"]
    lab = oracle.bgr2lab(img)
    assert sha(lab) == k["lab"]
    l = np.ascontiguousarray(lab[..., 0])
    out, hist, lut = oracle.clahe(l, return_tables=True)
    assert sha(out) == k["clahe_l"]
    assert hist[0, :4].tolist() == k["tile00_hist_0_4"]
    assert sha(oracle.correct_lighting(img)) == k["correct_lighting"]
    M = np.array(k["warp_matrix"]).reshape(3, 3)
    assert np.array_equal(oracle.get_perspective(synth.calib_points(H, W), [[0, 0], [620, 0], [0, 620], [620, 620]]), M)
    assert sha(oracle.warp(img, M, 620)) == k["warp"]


def test_stages_after_reference_bilateral_640x480(oracle, kat):
    k = kat["640x480"]
    bil = np.load(os.path.join(G, "bilateral_640x480.npz"))["reduce_noise"]
    assert sha(bil) == k["reduce_noise_ipp"]
    shp = oracle.sharpen(bil)
    assert sha(shp) == k["sharpen_of_ref_bilateral"]
    nrm = oracle.normalize(shp, True)
    assert sha(nrm) == k["normalize_of_ref"]
    gray, binary, T, blur = oracle.prepare_analysis(nrm, return_all=True)
    assert sha(gray) == k["gray_of_ref"] and sha(blur) == k["blur_of_ref"]
    assert T == k["otsu_t_of_ref"] and sha(binary) == k["binary_of_ref"]
    assert int(np.count_nonzero(binary)) == k["white_px_of_ref"]
    # and the oracle's own bilateral is within 1 LSB of the reference's on the full frame
    mine = oracle.bilateral(oracle.correct_lighting(synth.noise_frame(480, 640, 0)), use_fma=True)
    d = np.abs(mine.astype(int) - bil)
    assert d.max() <= 1 and np.count_nonzero(d) < 40


def test_warp_small(oracle):
    z = np.load(os.path.join(G, "warp_small.npz"))
    img = synth.noise_frame(270, 480, 9)
    w, M, S = oracle.warp_image(img, z["points"], display_size=(200, 180), margin=20)
    assert S == int(z["board_size"]) and np.array_equal(M, z["matrix"]) and np.array_equal(w, z["warped"])


def test_piece_detector_statistics(oracle):
    from chessboard_vision_b200 import hostapi
    z = np.load(os.path.join(G, "piece_detector.npz"))
    gray_sha = json.loads(str(z["gray_sha"]))
    board, pieces = synth.board_with_pieces(11, 7, 620)
    for row in z["stats"]:
        c, r = int(row[0]), int(row[1])
        y, x = (7 - r) * 77, c * 77
        g = oracle.square_preprocess(pieces[y:y + 77, x:x + 77], 5)
        assert sha(g) == gray_sha["%d_%d" % (c, r)]
        ref = oracle.square_preprocess(board[y:y + 77, x:x + 77], 5)
        st = oracle.pd_square_stats(g, ref)
        assert hostapi.mean_abs_diff(st) == row[2]
        assert float(hostapi.mean_abs_diff(st) > 25) == row[3]
        assert abs(hostapi.std_from_moments(st) - row[4]) <= 1e-12 * max(1.0, row[4])
        assert hostapi.std_below(st, 15) == (row[4] < 15)
        diff, cm, bm = hostapi.center_vs_border(st)
        assert (diff, cm, bm) == (row[5], row[6], row[7])
        assert hostapi.radial_symmetry(st) == row[8]


def test_change_detector_numerics(oracle):
    cdj = json.load(open(os.path.join(G, "change_detector.json")))
    board, pieces = synth.board_with_pieces(11, 7, 620)
    for key, want in cdj["detailed"].items():
        c, r = map(int, key.split("_"))
        y, x = (7 - r) * 77, c * 77
        m, v = oracle.cd_calibrate(oracle.square_preprocess(board[y:y + 77, x:x + 77], 5), 100.0)
        assert sha(m) == cdj["means_sha"][key]
        g = oracle.square_preprocess(pieces[y:y + 77, x:x + 77], 5)
        cnt, zmax = oracle.cd_detect(g, m, v, 2.5)
        assert (cnt / g.size) * 100 == want["pct_changed"]
        assert float(np.float32(zmax)) == want["z_score"]
        oracle.cd_update(g, m, v, 0.1)
        assert sha(m) == cdj["means_after_update_sha"][key] and sha(v) == cdj["vars_after_update_sha"][key]
    # squares below the 5 % gate are absent from `detailed`; their model still updates
    for key in ("0_0", "7_7", "3_4"):
        if key in cdj["detailed"]:
            continue
        c, r = map(int, key.split("_"))
        y, x = (7 - r) * 77, c * 77
        m, v = oracle.cd_calibrate(oracle.square_preprocess(board[y:y + 77, x:x + 77], 5), 100.0)
        g = oracle.square_preprocess(pieces[y:y + 77, x:x + 77], 5)
        cnt, _ = oracle.cd_detect(g, m, v, 2.5)
        assert cnt / g.size * 100 < 5.0
        oracle.cd_update(g, m, v, 0.1)
        assert sha(m) == cdj["means_after_update_sha"][key] and sha(v) == cdj["vars_after_update_sha"][key]


def test_reference_regression_vector(oracle):
    """test_change_detector_regression.py:31-54: zeros -> one square 255 => 100 % TOTAL."""
    cdj = json.load(open(os.path.join(G, "change_detector.json")))
    g0 = oracle.square_preprocess(np.zeros((50, 50), np.uint8), 5)
    m, v = oracle.cd_calibrate(g0, 100.0)
    g1 = oracle.square_preprocess(np.full((50, 50), 255, np.uint8), 5)
    cnt, zmax = oracle.cd_detect(g1, m, v, 2.5)
    want = cdj["regression_3_3"]
    assert cnt / g1.size * 100 == want["pct_changed"] == 100.0 and want["intensity"] == "TOTAL"
    assert float(np.float32(zmax)) == want["z_score"]
    half = np.zeros((77, 77), np.uint8); half[:40] = 255
    m, v = oracle.cd_calibrate(oracle.square_preprocess(np.zeros((77, 77), np.uint8), 5), 100.0)
    cnt, _ = oracle.cd_detect(oracle.square_preprocess(half, 5), m, v, 2.5)
    assert cnt / half.size * 100 == cdj["top40rows_1_1"]["pct_changed"]
