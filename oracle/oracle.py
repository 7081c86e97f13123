"""ctypes front-end of oracle/cvb_oracle.c (CPU oracle, test infrastructure).

Every function mirrors one stage of the reference hot path; the reference
file:line each stands in for is cited in cvb_oracle.c.  Parity pinning: see
tests/test_oracle_golden.py (golden vectors produced by the unmodified
reference, tools/make_golden.py) and tests/test_oracle_vs_cv2.py (live
cv2, when importable).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcvb_oracle.so")


def build(force=False):
    """Compile the C restatement (gcc, seconds)."""
    src = os.path.join(_HERE, "cvb_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_otsu_from_hist.restype = C.c_int
        _lib.orc_prepare_analysis.restype = C.c_int
        _lib.orc_bilateral_tables.restype = C.c_int
        _lib.orc_gaussian.restype = C.c_int
        _lib.orc_gaussian_kernel_q8.restype = C.c_int
        _lib.orc_get_perspective.restype = C.c_int
        _lib.orc_invert3.restype = C.c_int
        _lib.orc_square_preprocess.restype = C.c_int
        _lib.orc_tables_init()
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == np.uint8
    return a


# --- tables -----------------------------------------------------------------
def tables():
    g = np.empty(256, np.uint16); c = np.empty(3072, np.uint16)
    y = np.empty(512, np.int32); ig = np.empty(4096, np.uint8)
    lib().orc_get_tables(_p(g), _p(c), _p(y), _p(ig))
    return {"gamma": g, "cbrt": c, "lab2yf": y, "invgamma": ig}


def bilateral_tables(d=9, sigma_color=75.0, sigma_space=75.0):
    color = np.empty(768, np.float32); space = np.empty(512, np.float32)
    dy = np.empty(512, np.int32); dx = np.empty(512, np.int32)
    k = lib().orc_bilateral_tables(C.c_int(d), C.c_double(sigma_color), C.c_double(sigma_space),
                                   _p(color), _p(space), _p(dy), _p(dx))
    return color, space[:k].copy(), dy[:k].copy(), dx[:k].copy()


def gaussian_kernel_q8(k):
    q = np.zeros(31, np.int32)
    if lib().orc_gaussian_kernel_q8(C.c_int(k), _p(q)):
        raise ValueError("bad gaussian k=%r" % (k,))
    return q[:k].copy()


# --- enhancer stages ---------------------------------------------------------
def bgr2lab(bgr):
    bgr = _u8(bgr); out = np.empty_like(bgr)
    lib().orc_bgr2lab(_p(bgr), C.c_long(bgr.size // 3), _p(out))
    return out


def lab2bgr(lab):
    lab = _u8(lab); out = np.empty_like(lab)
    lib().orc_lab2bgr(_p(lab), C.c_long(lab.size // 3), _p(out))
    return out


def clahe_geometry(H, W, tiles=(8, 8)):
    tw = C.c_int(); th = C.c_int(); ew = C.c_int(); eh = C.c_int()
    lib().orc_clahe_geometry(C.c_int(H), C.c_int(W), C.c_int(tiles[0]), C.c_int(tiles[1]),
                             C.byref(tw), C.byref(th), C.byref(ew), C.byref(eh))
    return tw.value, th.value, ew.value, eh.value


def clahe(l, clip_limit=3.0, tiles=(8, 8), return_tables=False):
    """tiles = (tilesX, tilesY) like cv2.createCLAHE(tileGridSize)."""
    l = _u8(l); H, W = l.shape
    nt = tiles[0] * tiles[1]
    out = np.empty_like(l)
    hist = np.empty((nt, 256), np.int32); lut = np.empty((nt, 256), np.uint8)
    lib().orc_clahe(_p(l), C.c_int(H), C.c_int(W), C.c_double(clip_limit), C.c_int(tiles[0]), C.c_int(tiles[1]),
                    _p(out), _p(hist), _p(lut))
    return (out, hist, lut) if return_tables else out


def correct_lighting(bgr, clip_limit=3.0, tiles=(8, 8)):
    bgr = _u8(bgr); H, W, _ = bgr.shape; out = np.empty_like(bgr)
    lib().orc_correct_lighting(_p(bgr), C.c_int(H), C.c_int(W), C.c_double(clip_limit),
                               C.c_int(tiles[0]), C.c_int(tiles[1]), _p(out))
    return out


def bilateral(bgr, d=9, sigma_color=75.0, sigma_space=75.0, use_fma=True):
    bgr = _u8(bgr); H, W, _ = bgr.shape; out = np.empty_like(bgr)
    lib().orc_bilateral(_p(bgr), C.c_int(H), C.c_int(W), C.c_int(d), C.c_double(sigma_color),
                        C.c_double(sigma_space), C.c_int(int(use_fma)), _p(out))
    return out


def sharpen(img):
    img = _u8(img); H, W = img.shape[:2]; ch = 1 if img.ndim == 2 else img.shape[2]
    out = np.empty_like(img)
    lib().orc_sharpen(_p(img), C.c_int(H), C.c_int(W), C.c_int(ch), _p(out))
    return out


def normalize_lut(smin, smax, use_fma=True):
    lut = np.empty(256, np.uint8)
    lib().orc_normalize_lut(C.c_int(smin), C.c_int(smax), C.c_int(int(use_fma)), _p(lut))
    return lut


def normalize(img, use_fma=True, return_minmax=False):
    img = _u8(img); out = np.empty_like(img); mn = C.c_int(); mx = C.c_int()
    lib().orc_normalize(_p(img), C.c_long(img.size), C.c_int(int(use_fma)), _p(out), C.byref(mn), C.byref(mx))
    return (out, mn.value, mx.value) if return_minmax else out


def gray(bgr):
    H, W, _ = bgr.shape
    assert bgr.dtype == np.uint8 and bgr.strides[2] == 1 and bgr.strides[1] == 3
    out = np.empty((H, W), np.uint8)
    lib().orc_gray(_p(bgr), C.c_int(H), C.c_int(W), C.c_long(bgr.strides[0]), _p(out))
    return out


def gaussian(g, k=5, sigma=0.0):
    """cv2.GaussianBlur(g, (k, k), sigma) for u8 (sigma 0: OpenCV's default for the size)."""
    assert g.dtype == np.uint8 and g.ndim == 2 and (g.strides[1] == 1 or g.shape[1] == 1)
    H, W = g.shape; out = np.empty((H, W), np.uint8)
    if lib().orc_gaussian_sigma(_p(g), C.c_int(H), C.c_int(W), C.c_long(g.strides[0]), C.c_int(k), C.c_double(sigma), _p(out)):
        raise ValueError("bad gaussian k=%r" % (k,))
    return out


def gaussian_kernel_q8_sigma(k, sigma):
    q = np.zeros(31, np.int32)
    if lib().orc_gaussian_kernel_q8_sigma(C.c_int(k), C.c_double(sigma), _p(q)):
        raise ValueError("bad gaussian k=%r" % (k,))
    return q[:k].copy()


def dilate(plane, kw=5, kh=5, iterations=1):
    """cv2.dilate(plane, np.ones((kh, kw), np.uint8), iterations=iterations)"""
    g = _u8(plane); H, W = g.shape; out = np.empty_like(g)
    if lib().orc_dilate(_p(g), C.c_int(H), C.c_int(W), C.c_int(kw), C.c_int(kh), C.c_int(iterations), _p(out)):
        raise ValueError("bad dilate arguments")
    return out


def contour_mask(bgr):
    """The image part of board_detection.find_chessboard_corners (board_detection.py:9-14)."""
    bgr = _u8(bgr); H, W, _ = bgr.shape; out = np.empty((H, W), np.uint8)
    lib().orc_contour_mask(_p(bgr), C.c_int(H), C.c_int(W), _p(out))
    return out


def hist256(img):
    img = _u8(img); h = np.empty(256, np.int32)
    lib().orc_hist256(_p(img), C.c_long(img.size), _p(h))
    return h


def otsu_from_hist(h, n):
    h = np.ascontiguousarray(h, np.int32)
    return lib().orc_otsu_from_hist(_p(h), C.c_long(int(n)))


def threshold(img, T):
    img = _u8(img); out = np.empty_like(img)
    lib().orc_threshold(_p(img), C.c_long(img.size), C.c_int(int(T)), _p(out))
    return out


def prepare_analysis(bgr, return_all=False):
    bgr = _u8(bgr); H, W, _ = bgr.shape
    g = np.empty((H, W), np.uint8); bl = np.empty((H, W), np.uint8); b = np.empty((H, W), np.uint8)
    T = lib().orc_prepare_analysis(_p(bgr), C.c_int(H), C.c_int(W), _p(g), _p(bl), _p(b))
    return (g, b, T, bl) if return_all else (g, b)


def process_pipeline(bgr, use_fma=True):
    bgr = _u8(bgr); H, W, _ = bgr.shape; out = np.empty_like(bgr)
    lib().orc_process_pipeline(_p(bgr), C.c_int(H), C.c_int(W), C.c_int(int(use_fma)), _p(out))
    return out


# --- warp -----------------------------------------------------------------------
def get_perspective(src_pts, dst_pts):
    s = np.ascontiguousarray(np.asarray(src_pts, np.float32).reshape(4, 2))
    d = np.ascontiguousarray(np.asarray(dst_pts, np.float32).reshape(4, 2))
    M = np.empty(9, np.float64)
    if lib().orc_get_perspective(_p(s), _p(d), _p(M)):
        raise ValueError("degenerate quadrilateral")
    return M.reshape(3, 3)


def invert3(M):
    M = np.ascontiguousarray(M, np.float64); out = np.empty(9, np.float64)
    if lib().orc_invert3(_p(M), _p(out)):
        raise ValueError("singular matrix")
    return out.reshape(3, 3)


def warp(bgr, M, size):
    """cv2.warpPerspective(bgr, M, (size_w, size_h)) with INTER_LINEAR / BORDER_CONSTANT 0."""
    bgr = _u8(bgr); H, W, _ = bgr.shape
    sw, sh = (size, size) if np.isscalar(size) else size
    Mi = np.ascontiguousarray(invert3(M))
    out = np.empty((sh, sw, 3), np.uint8)
    lib().orc_warp(_p(bgr), C.c_int(H), C.c_int(W), _p(Mi), C.c_int(sh), C.c_int(sw), _p(out))
    return out


def warp_image(bgr, points, display_size=(1280, 720), margin=100):
    """board_detection.warp_image (board_detection.py:61-71)."""
    S = min(display_size) - margin
    M = get_perspective(points, [[0, 0], [S, 0], [0, S], [S, S]])
    return warp(bgr, M, S), M, S


# --- per-square --------------------------------------------------------------------
class PdStats(C.Structure):
    _fields_ = [("n", C.c_int64), ("sum", C.c_int64), ("sumsq", C.c_int64), ("sad", C.c_int64),
                ("center_sum", C.c_int64), ("center_cnt", C.c_int64),
                ("border_sum", C.c_int64), ("border_cnt", C.c_int64),
                ("ring_sum", C.c_int64 * 4), ("ring_cnt", C.c_int64 * 4)]


def square_preprocess(sq, k=5):
    assert sq.dtype == np.uint8
    h, w = sq.shape[:2]
    ch = 1 if sq.ndim == 2 else 3
    assert sq.strides[-1] == 1 and (ch == 1 or sq.strides[1] == 3)
    out = np.empty((h, w), np.uint8)
    if lib().orc_square_preprocess(_p(sq), C.c_int(h), C.c_int(w), C.c_int(ch), C.c_long(sq.strides[0]),
                                   C.c_int(k), _p(out)):
        raise ValueError("bad gaussian k")
    return out


def square_masks(h, w):
    m = np.empty((h, w), np.uint8)
    lib().orc_square_masks(C.c_int(h), C.c_int(w), _p(m))
    return m


def pd_square_stats(g, ref=None):
    g = _u8(g); h, w = g.shape
    st = PdStats()
    r = _u8(ref) if ref is not None else None
    lib().orc_pd_square_stats(_p(g), _p(r) if r is not None else None, C.c_int(h), C.c_int(w), C.byref(st))
    return {"n": st.n, "sum": st.sum, "sumsq": st.sumsq, "sad": st.sad,
            "center_sum": st.center_sum, "center_cnt": st.center_cnt,
            "border_sum": st.border_sum, "border_cnt": st.border_cnt,
            "ring_sum": list(st.ring_sum), "ring_cnt": list(st.ring_cnt)}


def cd_calibrate(g, initial_variance=100.0):
    g = _u8(g); m = np.empty(g.shape, np.float32); v = np.empty(g.shape, np.float32)
    lib().orc_cd_calibrate(_p(g), C.c_long(g.size), C.c_float(initial_variance), _p(m), _p(v))
    return m, v


def cd_update(g, mean, var, alpha=0.1):
    """In-place EMA update; alpha as the reference computes it: f32(alpha), f32(1 - alpha)."""
    g = _u8(g)
    assert mean.dtype == np.float32 and var.dtype == np.float32 and mean.flags.c_contiguous and var.flags.c_contiguous
    lib().orc_cd_update(_p(g), C.c_long(g.size), C.c_float(np.float32(alpha)), C.c_float(np.float32(1 - alpha)),
                        _p(mean), _p(var))


def cd_detect(g, mean, var, z_threshold=2.5):
    g = _u8(g); mean = np.ascontiguousarray(mean, np.float32); var = np.ascontiguousarray(var, np.float32)
    cnt = C.c_int64(); zmax = C.c_float()
    lib().orc_cd_detect(_p(g), C.c_long(g.size), _p(mean), _p(var), C.c_float(np.float32(z_threshold)),
                        C.byref(cnt), C.byref(zmax))
    return cnt.value, zmax.value


# --- S0 colour profile ("next" scope row) ----------------------------------------------
class ColorProfile(C.Structure):
    _fields_ = [("contrast", C.c_double), ("brightness", C.c_double),
                ("hue_shift", C.c_float), ("sat_scale", C.c_float), ("val_scale", C.c_float),
                ("radical_mode", C.c_int), ("target_hue", C.c_float), ("hue_window", C.c_float),
                ("simd_block", C.c_int)]


def color_profile_struct(profile, simd_block=32):
    """frame_enhancer.py:61-68 defaults; NumPy turns the python scalars into f32 (weak scalars)."""
    p = ColorProfile()
    p.contrast = float(profile.get("contrast", 1.0)); p.brightness = float(profile.get("brightness", 0))
    p.hue_shift = np.float32(profile.get("hue_shift", 0)); p.sat_scale = np.float32(profile.get("sat_scale", 1.0))
    p.val_scale = np.float32(profile.get("val_scale", 1.0)); p.radical_mode = int(bool(profile.get("radical_mode", 0)))
    p.target_hue = np.float32(profile.get("target_hue", 0)); p.hue_window = np.float32(profile.get("hue_window", 20))
    p.simd_block = int(simd_block)
    return p


def convert_scale_abs(img, alpha, beta):
    img = _u8(img); out = np.empty_like(img)
    lib().orc_convert_scale_abs(_p(img), C.c_long(img.size), C.c_double(alpha), C.c_double(beta), _p(out))
    return out


def bgr2hsv(bgr):
    bgr = _u8(bgr); out = np.empty_like(bgr)
    lib().orc_bgr2hsv(_p(bgr), C.c_long(bgr.size // 3), _p(out))
    return out


def hsv2bgr(hsv, simd_block=32):
    hsv = _u8(hsv); H, W, _ = hsv.shape; out = np.empty_like(hsv)
    lib().orc_hsv2bgr(_p(hsv), C.c_int(H), C.c_int(W), C.c_int(simd_block), _p(out))
    return out


def apply_color_profile(bgr, profile, simd_block=32):
    if not profile:
        return bgr
    bgr = _u8(bgr); H, W, _ = bgr.shape; out = np.empty_like(bgr)
    p = color_profile_struct(profile, simd_block)
    lib().orc_apply_color_profile(_p(bgr), C.c_int(H), C.c_int(W), C.byref(p), _p(out))
    return out


# --- Canny / grid refinement ("next" scope row) ---------------------------------------------
def canny(gray_u8, low=50, high=150):
    g = _u8(gray_u8); H, W = g.shape; out = np.empty_like(g)
    lib().orc_canny(_p(g), C.c_int(H), C.c_int(W), C.c_double(low), C.c_double(high), _p(out))
    return out


def rotate(img, code):
    """cv2.rotate(img, code) (0: 90 degrees clockwise, 1: 180, 2: 90 counter-clockwise) for (H,W) / (H,W,C) u8."""
    a = _u8(img); H, W = a.shape[:2]; ch = 1 if a.ndim == 2 else a.shape[2]
    shape = (H, W) if code == 1 else (W, H)
    out = np.empty(shape + (() if a.ndim == 2 else (ch,)), np.uint8)
    f = lib().orc_rotate; f.restype = C.c_int
    if f(_p(a), C.c_int(H), C.c_int(W), C.c_int(ch), C.c_int(int(code)), _p(out)) != 0:
        raise ValueError("rotate: bad code %r" % (code,))
    return out



def yuv_to_bgr(src, fmt):
    """cv2.cvtColor(src, COLOR_YUV2BGR_YUY2) for fmt 'yuy2' ((H, W, 2) u8) or COLOR_YUV2BGR_NV12 for 'nv12'
    ((H * 3 / 2, W) u8) -> (H, W, 3) BGR.  The camera ingest step either side of the path (play_lichess.py:16-18,45)."""
    a = _u8(src)
    code = {"yuy2": 1, "nv12": 2}[fmt]
    H, W = (a.shape[0], a.shape[1]) if code == 1 else (a.shape[0] * 2 // 3, a.shape[1])
    out = np.empty((H, W, 3), np.uint8)
    f = lib().orc_yuv_to_bgr; f.restype = C.c_int
    if f(_p(a), C.c_int(code), C.c_int(H), C.c_int(W), _p(out)) != 0:
        raise ValueError("yuv_to_bgr: bad shape %r for %s" % (a.shape, fmt))
    return out


def hough_circles(gray_u8, dp=1.2, min_dist=25, param1=100, param2=25, min_radius=0, max_radius=0, max_out=64,
                  return_info=False):
    """cv2.HoughCircles(gray, HOUGH_GRADIENT, ...) (piece_detector.py:232-241) -> (n,3) f32 (x, y, r) or None;
    with return_info also (support per circle, edge pixels, candidate centres)."""
    g = _u8(gray_u8); H, W = g.shape
    out = np.zeros((max_out, 4), np.float32); info = np.zeros(2, np.int32)
    f = lib().orc_hough_circles; f.restype = C.c_int
    n = f(_p(g), C.c_int(H), C.c_int(W), C.c_long(W), C.c_double(dp), C.c_double(min_dist), C.c_double(param1),
          C.c_double(param2), C.c_int(min_radius), C.c_int(max_radius), _p(out), C.c_int(max_out), _p(info))
    if n < 0:
        raise ValueError("hough_circles: bad arguments")
    n = min(n, max_out)
    circles = out[:n, :3].copy() if n else None
    if return_info:
        return circles, out[:n, 3].astype(np.int32), int(info[0]), int(info[1])
    return circles


def refine_grid(bgr, return_edges=False):
    """SmartGridExtractor.refine_grid (grid_extractor.py:66-121) -> (grid_x, grid_y) lists of 9 ints."""
    bgr = _u8(bgr); H, W, _ = bgr.shape
    gx = np.empty(9, np.int32); gy = np.empty(9, np.int32); e = np.empty((H, W), np.uint8)
    lib().orc_refine_grid(_p(bgr), C.c_int(H), C.c_int(W), _p(gx), _p(gy), _p(e))
    return (gx.tolist(), gy.tolist(), e) if return_edges else (gx.tolist(), gy.tolist())
