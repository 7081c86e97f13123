"""Thin Python host layer over the C ABI (include/cvb200.h).

``Engine`` owns one ``cvb_handle`` (one GPU, one stream).  Methods take and
return numpy arrays (host convenience: upload, launch, download) or
``DevArray`` device buffers (nothing copied).  Batches are leading-axis stacks
of frames: (n, H, W, 3) or a single (H, W, 3) frame.

Mirrors, stage by stage, what the reference's ImageEnhancer / warp_image /
ChangeDetector / PieceDetector compute through cv2 + numpy
(frame_enhancer.py:101-181, board_detection.py:61-71,
change_detector.py:36-167, piece_detector.py:82-207).
"""
import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import (ColorProfile, EnhanceParams, HoughParams, HoughResult, HoughSquare, PipelineParams, Rect, SquareParams,
                   SquareStats, check, HOUGH_MAX_CIRCLES,
                   SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE)

STATS_DTYPE = np.dtype([("n", "<i4"), ("has_ref", "<i4"), ("sum", "<u4"), ("sad", "<u4"), ("sumsq", "<u8"),
                        ("center_sum", "<u4"), ("center_cnt", "<u4"), ("border_sum", "<u4"), ("border_cnt", "<u4"),
                        ("ring_sum", "<u4", (4,)), ("ring_cnt", "<u4", (4,)),
                        ("cd_changed", "<i4"), ("cd_zmax", "<f4"), ("cd_valid", "<i4"), ("reserved", "<i4", (11,))])
assert STATS_DTYPE.itemsize == C.sizeof(SquareStats)
HOUGH_DTYPE = np.dtype([("count", "<i4"), ("n_edges", "<i4"), ("n_centers", "<i4"), ("status", "<i4"),
                        ("xyr", "<f4", (HOUGH_MAX_CIRCLES, 3)), ("support", "<i4", (HOUGH_MAX_CIRCLES,))])
HOUGH_SQUARE_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("w", "<i4"), ("h", "<i4"), ("min_radius", "<i4"),
                               ("max_radius", "<i4"), ("acc_rows", "<i4"), ("acc_cols", "<i4"), ("n_bins", "<i4"),
                               ("min_dist", "<f4")])
assert HOUGH_DTYPE.itemsize == C.sizeof(HoughResult) and HOUGH_SQUARE_DTYPE.itemsize == C.sizeof(HoughSquare)


class DevArray:
    """A device buffer with a shape and dtype (owned by the engine's GPU)."""

    def __init__(self, engine, shape, dtype=np.uint8):
        self.engine = engine
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        p = C.c_void_p()
        check(engine.lib.cvb_malloc(engine.h, self.nbytes, C.byref(p)))
        self.ptr = p.value

    def set(self, arr):
        arr = np.ascontiguousarray(arr, self.dtype)
        assert arr.nbytes == self.nbytes, (arr.shape, self.shape)
        check(self.engine.lib.cvb_memcpy_h2d(self.engine.h, self.ptr, arr.ctypes.data, self.nbytes))
        self.engine.synchronize()   # pageable source must outlive the copy
        return self

    def get(self):
        out = np.empty(self.shape, self.dtype)
        check(self.engine.lib.cvb_memcpy_d2h(self.engine.h, out.ctypes.data, self.ptr, self.nbytes))
        self.engine.synchronize()
        return out

    def free(self):
        if self.ptr:
            self.engine.lib.cvb_free(self.engine.h, self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            if self.ptr and self.engine.h:
                self.free()
        except Exception:
            pass


def _as_batch(a, channels):
    """(H,W[,C]) or (n,H,W[,C]) -> (n,H,W,C) view info."""
    if isinstance(a, DevArray):
        shape = a.shape
    else:
        shape = a.shape
    nd = 3 if channels > 1 else 2
    if len(shape) == nd:
        single = True
        n = 1; H, W = shape[0], shape[1]
    elif len(shape) == nd + 1:
        single = False
        n, H, W = shape[0], shape[1], shape[2]
    else:
        raise ValueError("expected a %d-channel image or batch, got shape %r" % (channels, shape))
    if channels > 1 and shape[-1] != channels:
        raise ValueError("expected %d channels, got shape %r" % (channels, shape))
    return single, n, H, W


class State:
    """Per-stream device state: PieceDetector references, ChangeDetector mean/variance planes."""

    def __init__(self, engine, n_streams, BH, BW):
        self.engine, self.n_streams, self.BH, self.BW = engine, n_streams, BH, BW
        p = C.c_void_p()
        check(engine.lib.cvb_state_create(engine.h, n_streams, BH, BW, C.byref(p)))
        self.ptr = p.value

    def get(self, stream, plane):
        dt = np.float32 if plane in (_lib.PLANE_CD_MEAN, _lib.PLANE_CD_VAR) else np.uint8
        out = np.empty((self.BH, self.BW), dt)
        check(self.engine.lib.cvb_state_get(self.engine.h, self.ptr, stream, plane, out.ctypes.data))
        return out

    def set(self, stream, plane, arr):
        dt = np.float32 if plane in (_lib.PLANE_CD_MEAN, _lib.PLANE_CD_VAR) else np.uint8
        arr = np.ascontiguousarray(arr, dt)
        assert arr.shape == (self.BH, self.BW)
        check(self.engine.lib.cvb_state_set(self.engine.h, self.ptr, stream, plane, arr.ctypes.data))

    def reset(self, stream=-1):
        check(self.engine.lib.cvb_state_reset(self.engine.h, self.ptr, stream))

    def free(self):
        if self.ptr:
            self.engine.lib.cvb_state_destroy(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            if self.engine.h:
                self.free()
        except Exception:
            pass


def grid_rects(board_size=620, grid_lines_x=None, grid_lines_y=None):
    """Rectangles (x, y, w, h) in the order GridExtractor / SmartGridExtractor
    insert squares (r-major, rank 8 first) and their (file, rank) keys
    (grid_extractor.py:33-56, 140-161)."""
    rects, keys = [], []
    if grid_lines_x is None or grid_lines_y is None:
        if np.isscalar(board_size):
            rows = cols = int(board_size)
        else:
            rows, cols = board_size
        sh, sw = rows // 8, cols // 8
        for r in range(8):
            for c in range(8):
                rects.append((c * sw, r * sh, sw, sh)); keys.append((c, 7 - r))
    else:
        for r in range(8):
            for c in range(8):
                x0, x1 = int(grid_lines_x[c]), int(grid_lines_x[c + 1])
                y0, y1 = int(grid_lines_y[r]), int(grid_lines_y[r + 1])
                if x0 >= x1 or y0 >= y1:
                    continue
                rects.append((x0, y0, x1 - x0, y1 - y0)); keys.append((c, 7 - r))
    return rects, keys


def _rect_array(rects):
    arr = (Rect * len(rects))()
    for i, (x, y, w, h) in enumerate(rects):
        arr[i].x, arr[i].y, arr[i].w, arr[i].h = int(x), int(y), int(w), int(h)
    return arr


class _LockedLib:
    """The C library behind one lock per engine.  A cvb_handle is not re-entrant (its workspaces and staging caches
    are unsynchronised) and ctypes releases the GIL during a call, so two Python threads sharing an engine -- the
    reference application has a second thread (lichess_session.py:36) -- must not be inside the library at once."""

    def __init__(self, lib, lock):
        self._lib, self._lock, self._cache = lib, lock, {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw, lock = getattr(self._lib, name), self._lock

            def fn(*args):
                with lock:
                    return raw(*args)
            self._cache[name] = fn
        return fn


class Engine:
    def __init__(self, device=0):
        self.lock = threading.RLock()
        self.lib = _LockedLib(_lib.load(), self.lock)
        self.h = None
        h = C.c_void_p()
        check(self.lib.cvb_create(int(device), C.byref(h)))
        self.h = h.value
        self.device = int(device)

    def close(self):
        if self.h:
            self.lib.cvb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- plumbing ---------------------------------------------------------------
    def synchronize(self):
        check(self.lib.cvb_synchronize(self.h))

    def set_stream(self, cuda_stream_ptr):
        check(self.lib.cvb_set_stream(self.h, cuda_stream_ptr))

    def launch_count(self):
        return int(self.lib.cvb_launch_count(self.h))

    def set_chunk_frames(self, frames):
        """Frames per chunk of the host-buffer pipeline (copy/compute overlap granularity)."""
        check(self.lib.cvb_set_chunk_frames(self.h, int(frames)))

    def profile(self, on=True):
        check(self.lib.cvb_profile_enable(self.h, int(bool(on))))

    def profile_read(self):
        """-> {kernel name: (total_ms, launches)} since profile(True)."""
        names = C.create_string_buffer(32 * 64)
        ms = (C.c_float * 64)(); cnt = (C.c_int * 64)(); n = C.c_int()
        check(self.lib.cvb_profile_read(self.h, names, ms, cnt, 64, C.byref(n)))
        return {names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode(): (ms[i], cnt[i]) for i in range(n.value)}

    def pinned(self, shape, dtype=np.uint8):
        """A page-locked host array (freed with the process)."""
        nbytes = int(np.prod(shape, dtype=np.int64)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        check(self.lib.cvb_host_alloc(nbytes, C.byref(p)))
        buf = (C.c_uint8 * nbytes).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def empty(self, shape, dtype=np.uint8):
        return DevArray(self, shape, dtype)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        return DevArray(self, arr.shape, arr.dtype).set(arr)

    def _in(self, a, dtype=np.uint8):
        if isinstance(a, DevArray):
            return a, False
        a = np.asarray(a)
        if a.dtype != dtype:
            raise ValueError("expected %s data, got %s" % (np.dtype(dtype), a.dtype))
        return self.upload(a), True

    def _out(self, dev, want_host, temps=()):
        res = dev.get() if want_host else dev
        if want_host:
            dev.free()
        for t in temps:
            t.free()
        return res

    def event(self):
        p = C.c_void_p()
        check(self.lib.cvb_event_create(C.byref(p)))
        return p.value

    def record(self, ev):
        check(self.lib.cvb_event_record(self.h, ev))

    def elapsed_ms(self, ev0, ev1):
        ms = C.c_float()
        check(self.lib.cvb_event_elapsed_ms(ev0, ev1, C.byref(ms)))
        return ms.value

    # -- parameters ----------------------------------------------------------------
    def color_profile_params(self, profile, simd_block=32):
        """dict with the keys of color_profile.json -> cvb_color_profile.  NumPy applies the python scalars to
        float32 arrays as float32 (frame_enhancer.py:79-89), hence the f32 fields."""
        p = ColorProfile()
        self.lib.cvb_color_profile_default(C.byref(p))
        p.contrast = float(profile.get("contrast", 1.0)); p.brightness = float(profile.get("brightness", 0))
        p.hue_shift = np.float32(profile.get("hue_shift", 0)); p.sat_scale = np.float32(profile.get("sat_scale", 1.0))
        p.val_scale = np.float32(profile.get("val_scale", 1.0))
        p.radical_mode = int(bool(profile.get("radical_mode", 0)))
        p.target_hue = np.float32(profile.get("target_hue", 0)); p.hue_window = np.float32(profile.get("hue_window", 20))
        p.simd_block = int(simd_block)
        return p

    def apply_color_profile(self, img, profile, simd_block=32):
        """ImageEnhancer.apply_color_profile (frame_enhancer.py:56-99); identity for an empty profile."""
        if not profile:
            return img
        return self._stage_p(self.lib.cvb_color_profile_dev, img, self.color_profile_params(profile, simd_block))

    def enhance_params(self, clip=3.0, tiles=(8, 8), d=9, sigma_color=75.0, sigma_space=75.0, profile=None):
        p = EnhanceParams()
        self.lib.cvb_enhance_params_default(C.byref(p))
        p.clahe_clip_limit = float(clip); p.tiles_x, p.tiles_y = int(tiles[0]), int(tiles[1])
        p.bilateral_d = int(d); p.sigma_color = float(sigma_color); p.sigma_space = float(sigma_space)
        if profile:
            p.use_color_profile = 1
            p.profile = self.color_profile_params(profile)
        return p

    # -- simple 1-in 1-out stages -----------------------------------------------------
    def _stage(self, fn, img, in_ch, out_ch, *extra):
        single, n, H, W = _as_batch(img, in_ch)
        src, tmp = self._in(img)
        shape = (n, H, W) + ((out_ch,) if out_ch > 1 else ())
        dst = self.empty(shape if not single else shape[1:])
        check(fn(self.h, src.ptr, n, H, W, *extra, dst.ptr))
        return self._out(dst, tmp, [src] if tmp else [])

    def bgr2lab(self, img):
        return self._stage(self.lib.cvb_bgr2lab_dev, img, 3, 3)

    def lab2bgr(self, img):
        return self._stage(self.lib.cvb_lab2bgr_dev, img, 3, 3)

    def bilateral(self, img, d=9, sigma_color=75.0, sigma_space=75.0):
        return self._stage(self.lib.cvb_bilateral_dev, img, 3, 3, int(d), float(sigma_color), float(sigma_space))

    def sharpen(self, img):
        return self._stage(self.lib.cvb_sharpen_dev, img, 3, 3)

    def gray(self, img):
        return self._stage(self.lib.cvb_gray_dev, img, 3, 1)

    def gaussian(self, plane, k=5, sigma=0.0):
        """cv2.GaussianBlur(plane, (k, k), sigma) for u8 planes."""
        if sigma and sigma > 0:
            return self._stage(self.lib.cvb_gaussian_sigma_dev, plane, 1, 1, int(k), float(sigma))
        return self._stage(self.lib.cvb_gaussian_dev, plane, 1, 1, int(k))

    def dilate(self, plane, kw=5, kh=5, iterations=1):
        """cv2.dilate(plane, np.ones((kh, kw), np.uint8), iterations=iterations)"""
        return self._stage(self.lib.cvb_dilate_dev, plane, 1, 1, int(kw), int(kh), int(iterations))

    def contour_mask(self, img):
        """The image part of board_detection.find_chessboard_corners (board_detection.py:9-14):
        gray -> GaussianBlur(7x7, 1) -> Canny(30, 100) -> dilate(5x5, 3 iterations)."""
        return self._stage(self.lib.cvb_contour_mask_dev, img, 3, 1)

    def normalize(self, img, return_minmax=False):
        a = np.asarray(img) if not isinstance(img, DevArray) else img
        C_ = 1 if len(a.shape) == 2 else a.shape[-1]
        # treat any (..., H, W, C) / (H, W) input as one frame of H*W*C bytes per leading index
        if len(a.shape) == 4:
            n, H, W = a.shape[0], a.shape[1], a.shape[2]
        elif len(a.shape) == 3 or len(a.shape) == 2:
            n, H, W = 1, a.shape[0], a.shape[1]
        else:
            raise ValueError("bad shape %r" % (a.shape,))
        src, tmp = self._in(img)
        dst = self.empty(a.shape)
        mm = self.empty((n, 2), np.int32)
        check(self.lib.cvb_normalize_dev(self.h, src.ptr, n, H, W, C_, dst.ptr, mm.ptr))
        res = self._out(dst, tmp, [src] if tmp else [])
        if return_minmax:
            m = mm.get(); mm.free()
            return res, m
        mm.free()
        return res

    def clahe(self, plane, clip=3.0, tiles=(8, 8), return_tables=False):
        single, n, H, W = _as_batch(plane, 1)
        src, tmp = self._in(plane)
        dst = self.empty((H, W) if single else (n, H, W))
        nt = tiles[0] * tiles[1]
        hist = self.empty((n, nt, 256), np.int32); lut = self.empty((n, nt, 256), np.uint8)
        check(self.lib.cvb_clahe_dev(self.h, src.ptr, n, H, W, float(clip), int(tiles[0]), int(tiles[1]),
                                     dst.ptr, hist.ptr, lut.ptr))
        res = self._out(dst, tmp, [src] if tmp else [])
        if return_tables:
            hh, ll = hist.get(), lut.get()
            hist.free(); lut.free()
            return res, (hh[0] if single else hh), (ll[0] if single else ll)
        hist.free(); lut.free()
        return res

    def correct_lighting(self, img, clip=3.0, tiles=(8, 8), return_tables=False):
        single, n, H, W = _as_batch(img, 3)
        src, tmp = self._in(img)
        dst = self.empty((H, W, 3) if single else (n, H, W, 3))
        nt = tiles[0] * tiles[1]
        hist = self.empty((n, nt, 256), np.int32); lut = self.empty((n, nt, 256), np.uint8)
        check(self.lib.cvb_correct_lighting_dev(self.h, src.ptr, n, H, W, float(clip), int(tiles[0]), int(tiles[1]),
                                                dst.ptr, hist.ptr, lut.ptr))
        res = self._out(dst, tmp, [src] if tmp else [])
        if return_tables:
            hh, ll = hist.get(), lut.get()
            hist.free(); lut.free()
            return res, (hh[0] if single else hh), (ll[0] if single else ll)
        hist.free(); lut.free()
        return res

    def prepare_analysis(self, img, return_all=False):
        """-> (gray, binary) like ImageEnhancer.prepare_analysis; return_all adds (otsu_t, blurred, hist)."""
        single, n, H, W = _as_batch(img, 3)
        src, tmp = self._in(img)
        shp = (H, W) if single else (n, H, W)
        g, b, bl = self.empty(shp), self.empty(shp), self.empty(shp)
        t = self.empty((n,), np.int32); hist = self.empty((n, 256), np.int32)
        check(self.lib.cvb_prepare_analysis_dev(self.h, src.ptr, n, H, W, g.ptr, b.ptr, bl.ptr, t.ptr, hist.ptr))
        if tmp:
            res = (g.get(), b.get(), t.get(), bl.get(), hist.get())
            for x in (g, b, bl, t, hist, src):
                x.free()
            if single:
                res = (res[0], res[1], int(res[2][0]), res[3], res[4][0])
            return res if return_all else res[:2]
        return (g, b, t, bl, hist) if return_all else (g, b)

    def process_pipeline(self, img, params=None):
        p = params or self.enhance_params()
        return self._stage_p(self.lib.cvb_process_pipeline_dev, img, p)

    def _stage_p(self, fn, img, p):
        single, n, H, W = _as_batch(img, 3)
        src, tmp = self._in(img)
        dst = self.empty((H, W, 3) if single else (n, H, W, 3))
        check(fn(self.h, src.ptr, n, H, W, C.byref(p), dst.ptr))
        return self._out(dst, tmp, [src] if tmp else [])

    def enhance(self, img, params=None):
        """process_pipeline + prepare_analysis through the HOST-buffer C entry point.
        -> (enhanced, gray, binary, otsu_t)"""
        p = params or self.enhance_params()
        img = np.ascontiguousarray(img, np.uint8)
        single, n, H, W = _as_batch(img, 3)
        enh = np.empty_like(img)
        shp = (H, W) if single else (n, H, W)
        g = np.empty(shp, np.uint8); b = np.empty(shp, np.uint8); t = np.empty(n, np.int32)
        check(self.lib.cvb_enhance(self.h, img.ctypes.data, n, H, W, C.byref(p), enh.ctypes.data, g.ctypes.data,
                                   b.ctypes.data, t.ctypes.data))
        return enh, g, b, (int(t[0]) if single else t)

    def enhance_dev(self, src, enhanced, gray=None, binary=None, otsu_t=None, params=None):
        p = params or self.enhance_params()
        n, H, W = src.shape[0], src.shape[1], src.shape[2]
        check(self.lib.cvb_enhance_dev(self.h, src.ptr, n, H, W, C.byref(p), enhanced.ptr if enhanced else None,
                                       gray.ptr if gray else None, binary.ptr if binary else None,
                                       otsu_t.ptr if otsu_t else None))

    # -- warp -------------------------------------------------------------------------------
    def get_perspective_transform(self, src_pts, dst_pts):
        s = np.ascontiguousarray(np.asarray(src_pts, np.float32).reshape(4, 2))
        d = np.ascontiguousarray(np.asarray(dst_pts, np.float32).reshape(4, 2))
        M = np.empty(9, np.float64)
        check(self.lib.cvb_get_perspective_transform(s.ctypes.data, d.ctypes.data, M.ctypes.data))
        return M.reshape(3, 3)

    def warp(self, img, M, size, rotate_180=False):
        """cv2.warpPerspective (board_detection.py:69); rotate_180 adds game_session.py's cv2.rotate(ROTATE_180)."""
        single, n, H, W = _as_batch(img, 3)
        M = np.ascontiguousarray(M, np.float64)
        n_mats = 1 if M.ndim == 2 else M.shape[0]
        ow, oh = (size, size) if np.isscalar(size) else size
        src, tmp = self._in(img)
        dst = self.empty((oh, ow, 3) if single else (n, oh, ow, 3))
        fn = self.lib.cvb_warp_rot180_dev if rotate_180 else self.lib.cvb_warp_dev
        check(fn(self.h, src.ptr, n, H, W, M.ctypes.data, n_mats, int(oh), int(ow), dst.ptr))
        return self._out(dst, tmp, [src] if tmp else [])

    def warp_dev(self, src, M, size, dst, rotate_180=False):
        """Device buffers: (n,H,W,3) -> (n,size,size,3); nothing is copied or synchronised."""
        M = np.ascontiguousarray(M, np.float64)
        n, H, W = src.shape[0], src.shape[1], src.shape[2]
        fn = self.lib.cvb_warp_rot180_dev if rotate_180 else self.lib.cvb_warp_dev
        check(fn(self.h, src.ptr, n, H, W, M.ctypes.data, 1 if M.ndim == 2 else M.shape[0], int(size), int(size), dst.ptr))

    def squares_dev(self, boards, rects, params, state, stream0, stats):
        """Device buffers: boards (n,BH,BW,3) -> stats (n, n_sq) on the device (may be None)."""
        n, BH, BW = boards.shape[0], boards.shape[1], boards.shape[2]
        ra = rects if not isinstance(rects, (list, tuple)) else _rect_array(rects)
        check(self.lib.cvb_squares_dev(self.h, boards.ptr, n, BH, BW, boards.shape[3] if len(boards.shape) == 4 else 1,
                                       C.cast(ra, C.c_void_p), len(ra), None, state.ptr if state is not None else None,
                                       int(stream0), C.byref(params), stats.ptr if stats is not None else None))

    def rotate(self, img, code):
        """cv2.rotate(img, code) for (H,W) / (H,W,3) / (n,H,W,3) u8; code 0: 90 cw, 1: 180, 2: 90 ccw."""
        shape = tuple(img.shape)
        if len(shape) == 2:
            n, H, W, ch, lead = 1, shape[0], shape[1], 1, ()
        elif len(shape) == 3 and shape[2] == 3:
            n, H, W, ch, lead = 1, shape[0], shape[1], 3, ()
        elif len(shape) == 4 and shape[3] == 3:
            n, H, W, ch, lead = shape[0], shape[1], shape[2], 3, (shape[0],)
        else:
            raise ValueError("rotate: expected (H,W), (H,W,3) or (n,H,W,3), got %r" % (shape,))
        oh, ow = (H, W) if int(code) == 1 else (W, H)
        src, tmp = self._in(img)
        dst = self.empty(lead + ((oh, ow, 3) if ch == 3 else (oh, ow)))
        check(self.lib.cvb_rotate_dev(self.h, src.ptr, n, H, W, ch, int(code), dst.ptr))
        return self._out(dst, tmp, [src] if tmp else [])

    def overlay(self, img, ops, n_ops, masks=b""):
        """Apply a display list (ctypes array of _lib.OverlayOp, first n_ops used) to (H,W,3) / (n,H,W,3) u8 images:
        cvb_overlay_dev.  A DevArray is drawn in place and returned; a NumPy image is uploaded, drawn and returned
        as a new array."""
        single, n, H, W = _as_batch(img, 3)
        dev, tmp = self._in(img)
        masks = bytes(masks)
        mbuf = (C.c_uint8 * max(1, len(masks))).from_buffer_copy(masks or b"\0")
        check(self.lib.cvb_overlay_dev(self.h, dev.ptr, n, H, W, C.cast(ops, C.c_void_p) if n_ops else None, int(n_ops),
                                       C.cast(mbuf, C.c_void_p), len(masks)))
        return self._out(dev, tmp)

    # -- Canny / grid refinement (calibration time) ----------------------------------------------
    def canny(self, gray, low=50, high=150):
        return self._canny(gray, low, high)

    def _canny(self, gray, low, high):
        single, n, H, W = _as_batch(gray, 1)
        src, tmp = self._in(gray)
        dst = self.empty((H, W) if single else (n, H, W))
        check(self.lib.cvb_canny_dev(self.h, src.ptr, n, H, W, float(low), float(high), dst.ptr))
        return self._out(dst, tmp, [src] if tmp else [])

    def projections(self, plane):
        """-> (row_sums, col_sums) like np.sum(plane, axis=1), np.sum(plane, axis=0) (as uint64)."""
        single, n, H, W = _as_batch(plane, 1)
        src, tmp = self._in(plane)
        rows, cols = self.empty((n, H), np.uint32), self.empty((n, W), np.uint32)
        check(self.lib.cvb_projections_dev(self.h, src.ptr, n, H, W, rows.ptr, cols.ptr))
        r, c = rows.get().astype(np.uint64), cols.get().astype(np.uint64)
        rows.free(); cols.free()
        if tmp:
            src.free()
        return (r[0], c[0]) if single else (r, c)

    def refine_grid(self, img_warped):
        """SmartGridExtractor.refine_grid (grid_extractor.py:66-121): gray -> Canny(50,150) -> edge projections on
        the GPU, then the seven window arg-max positions per axis (nine numbers each) on the host."""
        img = np.ascontiguousarray(img_warped, np.uint8)
        h, w = img.shape[:2]
        g = self.upload(img)
        gray = self.gray(g)
        edges = self._canny(gray, 50, 150)
        row_proj, col_proj = self.projections(edges)
        for t in (g, gray, edges):
            t.free()

        def lines(proj, length):
            step = length / 8.0
            out = [0]
            for i in range(1, 8):
                c, r = int(i * step), int(step * 0.3)
                window = proj[max(0, c - r):min(length, c + r)]
                out.append(max(0, c - r) + np.argmax(window) if len(window) > 0 else c)
            out.append(length)
            return out
        return lines(col_proj, w), lines(row_proj, h)

    # -- squares ---------------------------------------------------------------------------------
    def square_params(self, ops=SQ_PD_STATS, pd_blur=5, cd_blur=5, z_threshold=2.5, alpha=0.1,
                      initial_variance=100.0, min_variance=10.0):
        p = SquareParams()
        p.ops = int(ops); p.pd_blur = int(pd_blur); p.cd_blur = int(cd_blur)
        # NumPy semantics of change_detector.py:82-89 / 126-129: python floats are weak scalars -> f32
        p.z_threshold = np.float32(z_threshold)
        p.alpha = np.float32(alpha); p.one_minus_alpha = np.float32(1 - alpha)
        p.initial_variance = np.float32(initial_variance); p.min_variance = np.float32(min_variance)
        return p

    def new_state(self, n_streams, BH, BW):
        return State(self, n_streams, BH, BW)

    def squares(self, boards, rects, params, state=None, stream0=0, select=None, want_stats=True):
        """boards: (BH,BW[,3]) / (n,BH,BW[,3]) u8 (numpy or DevArray) -> structured array (n, n_sq)."""
        shape = boards.shape
        if len(shape) == 2:
            n, BH, BW, ch = 1, shape[0], shape[1], 1
        elif len(shape) == 3 and shape[-1] == 3:
            n, BH, BW, ch = 1, shape[0], shape[1], 3
        elif len(shape) == 3:
            n, BH, BW, ch = shape[0], shape[1], shape[2], 1
        elif len(shape) == 4:
            n, BH, BW, ch = shape[0], shape[1], shape[2], shape[3]
        else:
            raise ValueError("bad boards shape %r" % (shape,))
        src, tmp = self._in(boards)
        ra = _rect_array(rects)
        sel = None
        if select is not None:
            sel = np.ascontiguousarray(select, np.uint8)
            assert sel.size == len(rects)
        stats = self.empty((n, len(rects)), STATS_DTYPE) if want_stats else None
        check(self.lib.cvb_squares_dev(self.h, src.ptr, n, BH, BW, ch, C.cast(ra, C.c_void_p), len(rects),
                                       sel.ctypes.data if sel is not None else None,
                                       state.ptr if state is not None else None, int(stream0), C.byref(params),
                                       stats.ptr if stats is not None else None))
        out = None
        if stats is not None:
            out = stats.get(); stats.free()
        else:
            self.synchronize()
        if tmp:
            src.free()
        return out

    # -- Hough circles per square (piece_detector.py:210-270) ------------------------------------------
    def hough_params(self, dp=1.2, param1=100, param2=25, min_radius_ratio=0.20, max_radius_ratio=0.55,
                     min_dist_div=3, min_radius=None, max_radius=None, min_dist=None):
        """The arguments of the reference's cv2.HoughCircles call.  By default the radii and minDist
        follow each square (`int(min_dim * ratio)`, `min_dim // 3`); explicit `min_radius`,
        `max_radius`, `min_dist` give cv2.HoughCircles' own arguments instead."""
        p = HoughParams()
        self.lib.cvb_hough_params_default(C.byref(p))
        p.dp = float(dp); p.param1 = float(param1); p.param2 = float(param2)
        p.min_radius_ratio = float(min_radius_ratio); p.max_radius_ratio = float(max_radius_ratio)
        p.min_dist_div = int(min_dist_div)
        if min_radius is not None:
            p.min_radius_ratio = -1.0; p.min_radius = int(min_radius)
        if max_radius is not None:
            p.max_radius_ratio = -1.0; p.max_radius = int(max_radius)
        if min_dist is not None:
            p.min_dist_div = 0; p.min_dist = float(min_dist)
        return p

    def hough_geometry(self, rects, params):
        ra = _rect_array(rects)
        out = np.zeros(len(rects), HOUGH_SQUARE_DTYPE)
        check(self.lib.cvb_hough_geometry(C.cast(ra, C.c_void_p), len(rects), C.byref(params), out.ctypes.data))
        return out

    @staticmethod
    def _hough_select(select, n, n_sq):
        if select is None:
            return None
        sel = np.ascontiguousarray(select, np.uint8)
        if sel.size == n_sq and n > 1:
            sel = np.ascontiguousarray(np.broadcast_to(sel.reshape(1, n_sq), (n, n_sq)))
        assert sel.size == n * n_sq
        return sel

    def hough(self, planes, rects, params=None, select=None):
        """planes: (PH,PW) / (n,PH,PW) u8 gray (numpy or DevArray) -> structured array (n, n_sq) of HOUGH_DTYPE:
        cv2.HoughCircles of every selected square, circles in OpenCV's order."""
        params = params or self.hough_params()
        shape = planes.shape
        n, PH, PW = (1, shape[0], shape[1]) if len(shape) == 2 else shape
        src, tmp = self._in(planes)
        ra = _rect_array(rects)
        sel = self._hough_select(select, n, len(rects))
        res = self.empty((n, len(rects)), HOUGH_DTYPE)
        check(self.lib.cvb_hough_dev(self.h, src.ptr, n, PH, PW, C.cast(ra, C.c_void_p), len(rects),
                                     sel.ctypes.data if sel is not None else None, C.byref(params), res.ptr))
        out = res.get(); res.free()
        if tmp:
            src.free()
        return out

    def hough_state(self, state, rects, params=None, stream0=0, n=1, select=None):
        """The same on the state's last gray+blur squares (plane PLANE_PD_CUR) of n stream slots."""
        params = params or self.hough_params()
        ra = _rect_array(rects)
        sel = self._hough_select(select, n, len(rects))
        out = np.zeros((n, len(rects)), HOUGH_DTYPE)
        check(self.lib.cvb_hough_state(self.h, state.ptr, int(stream0), int(n), C.cast(ra, C.c_void_p), len(rects),
                                       sel.ctypes.data if sel is not None else None, C.byref(params), out.ctypes.data))
        return out

    @staticmethod
    def hough_circles(rec):
        """One HOUGH_DTYPE record -> what cv2.HoughCircles returns: (1, k, 3) f32 or None."""
        k = min(int(rec["count"]), HOUGH_MAX_CIRCLES)
        return rec["xyr"][:k].reshape(1, k, 3).copy() if k else None

    # -- whole path ----------------------------------------------------------------------------------
    def pipeline_params(self, enhance=None, squares=None, warp_enhanced=True, board_size=620, rotate_180=False):
        p = PipelineParams()
        self.lib.cvb_pipeline_params_default(C.byref(p))
        if enhance is not None:
            p.enhance = enhance
        if squares is not None:
            p.squares = squares
        p.warp_enhanced = int(bool(warp_enhanced)); p.board_size = int(board_size)
        p.rotate_180 = int(bool(rotate_180))
        return p

    # -- camera ingest (play_lichess.py:16-18,45) -------------------------------
    @staticmethod
    def _frame_geometry(frames, fmt):
        """-> (n, H, W) of host / device frames in `fmt`: 'bgr' (n,H,W,3), 'yuy2' (n,H,W,2), 'nv12' (n,H*3/2,W);
        a single frame may omit the leading axis."""
        shp = tuple(frames.shape)
        if fmt == "nv12":
            if len(shp) == 2:
                shp = (1,) + shp
            if len(shp) != 3 or shp[1] % 3 or (shp[1] * 2 // 3) % 2 or shp[2] % 2:
                raise ValueError("NV12 frames are (H * 3 / 2, W) with even H and W, got %r" % (frames.shape,))
            return shp[0], shp[1] * 2 // 3, shp[2]
        ch = {"bgr": 3, "yuy2": 2}[fmt]
        if len(shp) == 3:
            shp = (1,) + shp
        if len(shp) != 4 or shp[3] != ch or (fmt == "yuy2" and shp[2] % 2):
            raise ValueError("%s frames are (H, W, %d)%s, got %r" % (fmt, ch, " with even W" if fmt == "yuy2" else "", frames.shape))
        return shp[0], shp[1], shp[2]

    def cvt_to_bgr(self, frames, fmt):
        """cv2.cvtColor(frame, COLOR_YUV2BGR_YUY2 / COLOR_YUV2BGR_NV12) on the device -> BGR (n,H,W,3) or (H,W,3)."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, H, W = self._frame_geometry(frames, fmt)
        single = frames.ndim == (2 if fmt == "nv12" else 3)
        src = self.upload(frames)
        dst = self.empty((H, W, 3) if single else (n, H, W, 3))
        check(self.lib.cvb_cvt_to_bgr_dev(self.h, src.ptr, _lib.FORMATS[fmt], n, H, W, dst.ptr))
        out = dst.get()
        src.free(); dst.free()
        return out

    def cvt_to_bgr_dev(self, src, fmt, n, H, W, dst):
        """Device buffers: n native frames -> n BGR frames."""
        check(self.lib.cvb_cvt_to_bgr_dev(self.h, src.ptr, _lib.FORMATS[fmt], int(n), int(H), int(W), dst.ptr))

    def pipeline(self, frames, M, rects, params, state=None, stream0=0, select=None, fmt="bgr"):
        """HOST frames -> (otsu_t[n], stats[n, n_sq]); H2D and D2H happen inside the call.  `fmt`: 'bgr' (n,H,W,3),
        or the camera's native 'yuy2' (n,H,W,2) / 'nv12' (n,H*3/2,W), converted to BGR on the device exactly as
        cv2.cvtColor does, so that 2 / 1.5 instead of 3 bytes per pixel cross PCIe."""
        frames = np.ascontiguousarray(frames, np.uint8)
        n, H, W = self._frame_geometry(frames, fmt)
        M = np.ascontiguousarray(M, np.float64)
        n_mats = 1 if M.ndim == 2 else M.shape[0]
        ra = _rect_array(rects)
        stats = np.empty((n, len(rects)), STATS_DTYPE); t = np.empty(n, np.int32)
        sel = np.ascontiguousarray(select, np.uint8) if select is not None else None
        check(self.lib.cvb_pipeline_fmt(self.h, frames.ctypes.data, _lib.FORMATS[fmt], n, H, W, C.byref(params),
                                        M.ctypes.data, n_mats, C.cast(ra, C.c_void_p), len(rects),
                                        sel.ctypes.data if sel is not None else None,
                                        state.ptr if state is not None else None, int(stream0), t.ctypes.data,
                                        stats.ctypes.data))
        return t, stats

    def pipeline_submit(self, frames, M, rects, params, state=None, stream0=0, select=None, fmt="bgr"):
        """`pipeline` without the wait: returns a ticket; `pipeline_wait(ticket)` -> (otsu_t, stats).  Submit batch k+1
        before waiting for batch k and its first host->device copy runs beside the kernels of batch k.  `frames` must
        be a C-contiguous uint8 array (pinned: Engine.pinned) that is left alone until the wait."""
        if not (isinstance(frames, np.ndarray) and frames.dtype == np.uint8 and frames.flags.c_contiguous):
            raise ValueError("pipeline_submit: frames must be a C-contiguous uint8 array (it is read after the call returns)")
        n, H, W = self._frame_geometry(frames, fmt)
        M = np.ascontiguousarray(M, np.float64)
        n_mats = 1 if M.ndim == 2 else M.shape[0]
        ra = _rect_array(rects)
        stats = np.empty((n, len(rects)), STATS_DTYPE); t = np.empty(n, np.int32)
        sel = np.ascontiguousarray(select, np.uint8) if select is not None else None
        ticket = C.c_uint64(0)
        check(self.lib.cvb_pipeline_submit(self.h, frames.ctypes.data, _lib.FORMATS[fmt], n, H, W, C.byref(params),
                                           M.ctypes.data, n_mats, C.cast(ra, C.c_void_p), len(rects),
                                           sel.ctypes.data if sel is not None else None,
                                           state.ptr if state is not None else None, int(stream0), t.ctypes.data,
                                           stats.ctypes.data, C.byref(ticket)))
        return (ticket.value, t, stats, frames, M, ra, sel)     # keeps every buffer the enqueued work touches alive

    def pipeline_wait(self, ticket):
        """-> (otsu_t, stats) of that submission, complete"""
        check(self.lib.cvb_pipeline_wait(self.h, ticket[0]))
        return ticket[1], ticket[2]

    def pipeline_dev(self, src, M, rects, params, state=None, stream0=0, select=None, enhanced=None, gray=None,
                     binary=None, otsu_t=None, warped=None, stats=None):
        n, H, W = src.shape[0], src.shape[1], src.shape[2]
        M = np.ascontiguousarray(M, np.float64)
        n_mats = 1 if M.ndim == 2 else M.shape[0]
        ra = rects if not isinstance(rects, (list, tuple)) else _rect_array(rects)
        sel = np.ascontiguousarray(select, np.uint8) if select is not None else None
        g = lambda x: x.ptr if x is not None else None
        check(self.lib.cvb_pipeline_dev(self.h, src.ptr, n, H, W, C.byref(params), M.ctypes.data, n_mats,
                                        C.cast(ra, C.c_void_p), len(ra), sel.ctypes.data if sel is not None else None,
                                        state.ptr if state is not None else None, int(stream0),
                                        g(enhanced), g(gray), g(binary), g(otsu_t), g(warped), g(stats)))


_default = {}


def default_engine(device=0):
    """One shared engine per device for the drop-in module classes."""
    e = _default.get(device)
    if e is None or e.h is None:
        e = _default[device] = Engine(device)
    return e
