"""A numpy/oracle-backed stand-in for chessboard_vision_b200.engine.Engine, used ONLY by the CPU
tests to exercise the host logic of the drop-in modules (dict handling, state windows, gating)
where no GPU exists.  Every numeric result comes from the CPU oracle."""
import numpy as np

import oracle as O
from chessboard_vision_b200 import _lib
from chessboard_vision_b200.engine import STATS_DTYPE, HOUGH_DTYPE, Engine
from chessboard_vision_b200._lib import SquareParams, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE


class FakeState:
    def __init__(self, n, BH, BW):
        self.n_streams, self.BH, self.BW = n, BH, BW
        self.planes = {_lib.PLANE_PD_REF: np.zeros((n, BH, BW), np.uint8), _lib.PLANE_PD_CUR: np.zeros((n, BH, BW), np.uint8),
                       _lib.PLANE_FLAGS: np.zeros((n, BH, BW), np.uint8),
                       _lib.PLANE_CD_MEAN: np.zeros((n, BH, BW), np.float32), _lib.PLANE_CD_VAR: np.zeros((n, BH, BW), np.float32)}
        self.ptr = 1

    def get(self, stream, plane):
        return self.planes[plane][stream].copy()

    def set(self, stream, plane, arr):
        self.planes[plane][stream] = arr

    def reset(self, stream=-1):
        self.planes[_lib.PLANE_FLAGS][...] = 0

    def free(self):
        self.ptr = None


class FakeEngine:
    h = 1
    launches = 0

    def overlay(self, img, ops, n_ops, masks=b""):
        from oracle import overlay as OV
        return OV.apply_display_list(img, ops, n_ops, masks)

    def enhance_params(self, clip=3.0, tiles=(8, 8), d=9, sigma_color=75.0, sigma_space=75.0, profile=None):
        return dict(clip=clip, tiles=tiles, profile=profile)

    def apply_color_profile(self, img, profile, simd_block=32):
        return O.apply_color_profile(img, profile, simd_block)

    def clahe(self, plane, clip=3.0, tiles=(8, 8), return_tables=False):
        return O.clahe(plane, clip, tiles, return_tables)

    def correct_lighting(self, img, clip=3.0, tiles=(8, 8)):
        return O.correct_lighting(img, clip, tiles)

    def bilateral(self, img, d=9, sc=75.0, ss=75.0):
        return O.bilateral(img, d, sc, ss, True)

    def sharpen(self, img):
        return O.sharpen(img)

    def normalize(self, img):
        return O.normalize(img, True)

    def gray(self, img):
        return O.gray(img)

    def gaussian(self, g, k=5):
        return O.gaussian(g, k)

    def prepare_analysis(self, img):
        return O.prepare_analysis(img)

    def process_pipeline(self, img, params=None):
        if params and params.get("profile"):
            img = O.apply_color_profile(img, params["profile"])
        return O.process_pipeline(img, True)

    def enhance(self, img, params=None):
        enh = self.process_pipeline(img, params)
        g, b, t, _ = O.prepare_analysis(enh, True)
        return enh, g, b, t

    def refine_grid(self, img):
        return O.refine_grid(img)

    def get_perspective_transform(self, s, d):
        return O.get_perspective(s, d)

    def gaussian(self, g, k=5, sigma=0.0):
        return O.gaussian(np.ascontiguousarray(g), k, sigma)

    def dilate(self, plane, kw=5, kh=5, iterations=1):
        return O.dilate(plane, kw, kh, iterations)

    def contour_mask(self, img):
        FakeEngine.launches += 1
        return O.contour_mask(img)

    def warp(self, img, M, size):
        return O.warp(img, M, size)

    def square_params(self, ops=SQ_PD_STATS, pd_blur=5, cd_blur=5, z_threshold=2.5, alpha=0.1, initial_variance=100.0,
                      min_variance=10.0):
        p = SquareParams()
        p.ops, p.pd_blur, p.cd_blur = int(ops), int(pd_blur), int(cd_blur)
        p.z_threshold = np.float32(z_threshold); p.alpha = np.float32(alpha); p.one_minus_alpha = np.float32(1 - alpha)
        p.initial_variance = np.float32(initial_variance); p.min_variance = np.float32(min_variance)
        return p

    def new_state(self, n, BH, BW):
        return FakeState(n, BH, BW)

    def squares(self, boards, rects, p, state=None, stream0=0, select=None, want_stats=True):
        FakeEngine.launches += 1
        b = np.asarray(boards)
        if b.ndim == 2 or (b.ndim == 3 and b.shape[-1] == 3):
            b = b[None]
        out = np.zeros((b.shape[0], len(rects)), STATS_DTYPE)
        for f in range(b.shape[0]):
            s = stream0 + f
            for i, (x, y, w, h) in enumerate(rects):
                sel = True if select is None else bool(select[i])
                sq = b[f, y:y + h, x:x + w]
                st = out[f, i]
                st["n"] = w * h
                fl = state.planes[_lib.PLANE_FLAGS][s, y, x] if state is not None else 0
                if p.ops & (SQ_PD_STATS | SQ_PD_SET_REF):
                    g = O.square_preprocess(sq, p.pd_blur)
                    if p.ops & SQ_PD_STATS:
                        ref = state.planes[_lib.PLANE_PD_REF][s, y:y + h, x:x + w] if (fl & 1) else None
                        o = O.pd_square_stats(g, ref)
                        st["has_ref"] = 1 if ref is not None else 0
                        st["sum"], st["sumsq"], st["sad"] = o["sum"], o["sumsq"], max(o["sad"], 0)
                        st["center_sum"], st["center_cnt"] = o["center_sum"], o["center_cnt"]
                        st["border_sum"], st["border_cnt"] = o["border_sum"], o["border_cnt"]
                        st["ring_sum"], st["ring_cnt"] = o["ring_sum"], o["ring_cnt"]
                    if state is not None:
                        state.planes[_lib.PLANE_PD_CUR][s, y:y + h, x:x + w] = g
                        if (p.ops & SQ_PD_SET_REF) and sel:
                            state.planes[_lib.PLANE_PD_REF][s, y:y + h, x:x + w] = g
                            state.planes[_lib.PLANE_FLAGS][s, y:y + h, x:x + w] |= 1
                if (p.ops & (SQ_CD_CALIBRATE | SQ_CD_DETECT | SQ_CD_UPDATE)) and sel and state is not None:
                    g = O.square_preprocess(sq, p.cd_blur)
                    M = state.planes[_lib.PLANE_CD_MEAN][s, y:y + h, x:x + w]
                    V = state.planes[_lib.PLANE_CD_VAR][s, y:y + h, x:x + w]
                    has_cd = (fl & 6) == 6
                    if p.ops & SQ_CD_CALIBRATE:
                        m, v = O.cd_calibrate(g, p.initial_variance)
                        M[...] = m; V[...] = v
                        state.planes[_lib.PLANE_FLAGS][s, y:y + h, x:x + w] |= 6
                        has_cd = True
                    if has_cd:
                        m, v = np.array(M, np.float32), np.array(V, np.float32)
                        if p.ops & SQ_CD_DETECT:
                            cnt, zmax = O.cd_detect(g, m, v, p.z_threshold)
                            st["cd_changed"], st["cd_zmax"], st["cd_valid"] = cnt, zmax, 1
                        if p.ops & SQ_CD_UPDATE:
                            # the exact f32 (alpha, 1-alpha) pair the kernel receives
                            m2, v2 = np.array(M, np.float32), np.array(V, np.float32)
                            O.lib().orc_cd_update(g.ctypes.data_as(O.oracle.C.c_void_p), O.oracle.C.c_long(g.size),
                                                  O.oracle.C.c_float(p.alpha), O.oracle.C.c_float(p.one_minus_alpha),
                                                  m2.ctypes.data_as(O.oracle.C.c_void_p), v2.ctypes.data_as(O.oracle.C.c_void_p))
                            M[...] = m2; V[...] = v2
        return out if want_stats else None

    # -- Hough circles per square: geometry from the library's host code, circles from the oracle --
    @property
    def lib(self):
        return _lib.load()

    hough_params = Engine.hough_params
    hough_geometry = Engine.hough_geometry
    hough_circles = staticmethod(Engine.hough_circles)

    def hough(self, planes, rects, params=None, select=None):
        FakeEngine.launches += 1
        params = params or self.hough_params()
        pl = np.asarray(planes)
        if pl.ndim == 2:
            pl = pl[None]
        geo = self.hough_geometry(rects, params)
        sel = None if select is None else np.broadcast_to(np.asarray(select, np.uint8).reshape(-1, len(rects)), (pl.shape[0], len(rects)))
        out = np.zeros((pl.shape[0], len(rects)), HOUGH_DTYPE)
        for f in range(pl.shape[0]):
            for i, (x, y, w, h) in enumerate(rects):
                if w > _lib.HOUGH_MAX_DIM_GLOBAL or h > _lib.HOUGH_MAX_DIM_GLOBAL:
                    raise ValueError("square larger than %d" % _lib.HOUGH_MAX_DIM_GLOBAL)
                if sel is not None and not sel[f, i]:
                    out[f, i]["status"] = _lib.HOUGH_SKIPPED
                    continue
                c, sup, ne, nc = O.hough_circles(np.ascontiguousarray(pl[f, y:y + h, x:x + w]), dp=max(1.0, params.dp),
                                                 min_dist=float(geo[i]["min_dist"]), param1=params.param1, param2=params.param2,
                                                 min_radius=int(geo[i]["min_radius"]), max_radius=int(geo[i]["max_radius"]),
                                                 max_out=4096, return_info=True)
                k = 0 if c is None else len(c)
                out[f, i]["count"], out[f, i]["n_edges"], out[f, i]["n_centers"] = k, ne, nc
                k = min(k, _lib.HOUGH_MAX_CIRCLES)
                if k:
                    out[f, i]["xyr"][:k] = c[:k]
                    out[f, i]["support"][:k] = sup[:k]
        return out

    def hough_state(self, state, rects, params=None, stream0=0, n=1, select=None):
        return self.hough(state.planes[_lib.PLANE_PD_CUR][stream0:stream0 + n], rects, params, select)

    # -- whole path on host frames (cvb_pipeline): enhance -> warp -> squares, composed from the oracle --
    def pipeline_params(self, enhance=None, squares=None, warp_enhanced=True, board_size=620, rotate_180=False):
        return dict(enhance=enhance, squares=squares, warp_enhanced=warp_enhanced, board_size=board_size, rotate_180=rotate_180)

    def pipeline(self, frames, M, rects, params, state=None, stream0=0, select=None):
        frames = np.asarray(frames)
        n, S = frames.shape[0], params["board_size"]
        t = np.zeros(n, np.int32)
        boards = np.zeros((n, S, S, 3), np.uint8)
        for i in range(n):
            enh, _, _, t[i] = self.enhance(frames[i])
            b = O.warp(enh if params["warp_enhanced"] else frames[i], M, S)
            boards[i] = b[::-1, ::-1] if params["rotate_180"] else b
        return t, self.squares(boards, rects, params["squares"], state, stream0, select)
