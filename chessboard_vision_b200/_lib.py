"""ctypes binding of libcvb200.so (include/cvb200.h).

This is the seam the reference fills with ``from src.cython.<mod> import
<Class>`` (frame_enhancer.py:8-21, change_detector.py:7-19).  There is no CPU
fallback: a missing library or GPU raises, it never aliases a Python class.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CVB200_LIB") or os.path.join(_HERE, "libcvb200.so")   # override: a differently built library


class CvbError(RuntimeError):
    pass


class ColorProfile(C.Structure):
    _fields_ = [("contrast", C.c_double), ("brightness", C.c_double),
                ("hue_shift", C.c_float), ("sat_scale", C.c_float), ("val_scale", C.c_float),
                ("radical_mode", C.c_int), ("target_hue", C.c_float), ("hue_window", C.c_float),
                ("simd_block", C.c_int)]


class EnhanceParams(C.Structure):
    _fields_ = [("clahe_clip_limit", C.c_double), ("tiles_x", C.c_int), ("tiles_y", C.c_int),
                ("bilateral_d", C.c_int), ("sigma_color", C.c_double), ("sigma_space", C.c_double),
                ("use_color_profile", C.c_int), ("profile", ColorProfile)]


class Rect(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32)]


class SquareStats(C.Structure):
    _fields_ = [("n", C.c_int32), ("has_ref", C.c_int32), ("sum", C.c_uint32), ("sad", C.c_uint32),
                ("sumsq", C.c_uint64),
                ("center_sum", C.c_uint32), ("center_cnt", C.c_uint32),
                ("border_sum", C.c_uint32), ("border_cnt", C.c_uint32),
                ("ring_sum", C.c_uint32 * 4), ("ring_cnt", C.c_uint32 * 4),
                ("cd_changed", C.c_int32), ("cd_zmax", C.c_float), ("cd_valid", C.c_int32),
                ("reserved", C.c_int32 * 11)]


assert C.sizeof(SquareStats) == 128


class SquareParams(C.Structure):
    _fields_ = [("ops", C.c_int), ("pd_blur", C.c_int), ("cd_blur", C.c_int),
                ("z_threshold", C.c_float), ("alpha", C.c_float), ("one_minus_alpha", C.c_float),
                ("initial_variance", C.c_float), ("min_variance", C.c_float)]


class PipelineParams(C.Structure):
    _fields_ = [("enhance", EnhanceParams), ("squares", SquareParams),
                ("warp_enhanced", C.c_int), ("board_size", C.c_int), ("rotate_180", C.c_int), ("reserved", C.c_int)]


HOUGH_MAX_CIRCLES, HOUGH_MAX_DIM, HOUGH_MAX_DIM_GLOBAL = 16, 128, 254
HOUGH_OK, HOUGH_SKIPPED = 0, 1


class HoughParams(C.Structure):
    _fields_ = [("dp", C.c_float), ("min_dist", C.c_float), ("param1", C.c_double), ("param2", C.c_double),
                ("min_radius_ratio", C.c_double), ("max_radius_ratio", C.c_double),
                ("min_radius", C.c_int), ("max_radius", C.c_int), ("min_dist_div", C.c_int), ("reserved", C.c_int)]


class HoughSquare(C.Structure):
    _fields_ = [("x", C.c_int32), ("y", C.c_int32), ("w", C.c_int32), ("h", C.c_int32),
                ("min_radius", C.c_int32), ("max_radius", C.c_int32), ("acc_rows", C.c_int32), ("acc_cols", C.c_int32),
                ("n_bins", C.c_int32), ("min_dist", C.c_float)]


class HoughResult(C.Structure):
    _fields_ = [("count", C.c_int32), ("n_edges", C.c_int32), ("n_centers", C.c_int32), ("status", C.c_int32),
                ("xyr", (C.c_float * 3) * HOUGH_MAX_CIRCLES), ("support", C.c_int32 * HOUGH_MAX_CIRCLES)]


assert C.sizeof(HoughParams) == 56 and C.sizeof(HoughSquare) == 40 and C.sizeof(HoughResult) == 272

OV_RECT, OV_CIRCLE, OV_STAMP = 0, 1, 2


class OverlayOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("x0", C.c_int32), ("y0", C.c_int32), ("x1", C.c_int32), ("y1", C.c_int32),
                ("color", C.c_uint8 * 4), ("alpha", C.c_float), ("beta", C.c_float), ("group", C.c_int32),
                ("aux_ofs", C.c_uint32)]


assert C.sizeof(OverlayOp) == 40

SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE = 1, 2, 4, 8, 16
PLANE_PD_REF, PLANE_CD_MEAN, PLANE_CD_VAR, PLANE_PD_CUR, PLANE_FLAGS = 0, 1, 2, 3, 4
FMT_BGR, FMT_YUY2, FMT_NV12 = 0, 1, 2
FORMATS = {"bgr": FMT_BGR, "yuy2": FMT_YUY2, "nv12": FMT_NV12}

# every symbol include/cvb200.h declares: (name, restype, argtypes)
_P, _I, _D, _SZ = C.c_void_p, C.c_int, C.c_double, C.c_size_t
SYMBOLS = [
    ("cvb_version", _I, []),
    ("cvb_device_count", _I, []),
    ("cvb_last_error", C.c_char_p, []),
    ("cvb_create", _I, [_I, C.POINTER(_P)]),
    ("cvb_destroy", None, [_P]),
    ("cvb_set_stream", _I, [_P, _P]),
    ("cvb_synchronize", _I, [_P]),
    ("cvb_launch_count", C.c_int64, [_P]),
    ("cvb_profile_enable", _I, [_P, _I]),
    ("cvb_profile_read", _I, [_P, _P, _P, _P, _I, C.POINTER(_I)]),
    ("cvb_malloc", _I, [_P, _SZ, C.POINTER(_P)]),
    ("cvb_free", _I, [_P, _P]),
    ("cvb_host_alloc", _I, [_SZ, C.POINTER(_P)]),
    ("cvb_host_free", _I, [_P]),
    ("cvb_memcpy_h2d", _I, [_P, _P, _P, _SZ]),
    ("cvb_memcpy_d2h", _I, [_P, _P, _P, _SZ]),
    ("cvb_memset", _I, [_P, _P, _I, _SZ]),
    ("cvb_event_create", _I, [C.POINTER(_P)]),
    ("cvb_event_destroy", _I, [_P]),
    ("cvb_event_record", _I, [_P, _P]),
    ("cvb_event_elapsed_ms", _I, [_P, _P, C.POINTER(C.c_float)]),
    ("cvb_enhance_params_default", None, [C.POINTER(EnhanceParams)]),
    ("cvb_get_tables", _I, [_P, _P, _P, _P, _P]),
    ("cvb_get_bilateral_tables", _I, [_D, _D, _P, _P]),
    ("cvb_gaussian_kernel_q8", _I, [_I, _P]),
    ("cvb_color_profile_default", None, [C.POINTER(ColorProfile)]),
    ("cvb_color_profile_dev", _I, [_P, _P, _I, _I, _I, C.POINTER(ColorProfile), _P]),
    ("cvb_bgr2lab_dev", _I, [_P, _P, _I, _I, _I, _P]),
    ("cvb_lab2bgr_dev", _I, [_P, _P, _I, _I, _I, _P]),
    ("cvb_clahe_dev", _I, [_P, _P, _I, _I, _I, _D, _I, _I, _P, _P, _P]),
    ("cvb_correct_lighting_dev", _I, [_P, _P, _I, _I, _I, _D, _I, _I, _P, _P, _P]),
    ("cvb_bilateral_dev", _I, [_P, _P, _I, _I, _I, _I, _D, _D, _P]),
    ("cvb_sharpen_dev", _I, [_P, _P, _I, _I, _I, _P]),
    ("cvb_normalize_dev", _I, [_P, _P, _I, _I, _I, _I, _P, _P]),
    ("cvb_gray_dev", _I, [_P, _P, _I, _I, _I, _P]),
    ("cvb_gaussian_dev", _I, [_P, _P, _I, _I, _I, _I, _P]),
    ("cvb_prepare_analysis_dev", _I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    ("cvb_process_pipeline_dev", _I, [_P, _P, _I, _I, _I, C.POINTER(EnhanceParams), _P]),
    ("cvb_enhance_dev", _I, [_P, _P, _I, _I, _I, C.POINTER(EnhanceParams), _P, _P, _P, _P]),
    ("cvb_enhance", _I, [_P, _P, _I, _I, _I, C.POINTER(EnhanceParams), _P, _P, _P, _P]),
    ("cvb_get_perspective_transform", _I, [_P, _P, _P]),
    ("cvb_warp_dev", _I, [_P, _P, _I, _I, _I, _P, _I, _I, _I, _P]),
    ("cvb_warp_rot180_dev", _I, [_P, _P, _I, _I, _I, _P, _I, _I, _I, _P]),
    ("cvb_rotate_dev", _I, [_P, _P, _I, _I, _I, _I, _I, _P]),
    ("cvb_overlay_dev", _I, [_P, _P, _I, _I, _I, _P, _I, _P, _SZ]),
    ("cvb_gaussian_sigma_dev", _I, [_P, _P, _I, _I, _I, _I, _D, _P]),
    ("cvb_dilate_dev", _I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    ("cvb_contour_mask_dev", _I, [_P, _P, _I, _I, _I, _P]),
    ("cvb_canny_dev", _I, [_P, _P, _I, _I, _I, _D, _D, _P]),
    ("cvb_projections_dev", _I, [_P, _P, _I, _I, _I, _P, _P]),
    ("cvb_state_create", _I, [_P, _I, _I, _I, C.POINTER(_P)]),
    ("cvb_state_destroy", None, [_P]),
    ("cvb_square_params_default", None, [C.POINTER(SquareParams)]),
    ("cvb_squares_dev", _I, [_P, _P, _I, _I, _I, _I, _P, _I, _P, _P, _I, C.POINTER(SquareParams), _P]),
    ("cvb_state_get", _I, [_P, _P, _I, _I, _P]),
    ("cvb_state_set", _I, [_P, _P, _I, _I, _P]),
    ("cvb_state_reset", _I, [_P, _P, _I]),
    ("cvb_hough_params_default", None, [C.POINTER(HoughParams)]),
    ("cvb_hough_geometry", _I, [_P, _I, C.POINTER(HoughParams), _P]),
    ("cvb_hough_dev", _I, [_P, _P, _I, _I, _I, _P, _I, _P, C.POINTER(HoughParams), _P]),
    ("cvb_hough_state", _I, [_P, _P, _I, _I, _P, _I, _P, C.POINTER(HoughParams), _P]),
    ("cvb_pipeline_params_default", None, [C.POINTER(PipelineParams)]),
    ("cvb_pipeline_dev", _I, [_P, _P, _I, _I, _I, C.POINTER(PipelineParams), _P, _I, _P, _I, _P, _P, _I,
                              _P, _P, _P, _P, _P, _P]),
    ("cvb_set_chunk_frames", _I, [_P, _I]),
    ("cvb_pipeline", _I, [_P, _P, _I, _I, _I, C.POINTER(PipelineParams), _P, _I, _P, _I, _P, _P, _I, _P, _P]),
    ("cvb_debug_bounds_violations", C.c_longlong, [_P, C.POINTER(_I)]),
    ("cvb_frame_bytes", _SZ, [_I, _I, _I]),
    ("cvb_cvt_to_bgr_dev", _I, [_P, _P, _I, _I, _I, _I, _P]),
    ("cvb_pipeline_fmt", _I, [_P, _P, _I, _I, _I, _I, C.POINTER(PipelineParams), _P, _I, _P, _I, _P, _P, _I, _P, _P]),
    ("cvb_pipeline_submit", _I, [_P, _P, _I, _I, _I, _I, C.POINTER(PipelineParams), _P, _I, _P, _I, _P, _P, _I, _P, _P,
                                 C.POINTER(C.c_uint64)]),
    ("cvb_pipeline_wait", _I, [_P, C.c_uint64]),
]

_lib = None


def load():
    """Load libcvb200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = LIB_PATH      # CVB200_LIB=<...>/libcvb200_dbg.so selects the debug build (index checks, make debug)
    if not os.path.exists(path):
        raise ImportError(
            "libcvb200.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). chessboard_vision_b200 has no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)   # AttributeError if the .so is stale
        fn.restype = res
        fn.argtypes = args
    if lib.cvb_version() != 100:
        raise ImportError("libcvb200.so version %d does not match this binding (100)" % lib.cvb_version())
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().cvb_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(msg)
        raise CvbError("libcvb200 error %d: %s" % (rc, msg))
