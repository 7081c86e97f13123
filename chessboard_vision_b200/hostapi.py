"""Host-side glue shared by the drop-in modules (chessboard_vision_b200/dropin).

Pure Python / numpy: turning the reference's ``{(file, rank): ndarray view}``
square dictionaries (grid_extractor.py:32,56) into one board image plus
rectangles for the batched kernel, turning the kernel's integer sums back into
the floats the reference computes with numpy, and finding the reference's own
module for the names that are outside the hot path.
"""
import importlib.util
import os
import sys

import numpy as np

_DROPIN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")


def load_reference_module(name):
    """Import the reference's `<name>.py` from the first sys.path / cwd entry that is not
    the drop-in directory.  Used only for out-of-scope helpers (GUI drawing, corner
    finding, Hough-based logic, colour-profile step); returns None when absent."""
    alias = "_cvb_ref_" + name
    if alias in sys.modules:
        return sys.modules[alias]
    for entry in list(sys.path) + [os.getcwd()]:
        d = os.path.abspath(entry or os.getcwd())
        if d == _DROPIN_DIR:
            continue
        f = os.path.join(d, name + ".py")
        if os.path.isfile(f):
            spec = importlib.util.spec_from_file_location(alias, f)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[alias] = mod
            try:
                spec.loader.exec_module(mod)
            except Exception:
                del sys.modules[alias]
                raise
            return mod
    return None


def pack_squares(squares):
    """-> (board u8 (BH,BW[,3]) C-contiguous, rects [(x,y,w,h)], keys) for a dict of squares.

    Fast path: every square is a view into one C-contiguous parent image (what split_board
    returns) -> the parent is the board, rectangles come from the view offsets, nothing is
    copied.  Otherwise the squares are packed into an atlas, eight per row."""
    keys = list(squares.keys())
    if not keys:
        return None, [], []
    arrs = [np.asarray(squares[k]) for k in keys]
    ch = 1 if arrs[0].ndim == 2 else arrs[0].shape[2]
    for a in arrs:
        if a.dtype != np.uint8 or a.ndim not in (2, 3) or (1 if a.ndim == 2 else a.shape[2]) != ch or ch not in (1, 3):
            raise ValueError("squares must be uint8 HxW or HxWx3 arrays with one common channel count")
    base = arrs[0].base
    if base is not None and isinstance(base, np.ndarray) and base.flags.c_contiguous and base.dtype == np.uint8 \
            and base.ndim == arrs[0].ndim and all(a.base is base for a in arrs):
        b0 = base.__array_interface__["data"][0]
        rs, px = base.strides[0], base.strides[1]
        rects, ok = [], True
        for a in arrs:
            if a.strides[:2] != base.strides[:2] or (a.ndim == 3 and a.strides[2] != 1):
                ok = False
                break
            off = a.__array_interface__["data"][0] - b0
            y, rem = divmod(off, rs)
            x = rem // px
            rects.append((int(x), int(y), int(a.shape[1]), int(a.shape[0])))
        if ok:
            return base, rects, keys
    # atlas: shelves of eight squares
    rects, x, y, shelf_h, width = [], 0, 0, 0, 0
    for i, a in enumerate(arrs):
        if i % 8 == 0 and i:
            y += shelf_h; x = 0; shelf_h = 0
        rects.append((x, y, a.shape[1], a.shape[0]))
        x += a.shape[1]; shelf_h = max(shelf_h, a.shape[0]); width = max(width, x)
    height = y + shelf_h
    atlas = np.zeros((height, width) + ((3,) if ch == 3 else ()), np.uint8)
    for a, (rx, ry, w, h) in zip(arrs, rects):
        atlas[ry:ry + h, rx:rx + w] = a
    return atlas, rects, keys


# ---- integer sums -> the reference's float quantities -------------------------------------------
def mean_abs_diff(st):
    """np.mean(cv2.absdiff(cur, ref))  (piece_detector.py:88-91): exact integer sum / count in f64."""
    return float(st["sad"]) / float(st["n"])


def std_from_moments(st):
    """np.std(gray) (piece_detector.py:305) from the exact integer moments."""
    n, s, q = int(st["n"]), int(st["sum"]), int(st["sumsq"])
    return float(np.sqrt(max(n * q - s * s, 0)) / n)


def std_below(st, thr):
    """np.std(gray) < thr decided in integers for integral thr: n*sum(x^2) - sum(x)^2 < thr^2 * n^2."""
    n, s, q = int(st["n"]), int(st["sum"]), int(st["sumsq"])
    if float(thr).is_integer():
        return n * q - s * s < int(thr) * int(thr) * n * n
    return std_from_moments(st) < thr


def center_vs_border(st):
    """_detect_center_vs_border (piece_detector.py:177-207) -> (diff, center_mean, border_mean)."""
    with np.errstate(invalid="ignore", divide="ignore"):
        cm = np.float64(st["center_sum"]) / np.float64(st["center_cnt"])
        bm = np.float64(st["border_sum"]) / np.float64(st["border_cnt"])
    return abs(cm - bm), cm, bm


def radial_symmetry(st):
    """_analyze_radial_symmetry (piece_detector.py:141-175): np.var of the ring means / 500, capped at 1."""
    means = [np.float64(s) / np.float64(c) for s, c in zip(st["ring_sum"], st["ring_cnt"]) if c > 0]
    if len(means) < 2:
        return 0.0
    return min(1.0, np.var(means) / 500)
