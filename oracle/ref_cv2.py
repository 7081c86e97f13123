"""CPU ORACLE, second form (test infrastructure, NOT product code): the hot path
restated as the exact OpenCV / NumPy call sequence the reference executes, for
hosts where cv2 is importable.  It is what `bench.py --impl reference` and the
`cpu_baseline` leg time (the reference itself is Python source under
/root/reference and cannot travel to the GPU box; its arithmetic is these
library calls), and what tests/test_oracle_vs_cv2.py pins the C restatement
(oracle/cvb_oracle.c) against.

Each function names the reference lines whose call sequence it follows.
"""
import numpy as np

try:
    import cv2
    HAVE_CV2 = True
except Exception:  # pragma: no cover
    cv2 = None
    HAVE_CV2 = False

_SHARPEN = np.array([[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]])   # frame_enhancer.py:40-42
_clahe = None


def _get_clahe(clip=3.0, tiles=(8, 8)):
    global _clahe
    if _clahe is None or _clahe[0] != (clip, tiles):
        _clahe = ((clip, tiles), cv2.createCLAHE(clipLimit=clip, tileGridSize=tiles))   # frame_enhancer.py:36
    return _clahe[1]


def correct_lighting(frame, clip=3.0, tiles=(8, 8)):
    """frame_enhancer.py:101-120."""
    l, a, b = cv2.split(cv2.cvtColor(frame, cv2.COLOR_BGR2LAB))
    return cv2.cvtColor(cv2.merge((_get_clahe(clip, tiles).apply(l), a, b)), cv2.COLOR_LAB2BGR)


def process_pipeline(frame):
    """frame_enhancer.py:161-181 with no colour profile loaded."""
    x = correct_lighting(frame)
    x = cv2.bilateralFilter(x, d=9, sigmaColor=75, sigmaSpace=75)            # :131
    x = cv2.filter2D(x, -1, _SHARPEN)                                        # :138
    return cv2.normalize(x, None, alpha=0, beta=255, norm_type=cv2.NORM_MINMAX)   # :146


def prepare_analysis(frame):
    """frame_enhancer.py:148-159."""
    gray = cv2.cvtColor(frame, cv2.COLOR_BGR2GRAY)
    t, binary = cv2.threshold(cv2.GaussianBlur(gray, (5, 5), 0), 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
    return gray, binary, t


def warp_image(img, points, display_size=(1280, 720), margin=100):
    """board_detection.py:61-71."""
    s = min(display_size) - margin
    m = cv2.getPerspectiveTransform(np.float32(points), np.float32([[0, 0], [s, 0], [0, s], [s, s]]))
    return cv2.warpPerspective(img, m, (s, s)), m, s


def split_board(img, grid_x=None, grid_y=None):
    """grid_extractor.py:33-56 (linear) / 140-161 (calibrated lines): dict of views."""
    out = {}
    if grid_x is None or grid_y is None:
        sh, sw = img.shape[0] // 8, img.shape[1] // 8
        for r in range(8):
            for c in range(8):
                out[(c, 7 - r)] = img[r * sh:(r + 1) * sh, c * sw:(c + 1) * sw]
    else:
        for r in range(8):
            for c in range(8):
                if grid_x[c] < grid_x[c + 1] and grid_y[r] < grid_y[r + 1]:
                    out[(c, 7 - r)] = img[grid_y[r]:grid_y[r + 1], grid_x[c]:grid_x[c + 1]]
    return out


def preprocess_square(sq, k=5):
    """piece_detector.py:124-135 / change_detector.py:49-56."""
    g = cv2.cvtColor(sq, cv2.COLOR_BGR2GRAY) if sq.ndim == 3 else sq
    return cv2.GaussianBlur(g, (k | 1, k | 1), 0)


def pd_statistics(gray, ref=None):
    """The numeric part of PieceDetector on one preprocessed square:
    _has_changed (piece_detector.py:82-93), std gate (:305), centre-vs-border (:177-207),
    ring means / variance score (:141-175).  Same NumPy expressions, masks rebuilt per call."""
    h, w = gray.shape
    out = {}
    out["mean_diff"] = None if ref is None else float(np.mean(cv2.absdiff(gray, ref)))
    out["std"] = float(np.std(gray))
    cy, cx = h // 2, w // 2
    yy, xx = np.ogrid[:h, :w]
    cm = ((xx - cx) ** 2 + (yy - cy) ** 2) <= (min(h, w) // 4) ** 2
    cs = min(h, w) // 4
    bm = np.zeros((h, w), dtype=bool)
    bm[:cs, :cs] = True; bm[:cs, -cs:] = True; bm[-cs:, :cs] = True; bm[-cs:, -cs:] = True
    out["center_mean"] = float(np.mean(gray[cm])); out["border_mean"] = float(np.mean(gray[bm]))
    out["center_border_diff"] = abs(out["center_mean"] - out["border_mean"])
    ring_means = []
    for r in [min(h, w) * f for f in (0.15, 0.25, 0.35, 0.45)]:
        dist = np.sqrt((xx - cx) ** 2 + (yy - cy) ** 2)
        m = (dist >= r - 5) & (dist <= r + 5)
        if np.sum(m) > 0:
            ring_means.append(np.mean(gray[m]))
    out["ring_means"] = [float(v) for v in ring_means]
    out["symmetry"] = 0.0 if len(ring_means) < 2 else float(min(1.0, np.var(ring_means) / 500))
    return out


def detect_circle_unified(gray, min_radius_ratio=0.20, max_radius_ratio=0.55, param1=100, param2=25):
    """PieceDetector._detect_circle_unified (piece_detector.py:210-270): the cv2.HoughCircles call and
    the choice of the circle nearest the square centre -> (found, centre, radius, kind), circles."""
    h, w = gray.shape
    md = min(h, w)
    circles = cv2.HoughCircles(gray, cv2.HOUGH_GRADIENT, dp=1.2, minDist=md // 3, param1=param1, param2=param2,
                               minRadius=int(md * min_radius_ratio), maxRadius=int(md * max_radius_ratio))
    if circles is not None and len(circles[0]) > 0:
        best, best_d = None, float("inf")
        for c in circles[0]:
            d = np.sqrt((c[0] - w // 2) ** 2 + (c[1] - h // 2) ** 2)
            if d < md * 0.3 and d < best_d:
                best, best_d = c, d
        if best is not None:
            r = int(best[2])
            return (True, (int(best[0]), int(best[1])), r, "tower_top" if r < md * 0.20 else "hough"), circles
    return (False, None, None, None), circles


def cd_detect(gray_u8, mean, var, z_threshold=2.5):
    """change_detector.py:121-137,159 -> (changed_pixels, pct_changed, z_max)."""
    g = gray_u8.astype(np.float32)
    z = np.abs(g - mean) / np.sqrt(var)
    changed = int(np.count_nonzero(z > z_threshold))
    return changed, (changed / g.size) * 100, float(np.max(z))


def cd_update(gray_u8, mean, var, alpha=0.1):
    """change_detector.py:77-92 -> (new_mean, new_var)."""
    g = gray_u8.astype(np.float32)
    nm = (1 - alpha) * mean + alpha * g
    d = g - nm
    nv = np.maximum((1 - alpha) * var + alpha * (d ** 2), 10.0)
    return nm, nv


def full_frame(frame, points, cd_state=None, pd_ref=None, grid=None, keep=None):
    """One frame of the benchmark composition (BASELINE.json configs 1-3): enhance, analysis,
    warp, 64 squares, PieceDetector statistics, ChangeDetector detect + update.
    `keep` (a dict) receives the enhanced frame, the Otsu mask and the preprocessed squares (parity checks)."""
    enh = process_pipeline(frame)
    _, binary, t = prepare_analysis(enh)
    warped, _, _ = warp_image(enh, points)
    squares = split_board(warped, *(grid or (None, None)))
    res = {}
    if keep is not None:
        keep.update(enhanced=enh, binary=binary, gray_squares={})
    for pos, sq in squares.items():
        g = preprocess_square(sq, 5)
        if keep is not None:
            keep["gray_squares"][pos] = g
        st = pd_statistics(g, None if pd_ref is None else pd_ref.get(pos))
        if cd_state is not None:
            if pos not in cd_state:
                cd_state[pos] = (g.astype(np.float32), np.full(g.shape, 100, np.float32))   # change_detector.py:42-45
            m, v = cd_state[pos]
            st["cd"] = cd_detect(g, m, v)
            cd_state[pos] = cd_update(g, m, v)
        res[pos] = st
    return t, res
