"""Drop-in for the reference's piece_detector module (piece_detector.py:1-453).

The per-square numerics -- gray + 5x5 blur (piece_detector.py:124-135), the
absdiff-mean delta against the stored reference (:82-93), the std gate (:305),
centre-vs-border means (:177-207) and the four ring means (:141-175) -- run
for all 64 squares in ONE launch of the fused square kernel; the kernel
returns integer sums and this class turns them into the reference's floats
and flags.  The temporal logic (history vote, caching, gating: :99-122,
:348-440) is scalar Python and stays on the host, as in the reference.

`cv2.HoughCircles` (:210-270) runs on the GPU too: one CTA per
square on the gray+blur squares the statistics launch left on the device, for all
squares that need it in one launch, bit-identical to OpenCV's circle list; the
choice of the circle nearest the square centre (:245-268) stays on the host.
"""
import json
import os

import numpy as np

from chessboard_vision_b200 import _lib
from chessboard_vision_b200 import hostapi
from chessboard_vision_b200.engine import default_engine, SQ_PD_STATS, SQ_PD_SET_REF
from chessboard_vision_b200.squarestate import SquareRunner

SETTINGS_FILE = "piece_detector_settings.json"


def _empty_result():
    return {'has_piece': False, 'confidence': 0.0, 'center': None, 'radius': None, 'method': None,
            'center_border_diff': 0, 'is_ellipse': False, 'axes': None}


class PieceDetector:
    def __init__(self, device=0):
        self.min_radius_ratio = 0.20
        self.max_radius_ratio = 0.55
        self.edge_threshold = 50
        self.circle_threshold = 0.6
        self.history_size = 5
        self.min_presence = 0.6
        self.detection_history = {}
        self.load_settings()
        self._e = default_engine(device)
        self._run = SquareRunner(self._e)
        self._one_states = {}       # scratch states for the single-square API, by square shape
        # {(file, rank): gray image}: a dict-like window onto the device reference plane
        self.reference_squares = self._run.bind(_lib.PLANE_PD_REF, np.uint8, 1)
        self.cached_results = {}
        self.change_threshold = 25

    # -- settings (piece_detector.py:52-68) --
    def load_settings(self):
        if os.path.exists(SETTINGS_FILE):
            try:
                with open(SETTINGS_FILE, 'r') as f:
                    params = json.load(f)
                    if 'min_radius' in params:
                        self.min_radius_ratio = params['min_radius'] / 100.0
                    if 'max_radius' in params:
                        self.max_radius_ratio = params['max_radius'] / 100.0
                print(f"[PieceDetector] Settings loaded from {SETTINGS_FILE}")
            except Exception as e:
                print(f"[PieceDetector] Error loading settings: {e}")

    # -- device launches --
    def _stats(self, squares_dict, set_ref_keys=None, stats=True):
        ops = (SQ_PD_STATS if stats else 0) | (SQ_PD_SET_REF if set_ref_keys is not None else 0)
        p = self._e.square_params(ops=ops, pd_blur=5)
        out, keys = self._run.run(squares_dict, p, select_keys=set_ref_keys, want_stats=stats)
        if set_ref_keys is not None:
            self.reference_squares.device_changed([k for k in keys if k in set(set_ref_keys)])
        return out

    def _hough_batch(self, stats, keys):
        """One Hough launch for the squares of `keys` that pass the std gate (piece_detector.py:305)."""
        need = [k for k in keys if not hostapi.std_below(stats[k], 15)]
        return self._run.hough(need, self._hough_params()) if need else {}

    def _shape_of(self, pos):
        x, y, w, h = self._run.layout[pos]
        return np.empty((h, w), np.uint8)

    def _gray_squares(self, keys=None):
        plane = self._run.current_plane()
        lay = self._run.layout
        return {k: plane[y:y + h, x:x + w] for k, (x, y, w, h) in lay.items() if keys is None or k in keys}

    # -- reference management (piece_detector.py:70-97, 447-453) --
    def calibrate_reference(self, squares_dict):
        self.reference_squares.clear()
        self.cached_results.clear()
        stats = self._stats(squares_dict, set_ref_keys=list(squares_dict.keys()))
        circles = self._hough_batch(stats, list(squares_dict.keys()))
        for pos in squares_dict:
            self.cached_results[pos] = self._detect_from(stats[pos], self._shape_of(pos), circles.get(pos))

    def update_references(self, squares_dict):
        self._stats(squares_dict, set_ref_keys=list(squares_dict.keys()), stats=False)
        self.cached_results.clear()

    def _has_changed(self, pos, current_gray):
        if pos not in self.reference_squares:
            return True
        ref = self.reference_squares[pos]
        mean_diff = np.mean(np.abs(current_gray.astype(np.int16) - ref.astype(np.int16)))
        return mean_diff > self.change_threshold

    def _update_reference(self, pos, gray):
        self.reference_squares[pos] = np.array(gray, np.uint8)

    # -- temporal smoothing (piece_detector.py:99-122) --
    def _update_history(self, pos, has_piece):
        hist = self.detection_history.setdefault(pos, [])
        hist.append(has_piece)
        if len(hist) > self.history_size:
            hist.pop(0)

    def _get_stable_detection(self, pos):
        hist = self.detection_history.get(pos)
        if not hist:
            return False
        if len(hist) < 3:
            return hist[-1]
        return sum(hist) / len(hist) >= self.min_presence

    # -- single-square API (piece_detector.py:124-207) --
    def _one(self, square_img):
        sq = np.ascontiguousarray(square_img)
        if sq.dtype != np.uint8 or sq.ndim not in (2, 3):
            raise ValueError("expected a uint8 HxW or HxWx3 square, got %s %r" % (sq.dtype, sq.shape))
        key = sq.shape[:2]
        st = self._one_states.get(key)
        if st is None:
            if len(self._one_states) > 8:
                self._one_states.popitem()[1].free()
            st = self._one_states[key] = self._e.new_state(1, sq.shape[0], sq.shape[1])
        p = self._e.square_params(ops=SQ_PD_STATS, pd_blur=5)
        stats = self._e.squares(sq, [(0, 0, sq.shape[1], sq.shape[0])], p, st)[0, 0]
        return stats, st.get(0, _lib.PLANE_PD_CUR)

    def _preprocess_square(self, square_img):
        return self._one(square_img)[1]

    def _analyze_radial_symmetry(self, gray):
        # `gray` is already preprocessed in the reference's call chain; a 1x1 blur keeps it unchanged
        return hostapi.radial_symmetry(self._raw_stats(gray))

    def _detect_center_vs_border(self, gray):
        return hostapi.center_vs_border(self._raw_stats(gray))

    def _raw_stats(self, gray):
        g = np.ascontiguousarray(gray, np.uint8)
        p = self._e.square_params(ops=SQ_PD_STATS, pd_blur=1)
        return self._e.squares(g, [(0, 0, g.shape[1], g.shape[0])], p)[0, 0]

    def _hough_params(self):
        """The arguments of piece_detector.py:225-241."""
        return self._e.hough_params(dp=1.2, param1=getattr(self, 'hough_param1', 100), param2=getattr(self, 'hough_param2', 25),
                                    min_radius_ratio=self.min_radius_ratio, max_radius_ratio=self.max_radius_ratio,
                                    min_dist_div=3)

    @staticmethod
    def _pick_circle(rec, h, w):
        """piece_detector.py:243-270 on one Hough record -> (found, centre, radius, 'hough' | 'tower_top')"""
        count = int(rec["count"])
        if count > _lib.HOUGH_MAX_CIRCLES:
            raise RuntimeError("%d circles in one square: more than the %d the Hough kernel stores"
                               % (count, _lib.HOUGH_MAX_CIRCLES))
        md = min(h, w)
        best, best_d = None, float('inf')
        for c in rec["xyr"][:count]:
            d = np.sqrt((c[0] - w // 2) ** 2 + (c[1] - h // 2) ** 2)
            if d < md * 0.3 and d < best_d:
                best, best_d = c, d
        if best is None:
            return False, None, None, None
        r = int(best[2])
        return True, (int(best[0]), int(best[1])), r, ('tower_top' if r < md * 0.20 else 'hough')

    def _detect_circle_unified(self, gray):
        """piece_detector.py:210-270 -- cv2.HoughCircles on one gray square, on the GPU.
        -> (found, centre, radius, 'hough' | 'tower_top')"""
        g = np.ascontiguousarray(gray, np.uint8)
        h, w = g.shape
        if min(h, w) // 3 < 1:
            raise ValueError("square of %dx%d: minDist = min_dim // 3 must be positive (cv2.HoughCircles asserts)" % (h, w))
        rec = self._e.hough(g, [(0, 0, w, h)], self._hough_params())[0, 0]
        return self._pick_circle(rec, h, w)

    # -- detection (piece_detector.py:272-346) --
    def _detect_from(self, st, gray, hough_rec=None):
        """`gray` is only used for its shape when the square's Hough record is already there."""
        h, w = gray.shape
        result = _empty_result()
        if hostapi.std_below(st, 15):
            return result
        if hough_rec is not None:
            found, center, radius, kind = self._pick_circle(hough_rec, h, w)
        else:
            found, center, radius, kind = self._detect_circle_unified(gray)
        if found:
            result.update(has_piece=True, center=center, radius=radius, method=kind,
                          confidence=0.9 if kind == 'hough' else 0.75)
            return result
        diff, _, _ = hostapi.center_vs_border(st)
        result['center_border_diff'] = diff
        if diff > 40:
            result.update(has_piece=True, center=(w // 2, h // 2), radius=min(h, w) // 3, method='center_diff',
                          confidence=min(1.0, diff / 80))
            return result
        symmetry = hostapi.radial_symmetry(st)
        if symmetry > self.circle_threshold:
            result.update(has_piece=True, center=(w // 2, h // 2), radius=min(h, w) // 3, method='symmetry',
                          confidence=symmetry)
        return result

    def detect_piece(self, square_img, pos=None):
        st, gray = self._one(square_img)
        return self._detect_from(st, gray)

    def detect_pieces(self, squares_dict, positions):
        """`detect_piece` (piece_detector.py:272-346) for several squares of one board at once: ONE statistics launch
        and ONE Hough launch instead of two launches and a device-to-host read per square.  No reference, cache or
        history is touched, exactly like a sequence of detect_piece calls.  -> {pos: result}"""
        positions = [p for p in positions if p in squares_dict]
        if not positions:
            return {}
        stats = self._stats(squares_dict)
        circles = self._hough_batch(stats, positions)
        return {pos: self._detect_from(stats[pos], self._shape_of(pos), circles.get(pos)) for pos in positions}

    def detect_all_pieces(self, squares_dict, use_smoothing=True, use_delta=True, squares_to_check=None):
        """piece_detector.py:348-440 -> (results, visual_changes); one kernel launch for the 64 squares."""
        results, visual_changes, to_update = {}, set(), []
        if not squares_dict:
            return results, visual_changes
        stats = self._stats(squares_dict)
        plan = {}
        for pos in squares_dict:
            st = stats[pos]
            changed = (not st["has_ref"]) or hostapi.mean_abs_diff(st) > self.change_threshold
            if changed:
                visual_changes.add(pos)
            should_process = squares_to_check is not None and pos in squares_to_check
            if not should_process and (squares_to_check is None or use_delta):
                should_process = pos not in self.cached_results or changed
            plan[pos] = (should_process, should_process or pos not in self.cached_results)
        # every square that is detected this frame goes through one Hough launch
        circles = self._hough_batch(stats, [pos for pos, (_, detect) in plan.items() if detect])
        for pos in squares_dict:
            should_process, detect = plan[pos]
            if detect:
                raw = self._detect_from(stats[pos], self._shape_of(pos), circles.get(pos))
                self.cached_results[pos] = raw.copy()
            else:
                raw = self.cached_results[pos].copy()
            raw_has_piece = raw['has_piece']
            self._update_history(pos, raw_has_piece)
            stable_update = True
            if use_smoothing:
                stable = self._get_stable_detection(pos)
                raw['has_piece'] = stable
                stable_update = raw_has_piece == stable
            if should_process and stable_update:
                to_update.append(pos)
            results[pos] = raw
        if to_update:
            self._stats(squares_dict, set_ref_keys=to_update, stats=False)
        return results, visual_changes

    def get_occupied_squares(self, squares_dict, use_smoothing=True):
        results, _ = self.detect_all_pieces(squares_dict, use_smoothing)
        return {pos for pos, info in results.items() if info['has_piece']}
