"""Drop-in for the reference's change_detector module (change_detector.py:1-210)
and its Cython twin (src/cython/change_detector_cython.pyx), which this module
replaces at the selector seam (change_detector.py:7-19,203-208).

The running Gaussian background model -- gray + blur per square
(change_detector.py:49-56), f32 mean / variance EMA (:77-92) and the z-score
test (:121-137) -- lives on the GPU: one launch of the fused square kernel
handles all 64 squares, `means` / `variances` are dict-like windows onto the
device planes.  Classification thresholds (:141-150) and classify_hand_pattern
(:169-201) are dictionary logic and stay in Python, as in the reference.
"""
import numpy as np

from chessboard_vision_b200 import _lib
from chessboard_vision_b200.engine import default_engine, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE
from chessboard_vision_b200.squarestate import SquareRunner
from piece_detector import PieceDetector     # resolved like the reference does (change_detector.py:3)

USE_CYTHON = False
USE_B200 = True


class ChangeDetectorB200:
    def __init__(self, device=0):
        self.z_threshold = 2.5
        self.initial_variance = 100
        self.alpha = 0.1
        self.blur_kernel = 5
        self._kernel = 5
        self._e = default_engine(device)
        self._run = SquareRunner(self._e)
        self.means = self._run.bind(_lib.PLANE_CD_MEAN, np.float32, 2)
        self.variances = self._run.bind(_lib.PLANE_CD_VAR, np.float32, 4)
        self.is_calibrated = False
        self.focus_squares = set()
        self.piece_detector = PieceDetector()

    def _params(self, ops):
        return self._e.square_params(ops=ops, pd_blur=5, cd_blur=int(self.blur_kernel) | 1,
                                     z_threshold=self.z_threshold, alpha=self.alpha,
                                     initial_variance=self.initial_variance)

    def _launch(self, squares, ops, select_keys=None, want_stats=False):
        stats, keys = self._run.run(squares, self._params(ops), select_keys=select_keys, want_stats=want_stats)
        if ops & (SQ_CD_CALIBRATE | SQ_CD_UPDATE):
            touched = keys if select_keys is None else [k for k in keys if k in set(select_keys)]
            self.means.device_changed(touched if ops & SQ_CD_CALIBRATE else ())
            self.variances.device_changed(touched if ops & SQ_CD_CALIBRATE else ())
        return stats, keys

    def calibrate(self, squares):
        """change_detector.py:36-47."""
        self.means.clear()
        self.variances.clear()
        if squares:
            self._launch(squares, SQ_CD_CALIBRATE)
        self.is_calibrated = True

    def _preprocess(self, img):
        """change_detector.py:49-56 (single square, stand-alone API)."""
        img = np.ascontiguousarray(img)
        k = int(self.blur_kernel) | 1
        if img.ndim == 3:
            img = self._e.gray(img)
        return self._e.gaussian(img, k)

    def set_focus_squares(self, squares):
        self.focus_squares = set(squares)

    def clear_focus(self):
        self.focus_squares = set()

    def get_focus_count(self):
        return len(self.focus_squares) if self.focus_squares else 64

    def update_all_references(self, squares):
        """change_detector.py:67-92."""
        if not self.is_calibrated:
            self.calibrate(squares)
            return
        sel = [k for k in squares if k in self.focus_squares] if self.focus_squares else None
        if sel is not None and not sel:
            return
        for pos in (sel if sel is not None else squares):
            if pos not in self.means or pos not in self.variances:
                raise KeyError(pos)            # the reference indexes self.means[pos] (change_detector.py:81)
        self._launch(squares, SQ_CD_UPDATE, select_keys=sel)

    def detect_changes(self, squares):
        """change_detector.py:94-103."""
        return {pos: info['pct_changed'] for pos, info in self.detect_changes_detailed(squares).items()
                if info['intensity'] in ['PARCIAL', 'TOTAL']}

    def detect_changes_detailed(self, squares):
        """change_detector.py:105-167."""
        results = {}
        if not self.is_calibrated or not squares:
            return results
        to_check = list(self.focus_squares) if self.focus_squares else list(squares.keys())
        to_check = [p for p in to_check if p in squares]
        if not to_check:
            return results
        stats, _ = self._launch(squares, SQ_CD_DETECT, select_keys=to_check if self.focus_squares else None,
                                want_stats=True)
        hits = []
        for pos in to_check:
            st = stats[pos]
            if not st["cd_valid"]:
                continue                       # no model for this square (change_detector.py:125)
            pct_changed = (int(st["cd_changed"]) / int(st["n"])) * 100
            if pct_changed < 5.0:
                continue
            intensity = 'TOTAL' if pct_changed > 75 else 'PARCIAL' if pct_changed > 15 else 'LEVE'
            hits.append((pos, float(st["cd_zmax"]), pct_changed, intensity))
        # change_detector.py:156 asks the piece detector about every reported square; here all of them go through one
        # statistics launch and one Hough launch (a replaced piece_detector without the batch method is called per square)
        batch = getattr(self.piece_detector, 'detect_pieces', None)
        pieces = batch(squares, [h[0] for h in hits]) if (hits and callable(batch)) else None
        for pos, zmax, pct_changed, intensity in hits:
            pd_result = pieces[pos] if isinstance(pieces, dict) and pos in pieces else self.piece_detector.detect_piece(squares[pos], pos)
            results[pos] = {'z_score': zmax, 'pct_changed': pct_changed, 'intensity': intensity,
                            'is_circular': pd_result['has_piece'], 'center_ratio': 1.0}
        return results

    def classify_hand_pattern(self, detailed):
        """change_detector.py:169-201."""
        n = len(detailed)
        n_total = sum(1 for v in detailed.values() if v['intensity'] == 'TOTAL')
        if n_total >= 2 or n >= 4 or n > 2:
            return {'is_hand': True, 'is_move': False, 'move_candidates': set()}
        cands = set(detailed.keys())
        return {'is_hand': False, 'is_move': len(cands) == 2, 'move_candidates': cands}


def __getattr__(name):
    if name == "ChangeDetectorPython":
        from chessboard_vision_b200.hostapi import load_reference_module
        ref = load_reference_module("change_detector")
        if ref is not None:
            return ref.ChangeDetectorPython
    raise AttributeError(name)


ChangeDetector = ChangeDetectorB200
print("[INFO] ChangeDetector: B200 (sm_100a) backend")
