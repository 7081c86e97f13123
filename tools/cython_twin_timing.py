"""BASELINE.json configs[2] is quoted "vs src/cython path": this builds the reference's two Cython twins out of tree
(setup.py:5-22; /root/reference is read-only) and times ChangeDetectorCython next to ChangeDetectorPython on the
64 squares of consecutive warped boards.  Build container only:

    python tools/cython_twin_timing.py > profiles/r02_cython_twin.txt
"""
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chessboard_vision_b200 import synth


def main():
    work = tempfile.mkdtemp(prefix="ref_build_")
    dst = os.path.join(work, "ref")
    shutil.copytree("/root/reference", dst)
    r = subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=dst, capture_output=True, text=True)
    print("build_ext --inplace: rc", r.returncode)
    os.chdir(tempfile.mkdtemp(prefix="empty_cwd_"))
    sys.path.insert(0, dst)
    import cv2
    import change_detector as cdm                 # the selector picks the Cython twin when the .so is importable
    import grid_extractor as ge
    import board_detection as bd
    print("selected:", cdm.ChangeDetector, "| cv2", cv2.__version__, "threads", cv2.getNumThreads(), "| cores", os.cpu_count())
    pts = synth.calib_points(1080, 1920)
    boards = [bd.warp_image(synth.board_frame(1080, 1920, s), pts)[0] for s in (0, 100)]
    sq = [ge.GridExtractor().split_board(b) for b in boards]
    impls = [("cython twin", cdm.ChangeDetector)]
    if hasattr(cdm, "ChangeDetectorPython"):
        impls.append(("python class", cdm.ChangeDetectorPython))
    for name, cls in impls:
        det = cls()
        det.calibrate(sq[0])
        det.piece_detector.calibrate_reference(sq[0])
        t_det = t_upd = t_pd = 0.0
        reps = 30
        for _ in range(reps):
            t0 = time.perf_counter(); det.detect_changes_detailed(sq[1]); t1 = time.perf_counter()
            det.update_all_references(sq[1]); t2 = time.perf_counter()
            for p, s in sq[1].items():
                det.piece_detector._has_changed(p, det.piece_detector._preprocess_square(s))
            t3 = time.perf_counter()
            t_det += t1 - t0; t_upd += t2 - t1; t_pd += t3 - t2
        tot = (t_det + t_upd + t_pd) / reps
        print("%-12s per pair of frames (64 squares): detect_changes_detailed %.2f ms, update_all_references %.2f ms, "
              "_preprocess_square + _has_changed %.2f ms -> %.2f ms = %.0f pairs/s on this container's %d cores"
              % (name, t_det / reps * 1e3, t_upd / reps * 1e3, t_pd / reps * 1e3, tot * 1e3, 1 / tot, os.cpu_count()))
    print("(detect_changes_detailed includes PieceDetector.detect_piece with cv2.HoughCircles for every square >= 5 % changed, "
          "change_detector.py:156; the GPU number of `bench.py --config change64` covers the numeric part + statistics, "
          "Hough circles are timed separately as extras.next_hough_circles_64_squares)")
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
