"""Short device-resident run of the whole path for ncu (no CPU leg, no e2e leg).
usage: python tools/prof_run.py [frames] [steps]      (CVB_HOUGH=1 adds the Hough-circle launch on the squares of every frame,
CVB_CHECK=1 a small parity check against the oracle)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chessboard_vision_b200 import synth
from chessboard_vision_b200.engine import (Engine, grid_rects, STATS_DTYPE, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE,
                                           SQ_CD_DETECT, SQ_CD_UPDATE)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
H, W, S = 1080, 1920, 620
eng = Engine(0)
frames = synth.frame_batch(min(n, 8), H, W, "board", 0)
d_in = eng.upload(np.stack([frames[i % len(frames)] for i in range(n)]))
rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
M = eng.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
st = eng.new_state(n, S, S)
d_stats = eng.empty((n, 64), STATS_DTYPE)
d_otsu = eng.empty((n,), np.int32)
cal = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), board_size=S)
run = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), board_size=S)
eng.pipeline_dev(d_in, M, rects, cal, st, stats=d_stats, otsu_t=d_otsu)
eng.synchronize()
eng.profile(True)
for _ in range(steps):
    eng.pipeline_dev(d_in, M, rects, run, st, stats=d_stats, otsu_t=d_otsu)
    if os.environ.get("CVB_HOUGH"):
        hres = eng.hough_state(st, rects, None, 0, n)
prof = eng.profile_read()
tot = sum(v[0] for v in prof.values())
for k, (ms, c) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print("%-16s %8.3f ms/launch  %5.1f%%  (%.2f us/frame)" % (k, ms / c, 100 * ms / tot, ms / c / n * 1e3))
if os.environ.get("CVB_HOUGH"):
    print("hough: %.0f edge pixels, %.1f candidate centres, %.2f circles per square" % (
        hres["n_edges"].mean(), hres["n_centers"].mean(), hres["count"].mean()))
print("total %.3f ms/step  -> %.0f frames/s" % (tot / steps, n * steps / tot * 1e3))
if os.environ.get("CVB_CHECK"):
    import oracle as O
    for shp in ((270, 480), (133, 251), (480, 640), (64, 128), (16, 16), (200, 256), (720, 1280), (1080, 1920)):
        for kind in ("board", "noise"):
            f = synth.frame_batch(1, shp[0], shp[1], kind, 3)[0]
            got, ref = eng.process_pipeline(f), O.process_pipeline(f, True)
            bad = int((got != ref).sum())
            print("parity", shp, kind, "OK" if bad == 0 else "MISMATCH %d values, first at %s" % (bad, np.argwhere(got != ref)[0]))
