"""Small whole-path run for compute-sanitizer (memcheck / racecheck) and for the debug build's index checks
(CVB200_LIB=chessboard_vision_b200/libcvb200_dbg.so): odd sizes, partial tiles, all op masks, both ingest formats, the board overlay.
Prints `bounds violations: N` (-1: release library, no checks compiled in)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chessboard_vision_b200 import synth
from chessboard_vision_b200.engine import (Engine, grid_rects, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE)
eng = Engine(0)
for (H, W) in ((135, 241), (64, 64), (270, 480), (92, 160), (360, 640)):
    f = synth.frame_batch(2, H, W, "board", 1)
    eng.enhance(f)
    eng.bilateral(f[0]); eng.sharpen(f[0]); eng.correct_lighting(f[0]); eng.normalize(f[0])
    eng.prepare_analysis(f[0]); eng.gaussian(np.ascontiguousarray(f[0][..., 0]), 13)
    S = 100
    M = eng.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
    rects, _ = grid_rects(S)
    st = eng.new_state(2, S, S)
    for ops in (SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE, SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE, SQ_CD_DETECT, SQ_PD_SET_REF):
        pp = eng.pipeline_params(squares=eng.square_params(ops=ops, cd_blur=5 if ops != SQ_CD_DETECT else 13), board_size=S)
        eng.pipeline(f, M, rects, pp, st)
    st.free()
    if W % 2 == 0 and H % 2 == 0:
        for fmt in ("yuy2", "nv12"):
            eng.cvt_to_bgr(np.stack([synth.bgr_to_yuv(x, fmt) for x in f]), fmt)
from chessboard_vision_b200.overlay import BoardOverlay
for S in (203, 400):                                          # board overlay: tiles cut by the image border, clipped stamps
    BoardOverlay(eng).draw_interface(synth.frame_batch(1, S, S, "board", 2)[0], S, noise_active=True, pieces={(0, 0): "K", (7, 7): "q"},
                                     white_to_move=True, last_move=((0, 0), (7, 7)), lifted=(7, 0), radar=[(7, 7), (0, 7)], fps=1.0)
import ctypes
line = ctypes.c_int(0)
print("bounds violations:", eng.lib.cvb_debug_bounds_violations(eng.h, ctypes.byref(line)), "first at line", line.value)
print("sanitize run done")
