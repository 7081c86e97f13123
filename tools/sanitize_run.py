"""Small whole-path run for compute-sanitizer (memcheck / racecheck): odd sizes, partial tiles, all op masks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chessboard_vision_b200 import synth
from chessboard_vision_b200.engine import (Engine, grid_rects, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE)
eng = Engine(0)
for (H, W) in ((135, 241), (64, 64), (270, 480)):
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
print("sanitize run done")
