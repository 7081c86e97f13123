"""Host-buffer pipeline (cvb_pipeline_fmt) at different chunk sizes and arrival formats: frames/s end to end.
usage: python tools/chunk_sweep.py [frames] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chessboard_vision_b200 import synth
from chessboard_vision_b200.engine import Engine, grid_rects, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
H, W, S = 1080, 1920, 620
eng = Engine(0)
rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
M = eng.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
st = eng.new_state(n, S, S)
uniq = synth.frame_batch(8, H, W, "board", 0)
cal = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), board_size=S)
run = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), board_size=S)
for fmt in ("bgr", "yuy2", "nv12"):
    nat = uniq if fmt == "bgr" else np.stack([synth.bgr_to_yuv(f, fmt) for f in uniq])
    host = eng.pinned((n,) + nat.shape[1:])
    for i in range(n):
        host[i] = nat[i % 8]
    eng.pipeline(host, M, rects, cal, st, fmt=fmt)
    for chunk in (4, 8, 13, 16, 25, 37, 64):
        eng.set_chunk_frames(chunk)
        eng.pipeline(host, M, rects, run, st, fmt=fmt)
        e0, e1 = eng.event(), eng.event()
        eng.record(e0)
        for _ in range(steps):
            eng.pipeline(host, M, rects, run, st, fmt=fmt)
        eng.record(e1)
        ms = eng.elapsed_ms(e0, e1) / steps
        print("%-5s chunk %3d: %7.2f ms/step  %8.0f frames/s  (%.1f GB/s over PCIe)" % (fmt, chunk, ms, n / ms * 1e3, host.nbytes / ms / 1e6), flush=True)
    eng.lib.cvb_host_free(host.ctypes.data)
