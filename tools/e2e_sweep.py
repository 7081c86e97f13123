"""e2e (host buffers) throughput vs chunk size; also the raw pinned H2D bandwidth."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chessboard_vision_b200 import synth
from chessboard_vision_b200.engine import Engine, grid_rects, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE
n, H, W, S = 256, 1080, 1920, 620
eng = Engine(0)
fr = synth.frame_batch(8, H, W, "board", 0)
host = eng.pinned((n, H, W, 3))
for i in range(n): host[i] = fr[i % 8]
rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
M = eng.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
st = eng.new_state(n, S, S)
cal = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), board_size=S)
run = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), board_size=S)
eng.pipeline(host, M, rects, cal, st)
d = eng.empty((n, H, W, 3))
e0, e1 = eng.event(), eng.event()
for _ in range(2):
    eng.record(e0); eng.lib.cvb_memcpy_h2d(eng.h, d.ptr, host.ctypes.data, host.nbytes); eng.record(e1)
    ms = eng.elapsed_ms(e0, e1)
print("pinned H2D %.1f MB in %.2f ms = %.1f GB/s" % (host.nbytes / 1e6, ms, host.nbytes / ms / 1e6))
for chunk in (4, 8, 16, 32, 64, 256):
    eng.set_chunk_frames(chunk)
    for _ in range(2): eng.pipeline(host, M, rects, run, st)
    eng.record(e0)
    for _ in range(4): eng.pipeline(host, M, rects, run, st)
    eng.record(e1)
    ms = eng.elapsed_ms(e0, e1) / 4
    print("chunk %3d: %.2f ms/step  %.0f frames/s" % (chunk, ms, n / ms * 1e3))
