"""Where the gap between the resident-input rate and the end-to-end rate of the kernel-bound (NV12) case comes from:
device time of 256 resident 1080p frames processed in ONE pipeline_dev call against the same frames in slices of 25
(what the host-buffer pipeline launches per chunk), with and without a concurrent host->device copy stream.
    python tools/chunk_device_probe.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chessboard_vision_b200 import synth
from chessboard_vision_b200.engine import Engine, grid_rects, STATS_DTYPE, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE

eng = Engine(0)
H, W, S, n = 1080, 1920, 800, 256
rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
M = eng.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
u = synth.frame_batch(8, H, W, "board", 0)
d_in = eng.upload(np.stack([u[i % 8] for i in range(n)]))
st = eng.new_state(n, S, S)
stats = eng.empty((n, len(rects)), STATS_DTYPE); otsu = eng.empty((n,), np.int32)
cal = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), board_size=S)
run = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), board_size=S)
eng.pipeline_dev(d_in, M, rects, cal, st, stats=stats, otsu_t=otsu)


class View:          # a slice of a device array as the engine sees device arrays
    def __init__(self, base, f0, cnt, per):
        self.ptr = base.ptr + f0 * per; self.shape = (cnt,) + tuple(base.shape[1:])


def whole():
    eng.pipeline_dev(d_in, M, rects, run, st, stats=stats, otsu_t=otsu)


def sliced(chunk):
    def f():
        for f0 in range(0, n, chunk):
            cnt = min(chunk, n - f0)
            eng.pipeline_dev(View(d_in, f0, cnt, H * W * 3), M, rects, run, st, stream0=f0,
                             stats=View(stats, f0, cnt, len(rects) * STATS_DTYPE.itemsize), otsu_t=View(otsu, f0, cnt, 4))
    return f


def timed(fn, reps=5):
    fn(); eng.synchronize()
    e0, e1 = eng.event(), eng.event()
    eng.record(e0)
    for _ in range(reps):
        fn()
    eng.record(e1)
    return eng.elapsed_ms(e0, e1) / reps


print("one call, 256 frames:            %.2f ms" % timed(whole))
for c in (64, 32, 25, 16, 8):
    print("slices of %2d frames:             %.2f ms" % (c, timed(sliced(c))))
# per-kernel times in slices of 25 against one call
for name, fn in (("one call", whole), ("slices of 25", sliced(25))):
    eng.profile(True); fn(); p = eng.profile_read(); eng.profile(False)
    print(name, {k: round(v[0], 2) for k, v in sorted(p.items(), key=lambda kv: -kv[1][0])}, "sum %.2f ms" % sum(v[0] for v in p.values()))

# the same per-kernel sums while the host-buffer pipeline feeds NV12 frames over PCIe (copies concurrent with the kernels)
nat = np.stack([synth.bgr_to_yuv(x, "nv12") for x in u])
hn = eng.pinned((n,) + nat.shape[1:])
for i in range(n):
    hn[i] = nat[i % 8]
for _ in range(2):
    eng.pipeline(hn, M, rects, run, st, fmt="nv12")
t0 = time.perf_counter()
for _ in range(5):
    eng.pipeline(hn, M, rects, run, st, fmt="nv12")
print("e2e nv12 blocking: %.2f ms per step" % ((time.perf_counter() - t0) * 200))
eng.profile(True); eng.pipeline(hn, M, rects, run, st, fmt="nv12"); p = eng.profile_read(); eng.profile(False)
print("e2e nv12 kernels", {k: round(v[0], 2) for k, v in sorted(p.items(), key=lambda kv: -kv[1][0])}, "sum %.2f ms" % sum(v[0] for v in p.values()))
