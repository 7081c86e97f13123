#!/usr/bin/env python
"""Benchmark of the ChessVision hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (OpenCV)

A "step" is one pass of the full path (enhance chain + Otsu analysis + warp +
64 squares + PieceDetector statistics + ChangeDetector detect/update) over one
batch of synthetic 1080p BGR frames (BASELINE.json configs[3]: 256 frames per
GPU; weak scaling: every rank owns its own 256 frames and its own per-stream
state, no data-path collective).  Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "1080p frames/sec/GPU full enhance+grid+change pipeline; HBM GB/s vs roofline"
H, W, S = 1080, 1920, 620
UNIQUE = 16        # distinct synthetic frames, tiled to the batch (content does not change the work)


def measured_traffic():
    """DRAM bytes per 1080p frame per kernel from the committed ncu capture (profiles/), or {}."""
    best = {}
    d = os.path.join(ROOT, "profiles")
    if os.path.isdir(d):
        for f in sorted(os.listdir(d)):
            if f.endswith("_traffic.json"):
                try:
                    best = json.load(open(os.path.join(d, f)))
                except Exception:
                    pass
    return best


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# CPU reference leg (oracle/ is only touched here: cpu_baseline and --impl reference)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, seeds = args
    from chessboard_vision_b200 import synth
    pts = synth.calib_points(H, W)
    n = 0
    if kind == "cv2":
        import cv2
        from oracle import ref_cv2
        cv2.setNumThreads(1)
        cd = {}
        code = {"yuy2": cv2.COLOR_YUV2BGR_YUY2, "nv12": cv2.COLOR_YUV2BGR_NV12}.get(_INGEST)
        for s in seeds:
            f = _FRAMES[s % len(_FRAMES)]
            if code is not None:
                f = cv2.cvtColor(f, code)          # what cv2.VideoCapture does before read() returns (play_lichess.py:45)
            ref_cv2.full_frame(f, pts, cd_state=cd, pd_ref=None)
            n += 1
    else:
        import oracle as O
        M = O.get_perspective(pts, [[0, 0], [S, 0], [0, S], [S, S]])
        for s in seeds:
            f = _FRAMES[s % len(_FRAMES)]
            if _INGEST != "bgr":
                f = O.yuv_to_bgr(f, _INGEST)
            enh = O.process_pipeline(f, True)
            O.prepare_analysis(enh)
            board = O.warp(enh, M, S)
            for r in range(8):
                for c in range(8):
                    g = O.square_preprocess(board[r * 77:(r + 1) * 77, c * 77:(c + 1) * 77], 5)
                    O.pd_square_stats(g)
                    m, v = O.cd_calibrate(g)
                    O.cd_detect(g, m, v)
                    O.cd_update(g, m, v)
            n += 1
    return n


_FRAMES = None
_INGEST = "bgr"


def cpu_reference(frames, n_frames, pool, kind):
    """Frames/s of the reference CPU path over `n_frames` frames using every host core
    (one process per core, one OpenCV thread each: BASELINE.md section 3c)."""
    workers = pool._processes
    chunks = [list(range(i, n_frames, workers)) for i in range(workers)]
    chunks = [(kind, c) for c in chunks if c]
    t0 = time.perf_counter()
    done = sum(pool.map(_cpu_worker, chunks))
    dt = time.perf_counter() - t0
    return done / dt, dt


def make_pool(frames, ingest="bgr"):
    """Fork-based pool created BEFORE any CUDA context exists in this process."""
    import multiprocessing as mp
    global _FRAMES, _INGEST
    _FRAMES, _INGEST = frames, ingest
    try:
        import cv2  # noqa: F401
        kind = "cv2"
    except Exception:
        kind = "port"
        import oracle as O
        O.lib()
    cores = os.cpu_count() or 1
    pool = mp.get_context("fork").Pool(cores)
    pool.map(_cpu_worker, [(kind, [0])] * cores)     # warm every worker (imports, tables)
    return pool, kind, cores


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi in loop mode, started before the warm-up so that it is already sampling when the timed
    region begins; only the samples whose timestamp falls inside the timed region are summarised."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_ready(self, timeout=3.0):
        """Block until the first sample has been written (nvidia-smi takes a few 100 ms to start)."""
        t = time.time()
        while self.p is not None and time.time() - t < timeout:
            try:
                if os.path.getsize(self.f.name) > 0:
                    return
            except OSError:
                return
            time.sleep(0.02)

    def stop(self, t0, t1):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[1]), float(c[2]), c[5:9]))
            except ValueError:
                continue
        self.f.close()
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        inside = [r for r in rows if t0 <= r[0] <= t1] or [r for r in rows if t0 - 0.3 <= r[0] <= t1 + 0.3]
        if inside:
            reasons = set()
            for r in inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            out.update(sm_mhz=float(np.median([r[1] for r in inside])), sm_max_mhz=float(max(r[2] for r in inside)),
                       reasons=sorted(reasons), samples=len(inside))
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=256, help="frames per GPU per step")
    ap.add_argument("--kind", default="board", choices=["board", "noise"])
    ap.add_argument("--config", default="throughput", choices=["throughput", "enhance480", "latency", "change64", "streams4k"],
                    help="throughput = BASELINE.json configs[3] (the metric; default); enhance480 = configs[0]; latency = configs[1]; change64 = configs[2]; "
                         "streams4k = configs[4] (tools/bench_modes.py)")
    ap.add_argument("--ingest", default="bgr", choices=["bgr", "yuy2", "nv12"],
                    help="format the frames arrive in: BGR as cv2.VideoCapture.read() returns them (BASELINE configs), or the "
                         "camera's native YUY2 / NV12, converted to BGR on the device (this arm) or by cv2.cvtColor (reference arm)")
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames in the CPU sample (0: 4 per host core)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the single-frame latency / config-3 side measurements")
    ap.add_argument("--no-parity", action="store_true", help="skip the end-to-end parity block (reference cv2 sequence vs CUDA)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    # stdout carries exactly one JSON line: anything libraries print meanwhile (NCCL banner, ...) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    from chessboard_vision_b200 import synth

    config = {"workload": "BASELINE.json configs[3]: batch of %d synthetic 1080p BGR frames per GPU, full pipeline "
                          "(LAB+CLAHE+bilateral d=9+sharpen+normalize, gray+blur+Otsu mask, warp to 620x620, 64 squares, "
                          "piece_detector statistics + change_detector detect/update)" % args.frames,
              "frames_per_gpu": args.frames, "height": H, "width": W, "frame_kind": args.kind,
              "ingest": args.ingest + {"bgr": " (3 bytes / pixel, as cv2.VideoCapture.read() returns frames)",
                                       "yuy2": " (camera-native, 2 bytes / pixel, converted on the device like cv2.cvtColor)",
                                       "nv12": " (decoder-native, 1.5 bytes / pixel, converted on the device like cv2.cvtColor)"}[args.ingest],
              "unique_frames": UNIQUE, "parallelism": "frames sharded over %d GPU(s), no collective" % world,
              "numa": "each rank pinned to its GPU's CPUs (NVML affinity) before allocating page-locked buffers",
              "l2_policy": "inputs larger than L2 (%.0f MB per step per GPU vs 126 MB)" % (args.frames * H * W * 3 / 1e6)}

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return
        frames = synth.frame_batch(UNIQUE, H, W, args.kind, 0)
        if args.ingest != "bgr":
            frames = np.stack([synth.bgr_to_yuv(f, args.ingest) for f in frames])
        pool, kind, cores = make_pool(frames, args.ingest)
        per_step = args.cpu_frames or 2 * cores
        for _ in range(args.warmup):
            cpu_reference(frames, per_step, pool, kind)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_reference(frames, per_step, pool, kind)
        dt = time.perf_counter() - t0
        pool.close()
        fps = per_step * args.steps / dt
        sample = "%d frames per step (of the %d-frame batch), %d processes x 1 OpenCV thread" % (per_step, args.frames, cores)
        emit({
            "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 (f32 bilateral / CLAHE blend, f64 Otsu and warp coordinates)", "data": "synthetic",
            "config": config,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores,
                             "kind": "port", "sample": sample,
                             "what": ("oracle/ref_cv2.py: the reference's own cv2/numpy call sequence on OpenCV %s" %
                                      __import__("cv2").__version__) if kind == "cv2" else
                                     "oracle/cvb_oracle.c scalar C restatement (cv2 not importable)"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        return

    # ------------------------------------------------------------------ this repo's arm
    frames_u = synth.frame_batch(UNIQUE, H, W, args.kind, 0)
    native_u = frames_u if args.ingest == "bgr" else np.stack([synth.bgr_to_yuv(f, args.ingest) for f in frames_u])
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        pool, kind, cores = make_pool(native_u, args.ingest)          # fork before CUDA
        n_cpu = args.cpu_frames or 4 * cores
        fps_cpu, dt_cpu = cpu_reference(native_u, n_cpu, pool, kind)
        pool.close()
        cpu = {"value": fps_cpu, "unit": "frames/s", "cores": cores, "kind": "port",
               "sample": "%d of the %d frames of one step, %.1f s wall, %d processes x 1 OpenCV thread" %
                         (n_cpu, args.frames, dt_cpu, cores),
               "what": "oracle/ref_cv2.py (reference call sequence on cv2/numpy)" if kind == "cv2"
                       else "oracle/cvb_oracle.c scalar C restatement"}

    dist = None
    device = local_rank
    if world > 1:
        import torch
        import torch.distributed as dist
        from chessboard_vision_b200.sharding import place_rank
        # partial-node runs use the GPUs with the fast host path (profiles/r02_h2d_probe_8gpu.txt); identity at N = node size
        device = place_rank(local_rank, world, torch.cuda.device_count())
        torch.cuda.set_device(device)
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    config["placement"] = "rank r on GPU %s" % ("r" if device == local_rank else "(%d + r): GPUs 4-7 of this pool's boxes keep 55 GB/s "
                          "host->device each when used together, GPUs 0-3 share 115 GB/s (profiles/r02_h2d_probe_8gpu.txt)" % (device - local_rank))

    from chessboard_vision_b200.sharding import bind_to_gpu_numa
    numa_cpus = bind_to_gpu_numa(device) if world > 1 else 0
    from chessboard_vision_b200.engine import (Engine, grid_rects, _rect_array, STATS_DTYPE, SQ_PD_STATS, SQ_PD_SET_REF,
                                               SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE)
    eng = Engine(device)
    n = args.frames

    def barrier():
        eng.synchronize()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = eng.event(), eng.event()
        l0 = eng.launch_count()
        eng.record(e0)
        fn(k)
        eng.record(e1)
        ms = eng.elapsed_ms(e0, e1)
        barrier()
        launches = eng.launch_count() - l0
        if dist is not None:
            import torch
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    if args.config != "throughput":
        # the other BASELINE configurations (tools/bench_modes.py): same JSON contract, their own metric
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_modes
        sampler = ClockSampler(device) if rank == 0 else None
        if sampler:
            sampler.wait_ready()
        wall0 = time.time()
        out = bench_modes.MODES[args.config]({"eng": eng, "synth": synth, "args": args, "rank": rank, "world": world,
                                              "timed": timed})
        clocks = sampler.stop(wall0, time.time()) if sampler else None
        if rank == 0:
            out.update({"n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "scaling": "weak", "vs_baseline": None,
                        "dtype": "u8 (f32 bilateral / CLAHE blend, f64 Otsu and warp coordinates)", "data": "synthetic",
                        "clocks": clocks})
            emit(out)
        if dist is not None:
            dist.destroy_process_group()
        return

    rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
    M = eng.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
    state = eng.new_state(n if world == 1 else 2 * n, S, S)

    # host batch in pinned memory (e2e source) and a resident device copy (kernel-only source), both in the ingest format
    host = eng.pinned((n,) + native_u.shape[1:])
    for i in range(n):
        host[i] = native_u[(i + rank) % UNIQUE]
    d_native = eng.empty(host.shape)
    eng.lib.cvb_memcpy_h2d(eng.h, d_native.ptr, host.ctypes.data, host.nbytes)
    d_in = d_native if args.ingest == "bgr" else eng.empty((n, H, W, 3))
    d_stats = eng.empty((n, len(rects)), STATS_DTYPE)
    d_otsu = eng.empty((n,), np.int32)
    eng.synchronize()

    sq_cal = eng.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE)
    sq_run = eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE)
    pp_cal = eng.pipeline_params(squares=sq_cal, board_size=S)
    pp = eng.pipeline_params(squares=sq_run, board_size=S)
    # calibration frame: references + background model for every stream slot (untimed)
    if args.ingest != "bgr":
        eng.cvt_to_bgr_dev(d_native, args.ingest, n, H, W, d_in)
    eng.pipeline_dev(d_in, M, rects, pp_cal, state, stats=d_stats, otsu_t=d_otsu)
    eng.synchronize()

    def run_dev(k):
        for _ in range(k):
            if args.ingest != "bgr":
                eng.cvt_to_bgr_dev(d_native, args.ingest, n, H, W, d_in)
            eng.pipeline_dev(d_in, M, rects, pp, state, stats=d_stats, otsu_t=d_otsu)

    my_frames = [n]        # frames this rank feeds per e2e step (equal shares until the feed rates are known)

    def run_e2e(k):
        # the public host-buffer calls of a capture loop: submit a batch (H2D + kernels + D2H of the per-square results
        # enqueued), then collect the PREVIOUS batch's results while this one runs -- one batch in flight ahead, so the
        # first copy of a batch overlaps the kernels of the one before.  Everything submitted is waited for before the
        # function returns.  A share above the n frames of the pinned batch wraps around (a second call on the first
        # frames, their state in the upper stream slots).
        pending = []
        for _ in range(k):
            done = 0
            while done < my_frames[0]:
                cnt = min(n, my_frames[0] - done)
                pending.append(eng.pipeline_submit(host[:cnt], M, rects, pp, state, stream0=done, fmt=args.ingest))
                if len(pending) > 1:
                    eng.pipeline_wait(pending.pop(0))
                done += cnt
        while pending:
            eng.pipeline_wait(pending.pop(0))

    sampler = ClockSampler(device) if rank == 0 else None
    if sampler:
        sampler.wait_ready()
    run_dev(args.warmup)
    wall0 = time.time()
    ms_dev, launches = timed(run_dev, args.steps)
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if sampler else None

    # per-kernel device time (CUDA events around every launch, same K steps) -> roofline of the dominant kernel
    eng.profile(True)
    run_dev(args.steps)
    prof = eng.profile_read()
    eng.profile(False)

    run_e2e(args.warmup)
    e2e_shares = None
    if dist is not None:
        # The GPUs of a node need not share the host's PCIe / memory path equally (measured: profiles/r02_h2d_probe_8gpu.txt),
        # and every rank of this path is bound by its host->device copy: give each rank a share of the N * frames of a step
        # in proportion to the rate it just sustained, so that all ranks finish together.  No data moves between ranks.
        import torch
        from chessboard_vision_b200.sharding import proportional_split
        barrier()
        t0 = time.perf_counter(); run_e2e(3); eng.synchronize(); mine = 3 * n / (time.perf_counter() - t0)
        rates = torch.zeros(world, dtype=torch.float64, device="cuda"); rates[rank] = mine
        dist.all_reduce(rates)
        rl = rates.tolist()
        if max(rl) > 1.15 * min(rl):          # below that the differences are measurement noise: equal shares
            e2e_shares = [min(2 * n, x) for x in proportional_split(n * world, rl, multiple=8)]
            e2e_shares[e2e_shares.index(max(e2e_shares))] += n * world - sum(e2e_shares)
        else:
            e2e_shares = [n] * world
        my_frames[0] = e2e_shares[rank]
        eng.pipeline(host[:n], M, rects, pp_cal, state, stream0=n, fmt=args.ingest)      # models for the upper stream slots
        run_e2e(1)
    ms_e2e, _ = timed(run_e2e, args.steps)

    # ---- side measurements (rank 0, not part of value/e2e): the other BASELINE.json configurations through the very
    # functions of `--config latency|change64|streams4k` (short runs), the worst-case input kind, and the end-to-end
    # rate when the frames arrive in the camera's native format instead of BGR ----
    extras = {}
    if rank == 0 and not args.no_extras and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_modes
        side = argparse.Namespace(**vars(args)); side.steps, side.warmup, side.kind = 10, 3, "board"
        for name in ("enhance480", "latency", "change64", "streams4k"):
            r = bench_modes.MODES[name]({"eng": eng, "synth": synth, "args": side, "rank": 0, "world": 1, "timed": timed})
            extras["config_" + name] = {k: r[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "cpu_baseline",
                                                          "stages_us") if k in r}
            extras["config_" + name]["workload"] = r["config"]["workload"]
        if args.ingest == "bgr":
            other = "noise" if args.kind == "board" else "board"
            fo = synth.frame_batch(UNIQUE, H, W, other, 0)
            for i in range(n):
                host[i] = fo[i % UNIQUE]
            eng.lib.cvb_memcpy_h2d(eng.h, d_in.ptr, host.ctypes.data, host.nbytes)
            run_dev(3)
            ms_o, _ = timed(run_dev, 5)
            extras["frame_kind_" + other] = {"value": n * 5 / (ms_o / 1e3), "unit": "frames/s", "ms_per_step": ms_o / 5,
                                             "what": "the resident-input number of this line on '%s' frames (SURVEY.md 8d: uniform "
                                                     "noise is the worst case for the histogram atomics and the range-weight "
                                                     "lookups)" % other}
            for fmt in ("yuy2", "nv12"):
                nat = np.stack([synth.bgr_to_yuv(f, fmt) for f in frames_u])
                hn = eng.pinned((n,) + nat.shape[1:])
                for i in range(n):
                    hn[i] = nat[i % UNIQUE]
                def run_n(k, hn=hn, fmt=fmt):
                    pend = []
                    for _ in range(k):
                        pend.append(eng.pipeline_submit(hn, M, rects, pp, state, fmt=fmt))
                        if len(pend) > 1:
                            eng.pipeline_wait(pend.pop(0))
                    while pend:
                        eng.pipeline_wait(pend.pop(0))
                run_n(3)
                ms_n, _ = timed(run_n, 8)
                extras["e2e_ingest_" + fmt] = {"value": n * 8 / (ms_n / 1e3), "unit": "frames/s", "ms_per_step": ms_n / 8,
                                               "h2d_bytes_per_step": int(hn.nbytes),
                                               "what": "the e2e number of this line with frames arriving as %s (%.1f bytes per "
                                                       "pixel over PCIe), converted on the device bit-exactly as cv2.cvtColor "
                                                       "does (`--ingest %s` makes it the headline of both arms)" %
                                                       (fmt.upper(), hn.nbytes / (n * H * W), fmt)}
                eng.lib.cvb_host_free(hn.ctypes.data)
        # "next" row: Hough circles on the 64 gray+blur squares of warped boards with pieces (k_hough),
        # beside the reference's cv2.HoughCircles loop on this box's host cores
        nh = 64
        pboards = np.stack([synth.board_with_pieces(30 + i, 7, S)[1] for i in range(8)])
        sth = eng.new_state(nh, S, S)
        eng.squares(np.stack([pboards[i % 8] for i in range(nh)]), rects, eng.square_params(ops=SQ_PD_STATS), sth,
                    want_stats=False)
        hp = eng.hough_params()
        hres = eng.hough_state(sth, rects, hp, 0, nh)
        eng.profile(True)
        for _ in range(5):
            eng.hough_state(sth, rects, hp, 0, nh)
        hprof = eng.profile_read()
        eng.profile(False)
        ms_h = hprof["k_hough"][0] / hprof["k_hough"][1]
        cpu_h = None
        try:
            from oracle import ref_cv2
            planes_h = [sth.get(i, 3) for i in range(2)]
            t0 = time.perf_counter()
            n_found = 0
            for pl in planes_h:
                for (x, y, w, h) in rects:
                    n_found += ref_cv2.detect_circle_unified(np.ascontiguousarray(pl[y:y + h, x:x + w]))[0][0]
            cpu_h = (time.perf_counter() - t0) * 1e3 / len(planes_h)
        except ImportError:
            pass
        extras["next_hough_circles_64_squares"] = {
            "device_ms_per_batch": ms_h, "frames": nh, "us_per_frame": ms_h / nh * 1e3,
            "circles_per_frame": float(hres["count"].sum()) / nh, "edge_pixels_per_square": float(hres["n_edges"].mean()),
            "cpu_reference_ms_per_frame": cpu_h,
            "what": "cv2.HoughCircles(dp 1.2, 100/25, radii 20-55 %) on 64 squares per frame, one CTA per square; "
                    "CPU: the reference's loop over 64 squares with cv2 (its default threads), 2 frames"}
        sth.free()

    # ---- the other half of SURVEY 8f rank 4: GameSession._draw_interface on the device (csrc/cvb_overlay.cu)
    if rank == 0 and not args.no_extras and world == 1:
        from chessboard_vision_b200.overlay import BoardOverlay
        Sb = 800                                                     # the reference's board size (game_session.py:94)
        start = {}
        for f_, s_ in enumerate("RNBQKBNR"):
            start[(f_, 0)] = s_; start[(f_, 1)] = "P"; start[(f_, 6)] = "p"; start[(f_, 7)] = s_.lower()
        ov_state = dict(noise_active=True, last_move=((4, 1), (4, 3)), lifted=(6, 7), radar=[(5, 5), (7, 5)], pieces=start,
                        white_to_move=False, fps=30.0)
        board_img = np.ascontiguousarray(synth.frame_batch(1, Sb, Sb, "board", 3)[0])
        dl = BoardOverlay.display_list(Sb, **ov_state)
        ops_o, n_o, masks_o = dl.pack()
        boards_dev = eng.upload(np.stack([board_img] * 16))
        run_ov = lambda k: [eng.overlay(boards_dev, ops_o, n_o, masks_o) for _ in range(k)]
        run_ov(3)
        ms_ov, _ = timed(run_ov, 10)
        t0 = time.perf_counter()
        for _ in range(20):
            BoardOverlay.display_list(Sb, **ov_state).pack()
        host_list_ms = (time.perf_counter() - t0) * 1e3 / 20
        cpu_ov = same = None
        try:
            from oracle import overlay as ov_oracle
            ref_img = ov_oracle.draw_interface_cv2(board_img.copy(), Sb, **ov_state)
            same = bool(np.array_equal(eng.overlay(board_img, ops_o, n_o, masks_o), ref_img))
            t0 = time.perf_counter()
            for _ in range(10):
                ov_oracle.draw_interface_cv2(board_img.copy(), Sb, **ov_state)
            cpu_ov = (time.perf_counter() - t0) * 1e3 / 10
        except ImportError:
            pass
        extras["next_board_overlay"] = {
            "device_us_per_board": ms_ov / 10 / 16 * 1e3, "display_list_ops": n_o, "host_list_build_ms": host_list_ms,
            "cpu_reference_ms_per_board": cpu_ov, "identical_to_cv2_sequence": same,
            "what": "GameSession._draw_interface (grid, noise tint + text, last move, lifted square, 2 radar discs, 32 piece "
                    "letters, status texts) on an 800 x 800 warped board: one k_overlay launch per board batch of 16 "
                    "(list upload included); CPU: the same cv2 call sequence (oracle/overlay.py)"}
        boards_dev.free()

    # ---- parity beside the throughput number: the reference's cv2 / numpy call sequence against the
    # CUDA path on (previous, current) pairs of the very frame kinds that are timed: Otsu T, mask per pixel, change flags
    parity = None
    if rank == 0 and not args.no_parity:
        try:
            from oracle import parity as P
            parity = {"what": "oracle/parity.py: ref_cv2 (cv2 %s call sequence of the reference) vs the C ABI on 1080p pairs; "
                              "`cur` = `prev` with four blocks overwritten (LEVE / PARCIAL / TOTAL squares)" %
                              __import__("cv2").__version__}
            for kind in ("board", "noise"):
                parity[kind] = P.run(eng, list(synth.frame_batch(2, H, W, kind, 0)), H, W)
        except ImportError as e:
            parity = {"unavailable": "cv2 is not importable on this box (%s)" % e}

    total_frames = n * args.steps * world
    value = total_frames / (ms_dev / 1e3)
    e2e = total_frames / (ms_e2e / 1e3)
    if rank == 0:
        npx = H * W
        peak, peak_src = peaks()
        tot_ms = sum(v[0] for v in prof.values()) or 1.0
        # algorithmic bytes per frame per kernel
        alg = {"k_tile_hist": 6 * npx, "k_fused": 6 * npx, "k_finish": 8 * npx, "k_threshold": 2 * npx,
               "k_yuv2bgr": int({"yuy2": 5, "nv12": 4.5}.get(args.ingest, 0) * npx),
               "k_warp": 1012 * 916 * 3 + S * S * 3, "k_squares": S * S * 3 + 64 * 77 * 77 * (1 + 4 + 4 + 4 + 4 + 1),
               "k_clahe_lut": 64 * 256 * 5, "k_otsu": 1024}
        stages = {}
        for name, (ms, cnt) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
            per_launch_ms = ms / cnt
            b = alg.get(name, 0) * n
            stages[name] = {"ms_per_launch": per_launch_ms, "share": ms / tot_ms, "launches": cnt,
                            "alg_bytes_per_launch": b, "achieved_gbs": b / per_launch_ms / 1e6,
                            "frac_of_hbm_peak": b / per_launch_ms / 1e6 / peak}
        traffic = measured_traffic()
        per_frame = traffic.get("dram_bytes_per_frame", {})
        for name in stages:
            stages[name]["traffic_bytes_per_launch"] = per_frame[name] * n if name in per_frame else None
            pp_ = traffic.get("pipes_pct", {}).get(name)
            if pp_:           # the pipe ncu shows busiest for this kernel and its utilisation (percent of peak)
                k_, v_ = max(pp_.items(), key=lambda kv: kv[1])
                stages[name]["busiest_pipe"] = {"name": k_, "pct": v_}
        dom = max(prof.items(), key=lambda kv: kv[1][0])[0]
        d = stages[dom]
        pipes = traffic.get("pipes_pct", {}).get(dom, {})
        busiest = max(pipes.items(), key=lambda kv: kv[1]) if pipes else None
        pipe_names = {"lsu": "shared-memory (LSU) pipe", "issue": "instruction issue slots", "fma": "FMA pipe", "alu": "integer pipe",
                      "fp64": "FP64 pipe", "dram": "HBM"}
        path_alg = 19 * npx * n        # barrier-aware four-pass minimum of the enhance + analysis chain
        roofline = {"kernel": dom, "bound": ("hbm" if busiest is None or busiest[0] == "dram" else busiest[0]),
                    "achieved": d["achieved_gbs"], "peak": peak, "unit": "GB/s",
                    "frac": d["frac_of_hbm_peak"], "traffic": d["traffic_bytes_per_launch"],
                    "traffic_source": traffic.get("source"), "peak_source": peak_src,
                    "binding_pipe": None if busiest is None else {
                        "name": pipe_names.get(busiest[0], busiest[0]), "utilisation_pct": busiest[1], "all_pct": pipes,
                        "shared_wavefronts_that_are_conflict_replays": traffic.get("shared_conflict_share", {}).get(dom),
                        "source": "ncu --set full of the same kernel (profiles/, see traffic_source)"},
                    "path_frac_19N": {"alg_bytes_per_step": path_alg, "achieved_gbs": path_alg / (ms_dev / args.steps) / 1e6,
                                      "frac": path_alg / (ms_dev / args.steps) / 1e6 / peak,
                                      "what": "whole step against the 19 N bytes per frame an HBM-bound chain would move"},
                    "note": "achieved / frac are the dominant kernel's ALGORITHMIC HBM bytes over its duration, as the contract "
                            "defines them; that kernel (49-tap bilateral + lighting + sharpen) is not HBM-bound: `bound` names "
                            "the pipe ncu shows busiest; tools/ubench_taps.cu and tools/ubench_bilateral.cu (profiles/"
                            "r02_notes.md section 1) time the tap: 10.4 clk per warp and SM sub-partition for its 7 "
                            "instructions in isolation, 12.5 inside the stage in every exact encoding, i.e. the bilateral "
                            "stage alone is ~37 us per 1080p frame; the per-stage table gives the streaming kernels' HBM "
                            "fractions", "stages": stages}
        fp32_ops = 49 * 8 * npx * n       # 49 taps x (vabsdiff, LUT load, mul, add, 3 fma, convert) per pixel, lower bound
        roofline["fp32_pipe"] = {"ops_per_launch_lower_bound": fp32_ops,
                                 "achieved_tops": fp32_ops / (stages["k_fused"]["ms_per_launch"] / 1e3) / 1e12
                                 if "k_fused" in stages else None,
                                 "peak_tflops_fp32": 148 * 128 * 2 * 1.965e9 / 1e12}
        out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "u8 (f32 bilateral / CLAHE blend, f64 Otsu and warp coordinates)",
               "data": "synthetic", "config": config, "per_gpu": value / world,
               "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": int(host.nbytes) * world,
                       "d2h_bytes_per_step": int(n * len(rects) * STATS_DTYPE.itemsize + n * 4) * world,
                       "ms_per_step": ms_e2e / args.steps,
                       "frames_per_rank": e2e_shares or [n],
                       "sharding": "equal" if not e2e_shares or len(set(e2e_shares)) == 1 else "N x %d frames per step, each rank's share proportional to the "
                                   "host->device rate it sustained in the warm-up (the GPUs of this box do not share the host "
                                   "path equally); no data-path collective" % n,
                       "api": "Engine.pipeline_submit / pipeline_wait -> cvb_pipeline_submit / cvb_pipeline_wait (pinned host "
                              "frames in, per-square statistics + Otsu thresholds out of EVERY step, one batch in flight ahead "
                              "as in a capture loop; all K steps are submitted and completed inside the timed region)"},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
               "parity": parity, "extras": extras}
        emit(out)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
