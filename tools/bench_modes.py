"""The other BASELINE.json configurations as first-class bench modes (`python bench.py --config ...`):

  enhance480 configs[0]: frame_enhancer full chain (LAB CLAHE + bilateral d=9 + sharpen + Otsu) on ONE synthetic
             640x480 BGR frame -- the reference's own CPU-runnable case: the CPU time is the headline of its
             `cpu_baseline`, the CUDA latency of the same frame stands beside it;
  latency    configs[1]: ONE 1920x1080 BGR frame, full enhance chain + manual-ROI grid extraction to 64 squares (+ the
             per-square statistics) on one B200 -- a latency, so `higher_is_better` is false;
  change64   configs[2]: ChangeDetector / PieceDetector per-square statistics between consecutive 1080p frames, batch 64,
             next to the CPU path the Cython twin runs (src/cython/change_detector_cython.pyx:51-161 calls the same
             cv2 / numpy functions; its timing in the build container is in profiles/r02_cython_twin.txt);
  streams4k  configs[4]: 64 concurrent 3840x2160 camera streams over 8 GPUs = 8 streams per GPU, every stream with its
             own resident ChangeDetector / PieceDetector state; one step = one new frame of every stream.

Each mode prints the same JSON contract as the default mode (configs[3]); `metric` names what is measured."""
import time

import numpy as np

H, W, S = 1080, 1920, 620


def _median_ms(eng, fn, reps, warm):
    for _ in range(warm):
        fn()
    out = []
    for _ in range(reps):
        e0, e1 = eng.event(), eng.event()
        eng.record(e0); fn(); eng.record(e1)
        out.append(eng.elapsed_ms(e0, e1))
    return float(np.median(out)), float(np.min(out)), out


def enhance480(ctx):
    """configs[0]."""
    eng, synth, args = ctx["eng"], ctx["synth"], ctx["args"]
    h, w = 480, 640
    frames = synth.frame_batch(4, h, w, args.kind, 0)
    one = eng.upload(frames[:1])
    enh, gray, binary, otsu = eng.empty((1, h, w, 3)), eng.empty((1, h, w)), eng.empty((1, h, w)), eng.empty((1,), np.int32)
    l0 = eng.launch_count()
    med, mn, _ = _median_ms(eng, lambda: eng.enhance_dev(one, enh, gray, binary, otsu), args.steps, max(3, args.warmup))
    launches = (eng.launch_count() - l0) // (args.steps + max(3, args.warmup))
    host_ms = []
    for i in range(args.steps + 3):
        t0 = time.perf_counter(); eng.enhance(frames[i % 4]); host_ms.append((time.perf_counter() - t0) * 1e3)
    cpu = None
    try:
        import cv2
        from oracle import ref_cv2
        nthreads = cv2.getNumThreads()
        ref_cv2.prepare_analysis(ref_cv2.process_pipeline(frames[0]))
        ts = []
        for i in range(10):
            t0 = time.perf_counter(); ref_cv2.prepare_analysis(ref_cv2.process_pipeline(frames[i % 4])); ts.append((time.perf_counter() - t0) * 1e3)
        cv2.setNumThreads(1)
        t1 = []
        for i in range(5):
            t0 = time.perf_counter(); ref_cv2.prepare_analysis(ref_cv2.process_pipeline(frames[i % 4])); t1.append((time.perf_counter() - t0) * 1e3)
        cv2.setNumThreads(nthreads)
        cpu = {"value": float(np.median(ts)), "unit": "ms/frame", "cores": nthreads, "kind": "port",
               "one_thread_ms": float(np.median(t1)),
               "sample": "10 frames with OpenCV's thread pool, 5 frames with one thread",
               "what": "oracle/ref_cv2.py: process_pipeline + prepare_analysis (frame_enhancer.py:148-181) on cv2 %s" % cv2.__version__}
    except ImportError:
        pass
    return {"metric": "640x480 frame_enhancer full chain (LAB CLAHE + bilateral d=9 + sharpen + normalize + Otsu): latency",
            "value": med, "unit": "ms/frame", "higher_is_better": False, "ms_per_step": med,
            "config": {"workload": "BASELINE.json configs[0]: frame_enhancer full chain (LAB CLAHE + bilateral d=9 + sharpen + "
                                   "Otsu) on one synthetic 640x480 BGR frame on CPU (reference path); the CUDA path on the same frame",
                       "frame_kind": args.kind, "l2_policy": "one 0.9 MB frame: L2-resident by the nature of the configuration"},
            "e2e": {"value": float(np.median(host_ms[3:])), "unit": "ms/frame", "h2d_bytes_per_step": h * w * 3,
                    "d2h_bytes_per_step": h * w * 5 + 4, "api": "Engine.enhance (host frame in; enhanced, gray, mask, T out)"},
            "gpu_launches": int(launches), "cpu_baseline": cpu, "device_ms_min": mn}


def latency(ctx):
    """configs[1]."""
    eng, synth, args = ctx["eng"], ctx["synth"], ctx["args"]
    from chessboard_vision_b200.engine import grid_rects, STATS_DTYPE, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE
    rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
    pts = synth.calib_points(H, W)
    M = eng.get_perspective_transform(pts, [[0, 0], [S, 0], [0, S], [S, S]])
    frames = synth.frame_batch(4, H, W, args.kind, 0)
    host = eng.pinned((1, H, W, 3)); host[0] = frames[0]
    one = eng.upload(frames[:1])
    st = eng.new_state(1, S, S)
    stats = eng.empty((1, len(rects)), STATS_DTYPE); otsu = eng.empty((1,), np.int32)
    cal = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), board_size=S)
    run = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), board_size=S)
    eng.pipeline_dev(one, M, rects, cal, st, stats=stats, otsu_t=otsu)
    eng.synchronize()
    l0 = eng.launch_count()
    dev_med, dev_min, _ = _median_ms(eng, lambda: eng.pipeline_dev(one, M, rects, run, st, stats=stats, otsu_t=otsu),
                                     args.steps, max(3, args.warmup))
    launches = (eng.launch_count() - l0) // (args.steps + max(3, args.warmup))
    host_ms = []
    for i in range(args.steps + 3):
        t0 = time.perf_counter(); eng.pipeline(host, M, rects, run, st); host_ms.append((time.perf_counter() - t0) * 1e3)
    eng.profile(True)
    for _ in range(10):
        eng.pipeline_dev(one, M, rects, run, st, stats=stats, otsu_t=otsu)
    prof = eng.profile_read(); eng.profile(False)
    cpu = None
    try:
        import cv2
        from oracle import ref_cv2
        cd = {}
        ref_cv2.full_frame(frames[0], pts, cd_state=cd)
        ts = []
        for i in range(5):
            t0 = time.perf_counter(); ref_cv2.full_frame(frames[i % 4], pts, cd_state=cd); ts.append((time.perf_counter() - t0) * 1e3)
        cpu = {"value": float(np.median(ts)), "unit": "ms/frame", "cores": cv2.getNumThreads(), "kind": "port",
               "sample": "5 frames, OpenCV's own thread pool (%d threads): the reference's latency configuration" % cv2.getNumThreads(),
               "what": "oracle/ref_cv2.py (reference call sequence on cv2/numpy)"}
    except ImportError:
        pass
    return {"metric": "1080p single-frame latency: full enhance chain + ROI grid extraction to 64 squares + square statistics",
            "value": dev_med, "unit": "ms/frame", "higher_is_better": False, "ms_per_step": dev_med,
            "config": {"workload": "BASELINE.json configs[1]: single 1920x1080 BGR frame, full enhance chain + manual-ROI grid "
                                   "extraction to 64 squares on 1xB200", "frame_kind": args.kind,
                       "l2_policy": "one frame (6.2 MB) is L2-resident between repetitions by the nature of the configuration"},
            "e2e": {"value": float(np.median(host_ms[3:])), "unit": "ms/frame", "h2d_bytes_per_step": H * W * 3,
                    "d2h_bytes_per_step": len(rects) * STATS_DTYPE.itemsize + 4,
                    "api": "Engine.pipeline (pinned host frame in, statistics out), wall clock of the call"},
            "gpu_launches": int(launches), "cpu_baseline": cpu,
            "stages_us": {k: v[0] / v[1] * 1e3 for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
            "device_ms_min": dev_min}


def change64(ctx):
    """configs[2]."""
    eng, synth, args = ctx["eng"], ctx["synth"], ctx["args"]
    from chessboard_vision_b200.engine import grid_rects, STATS_DTYPE, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE
    nb = 64
    rects, keys = grid_rects(S)                       # GridExtractor (linear), as calibrate_sensitivity.py:79,146
    pts = synth.calib_points(H, W)
    M = eng.get_perspective_transform(pts, [[0, 0], [S, 0], [0, S], [S, S]])
    prev = synth.frame_batch(8, H, W, args.kind, 0)
    cur = synth.frame_batch(8, H, W, args.kind, 100)
    d_prev = eng.upload(np.stack([prev[i % 8] for i in range(nb)]))
    d_cur = eng.upload(np.stack([cur[i % 8] for i in range(nb)]))
    warped = eng.empty((nb, S, S, 3)); st = eng.new_state(nb, S, S); d_stats = eng.empty((nb, len(rects)), STATS_DTYPE)
    cal = eng.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE)
    run = eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE)
    eng.warp_dev(d_prev, M, S, warped); eng.squares_dev(warped, rects, cal, st, 0, d_stats)

    def step():
        eng.warp_dev(d_cur, M, S, warped); eng.squares_dev(warped, rects, run, st, 0, d_stats)
    med, mn, _ = _median_ms(eng, step, args.steps, max(3, args.warmup))
    # host-buffer form: 64 warped boards (what GridExtractor.split_board hands to the detectors) in, statistics out
    hb = eng.pinned((nb, S, S, 3)); hb[...] = warped.get()
    t = []
    for i in range(args.steps + 3):
        t0 = time.perf_counter(); eng.squares(hb, rects, run, st); t.append((time.perf_counter() - t0) * 1e3)
    cpu = None
    try:
        import cv2
        from oracle import ref_cv2
        nthreads = cv2.getNumThreads()
        cv2.setNumThreads(1)
        wb = [ref_cv2.warp_image(f, pts)[0] for f in (prev[0], cur[0])]
        sq = [ref_cv2.split_board(w) for w in wb]
        state = {p: (ref_cv2.preprocess_square(s, 5).astype(np.float32), np.full(s.shape[:2], 100, np.float32)) for p, s in sq[0].items()}
        ref = {p: ref_cv2.preprocess_square(s, 5) for p, s in sq[0].items()}
        t0 = time.perf_counter(); reps = 20
        for _ in range(reps):
            for p, s in sq[1].items():
                g = ref_cv2.preprocess_square(s, 5)
                ref_cv2.cd_detect(g, *state[p]); ref_cv2.cd_update(g, *state[p])
                float(np.mean(cv2.absdiff(g, ref[p])))
        per_pair = (time.perf_counter() - t0) / reps
        cv2.setNumThreads(nthreads)
        cpu = {"value": 1.0 / per_pair, "unit": "pairs/s", "cores": 1, "kind": "port",
               "sample": "%d pairs of 64 squares on one core" % reps,
               "what": "the cv2 / numpy calls of ChangeDetector._preprocess + detect_changes_detailed (numeric part) + "
                       "update_all_references and PieceDetector._has_changed per square: what both the Python class and its "
                       "Cython twin execute (profiles/r02_cython_twin.txt has the twin's own timing from the build container)"}
    except ImportError:
        pass
    return {"metric": "change_detector / piece_detector per-square statistics on consecutive 1080p frames (pairs/s, batch 64)",
            "value": nb / med * 1e3, "unit": "pairs/s", "higher_is_better": True, "ms_per_step": med,
            "config": {"workload": "BASELINE.json configs[2]: change_detector per-square absdiff/threshold between consecutive "
                                   "1080p frames, batch 64, vs src/cython path", "pairs_per_step": nb, "frame_kind": args.kind,
                       "step": "warp of 64 current frames to 620x620 + one k_squares launch: gray, 5x5 blur, absdiff delta, "
                               "moments, centre/border, rings, z-score detect + EMA update for 64 x 64 squares",
                       "l2_policy": "inputs larger than L2 (398 MB of frames per step vs 126 MB)"},
            "e2e": {"value": nb / float(np.median(t[3:])) * 1e3, "unit": "pairs/s", "h2d_bytes_per_step": nb * S * S * 3,
                    "d2h_bytes_per_step": nb * len(rects) * STATS_DTYPE.itemsize,
                    "api": "Engine.squares (64 warped boards in pinned host memory in, per-square records out)"},
            "gpu_launches": 2, "cpu_baseline": cpu}


def streams4k(ctx):
    """configs[4]: 8 streams of 3840x2160 per GPU (64 over 8 GPUs), weak scaling, per-stream state resident."""
    eng, synth, args, rank = ctx["eng"], ctx["synth"], ctx["args"], ctx["rank"]
    from chessboard_vision_b200.engine import grid_rects, STATS_DTYPE, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE
    ns, H4, W4 = 8, 2160, 3840
    rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
    M = eng.get_perspective_transform(synth.calib_points(H4, W4), [[0, 0], [S, 0], [0, S], [S, S]])
    uniq = np.stack([synth.frame_batch(1, H4, W4, args.kind, 40 + 2 * rank + i)[0] for i in range(2)])
    host = eng.pinned((ns, H4, W4, 3))
    for i in range(ns):
        host[i] = uniq[i % 2]
    d_in = eng.upload(host)
    st = eng.new_state(ns, S, S); d_stats = eng.empty((ns, len(rects)), STATS_DTYPE); d_otsu = eng.empty((ns,), np.int32)
    cal = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), board_size=S)
    run = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), board_size=S)
    eng.pipeline_dev(d_in, M, rects, cal, st, stats=d_stats, otsu_t=d_otsu)
    eng.synchronize()

    def run_dev(k):
        for _ in range(k):
            eng.pipeline_dev(d_in, M, rects, run, st, stats=d_stats, otsu_t=d_otsu)

    def run_e2e(k):        # a capture loop: one step (a new frame of every stream) in flight ahead of the one being collected
        pending = []
        for _ in range(k):
            pending.append(eng.pipeline_submit(host, M, rects, run, st))
            if len(pending) > 1:
                eng.pipeline_wait(pending.pop(0))
        while pending:
            eng.pipeline_wait(pending.pop(0))
    timed = ctx["timed"]
    run_dev(max(3, args.warmup))
    ms_dev, launches = timed(run_dev, args.steps)
    run_e2e(max(3, args.warmup))
    ms_e2e, _ = timed(run_e2e, args.steps)
    world = ctx["world"]
    return {"metric": "3840x2160 frames/s over concurrent camera streams, full pipeline with per-stream state",
            "value": ns * args.steps * world / (ms_dev / 1e3), "unit": "frames/s", "higher_is_better": True,
            "ms_per_step": ms_dev / args.steps,
            "config": {"workload": "BASELINE.json configs[4]: 64 concurrent 3840x2160 camera streams, full pipeline, sharded "
                                   "across 8xB200 = 8 streams per GPU (this run: %d GPU(s), %d streams)" % (world, ns * world),
                       "streams_per_gpu": ns, "frame_kind": args.kind, "step": "one new frame of every stream",
                       "parallelism": "streams sharded over %d GPU(s), state resident per stream, no collective" % world,
                       "l2_policy": "inputs larger than L2 (199 MB per step per GPU vs 126 MB)"},
            "e2e": {"value": ns * args.steps * world / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(host.nbytes),
                    "d2h_bytes_per_step": int(ns * len(rects) * STATS_DTYPE.itemsize + ns * 4), "ms_per_step": ms_e2e / args.steps,
                    "api": "Engine.pipeline_submit / pipeline_wait (pinned host frames in, per-square statistics + Otsu thresholds out "
                           "of every step, one step in flight ahead)"},
            "gpu_launches": int(launches), "mpixels_per_s": ns * args.steps * world * H4 * W4 / (ms_dev / 1e3) / 1e6}


MODES = {"enhance480": enhance480, "latency": latency, "change64": change64, "streams4k": streams4k}
