"""Randomised parity sweep on the GPU: random frame sizes and contents through the enhancer chain, the analysis tail,
the warp, the square statistics, the YUV ingest and random overlay display lists, each compared bit for bit with the CPU
oracle (test infrastructure).
usage: python tools/fuzz_parity.py [cases] [seed]      (tests/test_gpu_e2e_parity.py runs `sweep` under pytest)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import oracle as O
import overlay_common
from oracle import overlay as OV
from chessboard_vision_b200 import synth
from chessboard_vision_b200.engine import grid_rects, SQ_PD_STATS, SQ_CD_CALIBRATE, SQ_CD_DETECT, SQ_CD_UPDATE



def sweep(eng, cases=40, seed=0, verbose=True):
    """-> number of mismatching cases."""
    rng = np.random.default_rng(seed)
    bad = 0
    for c in range(cases):
        if c % 4 == 0:      # sizes that take the aligned fast paths
            H, W = int(rng.integers(2, 12)) * 36, int(rng.integers(1, 5)) * 120
        elif c % 4 == 1:    # row pitch a multiple of 16 bytes (the tensor-map path of the fused kernel, when selected)
            H, W = int(rng.integers(16, 300)), int(rng.integers(1, 30)) * 16
        else:
            H, W = int(rng.integers(9, 420)), int(rng.integers(9, 520))
        kind = rng.choice(["board", "noise", "flat", "ramp"])
        if kind == "board":
            f = synth.board_frame(H, W, int(rng.integers(0, 1000)))
        elif kind == "noise":
            f = synth.noise_frame(H, W, int(rng.integers(0, 1000)))
        elif kind == "flat":
            f = np.full((H, W, 3), int(rng.integers(0, 256)), np.uint8)
        else:
            f = np.broadcast_to((np.arange(W) * 255 // max(W - 1, 1)).astype(np.uint8)[None, :, None], (H, W, 3)).copy()
        ok = True
        enh = eng.process_pipeline(f)
        ref = O.process_pipeline(f, True)
        ok &= np.array_equal(enh, ref)
        g, b, T = eng.prepare_analysis(f, return_all=True)[:3]
        rg, rb, rT, _ = O.prepare_analysis(f, True)
        ok &= np.array_equal(g, rg) and np.array_equal(b, rb) and int(T) == int(rT)
        S = int(rng.integers(16, 200))
        pts = np.float32([[0, 0], [W, 0], [0, H], [W, H]]) + rng.uniform(-0.2, 0.2, (4, 2)).astype(np.float32) * [W, H]
        M = eng.get_perspective_transform(pts, [[0, 0], [S, 0], [0, S], [S, S]])
        flip = bool(rng.integers(0, 2))
        w = eng.warp(f, M, S, rotate_180=flip)
        rw = O.warp(f, M, S)
        ok &= np.array_equal(w, O.rotate(rw, 1) if flip else rw)
        rects, _ = grid_rects(S)
        rects = [r for r in rects if r[2] > 0 and r[3] > 0]
        if rects:
            st = eng.new_state(1, S, S)
            eng.squares(w, rects, eng.square_params(ops=SQ_PD_STATS | SQ_CD_CALIBRATE), st)
            w2 = np.clip(w.astype(np.int16) + rng.integers(-30, 31, w.shape), 0, 255).astype(np.uint8)
            stats = eng.squares(w2, rects, eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), st)
            for j, (x, y, ww, hh) in enumerate(rects):
                g2 = O.square_preprocess(w2[y:y + hh, x:x + ww], 5)
                o = O.pd_square_stats(g2)
                g1 = O.square_preprocess(w[y:y + hh, x:x + ww], 5)
                m, v = O.cd_calibrate(g1, 100.0)
                cnt, zmax = O.cd_detect(g2, m, v, 2.5)
                s = stats[0, j]
                ok &= int(s["sum"]) == o["sum"] and int(s["sumsq"]) == o["sumsq"] and int(s["cd_changed"]) == cnt
                ok &= (np.float32(s["cd_zmax"]) == np.float32(zmax)) or (np.isnan(s["cd_zmax"]) and np.isnan(zmax))
            st.free()
        if H % 2 == 0 and W % 2 == 0:      # camera formats: device conversion against the integer BT.601 of cv2.cvtColor
            for fmt in ("yuy2", "nv12"):
                raw = synth.bgr_to_yuv(f, fmt)
                ok &= np.array_equal(eng.cvt_to_bgr(raw, fmt), O.yuv_to_bgr(raw, fmt))
        dl = overlay_common.random_display_list(rng, S, S, int(rng.integers(1, 60)))
        ops, n_ops, masks = dl.pack()
        ok &= np.array_equal(eng.overlay(w, ops, n_ops, masks), OV.apply_display_list(w, ops, n_ops, masks))
        if not ok:
            bad += 1
            print("MISMATCH case", c, (H, W), kind, "S", S, "flip", flip)
    if verbose:
        print("fuzz: %d cases, %d mismatching" % (cases, bad))
    return bad


if __name__ == "__main__":
    from chessboard_vision_b200.engine import Engine
    n_bad = sweep(Engine(0), int(sys.argv[1]) if len(sys.argv) > 1 else 40, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    sys.exit(1 if n_bad else 0)
