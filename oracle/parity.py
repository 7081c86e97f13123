"""CHECKER (test infrastructure, not product code): end-to-end parity of the CUDA path against the reference's
own cv2 / numpy call sequence (oracle/ref_cv2.py) on (previous, current) frame pairs.

What is compared, per pair, is what `north_star` names: the Otsu threshold, the binary mask PIXEL BY PIXEL, the
PieceDetector change flags (`_has_changed`, piece_detector.py:82-93; std gate :305), the ChangeDetector changed-pixel
counts and LEVE / PARCIAL / TOTAL classes (change_detector.py:121-150), and the enhanced frame itself (whose only
admissible difference is the bilateral filter's <= 1 LSB, amplified at most 9x by the sharpen kernel: SURVEY.md 0.5).
Used by `bench.py` (the `parity` block of the JSON line) and by tests/test_gpu_e2e_parity.py.
"""
import numpy as np

S = 620


def change_blocks(H, W):
    """Axis-aligned blocks inside the calibrated ROI (frame coordinates) that turn a few squares into
    TOTAL / PARCIAL / LEVE changes: (x, y, w, h), scaled from 1080p."""
    sx, sy = W / 1920.0, H / 1080.0
    blk = [(672, 216, 146, 130), (1000, 500, 60, 55), (1300, 800, 36, 30), (900, 700, 100, 20)]
    return [(int(x * sx), int(y * sy), max(1, int(w * sx)), max(1, int(h * sy))) for x, y, w, h in blk]


def cd_class(changed, n):
    """change_detector.py:141-150 -> None (below 5 %), 'LEVE', 'PARCIAL' or 'TOTAL'."""
    pct = (changed / n) * 100
    if pct < 5.0:
        return None
    return "TOTAL" if pct > 75 else ("PARCIAL" if pct > 15 else "LEVE")


def cpu_pair(prev, cur, pts, grid):
    """The reference's call sequence on the pair -> per-frame masks / thresholds and per-square results of `cur`."""
    from oracle import ref_cv2
    cd, ka, kb = {}, {}, {}
    ta, _ = ref_cv2.full_frame(prev, pts, cd_state=cd, pd_ref=None, grid=grid, keep=ka)
    tb, res = ref_cv2.full_frame(cur, pts, cd_state=cd, pd_ref=ka["gray_squares"], grid=grid, keep=kb)
    sq = {}
    for pos, st in res.items():
        n = kb["gray_squares"][pos].size
        sq[pos] = {"has_changed": st["mean_diff"] > 25, "low_std": st["std"] < 15, "cd_changed": st["cd"][0],
                   "cd_class": cd_class(st["cd"][0], n), "cb_flag": st["center_border_diff"] > 40,
                   "sym_flag": st["symmetry"] > 0.6}
    return {"t": [int(ta), int(tb)], "mask": [ka["binary"], kb["binary"]], "enhanced": [ka["enhanced"], kb["enhanced"]],
            "squares": sq}


def gpu_pair(eng, prev, cur, pts, grid):
    """The same pair through the C ABI (host-buffer entry points)."""
    from chessboard_vision_b200.engine import (grid_rects, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT,
                                               SQ_CD_UPDATE)
    rects, keys = grid_rects(S, *grid) if grid[0] is not None else grid_rects(S)
    M = eng.get_perspective_transform(pts, [[0, 0], [S, 0], [0, S], [S, S]])
    st = eng.new_state(1, S, S)
    cal = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), board_size=S)
    run = eng.pipeline_params(squares=eng.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), board_size=S)
    eng.pipeline(prev[None], M, rects, cal, st)          # calibrate(): model + references from the first frame
    eng.pipeline(prev[None], M, rects, run, st)          # detect + update on it, as full_frame does
    _, stats = eng.pipeline(cur[None], M, rects, run, st)
    ea, _, ba, ta = eng.enhance(prev)
    eb, _, bb, tb = eng.enhance(cur)
    st.free()
    sq = {}
    for j, pos in enumerate(keys):
        s = stats[0, j]
        n = int(s["n"])
        cm = int(s["center_sum"]) / max(1, int(s["center_cnt"])); bm = int(s["border_sum"]) / max(1, int(s["border_cnt"]))
        rings = [int(s["ring_sum"][k]) / int(s["ring_cnt"][k]) for k in range(4) if int(s["ring_cnt"][k]) > 0]
        sym = 0.0 if len(rings) < 2 else min(1.0, float(np.var(rings)) / 500)
        sq[pos] = {"has_changed": int(s["sad"]) > 25 * n,
                   "low_std": n * int(s["sumsq"]) - int(s["sum"]) ** 2 < 225 * n * n,
                   "cd_changed": int(s["cd_changed"]), "cd_class": cd_class(int(s["cd_changed"]), n),
                   "cb_flag": abs(cm - bm) > 40, "sym_flag": sym > 0.6}
    return {"t": [int(ta), int(tb)], "mask": [ba, bb], "enhanced": [ea, eb], "squares": sq}


def compare(cpu, gpu):
    """-> dict of per-pair mismatch counts (all zeros and `otsu_t_equal` True = identical flags and masks)."""
    out = {"otsu_t": {"reference": cpu["t"], "cuda": gpu["t"]}, "otsu_t_equal": cpu["t"] == gpu["t"],
           "mask_px_diff": [int(np.count_nonzero(a != b)) for a, b in zip(cpu["mask"], gpu["mask"])],
           "enhanced_values_diff": [int(np.count_nonzero(a != b)) for a, b in zip(cpu["enhanced"], gpu["enhanced"])],
           "enhanced_max_abs_diff": max(int(np.abs(a.astype(np.int16) - b).max()) for a, b in zip(cpu["enhanced"], gpu["enhanced"]))}
    for key in ("has_changed", "low_std", "cd_class", "cb_flag", "sym_flag"):
        out[key + "_mismatch"] = sum(cpu["squares"][p][key] != gpu["squares"][p][key] for p in cpu["squares"])
    out["cd_changed_px_diff_max"] = max(abs(cpu["squares"][p]["cd_changed"] - gpu["squares"][p]["cd_changed"])
                                        for p in cpu["squares"])
    out["squares"] = len(cpu["squares"])
    out["squares_changed_ref"] = sum(1 for p in cpu["squares"] if cpu["squares"][p]["has_changed"])
    out["cd_classes_ref"] = {c: sum(1 for p in cpu["squares"] if cpu["squares"][p]["cd_class"] == c)
                             for c in ("LEVE", "PARCIAL", "TOTAL")}
    return out


def run(eng, frames, H, W, grid=None):
    """Parity of `frames` (list of BGR frames): each frame is paired with a copy carrying the change blocks."""
    from chessboard_vision_b200 import synth
    pts = synth.calib_points(H, W)
    grid = grid or (list(synth.CALIB_GRID_X), list(synth.CALIB_GRID_Y))
    pairs = []
    for f in frames:
        cur = synth.change_pair(f, change_blocks(H, W), 255)
        pairs.append(compare(cpu_pair(f, cur, pts, grid), gpu_pair(eng, f, cur, pts, grid)))
    agg = {"pairs": len(pairs), "frames": 2 * len(pairs),
           "otsu_t_equal": all(p["otsu_t_equal"] for p in pairs),
           "mask_px_diff_max": max(max(p["mask_px_diff"]) for p in pairs),
           "mask_px_diff": [p["mask_px_diff"] for p in pairs],
           "enhanced_values_diff_max": max(max(p["enhanced_values_diff"]) for p in pairs),
           "enhanced_max_abs_diff": max(p["enhanced_max_abs_diff"] for p in pairs),
           "flags_equal": all(p[k + "_mismatch"] == 0 for p in pairs for k in ("has_changed", "low_std", "cd_class", "cb_flag", "sym_flag")),
           "cd_changed_px_diff_max": max(p["cd_changed_px_diff_max"] for p in pairs),
           "squares_changed_ref": [p["squares_changed_ref"] for p in pairs],
           "cd_classes_ref": [p["cd_classes_ref"] for p in pairs]}
    for k in ("has_changed", "low_std", "cd_class", "cb_flag", "sym_flag"):
        agg[k + "_mismatch"] = sum(p[k + "_mismatch"] for p in pairs)
    return agg
