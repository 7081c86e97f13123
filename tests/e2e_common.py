"""Shared body of the end-to-end parity tests: the drop-in classes run the reference's own call sequence
(process_pipeline -> prepare_analysis -> warp_image -> split_board -> ChangeDetector / PieceDetector) on a
(previous, current) pair and the result is compared with what the UNMODIFIED reference produced
(tests/golden/e2e_reference.npz, tools/make_golden_e2e.py): Otsu thresholds, masks pixel by pixel, change flags."""
import json
import os

import numpy as np

from chessboard_vision_b200 import synth
from oracle.parity import change_blocks

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Mask pixels that may differ from the reference, per (kind, W, H, seed): the measured values of the bilateral
# filter's <= 1 LSB class (OpenCV's own IPP / SIMD / scalar paths differ from each other at this rate).
# The CUDA path is bit-identical to the CPU oracle, so these bounds are exact, not statistical.
MASK_PX_BOUND = {("noise", 1920, 1080, 0): 1, ("board", 1920, 1080, 0): 0, ("board", 1920, 1080, 5): 0,
                 ("noise", 640, 480, 0): 0, ("board", 1280, 720, 3): 0}


def load_cases():
    z = np.load(os.path.join(G, "e2e_reference.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return z, meta["cases"]


def run_case(mods, z, rec):
    """-> dict of measured differences against the reference for one case."""
    kind, H, W, seed = rec["kind"], rec["H"], rec["W"], rec["seed"]
    prev = (synth.board_frame if kind == "board" else synth.noise_frame)(H, W, seed)
    cur = synth.change_pair(prev, change_blocks(H, W), 255)
    enh = mods["frame_enhancer"].ImageEnhancer()
    grid = mods["grid_extractor"].SmartGridExtractor()
    grid.grid_lines_x, grid.grid_lines_y = list(synth.CALIB_GRID_X), list(synth.CALIB_GRID_Y)
    pts = synth.calib_points(H, W)
    out = {"mask_px": [], "t_ref": []}
    squares = []
    for tag, frame in (("prev", prev), ("cur", cur)):
        e = enh.process_pipeline(frame)
        gray, binary = enh.prepare_analysis(e)
        ref_mask = np.unpackbits(z["%s_%dx%d_%d/%s_mask_bits" % (kind, W, H, seed, tag)])[:H * W].reshape(H, W)
        assert set(np.unique(binary)) <= {0, 255}
        out["mask_px"].append(int(np.count_nonzero((binary > 0) != (ref_mask > 0))))
        out["t_ref"].append(rec[tag]["otsu_t"])
        # (a different Otsu threshold would change thousands of mask pixels, so the mask comparison covers T)
        warped, _, S = mods["board_detection"].warp_image(e, pts)
        squares.append(grid.split_board(warped))
    cd = mods["change_detector"].ChangeDetector()
    cd.calibrate(squares[0])
    pd = cd.piece_detector
    pd.calibrate_reference(squares[0])
    flags = {"%d_%d" % p: bool(pd._has_changed(p, pd._preprocess_square(sq))) for p, sq in squares[1].items()}
    det = cd.detect_changes_detailed(squares[1])
    out["has_changed_mismatch"] = sum(flags[k] != v for k, v in rec["has_changed"].items())
    got = {"%d_%d" % p: d for p, d in det.items()}
    out["detailed_keys_equal"] = sorted(got) == sorted(rec["detailed"])
    out["intensity_mismatch"] = sum(1 for k in rec["detailed"] if k not in got or got[k]["intensity"] != rec["detailed"][k]["intensity"])
    out["is_circular_mismatch"] = sum(1 for k in rec["detailed"] if k in got and bool(got[k]["is_circular"]) != rec["detailed"][k]["is_circular"])
    out["pct_abs_diff_max"] = max([abs(got[k]["pct_changed"] - rec["detailed"][k]["pct_changed"]) for k in rec["detailed"] if k in got] or [0.0])
    ch = {"%d_%d" % p: v for p, v in cd.detect_changes(squares[1]).items()}
    out["changes_keys_equal"] = sorted(ch) == sorted(rec["changes"])
    return out


def assert_case(rec, out):
    key = (rec["kind"], rec["W"], rec["H"], rec["seed"])
    assert max(out["mask_px"]) <= MASK_PX_BOUND[key], "mask differs from the reference on %s pixels (%s)" % (out["mask_px"], key)
    assert out["has_changed_mismatch"] == 0 and out["detailed_keys_equal"] and out["intensity_mismatch"] == 0, out
    assert out["is_circular_mismatch"] == 0 and out["changes_keys_equal"], out
    assert out["pct_abs_diff_max"] == 0.0, out           # measured: the changed-pixel counts are identical on these pairs
