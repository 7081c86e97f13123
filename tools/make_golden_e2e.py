"""Generate tests/golden/e2e_reference.npz by running the UNMODIFIED reference modules over the WHOLE chain
(process_pipeline -> prepare_analysis -> warp_image -> split_board -> ChangeDetector / PieceDetector) on
(previous, current) frame pairs.  Build container only (needs /root/reference):

    python tools/make_golden_e2e.py

Stored per pair: Otsu thresholds, the two binary masks bit-packed (so that the GPU test can compare PIXELS, not
counts), the sha of the enhanced frames, the `_has_changed` flag of every square, and the dictionary
`ChangeDetector.detect_changes_detailed` returns (pct_changed, intensity, z_score, is_circular) plus
`detect_changes`.  cwd is an empty temp dir: no colour profile, default PieceDetector settings.
"""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

CASES = [("noise", 1080, 1920, 0), ("board", 1080, 1920, 0), ("board", 1080, 1920, 5), ("noise", 480, 640, 0),
         ("board", 720, 1280, 3)]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    os.chdir(tempfile.mkdtemp(prefix="golden_cwd_"))
    sys.path.insert(0, REF)
    import cv2
    import frame_enhancer as fe
    import grid_extractor as ge
    import board_detection as bd
    import change_detector as cdm
    import piece_detector as pdm
    from chessboard_vision_b200 import synth
    from oracle.parity import change_blocks

    out, meta = {}, {"cv2": cv2.__version__, "numpy": np.__version__, "ipp": bool(cv2.ipp.useIPP()), "cases": []}
    for kind, H, W, seed in CASES:
        prev = (synth.board_frame if kind == "board" else synth.noise_frame)(H, W, seed)
        cur = synth.change_pair(prev, change_blocks(H, W), 255)
        enh = fe.ImageEnhancer()
        assert enh.profile == {}
        grid = ge.SmartGridExtractor()
        grid.grid_lines_x, grid.grid_lines_y = list(synth.CALIB_GRID_X), list(synth.CALIB_GRID_Y)
        pts = synth.calib_points(H, W)
        rec = {"kind": kind, "H": H, "W": W, "seed": seed}
        squares = []
        for tag, frame in (("prev", prev), ("cur", cur)):
            e = enh.process_pipeline(frame)
            gray, binary = enh.prepare_analysis(e)
            t, _ = cv2.threshold(cv2.GaussianBlur(gray, (5, 5), 0), 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
            warped, _, S = bd.warp_image(e, pts)
            squares.append(grid.split_board(warped))
            name = "%s_%dx%d_%d/%s" % (kind, W, H, seed, tag)
            out[name + "_mask_bits"] = np.packbits(binary > 0)
            rec[tag] = {"otsu_t": int(t), "white_px": int(np.count_nonzero(binary)), "enhanced_sha": sha(e),
                        "warped_sha": sha(warped), "input_sha": sha(frame)}
        cd = cdm.ChangeDetector()
        cd.calibrate(squares[0])
        pd = cd.piece_detector
        pd.calibrate_reference(squares[0])
        rec["has_changed"] = {"%d_%d" % p: bool(pd._has_changed(p, pd._preprocess_square(sq))) for p, sq in squares[1].items()}
        det = cd.detect_changes_detailed(squares[1])
        rec["detailed"] = {"%d_%d" % p: {"pct_changed": float(d["pct_changed"]), "intensity": d["intensity"],
                                         "z_score": float(d["z_score"]), "is_circular": bool(d["is_circular"])}
                           for p, d in det.items()}
        rec["changes"] = {"%d_%d" % p: float(v) for p, v in cd.detect_changes(squares[1]).items()}
        meta["cases"].append(rec)
    out["meta"] = np.frombuffer(json.dumps(meta).encode(), np.uint8)
    path = os.path.join(OUT, "e2e_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote %s (%d bytes)" % (path, os.path.getsize(path)))
    for r in meta["cases"]:
        print(r["kind"], r["W"], r["H"], r["seed"], "T", r["prev"]["otsu_t"], r["cur"]["otsu_t"], "changed squares",
              sum(r["has_changed"].values()), {k: v["intensity"] for k, v in r["detailed"].items()})


if __name__ == "__main__":
    main()
