"""Generate tests/golden/* by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, cv2, numpy):

    python tools/make_golden.py

The reference is imported from /root/reference with the working directory set
to an empty temp dir, so neither color_profile.json nor
piece_detector_settings.json is picked up (colour profile off, default radii),
exactly the oracle configuration SURVEY.md 8c defines.  Outputs are small .npz
/ .json fixtures; large arrays are stored as sha256 digests.
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


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    os.makedirs(OUT, exist_ok=True)
    os.chdir(tempfile.mkdtemp(prefix="golden_cwd_"))
    sys.path.insert(0, REF)
    import cv2
    import frame_enhancer as fe          # falls back to ImageEnhancerPython (no built .so in the read-only tree)
    import grid_extractor as ge
    import board_detection as bd
    import change_detector as cdm
    import piece_detector as pdm
    from chessboard_vision_b200 import synth

    enh = fe.ImageEnhancer()
    assert enh.profile == {}, "colour profile must be off for the golden run"
    meta = {"cv2": cv2.__version__, "numpy": np.__version__, "enhancer_class": type(enh).__name__,
            "ipp": bool(cv2.ipp.useIPP()), "reference": "hericmr/chessboard-vision @ /root/reference"}

    # ---- 1. enhancer, stage-isolated, small frames ------------------------------------------------
    small = {}
    for name, img in (("board_96x128", synth.board_frame(96, 128, 1)), ("noise_90x121", synth.noise_frame(90, 121, 2)),
                      ("board_120x160", synth.board_frame(120, 160, 5))):
        lab = cv2.cvtColor(img, cv2.COLOR_BGR2LAB)
        l = np.ascontiguousarray(lab[..., 0])
        lit = enh.correct_lighting(img)
        bil = enh.reduce_noise(lit)
        shp = enh.sharpen(bil)
        nrm = enh.normalize_intensity(shp)
        full = enh.process_pipeline(img)
        assert np.array_equal(full, nrm)
        gray, binary = enh.prepare_analysis(full)
        blur = cv2.GaussianBlur(gray, (5, 5), 0)
        t, _ = cv2.threshold(blur, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        small.update({name + "/input": img, name + "/lab": lab, name + "/clahe_l": enh.clahe.apply(l),
                      name + "/correct_lighting": lit, name + "/reduce_noise": bil, name + "/sharpen": shp,
                      name + "/normalize": nrm, name + "/gray": gray, name + "/blur": blur, name + "/binary": binary,
                      name + "/otsu_t": np.array(int(t))})
    np.savez_compressed(os.path.join(OUT, "enhancer_small.npz"), **small)

    # ---- 2. known-answer digests at full sizes (SURVEY.md 8c probe 16) ---------------------------
    kat = {}
    for (H, W) in ((480, 640), (1080, 1920)):
        img = synth.noise_frame(H, W, 0)
        lab = cv2.cvtColor(img, cv2.COLOR_BGR2LAB)
        l, a, b = cv2.split(lab)
        l2 = enh.clahe.apply(l)
        lit = cv2.cvtColor(cv2.merge((l2, a, b)), cv2.COLOR_LAB2BGR)
        assert np.array_equal(lit, enh.correct_lighting(img))
        bil = enh.reduce_noise(lit)
        shp = enh.sharpen(bil)
        nrm = enh.normalize_intensity(shp)
        gray, binary = enh.prepare_analysis(nrm)
        blur = cv2.GaussianBlur(gray, (5, 5), 0)
        t, _ = cv2.threshold(blur, 0, 255, cv2.THRESH_BINARY + cv2.THRESH_OTSU)
        pts = synth.calib_points(H, W)
        warped, M, S = bd.warp_image(img, pts)
        key = "%dx%d" % (W, H)
        kat[key] = {"input": sha(img), "lab": sha(lab), "clahe_l": sha(l2), "correct_lighting": sha(lit),
                    "reduce_noise_ipp": sha(bil),
                    # stages after the bilateral filter, each fed the reference's own bilateral output
                    "sharpen_of_ref_bilateral": sha(shp), "normalize_of_ref": sha(nrm), "gray_of_ref": sha(gray),
                    "blur_of_ref": sha(blur), "binary_of_ref": sha(binary), "otsu_t_of_ref": int(t),
                    "white_px_of_ref": int(np.count_nonzero(binary)),
                    "tile00_hist_0_4": np.bincount(l[:H // 8, :W // 8].ravel(), minlength=256)[:4].tolist(),
                    "warp": sha(warped), "warp_matrix": [float(v) for v in M.ravel()], "board_size": int(S)}
        if (H, W) == (480, 640):
            # the reference's bilateral output itself, so that later stages can be replayed bit-exactly
            np.savez_compressed(os.path.join(OUT, "bilateral_640x480.npz"), reduce_noise=bil)
    json.dump({"meta": meta, "kat": kat}, open(os.path.join(OUT, "kat.json"), "w"), indent=1)

    # ---- 3. grid extractor: key order and rectangles ---------------------------------------------
    board = synth.board_frame(620, 620, 7)
    lin = ge.GridExtractor().split_board(board)
    sg = ge.SmartGridExtractor()
    sg.grid_lines_x, sg.grid_lines_y = list(synth.CALIB_GRID_X), list(synth.CALIB_GRID_Y)
    smart = sg.split_board(board)

    def rect_of(view):
        off = view.__array_interface__["data"][0] - board.__array_interface__["data"][0]
        y, rem = divmod(off, board.strides[0])
        return [int(rem // 3), int(y), int(view.shape[1]), int(view.shape[0])]
    grid = {"linear": [[list(k), rect_of(v)] for k, v in lin.items()],
            "smart": [[list(k), rect_of(v)] for k, v in smart.items()],
            "smart_unset_falls_back": [list(k) for k in ge.SmartGridExtractor().split_board(board).keys()] ==
                                      [list(k) for k in lin.keys()]}
    json.dump(grid, open(os.path.join(OUT, "grid.json"), "w"))

    # ---- 4. piece detector statistics on 64 squares with some discs ------------------------------
    board_chk, boardp = synth.board_with_pieces(11, 7, 620)
    assert np.array_equal(board_chk, board)
    pd = pdm.PieceDetector()
    sq_ref = ge.GridExtractor().split_board(board)
    sq_cur = ge.GridExtractor().split_board(boardp)
    pd_out = {}
    gray_sha = {}
    for pos in sq_cur:
        pd.reference_squares[pos] = pd._preprocess_square(sq_ref[pos]).copy()
    rows = []
    for pos, sq in sq_cur.items():
        g = pd._preprocess_square(sq)
        diff, cm, bm = pd._detect_center_vs_border(g)
        rows.append([pos[0], pos[1], float(np.mean(cv2.absdiff(g, pd.reference_squares[pos]))),
                     float(pd._has_changed(pos, g)), float(np.std(g)), float(diff), float(cm), float(bm),
                     float(pd._analyze_radial_symmetry(g))])
        gray_sha["%d_%d" % pos] = sha(g)
    pd_out["stats"] = np.array(rows, np.float64)
    pd_out["gray_sha"] = np.array(json.dumps(gray_sha))
    np.savez_compressed(os.path.join(OUT, "piece_detector.npz"), **pd_out)

    # ---- 5. change detector: calibrate / detect / update -----------------------------------------
    cd = cdm.ChangeDetector()
    cd.calibrate(sq_ref)
    detailed = cd.detect_changes_detailed(sq_cur)
    changes = cd.detect_changes(sq_cur)
    cd_out = {"means_sha": {("%d_%d" % p): sha(v) for p, v in cd.means.items()},
              "detailed": {("%d_%d" % p): {k: (v if not isinstance(v, (np.bool_, bool)) else bool(v)) for k, v in d.items()}
                           for p, d in detailed.items()},
              "changes": {("%d_%d" % p): float(v) for p, v in changes.items()},
              "class": type(cd).__name__}
    cd.update_all_references(sq_cur)
    cd_out["means_after_update_sha"] = {("%d_%d" % p): sha(v) for p, v in cd.means.items()}
    cd_out["vars_after_update_sha"] = {("%d_%d" % p): sha(v) for p, v in cd.variances.items()}
    # non-default parameters (sensitivity_settings.json values: blur 13, z 2.55, alpha 0.13, var 600)
    cd2 = cdm.ChangeDetector()
    cd2.blur_kernel, cd2.z_threshold, cd2.alpha, cd2.initial_variance = 13, 2.55, 0.13, 600
    cd2.calibrate(sq_ref)
    cd2.set_focus_squares([(0, 7), (3, 3), (4, 4), (7, 0)])
    cd2.update_all_references(sq_cur)
    det2 = cd2.detect_changes_detailed(sq_cur)
    cd_out["tuned"] = {"means_sha": {("%d_%d" % p): sha(v) for p, v in cd2.means.items()},
                       "vars_sha": {("%d_%d" % p): sha(v) for p, v in cd2.variances.items()},
                       "detailed": {("%d_%d" % p): {"pct_changed": d["pct_changed"], "z_score": d["z_score"],
                                                    "intensity": d["intensity"]} for p, d in det2.items()}}
    # the reference's own behavioural test (test_change_detector_regression.py:31-54)
    det = cdm.ChangeDetector()
    z = {(c, r): np.zeros((50, 50), np.uint8) for r in range(8) for c in range(8)}
    det.calibrate(z)
    z[(3, 3)] = np.full((50, 50), 255, np.uint8)
    d3 = det.detect_changes_detailed(z)
    cd_out["regression_3_3"] = {"pct_changed": d3[(3, 3)]["pct_changed"], "z_score": d3[(3, 3)]["z_score"],
                                "intensity": d3[(3, 3)]["intensity"], "n_detected": len(d3)}
    z2 = {(c, r): np.zeros((77, 77), np.uint8) for r in range(8) for c in range(8)}
    det = cdm.ChangeDetector(); det.calibrate(z2)
    half = np.zeros((77, 77), np.uint8); half[:40] = 255
    z2[(1, 1)] = half
    d4 = det.detect_changes_detailed(z2)
    cd_out["top40rows_1_1"] = {"pct_changed": d4[(1, 1)]["pct_changed"], "intensity": d4[(1, 1)]["intensity"]}
    json.dump(cd_out, open(os.path.join(OUT, "change_detector.json"), "w"), indent=1)

    # ---- 7. colour profile (step 0 of process_pipeline; "next" scope row) ----------------------------
    profiles = {"repo": json.load(open(os.path.join(REF, "color_profile.json"))),
                "radical": {"hue_shift": 12, "sat_scale": 1.2, "val_scale": 0.9, "contrast": 1.1, "brightness": 5,
                            "radical_mode": 1, "target_hue": 30, "hue_window": 26},
                "fractional": {"hue_shift": 17.5, "sat_scale": 1.3, "val_scale": 0.8, "contrast": 0.9, "brightness": 12}}
    cp = {"profiles": np.array(json.dumps(profiles))}
    cp_kat = {}
    for pname, prof in profiles.items():
        e2 = fe.ImageEnhancer()
        e2.profile = prof
        for name, img in (("board_90x121", synth.board_frame(90, 121, 21)), ("noise_64x96", synth.noise_frame(64, 96, 22))):
            cp[pname + "/" + name] = e2.apply_color_profile(img)
        big = synth.noise_frame(1080, 1920, 0)
        cp_kat[pname] = {"apply_1920x1080": sha(e2.apply_color_profile(big)),
                         "process_pipeline_board_96x128": None}
        cp[pname + "/pipeline_board_96x128"] = e2.process_pipeline(synth.board_frame(96, 128, 1))
    cp["kat"] = np.array(json.dumps(cp_kat))
    np.savez_compressed(os.path.join(OUT, "color_profile.npz"), **cp)

    # ---- 8. Canny + refine_grid (calibration-time "next" scope row) -------------------------------
    rg = {}
    for seed in (7, 8, 9):
        b0, b1 = synth.board_with_pieces(seed, seed, 620)
        for tag, im in (("plain", b0), ("pieces", b1)):
            sgx = ge.SmartGridExtractor()
            gx, gy = sgx.refine_grid(im)
            edges = cv2.Canny(cv2.cvtColor(im, cv2.COLOR_BGR2GRAY), 50, 150)
            rg["%d_%s" % (seed, tag)] = {"grid_x": [int(v) for v in gx], "grid_y": [int(v) for v in gy],
                                         "edges_sha": sha(edges), "edge_px": int(np.count_nonzero(edges))}
    small = synth.board_frame(97, 133, 3)
    rg["canny_97x133_30_100"] = sha(cv2.Canny(cv2.cvtColor(small, cv2.COLOR_BGR2GRAY), 30, 100))
    json.dump(rg, open(os.path.join(OUT, "refine_grid.json"), "w"), indent=1)

    # ---- 9. Hough circles per square ("next" scope row, piece_detector.py:210-270) ----------------
    cv2.setNumThreads(1)      # cv2.HoughCircles' multi-threaded merge order is not needed for the result, but keep it fixed
    plane, hrects = synth.shape_atlas(3, 20, 8, 100, 512)
    cases = [{"params": dict(dp=1.2, param1=100, param2=25, min_radius_ratio=0.20, max_radius_ratio=0.55, min_dist_div=3)},
             {"params": dict(dp=1.2, param1=100, param2=30, min_radius_ratio=0.12, max_radius_ratio=0.55, min_dist_div=3)},
             {"params": dict(dp=1.0, param1=50, param2=10, min_radius=0, max_radius=0, min_dist=5.5)},
             {"params": dict(dp=2.0, param1=30, param2=5, min_radius=3, max_radius=2, min_dist=4.0)},
             {"params": dict(dp=1.7, param1=3, param2=2, min_radius=2, max_radius=40, min_dist=12.0)}]
    hz = {"plane": plane}
    for ci, case in enumerate(cases):
        q = case["params"]
        for i, (x, y, w, h) in enumerate(hrects):
            g = np.ascontiguousarray(plane[y:y + h, x:x + w])
            md = min(h, w)
            if "min_dist_div" in q:        # the reference's own argument expressions (piece_detector.py:225-241)
                args = dict(minDist=md // q["min_dist_div"], minRadius=int(md * q["min_radius_ratio"]),
                            maxRadius=int(md * q["max_radius_ratio"]))
            else:
                args = dict(minDist=q["min_dist"], minRadius=q["min_radius"], maxRadius=q["max_radius"])
            if args["minDist"] <= 0:
                hz["c%d_s%d" % (ci, i)] = np.zeros((0, 3), np.float32)
                continue
            c = cv2.HoughCircles(g, cv2.HOUGH_GRADIENT, dp=q["dp"], param1=q["param1"], param2=q["param2"], **args)
            hz["c%d_s%d" % (ci, i)] = np.zeros((0, 3), np.float32) if c is None else c[0].astype(np.float32)
    hz["meta"] = np.frombuffer(json.dumps({"rects": [list(r) for r in hrects], "cases": cases}).encode(), np.uint8)
    np.savez_compressed(os.path.join(OUT, "hough.npz"), **hz)
    # the reference's own method on the 64 preprocessed squares of boards with pieces
    hr = {}
    for seed in (11, 12, 13):
        _, bp = synth.board_with_pieces(seed, 7, 620)
        det = pdm.PieceDetector()
        rows = {}
        for pos, sq in ge.GridExtractor().split_board(bp).items():
            found, center, radius, kind = det._detect_circle_unified(det._preprocess_square(sq))
            full = det.detect_piece(sq, pos)
            rows["%d_%d" % pos] = {"found": bool(found), "center": None if center is None else [int(center[0]), int(center[1])],
                                   "radius": None if radius is None else int(radius), "kind": kind,
                                   "has_piece": bool(full["has_piece"]), "method": full["method"],
                                   "confidence": float(full["confidence"])}
        hr[str(seed)] = rows
    json.dump(hr, open(os.path.join(OUT, "hough_reference.json"), "w"), indent=0)
    cv2.setNumThreads(-1)

    # ---- 10. cv2.rotate (game_session.py:103-104,125-126: the warped board turned by 180 degrees) ----
    rot = {}
    for name, im in (("noise_37x53x3", synth.noise_frame(37, 53, 5)), ("noise_64x96x3", synth.noise_frame(64, 96, 6)),
                     ("gray_45x31", synth.noise_frame(45, 31, 7)[:, :, 0].copy())):
        for code, cname in ((cv2.ROTATE_90_CLOCKWISE, "cw"), (cv2.ROTATE_180, "180"), (cv2.ROTATE_90_COUNTERCLOCKWISE, "ccw")):
            rot["%s_%s" % (name, cname)] = {"code": int(code), "sha": sha(cv2.rotate(im, code)),
                                            "shape": list(cv2.rotate(im, code).shape)}
    # the reference's sequence: warp_image then rotate 180 (orientation_flipped)
    img = synth.noise_frame(270, 480, 9)
    w180 = cv2.rotate(bd.warp_image(img, synth.calib_points(270, 480), display_size=(200, 180), margin=20)[0], cv2.ROTATE_180)
    rot["warp_small_rot180_sha"] = sha(w180)
    json.dump(rot, open(os.path.join(OUT, "rotate.json"), "w"), indent=1)

    # ---- 11. find_chessboard_corners (calibration time; board_detection.py:4-27) --------------------
    fc = {}
    for (Hs, Ws, seed) in ((1080, 1920, 1), (720, 1280, 2), (1080, 1920, 3)):
        im = synth.table_scene(Hs, Ws, seed)
        g7 = cv2.GaussianBlur(cv2.cvtColor(im, cv2.COLOR_BGR2GRAY), (7, 7), 1)
        mask = cv2.dilate(cv2.Canny(g7, 30, 100), np.ones((5, 5), np.uint8), iterations=3)
        corners = bd.find_chessboard_corners(im)
        fc["%dx%d_%d" % (Hs, Ws, seed)] = {"blur_sha": sha(g7), "mask_sha": sha(mask), "mask_px": int(np.count_nonzero(mask)),
                                           "corners": corners.reshape(-1, 2).tolist() if corners.size else []}
    plain = synth.board_frame(270, 480, 3)
    fc["no_board_270x480"] = {"corners": bd.find_chessboard_corners(plain).reshape(-1, 2).tolist()}
    small = synth.noise_frame(97, 133, 4)[:, :, 1].copy()
    fc["blur_7_1_97x133"] = sha(cv2.GaussianBlur(small, (7, 7), 1))
    fc["blur_9_2.5_97x133"] = sha(cv2.GaussianBlur(small, (9, 9), 2.5))
    fc["dilate_5x5x3_97x133"] = sha(cv2.dilate(small, np.ones((5, 5), np.uint8), iterations=3))
    fc["dilate_7x3x2_97x133"] = sha(cv2.dilate(small, np.ones((7, 3), np.uint8), iterations=2))
    json.dump(fc, open(os.path.join(OUT, "find_corners.json"), "w"), indent=1)

    # ---- 6. warp on a small frame (full output) ---------------------------------------------------
    img = synth.noise_frame(270, 480, 9)
    pts = synth.calib_points(270, 480)
    w, M, S = bd.warp_image(img, pts, display_size=(200, 180), margin=20)
    np.savez_compressed(os.path.join(OUT, "warp_small.npz"), points=pts, warped=w, matrix=M,
                        board_size=np.array(S))
    print("golden fixtures written to", OUT)
    for f in sorted(os.listdir(OUT)):
        print("  %-28s %8d bytes" % (f, os.path.getsize(os.path.join(OUT, f))))


if __name__ == "__main__":
    main()
