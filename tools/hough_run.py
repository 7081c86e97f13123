"""Device-resident timing of the Hough-circle stage on warped boards with pieces.
usage: python tools/hough_run.py [frames] [steps]      (CVB_CHECK=1: compare a few squares with the oracle)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from chessboard_vision_b200 import synth, hostapi, _lib
from chessboard_vision_b200.engine import Engine, grid_rects, SQ_PD_STATS

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
S = 620
eng = Engine(0)
rects, _ = grid_rects(S, synth.CALIB_GRID_X, synth.CALIB_GRID_Y)
boards = np.stack([synth.board_with_pieces(30 + i, 7, S)[1] for i in range(min(n, 8))])
boards = np.stack([boards[i % len(boards)] for i in range(n)])
st = eng.new_state(n, S, S)
stats = eng.squares(boards, rects, eng.square_params(ops=SQ_PD_STATS), state=st)
gate = np.array([[0 if hostapi.std_below(stats[f, i], 15) else 1 for i in range(64)] for f in range(n)], np.uint8)
planes = eng.upload(np.stack([st.get(f, _lib.PLANE_PD_CUR) for f in range(n)]))
p = eng.hough_params()
for name, sel in (("all 64 squares", None), ("std-gated (%.1f squares/frame)" % gate.sum(1).mean(), gate)):
    out = eng.hough(planes, rects, p, select=sel)
    eng.synchronize()
    eng.profile(True)
    for _ in range(steps):
        out = eng.hough(planes, rects, p, select=sel)
    prof = eng.profile_read()
    eng.profile(False)
    ms, c = prof["k_hough"]
    done = out["status"] == 0
    print("%-36s %8.3f ms/launch  %7.2f us/frame  %6.2f us/square   circles/frame %.1f  edges/square %.0f  centres/square %.1f"
          % (name, ms / c, ms / c / n * 1e3, ms / c / max(done.sum(), 1) * 1e3, out["count"].sum() / n,
             out["n_edges"][done].mean(), out["n_centers"][done].mean()))
if os.environ.get("CVB_CHECK"):
    import oracle as O
    geo = eng.hough_geometry(rects, p)
    pl = planes.get()
    bad = 0
    for f in range(min(n, 2)):
        for i, (x, y, w, h) in enumerate(rects):
            c = O.hough_circles(pl[f, y:y + h, x:x + w], min_dist=float(geo[i]["min_dist"]), min_radius=int(geo[i]["min_radius"]),
                                max_radius=int(geo[i]["max_radius"]))
            k = 0 if c is None else len(c)
            bad += not (out[f, i]["count"] == k and (k == 0 or np.array_equal(out[f, i]["xyr"][:k], c))) if out[f, i]["status"] == 0 else 0
    print("parity", "OK" if not bad else "MISMATCH %d" % bad)
