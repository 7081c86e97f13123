"""CPU oracle of the board overlay -- TEST INFRASTRUCTURE, never on the product path.

Two independent restatements of GameSession._draw_interface (game_session.py:293-388):

* `draw_interface_cv2` is the reference's own cv2 call sequence (line / rectangle / circle / addWeighted / putText on
  `vis` and on `overlay = vis.copy()`), minus the two cv2.imshow calls.  Pinned: tests/golden/overlay.json holds
  digests of what the UNMODIFIED GameSession._draw_interface drew for the same states (tools/make_golden_overlay.py).
* `apply_display_list` interprets the display list of include/cvb200.h (cvb_overlay_op) in NumPy with the arithmetic
  of OpenCV's drawing.cpp / arithm: inclusive clipped rectangles, the midpoint spans of Circle(), 1-bit stamps, and
  addWeighted for 8-bit data = saturate(rint(fma(src1, (float)alpha, src2 * (float)beta))) (checked against
  cv2.addWeighted on all 65536 byte pairs in tests/test_overlay_cpu.py).
"""
import cv2
import numpy as np

STATE_KEYS = ("noise_active", "grid_lines_x", "grid_lines_y", "last_move", "lifted", "radar", "pieces", "white_to_move", "fps")


def draw_interface_cv2(vis, board_size, noise_active=False, grid_lines_x=None, grid_lines_y=None, last_move=None,
                       lifted=None, radar=(), pieces=None, white_to_move=None, fps=0.0):
    """game_session.py:293-388 with the session state passed in; draws on `vis` in place"""
    sq_size = board_size // 8
    if grid_lines_x and grid_lines_y:                                            # :297-303
        for x in grid_lines_x:
            cv2.line(vis, (int(x), 0), (int(x), board_size), (0, 200, 100), 1)
        for y in grid_lines_y:
            cv2.line(vis, (0, int(y)), (board_size, int(y)), (0, 200, 100), 1)
    else:                                                                        # :304-308
        for i in range(9):
            cv2.line(vis, (i * sq_size, 0), (i * sq_size, board_size), (50, 50, 50), 1)
            cv2.line(vis, (0, i * sq_size), (board_size, i * sq_size), (50, 50, 50), 1)
    if noise_active:                                                             # :311-316
        overlay = vis.copy()
        overlay[:] = (0, 0, 80)
        cv2.addWeighted(overlay, 0.3, vis, 0.7, 0, vis)
        cv2.putText(vis, "jogada em andamento", (board_size // 2 - 120, board_size // 2), cv2.FONT_HERSHEY_SIMPLEX, 1.0,
                    (0, 0, 255), 3)
    if last_move is not None:                                                    # :319-338
        overlay = vis.copy()
        for f, r in last_move:
            x1, y1 = f * sq_size, (7 - r) * sq_size
            cv2.rectangle(overlay, (x1, y1), (x1 + sq_size, y1 + sq_size), (100, 50, 0), -1)
        cv2.addWeighted(overlay, 0.5, vis, 0.5, 0, vis)
    if lifted:                                                                   # :341-347
        x1, y1 = lifted[0] * sq_size, (7 - lifted[1]) * sq_size
        overlay = vis.copy()
        cv2.rectangle(overlay, (x1, y1), (x1 + sq_size, y1 + sq_size), (0, 0, 200), -1)
        cv2.addWeighted(overlay, 0.4, vis, 0.6, 0, vis)
    for f, r in radar:                                                           # :349-357
        x1, y1 = f * sq_size, (7 - r) * sq_size
        overlay = vis.copy()
        cv2.circle(overlay, (x1 + sq_size // 2, y1 + sq_size // 2), int(sq_size * 0.4 / 2), (0, 100, 0), -1)
        cv2.addWeighted(overlay, 0.6, vis, 0.4, 0, vis)
    if pieces is not None:                                                       # :360-378
        for f in range(8):
            for r in range(8):
                sym = pieces.get((f, r))
                if not sym:
                    continue
                x, y = f * sq_size + sq_size // 2, (7 - r) * sq_size + sq_size // 2
                white = sym.isupper()
                color = (255, 255, 255) if white else (0, 0, 0)
                bg = (0, 0, 0) if white else (255, 255, 255)
                cv2.putText(vis, sym, (x - 15, y + 10), cv2.FONT_HERSHEY_SIMPLEX, 1.2, bg, 4)
                cv2.putText(vis, sym, (x - 15, y + 10), cv2.FONT_HERSHEY_SIMPLEX, 1.2, color, 2)
    turn_text = "Brancas" if (pieces is not None and white_to_move) else "Pretas"   # :381-383
    cv2.putText(vis, "Turno: %s" % turn_text, (10, 30), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (0, 255, 0), 2)
    cv2.putText(vis, "FPS: %.1f" % fps, (board_size - 150, 30), cv2.FONT_HERSHEY_SIMPLEX, 0.6, (0, 255, 255), 2)
    return vis


def circle_half_widths(r):
    """half width of the rows dy = 0..r of cv2.circle(.., r, .., -1): drawing.cpp Circle(), fill branch"""
    half = [0] * (r + 1)
    err, dx, dy, plus, minus = 0, r, 0, 1, (r << 1) - 1
    while dx >= dy:
        half[dy] = max(half[dy], dx)
        half[dx] = max(half[dx], dy)
        dy += 1
        err += plus
        plus += 2
        mask = (1 if err <= 0 else 0) - 1
        err -= minus & mask
        dx += mask
        minus -= mask & 2
    return half


def add_weighted_u8(src1, alpha, src2, beta):
    """cv2.addWeighted(src1, alpha, src2, beta, 0) for u8: one fused multiply-add in f32 around a rounded product"""
    a, b = np.float32(alpha), np.float32(beta)
    t = (np.asarray(src2, np.float32) * b).astype(np.float64)                     # rounded to f32
    exact = np.asarray(src1, np.float64) * float(a) + t                           # exact in f64 (35 significant bits)
    return np.clip(np.rint(exact.astype(np.float32)), 0, 255).astype(np.uint8)


def apply_display_list(img, ops, n_ops, masks):
    """ops: sequence of objects with the fields of cvb_overlay_op (e.g. _lib.OverlayOp); draws on a copy of img"""
    out = np.array(img, np.uint8, copy=True)
    H, W = out.shape[:2]
    masks = np.frombuffer(bytes(masks), np.uint8)
    done = np.zeros((H, W), np.int64)            # group that last touched the pixel
    for i in range(n_ops):
        o = ops[i]
        cover = np.zeros((H, W), bool)
        if o.kind == 0:
            x0, x1 = sorted((o.x0, o.x1))
            y0, y1 = sorted((o.y0, o.y1))
            cover[max(0, y0):max(0, y1 + 1), max(0, x0):max(0, x1 + 1)] = True
        elif o.kind == 1:
            for dy, hw in enumerate(circle_half_widths(o.x1)):
                for y in {o.y0 - dy, o.y0 + dy}:
                    if 0 <= y < H:
                        cover[y, max(0, o.x0 - hw):max(0, o.x0 + hw + 1)] = True
        else:
            w, h = o.x1, o.y1
            rb = (w + 7) // 8
            bits = np.unpackbits(masks[o.aux_ofs:o.aux_ofs + rb * h].reshape(h, rb), axis=1, bitorder="little")[:, :w].astype(bool)
            ys, xs = np.nonzero(bits)
            ys, xs = ys + o.y0, xs + o.x0
            ok = (ys >= 0) & (ys < H) & (xs >= 0) & (xs < W)
            cover[ys[ok], xs[ok]] = True
        if o.group:
            cover &= done != o.group
        color = np.array(list(o.color)[:3], np.uint8)
        if o.alpha == 1.0 and o.beta == 0.0:
            out[cover] = color
        else:
            out[cover] = add_weighted_u8(np.broadcast_to(color, out[cover].shape), o.alpha, out[cover], o.beta)
        done[cover] = o.group
    return out
