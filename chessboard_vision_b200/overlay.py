"""Device-side board overlay: GameSession._draw_interface (game_session.py:293-388) as a
display list that one kernel applies to the warped board on the GPU (`cvb_overlay_dev`, csrc/cvb_overlay.cu),
pixel for pixel what the reference's cv2.line / rectangle / circle / addWeighted / putText sequence draws.

`DisplayList` is the general form (one method per cv2 drawing call the reference uses); `BoardOverlay.draw_interface`
builds the list of `_draw_interface` from the session state it reads (grid lines, noise flag, last move, lifted
square, radar destinations, pieces, side to move, fps).  Text is rasterised once per string with cv2.putText itself
(the glyph cache below) and applied as 1-bit stamps, so the letters are OpenCV's own; nothing else touches the host.
There is no CPU drawing fallback: without the library / a GPU the Engine raises.
"""
import ctypes as C

import cv2
import numpy as np

from ._lib import OV_CIRCLE, OV_RECT, OV_STAMP, OverlayOp

_FAR = 1 << 19       # "the whole image" as a rectangle; the kernel clips to the image


class DisplayList:
    """Drawing calls in order.  Colours are (B, G, R); `alpha`/`beta` are the addWeighted weights of
    `overlay = vis.copy(); <shape on overlay>; cv2.addWeighted(overlay, alpha, vis, beta, 0, vis)`."""

    _glyphs = {}       # (text, font, scale, thickness) -> (dx, dy, w, h, packed rows)

    def __init__(self):
        self.ops, self.masks, self._groups = [], bytearray(), 0

    def group(self):
        """id for shapes drawn on ONE overlay copy before a single addWeighted (game_session.py:322-338)"""
        self._groups += 1
        return self._groups

    def _add(self, kind, x0, y0, x1, y1, color, alpha, beta, group, aux=0):
        op = OverlayOp(kind, int(x0), int(y0), int(x1), int(y1), (C.c_uint8 * 4)(*[int(c) & 255 for c in color][:3], 0),
                       np.float32(alpha), np.float32(beta), int(group), int(aux))
        self.ops.append(op)

    def rectangle(self, pt1, pt2, color, alpha=1.0, beta=0.0, group=0):
        """cv2.rectangle(img, pt1, pt2, color, -1)"""
        self._add(OV_RECT, pt1[0], pt1[1], pt2[0], pt2[1], color, alpha, beta, group)

    def line(self, pt1, pt2, color):
        """cv2.line(img, pt1, pt2, color, 1) for the axis-aligned lines the reference draws"""
        if pt1[0] != pt2[0] and pt1[1] != pt2[1]:
            raise ValueError("only axis-aligned lines are part of the board overlay")
        self._add(OV_RECT, pt1[0], pt1[1], pt2[0], pt2[1], color, 1.0, 0.0, 0)

    def fill(self, color, alpha, beta):
        """overlay[:] = color; addWeighted(overlay, alpha, vis, beta, 0, vis)   (game_session.py:313-315)"""
        self._add(OV_RECT, -_FAR, -_FAR, _FAR, _FAR, color, alpha, beta, 0)

    def circle(self, center, radius, color, alpha=1.0, beta=0.0, group=0):
        """cv2.circle(img, center, radius, color, -1)"""
        self._add(OV_CIRCLE, center[0], center[1], int(radius), 0, color, alpha, beta, group)

    @classmethod
    def _glyph(cls, text, font, scale, thickness):
        key = (text, int(font), float(scale), int(thickness))
        g = cls._glyphs.get(key)
        if g is None:
            (w, h), base = cv2.getTextSize(text, font, scale, thickness)
            pad = thickness + 8
            while True:
                canvas = np.zeros((h + base + 2 * pad, w + 2 * pad), np.uint8)
                cv2.putText(canvas, text, (pad, pad + h), font, scale, 255, thickness)
                ys, xs = np.nonzero(canvas)
                if len(ys) == 0:
                    g = (0, 0, 0, 0, b"")
                    break
                if ys.min() > 0 and xs.min() > 0 and ys.max() < canvas.shape[0] - 1 and xs.max() < canvas.shape[1] - 1:
                    box = canvas[ys.min():ys.max() + 1, xs.min():xs.max() + 1] > 0
                    rows = np.packbits(box, axis=1, bitorder="little")
                    g = (int(xs.min()) - pad, int(ys.min()) - (pad + h), box.shape[1], box.shape[0], rows.tobytes())
                    break
                pad *= 2        # ink reached the canvas border: the text box under-estimated it
            if len(cls._glyphs) >= 1024:      # an FPS read-out makes new strings for hours: keep the cache bounded
                cls._glyphs.clear()
            cls._glyphs[key] = g
        return g

    def put_text(self, text, org, font, scale, color, thickness=1):
        """cv2.putText(img, text, org, font, scale, color, thickness) (LINE_8, bottomLeftOrigin False)"""
        dx, dy, w, h, rows = self._glyph(text, font, scale, thickness)
        if w == 0:
            return
        ofs = len(self.masks)
        self.masks += rows
        self._add(OV_STAMP, org[0] + dx, org[1] + dy, w, h, color, 1.0, 0.0, 0, ofs)

    def pack(self):
        arr = (OverlayOp * max(1, len(self.ops)))(*self.ops)
        return arr, len(self.ops), bytes(self.masks)


class BoardOverlay:
    """What GameSession._draw_interface draws on `vis` (game_session.py:293-388), minus the two cv2.imshow calls."""

    GRID_SMART, GRID_REGULAR = (0, 200, 100), (50, 50, 50)

    def __init__(self, engine=None, device=0):
        if engine is None:
            from .engine import default_engine
            engine = default_engine(device)
        self._e = engine

    @staticmethod
    def display_list(board_size, noise_active=False, grid_lines_x=None, grid_lines_y=None, last_move=None, lifted=None,
                     radar=(), pieces=None, white_to_move=None, fps=0.0):
        """The drawing calls of _draw_interface in the reference's order.
        last_move: ((file, rank) from, (file, rank) to) or None; lifted: (file, rank) or None; radar: [(file, rank)];
        pieces: {(file, rank): symbol} with upper case = white (chess.Piece.symbol()), None = no board yet;
        white_to_move: board.turn (None / no board reads 'Pretas' exactly as the reference's expression does)."""
        dl = DisplayList()
        S = int(board_size)
        sq = S // 8
        font = cv2.FONT_HERSHEY_SIMPLEX
        if grid_lines_x and grid_lines_y:                                         # game_session.py:297-303
            for x in grid_lines_x:
                dl.line((int(x), 0), (int(x), S), BoardOverlay.GRID_SMART)
            for y in grid_lines_y:
                dl.line((0, int(y)), (S, int(y)), BoardOverlay.GRID_SMART)
        else:                                                                     # :304-308
            for i in range(9):
                dl.line((i * sq, 0), (i * sq, S), BoardOverlay.GRID_REGULAR)
                dl.line((0, i * sq), (S, i * sq), BoardOverlay.GRID_REGULAR)
        if noise_active:                                                          # :311-316
            dl.fill((0, 0, 80), 0.3, 0.7)
            dl.put_text("jogada em andamento", (S // 2 - 120, S // 2), font, 1.0, (0, 0, 255), 3)
        if last_move is not None:                                                 # :319-338
            g = dl.group()
            for f, r in last_move:
                x1, y1 = f * sq, (7 - r) * sq
                dl.rectangle((x1, y1), (x1 + sq, y1 + sq), (100, 50, 0), 0.5, 0.5, g)
        if lifted:                                                                # :341-347
            x1, y1 = lifted[0] * sq, (7 - lifted[1]) * sq
            dl.rectangle((x1, y1), (x1 + sq, y1 + sq), (0, 0, 200), 0.4, 0.6)
        for f, r in radar:                                                        # :349-357
            x1, y1 = f * sq, (7 - r) * sq
            dl.circle((x1 + sq // 2, y1 + sq // 2), int(sq * 0.4 / 2), (0, 100, 0), 0.6, 0.4)
        if pieces is not None:                                                    # :360-378
            for f in range(8):
                for r in range(8):
                    sym = pieces.get((f, r))
                    if not sym:
                        continue
                    x, y = f * sq + sq // 2, (7 - r) * sq + sq // 2
                    white = sym.isupper()
                    color, bg = ((255, 255, 255), (0, 0, 0)) if white else ((0, 0, 0), (255, 255, 255))
                    dl.put_text(sym, (x - 15, y + 10), font, 1.2, bg, 4)
                    dl.put_text(sym, (x - 15, y + 10), font, 1.2, color, 2)
        turn = "Brancas" if (pieces is not None and white_to_move) else "Pretas"    # :381-383
        dl.put_text("Turno: %s" % turn, (10, 30), font, 0.6, (0, 255, 0), 2)
        dl.put_text("FPS: %.1f" % fps, (S - 150, 30), font, 0.6, (0, 255, 255), 2)  # :385-386
        return dl

    @staticmethod
    def session_state(session, noise_active):
        """The arguments of display_list read off a GameSession exactly where _draw_interface reads them
        (game_session.py:297-386): grid lines, last move, lifted square, radar destinations, pieces, side to move, fps.
        `session.game.board` is a python-chess Board (or None)."""
        board = session.game.board
        fr = lambda sq: (sq & 7, sq >> 3)                 # chess.square_file / chess.square_rank
        last = None
        if board and len(board.move_stack) > 0:
            mv = board.peek()
            last = (fr(mv.from_square), fr(mv.to_square))
        return dict(noise_active=bool(noise_active), grid_lines_x=session.grid.grid_lines_x, grid_lines_y=session.grid.grid_lines_y,
                    last_move=last, lifted=session.lifted_piece_square, radar=list(session.current_radar_destinations),
                    pieces=(None if not board else {fr(sq): p.symbol() for sq, p in board.piece_map().items()}),
                    white_to_move=bool(board and board.turn), fps=session.fps_display)

    def draw(self, vis, dl):
        """Apply a DisplayList.  A NumPy image is drawn in place (as cv2 does) and returned; a DevArray stays on the device."""
        ops, n, masks = dl.pack()
        if n == 0:
            return vis
        if isinstance(vis, np.ndarray):
            vis[...] = self._e.overlay(np.ascontiguousarray(vis), ops, n, masks)
            return vis
        self._e.overlay(vis, ops, n, masks)
        return vis

    def draw_interface(self, vis, board_size, **state):
        return self.draw(vis, self.display_list(board_size, **state))
