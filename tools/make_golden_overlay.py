"""What the reference's UNMODIFIED GameSession._draw_interface (game_session.py:293-388) draws for the scenarios of
tests/overlay_common.py: tests/golden/overlay.json (a digest per scenario, the image shown by cv2.imshow("Tabuleiro"))
and tests/golden/overlay_small.npz (the whole image of the smallest scenario).  Build container only (/root/reference).

    python tools/make_golden_overlay.py
"""
import json
import os
import sys
import threading
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import caller_harness as ch
import overlay_common as oc


def reference_draw(gs_mod, chess, size, state, shown):
    """call the unbound method on a stand-in `self` that carries exactly the attributes it reads"""
    board = None
    if state.get("pieces") is not None:
        board = types.SimpleNamespace()
        pieces = state["pieces"]

        def piece_at(sq):
            sym = pieces.get((chess.square_file(sq), chess.square_rank(sq)))
            if not sym:
                return None
            kind = {"p": chess.PAWN, "n": chess.KNIGHT, "b": chess.BISHOP, "r": chess.ROOK, "q": chess.QUEEN, "k": chess.KING}[sym.lower()]
            return chess.Piece(kind, sym.isupper())
        board.piece_at = piece_at
        board.turn = bool(state.get("white_to_move"))
        lm = state.get("last_move")
        board.move_stack = [1] if lm else []
        board.peek = lambda: chess.Move(chess.square(*lm[0]), chess.square(*lm[1]))
    fake = types.SimpleNamespace(
        grid=types.SimpleNamespace(grid_lines_x=state.get("grid_lines_x"), grid_lines_y=state.get("grid_lines_y")),
        board_lock=threading.RLock(), game=types.SimpleNamespace(board=board),
        lifted_piece_square=state.get("lifted"), current_radar_destinations=list(state.get("radar", ())),
        fps_display=state.get("fps", 0.0))
    vis = oc.board_image(size)
    noise = gs_mod.NoiseState.NOISE_ACTIVE if state.get("noise_active") else gs_mod.NoiseState.IDLE
    shown.clear()
    gs_mod.GameSession._draw_interface(fake, vis, size, noise, np.zeros((4, 4, 3), np.uint8))
    assert shown and shown[0][0] == "Tabuleiro"
    return shown[0][1]


def main():
    import cv2
    out = {}
    with ch.caller_env("reference"):
        import chess
        import game_session
        shown = []
        cv2.imshow = lambda name, img: shown.append((name, img.copy()))
        small = None
        for name, size, state in oc.SCENARIOS:
            img = reference_draw(game_session, chess, size, state, shown)
            out[name] = {"size": size, "sha256": oc.digest(img), "input_sha256": oc.digest(oc.board_image(size))}
            if name == "small_board":
                small = img
    gdir = os.path.join(ROOT, "tests", "golden")
    json.dump(out, open(os.path.join(gdir, "overlay.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(gdir, "overlay_small.npz"), small_board=small)
    print("wrote overlay.json (%d scenarios), overlay_small.npz %d bytes" % (len(out), os.path.getsize(os.path.join(gdir, "overlay_small.npz"))))


if __name__ == "__main__":
    main()
