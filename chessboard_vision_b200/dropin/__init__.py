"""Reference-named modules (frame_enhancer, grid_extractor, board_detection,
change_detector, piece_detector).  Put THIS directory ahead of the reference
checkout on sys.path and game_session.py / play_lichess.py /
calibrate_sensitivity.py import the B200 implementations unchanged:

    import sys, chessboard_vision_b200.dropin as d
    sys.path.insert(0, d.PATH)
"""
import os

PATH = os.path.dirname(os.path.abspath(__file__))


def install():
    """Insert the drop-in directory at the front of sys.path (idempotent)."""
    import sys
    if PATH in sys.path:
        sys.path.remove(PATH)
    sys.path.insert(0, PATH)
    return PATH
