"""chessboard_vision_b200 -- B200-native (sm_100a) implementation of the
ChessVision per-frame hot path: frame_enhancer chain, board warp, 64-square
split and the change_detector / piece_detector square statistics, behind the
reference's own module API (hericmr/chessboard-vision).

(The distribution is named chessboard-vision_b200; a Python package name
cannot contain '-', hence the underscore.)

    from chessboard_vision_b200 import Engine          # C-ABI host layer
    from chessboard_vision_b200.dropin import ...      # reference-named modules

Importing this package does not touch the GPU; creating an Engine does, and
fails loudly when libcvb200.so or a B200 is missing (no CPU fallback).
"""
from ._lib import CvbError, LIB_PATH  # noqa: F401


def __getattr__(name):
    if name in ("Engine", "DevArray", "State", "grid_rects", "default_engine", "STATS_DTYPE"):
        from . import engine
        return getattr(engine, name)
    raise AttributeError(name)


__version__ = "0.1.0"
