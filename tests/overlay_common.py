"""Overlay scenarios shared by the golden generator (tools/make_golden_overlay.py) and the tests: session states that
exercise every branch of GameSession._draw_interface (game_session.py:293-388) and a deterministic warped-board image."""
import hashlib

import numpy as np

START = {}
for _f, _s in enumerate("RNBQKBNR"):
    START[(_f, 0)] = _s
    START[(_f, 1)] = "P"
    START[(_f, 6)] = "p"
    START[(_f, 7)] = _s.lower()
AFTER_E4 = dict(START)
del AFTER_E4[(4, 1)]
AFTER_E4[(4, 3)] = "P"

SMART_X = [0, 79.6, 157.2, 234.9, 310.1, 386.5, 464.8, 541.3, 619.9]
SMART_Y = [0.4, 80.2, 158.7, 235.1, 311.9, 388.0, 465.5, 542.6, 620.0]

# (name, board_size, state)
SCENARIOS = [
    ("idle_no_board", 800, dict(pieces=None, fps=0.0)),
    ("start_position", 800, dict(pieces=START, white_to_move=True, fps=29.97)),
    ("last_move", 800, dict(pieces=AFTER_E4, white_to_move=False, last_move=((4, 1), (4, 3)), fps=31.2)),
    ("adjacent_last_move", 800, dict(pieces=START, white_to_move=True, last_move=((3, 0), (4, 0)), fps=12.5)),
    ("noise_lifted_radar", 800, dict(pieces=AFTER_E4, white_to_move=False, noise_active=True, lifted=(6, 7),
                                     radar=[(5, 5), (7, 5)], last_move=((4, 1), (4, 3)), fps=101.25)),
    ("smart_grid", 620, dict(pieces=AFTER_E4, white_to_move=True, grid_lines_x=SMART_X, grid_lines_y=SMART_Y,
                             last_move=((6, 7), (5, 5)), lifted=(0, 0), radar=[(0, 2), (0, 3), (1, 1)], fps=59.94)),
    ("small_board", 203, dict(pieces=START, white_to_move=True, noise_active=True, last_move=((1, 0), (2, 2)),
                              lifted=(7, 7), radar=[(7, 0), (0, 7)], fps=7.0)),
    ("corner_highlights", 401, dict(pieces={(0, 0): "K", (7, 7): "k", (7, 0): "Q", (0, 7): "q"}, white_to_move=False,
                                    last_move=((7, 0), (7, 7)), lifted=(7, 0), radar=[(7, 7), (0, 0), (7, 0)], fps=0.04)),
]


def board_image(size, seed=0):
    """a warped-board stand-in: checker squares plus seeded noise so that every blend sees all byte values"""
    rng = np.random.default_rng(seed)
    sq = max(1, size // 8)
    yy, xx = np.mgrid[:size, :size]
    light = ((yy // sq + xx // sq) % 2 == 0)[..., None]
    base = np.where(light, np.array([150, 170, 180]), np.array([70, 100, 120]))
    img = base + rng.integers(-70, 76, (size, size, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def digest(img):
    return hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()


def random_display_list(rng, H, W, n_ops):
    """every op kind, opaque and blended, grouped shapes, text stamps, partly or wholly outside the image"""
    from chessboard_vision_b200.overlay import DisplayList
    import cv2
    dl = DisplayList()
    weights = ((1.0, 0.0), (0.3, 0.7), (0.5, 0.5), (0.4, 0.6), (0.6, 0.4), (0.9, 0.35), (0.0, 1.0))
    while len(dl.ops) < n_ops:
        kind = int(rng.integers(0, 5))
        color = tuple(int(c) for c in rng.integers(0, 256, 3))
        a, b = weights[int(rng.integers(0, len(weights)))]
        x, y = int(rng.integers(-30, W + 30)), int(rng.integers(-30, H + 30))
        if kind == 0:
            dl.rectangle((x, y), (x + int(rng.integers(-40, 90)), y + int(rng.integers(-40, 90))), color, a, b)
        elif kind == 1:
            dl.circle((x, y), int(rng.integers(0, 60)), color, a, b)
        elif kind == 2:
            dl.put_text("".join(chr(int(c)) for c in rng.integers(33, 127, int(rng.integers(1, 9)))), (x, y),
                        cv2.FONT_HERSHEY_SIMPLEX, float(rng.uniform(0.4, 1.6)), color, int(rng.integers(1, 5)))
        elif kind == 3:                                          # shapes on one overlay copy
            g = dl.group()
            for _ in range(int(rng.integers(2, 5))):
                if rng.integers(0, 2):
                    dl.rectangle((x, y), (x + int(rng.integers(0, 70)), y + int(rng.integers(0, 70))), color, a, b, g)
                else:
                    dl.circle((x, y), int(rng.integers(0, 40)), color, a, b, g)
                x, y = x + int(rng.integers(-30, 31)), y + int(rng.integers(-30, 31))
        else:
            dl.line((x, 0), (x, H), color) if rng.integers(0, 2) else dl.line((0, y), (W, y), color)
    return dl
