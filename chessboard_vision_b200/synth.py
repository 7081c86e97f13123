"""Seeded synthetic inputs: the same frames feed
the CUDA path, the oracle, the golden-vector generator and the benchmark."""
import numpy as np

CALIB_CORNERS_1080P = [[556, 112], [1560, 108], [1562, 1024], [550, 1005]]   # calibration.json:2-19 (TL,TR,BR,BL)
CALIB_GRID_X = [0, 79, 157, 234, 310, 386, 464, 541, 620]                    # calibration.json:22-32
CALIB_GRID_Y = [0, 80, 158, 235, 311, 388, 465, 542, 620]                    # calibration.json:33-42


def noise_frame(H, W, seed=0):
    """Uniform noise: worst case for the histograms and the colour-weight LUT."""
    return np.random.default_rng(seed).integers(0, 256, (H, W, 3), dtype=np.uint8)


def board_frame(H, W, seed=0):
    """8x8 checker (60/180) with per-channel low-frequency gradients + N(0,8) noise."""
    rng = np.random.default_rng(seed)
    x = np.arange(W)[None, :]
    y = np.arange(H)[:, None]
    base = ((x // max(W // 8, 1) + y // max(H // 8, 1)) % 2) * 120 + 60
    b = base + 20 * np.sin(x / 37) + 0 * y
    g = 0.8 * base + 10 * np.cos(y / 23) + 0 * x
    r = 0.6 * base + 0.05 * x + 0 * y
    img = np.stack([b, g, r], -1) + rng.normal(0, 8, (H, W, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def frame_batch(n, H, W, kind="board", seed0=0):
    gen = board_frame if kind == "board" else noise_frame
    return np.stack([gen(H, W, seed0 + i) for i in range(n)])


def calib_points(H, W):
    """calibration.json corners scaled to the frame, reordered TL,TR,BL,BR as
    board_detection.reorder does before warp_image (board_detection.py:49-58)."""
    c = np.array(CALIB_CORNERS_1080P, np.float64) * [W / 1920.0, H / 1080.0]
    tl, tr, br, bl = c
    return np.float32([tl, tr, bl, br])


def change_pair(frame, squares, value=255, board_pts=None):
    """Copy of `frame` with axis-aligned blocks overwritten (drives LEVE/PARCIAL/TOTAL)."""
    out = frame.copy()
    for (x, y, w, h) in squares:
        out[y:y + h, x:x + w] = value
    return out


def board_with_pieces(seed=11, base_seed=7, size=620):
    """A warped-board-like image (board_frame) with 20 random coloured discs near square centres."""
    rng = np.random.default_rng(seed)
    board = board_frame(size, size, base_seed)
    out = board.copy()
    yy, xx = np.ogrid[:size, :size]
    sq = size // 8
    for _ in range(20):
        c, r = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        cx, cy = c * sq + sq // 2 + int(rng.integers(-4, 5)), r * sq + sq // 2 + int(rng.integers(-4, 5))
        rad = int(rng.integers(18, 30))
        col = rng.integers(0, 256, 3)
        out[((xx - cx) ** 2 + (yy - cy) ** 2) <= rad * rad] = col
    return board, out
