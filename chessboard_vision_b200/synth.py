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


def shape_atlas(seed=0, n_squares=48, min_side=8, max_side=128, plane_w=1024):
    """A gray plane packed with `n_squares` rectangles of random size, each holding random discs,
    rings and bars on a flat background plus noise and a 3x3 box blur -- varied input for the
    Hough-circle stage.  -> (plane u8, [(x, y, w, h), ...])"""
    rng = np.random.default_rng(seed)
    rects, x, y, row_h = [], 0, 0, 0
    for _ in range(n_squares):
        w, h = int(rng.integers(min_side, max_side + 1)), int(rng.integers(min_side, max_side + 1))
        if x + w > plane_w:
            x, y, row_h = 0, y + row_h, 0
        rects.append((x, y, w, h))
        x, row_h = x + w, max(row_h, h)
    plane = np.zeros((y + row_h, plane_w), np.uint8)
    for (x, y, w, h) in rects:
        img = np.full((h, w), float(rng.integers(40, 200)))
        yy, xx = np.ogrid[:h, :w]
        for _ in range(int(rng.integers(0, 4))):
            cx, cy = int(rng.integers(0, w)), int(rng.integers(0, h))
            rad = int(rng.integers(3, max(4, min(h, w) // 2)))
            d2 = (xx - cx) ** 2 + (yy - cy) ** 2
            mask = d2 <= rad * rad
            if rng.random() < 0.3:
                mask &= d2 >= max(rad - 2, 0) ** 2
            img[mask] = float(rng.integers(0, 256))
        if rng.random() < 0.3:
            x0, x1 = sorted(int(v) for v in rng.integers(0, w, 2))
            y0, y1 = sorted(int(v) for v in rng.integers(0, h, 2))
            img[y0:y1 + 1, x0:x1 + 1] = float(rng.integers(0, 256))
        img += rng.normal(0, float(rng.choice([0, 2, 8, 20])), img.shape)
        img = np.clip(img, 0, 255)
        if rng.random() < 0.7:
            p = np.pad(img, 1, mode="edge")
            img = sum(p[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)) / 9.0
        plane[y:y + h, x:x + w] = img.astype(np.uint8)
    return plane, rects


def table_scene(H, W, seed=1):
    """A bright board-sized quadrilateral (the calibration corners) on a darker table plus sensor noise:
    the kind of frame board_detection.find_chessboard_corners is run on."""
    rng = np.random.default_rng(seed)
    tl, tr, bl, br = calib_points(H, W).astype(np.float64)
    quad = [tl, tr, br, bl]
    yy, xx = np.mgrid[:H, :W].astype(np.float64)
    inside = np.ones((H, W), bool)
    for a, b in zip(quad, quad[1:] + quad[:1]):
        inside &= (b[0] - a[0]) * (yy - a[1]) - (b[1] - a[1]) * (xx - a[0]) >= 0
    img = np.empty((H, W, 3), np.float32)
    img[...] = (40, 50, 45)
    img[inside] = (190, 200, 210)
    img += rng.normal(0, 3, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def bgr_to_yuv(frame, fmt):
    """A camera-native frame ('yuy2': (H,W,2), 'nv12': (H*3/2,W)) showing roughly `frame` (BT.601 studio range, chroma
    of each pixel pair / 2x2 block averaged).  Only used to synthesise ingest inputs: what matters downstream is the
    exact YUV -> BGR step, not this one."""
    f = frame.astype(np.float32)
    b, g, r = f[..., 0], f[..., 1], f[..., 2]
    y = 16 + 0.257 * r + 0.504 * g + 0.098 * b
    u = 128 - 0.148 * r - 0.291 * g + 0.439 * b
    v = 128 + 0.439 * r - 0.368 * g - 0.071 * b
    q = lambda a: np.clip(np.rint(a), 0, 255).astype(np.uint8)
    H, W = y.shape
    if fmt == "yuy2":
        out = np.empty((H, W, 2), np.uint8)
        out[..., 0] = q(y)
        out[:, 0::2, 1] = q((u[:, 0::2] + u[:, 1::2]) / 2)
        out[:, 1::2, 1] = q((v[:, 0::2] + v[:, 1::2]) / 2)
        return out
    if fmt == "nv12":
        out = np.empty((H * 3 // 2, W), np.uint8)
        out[:H] = q(y)
        uv = out[H:].reshape(H // 2, W // 2, 2)
        uv[..., 0] = q((u[0::2, 0::2] + u[0::2, 1::2] + u[1::2, 0::2] + u[1::2, 1::2]) / 4)
        uv[..., 1] = q((v[0::2, 0::2] + v[0::2, 1::2] + v[1::2, 0::2] + v[1::2, 1::2]) / 4)
        return out
    raise ValueError(fmt)
