"""cv2.rotate (the 180-degree turn of the warped board when the session's orientation is flipped,
game_session.py:103-104,125-126): oracle vs the committed cv2 digests on CPU, GPU kernel and the
pipeline's rotate_180 flag vs the oracle on the GPU."""
import hashlib
import json
import os

import numpy as np
import pytest

from chessboard_vision_b200 import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
IMAGES = {"noise_37x53x3": lambda: synth.noise_frame(37, 53, 5), "noise_64x96x3": lambda: synth.noise_frame(64, 96, 6),
          "gray_45x31": lambda: synth.noise_frame(45, 31, 7)[:, :, 0].copy()}
CODES = {"cw": 0, "180": 1, "ccw": 2}


def test_oracle_rotate_matches_cv2_digests(oracle):
    gold = json.load(open(os.path.join(G, "rotate.json")))
    for name, gen in IMAGES.items():
        im = gen()
        for cname, code in CODES.items():
            want = gold["%s_%s" % (name, cname)]
            assert want["code"] == code
            got = oracle.rotate(im, code)
            assert list(got.shape) == want["shape"] and sha(got) == want["sha"], (name, cname)
        assert np.array_equal(oracle.rotate(im, 1), im[::-1, ::-1])
        assert np.array_equal(oracle.rotate(im, 0), np.rot90(im, -1))
        assert np.array_equal(oracle.rotate(im, 2), np.rot90(im, 1))
    with pytest.raises(ValueError):
        oracle.rotate(IMAGES["gray_45x31"](), 3)
    # warp_image followed by the 180-degree turn
    img = synth.noise_frame(270, 480, 9)
    z = np.load(os.path.join(G, "warp_small.npz"))
    S = int(z["board_size"])
    assert sha(oracle.rotate(oracle.warp(img, z["matrix"], S), 1)) == gold["warp_small_rot180_sha"]


@pytest.mark.gpu
def test_gpu_rotate_all_codes_and_shapes(engine, oracle):
    gold = json.load(open(os.path.join(G, "rotate.json")))
    for name, gen in IMAGES.items():
        im = gen()
        for cname, code in CODES.items():
            assert sha(engine.rotate(im, code)) == gold["%s_%s" % (name, cname)]["sha"], (name, cname)
    batch = synth.frame_batch(3, 77, 130, "noise", 4)
    for code in (0, 1, 2):
        got = engine.rotate(batch, code)
        for i in range(3):
            assert np.array_equal(got[i], oracle.rotate(batch[i], code))
    big = synth.noise_frame(1080, 1920, 1)
    assert np.array_equal(engine.rotate(big, 0), np.rot90(big, -1))
    with pytest.raises(ValueError):
        engine.rotate(batch, 3)
    with pytest.raises(ValueError):
        engine.rotate(np.zeros((4, 4, 2), np.uint8), 1)


@pytest.mark.gpu
def test_gpu_warp_and_pipeline_with_the_board_turned(engine, oracle):
    from chessboard_vision_b200.engine import grid_rects, SQ_PD_STATS
    gold = json.load(open(os.path.join(G, "rotate.json")))
    img = synth.noise_frame(270, 480, 9)
    z = np.load(os.path.join(G, "warp_small.npz"))
    S = int(z["board_size"])
    assert sha(engine.warp(img, z["matrix"], S, rotate_180=True)) == gold["warp_small_rot180_sha"]
    # whole path: statistics of the squares of the turned board
    H, W, S = 270, 480, 160
    frames = synth.frame_batch(2, H, W, "board", 5)
    M = engine.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
    rects, _ = grid_rects(S)
    st = engine.new_state(2, S, S)
    pp = engine.pipeline_params(squares=engine.square_params(ops=SQ_PD_STATS), board_size=S, rotate_180=True)
    _, stats = engine.pipeline(frames, M, rects, pp, st)
    for f in range(2):
        board = oracle.rotate(oracle.warp(oracle.process_pipeline(frames[f], True), M, S), 1)
        for j, (x, y, w, h) in enumerate(rects):
            o = oracle.pd_square_stats(oracle.square_preprocess(board[y:y + h, x:x + w], 5))
            assert stats[f, j]["sum"] == o["sum"] and stats[f, j]["sumsq"] == o["sumsq"]
            assert list(stats[f, j]["ring_sum"]) == o["ring_sum"]
    st.free()
