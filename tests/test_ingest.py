"""Camera ingest: YUY2 / NV12 -> BGR.  CPU: the oracle against cv2.cvtColor on ALL 2^24
(y, u, v) and on random frames.  GPU: the kernels against the oracle (exhaustive, aligned / generic paths) and the
host-buffer pipeline fed with native frames against the same pipeline fed with cv2's BGR."""
import numpy as np
import pytest

from chessboard_vision_b200 import synth


def _all_yuv_yuy2(u):
    """(32, 4096, 2) YUY2 image: every (y0, v) with this u, y1 a permutation of y0."""
    v = np.repeat(np.arange(256), 256); y0 = np.tile(np.arange(256), 256); y1 = (y0 * 7 + 13) % 256
    img = np.empty((1, 131072, 2), np.uint8)
    img[0, 0::2, 0] = y0; img[0, 1::2, 0] = y1; img[0, 0::2, 1] = u; img[0, 1::2, 1] = v
    return img.reshape(32, 4096, 2)


def test_oracle_equals_cv2_on_every_yuv(oracle):
    cv2 = pytest.importorskip("cv2")
    for u in range(256):
        img = _all_yuv_yuy2(u)
        assert np.array_equal(oracle.yuv_to_bgr(img, "yuy2"), cv2.cvtColor(img, cv2.COLOR_YUV2BGR_YUY2)), u


@pytest.mark.parametrize("shape", [(1080, 1920), (480, 640), (6, 10), (2, 2), (34, 70)])
def test_oracle_equals_cv2_on_frames(oracle, shape):
    cv2 = pytest.importorskip("cv2")
    H, W = shape
    rng = np.random.default_rng(H * W)
    nv = rng.integers(0, 256, (H * 3 // 2, W), dtype=np.uint8)
    yu = rng.integers(0, 256, (H, W, 2), dtype=np.uint8)
    assert np.array_equal(oracle.yuv_to_bgr(nv, "nv12"), cv2.cvtColor(nv, cv2.COLOR_YUV2BGR_NV12))
    assert np.array_equal(oracle.yuv_to_bgr(yu, "yuy2"), cv2.cvtColor(yu, cv2.COLOR_YUV2BGR_YUY2))
    with pytest.raises(ValueError):
        oracle.yuv_to_bgr(np.zeros((4, 3, 2), np.uint8), "yuy2")


def test_synthetic_yuv_frames_look_like_their_bgr_source(oracle):
    f = synth.board_frame(64, 96, 1)
    for fmt in ("yuy2", "nv12"):
        back = oracle.yuv_to_bgr(synth.bgr_to_yuv(f, fmt), fmt)
        assert np.abs(back.astype(int) - f).mean() < 6


@pytest.mark.gpu
def test_gpu_equals_oracle_on_every_yuv(engine, oracle):
    for u in range(0, 256, 5):
        img = _all_yuv_yuy2(u)
        assert np.array_equal(engine.cvt_to_bgr(img, "yuy2"), oracle.yuv_to_bgr(img, "yuy2")), u


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1080, 1920), (480, 640), (6, 10), (2, 2), (34, 70), (90, 136)])
def test_gpu_equals_oracle_on_frames(engine, oracle, shape):
    H, W = shape
    rng = np.random.default_rng(H + W)
    for fmt, frames in (("nv12", rng.integers(0, 256, (3, H * 3 // 2, W), dtype=np.uint8)),
                        ("yuy2", rng.integers(0, 256, (3, H, W, 2), dtype=np.uint8))):
        got = engine.cvt_to_bgr(frames, fmt)
        assert got.shape == (3, H, W, 3)
        for i in range(3):
            assert np.array_equal(got[i], oracle.yuv_to_bgr(frames[i], fmt)), (fmt, i)
        assert np.array_equal(engine.cvt_to_bgr(frames[0], fmt), got[0])
    with pytest.raises(ValueError):
        engine.cvt_to_bgr(np.zeros((5, 4), np.uint8), "nv12")


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["yuy2", "nv12"])
def test_pipeline_on_native_frames_equals_pipeline_on_converted_frames(engine, oracle, fmt):
    from chessboard_vision_b200.engine import grid_rects, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT
    H, W, S, n = 270, 480, 160, 5
    native = np.stack([synth.bgr_to_yuv(synth.board_frame(H, W, i), fmt) for i in range(n)])
    bgr = np.stack([oracle.yuv_to_bgr(native[i], fmt) for i in range(n)])
    M = engine.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
    rects, _ = grid_rects(S)
    pp = engine.pipeline_params(squares=engine.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE | SQ_CD_DETECT),
                                board_size=S)
    engine.set_chunk_frames(2)               # several chunks, the last one short
    try:
        st_a, st_b = engine.new_state(n, S, S), engine.new_state(n, S, S)
        ta, sa = engine.pipeline(native, M, rects, pp, st_a, fmt=fmt)
        tb, sb = engine.pipeline(bgr, M, rects, pp, st_b)
    finally:
        engine.set_chunk_frames(0)
    assert np.array_equal(ta, tb) and sa.tobytes() == sb.tobytes()
    st_a.free(); st_b.free()


@pytest.mark.gpu
def test_submit_wait_equals_the_blocking_pipeline(engine, oracle):
    """cvb_pipeline_submit / cvb_pipeline_wait: batches in flight back to back (formats and chunk layouts changing between
    them, the change-detector state evolving from batch to batch) give what the blocking calls give."""
    from chessboard_vision_b200.engine import (grid_rects, SQ_PD_STATS, SQ_PD_SET_REF, SQ_CD_CALIBRATE, SQ_CD_DETECT,
                                               SQ_CD_UPDATE)
    H, W, S, n = 270, 480, 160, 6
    M = engine.get_perspective_transform(synth.calib_points(H, W), [[0, 0], [S, 0], [0, S], [S, S]])
    rects, _ = grid_rects(S)
    cal = engine.pipeline_params(squares=engine.square_params(ops=SQ_PD_STATS | SQ_PD_SET_REF | SQ_CD_CALIBRATE), board_size=S)
    run = engine.pipeline_params(squares=engine.square_params(ops=SQ_PD_STATS | SQ_CD_DETECT | SQ_CD_UPDATE), board_size=S)
    batches = []
    for k, fmt in enumerate(["bgr", "yuy2", "bgr", "nv12", "nv12", "bgr"]):
        f = np.stack([synth.board_frame(H, W, 10 * k + i) for i in range(n)])
        if fmt != "bgr":
            f = np.stack([synth.bgr_to_yuv(x, fmt) for x in f])
        host = engine.pinned(f.shape)
        host[...] = f
        batches.append((host, fmt, cal if k == 0 else run))
    st_a, st_b = engine.new_state(n, S, S), engine.new_state(n, S, S)
    for chunk in (0, 4):
        engine.set_chunk_frames(chunk)
        try:
            ref = [engine.pipeline(h, M, rects, p, st_a, fmt=fmt) for h, fmt, p in batches]
            tickets = [engine.pipeline_submit(h, M, rects, p, st_b, fmt=fmt) for h, fmt, p in batches]   # all in flight
            got = [engine.pipeline_wait(t) for t in tickets]
        finally:
            engine.set_chunk_frames(0)
        for (ta, sa), (tb, sb) in zip(ref, got):
            assert np.array_equal(ta, tb) and sa.tobytes() == sb.tobytes()
        for plane in range(5):
            assert np.array_equal(st_a.get(0, plane), st_b.get(0, plane))
    with pytest.raises(ValueError):
        engine.pipeline_submit(batches[0][0][:, ::2], M, rects, run, st_b)        # not contiguous: read after the call returns
    st_a.free(); st_b.free()
    for h, _, _ in batches:
        engine.lib.cvb_host_free(h.ctypes.data)
