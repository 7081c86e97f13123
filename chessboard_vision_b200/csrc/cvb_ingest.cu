// Camera ingest: the camera's native YUV frame -> BGR on the device.
//
// Reference: play_lichess.py:16-18,45 and game_session.py:99,113 take BGR frames from cv2.VideoCapture.read(),
// which converts the camera's packed YUY2 (or a decoder's NV12) with cv2.cvtColor(COLOR_YUV2BGR_YUY2 / _NV12) on
// the host.  Doing it here moves 2 (YUY2) or 1.5 (NV12) bytes per pixel over PCIe instead of 3, which is what the
// host-buffer path is bound by.  Integer BT.601 with 20 fractional bits (OpenCV color_yuv.simd.hpp), bit-exact
// against cv2 on all 2^24 (y, u, v) (tests/test_ingest.py).
#include "cvb_device.cuh"

namespace {

CVB_DEV uint32_t yuv_px(int y, int ruv, int guv, int buv)
{
    const int yy = max(0, y - 16) * 1220542;
    return pack_bgr(clamp_u8((yy + buv) >> 20), clamp_u8((yy + guv) >> 20), clamp_u8((yy + ruv) >> 20));
}
CVB_DEV void uv_terms(int u, int v, int &ruv, int &guv, int &buv)
{
    const int uu = u - 128, vv = v - 128;
    ruv = (1 << 19) + 1673527 * vv;
    guv = (1 << 19) - 852492 * vv - 409993 * uu;
    buv = (1 << 19) + 2116026 * uu;
}
// four pixels (b g r, 12 bytes) as three words
CVB_DEV void store4(uint32_t *o, uint32_t q0, uint32_t q1, uint32_t q2, uint32_t q3)
{
    o[0] = q0 | (q1 << 24); o[1] = (q1 >> 8) | (q2 << 16); o[2] = (q2 >> 16) | (q3 << 8);
}

// YUY2: one thread converts 8 pixels = one 16-byte load (Y0 U0 Y1 V0 | Y2 U1 Y3 V1 | ...) -> 24 bytes
__global__ void __launch_bounds__(256) k_yuy2_bgr(const uint8_t *__restrict__ src, long groups, uint8_t *__restrict__ dst)
{
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < groups; i += (long)gridDim.x * 256) {
        const uint4 q = __ldg(reinterpret_cast<const uint4 *>(src) + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
        uint32_t px[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int ruv, guv, buv;
            uv_terms((w[k] >> 8) & 0xff, w[k] >> 24, ruv, guv, buv);
            px[2 * k] = yuv_px(w[k] & 0xff, ruv, guv, buv);
            px[2 * k + 1] = yuv_px((w[k] >> 16) & 0xff, ruv, guv, buv);
        }
        uint32_t o[6];
        store4(o, px[0], px[1], px[2], px[3]); store4(o + 3, px[4], px[5], px[6], px[7]);
        uint2 *op = reinterpret_cast<uint2 *>(dst + i * 24);          // 24 i is a multiple of 8
        op[0] = make_uint2(o[0], o[1]); op[1] = make_uint2(o[2], o[3]); op[2] = make_uint2(o[4], o[5]);
    }
}
// NV12: one thread converts 8 pixels of two rows that share their (U, V) samples
__global__ void __launch_bounds__(256) k_nv12_bgr(const uint8_t *__restrict__ src, int H, int W, uint8_t *__restrict__ dst)
{
    const int frame = blockIdx.y;
    const int gpr = W >> 3;                                   // groups per row
    const long groups = (long)gpr * (H >> 1);
    const uint8_t *fy = src + (size_t)frame * H * W * 3 / 2, *fuv = fy + (size_t)H * W;
    uint8_t *out = dst + (size_t)frame * H * W * 3;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < groups; i += (long)gridDim.x * 256) {
        const int ry = (int)(i / gpr), gx = (int)(i - (long)ry * gpr);
        const uint2 y0 = __ldg(reinterpret_cast<const uint2 *>(fy + (size_t)(2 * ry) * W) + gx);
        const uint2 y1 = __ldg(reinterpret_cast<const uint2 *>(fy + (size_t)(2 * ry + 1) * W) + gx);
        const uint2 uv = __ldg(reinterpret_cast<const uint2 *>(fuv + (size_t)ry * W) + gx);
        const uint32_t uvw[2] = {uv.x, uv.y}, ya[2] = {y0.x, y0.y}, yb[2] = {y1.x, y1.y};
        uint32_t pa[8], pb[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {                          // pixel pair k: bytes 2k (U), 2k + 1 (V) of the UV group
            const uint32_t w = uvw[k >> 1] >> (16 * (k & 1));
            int ruv, guv, buv;
            uv_terms(w & 0xff, (w >> 8) & 0xff, ruv, guv, buv);
            const uint32_t a = ya[k >> 1] >> (16 * (k & 1)), b = yb[k >> 1] >> (16 * (k & 1));
            pa[2 * k] = yuv_px(a & 0xff, ruv, guv, buv); pa[2 * k + 1] = yuv_px((a >> 8) & 0xff, ruv, guv, buv);
            pb[2 * k] = yuv_px(b & 0xff, ruv, guv, buv); pb[2 * k + 1] = yuv_px((b >> 8) & 0xff, ruv, guv, buv);
        }
        uint32_t o[6];
        uint2 *op = reinterpret_cast<uint2 *>(out + ((size_t)(2 * ry) * W + 8 * gx) * 3);
        store4(o, pa[0], pa[1], pa[2], pa[3]); store4(o + 3, pa[4], pa[5], pa[6], pa[7]);
        op[0] = make_uint2(o[0], o[1]); op[1] = make_uint2(o[2], o[3]); op[2] = make_uint2(o[4], o[5]);
        op = reinterpret_cast<uint2 *>(out + ((size_t)(2 * ry + 1) * W + 8 * gx) * 3);
        store4(o, pb[0], pb[1], pb[2], pb[3]); store4(o + 3, pb[4], pb[5], pb[6], pb[7]);
        op[0] = make_uint2(o[0], o[1]); op[1] = make_uint2(o[2], o[3]); op[2] = make_uint2(o[4], o[5]);
    }
}
// any even width (and unaligned pointers): one pixel pair per thread, byte accesses
__global__ void __launch_bounds__(256) k_yuv_bgr_generic(const uint8_t *__restrict__ src, int fmt, int H, int W,
                                                         uint8_t *__restrict__ dst)
{
    const int frame = blockIdx.y;
    const long pairs = (long)H * (W >> 1);
    const size_t fbytes = fmt == CVB_FMT_YUY2 ? (size_t)H * W * 2 : (size_t)H * W * 3 / 2;
    const uint8_t *f = src + (size_t)frame * fbytes;
    uint8_t *out = dst + (size_t)frame * H * W * 3;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < pairs; i += (long)gridDim.x * 256) {
        const int y = (int)(i / (W >> 1)), x = 2 * (int)(i - (long)y * (W >> 1));
        int y0, y1, u, v;
        if (fmt == CVB_FMT_YUY2) {
            const uint8_t *p = f + ((size_t)y * W + x) * 2;
            y0 = p[0]; u = p[1]; y1 = p[2]; v = p[3];
        } else {
            const uint8_t *uv = f + (size_t)H * W + (size_t)(y >> 1) * W + x;
            y0 = f[(size_t)y * W + x]; y1 = f[(size_t)y * W + x + 1]; u = uv[0]; v = uv[1];
        }
        int ruv, guv, buv;
        uv_terms(u, v, ruv, guv, buv);
        const uint32_t q0 = yuv_px(y0, ruv, guv, buv), q1 = yuv_px(y1, ruv, guv, buv);
        uint8_t *o = out + ((size_t)y * W + x) * 3;
        o[0] = (uint8_t)q0; o[1] = (uint8_t)(q0 >> 8); o[2] = (uint8_t)(q0 >> 16);
        o[3] = (uint8_t)q1; o[4] = (uint8_t)(q1 >> 8); o[5] = (uint8_t)(q1 >> 16);
    }
}

}  // namespace

size_t cvb_host_frame_bytes(int format, int H, int W)
{
    const size_t n = (size_t)H * W;
    return format == CVB_FMT_YUY2 ? 2 * n : format == CVB_FMT_NV12 ? n * 3 / 2 : 3 * n;
}

int launch_yuv_to_bgr(cvb_handle *h, const uint8_t *src, int format, int n, int H, int W, uint8_t *bgr)
{
    CVB_REQUIRE(format == CVB_FMT_YUY2 || format == CVB_FMT_NV12, "unknown ingest format %d", format);
    CVB_REQUIRE(W >= 2 && W % 2 == 0, "YUV 4:2:x frames need an even width, got %d", W);
    CVB_REQUIRE(format != CVB_FMT_NV12 || H % 2 == 0, "NV12 frames need an even height, got %d", H);
    const bool aligned = W % 8 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(bgr)) & 15) == 0;
    const int cap = h->sm_count * 16;
    if (aligned && format == CVB_FMT_YUY2) {
        const long groups = (long)n * H * W / 8;              // frames are contiguous and W % 8 == 0: one flat run
        PROF(h, "k_yuv2bgr");
        k_yuy2_bgr<<<(int)std::min<long>((groups + 255) / 256, cap), 256, 0, h->stream>>>(src, groups, bgr);
    } else if (aligned && format == CVB_FMT_NV12 && ((size_t)H * W * 3 / 2) % 16 == 0) {
        const long groups = (long)(W / 8) * (H / 2);
        dim3 grid((unsigned)std::max<long>(1, std::min<long>((groups + 255) / 256, cap / std::max(1, n) + 1)), n);
        PROF(h, "k_yuv2bgr");
        k_nv12_bgr<<<grid, 256, 0, h->stream>>>(src, H, W, bgr);
    } else {
        const long pairs = (long)H * (W / 2);
        dim3 grid((unsigned)std::max<long>(1, std::min<long>((pairs + 255) / 256, cap / std::max(1, n) + 1)), n);
        PROF(h, "k_yuv2bgr");
        k_yuv_bgr_generic<<<grid, 256, 0, h->stream>>>(src, format, H, W, bgr);
    }
    LAUNCH_CHECK(h);
    return CVB_OK;
}

CVB_BOUNDS_TU(ingest)
