// board_detection.warp_image, grid_extractor squares and the per-square
// change_detector / piece_detector statistics as sm_100a kernels.
//
// Reference: board_detection.py:61-71 (warp), grid_extractor.py:33-56,140-161
// (square rectangles), change_detector.py:36-167, piece_detector.py:82-97,
// 124-207,305 (statistics).  Compiled with -fmad=false.
#include "cvb_device.cuh"
#include <cmath>
#include <cstdlib>
#include <algorithm>

static_assert(sizeof(cvb_square_stats) == 128, "cvb_square_stats is part of the ABI");


// ---------------------------------------------------------------------------------------
// cv2.warpPerspective, INTER_LINEAR, BORDER_CONSTANT(0)  (imgwarp.cpp
// WarpPerspectiveInvoker + remapBilinear): destination -> source coordinates in
// f64 with the invoker's 64-column block association, fixed point with 5
// fractional bits, weights (32-ax)(32-ay).., (sum + 512) >> 10.
// ---------------------------------------------------------------------------------------
// One thread per destination column and four destination rows (a CTA covers 64 x 16 pixels): the block set-up and the
// barrier are shared by four pixels and the twelve source bytes of four independent pixels are in flight together
// (the kernel was bound by issue slots and load latency: 204 instructions per pixel, 31 % long-scoreboard stalls).
// Loading the source bytes of all four pixels before any arithmetic was measured too: 92 registers, 3.6 us instead of 2.3.
constexpr int WARP_ROWS = 16;
__global__ void __launch_bounds__(256) k_warp(const uint8_t *__restrict__ src, int H, int W,
                                              const double *__restrict__ minv, int n_mats, int OH, int OW, int rot180,
                                              uint8_t *__restrict__ dst)
{
    // X0, Y0, W0 of a 64-column block row (imgwarp.cpp forms them once per block and adds M * x1 per pixel):
    // computed by one thread per row of the CTA, read back by the 64 threads of that row
    __shared__ double s_row[WARP_ROWS][3];
    __shared__ double s_m[3];
    const int frame = blockIdx.z;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), ty = threadIdx.x >> 6, y0 = blockIdx.y * WARP_ROWS;
    if (threadIdx.x < WARP_ROWS) {
        const double *M = minv + (n_mats == 1 ? 0 : (size_t)frame * 9);
        const double bx = (double)(blockIdx.x * 64), yy = (double)(y0 + threadIdx.x);
        s_row[threadIdx.x][0] = __dadd_rn(__dadd_rn(__dmul_rn(M[0], bx), __dmul_rn(M[1], yy)), M[2]);
        s_row[threadIdx.x][1] = __dadd_rn(__dadd_rn(__dmul_rn(M[3], bx), __dmul_rn(M[4], yy)), M[5]);
        s_row[threadIdx.x][2] = __dadd_rn(__dadd_rn(__dmul_rn(M[6], bx), __dmul_rn(M[7], yy)), M[8]);
        if (threadIdx.x == 0) { s_m[0] = M[0]; s_m[1] = M[3]; s_m[2] = M[6]; }
    }
    __syncthreads();
    if (x >= OW) return;
    const double x1 = (double)(x & 63);
    const double mx0 = __dmul_rn(s_m[0], x1), mx1 = __dmul_rn(s_m[1], x1), mx2 = __dmul_rn(s_m[2], x1);
    const uint8_t *img = src + (size_t)frame * H * W * 3;
#pragma unroll
    for (int k = 0; k < WARP_ROWS / 4; ++k) {
        const int ly = ty + 4 * k, y = y0 + ly;
        if (y >= OH) break;
        double Wd = __dadd_rn(s_row[ly][2], mx2);
        Wd = Wd != 0.0 ? __ddiv_rn(32.0, Wd) : 0.0;
        double fX = __dmul_rn(__dadd_rn(s_row[ly][0], mx0), Wd);
        double fY = __dmul_rn(__dadd_rn(s_row[ly][1], mx1), Wd);
        // imgwarp.cpp clamps fX / fY to [INT_MIN, INT_MAX] with std::min / std::max and rounds (saturate_cast<int>):
        // cvt.rni.s32.f64 saturates to the same two values, and a NaN -- which std::min((double)INT_MAX, NaN) turns into
        // INT_MAX -- is the one case it maps elsewhere (to 0).  Same integers, ~25 instructions less than fmin / fmax.
        int X = __double2int_rn(fX), Y = __double2int_rn(fY);
        if (fX != fX) X = 2147483647;
        if (fY != fY) Y = 2147483647;
        const int sx = max(-32768, min(32767, X >> 5)), sy = max(-32768, min(32767, Y >> 5));
        const int ax = X & 31, ay = Y & 31;
        const int w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay), w10 = (32 - ax) * ay, w11 = ax * ay;
        int acc[3] = {0, 0, 0};
        if (sx >= 0 && sx + 1 < W && sy >= 0 && sy + 1 < H) {
            // all four taps inside (almost every pixel): two row pointers, constant byte offsets
            const uint8_t *p = img + ((size_t)sy * W + sx) * 3, *q = p + (size_t)W * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) acc[c] = w00 * p[c] + w01 * p[3 + c] + w10 * q[c] + w11 * q[3 + c];
        } else {
            const bool x0in = sx >= 0 && sx < W, x1in = sx + 1 >= 0 && sx + 1 < W;
            const bool y0in = sy >= 0 && sy < H, y1in = sy + 1 >= 0 && sy + 1 < H;
            if (y0in) {
                const uint8_t *r = img + (size_t)sy * W * 3;
                if (x0in) { const uint8_t *p = r + (size_t)sx * 3; acc[0] += w00 * p[0]; acc[1] += w00 * p[1]; acc[2] += w00 * p[2]; }
                if (x1in) { const uint8_t *p = r + (size_t)(sx + 1) * 3; acc[0] += w01 * p[0]; acc[1] += w01 * p[1]; acc[2] += w01 * p[2]; }
            }
            if (y1in) {
                const uint8_t *r = img + (size_t)(sy + 1) * W * 3;
                if (x0in) { const uint8_t *p = r + (size_t)sx * 3; acc[0] += w10 * p[0]; acc[1] += w10 * p[1]; acc[2] += w10 * p[2]; }
                if (x1in) { const uint8_t *p = r + (size_t)(sx + 1) * 3; acc[0] += w11 * p[0]; acc[1] += w11 * p[1]; acc[2] += w11 * p[2]; }
            }
        }
        // cv2.rotate(warped, ROTATE_180) (game_session.py:125-126) is a permutation of the destination pixels
        const int oy = rot180 ? OH - 1 - y : y, ox = rot180 ? OW - 1 - x : x;
        uint8_t *o = dst + ((size_t)frame * OH * OW + (size_t)oy * OW + ox) * 3;
        o[0] = (uint8_t)((acc[0] + 512) >> 10);
        o[1] = (uint8_t)((acc[1] + 512) >> 10);
        o[2] = (uint8_t)((acc[2] + 512) >> 10);
    }
}

int launch_warp(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const double *d_minv, int n_mats, int out_h,
                int out_w, int rot180, uint8_t *warped)
{
    dim3 grid((out_w + 63) / 64, (out_h + WARP_ROWS - 1) / WARP_ROWS, n);
    PROF(h, "k_warp");
    k_warp<<<grid, 256, 0, h->stream>>>(bgr, H, W, d_minv, n_mats, out_h, out_w, rot180, warped);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// cv2.rotate (ROTATE_90_CLOCKWISE 0, ROTATE_180 1, ROTATE_90_COUNTERCLOCKWISE 2) of n images with C-byte pixels:
// 32x32-pixel tiles through shared memory so that both the reads and the writes walk rows.
template <int C>
__global__ void __launch_bounds__(256) k_rotate(const uint8_t *__restrict__ src, int H, int W, int code, uint8_t *__restrict__ dst)
{
    __shared__ uint8_t tile[32][32 * C + 4];
    const int frame = blockIdx.z, x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
    const uint8_t *img = src + (size_t)frame * H * W * C;
    uint8_t *out = dst + (size_t)frame * H * W * C;
    const int tw = min(32, W - x0), th = min(32, H - y0);
    for (int i = threadIdx.x; i < th * tw * C; i += 256) {
        const int r = i / (tw * C), b = i - r * (tw * C);
        tile[r][b] = img[((size_t)(y0 + r) * W + x0) * C + b];
    }
    __syncthreads();
    if (code == 1) {                       // dst(y, x) = src(H-1-y, W-1-x): tile rows written right to left
        for (int i = threadIdx.x; i < th * tw * C; i += 256) {
            const int r = i / (tw * C), b = i - r * (tw * C), px = b / C, ch = b - px * C;
            // destination row H-1-(y0+r), destination pixels W-1-(x0+tw-1) ... ascending
            out[((size_t)(H - 1 - y0 - r) * W + (W - x0 - tw) + px) * C + ch] = tile[r][(tw - 1 - px) * C + ch];
        }
    } else {                               // output is W rows of H pixels
        for (int i = threadIdx.x; i < th * tw * C; i += 256) {
            const int orow = i / (th * C), b = i - orow * (th * C), px = b / C, ch = b - px * C;
            // clockwise: dst(x, H-1-y) = src(y, x); counter-clockwise: dst(W-1-x, y) = src(y, x)
            if (code == 0)
                out[((size_t)(x0 + orow) * H + (H - y0 - th) + px) * C + ch] = tile[th - 1 - px][orow * C + ch];
            else
                out[((size_t)(W - 1 - x0 - orow) * H + y0 + px) * C + ch] = tile[px][orow * C + ch];
        }
    }
}

int launch_rotate(cvb_handle *h, const uint8_t *src, int n, int H, int W, int C, int code, uint8_t *dst)
{
    dim3 grid((W + 31) / 32, (H + 31) / 32, n);
    PROF(h, "k_rotate");
    if (C == 3) k_rotate<3><<<grid, 256, 0, h->stream>>>(src, H, W, code, dst);
    else k_rotate<1><<<grid, 256, 0, h->stream>>>(src, H, W, code, dst);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// ---------------------------------------------------------------------------------------
// Per-square kernel: one block per (square, frame).
//   gray (S7) -> Gaussian k (S8, borders reflected inside the square, as the
//   reference blurs each square view on its own) -> fused reductions:
//     PieceDetector: sum, sum^2, SAD vs reference, centre/corner/ring sums
//     ChangeDetector: f32 z-score count/max, EMA update of mean/variance
// ---------------------------------------------------------------------------------------
struct SquareArgs {
    const uint8_t *boards;
    int BH, BW, C;
    const cvb_rect *rects;
    const int32_t *mask_ofs;
    const uint8_t *masks;
    const uint8_t *select;
    uint8_t *pd_ref, *pd_cur, *flags;
    float *cd_mean, *cd_var;
    int stream0;
    cvb_square_params p;
    int pd_q[31], cd_q[31];
    cvb_square_stats *stats;
};

// single-bounce REFLECT_101, valid while the overshoot is smaller than n
CVB_DEV int reflect_near(int p, int n)
{
    p = p < 0 ? -p : p;
    return p >= n ? 2 * n - 2 - p : p;
}

// horizontal Q8 pass over the whole square; pixels are dealt to the threads in row-major order so that a
// square of any width keeps all lanes busy
template <int K>
CVB_DEV void hpass(const uint8_t *s_g, uint16_t *s_h, const int *q, int k_rt, int w, int h, unsigned inv_w, bool near_ok)
{
    const int k = K ? K : k_rt, r = k >> 1, n = w * h;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int y = div_magic(i, inv_w), x = i - y * w;
        const uint8_t *row = s_g + y * w;
        uint32_t s = 0;
        if (K == 5) {
            if (x >= 2 && x + 2 < w) {
                s = q[0] * row[x - 2] + q[1] * row[x - 1] + q[2] * row[x] + q[3] * row[x + 1] + q[4] * row[x + 2];
            } else {
#pragma unroll
                for (int j = 0; j < 5; ++j)
                    s += (uint32_t)q[j] * row[near_ok ? reflect_near(x + j - 2, w) : reflect101(x + j - 2, w)];
            }
        } else {
            for (int j = 0; j < k; ++j)
                s += (uint32_t)q[j] * row[near_ok ? reflect_near(x + j - r, w) : reflect101(x + j - r, w)];
        }
        s_h[i] = (uint16_t)s;
    }
}
template <int K>
CVB_DEV int blur_at(const uint16_t *s_h, const int *q, int k_rt, int x, int y, int w, int h, bool near_ok)
{
    const int k = K ? K : k_rt, r = k >> 1;
    uint32_t s = 0;
    if (K == 5) {
        if (y >= 2 && y + 2 < h) {
            const uint16_t *c = s_h + (y - 2) * w + x;
            s = q[0] * c[0] + q[1] * c[w] + q[2] * c[2 * w] + q[3] * c[3 * w] + q[4] * c[4 * w];
        } else {
#pragma unroll
            for (int i = 0; i < 5; ++i)
                s += (uint32_t)q[i] * s_h[(near_ok ? reflect_near(y + i - 2, h) : reflect101(y + i - 2, h)) * w + x];
        }
    } else {
        for (int i = 0; i < k; ++i)
            s += (uint32_t)q[i] * s_h[(near_ok ? reflect_near(y + i - r, h) : reflect101(y + i - r, h)) * w + x];
    }
    return (int)((s + 32768u) >> 16);
}

// |g - m| / sqrt(v) in IEEE f32 (change_detector.py:128-129).  A zero numerator is common (a static square whose mean has
// converged) and sends div.rn's special-operand path through a ~30-instruction subroutine for the whole warp: its
// quotient is known without dividing (+0 for v > 0 incl. +inf, NaN for v <= 0, -0 or NaN: 0 / sqrt(v) in every case).
CVB_DEV float cd_zscore(float gf, float m, float v)
{
    const float d = fabsf(__fsub_rn(gf, m));
    if (d == 0.0f) return v > 0.0f ? 0.0f : __int_as_float(0x7fc00000);
    return __fdiv_rn(d, __fsqrt_rn(v));
}

template <int OPS>
// measured on B200 (77-px squares): pixels per thread and round / resident CTAs per SM 4/4: 4.66, 2/6: 4.38, 3/5: 4.60,
// 1/8: 4.55, 4/3: 5.83 us per frame -- occupancy beats batch depth here
#ifndef CVB_SQ_U
#define CVB_SQ_U 2
#endif
#ifndef CVB_SQ_MINB
#define CVB_SQ_MINB 6
#endif
__global__ void __launch_bounds__(256, CVB_SQ_MINB) k_squares(const SquareArgs a)
{
    extern __shared__ __align__(16) uint8_t sq_smem[];
    __shared__ unsigned long long s_acc[16];
    __shared__ unsigned s_cd_nan;
    __shared__ int s_cd_zbits;
    const int tid = threadIdx.x, lane = tid & 31;
    const int sq = blockIdx.x, frame = blockIdx.y;
    const cvb_rect rc = a.rects[sq];
    const int w = rc.w, h = rc.h, n = w * h;
    uint8_t *s_g = sq_smem;
    uint16_t *s_h = reinterpret_cast<uint16_t *>(sq_smem + ((n + 15) & ~15));
    const size_t plane = (size_t)a.BH * a.BW;
    const size_t so = (size_t)(a.stream0 + frame) * plane;       // state slot offset
    const uint8_t *board = a.boards + (size_t)frame * plane * a.C;
    const bool selected = a.select ? a.select[sq] != 0 : true;
    const int ops = OPS >= 0 ? OPS : a.p.ops;     // compile-time mask strips the per-pixel flag tests

    if (tid < 16) s_acc[tid] = 0ull;
    if (tid == 0) { s_cd_nan = 0; s_cd_zbits = __float_as_int(-INFINITY); }
    const unsigned inv_w = magic_of(w);
    {
        const uint8_t *org = board + ((size_t)rc.y * a.BW + rc.x) * a.C;
        const int pitch = a.BW * a.C;
        if (a.C == 3) {
            for (int i0 = tid; i0 < n; i0 += 256 * 4) {
                int c0[4], c1[4], c2[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + 256 * j;
                    if (i < n) {
                        const int y = div_magic(i, inv_w), x = i - y * w;
                        const uint8_t *p = org + (size_t)y * pitch + 3 * x;
                        c0[j] = __ldg(p); c1[j] = __ldg(p + 1); c2[j] = __ldg(p + 2);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int i = i0 + 256 * j;
                    CVB_BOUNDS(i >= n || (rc.y + div_magic(i, inv_w) < a.BH && rc.x + (i - div_magic(i, inv_w) * w) < a.BW));
                    if (i < n) s_g[i] = (uint8_t)gray_px(c0[j], c1[j], c2[j]);
                }
            }
        } else {
            for (int i = tid; i < n; i += 256) {
                const int y = div_magic(i, inv_w), x = i - y * w;
                s_g[i] = __ldg(org + (size_t)y * pitch + x);
            }
        }
    }
    __syncthreads();

    const size_t first = so + (size_t)rc.y * a.BW + rc.x;
    const bool state = a.flags != nullptr;
    const int fl0 = state ? a.flags[first] : 0;
    const bool has_ref = (fl0 & 1) != 0;
    const bool has_cd = (fl0 & 6) == 6;         // bit 1: a mean is stored, bit 2: a variance is stored

    const bool need_pd = (ops & (CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF)) != 0;
    const bool need_cd = (ops & (CVB_SQ_CD_CALIBRATE | CVB_SQ_CD_DETECT | CVB_SQ_CD_UPDATE)) != 0 && selected && state;
    const bool same_blur = a.p.pd_blur == a.p.cd_blur;
    const bool pass1 = need_pd || (need_cd && same_blur);
    const bool near_pd = w > (a.p.pd_blur >> 1) && h > (a.p.pd_blur >> 1);
    const bool near_cd = w > (a.p.cd_blur >> 1) && h > (a.p.cd_blur >> 1);
    if (pass1) {
        if (a.p.pd_blur == 5) hpass<5>(s_g, s_h, a.pd_q, 5, w, h, inv_w, near_pd);
        else hpass<0>(s_g, s_h, a.pd_q, a.p.pd_blur, w, h, inv_w, near_pd);
    }
    __syncthreads();

    // state planes of this stream slot, addressed from the square's first pixel with 32-bit offsets
    const size_t org = so + (size_t)rc.y * a.BW + rc.x;
    uint8_t *const pd_ref = a.pd_ref ? a.pd_ref + org : nullptr, *const pd_cur = a.pd_cur ? a.pd_cur + org : nullptr,
                   *const flags = a.flags ? a.flags + org : nullptr;
    float *const cd_mean = a.cd_mean ? a.cd_mean + org : nullptr, *const cd_var = a.cd_var ? a.cd_var + org : nullptr;
    const float zthr = a.p.z_threshold, alpha = a.p.alpha, oma = a.p.one_minus_alpha, minvar = a.p.min_variance,
                initvar = a.p.initial_variance;
    unsigned sum = 0, sad = 0, csum = 0, ccnt = 0, bsum = 0, bcnt = 0;
    unsigned rsum[4] = {0, 0, 0, 0}, rcnt[4] = {0, 0, 0, 0};
    unsigned long long sumsq = 0;
    unsigned cd_cnt = 0;
    float cd_zmax = -INFINITY;
    bool cd_nan = false;
    const uint8_t *mask = a.masks + a.mask_ofs[sq];

    // Pixels are dealt to the threads in row-major order, U per thread and round: all state loads of a batch
    // are issued before any arithmetic or store, so several memory round trips overlap (the loop is latency bound).
    constexpr int U = CVB_SQ_U;
    auto cd_batch = [&](const int (&gv)[U], const bool (&ok)[U], const unsigned (&ofs)[U]) {
        float m[U], v[U];
        const bool calib = (ops & CVB_SQ_CD_CALIBRATE) != 0;
        if (!calib && !has_cd) return;
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const unsigned o = ofs[j];
            if (calib) { m[j] = (float)gv[j]; v[j] = initvar; }
            else if (ok[j]) { m[j] = cd_mean[o]; v[j] = cd_var[o]; }
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            if (!ok[j]) continue;
            const unsigned o = ofs[j];
            const float gf = (float)gv[j];
            if (calib) { cd_mean[o] = m[j]; cd_var[o] = v[j]; flags[o] |= 6; }
            if (ops & CVB_SQ_CD_DETECT) {
                const float z = cd_zscore(gf, m[j], v[j]);
                if (z > zthr) ++cd_cnt;
                if (z != z) cd_nan = true; else cd_zmax = fmaxf(cd_zmax, z);
            }
            if (ops & CVB_SQ_CD_UPDATE) {
                // change_detector.py:82-89, every product and sum rounded on its own
                const float nm = __fadd_rn(__fmul_rn(oma, m[j]), __fmul_rn(alpha, gf));
                const float d = __fsub_rn(gf, nm);
                float nv = __fadd_rn(__fmul_rn(oma, v[j]), __fmul_rn(alpha, __fmul_rn(d, d)));
                if (!(nv > minvar) && nv == nv) nv = minvar;   // np.maximum keeps NaN
                cd_mean[o] = nm; cd_var[o] = nv;
            }
        }
    };

    if (pass1) {
        const bool k5 = a.p.pd_blur == 5;
        for (int i0 = tid; i0 < n; i0 += 256 * U) {
            int gv[U], mk[U], rf[U];
            bool ok[U];
            unsigned ofs[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int i = i0 + 256 * j;
                ok[j] = i < n;
                mk[j] = 0; rf[j] = 0; gv[j] = 0; ofs[j] = 0;
                if (ok[j]) {
                    const int y = div_magic(i, inv_w), x = i - y * w;
                    ofs[j] = (unsigned)(y * a.BW + x);
                    if (ops & CVB_SQ_PD_STATS) {
                        mk[j] = mask[i];
                        if (has_ref) rf[j] = pd_ref[ofs[j]];
                    }
                    gv[j] = k5 ? blur_at<5>(s_h, a.pd_q, 5, x, y, w, h, near_pd)
                               : blur_at<0>(s_h, a.pd_q, a.p.pd_blur, x, y, w, h, near_pd);
                }
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                if (!ok[j]) continue;
                const unsigned o = ofs[j];
                const int g1 = gv[j];
                if (ops & CVB_SQ_PD_STATS) {
                    const int m = mk[j];
                    sum += g1; sumsq += (unsigned)(g1 * g1);
                    if (has_ref) sad += (unsigned)abs(g1 - rf[j]);
                    if (m & 1) { csum += g1; ++ccnt; }
                    if (m & 2) { bsum += g1; ++bcnt; }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (m & (4 << k)) { rsum[k] += g1; ++rcnt[k]; }
                }
                if (state && need_pd) pd_cur[o] = (uint8_t)g1;
                if ((ops & CVB_SQ_PD_SET_REF) && selected && state) { pd_ref[o] = (uint8_t)g1; flags[o] |= 1; }
            }
            if (need_cd && same_blur) cd_batch(gv, ok, ofs);
        }
    }
    // ---- ChangeDetector pass when its blur differs ----
    if (need_cd && !same_blur) {
        __syncthreads();
        if (a.p.cd_blur == 5) hpass<5>(s_g, s_h, a.cd_q, 5, w, h, inv_w, near_cd);
        else hpass<0>(s_g, s_h, a.cd_q, a.p.cd_blur, w, h, inv_w, near_cd);
        __syncthreads();
        for (int i0 = tid; i0 < n; i0 += 256 * U) {
            int gv[U];
            bool ok[U];
            unsigned ofs[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const int i = i0 + 256 * j;
                ok[j] = i < n;
                gv[j] = 0; ofs[j] = 0;
                if (ok[j]) {
                    const int y = div_magic(i, inv_w), x = i - y * w;
                    ofs[j] = (unsigned)(y * a.BW + x);
                    gv[j] = a.p.cd_blur == 5 ? blur_at<5>(s_h, a.cd_q, 5, x, y, w, h, near_cd)
                                             : blur_at<0>(s_h, a.cd_q, a.p.cd_blur, x, y, w, h, near_cd);
                }
            }
            cd_batch(gv, ok, ofs);
        }
    }
    if (!a.stats) return;

    // ---- block reduction ----
    unsigned long long vals[16] = {sum, sumsq, sad, csum, ccnt, bsum, bcnt, rsum[0], rsum[1], rsum[2], rsum[3],
                                   rcnt[0], rcnt[1], rcnt[2], rcnt[3], cd_cnt};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const unsigned long long r = warp_sum_ull(vals[k]);
        if (lane == 0 && r) atomicAdd(&s_acc[k], r);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cd_zmax = fmaxf(cd_zmax, __shfl_xor_sync(0xffffffffu, cd_zmax, o));
    const unsigned any_nan = __ballot_sync(0xffffffffu, cd_nan);
    if (lane == 0) {
        // non-NaN z is >= +0, so its bit pattern orders like a signed int; -inf (no pixel) is negative
        atomicMax(&s_cd_zbits, __float_as_int(cd_zmax));
        if (any_nan) atomicOr(&s_cd_nan, 1u);
    }
    __syncthreads();
    if (tid == 0) {
        cvb_square_stats st;
        memset(&st, 0, sizeof st);
        st.n = n;
        st.has_ref = has_ref ? 1 : 0;
        st.sum = (uint32_t)s_acc[0]; st.sumsq = s_acc[1]; st.sad = (uint32_t)s_acc[2];
        st.center_sum = (uint32_t)s_acc[3]; st.center_cnt = (uint32_t)s_acc[4];
        st.border_sum = (uint32_t)s_acc[5]; st.border_cnt = (uint32_t)s_acc[6];
        for (int k = 0; k < 4; ++k) { st.ring_sum[k] = (uint32_t)s_acc[7 + k]; st.ring_cnt[k] = (uint32_t)s_acc[11 + k]; }
        const bool cd_ran = need_cd && (ops & CVB_SQ_CD_DETECT) && (has_cd || (ops & CVB_SQ_CD_CALIBRATE));
        st.cd_valid = cd_ran ? 1 : 0;
        st.cd_changed = (int32_t)s_acc[15];
        st.cd_zmax = s_cd_nan ? __int_as_float(0x7fc00000) : __int_as_float(s_cd_zbits);   // np.max propagates NaN
        a.stats[(size_t)frame * gridDim.x + sq] = st;
    }
}

// ---------------------------------------------------------------------------------------
// The same per-square work for the case every live caller uses (3-channel boards, both blurs 5 x 5, board pitch a
// multiple of 4 pixels), four pixels per thread.  Groups of 4 pixels are aligned to the BOARD's x grid (x % 4 == 0), not
// to the square: a group of BGR pixels is then three aligned words, the state planes move as one word (references,
// gray squares, flags) or one float4 (mean, variance) per group, and the statistics become byte-wise dot products
// (sum = dp4a(g, 1), sum of squares = dp4a(g, g), SAD = vsadu4, masked region sums = dp4a(g, mask bit)).  Pixels of a
// group that lie outside the square (its first / last group when the square does not start on the grid) are masked
// out of every sum and never stored.  Gaussian 5 x 5 as in k_finish: rows by dp4a on packed bytes, columns on packed
// 16-bit lanes, (sum + 128) >> 8 -- the Q8 x Q8 form of smooth.dispatch.cpp with q = 16 [1 4 6 4 1] reduces to it exactly.
// REFLECT_101 inside the square: two mirrored rows above / below are staged from their source rows, two mirrored
// columns left / right are copied in shared memory after the gray pass.
// ---------------------------------------------------------------------------------------
template <int OPS>
#ifndef CVB_SQ4_MINB
#define CVB_SQ4_MINB 6
#endif
__global__ void __launch_bounds__(256, CVB_SQ4_MINB) k_squares4(const SquareArgs a)
{
    extern __shared__ __align__(16) uint8_t sq_smem[];
    __shared__ unsigned long long s_acc[16];
    __shared__ unsigned s_cd_nan;
    __shared__ int s_cd_zbits;
    const int tid = threadIdx.x, lane = tid & 31;
    const int sq = blockIdx.x, frame = blockIdx.y;
    const cvb_rect rc = a.rects[sq];
    const int w = rc.w, h = rc.h, n = w * h;
    const int X0 = rc.x & ~3, X1 = (rc.x + w + 3) & ~3, span = X1 - X0, G = span >> 2;
    const int R = h + 4;                                   // staged rows: local y = -2 .. h + 1
    const int pitchG = span + 8;                           // gray bytes per staged row: board x = X0 - 4 .. X1 + 3
    uint8_t *s_g = sq_smem;
    uint16_t *s_h = reinterpret_cast<uint16_t *>(sq_smem + (((size_t)R * pitchG + 15) & ~(size_t)15));   // [R][span]
    const size_t plane = (size_t)a.BH * a.BW;
    const size_t so = (size_t)(a.stream0 + frame) * plane;
    const uint8_t *board = a.boards + (size_t)frame * plane * 3;
    const bool selected = a.select ? a.select[sq] != 0 : true;
    constexpr int ops = OPS;

    if (tid < 16) s_acc[tid] = 0ull;
    if (tid == 0) { s_cd_nan = 0; s_cd_zbits = __float_as_int(-INFINITY); }
    const unsigned inv_g = magic_of(G);
    // ---- 1: gray of the staged rows, whole aligned groups (neighbouring squares' pixels inside a group are harmless) ----
    for (int i = tid; i < R * G; i += 256) {
        const int r = div_magic(i, inv_g), g = i - r * G;
        const int ys = rc.y + reflect_near(r - 2, h);
        const uint32_t *p = reinterpret_cast<const uint32_t *>(board + ((size_t)ys * a.BW + X0 + 4 * g) * 3);
        CVB_BOUNDS(ys >= 0 && ys < a.BH && X0 + 4 * g + 3 < a.BW && r * pitchG + 4 + 4 * g + 3 < R * pitchG);
        const uint32_t w0 = __ldg(p), w1 = __ldg(p + 1), w2 = __ldg(p + 2);
        *reinterpret_cast<uint32_t *>(s_g + r * pitchG + 4 + 4 * g) = gray4_from_words(w0, w1, w2);
    }
    __syncthreads();
    // ---- mirrored columns (local x = -1, -2, w, w + 1) ----
    {
        const int cl = rc.x - X0 + 4, cr = cl + w - 1;
        for (int r = tid; r < R; r += 256) {
            uint8_t *row = s_g + r * pitchG;
            const uint8_t l1 = row[cl + 1], l2 = row[cl + 2], r1 = row[cr - 1], r2 = row[cr - 2];
            row[cl - 1] = l1; row[cl - 2] = l2; row[cr + 1] = r1; row[cr + 2] = r2;
        }
    }
    __syncthreads();
    // ---- 2: horizontal [1 4 6 4 1], four outputs per item ----
    for (int i = tid; i < R * G; i += 256) {
        const int r = div_magic(i, inv_g), g = i - r * G;
        const uint32_t *gp = reinterpret_cast<const uint32_t *>(s_g + r * pitchG + 4 * g);
        const uint32_t wa = gp[0], wb = gp[1], wc = gp[2];          // board x = X0 + 4g - 4 .. X0 + 4g + 7
        constexpr uint32_t kW = 1u | (4u << 8) | (6u << 16) | (4u << 24);
        const uint32_t h0 = __dp4a(__byte_perm(wa, wb, 0x5432), kW, (wb >> 16) & 0xffu);
        const uint32_t h1 = __dp4a(__byte_perm(wa, wb, 0x6543), kW, wb >> 24);
        const uint32_t h2 = __dp4a(wb, kW, wc & 0xffu);
        const uint32_t h3 = __dp4a(__byte_perm(wb, wc, 0x4321), kW, (wc >> 8) & 0xffu);
        *reinterpret_cast<uint2 *>(s_h + r * span + 4 * g) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
    }
    __syncthreads();

    const size_t first = so + (size_t)rc.y * a.BW + rc.x;
    const bool state = a.flags != nullptr;
    const int fl0 = state ? a.flags[first] : 0;
    const bool has_ref = (fl0 & 1) != 0;
    const bool has_cd = (fl0 & 6) == 6;
    const bool need_pd = (ops & (CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF)) != 0;
    const bool need_cd = (ops & (CVB_SQ_CD_CALIBRATE | CVB_SQ_CD_DETECT | CVB_SQ_CD_UPDATE)) != 0 && selected && state;
    const bool calib = (ops & CVB_SQ_CD_CALIBRATE) != 0;
    const bool cd_live = need_cd && (calib || has_cd);
    const float zthr = a.p.z_threshold, alpha = a.p.alpha, oma = a.p.one_minus_alpha, minvar = a.p.min_variance,
                initvar = a.p.initial_variance;
    unsigned sum = 0, sad = 0, rsum[6] = {0, 0, 0, 0, 0, 0}, rcnt[6] = {0, 0, 0, 0, 0, 0};
    unsigned long long sumsq = 0;
    unsigned cd_cnt = 0;
    float cd_zmax = -INFINITY;
    bool cd_nan = false;
    const uint8_t *mask = a.masks + a.mask_ofs[sq];
    // ---- 3: vertical pass, statistics, state ----
    for (int i = tid; i < h * G; i += 256) {
        const int y = div_magic(i, inv_g), g = i - y * G;
        const uint16_t *hrow = s_h + y * span + 4 * g;              // staged row y is local y - 2
        const uint2 r0 = *reinterpret_cast<const uint2 *>(hrow), r1 = *reinterpret_cast<const uint2 *>(hrow + span),
                    r2 = *reinterpret_cast<const uint2 *>(hrow + 2 * span), r3 = *reinterpret_cast<const uint2 *>(hrow + 3 * span),
                    r4 = *reinterpret_cast<const uint2 *>(hrow + 4 * span);
        const uint32_t lo = r0.x + 4 * r1.x + 6 * r2.x + 4 * r3.x + r4.x, hi = r0.y + 4 * r1.y + 6 * r2.y + 4 * r3.y + r4.y;
        const uint32_t o0 = ((lo & 0xffff) + 128) >> 8, o1 = ((lo >> 16) + 128) >> 8, o2 = ((hi & 0xffff) + 128) >> 8,
                       o3 = ((hi >> 16) + 128) >> 8;
        const int X = X0 + 4 * g, xl = X - rc.x;                    // local x of the group's first pixel (may be -3 .. -1)
        uint32_t vbytes = 0;                                        // 0xff in the byte lanes of pixels inside the square
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (xl + j >= 0 && xl + j < w) vbytes |= 0xffu << (8 * j);
        if (!vbytes) continue;
        const bool full = vbytes == 0xffffffffu;
        const uint32_t gw = (o0 | (o1 << 8) | (o2 << 16) | (o3 << 24)) & vbytes;
        const unsigned ofs = (unsigned)((rc.y + y) * a.BW + X);      // offset inside a state plane (multiple of 4)
        CVB_BOUNDS(rc.y + y < a.BH && X + 3 < a.BW && (ofs & 3u) == 0u);
        if (ops & CVB_SQ_PD_STATS) {
            uint32_t mw = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (vbytes & (0xffu << (8 * j))) mw |= (uint32_t)mask[y * w + xl + j] << (8 * j);
            sum += __dp4a(gw, 0x01010101u, 0u);
            sumsq += __dp4a(gw, gw, 0u);
            if (has_ref) sad += __vsadu4(gw, *reinterpret_cast<const uint32_t *>(a.pd_ref + so + ofs) & vbytes);
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const uint32_t sel = (mw >> k) & 0x01010101u;
                rsum[k] += __dp4a(gw, sel, 0u);
                rcnt[k] += __popc(sel);
            }
        }
        if (state && need_pd) {
            uint8_t *pc = a.pd_cur + so + ofs;
            const bool setref = (ops & CVB_SQ_PD_SET_REF) && selected;
            if (full) {
                *reinterpret_cast<uint32_t *>(pc) = gw;
                if (setref) {
                    *reinterpret_cast<uint32_t *>(a.pd_ref + so + ofs) = gw;
                    uint32_t *fp = reinterpret_cast<uint32_t *>(a.flags + so + ofs);
                    *fp = *fp | 0x01010101u;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (vbytes & (0xffu << (8 * j))) {
                        const uint8_t v = (uint8_t)(gw >> (8 * j));
                        pc[j] = v;
                        if (setref) { a.pd_ref[so + ofs + j] = v; a.flags[so + ofs + j] |= 1; }
                    }
            }
        }
        if (cd_live) {
            float *pm = a.cd_mean + so + ofs, *pv = a.cd_var + so + ofs;
            float m[4], v[4];
            if (calib) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { m[j] = (float)((gw >> (8 * j)) & 0xffu); v[j] = initvar; }
            } else if (full) {
                const float4 m4 = *reinterpret_cast<const float4 *>(pm), v4 = *reinterpret_cast<const float4 *>(pv);
                m[0] = m4.x; m[1] = m4.y; m[2] = m4.z; m[3] = m4.w; v[0] = v4.x; v[1] = v4.y; v[2] = v4.z; v[3] = v4.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    m[j] = 0.f; v[j] = 1.f;
                    if (vbytes & (0xffu << (8 * j))) { m[j] = pm[j]; v[j] = pv[j]; }
                }
            }
            float nm[4], nv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                nm[j] = m[j]; nv[j] = v[j];
                if (!(vbytes & (0xffu << (8 * j)))) continue;
                const float gf = (float)((gw >> (8 * j)) & 0xffu);
                if (ops & CVB_SQ_CD_DETECT) {
                    const float z = cd_zscore(gf, m[j], v[j]);
                    if (z > zthr) ++cd_cnt;
                    if (z != z) cd_nan = true; else cd_zmax = fmaxf(cd_zmax, z);
                }
                if (ops & CVB_SQ_CD_UPDATE) {
                    // change_detector.py:82-89, every product and sum rounded on its own
                    nm[j] = __fadd_rn(__fmul_rn(oma, m[j]), __fmul_rn(alpha, gf));
                    const float d = __fsub_rn(gf, nm[j]);
                    float t = __fadd_rn(__fmul_rn(oma, v[j]), __fmul_rn(alpha, __fmul_rn(d, d)));
                    if (!(t > minvar) && t == t) t = minvar;       // np.maximum keeps NaN
                    nv[j] = t;
                }
            }
            if (calib || (ops & CVB_SQ_CD_UPDATE)) {
                if (full) {
                    *reinterpret_cast<float4 *>(pm) = make_float4(nm[0], nm[1], nm[2], nm[3]);
                    *reinterpret_cast<float4 *>(pv) = make_float4(nv[0], nv[1], nv[2], nv[3]);
                    if (calib) {
                        uint32_t *fp = reinterpret_cast<uint32_t *>(a.flags + so + ofs);
                        *fp = *fp | 0x06060606u;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (vbytes & (0xffu << (8 * j))) {
                            pm[j] = nm[j]; pv[j] = nv[j];
                            if (calib) a.flags[so + ofs + j] |= 6;
                        }
                }
            }
        }
    }
    if (!a.stats) return;

    // ---- block reduction (same record as k_squares) ----
    unsigned long long vals[16] = {sum, sumsq, sad, rsum[0], rcnt[0], rsum[1], rcnt[1], rsum[2], rsum[3], rsum[4], rsum[5],
                                   rcnt[2], rcnt[3], rcnt[4], rcnt[5], cd_cnt};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const unsigned long long r = warp_sum_ull(vals[k]);
        if (lane == 0 && r) atomicAdd(&s_acc[k], r);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cd_zmax = fmaxf(cd_zmax, __shfl_xor_sync(0xffffffffu, cd_zmax, o));
    const unsigned any_nan = __ballot_sync(0xffffffffu, cd_nan);
    if (lane == 0) {
        atomicMax(&s_cd_zbits, __float_as_int(cd_zmax));
        if (any_nan) atomicOr(&s_cd_nan, 1u);
    }
    __syncthreads();
    if (tid == 0) {
        cvb_square_stats st;
        memset(&st, 0, sizeof st);
        st.n = n;
        st.has_ref = has_ref ? 1 : 0;
        st.sum = (uint32_t)s_acc[0]; st.sumsq = s_acc[1]; st.sad = (uint32_t)s_acc[2];
        st.center_sum = (uint32_t)s_acc[3]; st.center_cnt = (uint32_t)s_acc[4];
        st.border_sum = (uint32_t)s_acc[5]; st.border_cnt = (uint32_t)s_acc[6];
        for (int k = 0; k < 4; ++k) { st.ring_sum[k] = (uint32_t)s_acc[7 + k]; st.ring_cnt[k] = (uint32_t)s_acc[11 + k]; }
        const bool cd_ran = need_cd && (ops & CVB_SQ_CD_DETECT) && (has_cd || calib);
        st.cd_valid = cd_ran ? 1 : 0;
        st.cd_changed = (int32_t)s_acc[15];
        st.cd_zmax = s_cd_nan ? __int_as_float(0x7fc00000) : __int_as_float(s_cd_zbits);
        a.stats[(size_t)frame * gridDim.x + sq] = st;
    }
}

int launch_squares(cvb_handle *h, const uint8_t *boards, int n, int BH, int BW, int C, const cvb_rect *d_rects,
                   const int32_t *d_mask_ofs, const uint8_t *d_masks, int n_sq, int max_px, const uint8_t *d_select,
                   cvb_state *st, int stream0, const cvb_square_params &p, const int *pd_q, const int *cd_q,
                   cvb_square_stats *stats)
{
    SquareArgs a;
    memset(&a, 0, sizeof a);
    a.boards = boards; a.BH = BH; a.BW = BW; a.C = C; a.rects = d_rects; a.mask_ofs = d_mask_ofs; a.masks = d_masks;
    a.select = d_select; a.stream0 = stream0; a.p = p; a.stats = stats;
    if (st) { a.pd_ref = st->pd_ref; a.pd_cur = st->pd_cur; a.flags = st->flags; a.cd_mean = st->cd_mean; a.cd_var = st->cd_var; }
    memcpy(a.pd_q, pd_q, sizeof a.pd_q);
    memcpy(a.cd_q, cd_q, sizeof a.cd_q);
    const size_t smem = (((size_t)max_px + 15) & ~(size_t)15) + (size_t)max_px * 2;
    if (smem > 200 * 1024) {
        cvb_set_error("square of %d pixels needs %zu bytes of shared memory (max 204800)", max_px, smem);
        return CVB_ERR_INVALID;
    }
    dim3 grid(n_sq, n);
    // four-pixels-per-thread form: 3-channel boards, both blurs 5 x 5, board rows and planes that keep 4-pixel groups
    // aligned, squares of at least 3 x 3 pixels (single-bounce mirrors), one of the op masks the callers use
    {
        bool fast = C == 3 && p.pd_blur == 5 && p.cd_blur == 5 && BW % 4 == 0 && ((size_t)BH * BW) % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(boards) & 3) == 0 && !getenv("CVB_SQUARES_OLD");
        int max_w = 0, max_h = 0;
        for (const cvb_rect &r : h->rect_cache.rects) {
            fast = fast && r.w >= 3 && r.h >= 3;
            max_w = std::max(max_w, r.w); max_h = std::max(max_h, r.h);
        }
        fast = fast && (int)h->rect_cache.rects.size() == n_sq;
        void (*k4)(const SquareArgs) = nullptr;
        switch (p.ops) {
        case CVB_SQ_PD_STATS: k4 = k_squares4<CVB_SQ_PD_STATS>; break;
        case CVB_SQ_PD_SET_REF: k4 = k_squares4<CVB_SQ_PD_SET_REF>; break;
        case CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF: k4 = k_squares4<CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF>; break;
        case CVB_SQ_CD_CALIBRATE: k4 = k_squares4<CVB_SQ_CD_CALIBRATE>; break;
        case CVB_SQ_CD_DETECT: k4 = k_squares4<CVB_SQ_CD_DETECT>; break;
        case CVB_SQ_CD_UPDATE: k4 = k_squares4<CVB_SQ_CD_UPDATE>; break;
        case CVB_SQ_PD_STATS | CVB_SQ_CD_DETECT | CVB_SQ_CD_UPDATE:
            k4 = k_squares4<CVB_SQ_PD_STATS | CVB_SQ_CD_DETECT | CVB_SQ_CD_UPDATE>; break;
        case CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF | CVB_SQ_CD_CALIBRATE:
            k4 = k_squares4<CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF | CVB_SQ_CD_CALIBRATE>; break;
        case CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF | CVB_SQ_CD_CALIBRATE | CVB_SQ_CD_DETECT:
            k4 = k_squares4<CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF | CVB_SQ_CD_CALIBRATE | CVB_SQ_CD_DETECT>; break;
        default: break;
        }
        if (fast && k4) {
            const int span = max_w + 6, rows = max_h + 4;
            const size_t smem4 = (((size_t)rows * (span + 8) + 15) & ~(size_t)15) + (size_t)rows * span * 2;
            if (smem4 <= 200 * 1024) {
                size_t &cur4 = h->squares_smem_attr[(const void *)k4];
                if (smem4 > 48 * 1024 && smem4 > cur4) {
                    CVB_CHECK_CUDA(cudaFuncSetAttribute(k4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
                    cur4 = smem4;
                }
                PROF(h, "k_squares");
                k4<<<grid, 256, smem4, h->stream>>>(a);
                LAUNCH_CHECK(h);
                return CVB_OK;
            }
        }
    }
    // the common op masks get a specialised instance (flag tests folded at compile time)
    void (*kern)(const SquareArgs) = k_squares<-1>;
    switch (p.ops) {
    case CVB_SQ_PD_STATS: kern = k_squares<CVB_SQ_PD_STATS>; break;
    case CVB_SQ_PD_SET_REF: kern = k_squares<CVB_SQ_PD_SET_REF>; break;
    case CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF: kern = k_squares<CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF>; break;
    case CVB_SQ_CD_CALIBRATE: kern = k_squares<CVB_SQ_CD_CALIBRATE>; break;
    case CVB_SQ_CD_DETECT: kern = k_squares<CVB_SQ_CD_DETECT>; break;
    case CVB_SQ_CD_UPDATE: kern = k_squares<CVB_SQ_CD_UPDATE>; break;
    case CVB_SQ_PD_STATS | CVB_SQ_CD_DETECT | CVB_SQ_CD_UPDATE:
        kern = k_squares<CVB_SQ_PD_STATS | CVB_SQ_CD_DETECT | CVB_SQ_CD_UPDATE>; break;
    case CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF | CVB_SQ_CD_CALIBRATE:
        kern = k_squares<CVB_SQ_PD_STATS | CVB_SQ_PD_SET_REF | CVB_SQ_CD_CALIBRATE>; break;
    default: break;
    }
    size_t &cur = h->squares_smem_attr[(const void *)kern];     // per handle = per device
    if (smem > 48 * 1024 && smem > cur) {
        CVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cur = smem;
    }
    PROF(h, "k_squares");
    kern<<<grid, 256, smem, h->stream>>>(a);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

int launch_state_reset(cvb_handle *h, cvb_state *s, int stream)
{
    const size_t plane = (size_t)s->BH * s->BW;
    uint8_t *p = s->flags + (stream < 0 ? 0 : (size_t)stream * plane);
    const size_t n = stream < 0 ? plane * s->n_streams : plane;
    CVB_CHECK_CUDA(cudaMemsetAsync(p, 0, n, h->stream));
    return CVB_OK;
}

CVB_BOUNDS_TU(grid)
