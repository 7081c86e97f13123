// Device-side helpers shared by the kernels of libcvb200 (sm_100a).
#pragma once
#include "cvb_internal.h"

#define CVB_DEV __device__ __forceinline__

// ---- index checks of the debug build (make debug -> libcvb200_dbg.so, -DCVB_DEBUG_BOUNDS) ---------------------------
// compute-sanitizer is closed on this pool, so the shared-memory tile / halo indices and the global offsets of the
// tiled kernels carry explicit range checks in a debug build: a violation bumps a per-translation-unit counter (and
// records its source line) instead of trapping, cvb_debug_bounds_violations() sums the counters, and
// tests/test_gpu_debug_bounds.py runs the odd-size whole-path script on that library and expects zero.
#ifdef CVB_DEBUG_BOUNDS
static __device__ unsigned long long cvb_bounds_fail_ctr;
static __device__ int cvb_bounds_first_line;
#define CVB_BOUNDS(cond)                                                                          \
    do {                                                                                          \
        if (!(cond)) {                                                                            \
            if (atomicAdd(&cvb_bounds_fail_ctr, 1ull) == 0ull) cvb_bounds_first_line = __LINE__;  \
        }                                                                                         \
    } while (0)
#define CVB_BOUNDS_TU(name)                                                                       \
    void cvb_bounds_##name(unsigned long long *n, int *line)                                      \
    {                                                                                             \
        *n = 0; *line = 0;                                                                        \
        cudaMemcpyFromSymbol(n, cvb_bounds_fail_ctr, sizeof *n);                                  \
        cudaMemcpyFromSymbol(line, cvb_bounds_first_line, sizeof *line);                          \
    }
#else
#define CVB_BOUNDS(cond) ((void)0)
#define CVB_BOUNDS_TU(name) void cvb_bounds_##name(unsigned long long *n, int *line) { *n = 0; *line = 0; }
#endif

// cv::borderInterpolate(BORDER_REFLECT_101)
CVB_DEV int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) p = p < 0 ? -p : 2 * n - 2 - p;
    return p;
}
// i / d for 0 <= i, i * d < 2^32, with inv = ceil(2^32 / d): one IMAD.HI instead of a division (inv == 0 means d == 1)
CVB_DEV int div_magic(int i, unsigned inv) { return inv ? (int)__umulhi((unsigned)i, inv) : i; }
CVB_DEV unsigned magic_of(int d) { return d > 1 ? 0xFFFFFFFFu / (unsigned)d + 1u : 0u; }
CVB_DEV int clamp_u8(int v) { return min(max(v, 0), 255); }
// saturate_cast<uchar>(float): round half to even, then clamp
CVB_DEV int round_u8(float v) { return clamp_u8(__float2int_rn(v)); }

CVB_DEV uint32_t pack_bgr(int b, int g, int r) { return (uint32_t)b | ((uint32_t)g << 8) | ((uint32_t)r << 16); }

// ---- shared-memory copy of the colour tables ----------------------------------
struct __align__(16) SmemColorTables {
    uint16_t gamma[256];
    uint16_t cbrt[2048];
    int32_t  lab2yf[512];      // [0..255]: y | ify << 16 (one load per pixel); [256..511] unused padding
    uint8_t  invgamma[4096];
};
static_assert(sizeof(SmemColorTables) % 16 == 0, "vector copy");
static_assert(offsetof(CvbTables, ltab) == sizeof(SmemColorTables), "CvbTables prefix must match");

CVB_DEV void load_color_tables(SmemColorTables *s, const CvbTables *g)
{
    const uint4 *src = reinterpret_cast<const uint4 *>(g);
    uint4 *dst = reinterpret_cast<uint4 *>(s);
    for (int i = threadIdx.x; i < (int)(sizeof(SmemColorTables) / 16); i += blockDim.x) dst[i] = __ldg(src + i);
}

// S1: RGB2Lab_b (OpenCV color_lab.cpp) -- reference call frame_enhancer.py:108
CVB_DEV void bgr2lab_px(const SmemColorTables *t, int b, int g, int r, int &L, int &A, int &Bc)
{
    const int Bl = t->gamma[b], Gl = t->gamma[g], Rl = t->gamma[r];
    const int fX = t->cbrt[(Rl * 1777 + Gl * 1541 + Bl * 778 + 2048) >> 12];
    const int fY = t->cbrt[(Rl * 871 + Gl * 2929 + Bl * 296 + 2048) >> 12];
    const int fZ = t->cbrt[(Rl * 73 + Gl * 448 + Bl * 3575 + 2048) >> 12];
    // OpenCV saturates these to 0..255; over all 2^24 inputs the values stay inside L 0..255, a 42..226,
    // b 20..223 (tests/test_abi.py::test_lab_forward_never_saturates), so the clamps are dropped
    L = (296 * fY - 1336934 + 16384) >> 15;
    A = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
    Bc = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
}

// abToXZ_b as a function (C integer division truncates toward zero)
CVB_DEV int ab_to_xz(int t)
{
    if (t <= 3390) return t * 108 / 841 - 290;
    return (((t * t) >> 14) * t) >> 14;   // t > 0 here, so >> is the C division
}

// S3: Lab2RGBinteger (8-bit) -- reference call frame_enhancer.py:120
CVB_DEV uint32_t lab2bgr_px(const SmemColorTables *t, int L, int a, int b)
{
    const uint32_t yf = (uint32_t)t->lab2yf[L];          // both values are < 2^15
    const int y = yf & 0xffff, ify = yf >> 16;
    const int adiv = ((5 * a * 53687 + 128) >> 13) - 4194;
    const int bdiv = ((b * 41943 + 16) >> 9) - 10485 + 1;
    const int x = ab_to_xz(ify + adiv), z = ab_to_xz(ify - bdiv);
    int ro = (12615 * x - 6296 * y - 2223 * z + 8192) >> 14;
    int go = (-3773 * x + 7684 * y + 185 * z + 8192) >> 14;
    int bo = (217 * x - 836 * y + 4715 * z + 8192) >> 14;
    ro = min(max(ro, 0), 4095); go = min(max(go, 0), 4095); bo = min(max(bo, 0), 4095);
    return pack_bgr(t->invgamma[bo], t->invgamma[go], t->invgamma[ro]);
}

// The CLAHE blend is a convex combination of four bytes with f32 weights a, 1 - a: it lies in [0, 255 * (1 + 2^-22)],
// so saturate_cast<uchar> reduces to the rounding
CVB_DEV int blend_u8(float v) { return __float2int_rn(v); }

// S2 interpolation (clahe.cpp CLAHE_Interpolation_Body); unfused f32, this association
struct ClaheAxis { int i1, i2; float a, a1; };
CVB_DEV ClaheAxis clahe_axis(int p, float inv_t, int tiles)
{
    ClaheAxis r;
    const float tf = __fsub_rn(__fmul_rn((float)p, inv_t), 0.5f);
    const int t1 = __float2int_rd(tf);
    r.a = __fsub_rn(tf, (float)t1);
    r.a1 = __fsub_rn(1.0f, r.a);
    r.i1 = max(t1, 0);
    r.i2 = min(t1 + 1, tiles - 1);
    return r;
}
CVB_DEV int clahe_interp(const uint8_t *__restrict__ lut, int tiles_x, const ClaheAxis &ax, const ClaheAxis &ay, int v)
{
    const float l11 = (float)__ldg(lut + (ay.i1 * tiles_x + ax.i1) * 256 + v);
    const float l12 = (float)__ldg(lut + (ay.i1 * tiles_x + ax.i2) * 256 + v);
    const float l21 = (float)__ldg(lut + (ay.i2 * tiles_x + ax.i1) * 256 + v);
    const float l22 = (float)__ldg(lut + (ay.i2 * tiles_x + ax.i2) * 256 + v);
    const float top = __fmul_rn(__fadd_rn(__fmul_rn(l11, ax.a1), __fmul_rn(l12, ax.a)), ay.a1);
    const float bot = __fmul_rn(__fadd_rn(__fmul_rn(l21, ax.a1), __fmul_rn(l22, ax.a)), ay.a);
    return blend_u8(__fadd_rn(top, bot));
}

// S7: RGB2Gray<uchar>, 15-bit coefficients -- frame_enhancer.py:154
CVB_DEV int gray_px(int b, int g, int r) { return (3735 * b + 19235 * g + 9798 * r + 16384) >> 15; }
// the same on a packed (b, g, r, x) word: two 16x8-bit dot products (IDP.2A), no byte extraction
CVB_DEV unsigned gray_packed(uint32_t bgrx)
{
    const unsigned bg = __dp2a_lo(3735u | (19235u << 16), bgrx, 16384u);
    return __dp2a_hi(9798u, bgrx, bg) >> 15;
}
// gray of the four pixels held in three consecutive words (b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3), packed
CVB_DEV uint32_t gray4_from_words(uint32_t w0, uint32_t w1, uint32_t w2)
{
    const unsigned g0 = gray_packed(w0);
    const unsigned g1 = gray_packed(__byte_perm(w0, w1, 0x6543));
    const unsigned g2 = gray_packed(__byte_perm(w1, w2, 0x5432));
    const unsigned g3 = gray_packed(w2 >> 8);
    return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
}

// S6: the 256-entry map of cv2.normalize(NORM_MINMAX,0,255) for one (min,max):
// f64 scale/shift, then f32 fma per value (convert_scale.simd.hpp, FMA3 hosts)
CVB_DEV int normalize_value(int v, int smin, int smax)
{
    const double range = (double)(smax - smin);
    const double scale = __dmul_rn(255.0, range > 2.220446049250313e-16 ? __ddiv_rn(1.0, range) : 0.0);
    const double shift = __dsub_rn(0.0, __dmul_rn((double)smin, scale));
    return round_u8(__fmaf_rn((float)v, (float)scale, (float)shift));
}

CVB_DEV int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CVB_DEV unsigned warp_sum_u(unsigned v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CVB_DEV unsigned long long warp_sum_ull(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
CVB_DEV int warp_min(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
CVB_DEV int warp_max(int v)
{
#pragma unroll
    for (int o = 16; o; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- TMA bulk copy (cp.async.bulk) of a 16-byte-aligned table into shared memory -------------
// One thread arms an mbarrier with the byte count and issues the copy; the copy engine moves the bytes and
// completes the barrier's transaction; readers wait on the barrier's phase.  No thread touches the data on the way.
CVB_DEV uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
CVB_DEV void mbar_init(uint64_t *bar, int arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
CVB_DEV void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
CVB_DEV void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
CVB_DEV void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(phase)
        : "memory");
}

// ---- byte-span staging between global and shared memory ---------------------------
// Copies nbytes from g into a 16B-aligned shared buffer so that g[i] lands at
// s_base[phase + i] with phase = (address of g) & 15: the 16-byte body then moves
// as aligned uint4 on both sides.  The buffer needs nbytes + 16 bytes.
CVB_DEV int span_phase(const void *g) { return (int)(reinterpret_cast<uintptr_t>(g) & 15); }

CVB_DEV void g2s_span(uint8_t *s_base, const uint8_t *__restrict__ g, int nbytes, int tid, int nthreads)
{
    const int phase = span_phase(g);
    const int head = min(nbytes, (16 - phase) & 15);
    const int body = (nbytes - head) >> 4;
    const int tail0 = head + (body << 4);
    for (int i = tid; i < head; i += nthreads) s_base[phase + i] = __ldg(g + i);
    const uint4 *gv = reinterpret_cast<const uint4 *>(g + head);
    uint4 *sv = reinterpret_cast<uint4 *>(s_base + phase + head);
    for (int i = tid; i < body; i += nthreads) sv[i] = __ldg(gv + i);
    for (int i = tail0 + tid; i < nbytes; i += nthreads) s_base[phase + i] = __ldg(g + i);
}
CVB_DEV void s2g_span(uint8_t *__restrict__ g, const uint8_t *s_base, int nbytes, int tid, int nthreads)
{
    const int phase = span_phase(g);
    const int head = min(nbytes, (16 - phase) & 15);
    const int body = (nbytes - head) >> 4;
    const int tail0 = head + (body << 4);
    for (int i = tid; i < head; i += nthreads) g[i] = s_base[phase + i];
    uint4 *gv = reinterpret_cast<uint4 *>(g + head);
    const uint4 *sv = reinterpret_cast<const uint4 *>(s_base + phase + head);
    for (int i = tid; i < body; i += nthreads) gv[i] = sv[i];
    for (int i = tail0 + tid; i < nbytes; i += nthreads) g[i] = s_base[phase + i];
}
