// frame_enhancer hot path as sm_100a kernels.
//
//   pass 1  k_tile_hist      BGR -> L (LUT) -> 8x8 tile histograms        (read 3N)
//           k_clahe_lut      clip / redistribute / scan -> 64 LUTs         (tiny)
//   pass 2  k_fused          LAB, CLAHE blend, LAB->BGR into a smem halo
//                            tile; bilateral d=9 from smem; 3x3 sharpen;
//                            global min/max                                (read 3N, write 3N)
//   pass 3  k_finish         normalize, gray, blur 5x5, Otsu histogram     (read 3N, write 5N)
//           k_otsu           f64 Otsu scan                                 (tiny)
//   pass 4  k_threshold      binary mask                                   (read N, write N)
//
// Reference: frame_enhancer.py:101-181 (ImageEnhancerPython) and its Cython
// twin src/cython/frame_enhancer_cython.pyx:86-153, which call cv2 for every
// stage; the arithmetic restated here is OpenCV 4.13's (file names per kernel).
// Compiled with -fmad=false: every fused multiply-add below is explicit.
#include "cvb_device.cuh"
#include <cfloat>
#include <cstdlib>


int cvb_clahe_geom(int H, int W, double clip_limit, int tx, int ty, ClaheGeom *g)
{
    CVB_REQUIRE(tx >= 1 && ty >= 1 && tx <= 16 && ty <= 16, "CLAHE tile grid %dx%d unsupported (1..16)", tx, ty);
    CVB_REQUIRE(H >= 1 && W >= 1, "empty image");
    g->tiles_x = tx; g->tiles_y = ty;
    g->ext_w = W; g->ext_h = H;
    if (W % tx != 0 || H % ty != 0) {  // clahe.cpp pads both axes, even one that divides
        g->ext_w = W + (tx - W % tx);
        g->ext_h = H + (ty - H % ty);
    }
    g->tile_w = g->ext_w / tx; g->tile_h = g->ext_h / ty;
    const int area = g->tile_w * g->tile_h;
    g->clip = 0;
    if (clip_limit > 0.0) {
        g->clip = (int)(clip_limit * area / 256);
        if (g->clip < 1) g->clip = 1;
    }
    g->lut_scale = (float)255 / (float)area;
    g->inv_tw = 1.0f / (float)g->tile_w;
    g->inv_th = 1.0f / (float)g->tile_h;
    return CVB_OK;
}

// ---------------------------------------------------------------------------------------
// S0: apply_color_profile (frame_enhancer.py:56-99), pointwise.
//   convertScaleAbs : |fma(v, alpha, beta)| rounded, saturated          (convert_scale.simd.hpp)
//   BGR2HSV (8-bit) : integer, sdiv / hdiv tables, 12 fractional bits   (color_hsv RGB2HSV_b)
//   NumPy step      : f32: radical-mode saturation, (h + shift) mod 180, scales, clip, truncate to u8
//   HSV2BGR (8-bit) : f32 sectors with 1 - s*h formed by fnma; the 32-pixel vector body of a row
//                     truncates, the row tail rounds to nearest even (measured against cv2 4.13)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_color_profile(const uint8_t *__restrict__ src, int H, int W, cvb_color_profile c,
                                                       uint8_t *__restrict__ dst)
{
    __shared__ int s_sdiv[256], s_hdiv[256];
    {
        const int i = threadIdx.x;
        // saturate_cast<int>((255 << 12) / (1. * i)), ((180 << 12) / (6. * i)): f64 quotient rounded to nearest even
        s_sdiv[i] = i ? __double2int_rn(__ddiv_rn((double)(255 << 12), (double)i)) : 0;
        s_hdiv[i] = i ? __double2int_rn(__ddiv_rn((double)(180 << 12), __dmul_rn(6.0, (double)i))) : 0;
    }
    __syncthreads();
    const float a = (float)c.contrast, b = (float)c.brightness;
    const int frame = blockIdx.z;
    const int nvec = c.simd_block > 0 ? (W / c.simd_block) * c.simd_block : 0;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= W || y >= H) return;
    const size_t o = ((size_t)frame * H * W + (size_t)y * W + x) * 3;
    int ch[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) ch[k] = round_u8(fabsf(__fmaf_rn((float)src[o + k], a, b)));
    // BGR -> HSV
    const int B = ch[0], G = ch[1], R = ch[2];
    const int v = max(B, max(G, R)), vmin = min(B, min(G, R)), diff = v - vmin;
    const int sat = (diff * s_sdiv[v] + (1 << 11)) >> 12;
    int hue = v == R ? G - B : (v == G ? B - R + 2 * diff : R - G + 4 * diff);
    hue = (hue * s_hdiv[diff] + (1 << 11)) >> 12;
    if (hue < 0) hue += 180;
    // NumPy f32 step
    float hf = (float)clamp_u8(hue), sf = (float)sat, vf = (float)v;
    if (c.radical_mode) {
        float hd = fabsf(__fsub_rn(hf, c.target_hue));
        hd = fminf(hd, __fsub_rn(180.f, hd));
        sf = hd < c.hue_window ? __fmul_rn(sf, 2.0f) : __fmul_rn(sf, 0.5f);
    }
    float m = fmodf(__fadd_rn(hf, c.hue_shift), 180.f);      // np.remainder: fmod, then shifted into [0, 180)
    if (m != 0.f) { if (m < 0.f) m = __fadd_rn(m, 180.f); } else m = 0.f;
    hf = fminf(fmaxf(m, 0.f), 179.f);
    sf = fminf(fmaxf(__fmul_rn(sf, c.sat_scale), 0.f), 255.f);
    vf = fminf(fmaxf(__fmul_rn(vf, c.val_scale), 0.f), 255.f);
    const int h8 = (int)hf, s8 = (int)sf, v8 = (int)vf;      // astype(np.uint8) truncates
    // HSV -> BGR
    const float s1 = __fmul_rn((float)s8, 1.f / 255.f), v1 = __fmul_rn((float)v8, 1.f / 255.f);
    float hh = __fmul_rn((float)h8, 6.f / 180.f);
    const float fl = floorf(hh);
    int sector = (int)fl % 6;
    hh = __fsub_rn(hh, fl);
    float tab[4];
    tab[0] = v1;
    tab[1] = __fmul_rn(v1, __fsub_rn(1.f, s1));
    tab[2] = __fmul_rn(v1, __fmaf_rn(-s1, hh, 1.f));
    tab[3] = __fmul_rn(v1, __fmaf_rn(-s1, __fsub_rn(1.f, hh), 1.f));
    // sector_data of HSV2RGB_native: {1,3,0},{1,0,2},{3,0,1},{0,2,1},{0,1,3},{2,1,0} packed two bits per entry
    const unsigned sel = sector == 0 ? 0x0Du : sector == 1 ? 0x21u : sector == 2 ? 0x13u : sector == 3 ? 0x18u
                       : sector == 4 ? 0x34u : 0x06u;       // bits: b | g << 2 | r << 4
    const bool vec = x < nvec;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float t = __fmul_rn(tab[(sel >> (2 * k)) & 3], 255.f);
        dst[o + k] = (uint8_t)clamp_u8(vec ? (int)t : __float2int_rn(t));
    }
}
int launch_color_profile(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_color_profile &p, uint8_t *out)
{
    dim3 grid((W + 63) / 64, (H + 3) / 4, n);
    PROF(h, "k_color_profile");
    k_color_profile<<<grid, 256, 0, h->stream>>>(bgr, H, W, p, out);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// ---------------------------------------------------------------------------------------
// Pointwise colour conversions (stage-isolated API; the hot path uses k_fused)
// ---------------------------------------------------------------------------------------
template <bool FWD>
__global__ void __launch_bounds__(256) k_lab_pointwise(const uint8_t *__restrict__ src, long npx,
                                                       const CvbTables *__restrict__ tabs, uint8_t *__restrict__ dst)
{
    __shared__ SmemColorTables st;
    load_color_tables(&st, tabs);
    __syncthreads();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (long)gridDim.x * blockDim.x) {
        const int c0 = src[3 * i], c1 = src[3 * i + 1], c2 = src[3 * i + 2];
        if (FWD) {
            int L, A, B;
            bgr2lab_px(&st, c0, c1, c2, L, A, B);
            dst[3 * i] = (uint8_t)L; dst[3 * i + 1] = (uint8_t)A; dst[3 * i + 2] = (uint8_t)B;
        } else {
            const uint32_t q = lab2bgr_px(&st, c0, c1, c2);
            dst[3 * i] = (uint8_t)q; dst[3 * i + 1] = (uint8_t)(q >> 8); dst[3 * i + 2] = (uint8_t)(q >> 16);
        }
    }
}
// blocks along x for a grid-stride streaming kernel: enough for `vecs` 16-byte
// vectors at 256 per block, capped so that the whole launch stays near `cap`
static int stream_blocks(long vecs, int cap)
{
    long b = (vecs + 255) / 256 + 1;
    if (cap < 1) cap = 1;
    if (b > cap) b = cap;
    return (int)b;
}
static int grid_for(cvb_handle *h, long items, int per_block)
{
    long b = (items + per_block - 1) / per_block;
    long cap = (long)h->sm_count * 16;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}
int launch_bgr2lab(cvb_handle *h, const uint8_t *bgr, long npx, uint8_t *lab)
{
    PROF(h, "k_lab_pointwise");
    k_lab_pointwise<true><<<grid_for(h, npx, 256), 256, 0, h->stream>>>(bgr, npx, h->d_tables, lab);
    LAUNCH_CHECK(h);
    return CVB_OK;
}
int launch_lab2bgr(cvb_handle *h, const uint8_t *lab, long npx, uint8_t *bgr)
{
    PROF(h, "k_lab_pointwise");
    k_lab_pointwise<false><<<grid_for(h, npx, 256), 256, 0, h->stream>>>(lab, npx, h->d_tables, bgr);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// ---------------------------------------------------------------------------------------
// Pass 1: per-tile histograms of L (CLAHE_CalcLut_Body, histogram part)
//   grid = (tiles, row-splits, frames); warp-private 256-bin histograms in smem.
//   Tiles are laid over the REFLECT_101-extended image (clahe.cpp), so right /
//   bottom tiles of non-divisible sizes read mirrored pixels.
//   Fast path (tile inside the image, 16-pixel groups 16-byte aligned): each thread
//   pulls 16 pixels as three 128-bit loads; L comes from two shared LUTs
//   (per-channel Y contributions, then Y-index -> L).
// ---------------------------------------------------------------------------------------
CVB_DEV int l_of_bgr(const int (*tY)[256], const uint8_t *ltab, uint32_t b, uint32_t g, uint32_t r)
{
    return ltab[(tY[0][b] + tY[1][g] + tY[2][r] + 2048) >> 12];
}

// FROM_BGR && lab_out != nullptr: the pass also stores the whole Lab pixel (RGB2Lab_b, the same integers stage A of
// k_fused would recompute for every tile + halo pixel), so that k_fused starts from Lab.
template <bool FROM_BGR>
__global__ void __launch_bounds__(256) k_tile_hist(const uint8_t *__restrict__ src, int H, int W, ClaheGeom g,
                                                   const CvbTables *__restrict__ tabs, int32_t *__restrict__ hist,
                                                   int32_t *__restrict__ minmax_init, uint8_t *__restrict__ lab_out)
{
    __shared__ int s_hist[8][256];
    __shared__ int s_tY[3][256];          // per-channel contribution to the Y index numerator
    __shared__ uint8_t s_ltab[2048];
    __shared__ uint16_t s_gam[FROM_BGR ? 256 : 1], s_cbrt[FROM_BGR ? 2048 : 1];   // for the Lab output
    const int tid = threadIdx.x, warp = tid >> 5;
    const int tile = blockIdx.x, tx = tile % g.tiles_x, ty = tile / g.tiles_x;
    const int frame = blockIdx.z;
    const bool want_lab = FROM_BGR && lab_out != nullptr;
    for (int i = tid; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
    if (FROM_BGR) {
        const int gv = tabs->gamma[tid];
        s_tY[0][tid] = gv * 296; s_tY[1][tid] = gv * 2929; s_tY[2][tid] = gv * 871;
        for (int i = tid; i < 2048 / 4; i += 256)
            reinterpret_cast<uint32_t *>(s_ltab)[i] = reinterpret_cast<const uint32_t *>(tabs->ltab)[i];
        if (want_lab) {
            s_gam[tid] = (uint16_t)gv;
            for (int i = tid; i < 2048 / 2; i += 256)
                reinterpret_cast<uint32_t *>(s_cbrt)[i] = reinterpret_cast<const uint32_t *>(tabs->cbrt)[i];
        }
    }
    if (minmax_init && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {
        minmax_init[2 * frame] = 255; minmax_init[2 * frame + 1] = 0;
    }
    __syncthreads();
    const int rows_per = (g.tile_h + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(r0 + rows_per, g.tile_h);
    constexpr int CH = FROM_BGR ? 3 : 1;
    const uint8_t *img = src + (size_t)frame * H * W * CH;
    uint8_t *lab = want_lab ? lab_out + (size_t)frame * H * W * 3 : nullptr;
    int *my = s_hist[warp];
    // one pixel: histogram of L; with want_lab the packed (L, a, b) word is returned as well
    auto lab_px = [&](uint32_t b, uint32_t gch, uint32_t r) -> uint32_t {
        const int Bl = s_gam[b], Gl = s_gam[gch], Rl = s_gam[r];
        const int fX = s_cbrt[(Rl * 1777 + Gl * 1541 + Bl * 778 + 2048) >> 12];
        const int fY = s_cbrt[(Rl * 871 + Gl * 2929 + Bl * 296 + 2048) >> 12];
        const int fZ = s_cbrt[(Rl * 73 + Gl * 448 + Bl * 3575 + 2048) >> 12];
        const int L = (296 * fY - 1336934 + 16384) >> 15;                    // never saturate (tests/test_abi.py)
        const int A = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
        const int Bc = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
        CVB_BOUNDS(L >= 0 && L < 256 && A >= 0 && A < 256 && Bc >= 0 && Bc < 256);
        atomicAdd(&my[L], 1);
        return pack_bgr(L, A, Bc);
    };
    const int x_lo = tx * g.tile_w, y_lo = ty * g.tile_h + r0;
    const bool inside = x_lo + g.tile_w <= W && ty * g.tile_h + r1 <= H;
    const bool fast = inside && (g.tile_w % 16 == 0) && (W % 16 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) &&
                      (!want_lab || (reinterpret_cast<uintptr_t>(lab_out) & 15) == 0);
    if (fast) {
        const int gpr = g.tile_w >> 4;                       // 16-pixel groups per tile row
        const int ngroups = (r1 - r0) * gpr;
        for (int i = tid; i < ngroups; i += 256) {
            const int ry = i / gpr, gx = i - ry * gpr;
            const size_t px = (size_t)(y_lo + ry) * W + x_lo + (gx << 4);
            if (FROM_BGR) {
                const uint4 *p = reinterpret_cast<const uint4 *>(img + px * 3);
                const uint4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
                const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
                if (want_lab) {
                    uint32_t o[12];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {       // 4 pixels per 3 words: b0 g0 r0 b1 | g1 r1 b2 g2 | r2 b3 g3 r3
                        const uint32_t w0 = w[3 * k], w1 = w[3 * k + 1], w2 = w[3 * k + 2];
                        const uint32_t q0 = lab_px(w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff);
                        const uint32_t q1 = lab_px(w0 >> 24, w1 & 0xff, (w1 >> 8) & 0xff);
                        const uint32_t q2 = lab_px((w1 >> 16) & 0xff, w1 >> 24, w2 & 0xff);
                        const uint32_t q3 = lab_px((w2 >> 8) & 0xff, (w2 >> 16) & 0xff, w2 >> 24);
                        o[3 * k] = q0 | (q1 << 24); o[3 * k + 1] = (q1 >> 8) | (q2 << 16); o[3 * k + 2] = (q2 >> 16) | (q3 << 8);
                    }
                    uint4 *op = reinterpret_cast<uint4 *>(lab + px * 3);
                    op[0] = make_uint4(o[0], o[1], o[2], o[3]); op[1] = make_uint4(o[4], o[5], o[6], o[7]);
                    op[2] = make_uint4(o[8], o[9], o[10], o[11]);
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t w0 = w[3 * k], w1 = w[3 * k + 1], w2 = w[3 * k + 2];
                        atomicAdd(&my[l_of_bgr(s_tY, s_ltab, w0 & 0xff, (w0 >> 8) & 0xff, (w0 >> 16) & 0xff)], 1);
                        atomicAdd(&my[l_of_bgr(s_tY, s_ltab, w0 >> 24, w1 & 0xff, (w1 >> 8) & 0xff)], 1);
                        atomicAdd(&my[l_of_bgr(s_tY, s_ltab, (w1 >> 16) & 0xff, w1 >> 24, w2 & 0xff)], 1);
                        atomicAdd(&my[l_of_bgr(s_tY, s_ltab, (w2 >> 8) & 0xff, (w2 >> 16) & 0xff, w2 >> 24)], 1);
                    }
                }
            } else {
                const uint4 a = __ldg(reinterpret_cast<const uint4 *>(img + px));
                const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    atomicAdd(&my[w[k] & 0xff], 1); atomicAdd(&my[(w[k] >> 8) & 0xff], 1);
                    atomicAdd(&my[(w[k] >> 16) & 0xff], 1); atomicAdd(&my[w[k] >> 24], 1);
                }
            }
        }
    } else {
        const int npx = (r1 - r0) * g.tile_w;
        for (int i = tid; i < npx; i += 256) {
            const int ry = i / g.tile_w, rx = i - ry * g.tile_w;
            const int sy = reflect101(y_lo + ry, H), sx = reflect101(x_lo + rx, W);
            if (FROM_BGR) {
                const uint8_t *p = img + ((size_t)sy * W + sx) * 3;
                if (want_lab) {
                    // mirrored positions of the padded tile grid store their source pixel again: same value
                    const uint32_t q = lab_px(p[0], p[1], p[2]);
                    uint8_t *o = lab + ((size_t)sy * W + sx) * 3;
                    o[0] = (uint8_t)q; o[1] = (uint8_t)(q >> 8); o[2] = (uint8_t)(q >> 16);
                } else {
                    atomicAdd(&my[l_of_bgr(s_tY, s_ltab, p[0], p[1], p[2])], 1);
                }
            } else {
                atomicAdd(&my[img[(size_t)sy * W + sx]], 1);
            }
        }
    }
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_hist[w][tid];
    if (tot) atomicAdd(&hist[((size_t)frame * g.tiles_x * g.tiles_y + tile) * 256 + tid], tot);
}

int launch_tile_hist(cvb_handle *h, const uint8_t *src, int from_bgr, int n, int H, int W, const ClaheGeom &g,
                     int32_t *hist, int32_t *minmax_init, uint8_t *lab_out)
{
    const int tiles = g.tiles_x * g.tiles_y;
    CVB_CHECK_CUDA(cudaMemsetAsync(hist, 0, sizeof(int32_t) * 256 * tiles * (size_t)n, h->stream));
    // enough blocks per frame to fill the machine at n == 1, fewer splits for big batches
    int splits = (4 * h->sm_count + tiles * n - 1) / (tiles * n);
    splits = max(1, min(splits, (g.tile_h + 7) / 8));
    dim3 grid(tiles, splits, n);
    PROF(h, "k_tile_hist");
    if (from_bgr) k_tile_hist<true><<<grid, 256, 0, h->stream>>>(src, H, W, g, h->d_tables, hist, minmax_init, lab_out);
    else k_tile_hist<false><<<grid, 256, 0, h->stream>>>(src, H, W, g, h->d_tables, hist, minmax_init, nullptr);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// clip, redistribute, prefix-sum, scale: one block per tile, one thread per bin
__global__ void __launch_bounds__(256) k_clahe_lut(const int32_t *__restrict__ hist, ClaheGeom g, uint8_t *__restrict__ lut)
{
    __shared__ int s_part[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t base = (size_t)blockIdx.x * 256;
    int v = hist[base + tid];
    if (g.clip > 0) {
        int over = max(v - g.clip, 0);
        v -= over;
        over = warp_sum(over);
        if (lane == 0) s_part[warp] = over;
        __syncthreads();
        int clipped = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) clipped += s_part[w];
        const int batch = clipped / 256;
        int resid = clipped - batch * 256;
        v += batch;
        if (resid != 0) {
            // for (i = 0; i < 256 && resid > 0; i += step, --resid) h[i]++
            const int step = max(256 / resid, 1);
            if (tid % step == 0 && tid / step < resid) v += 1;
        }
        __syncthreads();
    }
    // inclusive scan over 256 bins
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_part[warp] = x;
    __syncthreads();
    int add = 0;
    for (int w = 0; w < warp; ++w) add += s_part[w];
    x += add;
    lut[base + tid] = (uint8_t)round_u8(__fmul_rn((float)x, g.lut_scale));
}
int launch_clahe_lut(cvb_handle *h, const int32_t *hist, int n, const ClaheGeom &g, uint8_t *lut)
{
    PROF(h, "k_clahe_lut");
    k_clahe_lut<<<g.tiles_x * g.tiles_y * n, 256, 0, h->stream>>>(hist, g, lut);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// CLAHE interpolation on a plain u8 plane (stage-isolated API)
__global__ void __launch_bounds__(256) k_clahe_apply_plane(const uint8_t *__restrict__ src, int H, int W, ClaheGeom g,
                                                           const uint8_t *__restrict__ lut, uint8_t *__restrict__ dst)
{
    const int frame = blockIdx.z;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= W || y >= H) return;
    const size_t o = (size_t)frame * H * W + (size_t)y * W + x;
    const ClaheAxis ax = clahe_axis(x, g.inv_tw, g.tiles_x), ay = clahe_axis(y, g.inv_th, g.tiles_y);
    dst[o] = (uint8_t)clahe_interp(lut + (size_t)frame * g.tiles_x * g.tiles_y * 256, g.tiles_x, ax, ay, src[o]);
}
int launch_clahe_apply_plane(cvb_handle *h, const uint8_t *src, int n, int H, int W, const ClaheGeom &g,
                             const uint8_t *lut, uint8_t *dst)
{
    dim3 grid((W + 63) / 64, (H + 3) / 4, n);
    PROF(h, "k_clahe_apply_plane");
    k_clahe_apply_plane<<<grid, 256, 0, h->stream>>>(src, H, W, g, lut, dst);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// ---------------------------------------------------------------------------------------
// Pass 2: the fused tile kernel.
//   A  stage the (optionally lighting-corrected) pixels of the tile plus halo in
//      shared memory as packed BGRx words; out-of-image positions are filled by
//      REFLECT_101, i.e. exactly the copyMakeBorder image bilateralFilter sees.
//      Per-column / per-row work (mirror index, CLAHE cell and blend factors) is
//      computed once per block.
//   B  bilateral d=9 (bilateral_filter.dispatch.cpp / .simd.hpp): circular support
//      r<=4 (49 taps), weight = f32(space[k] * color[|db|+|dg|+|dr|]) read from a
//      shared-memory table that already holds the product for each of the 10
//      distinct radii, taps accumulated in row-major order with fmaf,
//      result = rint(sum * (1/wsum)).  Each thread owns runs of 4 adjacent pixels;
//      a row of the window is three LDS.128.
//   C  3x3 sharpen 10*c - sum9, saturate (filter2D, REFLECT_101 of the *filtered*
//      image: mirrored neighbours index already-computed B pixels) + min/max,
//      then 4 pixels -> 3 words stores.
// ---------------------------------------------------------------------------------------
struct FusedArgs {
    const uint8_t *src;
    uint8_t *dst;
    int H, W;
    const CvbTables *tabs;
    const uint8_t *lut;       // CLAHE LUTs of all frames (LIGHT)
    ClaheGeom g;
    const float *wlut;        // [10][768] space*colour weights, or [768] colour weights when !FOLD (device)
    float sw[81];             // spatial weights [dy+4][dx+4] (used when !FOLD)
    int32_t *minmax;          // per frame {min,max} or null
    int src_is_lab;           // LIGHT: src already holds (L, a, b) bytes (stored by the tile-histogram pass)
};

// class of a tap by squared radius: 0,1,2,4,5,8,9,10,13,16 -> 0..9
__host__ __device__ constexpr int r2_class(int r2)
{
    return r2 == 0 ? 0 : r2 == 1 ? 1 : r2 == 2 ? 2 : r2 == 4 ? 3 : r2 == 5 ? 4 : r2 == 8 ? 5 : r2 == 9 ? 6 : r2 == 10 ? 7
         : r2 == 13 ? 8 : 9;
}
static const int kR2[10] = {0, 1, 2, 4, 5, 8, 9, 10, 13, 16};

struct __align__(16) AxisInfo {  // one per staged column / row: a single 128-bit shared load
    float a;                     // CLAHE blend factor towards the second tile (a1 = 1 - a)
    uint32_t src;                // byte offset of the mirrored source column in a row (x * 3) / of the source row in the frame
    uint32_t t1, t2;             // byte offsets of the two CLAHE tile LUTs along this axis (column: tile * 256, row: tile * tiles_x * 256)
};

template <int TW, int TH, bool LIGHT, bool BIL, bool SHARP, int LUTMODE = 1>
struct FusedCfg {
    static constexpr int BX = SHARP ? 2 : 0, BY = SHARP ? 1 : 0;   // B halo around the tile
    static constexpr int BW = TW + 2 * BX, BH = TH + 2 * BY;
    static constexpr int AR = BIL ? 4 : 0;
    static constexpr int AW = BW + 2 * AR, AH = BH + 2 * AR;
    static constexpr int RUNS = BW / 4;
    static_assert(TW % 4 == 0 && BW % 4 == 0 && AW % 4 == 0, "runs of 4 / LDS.128 alignment");
    static constexpr size_t offA = 0;
    static constexpr size_t offB = offA + (size_t)AW * AH * 4;
    static constexpr size_t offW = offB + (BIL ? (size_t)BW * BH * 4 : 0);
    static constexpr int NLUT = LUTMODE == 1 ? 10 : 1;
    // the colour-conversion tables are only read in stage A, the B tile only written from stage B on: with a
    // bilateral stage they share memory
    static constexpr bool TAB_IN_B = BIL && LIGHT && (size_t)BW * BH * 4 >= sizeof(SmemColorTables);
    static constexpr size_t endW = offW + (BIL ? NLUT * 768 * 4 : 0);
    static constexpr size_t offT = TAB_IN_B ? offB : endW;
    static constexpr size_t offX = endW + (LIGHT && !TAB_IN_B ? sizeof(SmemColorTables) : 0);
    static constexpr size_t smem_bytes = offX + (size_t)(AW + AH) * sizeof(AxisInfo) + (size_t)(2 * TW + 2 * TH) * 2;
};

// u8 lane k of a packed word -> float, exact: one PRMT builds 2^23 + v, one FADD removes 2^23
CVB_DEV float byte_to_float(uint32_t q, int k)
{
    return __uint_as_float(__byte_perm(q, 0x4B000000u, 0x7540u | (unsigned)k)) - 8388608.0f;
}

#ifndef CONV_MIX
#define CONV_MIX 1
#endif
#ifndef CONV_PAIR
#define CONV_PAIR 1
#endif
// (Fetching the weights through the texture unit instead -- tex1Dfetch on the same table, no address arithmetic, a pipe
// the kernel does not use otherwise -- was measured and dropped: 123 / 65 / 54 us per frame with all / half / a third of
// the taps, the TEX pipe delivers about 4 texels per clock and SM: profiles/r02_notes.md.)
template <int TW, int TH, bool LIGHT, bool BIL, bool SHARP, int NT = 256, int LUTMODE = 1, int ROWS = 1>
__global__ void __launch_bounds__(NT) k_fused(const FusedArgs a)
{
    using Cfg = FusedCfg<TW, TH, LIGHT, BIL, SHARP, LUTMODE>;
    constexpr int AW = Cfg::AW, AH = Cfg::AH, BW = Cfg::BW, BH = Cfg::BH, AR = Cfg::AR;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *sA = reinterpret_cast<uint32_t *>(smem_raw + Cfg::offA);
    uint32_t *sB = BIL ? reinterpret_cast<uint32_t *>(smem_raw + Cfg::offB) : sA;
    const float *sW = reinterpret_cast<const float *>(smem_raw + Cfg::offW);
    SmemColorTables *sTab = reinterpret_cast<SmemColorTables *>(smem_raw + Cfg::offT);
    AxisInfo *sAx = reinterpret_cast<AxisInfo *>(smem_raw + Cfg::offX);          // [AW] columns then [AH] rows
    int16_t *sNb = reinterpret_cast<int16_t *>(sAx + AW + AH);                    // xm[TW] xp[TW] ym[TH] yp[TH]
    const int tid = threadIdx.x;
    const int frame = blockIdx.z;
    const int H = a.H, W = a.W;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const int bx0 = x0 - Cfg::BX, by0 = y0 - Cfg::BY, ax0 = bx0 - AR, ay0 = by0 - AR;
    const uint8_t *img = a.src + (size_t)frame * H * W * 3;

    // The two constant tables come in through the copy engine (TMA bulk copies): the colour-conversion tables are
    // awaited before stage A, the bilateral weight table before stage B, whose load so overlaps stage A.
    __shared__ uint64_t s_bar[2];
    if (tid == 0) {
        mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1);
        mbar_init_fence();
        if (LIGHT) bulk_g2s(sTab, a.tabs, (uint32_t)sizeof(SmemColorTables), &s_bar[0]);
        if (BIL) bulk_g2s(smem_raw + Cfg::offW, a.wlut, (uint32_t)(Cfg::NLUT * 768 * 4), &s_bar[1]);
    }
    for (int i = tid; i < AW + AH; i += NT) {
        AxisInfo ai;
        const bool col = i < AW;
        const int p = col ? reflect101(ax0 + i, W) : reflect101(ay0 + (i - AW), H);
        ai.a = 0.f;
        ai.src = col ? (uint32_t)p * 3u : (uint32_t)p * (uint32_t)W * 3u;       // a frame is < 4 GB
        ai.t1 = ai.t2 = 0;
        if (LIGHT) {
            const ClaheAxis ca = col ? clahe_axis(p, a.g.inv_tw, a.g.tiles_x) : clahe_axis(p, a.g.inv_th, a.g.tiles_y);
            const uint32_t unit = col ? 256u : 256u * (uint32_t)a.g.tiles_x;
            ai.a = ca.a;
            ai.t1 = (uint32_t)ca.i1 * unit; ai.t2 = (uint32_t)ca.i2 * unit;
        }
        sAx[i] = ai;
    }
    if (SHARP) {
        for (int i = tid; i < TW; i += NT) {
            const int X = min(x0 + i, W - 1);
            sNb[i] = (int16_t)(reflect101(X - 1, W) - bx0);
            sNb[TW + i] = (int16_t)(reflect101(X + 1, W) - bx0);
        }
        for (int i = tid; i < TH; i += NT) {
            const int Y = min(y0 + i, H - 1);
            sNb[2 * TW + i] = (int16_t)(reflect101(Y - 1, H) - by0);
            sNb[2 * TW + TH + i] = (int16_t)(reflect101(Y + 1, H) - by0);
        }
    }
    __syncthreads();
    if (LIGHT) mbar_wait(&s_bar[0], 0);

    // ---- A ----
    {
        const uint8_t *lut = LIGHT ? a.lut + (size_t)frame * a.g.tiles_x * a.g.tiles_y * 256 : nullptr;
        for (int i = tid; i < AW * AH; i += NT) {
            const int ly = i / AW, lx = i - ly * AW;
            const AxisInfo cx = sAx[lx], cy = sAx[AW + ly];
            const uint8_t *p = img + (cy.src + cx.src);                 // 32-bit sum: one 64-bit add per pixel
            const int c0 = __ldg(p), c1 = __ldg(p + 1), c2 = __ldg(p + 2);
            uint32_t q;
            if (LIGHT) {
                int L, A, B;
                if (a.src_is_lab) { L = c0; A = c1; B = c2; }
                else bgr2lab_px(sTab, c0, c1, c2, L, A, B);
                const uint32_t o1 = cy.t1 + (uint32_t)L, o2 = cy.t2 + (uint32_t)L;
                const float l11 = (float)__ldg(lut + (o1 + cx.t1)), l12 = (float)__ldg(lut + (o1 + cx.t2));
                const float l21 = (float)__ldg(lut + (o2 + cx.t1)), l22 = (float)__ldg(lut + (o2 + cx.t2));
                const float xa1 = __fsub_rn(1.0f, cx.a), ya1 = __fsub_rn(1.0f, cy.a);
                const float top = __fmul_rn(__fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, cx.a)), ya1);
                const float bot = __fmul_rn(__fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, cx.a)), cy.a);
                L = blend_u8(__fadd_rn(top, bot));
                q = lab2bgr_px(sTab, L, A, B);
            } else {
                q = pack_bgr(c0, c1, c2);
            }
            CVB_BOUNDS(i >= 0 && (size_t)i < (size_t)AW * AH && (size_t)(cy.src + cx.src) + 2 < (size_t)H * W * 3);
            sA[i] = q;
        }
    }
    __syncthreads();

    // ---- B ----
    if (BIL) mbar_wait(&s_bar[1], 0);
    // ROWS output rows per thread: a window row is loaded and converted to float once and feeds every output
    // row of the thread it is in range of (ROWS == 2 halves the byte->float conversions and the LDS.128s).
    if (BIL) {
        constexpr int RUNS = Cfg::RUNS;
        static_assert(BH % ROWS == 0, "B rows must split evenly over the rows of a thread");
        for (int item = tid; item < (BH / ROWS) * RUNS; item += NT) {
            const int rg = item / RUNS, r4 = (item - rg * RUNS) * 4;
            const int row0 = rg * ROWS;
            const int Y0 = by0 + row0, X = bx0 + r4;
            if (Y0 + ROWS - 1 < 0 || Y0 >= H || X + 3 < 0 || X >= W) continue;
            float wsum[ROWS][4], sb[ROWS][4], sg[ROWS][4], sr[ROWS][4];
            uint32_t ctr[ROWS][4];
#pragma unroll
            for (int t = 0; t < ROWS; ++t) {
                CVB_BOUNDS((row0 + t + 4) * AW + r4 + 4 + 3 < AW * AH && row0 >= 0 && r4 >= 0);
                const uint4 c = *reinterpret_cast<const uint4 *>(sA + (row0 + t + 4) * AW + r4 + 4);
                ctr[t][0] = c.x; ctr[t][1] = c.y; ctr[t][2] = c.z; ctr[t][3] = c.w;
#pragma unroll
                for (int j = 0; j < 4; ++j) wsum[t][j] = sb[t][j] = sg[t][j] = sr[t][j] = 0.f;
            }
#pragma unroll
            for (int k = 0; k < 9 + ROWS - 1; ++k) {          // window row k is A row (row0 + k): dy = k - 4 - t for output row t
                const uint32_t *rowp = sA + (row0 + k) * AW + r4;
                uint32_t px[12];
                float fb[12], fg[12], fr[12];
                // columns r4 .. r4+11 of the A tile hold image x = X-4 .. X+7; first needed column over the rows fed:
                // 4 - (largest |dx| of the disc on that dy)
                int lo = 4;
#pragma unroll
                for (int t = 0; t < ROWS; ++t) {
                    const int dy = k - 4 - t, ady = dy < 0 ? -dy : dy;
                    if (ady <= 4) lo = min(lo, ady == 4 ? 4 : ady == 3 ? 2 : ady >= 1 ? 1 : 0);
                }
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    if (v != 1 && lo >= 4) continue;                 // |dy| == 4 needs only the middle quad
                    CVB_BOUNDS((row0 + k) * AW + r4 + 4 * v + 3 < AW * AH);
                    const uint4 q = *reinterpret_cast<const uint4 *>(rowp + 4 * v);
                    px[4 * v] = q.x; px[4 * v + 1] = q.y; px[4 * v + 2] = q.z; px[4 * v + 3] = q.w;
                }
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    if (c < lo || c > 11 - lo) continue;
#if CONV_PAIR
                    {   // two byte->float conversions share one packed add (FADD2): 2 PRMT + 1 FADD2 instead of 2 PRMT + 2 FADD
                        unsigned long long pr, res;
                        const uint32_t mb = __byte_perm(px[c], 0x4B000000u, 0x7540u), mg = __byte_perm(px[c], 0x4B000000u, 0x7541u);
                        asm("mov.b64 %0, {%1, %2};" : "=l"(pr) : "r"(mb), "r"(mg));
                        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(res) : "l"(pr), "l"(0xCB000000CB000000ull));
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(fb[c]), "=f"(fg[c]) : "l"(res));
                    }
#else
                    fb[c] = byte_to_float(px[c], 0);
                    fg[c] = byte_to_float(px[c], 1);
#endif
                    // one channel goes through the otherwise idle conversion unit (I2F.U8): one issue slot instead of two
                    fr[c] = CONV_MIX ? (float)((px[c] >> 16) & 0xffu) : byte_to_float(px[c], 2);
                }
#pragma unroll
                for (int t = 0; t < ROWS; ++t) {
                    const int dy = k - 4 - t;
                    if (dy < -4 || dy > 4) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
#pragma unroll
                        for (int dx = -4; dx <= 4; ++dx) {
                            if (dy * dy + dx * dx > 16) continue;
                            const int c = j + 4 + dx;
                            float w;
                            if (dy == 0 && dx == 0) {
                                w = 1.0f;      // the centre tap: distance 0 in space and colour, exp(0) * exp(0)
                            } else {
                                const unsigned sad = __vsadu4(px[c], ctr[t][j]);
                                CVB_BOUNDS(sad < 768u);
                                w = LUTMODE == 1 ? sW[r2_class(dy * dy + dx * dx) * 768 + sad]
                                                 : __fmul_rn(a.sw[(dy + 4) * 9 + dx + 4], sW[sad]);
                            }
                            wsum[t][j] = __fadd_rn(wsum[t][j], w);
                            sb[t][j] = __fmaf_rn(fb[c], w, sb[t][j]);
                            sg[t][j] = __fmaf_rn(fg[c], w, sg[t][j]);
                            sr[t][j] = __fmaf_rn(fr[c], w, sr[t][j]);
                        }
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < ROWS; ++t) {
                uint4 o;
                uint32_t *op = &o.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float inv = __fdiv_rn(1.0f, wsum[t][j]);
                    op[j] = pack_bgr(round_u8(__fmul_rn(sb[t][j], inv)), round_u8(__fmul_rn(sg[t][j], inv)),
                                     round_u8(__fmul_rn(sr[t][j], inv)));
                }
                CVB_BOUNDS((row0 + t) * BW + r4 + 3 < BW * BH);
                *reinterpret_cast<uint4 *>(sB + (row0 + t) * BW + r4) = o;
            }
        }
        __syncthreads();
    }

    // ---- C ----
    int vmin = 255, vmax = 0;
    // Sharpen, fast path (aligned rows, tile not cut on the right): the one-pixel ring of the B tile that lies outside
    // the image is first filled with its REFLECT_101 mirror, then a thread sharpens 4 pixels of a row from two LDS.128
    // per neighbour row (column sums shared between the 4 outputs) and stores its 12 bytes straight to the frame.
    const bool fastC = SHARP && BIL && (W % 4 == 0) && (x0 + TW <= W) && W >= 2 && H >= 2 &&
                       ((reinterpret_cast<uintptr_t>(a.dst) & 3) == 0);
    if (fastC) {
        constexpr int BX = Cfg::BX, BY = Cfg::BY;
        const int vh = min(TH, H - y0);                        // valid output rows of this tile
        const bool top = y0 == 0, bottom = y0 + vh == H, left = x0 == 0, right = x0 + TW == W;
        if (top || bottom) {
            for (int i = tid; i < BW; i += NT) {
                if (top) sB[(BY - 1) * BW + i] = sB[(BY + 1) * BW + i];
                if (bottom) sB[(BY + vh) * BW + i] = sB[(BY + vh - 2) * BW + i];
            }
            __syncthreads();
        }
        if (left || right) {
            for (int i = tid; i < vh + 2; i += NT) {
                uint32_t *row = sB + (BY - 1 + i) * BW;
                if (left) row[BX - 1] = row[BX + 1];
                if (right) row[BX + TW] = row[BX + TW - 2];
            }
            __syncthreads();
        }
        uint8_t *out = a.dst + (size_t)frame * H * W * 3;
        constexpr int GROUPS = TW / 4;
        for (int item = tid; item < TH * GROUPS; item += NT) {
            const int ty = item / GROUPS, tx4 = (item - ty * GROUPS) * 4;
            if (ty >= vh) break;
            // columns tx4 .. tx4+7 of the B tile = image x - 2 .. x + 5 for the group's first pixel x (BX == 2)
            uint32_t cbr[6], cg[6], ctr_w[4];
#pragma unroll
            for (int k = 0; k < 6; ++k) cbr[k] = cg[k] = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const uint32_t *rp = sB + (ty + BY - 1 + r) * BW + tx4;
                CVB_BOUNDS((ty + BY - 1 + r) >= 0 && (ty + BY - 1 + r) * BW + tx4 + 7 < BW * BH);
                const uint4 q0 = *reinterpret_cast<const uint4 *>(rp), q1 = *reinterpret_cast<const uint4 *>(rp + 4);
                const uint32_t w[6] = {q0.y, q0.z, q0.w, q1.x, q1.y, q1.z};      // image x - 1 .. x + 4
#pragma unroll
                for (int k = 0; k < 6; ++k) { cbr[k] += w[k] & 0x00ff00ffu; cg[k] += (w[k] >> 8) & 0xffu; }
                if (r == 1) { ctr_w[0] = w[1]; ctr_w[1] = w[2]; ctr_w[2] = w[3]; ctr_w[3] = w[4]; }
            }
            CVB_BOUNDS(y0 + ty < H && x0 + tx4 + 3 < W);
            uint32_t res[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                // 16-bit lanes: (b, r) in one word, g in the other; nine bytes sum to <= 2295
                const uint32_t s02 = cbr[j] + cbr[j + 1] + cbr[j + 2], s1 = cg[j] + cg[j + 1] + cg[j + 2];
                const uint32_t c = ctr_w[j];
                const int b = clamp_u8(10 * (int)(c & 0xff) - (int)(s02 & 0xffff));
                const int g = clamp_u8(10 * (int)((c >> 8) & 0xff) - (int)s1);
                const int r = clamp_u8(10 * (int)((c >> 16) & 0xff) - (int)(s02 >> 16));
                res[j] = pack_bgr(b, g, r);
                if (a.minmax) {
                    vmin = min(vmin, min(b, min(g, r)));
                    vmax = max(vmax, max(b, max(g, r)));
                }
            }
            uint32_t *o32 = reinterpret_cast<uint32_t *>(out + ((size_t)(y0 + ty) * W + x0 + tx4) * 3);
            o32[0] = (res[0] & 0xffffffu) | (res[1] << 24);
            o32[1] = ((res[1] >> 8) & 0xffffu) | (res[2] << 16);
            o32[2] = ((res[2] >> 16) & 0xffu) | (res[3] << 8);
        }
        if (a.minmax) {
            vmin = warp_min(vmin); vmax = warp_max(vmax);
            if ((tid & 31) == 0) {
                atomicMin(a.minmax + 2 * frame, vmin);
                atomicMax(a.minmax + 2 * frame + 1, vmax);
            }
        }
        return;
    }
    // general path (results of the tile go to sOut = the start of sA, free once B is done)
    uint32_t *sOut = sA;
    uint32_t keep[(TW * TH + NT - 1) / NT];
#pragma unroll
    for (int it = 0; it < (TW * TH + NT - 1) / NT; ++it) {
        const int i = tid + it * NT;
        uint32_t q = 0;
        if (i < TW * TH) {
            const int ty = i / TW, tx = i - ty * TW;
            if (y0 + ty < H && x0 + tx < W) {
                if (SHARP) {
                    const int xm = sNb[tx], xp = sNb[TW + tx], xc = tx + Cfg::BX;
                    const int ym = sNb[2 * TW + ty], yp = sNb[2 * TW + TH + ty], yc = ty + Cfg::BY;
                    const uint32_t *r0 = sB + ym * BW, *r1 = sB + yc * BW, *r2 = sB + yp * BW;
                    const uint32_t c = r1[xc];
                    const uint32_t t0 = r0[xm], t1 = r0[xc], t2 = r0[xp], t3 = r1[xm], t5 = r1[xp], t6 = r2[xm], t7 = r2[xc],
                                   t8 = r2[xp];
                    // 16-bit lanes: (b, r) in one word, g in the other; nine bytes sum to <= 2295
                    const uint32_t s02 = (t0 & 0x00ff00ffu) + (t1 & 0x00ff00ffu) + (t2 & 0x00ff00ffu) + (t3 & 0x00ff00ffu) +
                                         (c & 0x00ff00ffu) + (t5 & 0x00ff00ffu) + (t6 & 0x00ff00ffu) + (t7 & 0x00ff00ffu) +
                                         (t8 & 0x00ff00ffu);
                    const uint32_t s1 = ((t0 >> 8) & 0xffu) + ((t1 >> 8) & 0xffu) + ((t2 >> 8) & 0xffu) + ((t3 >> 8) & 0xffu) +
                                        ((c >> 8) & 0xffu) + ((t5 >> 8) & 0xffu) + ((t6 >> 8) & 0xffu) + ((t7 >> 8) & 0xffu) +
                                        ((t8 >> 8) & 0xffu);
                    const int b = clamp_u8(10 * (int)(c & 0xff) - (int)(s02 & 0xffff));
                    const int g = clamp_u8(10 * (int)((c >> 8) & 0xff) - (int)s1);
                    const int r = clamp_u8(10 * (int)((c >> 16) & 0xff) - (int)(s02 >> 16));
                    q = pack_bgr(b, g, r);
                } else {
                    q = sB[(ty + Cfg::BY) * BW + tx + Cfg::BX];
                }
                if (a.minmax) {
                    const int b = q & 0xff, g = (q >> 8) & 0xff, r = (q >> 16) & 0xff;
                    vmin = min(vmin, min(b, min(g, r)));
                    vmax = max(vmax, max(b, max(g, r)));
                }
            }
        }
        keep[it] = q;
    }
    __syncthreads();                      // every read of sB (may alias sA) is done
#pragma unroll
    for (int it = 0; it < (TW * TH + NT - 1) / NT; ++it) {
        const int i = tid + it * NT;
        if (i < TW * TH) sOut[i] = keep[it];
    }
    __syncthreads();
    uint8_t *out = a.dst + (size_t)frame * H * W * 3;
    const bool fast = (W % 4 == 0) && (x0 + TW <= W) && ((reinterpret_cast<uintptr_t>(a.dst) & 3) == 0);
    constexpr int GROUPS = TW / 4;
    for (int item = tid; item < TH * GROUPS; item += NT) {
        const int ty = item / GROUPS, tx4 = (item - ty * GROUPS) * 4;
        const int Y = y0 + ty;
        if (Y >= H) break;
        const uint4 rq = *reinterpret_cast<const uint4 *>(sOut + ty * TW + tx4);
        const uint32_t res[4] = {rq.x, rq.y, rq.z, rq.w};
        uint8_t *o = out + ((size_t)Y * W + x0 + tx4) * 3;
        if (fast) {
            // 4 pixels = 12 bytes = 3 aligned words
            uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
            o32[0] = (res[0] & 0xffffffu) | (res[1] << 24);
            o32[1] = ((res[1] >> 8) & 0xffffu) | (res[2] << 16);
            o32[2] = ((res[2] >> 16) & 0xffu) | (res[3] << 8);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (x0 + tx4 + j < W) {
                    o[3 * j] = (uint8_t)res[j]; o[3 * j + 1] = (uint8_t)(res[j] >> 8); o[3 * j + 2] = (uint8_t)(res[j] >> 16);
                }
        }
    }
    if (a.minmax) {
        vmin = warp_min(vmin); vmax = warp_max(vmax);
        if ((tid & 31) == 0) {
            atomicMin(a.minmax + 2 * frame, vmin);
            atomicMax(a.minmax + 2 * frame + 1, vmax);
        }
    }
}

template <int TW, int TH, bool LIGHT, bool BIL, bool SHARP, int NT = 256, int LUTMODE = 1, int ROWS = 1>
static int launch_fused_t(cvb_handle *h, const FusedArgs &a, int n)
{
    using Cfg = FusedCfg<TW, TH, LIGHT, BIL, SHARP, LUTMODE>;
    auto kern = k_fused<TW, TH, LIGHT, BIL, SHARP, NT, LUTMODE, ROWS>;
    // per device: one handle per GPU, possibly several GPUs in one process
    if (!h->fused_attr_done.count((const void *)kern)) {
        CVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes));
        h->fused_attr_done.insert((const void *)kern);
    }
    dim3 grid((a.W + TW - 1) / TW, (a.H + TH - 1) / TH, n);
    PROF(h, "k_fused");
    kern<<<grid, NT, Cfg::smem_bytes, h->stream>>>(a);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

int launch_fused(cvb_handle *h, const uint8_t *src, int n, int H, int W, bool light, bool bilateral, bool sharpen,
                 const ClaheGeom *g, const uint8_t *lut, double sigma_color, double sigma_space, uint8_t *out,
                 int32_t *minmax, bool src_is_lab)
{
    FusedArgs a;
    a.src_is_lab = src_is_lab ? 1 : 0;
    a.src = src; a.dst = out; a.H = H; a.W = W; a.tabs = h->d_tables; a.lut = lut; a.minmax = minmax;
    a.wlut = nullptr;
    if (g) a.g = *g; else memset(&a.g, 0, sizeof a.g);
    if (bilateral) {
        if (h->color_sigma != sigma_color || h->space_sigma != sigma_space || !h->d_color) {
            std::vector<float> color(768), space(81), wl(11 * 768);
            cvb_host_bilateral_tables(sigma_color, sigma_space, color.data(), space.data());
            for (int c = 0; c < 10; ++c) {
                // any tap of that radius (weights depend on dy^2+dx^2 only)
                float sw = 0.f;
                for (int dy = -4; dy <= 4; ++dy)
                    for (int dx = -4; dx <= 4; ++dx)
                        if (dy * dy + dx * dx == kR2[c]) sw = space[(dy + 4) * 9 + dx + 4];
                for (int i = 0; i < 768; ++i) {
                    volatile float prod = sw * color[i];      // one f32 rounding, as space_weight[k]*color_weight[i]
                    wl[c * 768 + i] = prod;
                }
            }
            for (int i = 0; i < 768; ++i) wl[10 * 768 + i] = color[i];
            if (!h->d_color) CVB_CHECK_CUDA(cudaMalloc(&h->d_color, wl.size() * sizeof(float)));
            CVB_CHECK_CUDA(cudaMemcpyAsync(h->d_color, wl.data(), wl.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));   // pageable source dies at scope end
            h->color_sigma = sigma_color; h->space_sigma = sigma_space;
        }
        a.wlut = h->d_color;
    }
    memset(a.sw, 0, sizeof a.sw);
    if (bilateral) cvb_host_bilateral_tables(sigma_color, sigma_space, nullptr, a.sw);
    // Tile shapes measured on B200 at 1080p (profiles/r01_notes.md): 120x64 outputs per 512-thread CTA.  With the
    // sharpen halo the bilateral stage then has 31 runs x 33 row pairs = 1023 work items for its two rounds of 512
    // threads, the sharpen stage exactly 15 pixels per thread; two CTAs fit an SM because the colour tables share
    // memory with the B tile.  The folded 30 KB weight table beat the 3 KB one.
    if (light && bilateral && sharpen) {
        // CVB_FUSED=<variant> routes the process_pipeline path through the persistent TMA-fed kernel of cvb_fused2.cu
        // (needs Lab rows with a 16-byte pitch); unset = this file's per-tile kernel, which is still the faster one
        // (profiles/r02_notes.md)
        static const char *env = getenv("CVB_FUSED");
        if (env && env[0] >= '0' && env[0] <= '9' && src_is_lab && minmax && fused_tma_applicable(H, W, src, out))
            return launch_fused_tma(h, src, n, H, W, *g, lut, h->d_color, a.sw, out, minmax, atoi(env));
        return launch_fused_t<120, 64, true, true, true, 512, 1, 2>(h, a, n);
    }
    if (!light && bilateral && !sharpen) return launch_fused_t<120, 64, false, true, false, 512, 1, 2>(h, a, n);
    if (!light && bilateral && sharpen) return launch_fused_t<120, 64, false, true, true, 512, 1, 2>(h, a, n);
    if (light && bilateral && !sharpen) return launch_fused_t<120, 64, true, true, false, 512, 1, 2>(h, a, n);
    if (!light && !bilateral && sharpen) return launch_fused_t<60, 30, false, false, true>(h, a, n);
    if (light && !bilateral && !sharpen) return launch_fused_t<60, 30, true, false, false>(h, a, n);
    cvb_set_error("unsupported fused stage combination");
    return CVB_ERR_INVALID;
}
int launch_correct_lighting(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const ClaheGeom &g,
                            const uint8_t *lut, uint8_t *out)
{
    return launch_fused(h, bgr, n, H, W, true, false, false, &g, lut, 0, 0, out, nullptr, false);
}

// ---------------------------------------------------------------------------------------
// S6 stand-alone: global min/max over all channels, then the 256-entry map
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_minmax_init(int32_t *minmax, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { minmax[2 * i] = 255; minmax[2 * i + 1] = 0; }
}
__global__ void __launch_bounds__(256) k_minmax(const uint8_t *__restrict__ src, long bytes, int32_t *__restrict__ minmax)
{
    const int frame = blockIdx.y;
    const uint8_t *p = src + (size_t)frame * bytes;
    int lo = 255, hi = 0;
    const long nvec = ((reinterpret_cast<uintptr_t>(p) & 15) == 0) ? bytes / 16 : 0;
    const uint4 *pv = reinterpret_cast<const uint4 *>(p);
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
        const uint4 q = __ldg(pv + i);
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // bytewise min/max of the word against the running packed extrema
            lo = min(lo, (int)min(min(w[k] & 0xff, (w[k] >> 8) & 0xff), min((w[k] >> 16) & 0xff, w[k] >> 24)));
            hi = max(hi, (int)max(max(w[k] & 0xff, (w[k] >> 8) & 0xff), max((w[k] >> 16) & 0xff, w[k] >> 24)));
        }
    }
    for (long i = nvec * 16 + (long)blockIdx.x * 256 + threadIdx.x; i < bytes; i += (long)gridDim.x * 256) {
        const int v = p[i];
        lo = min(lo, v); hi = max(hi, v);
    }
    lo = warp_min(lo); hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(minmax + 2 * frame, lo);
        atomicMax(minmax + 2 * frame + 1, hi);
    }
}
int launch_minmax(cvb_handle *h, const uint8_t *src, int n, long bytes_per_frame, int32_t *minmax)
{
    PROF(h, "k_minmax_init");
    k_minmax_init<<<(n + 255) / 256, 256, 0, h->stream>>>(minmax, n);
    LAUNCH_CHECK(h);
    int bx = stream_blocks(bytes_per_frame / 16, 8 * h->sm_count / (n > 0 ? n : 1));
    dim3 grid(bx, n);
    PROF(h, "k_minmax");
    k_minmax<<<grid, 256, 0, h->stream>>>(src, bytes_per_frame, minmax);
    LAUNCH_CHECK(h);
    return CVB_OK;
}
__global__ void __launch_bounds__(256) k_normalize(const uint8_t *__restrict__ src, long bytes,
                                                   const int32_t *__restrict__ minmax, uint8_t *__restrict__ dst)
{
    __shared__ uint8_t s_map[256];
    const int frame = blockIdx.y;
    s_map[threadIdx.x] = (uint8_t)normalize_value(threadIdx.x, minmax[2 * frame], minmax[2 * frame + 1]);
    __syncthreads();
    const uint8_t *p = src + (size_t)frame * bytes;
    uint8_t *o = dst + (size_t)frame * bytes;
    const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(o)) & 15) == 0;
    const long nvec = al ? bytes / 16 : 0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
        uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + i);
        uint32_t *w = &q.x;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            w[k] = (uint32_t)s_map[w[k] & 0xff] | ((uint32_t)s_map[(w[k] >> 8) & 0xff] << 8) |
                   ((uint32_t)s_map[(w[k] >> 16) & 0xff] << 16) | ((uint32_t)s_map[w[k] >> 24] << 24);
        reinterpret_cast<uint4 *>(o)[i] = q;
    }
    for (long i = nvec * 16 + (long)blockIdx.x * 256 + threadIdx.x; i < bytes; i += (long)gridDim.x * 256)
        o[i] = s_map[p[i]];
}
int launch_normalize(cvb_handle *h, const uint8_t *src, int n, long bytes_per_frame, const int32_t *minmax, uint8_t *out)
{
    int bx = stream_blocks(bytes_per_frame / 16, 16 * h->sm_count / (n > 0 ? n : 1));
    dim3 grid(bx, n);
    PROF(h, "k_normalize");
    k_normalize<<<grid, 256, 0, h->stream>>>(src, bytes_per_frame, minmax, out);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// S7 stand-alone
__global__ void __launch_bounds__(256) k_gray(const uint8_t *__restrict__ bgr, long npx, uint8_t *__restrict__ gray)
{
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < npx; i += (long)gridDim.x * 256)
        gray[i] = (uint8_t)gray_px(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
}
int launch_gray(cvb_handle *h, const uint8_t *bgr, long npx, uint8_t *gray)
{
    PROF(h, "k_gray");
    k_gray<<<grid_for(h, npx, 256), 256, 0, h->stream>>>(bgr, npx, gray);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// S8 stand-alone, any odd k <= 31: Q8 x Q8 fixed point (smooth.dispatch.cpp /
// fixedpoint.inl.hpp), REFLECT_101.  64x16 tile + halo in shared memory.
struct GaussK { int q[31]; int k; };
__global__ void __launch_bounds__(256) k_gaussian(const uint8_t *__restrict__ src, int H, int W, GaussK gk,
                                                  uint8_t *__restrict__ dst)
{
    constexpr int TWG = 64, THG = 16, RMAX = 15;
    __shared__ uint8_t s_in[(THG + 2 * RMAX) * (TWG + 2 * RMAX)];
    __shared__ uint16_t s_h[(THG + 2 * RMAX) * TWG];
    const int r = gk.k / 2, aw = TWG + 2 * r, ah = THG + 2 * r;
    const int frame = blockIdx.z, x0 = blockIdx.x * TWG, y0 = blockIdx.y * THG, tid = threadIdx.x;
    const uint8_t *img = src + (size_t)frame * H * W;
    for (int i = tid; i < aw * ah; i += 256) {
        const int ly = i / aw, lx = i - ly * aw;
        s_in[i] = img[(size_t)reflect101(y0 - r + ly, H) * W + reflect101(x0 - r + lx, W)];
    }
    __syncthreads();
    for (int i = tid; i < ah * TWG; i += 256) {
        const int ly = i / TWG, lx = i - ly * TWG;
        uint32_t s = 0;
        for (int j = 0; j < gk.k; ++j) s += (uint32_t)gk.q[j] * s_in[ly * aw + lx + j];
        s_h[i] = (uint16_t)s;
    }
    __syncthreads();
    for (int i = tid; i < THG * TWG; i += 256) {
        const int ly = i / TWG, lx = i - ly * TWG;
        const int X = x0 + lx, Y = y0 + ly;
        if (X >= W || Y >= H) continue;
        uint32_t s = 0;
        for (int j = 0; j < gk.k; ++j) s += (uint32_t)gk.q[j] * s_h[(ly + j) * TWG + lx];
        dst[(size_t)frame * H * W + (size_t)Y * W + X] = (uint8_t)((s + 32768u) >> 16);
    }
}
int launch_gaussian(cvb_handle *h, const uint8_t *src, int n, int H, int W, int ksize, double sigma, uint8_t *dst)
{
    GaussK gk;
    memset(&gk, 0, sizeof gk);
    if (cvb_host_gaussian_q8_sigma(ksize, sigma, gk.q) != CVB_OK) {
        cvb_set_error("GaussianBlur ksize %d unsupported (odd, 1..31)", ksize);
        return CVB_ERR_INVALID;
    }
    gk.k = ksize;
    dim3 grid((W + 63) / 64, (H + 15) / 16, n);
    PROF(h, "k_gaussian");
    k_gaussian<<<grid, 256, 0, h->stream>>>(src, H, W, gk, dst);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// ---------------------------------------------------------------------------------------
// Pass 3: normalize -> gray -> blur5 -> histogram.  Tile 120x36 (tiles 1080p and 4K exactly), halo 2.
//   Thread mapping: a warp owns a row, a lane owns a group of 4 pixels (120 + 2*4 columns = 32 groups),
//   so row pointers are warp-uniform and a warp moves 384 contiguous bytes per row.
//   1  groups of 4 pixels (three 32-bit words) straight from the sharpened frame: min-max map
//      (skipped when it is the identity), write the enhanced pixels (3 words) and gray (1 word),
//      keep gray (+ halo) in shared memory;
//   2  horizontal [1 4 6 4 1];  3  vertical [1 4 6 4 1], (sum+128)>>8, histogram.
// Mirror rows / columns (REFLECT_101 of GaussianBlur) are resolved when gray is staged.
// FAST: W % 4 == 0 and every pointer 4-byte aligned, so a group is three aligned words and never
// straddles the right image edge; the generic instance keeps the byte paths.
// ---------------------------------------------------------------------------------------
template <bool NORM, bool FAST>
__global__ void __launch_bounds__(256, 6) k_finish(const uint8_t *__restrict__ src, int H, int W,
                                                const int32_t *__restrict__ minmax, uint8_t *__restrict__ enhanced,
                                                uint8_t *__restrict__ gray, uint8_t *__restrict__ blurred,
                                                int32_t *__restrict__ hist)
{
    constexpr int FW = 120, FH = 36, R = 2, PADX = 4;
    constexpr int GH = FH + 2 * R;                 // 40 staged rows: 5 per warp
    constexpr int SW = FW + 2 * PADX;              // 128 gray columns: X = x0-4 .. x0+123, 32 groups
    __shared__ __align__(16) uint8_t s_g[GH][SW];
    __shared__ __align__(8) uint16_t s_h[GH][FW];
    __shared__ int s_hist[8][256];
    __shared__ uint8_t s_map[256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, frame = blockIdx.z;
    const int x0 = blockIdx.x * FW, y0 = blockIdx.y * FH;
    const size_t fo = (size_t)frame * H * W;
    const uint8_t *img = src + fo * 3;
    if (hist)
        for (int i = tid; i < 8 * 256; i += 256) (&s_hist[0][0])[i] = 0;
    bool identity = true;
    if (NORM) {
        const int lo = minmax[2 * frame], hi = minmax[2 * frame + 1];
        identity = (lo == 0 && hi == 255);
        if (!identity) s_map[tid] = (uint8_t)normalize_value(tid, lo, hi);
        __syncthreads();
    }
    // ---- 1 ----
    const int X = x0 - PADX + 4 * lane;            // first pixel of this lane's group
    const bool inside = X >= 0 && X + 3 < W;
    const bool interior_col = lane >= 1 && lane <= FW / 4 && X < W;
    // Tiles whose 40 staged rows all exist (no mirrored or missing rows): an aligned 4-pixel group advances by a
    // constant number of words per staged row, so the five rows of a warp are loaded up front and every address
    // is one pointer plus a compile-time multiple of the row stride.
    const bool rows_inside = FAST && y0 >= R && y0 + FH + R <= H;          // uniform over the block
    const bool fast_lane = rows_inside && inside;
    if (fast_lane) {
        const size_t row32 = (size_t)W * 3 / 4, grow32 = (size_t)W / 4;    // words per image row (W % 4 == 0)
        const size_t p0 = (size_t)(y0 - R + warp) * W + X;
        const uint32_t *p32 = reinterpret_cast<const uint32_t *>(img + p0 * 3);
        uint32_t *e32 = enhanced ? reinterpret_cast<uint32_t *>(enhanced + (fo + p0) * 3) : nullptr;
        uint32_t *g32 = gray ? reinterpret_cast<uint32_t *>(gray + fo + p0) : nullptr;
        uint32_t wv[5][3];
#pragma unroll
        for (int it = 0; it < 5; ++it) {
            wv[it][0] = __ldg(p32 + it * 8 * row32); wv[it][1] = __ldg(p32 + it * 8 * row32 + 1);
            wv[it][2] = __ldg(p32 + it * 8 * row32 + 2);
        }
#pragma unroll
        for (int it = 0; it < 5; ++it) {
            const int ly = warp + 8 * it;
            uint32_t w0 = wv[it][0], w1 = wv[it][1], w2 = wv[it][2];
            if (NORM && !identity) {
                w0 = s_map[w0 & 0xff] | (s_map[(w0 >> 8) & 0xff] << 8) | (s_map[(w0 >> 16) & 0xff] << 16) | ((uint32_t)s_map[w0 >> 24] << 24);
                w1 = s_map[w1 & 0xff] | (s_map[(w1 >> 8) & 0xff] << 8) | (s_map[(w1 >> 16) & 0xff] << 16) | ((uint32_t)s_map[w1 >> 24] << 24);
                w2 = s_map[w2 & 0xff] | (s_map[(w2 >> 8) & 0xff] << 8) | (s_map[(w2 >> 16) & 0xff] << 16) | ((uint32_t)s_map[w2 >> 24] << 24);
            }
            const uint32_t gpack = gray4_from_words(w0, w1, w2);
            if (interior_col && ly >= R && ly < GH - R) {
                if (enhanced) { e32[it * 8 * row32] = w0; e32[it * 8 * row32 + 1] = w1; e32[it * 8 * row32 + 2] = w2; }
                if (gray) g32[it * 8 * grow32] = gpack;
            }
            CVB_BOUNDS(ly >= 0 && ly < GH && 4 * lane + 3 < SW);
            *reinterpret_cast<uint32_t *>(&s_g[ly][4 * lane]) = gpack;
        }
    }
    for (int ly = warp; ly < GH && !fast_lane; ly += 8) {
        const int Y = y0 - R + ly;
        const int sy = reflect101(Y, H);
        const uint8_t *rowp = img + (size_t)sy * W * 3;
        const bool interior = interior_col && ly >= R && ly < GH - R && Y < H;
        uint32_t gpack = 0;
        if (inside) {
            const uint8_t *p = rowp + (size_t)X * 3;
            uint32_t w0, w1, w2;
            if (FAST) {
                const uint32_t *p32 = reinterpret_cast<const uint32_t *>(p);
                w0 = __ldg(p32); w1 = __ldg(p32 + 1); w2 = __ldg(p32 + 2);
            } else {
                w0 = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
                w1 = p[4] | (p[5] << 8) | (p[6] << 16) | ((uint32_t)p[7] << 24);
                w2 = p[8] | (p[9] << 8) | (p[10] << 16) | ((uint32_t)p[11] << 24);
            }
            if (NORM && !identity) {
                w0 = s_map[w0 & 0xff] | (s_map[(w0 >> 8) & 0xff] << 8) | (s_map[(w0 >> 16) & 0xff] << 16) | ((uint32_t)s_map[w0 >> 24] << 24);
                w1 = s_map[w1 & 0xff] | (s_map[(w1 >> 8) & 0xff] << 8) | (s_map[(w1 >> 16) & 0xff] << 16) | ((uint32_t)s_map[w1 >> 24] << 24);
                w2 = s_map[w2 & 0xff] | (s_map[(w2 >> 8) & 0xff] << 8) | (s_map[(w2 >> 16) & 0xff] << 16) | ((uint32_t)s_map[w2 >> 24] << 24);
            }
            gpack = gray4_from_words(w0, w1, w2);
            if (interior) {
                const size_t po = (size_t)Y * W + X;
                if (enhanced) {
                    uint8_t *o = enhanced + (fo + po) * 3;
                    if (FAST) {
                        uint32_t *o32 = reinterpret_cast<uint32_t *>(o);
                        o32[0] = w0; o32[1] = w1; o32[2] = w2;
                    } else {
                        const uint32_t ww[3] = {w0, w1, w2};
                        for (int i = 0; i < 12; ++i) o[i] = (uint8_t)(ww[i >> 2] >> (8 * (i & 3)));
                    }
                }
                if (gray) {
                    uint8_t *o = gray + fo + po;
                    if (FAST) *reinterpret_cast<uint32_t *>(o) = gpack;
                    else
                        for (int i = 0; i < 4; ++i) o[i] = (uint8_t)(gpack >> (8 * i));
                }
            }
        } else {
            // group touching / beyond an image edge: per pixel, with mirrored columns
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int Xj = X + j;
                if (Xj < x0 - R || Xj >= x0 + FW + R || Xj >= W + R) continue;   // never read by a valid output
                const uint8_t *p = rowp + (size_t)reflect101(Xj, W) * 3;
                int b = p[0], g = p[1], r = p[2];
                if (NORM && !identity) { b = s_map[b]; g = s_map[g]; r = s_map[r]; }
                const int gv = gray_px(b, g, r);
                gpack |= (uint32_t)gv << (8 * j);
                if (Xj >= x0 && Xj < W && Xj < x0 + FW && ly >= R && ly < GH - R && Y < H) {
                    const size_t po = (size_t)Y * W + Xj;
                    if (enhanced) {
                        uint8_t *o = enhanced + (fo + po) * 3;
                        o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r;
                    }
                    if (gray) gray[fo + po] = (uint8_t)gv;
                }
            }
        }
        *reinterpret_cast<uint32_t *>(&s_g[ly][4 * lane]) = gpack;
    }
    __syncthreads();
    // ---- 2: horizontal pass, 4 outputs per lane (output x <-> gray columns x+2 .. x+6), lanes 0..29 ----
    if (lane < FW / 4) {
        const int x = 4 * lane;
        for (int ly = warp; ly < GH; ly += 8) {
            const uint32_t *gp = reinterpret_cast<const uint32_t *>(&s_g[ly][x]);
            const uint32_t a = gp[0], b = gp[1], c = gp[2];          // columns x .. x+11
            // h_k = v[k+2] + 4 v[k+3] + 6 v[k+4] + 4 v[k+5] + v[k+6]: one 4x8-bit dot product (weights 1,4,6,4) + the fifth tap
            constexpr uint32_t kW = 1u | (4u << 8) | (6u << 16) | (4u << 24);
            const uint32_t h0 = __dp4a(__byte_perm(a, b, 0x5432), kW, (b >> 16) & 0xffu);
            const uint32_t h1 = __dp4a(__byte_perm(a, b, 0x6543), kW, b >> 24);
            const uint32_t h2 = __dp4a(b, kW, c & 0xffu);
            const uint32_t h3 = __dp4a(__byte_perm(b, c, 0x4321), kW, (c >> 8) & 0xffu);
            *reinterpret_cast<uint2 *>(&s_h[ly][x]) = make_uint2(h0 | (h1 << 16), h2 | (h3 << 16));
        }
    }
    __syncthreads();
    // ---- 3: vertical pass + histogram ----
    int *myh = s_hist[warp];
    if (lane < FW / 4) {
        const int x = 4 * lane, Xo = x0 + x;
        if (Xo < W) {
            const int nvalid = FAST ? 4 : min(4, W - Xo);
            uint8_t *orow = blurred ? blurred + fo + (size_t)(y0 + warp) * W + Xo : nullptr;
            const uint16_t *hrow = &s_h[warp][x];
#pragma unroll
            for (int it = 0; it < (FH + 7) / 8; ++it, hrow += 8 * FW, orow += (size_t)8 * W) {
                const int ly = warp + 8 * it;
                if (ly >= FH || y0 + ly >= H) break;
                const uint2 r0 = *reinterpret_cast<const uint2 *>(hrow);
                const uint2 r1 = *reinterpret_cast<const uint2 *>(hrow + FW);
                const uint2 r2 = *reinterpret_cast<const uint2 *>(hrow + 2 * FW);
                const uint2 r3 = *reinterpret_cast<const uint2 *>(hrow + 3 * FW);
                const uint2 r4 = *reinterpret_cast<const uint2 *>(hrow + 4 * FW);
                CVB_BOUNDS(ly + 4 < GH && x + 3 < FW && y0 + ly < H && (FAST ? Xo + 3 < W : Xo < W));
                // two 16-bit lanes per word; each lane's sum is <= 255*256 = 65280, so no carry crosses lanes
                const uint32_t lo = r0.x + 4 * r1.x + 6 * r2.x + 4 * r3.x + r4.x;
                const uint32_t hi = r0.y + 4 * r1.y + 6 * r2.y + 4 * r3.y + r4.y;
                const int o0 = ((lo & 0xffff) + 128) >> 8, o1 = ((lo >> 16) + 128) >> 8;
                const int o2 = ((hi & 0xffff) + 128) >> 8, o3 = ((hi >> 16) + 128) >> 8;
                const int ov[4] = {o0, o1, o2, o3};
                if (hist) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (j < nvalid) atomicAdd(&myh[ov[j]], 1);
                }
                if (blurred) {
                    if (FAST)
                        *reinterpret_cast<uint32_t *>(orow) = (uint32_t)o0 | ((uint32_t)o1 << 8) | ((uint32_t)o2 << 16) | ((uint32_t)o3 << 24);
                    else
                        for (int j = 0; j < nvalid; ++j) orow[j] = (uint8_t)ov[j];
                }
            }
        }
    }
    if (hist) {
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) tot += s_hist[w][tid];
        if (tot) atomicAdd(&hist[(size_t)frame * 256 + tid], tot);
    }
}
int launch_finish(cvb_handle *h, const uint8_t *src, int n, int H, int W, const int32_t *minmax, uint8_t *enhanced,
                  uint8_t *gray, uint8_t *blurred, int32_t *hist)
{
    if (hist) CVB_CHECK_CUDA(cudaMemsetAsync(hist, 0, sizeof(int32_t) * 256 * (size_t)n, h->stream));
    dim3 grid((W + 119) / 120, (H + 35) / 36, n);
    const uintptr_t bits = reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(enhanced) |
                           reinterpret_cast<uintptr_t>(gray) | reinterpret_cast<uintptr_t>(blurred);
    const bool fast = (W % 4 == 0) && (bits & 3) == 0;
    PROF(h, "k_finish");
    if (minmax) {
        if (fast) k_finish<true, true><<<grid, 256, 0, h->stream>>>(src, H, W, minmax, enhanced, gray, blurred, hist);
        else k_finish<true, false><<<grid, 256, 0, h->stream>>>(src, H, W, minmax, enhanced, gray, blurred, hist);
    } else {
        if (fast) k_finish<false, true><<<grid, 256, 0, h->stream>>>(src, H, W, nullptr, enhanced, gray, blurred, hist);
        else k_finish<false, false><<<grid, 256, 0, h->stream>>>(src, H, W, nullptr, enhanced, gray, blurred, hist);
    }
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// ---------------------------------------------------------------------------------------
// S9: Otsu threshold, OpenCV's f64 loop verbatim in structure (thresh.cpp
// getThreshVal_Otsu_8u): sequential recurrences, strict '>' so the first
// maximum wins.  One warp per frame; lane 0 walks the 256 bins.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_otsu(const int32_t *__restrict__ hist, long npx, int32_t *__restrict__ otsu_t)
{
    // Only the (q1, mu1) recurrence is sequential; p_i, i*p_i and mu before it, and mu2 / sigma / arg-max
    // after it, are spread over the warp.  Every operation keeps OpenCV's operands and rounding.
    __shared__ double s_p[256], s_ip[256], s_q1[256], s_mu1[256];
    __shared__ unsigned char s_ok[256];
    const int frame = blockIdx.x, lane = threadIdx.x;
    const int32_t *hh = hist + (size_t)frame * 256;
    const double scale = __ddiv_rn(1.0, (double)npx);
    double part = 0;     // sum of i*h[i]: integers below 2^53, exact in any order
    for (int i = lane; i < 256; i += 32) {
        const double hv = (double)hh[i];
        const double p = __dmul_rn(hv, scale);
        s_p[i] = p;
        s_ip[i] = __dmul_rn((double)i, p);
        part += __dmul_rn((double)i, hv);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const double mu = __dmul_rn(part, scale);
    __syncwarp();
    if (lane == 0) {
        double mu1 = 0, q1 = 0;
        for (int i = 0; i < 256; ++i) {
            mu1 = __dmul_rn(mu1, q1);
            q1 = __dadd_rn(q1, s_p[i]);
            const double q2 = __dsub_rn(1.0, q1);
            const bool ok = !(fmin(q1, q2) < (double)FLT_EPSILON || fmax(q1, q2) > 1.0 - (double)FLT_EPSILON);
            if (ok) mu1 = __ddiv_rn(__dadd_rn(mu1, s_ip[i]), q1);
            s_q1[i] = q1; s_mu1[i] = mu1; s_ok[i] = ok;
        }
    }
    __syncwarp();
    double best = 0;      // `sigma > max_sigma` with max_sigma starting at 0: the first strict maximum wins
    int best_i = 0;
    for (int i = lane; i < 256; i += 32) {
        if (!s_ok[i]) continue;
        const double q1 = s_q1[i], mu1 = s_mu1[i], q2 = __dsub_rn(1.0, q1);
        const double mu2 = __ddiv_rn(__dsub_rn(mu, __dmul_rn(q1, mu1)), q2);
        const double d = __dsub_rn(mu1, mu2);
        const double sigma = __dmul_rn(__dmul_rn(__dmul_rn(q1, q2), d), d);
        if (sigma > best) { best = sigma; best_i = i; }      // ascending i per lane: first maximum of the lane
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob > best || (ob == best && ob > 0 && oi < best_i)) { best = ob; best_i = oi; }
    }
    if (lane == 0) otsu_t[frame] = best > 0 ? best_i : 0;
}
int launch_otsu(cvb_handle *h, const int32_t *hist, int n, long npx, int32_t *otsu_t)
{
    PROF(h, "k_otsu");
    k_otsu<<<n, 32, 0, h->stream>>>(hist, npx, otsu_t);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

// Pass 4: binary mask, 16 pixels per thread per step
__global__ void __launch_bounds__(256) k_threshold(const uint8_t *__restrict__ src, long npx,
                                                   const int32_t *__restrict__ otsu_t, uint8_t *__restrict__ dst)
{
    const int frame = blockIdx.y;
    const uint32_t T = (uint32_t)otsu_t[frame];
    const uint8_t *p = src + (size_t)frame * npx;
    uint8_t *o = dst + (size_t)frame * npx;
    const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(o)) & 15) == 0;
    const long nvec = al ? npx / 16 : 0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (long)gridDim.x * 256) {
        uint4 q = __ldg(reinterpret_cast<const uint4 *>(p) + i);
        uint32_t *w = &q.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t v = w[k];
            w[k] = ((v & 0xff) > T ? 0xffu : 0u) | (((v >> 8) & 0xff) > T ? 0xff00u : 0u) |
                   (((v >> 16) & 0xff) > T ? 0xff0000u : 0u) | ((v >> 24) > T ? 0xff000000u : 0u);
        }
        reinterpret_cast<uint4 *>(o)[i] = q;
    }
    for (long i = nvec * 16 + (long)blockIdx.x * 256 + threadIdx.x; i < npx; i += (long)gridDim.x * 256)
        o[i] = p[i] > T ? 255 : 0;
}
int launch_threshold(cvb_handle *h, const uint8_t *src, int n, long npx, const int32_t *otsu_t, uint8_t *dst)
{
    int bx = stream_blocks(npx / 16, 16 * h->sm_count / (n > 0 ? n : 1));
    dim3 grid(bx, n);
    PROF(h, "k_threshold");
    k_threshold<<<grid, 256, 0, h->stream>>>(src, npx, otsu_t, dst);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

CVB_BOUNDS_TU(enhance)
