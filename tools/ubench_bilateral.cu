// Bilateral stage alone: every CTA runs the item functions of csrc/cvb_fused2.cu on one synthetic lighting tile again
// and again (no lighting stage, no sharpen stage, no barriers in the loop), to separate what the stage costs by itself
// from what it costs inside the fused kernel.  Prints clocks per output pixel per warp and the frame time they imply.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I chessboard_vision_b200/csrc -o ubench_bilateral tools/ubench_bilateral.cu
#include "../chessboard_vision_b200/csrc/cvb_fused2.cu"
#include <cstdio>
#include <cstdarg>
#include <vector>
#include <random>

void cvb_set_error(const char *, ...) {}
void cvb_prof_begin(cvb_handle *, const char *) {}
void cvb_prof_end(cvb_handle *) {}

namespace {
template <bool LUTP, int ABL>
CVB_DEV void bilateral_item_abl(const uint32_t *sA, uint32_t *sB, const float *sW, const float *myW, const Fused2Args &a,
                                   int row0, int r4)
{
    unsigned dummy = 0;
    float wsum[2][4], sb[2][4], sg[2][4], sr[2][4];
    uint32_t ctr[2][4];
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
        const uint4 c = *reinterpret_cast<const uint4 *>(sA + (row0 + tt + 4) * AW + r4 + 4);
        ctr[tt][0] = c.x; ctr[tt][1] = c.y; ctr[tt][2] = c.z; ctr[tt][3] = c.w;
#pragma unroll
        for (int j = 0; j < 4; ++j) wsum[tt][j] = sb[tt][j] = sg[tt][j] = sr[tt][j] = 0.f;
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        const uint32_t *rowp = sA + (row0 + k) * AW + r4;
        uint32_t px[12];
        float fb[12], fg[12], fr[12];
        int lo = 4;
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
            const int dy = k - 4 - tt, ady = dy < 0 ? -dy : dy;
            if (ady <= 4) lo = min(lo, 4 - row_half(dy));
        }
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            if (v != 1 && lo >= 4) continue;
            const uint4 q = *reinterpret_cast<const uint4 *>(rowp + 4 * v);
            px[4 * v] = q.x; px[4 * v + 1] = q.y; px[4 * v + 2] = q.z; px[4 * v + 3] = q.w;
        }
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            if (c < lo || c > 11 - lo) continue;
            if (ABL == 3) { fb[c] = __uint_as_float(px[c]); fg[c] = __uint_as_float(px[c] ^ 0x3f000000u); fr[c] = __uint_as_float(px[c] | 0x3f800000u); }
            else {
            unpk2(add2(pk2u(__byte_perm(px[c], 0x4B000000u, 0x7540u), __byte_perm(px[c], 0x4B000000u, 0x7541u)),
                       0xCB000000CB000000ull), fb[c], fg[c]);
            fr[c] = (float)((px[c] >> 16) & 0xffu);      // through the conversion unit (I2F.U8), idle otherwise
            }
        }
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
            const int dy = k - 4 - tt;
            if (dy < -4 || dy > 4) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int dx = -4; dx <= 4; ++dx) {
                    if (dy * dy + dx * dx > 16) continue;
                    const int c = j + 4 + dx;
                    float w;
                    if (dy == 0 && dx == 0) {
                        w = 1.0f;
                    } else {
                        if (ABL == 1) w = a.sw81[(dy + 4) * 9 + dx + 4];                         // no distance, no lookup
                        else if (ABL == 5) w = __uint_as_float(px[c] & 0x3fffffffu);             // a register that exists anyway as the weight
                        else if (ABL == 6) { dummy += __vsadu4(px[c], ctr[tt][j]); w = a.sw81[(dy + 4) * 9 + dx + 4]; }   // distance computed, not used
                        else {
                        const unsigned sad = __vsadu4(px[c], ctr[tt][j]);
                        if (ABL == 4) w = __uint_as_float(0x3f000000u + sad);                    // distance, no lookup
                        else
                        w = LUTP ? __fmul_rn(myW[sad * 32], a.sw81[(dy + 4) * 9 + dx + 4])
                                 : sW[r2_class2(dy * dy + dx * dx) * 768 + sad];
                        }
                    }
                    wsum[tt][j] = __fadd_rn(wsum[tt][j], w);
                    if (ABL != 2) {
                    sb[tt][j] = __fmaf_rn(fb[c], w, sb[tt][j]);
                    sg[tt][j] = __fmaf_rn(fg[c], w, sg[tt][j]);
                    sr[tt][j] = __fmaf_rn(fr[c], w, sr[tt][j]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
        uint4 o;
        uint32_t *op = &o.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float inv = __fdiv_rn(1.0f, wsum[tt][j]);
            op[j] = pack_bgr(round_u8(__fmul_rn(sb[tt][j], inv)), round_u8(__fmul_rn(sg[tt][j], inv)),
                             round_u8(__fmul_rn(sr[tt][j], inv)));
        }
        *reinterpret_cast<uint4 *>(sB + (row0 + tt) * BW + r4) = o;
    }
}

}

template <int NT, bool LUTP, int ACC, int MINB = 1>
__global__ void __launch_bounds__(NT, MINB) k_bil_only(const uint32_t *tile, const float *wlut, const __grid_constant__ Fused2Args a,
                                                   int reps, uint32_t *out, long long *cyc)
{
    constexpr int TH = 64, BH = TH + 2, AH = BH + 8;
    extern __shared__ __align__(16) uint8_t smem[];
    uint32_t *sA = reinterpret_cast<uint32_t *>(smem);
    uint32_t *sB = sA + AW * AH;
    float *sW = reinterpret_cast<float *>(sB + BW * BH);
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < AW * AH; i += NT) sA[i] = tile[i];
    if (LUTP) { for (int k = tid; k < 768 * 32; k += NT) sW[k] = wlut[10 * 768 + (k >> 5)]; }
    else { for (int k = tid; k < 10 * 768; k += NT) sW[k] = wlut[k]; }
    __syncthreads();
    const float *myW = sW + lane;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
        for (int item = tid; item < (BH / 2) * RUNS; item += NT) {
            const int rg = item / RUNS, r4 = (item - rg * RUNS) * 4;
            if (ACC == 2) bilateral_item_packed<LUTP>(sA, sB, sW, myW, a, rg * 2, r4);
            else if (ACC >= 10) bilateral_item_abl<LUTP, ACC - 10>(sA, sB, sW, myW, a, rg * 2, r4);
            else bilateral_item_scalar<LUTP>(sA, sB, sW, myW, a, rg * 2, r4);
        }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * NT + tid] = sB[tid];
}

template <int NT, bool LUTP, int ACC, int MINB = 1>
void run(const char *name, const uint32_t *d_tile, const float *d_w, const Fused2Args &a)
{
    constexpr int TH = 64, BH = TH + 2, AH = BH + 8;
    const size_t smem = (size_t)AW * AH * 4 + (size_t)BW * BH * 4 + (LUTP ? 768 * 32 * 4 : 10 * 768 * 4);
    auto k = k_bil_only<NT, LUTP, ACC, MINB>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    uint32_t *out; long long *cyc;
    constexpr int NCTA = 148 * MINB;                             // MINB CTAs resident per SM, each on its own tile
    cudaMalloc(&out, NCTA * NT * 4); cudaMalloc(&cyc, NCTA * 8);
    const int reps = 40;
    k<<<NCTA, NT, smem>>>(d_tile, d_w, a, 2, out, cyc);
    k<<<NCTA, NT, smem>>>(d_tile, d_w, a, reps, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long c[NCTA]; cudaMemcpy(c, cyc, sizeof c, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < NCTA; ++i) avg += c[i]; avg /= NCTA;
    const double px = (double)BW * BH * reps * MINB;             // bilateral outputs per SM in that time
    const double clk_per_px_sm = avg / px;                        // SM clocks per output pixel
    // a 1080p frame: 272 tiles of 124 x 66 bilateral outputs over 148 SMs at 1.965 GHz
    printf("%-52s %s  %.3f SM-clk per pixel = %.1f clk per warp-pixel-group and SMSP -> %.1f us per 1080p frame\n", name,
           e == cudaSuccess ? "" : cudaGetErrorString(e), clk_per_px_sm, clk_per_px_sm * 128, clk_per_px_sm * 272.0 * BW * BH / 148 / 1965.0);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    constexpr int AH = 74;
    std::vector<uint32_t> tile(AW * AH);
    std::mt19937 rng(1);
    std::normal_distribution<float> noise(0.f, 20.f);            // sigma 8 amplified by CLAHE, as in the board frames
    for (int y = 0; y < AH; ++y)
        for (int x = 0; x < AW; ++x) {
            const int base = ((x / 40 + y / 30) & 1) ? 170 : 70;
            auto ch = [&](float s) { int v = (int)(base * s + noise(rng)); return (uint32_t)(v < 0 ? 0 : v > 255 ? 255 : v); };
            tile[y * AW + x] = ch(1.f) | (ch(0.8f) << 8) | (ch(0.6f) << 16) | 0x01000000u;
        }
    std::vector<float> color(768), space(81), wl(11 * 768);
    for (int i = 0; i < 768; ++i) color[i] = (float)exp((double)i * i * (-0.5 / (75.0 * 75.0)));
    for (int dy = -4; dy <= 4; ++dy) for (int dx = -4; dx <= 4; ++dx) space[(dy + 4) * 9 + dx + 4] = (float)exp((dy * dy + dx * dx) * (-0.5 / (75.0 * 75.0)));
    const int r2[10] = {0, 1, 2, 4, 5, 8, 9, 10, 13, 16};
    for (int c = 0; c < 10; ++c) for (int i = 0; i < 768; ++i) wl[c * 768 + i] = (float)exp(r2[c] * (-0.5 / (75.0 * 75.0))) * color[i];
    for (int i = 0; i < 768; ++i) wl[10 * 768 + i] = color[i];
    uint32_t *d_tile; float *d_w;
    cudaMalloc(&d_tile, tile.size() * 4); cudaMemcpy(d_tile, tile.data(), tile.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&d_w, wl.size() * 4); cudaMemcpy(d_w, wl.data(), wl.size() * 4, cudaMemcpyHostToDevice);
    Fused2Args a; memset(&a, 0, sizeof a);
    memcpy(a.sw81, space.data(), sizeof a.sw81);
    for (int dy = -4; dy <= 4; ++dy) {
        const int nt = row_taps(dy);
        for (int p = 0; 2 * p + 1 < nt; ++p) {
            a.swp[dy + 4][p].x = space[(dy + 4) * 9 + tap_dx(dy, 2 * p) + 4];
            a.swp[dy + 4][p].y = space[(dy + 4) * 9 + tap_dx(dy, 2 * p + 1) + 4];
        }
        a.sws[dy + 4] = (nt & 1) ? space[(dy + 4) * 9 + tap_dx(dy, nt - 1) + 4] : 0.f;
    }
    run<1024, true, 0>("private table, scalar accumulation, 1024 threads", d_tile, d_w, a);
    run<512, true, 0>("private table, scalar accumulation,  512 threads", d_tile, d_w, a);
    run<1024, false, 0>("folded table,  scalar accumulation, 1024 threads", d_tile, d_w, a);
    run<512, true, 2>("private table, packed accumulation,  512 threads", d_tile, d_w, a);
    run<1024, true, 2>("private table, packed accumulation, 1024 threads", d_tile, d_w, a);
    run<1024, false, 2>("folded table,  packed accumulation, 1024 threads", d_tile, d_w, a);
    // ablations of the scalar private-table form (not the filter any more: which part costs what)
    // two CTAs per SM, as the shipped kernel runs (product table: two private tables do not fit)
    run<512, false, 0, 2>("2 CTAs x 512, product table, scalar (the shipped form)", d_tile, d_w, a);
    run<512, false, 2, 2>("2 CTAs x 512, product table, packed (64 registers)", d_tile, d_w, a);
    run<384, false, 2, 2>("2 CTAs x 384, product table, packed (85 registers)", d_tile, d_w, a);
    run<352, false, 2, 2>("2 CTAs x 352, product table, packed (93 registers)", d_tile, d_w, a);
    run<256, false, 2, 2>("2 CTAs x 256, product table, packed (128 registers)", d_tile, d_w, a);
    run<352, false, 0, 2>("2 CTAs x 352, product table, scalar", d_tile, d_w, a);
    run<1024, true, 10>("ablation: nothing removed", d_tile, d_w, a);
    run<1024, true, 11>("ablation: constant weights (no VABSDIFF4 / LEA / LDS / FMUL)", d_tile, d_w, a);
    run<1024, true, 14>("ablation: distance but no table lookup (no LEA / LDS / FMUL)", d_tile, d_w, a);
    run<1024, true, 12>("ablation: weight sum only (no 3 FFMA per tap)", d_tile, d_w, a);
    run<1024, true, 13>("ablation: no byte -> float conversions", d_tile, d_w, a);
    run<1024, true, 15>("ablation: weight = an existing per-lane register (no VABSDIFF4 / LEA / LDS / FMUL)", d_tile, d_w, a);
    run<1024, true, 16>("ablation: constant weights, distance computed but unused", d_tile, d_w, a);
    return 0;
}
