// Pass 2 of the frame_enhancer path as ONE persistent, TMA-fed kernel (sm_100a):
//   CLAHE blend of L + Lab->BGR (frame_enhancer.py:111-120)  ->  bilateral d=9 (:131)  ->  3x3 sharpen (:138)
//   + global min/max for the normalize step (:146).
//
// Same arithmetic, tap order and roundings as k_fused in cvb_enhance.cu (which stays the general path for odd sizes
// and the stage-isolated entry points); what differs is how the work is fed and issued:
//
//   * one CTA per SM walks the (frame, tile) list; the colour tables and the bilateral weight table are loaded once
//     per CTA instead of once per tile;
//   * the Lab tile + halo (132 x 74 pixels, 3 bytes each) is fetched by the copy engine from a 3-D tensor map over
//     the Lab frames (cp.async.bulk.tensor, UTMALDG in SASS; u16 elements, box 208 x 74 x 1 = 416-byte rows) into
//     shared memory while the previous tile is still in its bilateral / sharpen stages.  The copy engine wants the
//     first byte of a box row on a 16-byte boundary of the frame row (measured: any other start raises an illegal
//     instruction), so the box starts at the boundary at or below the tile's first byte and the tile begins 6 or
//     14 bytes into each staged row; out-of-image parts arrive as zeros and the REFLECT_101 ring is patched in
//     shared memory for the tiles on an image edge;
//   * the range weights sit in a lane-private table (768 entries x 32 lanes, one bank per lane), so a lookup is one
//     shared-memory wavefront whatever the colour distances of the warp are (the folded 10 x 768 table of k_fused
//     averaged 2.5); the spatial factor is applied by a packed multiply for two taps at a time (FMUL2) and gives
//     the same f32 product the folded table holds;
//   * pixels carry a constant 1 in their fourth byte: one PRMT/FADD2 sequence turns a word into the float pairs
//     (b, g) and (r, 1), and a tap is two packed FMAs, (sb, sg) += (b, g) * w and (sr, wsum) += (r, 1) * w, with
//     the weight broadcast by the instruction (FFMA2 ..., R.F32).  fma(1, w, wsum) is the IEEE sum w + wsum, so
//     the four running sums are bit-identical to the scalar version and to the oracle;
//   * only two block-wide barriers per tile: the sharpen stage of tile n and the lighting stage of tile n+1 touch
//     disjoint buffers and run back to back.
#include "cvb_device.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>

namespace {

constexpr int TW = 120, BX = 2, BY = 1, AR = 4;
constexpr int BW = TW + 2 * BX;                            // bilateral outputs kept per tile row (sharpen halo)
constexpr int AW = BW + 2 * AR;                            // lighting outputs per tile row (bilateral halo)
constexpr int RAW_PITCH = 416;                             // bytes per staged Lab row: 132 * 3 = 396, + up to 15 bytes of lead-in (see raw_lead)
constexpr int RUNS = BW / 4;                               // runs of 4 pixels per bilateral row
static_assert(AW % 2 == 0 && BW % 4 == 0 && AW * 3 + 15 <= RAW_PITCH && RAW_PITCH % 16 == 0, "tile geometry");

struct __align__(16) Axis2 {     // one per staged column / row
    float a;                     // CLAHE blend factor towards the second tile
    uint32_t t1, t2;             // byte offsets of the two CLAHE tile LUTs along this axis
    uint32_t pad;
};

struct Fused2Args {
    uint8_t *dst;
    int H, W, n;
    const CvbTables *tabs;
    const uint8_t *lut;          // CLAHE LUTs of all frames
    ClaheGeom g;
    const float *wlut;           // device: [10][768] space x colour products, then [768] colour weights
    int32_t *minmax;             // per frame {min, max}
    int tiles_x, tiles_y, total_tiles;
    float2 swp[9][4];            // spatial weights of the tap pairs of window row dy + 4 (see tap_dx)
    float sws[9];                // spatial weight of the row's unpaired last tap
    float sw81[81];              // spatial weights [dy + 4][dx + 4]
};

// Shared memory of a CTA: NG groups of threads, each with its own staged Lab tile, lighting tile (A), bilateral
// tile (B), axis records and TMA barrier; one weight table and one set of colour tables for all groups.
template <int NG, int TH, bool LUTP>
struct Smem2 {
    static constexpr int BH = TH + 2 * BY, AH = BH + 2 * AR;
    static constexpr size_t rawBytes = (size_t)RAW_PITCH * AH;
    static constexpr size_t gRaw = 0;
    static constexpr size_t gA = (rawBytes + 127) / 128 * 128;
    static constexpr size_t gB = gA + (size_t)AW * AH * 4;
    static constexpr size_t gX = gB + (size_t)BW * BH * 4;
    static constexpr size_t gBytes = (gX + (size_t)(AW + AH) * sizeof(Axis2) + 127) / 128 * 128;
    static constexpr size_t offW = gBytes * NG;
    static constexpr size_t offT = offW + (LUTP ? 768 * 32 * 4 : 10 * 768 * 4);
    static constexpr size_t offBar = offT + sizeof(SmemColorTables);
    static constexpr size_t bytes = offBar + 8 * sizeof(uint64_t);
    static_assert(gA % 16 == 0 && gB % 16 == 0 && gX % 16 == 0 && offW % 16 == 0 && offT % 16 == 0 && offBar % 8 == 0, "alignment");
    static_assert(bytes <= 232448, "more than 227 KB of shared memory");
};

// taps of window row dy in accumulation order, the centre of row 0 left out: half-width of the disc r <= 4
__host__ __device__ constexpr int row_half(int dy) { return dy == 4 || dy == -4 ? 0 : (dy == 3 || dy == -3 ? 2 : (dy == 0 ? 4 : 3)); }
__host__ __device__ constexpr int row_taps(int dy) { return 2 * row_half(dy) + (dy == 0 ? 0 : 1); }
__host__ __device__ constexpr int tap_dx(int dy, int idx)
{
    return dy == 0 ? (idx < 4 ? idx - 4 : idx - 3) : idx - row_half(dy);
}
__host__ __device__ constexpr int r2_class2(int r2)
{
    return r2 == 0 ? 0 : r2 == 1 ? 1 : r2 == 2 ? 2 : r2 == 4 ? 3 : r2 == 5 ? 4 : r2 == 8 ? 5 : r2 == 9 ? 6 : r2 == 10 ? 7
         : r2 == 13 ? 8 : 9;
}

typedef unsigned long long u64;
CVB_DEV u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
CVB_DEV u64 pk2u(uint32_t lo, uint32_t hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r; }
CVB_DEV void unpk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
CVB_DEV u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
CVB_DEV u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
CVB_DEV u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

CVB_DEV void tma_load_tile(void *dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_addr(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_addr(bar))
                 : "memory");
}

// bytes between the 16-byte boundary the box starts on and the first byte of staged column 0 (x = x0 - 6)
CVB_DEV int raw_lead(int x0) { return ((x0 - BX - AR) * 3) & 15; }
// first box coordinate (u16 elements) of the tile at x0
CVB_DEV int raw_c0(int x0) { return ((x0 - BX - AR) * 3 - raw_lead(x0)) / 2; }

// barrier over the threads of one group (named barrier 1 + group) or over the CTA
template <int NG, int NTG>
CVB_DEV void group_sync(int group)
{
    if (NG == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(NTG) : "memory");
}

template <int TH>
struct TilePos { int frame, x0, y0; };
template <int TH>
CVB_DEV TilePos<TH> tile_pos(int t, const Fused2Args &a)
{
    const int tpf = a.tiles_x * a.tiles_y;
    TilePos<TH> p;
    p.frame = t / tpf;
    const int r = t - p.frame * tpf, ty = r / a.tiles_x;
    p.y0 = ty * TH; p.x0 = (r - ty * a.tiles_x) * TW;
    return p;
}

template <int NTG, int AH>
CVB_DEV void fill_axis(Axis2 *sAx, const Fused2Args &a, int x0, int y0, int tid)
{
    const int ax0 = x0 - BX - AR, ay0 = y0 - BY - AR;
    for (int i = tid; i < AW + AH; i += NTG) {
        const bool col = i < AW;
        const int p = col ? reflect101(ax0 + i, a.W) : reflect101(ay0 + (i - AW), a.H);
        const ClaheAxis ca = col ? clahe_axis(p, a.g.inv_tw, a.g.tiles_x) : clahe_axis(p, a.g.inv_th, a.g.tiles_y);
        const uint32_t unit = col ? 256u : 256u * (uint32_t)a.g.tiles_x;
        Axis2 ai;
        ai.a = ca.a; ai.t1 = (uint32_t)ca.i1 * unit; ai.t2 = (uint32_t)ca.i2 * unit; ai.pad = 0;
        sAx[i] = ai;
    }
}

// one staged pixel: CLAHE blend of L (clahe.cpp interpolation body, unfused f32) and Lab -> BGR; byte 3 = 1
CVB_DEV uint32_t light_px(const SmemColorTables *sTab, const uint8_t *__restrict__ lut, const Axis2 &cx, const Axis2 &cy,
                          int L, int A, int B)
{
    const uint32_t o1 = cy.t1 + (uint32_t)L, o2 = cy.t2 + (uint32_t)L;
    const float l11 = (float)__ldg(lut + (o1 + cx.t1)), l12 = (float)__ldg(lut + (o1 + cx.t2));
    const float l21 = (float)__ldg(lut + (o2 + cx.t1)), l22 = (float)__ldg(lut + (o2 + cx.t2));
    const float xa1 = __fsub_rn(1.0f, cx.a), ya1 = __fsub_rn(1.0f, cy.a);
    const float top = __fmul_rn(__fadd_rn(__fmul_rn(l11, xa1), __fmul_rn(l12, cx.a)), ya1);
    const float bot = __fmul_rn(__fadd_rn(__fmul_rn(l21, xa1), __fmul_rn(l22, cx.a)), cy.a);
    return lab2bgr_px(sTab, blend_u8(__fadd_rn(top, bot)), A, B) | 0x01000000u;
}

// ---- B, packed accumulation: (sb, sg) += (b, g) * w, (sr, wsum) += (r, 1) * w, two taps' spatial factors per FMUL2 ----
template <bool LUTP>
CVB_DEV void bilateral_item_packed(const uint32_t *sA, uint32_t *sB, const float *sW, const float *myW, const Fused2Args &a,
                                   int row0, int r4)
{
    u64 accA[2][4], accB[2][4];         // (sb, sg), (sr, wsum)
    uint32_t ctr[2][4];
#pragma unroll
    for (int tt = 0; tt < 2; ++tt) {
        const uint4 c = *reinterpret_cast<const uint4 *>(sA + (row0 + tt + 4) * AW + r4 + 4);
        ctr[tt][0] = c.x; ctr[tt][1] = c.y; ctr[tt][2] = c.z; ctr[tt][3] = c.w;
#pragma unroll
        for (int j = 0; j < 4; ++j) accA[tt][j] = accB[tt][j] = 0ull;
    }
#pragma unroll
    for (int k = 0; k < 10; ++k) {           // window row k is A row (row0 + k): dy = k - 4 - tt for output row tt
        const uint32_t *rowp = sA + (row0 + k) * AW + r4;
        uint32_t px[12];
        u64 bg[12], r1[12];
        int lo = 4;                           // first needed column (of 12) over the output rows this row feeds
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
            // bytes -> floats: PRMT builds 2^23 + v in each half, one packed add removes 2^23
            bg[c] = add2(pk2u(__byte_perm(px[c], 0x4B000000u, 0x7540u), __byte_perm(px[c], 0x4B000000u, 0x7541u)),
                         0xCB000000CB000000ull);
            r1[c] = add2(pk2u(__byte_perm(px[c], 0x4B000000u, 0x7542u), __byte_perm(px[c], 0x4B000000u, 0x7543u)),
                         0xCB000000CB000000ull);
        }
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
            const int dy = k - 4 - tt;
            if (dy < -4 || dy > 4) continue;
            const int nt = row_taps(dy);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int p = 0; p <= nt / 2; ++p) {
                    if (dy == 0 && p == 2) {            // the centre tap sits between the pairs (-2,-1) and (1,2)
                        const u64 one = pk2(1.0f, 1.0f);
                        accA[tt][j] = fma2(bg[j + 4], one, accA[tt][j]);
                        accB[tt][j] = fma2(r1[j + 4], one, accB[tt][j]);
                    }
                    if (2 * p + 1 < nt) {
                        const int ca = j + 4 + tap_dx(dy, 2 * p), cb = j + 4 + tap_dx(dy, 2 * p + 1);
                        const unsigned sa = __vsadu4(px[ca], ctr[tt][j]), sb = __vsadu4(px[cb], ctr[tt][j]);
                        float wa, wb;
                        if (LUTP) {
                            const u64 w2 = mul2(pk2(myW[sa * 32], myW[sb * 32]), pk2(a.swp[dy + 4][p].x, a.swp[dy + 4][p].y));
                            unpk2(w2, wa, wb);
                        } else {
                            wa = sW[r2_class2(dy * dy + (ca - j - 4) * (ca - j - 4)) * 768 + sa];
                            wb = sW[r2_class2(dy * dy + (cb - j - 4) * (cb - j - 4)) * 768 + sb];
                        }
                        accA[tt][j] = fma2(bg[ca], pk2(wa, wa), accA[tt][j]);
                        accB[tt][j] = fma2(r1[ca], pk2(wa, wa), accB[tt][j]);
                        accA[tt][j] = fma2(bg[cb], pk2(wb, wb), accA[tt][j]);
                        accB[tt][j] = fma2(r1[cb], pk2(wb, wb), accB[tt][j]);
                    } else if (2 * p < nt) {            // the row's unpaired last tap
                        const int ca = j + 4 + tap_dx(dy, 2 * p);
                        const unsigned sa = __vsadu4(px[ca], ctr[tt][j]);
                        const float wa = LUTP ? __fmul_rn(myW[sa * 32], a.sws[dy + 4])
                                              : sW[r2_class2(dy * dy + (ca - j - 4) * (ca - j - 4)) * 768 + sa];
                        accA[tt][j] = fma2(bg[ca], pk2(wa, wa), accA[tt][j]);
                        accB[tt][j] = fma2(r1[ca], pk2(wa, wa), accB[tt][j]);
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
            float sb, sg, sr, ws;
            unpk2(accA[tt][j], sb, sg); unpk2(accB[tt][j], sr, ws);
            const float inv = __fdiv_rn(1.0f, ws);
            op[j] = pack_bgr(round_u8(__fmul_rn(sb, inv)), round_u8(__fmul_rn(sg, inv)), round_u8(__fmul_rn(sr, inv)));
        }
        *reinterpret_cast<uint4 *>(sB + (row0 + tt) * BW + r4) = o;
    }
}

// ---- B, scalar accumulation (the loop of k_fused): wsum += w, three FFMA; the FMA pipe does 5 operations per tap ----
template <bool LUTP>
CVB_DEV void bilateral_item_scalar(const uint32_t *sA, uint32_t *sB, const float *sW, const float *myW, const Fused2Args &a,
                                   int row0, int r4)
{
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
            unpk2(add2(pk2u(__byte_perm(px[c], 0x4B000000u, 0x7540u), __byte_perm(px[c], 0x4B000000u, 0x7541u)),
                       0xCB000000CB000000ull), fb[c], fg[c]);
            fr[c] = (float)((px[c] >> 16) & 0xffu);      // through the conversion unit (I2F.U8), idle otherwise
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
                        const unsigned sad = __vsadu4(px[c], ctr[tt][j]);
                        w = LUTP ? __fmul_rn(myW[sad * 32], a.sw81[(dy + 4) * 9 + dx + 4])
                                 : sW[r2_class2(dy * dy + dx * dx) * 768 + sad];
                    }
                    wsum[tt][j] = __fadd_rn(wsum[tt][j], w);
                    sb[tt][j] = __fmaf_rn(fb[c], w, sb[tt][j]);
                    sg[tt][j] = __fmaf_rn(fg[c], w, sg[tt][j]);
                    sr[tt][j] = __fmaf_rn(fr[c], w, sr[tt][j]);
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

// NG groups of NTG threads per CTA; every group walks its own list of TW x TH tiles through the three stages with
// barriers of its own, so that the groups of an SM sit in different stages (the lighting stage is bound by the
// integer pipes and global-load latency, the bilateral stage by the FMA / shared-memory pipes).
template <int NG, int NTG, int TH, bool LUTP, int ACC>
__global__ void __launch_bounds__(NG * NTG, 1) k_fused_tma(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Fused2Args a)
{
    using S = Smem2<NG, TH, LUTP>;
    constexpr int BH = S::BH, AH = S::AH, NT = NG * NTG;
    constexpr int ITEMS_B = (BH / 2) * RUNS;
    static_assert(BH % 2 == 0, "row pairs");
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid_cta = threadIdx.x, lane = tid_cta & 31;
    const int group = NG == 1 ? 0 : tid_cta / NTG, tid = NG == 1 ? tid_cta : tid_cta - group * NTG;
    uint8_t *gbase = smem + S::gBytes * group;
    uint8_t *sRaw = gbase + S::gRaw;
    uint32_t *sA = reinterpret_cast<uint32_t *>(gbase + S::gA);
    uint32_t *sB = reinterpret_cast<uint32_t *>(gbase + S::gB);
    Axis2 *sAx = reinterpret_cast<Axis2 *>(gbase + S::gX);
    float *sW = reinterpret_cast<float *>(smem + S::offW);
    SmemColorTables *sTab = reinterpret_cast<SmemColorTables *>(smem + S::offT);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::offBar);     // 0: colour tables, 1: folded weights, 2 + g: raw tile of group g
    uint64_t *bar_raw = &bars[2 + group];
    const int H = a.H, W = a.W;
    const int stride = gridDim.x * NG;

    int t = blockIdx.x * NG + group;
    const bool active = t < a.total_tiles;
    TilePos<TH> cur = tile_pos<TH>(active ? t : 0, a);
    if (tid_cta == 0) {
        for (int i = 0; i < 2 + NG; ++i) mbar_init(&bars[i], 1);
        mbar_init_fence();
        bulk_g2s(sTab, a.tabs, (uint32_t)sizeof(SmemColorTables), &bars[0]);
        if (!LUTP) bulk_g2s(sW, a.wlut, 10 * 768 * 4, &bars[1]);
    }
    if (LUTP) {
        // lane-private copies of the range weights: entry i of lane l at word 32 * i + l (bank l)
        const float *col = a.wlut + 10 * 768;
        for (int k = tid_cta; k < 768 * 32; k += NT) sW[k] = __ldg(col + (k >> 5));
    }
    if (active) fill_axis<NTG, AH>(sAx, a, cur.x0, cur.y0, tid);
    __syncthreads();
    if (active && tid == 0) tma_load_tile(sRaw, &tmap, raw_c0(cur.x0), cur.y0 - BY - AR, cur.frame, bar_raw, (uint32_t)S::rawBytes);
    mbar_wait(&bars[0], 0);
    if (!LUTP) mbar_wait(&bars[1], 0);
    const float *myW = sW + lane;

    uint32_t phase = 0;
    for (; t < a.total_tiles; t += stride) {
        const int x0 = cur.x0, y0 = cur.y0, frame = cur.frame;
        const int bx0 = x0 - BX, by0 = y0 - BY, ax0 = bx0 - AR, ay0 = by0 - AR;
        const int tn = t + stride;
        const bool has_next = tn < a.total_tiles;
        TilePos<TH> nxt = cur;
        if (has_next) nxt = tile_pos<TH>(tn, a);

        // ---- the Lab tile of this iteration has landed? ----
        mbar_wait(bar_raw, phase);
        phase ^= 1;
        uint8_t *raw = sRaw + raw_lead(x0);           // staged column 0 of every row starts here (an even offset)
        if (ax0 < 0 || ay0 < 0 || ax0 + AW > W || ay0 + AH > H) {        // group-uniform: tile on an image edge
            // REFLECT_101 ring (6 columns / 5 rows beyond the frame are read at most): copy from the in-image pixel.
            // Only the ring is visited: rows above / below the frame over the whole width, then the side columns.
            const int top_n = max(0, min(AH, -ay0)), bot0 = max(0, min(AH, H - ay0));
            const int left_n = max(0, min(AW, -ax0)), right0 = max(0, min(AW, W - ax0));
            const int ring_rows = top_n + (AH - bot0), ring_cols = left_n + (AW - right0);
            for (int i = tid; i < ring_rows * AW + (AH - ring_rows) * ring_cols; i += NTG) {
                int ly, lx;
                if (i < ring_rows * AW) {
                    const int r = i / AW;
                    lx = i - r * AW; ly = r < top_n ? r : bot0 + (r - top_n);
                } else {
                    const int k2 = i - ring_rows * AW, r = k2 / ring_cols, c = k2 - r * ring_cols;
                    ly = top_n + r; lx = c < left_n ? c : right0 + (c - left_n);
                }
                const int x = ax0 + lx, y = ay0 + ly;
                if (x < -6 || x > W + 5 || y < -5 || y > H + 5) continue;
                const int sx = reflect101(x, W) - ax0, sy = reflect101(y, H) - ay0;
                if (sx < 0 || sx >= AW || sy < 0 || sy >= AH) continue;
                const uint8_t *s = raw + sy * RAW_PITCH + sx * 3;
                uint8_t *d = raw + ly * RAW_PITCH + lx * 3;
                CVB_BOUNDS(ly >= 0 && ly < AH && lx >= 0 && lx < AW && (d - sRaw) + 2 < (int)S::rawBytes && (s - sRaw) + 2 < (int)S::rawBytes);
                d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
            }
            group_sync<NG, NTG>(group);
        }

        // ---- A: lighting of tile + halo, two pixels per item ----
        {
            const uint8_t *lut = a.lut + (size_t)frame * a.g.tiles_x * a.g.tiles_y * 256;
            constexpr int PAIRS = AW / 2;
            for (int item = tid; item < AH * PAIRS; item += NTG) {
                const int ly = item / PAIRS, lx = 2 * (item - ly * PAIRS);
                const uint16_t *rp = reinterpret_cast<const uint16_t *>(raw + ly * RAW_PITCH + lx * 3);
                CVB_BOUNDS(ly < AH && lx + 1 < AW && (raw - sRaw) + ly * RAW_PITCH + lx * 3 + 5 < (int)S::rawBytes);
                const uint32_t u0 = rp[0], u1 = rp[1], u2 = rp[2];                    // L0 a0 | b0 L1 | a1 b1
                const Axis2 cy = sAx[AW + ly], c0 = sAx[lx], c1 = sAx[lx + 1];
                const uint32_t q0 = light_px(sTab, lut, c0, cy, u0 & 0xff, u0 >> 8, u1 & 0xff);
                const uint32_t q1 = light_px(sTab, lut, c1, cy, u1 >> 8, u2 & 0xff, u2 >> 8);
                *reinterpret_cast<uint2 *>(sA + ly * AW + lx) = make_uint2(q0, q1);
            }
        }
        group_sync<NG, NTG>(group);  // sA complete; the raw tile, the axis records and sB (sharpen of the previous tile) are free

        if (has_next) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                tma_load_tile(sRaw, &tmap, raw_c0(nxt.x0), nxt.y0 - BY - AR, nxt.frame, bar_raw, (uint32_t)S::rawBytes);
            }
            fill_axis<NTG, AH>(sAx, a, nxt.x0, nxt.y0, tid);
        }

        // ---- B: bilateral, runs of 4 pixels x 2 rows per item ----
        for (int item = tid; item < ITEMS_B; item += NTG) {
            const int rg = item / RUNS, r4 = (item - rg * RUNS) * 4;
            const int row0 = rg * 2;
            const int Y0 = by0 + row0, X = bx0 + r4;
            if (Y0 + 1 < 0 || Y0 >= H || X + 3 < 0 || X >= W) continue;
            if (ACC == 2) bilateral_item_packed<LUTP>(sA, sB, sW, myW, a, row0, r4);
            else bilateral_item_scalar<LUTP>(sA, sB, sW, myW, a, row0, r4);
        }
        group_sync<NG, NTG>(group);  // sB complete; sA free for the lighting stage of the next tile

        // ---- C: 3x3 sharpen + min/max, 4 pixels per item, stored straight to the frame ----
        {
            const int vh = min(TH, H - y0), vw = min(TW, W - x0);      // valid outputs of this tile (vw % 4 == 0)
            const bool top = y0 == 0, bottom = y0 + vh == H, left = x0 == 0, right = x0 + vw == W;
            if (top || bottom) {
                for (int i = tid; i < BW; i += NTG) {
                    if (top) sB[(BY - 1) * BW + i] = sB[(BY + 1) * BW + i];
                    if (bottom) sB[(BY + vh) * BW + i] = sB[(BY + vh - 2) * BW + i];
                }
                group_sync<NG, NTG>(group);
            }
            if (left || right) {
                for (int i = tid; i < vh + 2; i += NTG) {
                    uint32_t *row = sB + (BY - 1 + i) * BW;
                    if (left) row[BX - 1] = row[BX + 1];
                    if (right) row[BX + vw] = row[BX + vw - 2];
                }
                group_sync<NG, NTG>(group);
            }
            uint8_t *out = a.dst + (size_t)frame * H * W * 3;
            constexpr int GROUPS = TW / 4;
            int vmin = 255, vmax = 0;
            for (int item = tid; item < TH * GROUPS; item += NTG) {
                const int ty = item / GROUPS, tx4 = (item - ty * GROUPS) * 4;
                if (ty >= vh) break;
                if (tx4 >= vw) continue;
                uint32_t cbr[6], cg[6], ctr_w[4];
#pragma unroll
                for (int k = 0; k < 6; ++k) cbr[k] = cg[k] = 0;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const uint32_t *rp = sB + (ty + BY - 1 + r) * BW + tx4;
                    CVB_BOUNDS(ty + BY - 1 + r >= 0 && (ty + BY - 1 + r) * BW + tx4 + 7 < BW * BH && y0 + ty < H && x0 + tx4 + 3 < W);
                    const uint4 q0 = *reinterpret_cast<const uint4 *>(rp), q1 = *reinterpret_cast<const uint4 *>(rp + 4);
                    const uint32_t w[6] = {q0.y, q0.z, q0.w, q1.x, q1.y, q1.z};      // image x - 1 .. x + 4
#pragma unroll
                    for (int k = 0; k < 6; ++k) { cbr[k] += w[k] & 0x00ff00ffu; cg[k] += (w[k] >> 8) & 0xffu; }
                    if (r == 1) { ctr_w[0] = w[1]; ctr_w[1] = w[2]; ctr_w[2] = w[3]; ctr_w[3] = w[4]; }
                }
                uint32_t res[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t s02 = cbr[j] + cbr[j + 1] + cbr[j + 2], s1 = cg[j] + cg[j + 1] + cg[j + 2];
                    const uint32_t c = ctr_w[j];
                    const int b = clamp_u8(10 * (int)(c & 0xff) - (int)(s02 & 0xffff));
                    const int g = clamp_u8(10 * (int)((c >> 8) & 0xff) - (int)s1);
                    const int r = clamp_u8(10 * (int)((c >> 16) & 0xff) - (int)(s02 >> 16));
                    res[j] = pack_bgr(b, g, r);
                    vmin = min(vmin, min(b, min(g, r)));
                    vmax = max(vmax, max(b, max(g, r)));
                }
                uint32_t *o32 = reinterpret_cast<uint32_t *>(out + ((size_t)(y0 + ty) * W + x0 + tx4) * 3);
                o32[0] = (res[0] & 0xffffffu) | (res[1] << 24);
                o32[1] = ((res[1] >> 8) & 0xffffu) | (res[2] << 16);
                o32[2] = ((res[2] >> 16) & 0xffu) | (res[3] << 8);
            }
            if (a.minmax) {
                vmin = warp_min(vmin); vmax = warp_max(vmax);
                if (lane == 0 && vmin <= vmax) {
                    atomicMin(a.minmax + 2 * frame, vmin);
                    atomicMax(a.minmax + 2 * frame + 1, vmax);
                }
            }
        }
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Producer / consumer form: NP lighting threads run one tile ahead of NQ bilateral + sharpen threads.
//   producers: wait for the Lab tile (TMA) -> patch the mirror ring -> lighting into sA[k & 1] -> signal `full`
//              -> start the TMA load of their next tile;
//   consumers: wait `full` -> bilateral from sA[k & 1] into sB -> signal `empty` -> sharpen + store + min/max.
// The two stages are bound by different pipes (lighting: integer pipe and global-load latency; bilateral: FMA and
// shared-memory pipes), so running them side by side on one SM lets their pipe demands overlap the way two resident
// CTAs of k_fused do -- with ONE lane-private weight table instead of two conflict-ridden folded ones.
// ---------------------------------------------------------------------------------------------------------------
template <int TH, bool LUTP>
struct SmemPC {
    static constexpr int BH = TH + 2 * BY, AH = BH + 2 * AR;
    static constexpr size_t rawBytes = (size_t)RAW_PITCH * AH;
    static constexpr size_t offRaw = 0;
    static constexpr size_t offA0 = (rawBytes + 127) / 128 * 128;
    static constexpr size_t aBytes = (size_t)AW * AH * 4;
    static constexpr size_t offB = offA0 + 2 * aBytes;
    static constexpr size_t offX = offB + (size_t)BW * BH * 4;
    static constexpr size_t offW = (offX + (size_t)(AW + AH) * sizeof(Axis2) + 15) / 16 * 16;
    static constexpr size_t offT = offW + (LUTP ? 768 * 32 * 4 : 10 * 768 * 4);
    static constexpr size_t offBar = offT + sizeof(SmemColorTables);
    static constexpr size_t bytes = offBar + 8 * sizeof(uint64_t);
    static_assert(offA0 % 16 == 0 && aBytes % 16 == 0 && offB % 16 == 0 && offX % 16 == 0 && offT % 16 == 0 && offBar % 8 == 0, "alignment");
    static_assert(bytes <= 232448, "more than 227 KB of shared memory");
};

// wait that does not burn issue slots: the spinning warps of the producer / consumer kernel otherwise execute 10 % of
// all instructions (ncu: BRA / SYNCS / YIELD), on a kernel that is bound by instruction issue
CVB_DEV void mbar_wait_sleep(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITS_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x400;\n"
        "@p bra DONES_%=;\n"
        "nanosleep.u32 256;\n"
        "bra WAITS_%=;\n"
        "DONES_%=:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(phase)
        : "memory");
}
CVB_DEV void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
template <int COUNT>
CVB_DEV void named_sync(int id)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(COUNT) : "memory");
}

template <int TH, int NP, int NQ, bool LUTP, int ACC>
__global__ void __launch_bounds__(NP + NQ, 1) k_fused_pc(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Fused2Args a)
{
    using S = SmemPC<TH, LUTP>;
    constexpr int BH = S::BH, AH = S::AH, NT = NP + NQ;
    constexpr int ITEMS_B = (BH / 2) * RUNS;
    static_assert(BH % 2 == 0 && NP % 32 == 0 && NQ % 32 == 0, "whole warps, row pairs");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *sRaw = smem + S::offRaw;
    uint32_t *sB = reinterpret_cast<uint32_t *>(smem + S::offB);
    Axis2 *sAx = reinterpret_cast<Axis2 *>(smem + S::offX);
    float *sW = reinterpret_cast<float *>(smem + S::offW);
    SmemColorTables *sTab = reinterpret_cast<SmemColorTables *>(smem + S::offT);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + S::offBar);   // 0 tables, 1 folded weights, 2 raw tile, 3-4 full, 5-6 empty
    uint64_t *bar_raw = &bars[2], *bar_full = &bars[3], *bar_empty = &bars[5];
    const int tid = threadIdx.x, lane = tid & 31;
    const int H = a.H, W = a.W;
    const bool producer = tid < NP;

    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(bar_raw, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&bar_full[b], NP / 32); mbar_init(&bar_empty[b], NQ / 32); }
        mbar_init_fence();
        bulk_g2s(sTab, a.tabs, (uint32_t)sizeof(SmemColorTables), &bars[0]);
        if (!LUTP) bulk_g2s(sW, a.wlut, 10 * 768 * 4, &bars[1]);
    }
    if (LUTP) {
        const float *col = a.wlut + 10 * 768;
        for (int k = tid; k < 768 * 32; k += NT) sW[k] = __ldg(col + (k >> 5));
    }
    __syncthreads();
    mbar_wait(&bars[0], 0);
    if (!LUTP) mbar_wait(&bars[1], 0);

    if (producer) {
        // ------------------------------------------------ lighting warps ------------------------------------------------
        int t = blockIdx.x;
        if (t >= a.total_tiles) return;
        TilePos<TH> cur = tile_pos<TH>(t, a);
        fill_axis<NP, AH>(sAx, a, cur.x0, cur.y0, tid);
        if (tid == 0) tma_load_tile(sRaw, &tmap, raw_c0(cur.x0), cur.y0 - BY - AR, cur.frame, bar_raw, (uint32_t)S::rawBytes);
        named_sync<NP>(1);
        for (int k = 0; t < a.total_tiles; t += gridDim.x, ++k) {
            const int x0 = cur.x0, y0 = cur.y0, frame = cur.frame;
            const int ax0 = x0 - BX - AR, ay0 = y0 - BY - AR;
            const int tn = t + gridDim.x;
            const bool has_next = tn < a.total_tiles;
            TilePos<TH> nxt = cur;
            if (has_next) nxt = tile_pos<TH>(tn, a);
            uint32_t *sA = reinterpret_cast<uint32_t *>(smem + S::offA0 + (size_t)(k & 1) * S::aBytes);
            mbar_wait_sleep(bar_raw, k & 1);
            uint8_t *raw = sRaw + raw_lead(x0);
            if (ax0 < 0 || ay0 < 0 || ax0 + AW > W || ay0 + AH > H) {
                const int top_n = max(0, min(AH, -ay0)), bot0 = max(0, min(AH, H - ay0));
                const int left_n = max(0, min(AW, -ax0)), right0 = max(0, min(AW, W - ax0));
                const int ring_rows = top_n + (AH - bot0), ring_cols = left_n + (AW - right0);
                for (int i = tid; i < ring_rows * AW + (AH - ring_rows) * ring_cols; i += NP) {
                    int ly, lx;
                    if (i < ring_rows * AW) {
                        const int r = i / AW;
                        lx = i - r * AW; ly = r < top_n ? r : bot0 + (r - top_n);
                    } else {
                        const int k2 = i - ring_rows * AW, r = k2 / ring_cols, c = k2 - r * ring_cols;
                        ly = top_n + r; lx = c < left_n ? c : right0 + (c - left_n);
                    }
                    const int x = ax0 + lx, y = ay0 + ly;
                    if (x < -6 || x > W + 5 || y < -5 || y > H + 5) continue;
                    const int sx = reflect101(x, W) - ax0, sy = reflect101(y, H) - ay0;
                    if (sx < 0 || sx >= AW || sy < 0 || sy >= AH) continue;
                    const uint8_t *s = raw + sy * RAW_PITCH + sx * 3;
                    uint8_t *d = raw + ly * RAW_PITCH + lx * 3;
                    CVB_BOUNDS(ly >= 0 && ly < AH && lx >= 0 && lx < AW && (d - sRaw) + 2 < (int)S::rawBytes && (s - sRaw) + 2 < (int)S::rawBytes);
                    d[0] = s[0]; d[1] = s[1]; d[2] = s[2];
                }
                named_sync<NP>(1);
            }
            mbar_wait_sleep(&bar_empty[k & 1], ((k >> 1) & 1) ^ 1);        // the consumers are done with this buffer's previous tile
            {
                const uint8_t *lut = a.lut + (size_t)frame * a.g.tiles_x * a.g.tiles_y * 256;
                constexpr int PAIRS = AW / 2;
                for (int item = tid; item < AH * PAIRS; item += NP) {
                    const int ly = item / PAIRS, lx = 2 * (item - ly * PAIRS);
                    const uint16_t *rp = reinterpret_cast<const uint16_t *>(raw + ly * RAW_PITCH + lx * 3);
                    CVB_BOUNDS(ly < AH && lx + 1 < AW && (raw - sRaw) + ly * RAW_PITCH + lx * 3 + 5 < (int)S::rawBytes);
                    const uint32_t u0 = rp[0], u1 = rp[1], u2 = rp[2];
                    const Axis2 cy = sAx[AW + ly], c0 = sAx[lx], c1 = sAx[lx + 1];
                    const uint32_t q0 = light_px(sTab, lut, c0, cy, u0 & 0xff, u0 >> 8, u1 & 0xff);
                    const uint32_t q1 = light_px(sTab, lut, c1, cy, u1 >> 8, u2 & 0xff, u2 >> 8);
                    *reinterpret_cast<uint2 *>(sA + ly * AW + lx) = make_uint2(q0, q1);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[k & 1]);
            named_sync<NP>(1);                       // every producer is done with the raw tile and the axis records
            if (has_next) {
                if (tid == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    tma_load_tile(sRaw, &tmap, raw_c0(nxt.x0), nxt.y0 - BY - AR, nxt.frame, bar_raw, (uint32_t)S::rawBytes);
                }
                fill_axis<NP, AH>(sAx, a, nxt.x0, nxt.y0, tid);
                named_sync<NP>(1);
            }
            cur = nxt;
        }
    } else {
        // ------------------------------------------- bilateral + sharpen warps -------------------------------------------
        const int ctid = tid - NP;
        const float *myW = sW + lane;
        int t = blockIdx.x;
        for (int k = 0; t < a.total_tiles; t += gridDim.x, ++k) {
            const TilePos<TH> cur = tile_pos<TH>(t, a);
            const int x0 = cur.x0, y0 = cur.y0, frame = cur.frame;
            const int bx0 = x0 - BX, by0 = y0 - BY;
            const uint32_t *sA = reinterpret_cast<const uint32_t *>(smem + S::offA0 + (size_t)(k & 1) * S::aBytes);
            mbar_wait_sleep(&bar_full[k & 1], (k >> 1) & 1);
            for (int item = ctid; item < ITEMS_B; item += NQ) {
                const int rg = item / RUNS, r4 = (item - rg * RUNS) * 4;
                const int row0 = rg * 2;
                const int Y0 = by0 + row0, X = bx0 + r4;
                if (Y0 + 1 < 0 || Y0 >= H || X + 3 < 0 || X >= W) continue;
                if (ACC == 2) bilateral_item_packed<LUTP>(sA, sB, sW, myW, a, row0, r4);
                else bilateral_item_scalar<LUTP>(sA, sB, sW, myW, a, row0, r4);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[k & 1]);     // this warp no longer reads the lighting tile
            named_sync<NQ>(2);                                 // sB complete
            {
                const int vh = min(TH, H - y0), vw = min(TW, W - x0);
                const bool top = y0 == 0, bottom = y0 + vh == H, left = x0 == 0, right = x0 + vw == W;
                if (top || bottom) {
                    for (int i = ctid; i < BW; i += NQ) {
                        if (top) sB[(BY - 1) * BW + i] = sB[(BY + 1) * BW + i];
                        if (bottom) sB[(BY + vh) * BW + i] = sB[(BY + vh - 2) * BW + i];
                    }
                    named_sync<NQ>(2);
                }
                if (left || right) {
                    for (int i = ctid; i < vh + 2; i += NQ) {
                        uint32_t *row = sB + (BY - 1 + i) * BW;
                        if (left) row[BX - 1] = row[BX + 1];
                        if (right) row[BX + vw] = row[BX + vw - 2];
                    }
                    named_sync<NQ>(2);
                }
                uint8_t *out = a.dst + (size_t)frame * H * W * 3;
                constexpr int GROUPS = TW / 4;
                int vmin = 255, vmax = 0;
                for (int item = ctid; item < TH * GROUPS; item += NQ) {
                    const int ty = item / GROUPS, tx4 = (item - ty * GROUPS) * 4;
                    if (ty >= vh) break;
                    if (tx4 >= vw) continue;
                    uint32_t cbr[6], cg[6], ctr_w[4];
#pragma unroll
                    for (int q = 0; q < 6; ++q) cbr[q] = cg[q] = 0;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const uint32_t *rp = sB + (ty + BY - 1 + r) * BW + tx4;
                        const uint4 q0 = *reinterpret_cast<const uint4 *>(rp), q1 = *reinterpret_cast<const uint4 *>(rp + 4);
                        const uint32_t w[6] = {q0.y, q0.z, q0.w, q1.x, q1.y, q1.z};
#pragma unroll
                        for (int q = 0; q < 6; ++q) { cbr[q] += w[q] & 0x00ff00ffu; cg[q] += (w[q] >> 8) & 0xffu; }
                        if (r == 1) { ctr_w[0] = w[1]; ctr_w[1] = w[2]; ctr_w[2] = w[3]; ctr_w[3] = w[4]; }
                    }
                    uint32_t res[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t s02 = cbr[j] + cbr[j + 1] + cbr[j + 2], s1 = cg[j] + cg[j + 1] + cg[j + 2];
                        const uint32_t c = ctr_w[j];
                        const int b = clamp_u8(10 * (int)(c & 0xff) - (int)(s02 & 0xffff));
                        const int g = clamp_u8(10 * (int)((c >> 8) & 0xff) - (int)s1);
                        const int r = clamp_u8(10 * (int)((c >> 16) & 0xff) - (int)(s02 >> 16));
                        res[j] = pack_bgr(b, g, r);
                        vmin = min(vmin, min(b, min(g, r)));
                        vmax = max(vmax, max(b, max(g, r)));
                    }
                    uint32_t *o32 = reinterpret_cast<uint32_t *>(out + ((size_t)(y0 + ty) * W + x0 + tx4) * 3);
                    o32[0] = (res[0] & 0xffffffu) | (res[1] << 24);
                    o32[1] = ((res[1] >> 8) & 0xffffu) | (res[2] << 16);
                    o32[2] = ((res[2] >> 16) & 0xffu) | (res[3] << 8);
                }
                if (a.minmax) {
                    vmin = warp_min(vmin); vmax = warp_max(vmax);
                    if (lane == 0 && vmin <= vmax) {
                        atomicMin(a.minmax + 2 * frame, vmin);
                        atomicMax(a.minmax + 2 * frame + 1, vmax);
                    }
                }
            }
            named_sync<NQ>(2);                                 // sB free for the bilateral stage of the next tile
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_lab_tensor_map(const uint8_t *lab, int n, int H, int W, int box_rows, CUtensorMap *map)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        CVB_CHECK_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        CVB_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        encode = (EncodeTiledFn)fn;
    }
    // the Lab frames as u16 elements: [n][H][W * 3 / 2]
    const cuuint64_t dims[3] = {(cuuint64_t)W * 3 / 2, (cuuint64_t)H, (cuuint64_t)n};
    const cuuint64_t strides[2] = {(cuuint64_t)W * 3, (cuuint64_t)H * W * 3};
    const cuuint32_t box[3] = {RAW_PITCH / 2, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<uint8_t *>(lab), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CVB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (Lab frames %d x %d x %d)", (int)r, n, H, W);
    return CVB_OK;
}

template <int NG, int NTG, int TH, bool LUTP, int ACC>
int launch_variant(cvb_handle *h, const uint8_t *lab, Fused2Args &a)
{
    using S = Smem2<NG, TH, LUTP>;
    auto kern = k_fused_tma<NG, NTG, TH, LUTP, ACC>;
    if (!h->fused_attr_done.count((const void *)kern)) {
        CVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
        h->fused_attr_done.insert((const void *)kern);
    }
    a.tiles_x = (a.W + TW - 1) / TW; a.tiles_y = (a.H + TH - 1) / TH;
    const long total = (long)a.tiles_x * a.tiles_y * a.n;
    CVB_REQUIRE(total < (1l << 31), "too many tiles");
    a.total_tiles = (int)total;
    alignas(64) CUtensorMap map;
    CVB_TRY(make_lab_tensor_map(lab, a.n, a.H, a.W, S::AH, &map));
    const int grid = (int)std::min<long>((total + NG - 1) / NG, h->sm_count);
    PROF(h, "k_fused");
    kern<<<grid, NG * NTG, S::bytes, h->stream>>>(map, a);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

template <int TH, int NP, int NQ, bool LUTP, int ACC>
int launch_pc(cvb_handle *h, const uint8_t *lab, Fused2Args &a)
{
    using S = SmemPC<TH, LUTP>;
    auto kern = k_fused_pc<TH, NP, NQ, LUTP, ACC>;
    if (!h->fused_attr_done.count((const void *)kern)) {
        CVB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
        h->fused_attr_done.insert((const void *)kern);
    }
    a.tiles_x = (a.W + TW - 1) / TW; a.tiles_y = (a.H + TH - 1) / TH;
    const long total = (long)a.tiles_x * a.tiles_y * a.n;
    CVB_REQUIRE(total < (1l << 31), "too many tiles");
    a.total_tiles = (int)total;
    alignas(64) CUtensorMap map;
    CVB_TRY(make_lab_tensor_map(lab, a.n, a.H, a.W, S::AH, &map));
    const int grid = (int)std::min<long>(total, h->sm_count);
    PROF(h, "k_fused");
    kern<<<grid, NP + NQ, S::bytes, h->stream>>>(map, a);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

}  // namespace

CVB_BOUNDS_TU(fused2)

bool fused_tma_applicable(int H, int W, const uint8_t *lab, const uint8_t *out)
{
    return W % 16 == 0 && W >= 16 && H >= 16 && ((reinterpret_cast<uintptr_t>(lab) & 15) == 0) &&
           ((reinterpret_cast<uintptr_t>(out) & 3) == 0) && (size_t)H * W * 3 < (1ull << 32);
}

// variant (CVB_FUSED): the experiments of profiles/r02_notes.md; 0 is the default
int launch_fused_tma(cvb_handle *h, const uint8_t *lab, int n, int H, int W, const ClaheGeom &g, const uint8_t *lut,
                     const float *d_wlut, const float *space81, uint8_t *out, int32_t *minmax, int variant)
{
    Fused2Args a;
    memset(&a, 0, sizeof a);
    a.dst = out; a.H = H; a.W = W; a.n = n; a.tabs = h->d_tables; a.lut = lut; a.g = g; a.wlut = d_wlut; a.minmax = minmax;
    memcpy(a.sw81, space81, sizeof a.sw81);
    for (int dy = -4; dy <= 4; ++dy) {
        const int nt = row_taps(dy);
        for (int p = 0; 2 * p + 1 < nt; ++p) {
            a.swp[dy + 4][p].x = space81[(dy + 4) * 9 + tap_dx(dy, 2 * p) + 4];
            a.swp[dy + 4][p].y = space81[(dy + 4) * 9 + tap_dx(dy, 2 * p + 1) + 4];
        }
        a.sws[dy + 4] = (nt & 1) ? space81[(dy + 4) * 9 + tap_dx(dy, nt - 1) + 4] : 0.f;
    }
    // the variants that stay built (every one bit-identical to the oracle; times in profiles/r02_notes.md, where the
    // results of the variants that were measured and dropped are kept too)
    switch (variant) {
    case 1: return launch_variant<1, 1024, 64, false, 2>(h, lab, a);   // one group of 1024, folded table, packed accumulation
    case 2: return launch_variant<1, 512, 64, true, 2>(h, lab, a);     // one group of 512, private table, packed
    case 3: return launch_variant<2, 512, 30, true, 0>(h, lab, a);     // two groups of 512 on 120 x 30 tiles, private table, scalar
    case 4: return launch_variant<1, 512, 64, true, 0>(h, lab, a);     // one group of 512, private table, scalar
    case 5: return launch_pc<64, 256, 512, false, 2>(h, lab, a);       // producer / consumer, 8 + 16 warps of 80 registers, 120 x 64 tiles, product table, packed
    default: return launch_pc<46, 256, 768, true, 0>(h, lab, a);       // producer / consumer warps, 120 x 46 tiles, private table
    }
}
