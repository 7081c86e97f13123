// cv2.HoughCircles(gray, HOUGH_GRADIENT, ...) for a batch of small squares -- the step after the
// per-square statistics in PieceDetector._detect_circle_unified (piece_detector.py:210-270),
// SURVEY.md 8f rank 1.  One CTA owns one square; the whole transform lives in shared memory (squares up to
// CVB_HOUGH_MAX_DIM a side) or, for larger squares, in a global-memory slice of the same layout.
//
// OpenCV imgproc/src/hough.cpp (HoughCirclesGradient), restated (oracle: orc_hough_circles):
//   A  Sobel 3x3 (replicate border), Canny(dx, dy, max(1, param1/2), param1) with L1 magnitude
//   B  every edge pixel with a non-zero gradient walks a Q10 ray along +-gradient for the radii
//      minRadius..maxRadius through an accumulator of 1/dp resolution; a ray stops at its first
//      step outside (one thread per ray, shared-memory atomics)
//   C  centres: cells (not in accumulator row 0 / column 0) > param2, > left, >= right, > up, >= down
//   D  per centre (one warp each): distances to the edge pixels, 10 bins per dp, best 10-bin
//      window scanning downwards -> (radius, support); kept when support > param2
//   E  OpenCV sorts by (support desc, radius desc, x asc, y asc) and greedily drops circles closer
//      than minDist to a kept one == repeatedly take the best live candidate and kill its
//      neighbourhood, which needs no sort.
// All float steps are single IEEE operations (the library is built with -fmad=false).
#include "cvb_device.cuh"
#include <cmath>
#include <cstring>
#include <algorithm>

static_assert(sizeof(cvb_hough_params) == 56 && sizeof(cvb_hough_square) == 40 && sizeof(cvb_hough_result) == 272,
              "ABI struct layout (see _lib.py)");

namespace {

constexpr int HNT = 256;             // threads per square
constexpr int HW_ = HNT / 32;        // warps

// Byte offsets into dynamic shared memory, sized for the largest square of the call.  Two regions are reused:
//   [off_acc]    the vote accumulator (phases B, C), then the per-warp radius histograms (phase D)
//   [off_union]  gray | Canny map | ray steps (phases A, B), then the candidate list (phases C-E)
struct HoughLayout {
    int off_acc, off_nz, off_union, total;
    int bins_pitch;                  // ints per warp: the bins, then one mask bit per bin
    int bins_words_at;               // where the mask words start inside a warp's slice
    int dir_cap;                     // ray-step entries that fit behind the early view of the union
};

struct HoughMisc {
    int nnz, ncent, kept, pad;
    unsigned long long best[HW_];
};

CVB_DEV void sobel_at(const uint8_t *g, int gp, int y, int x, int &gx, int &gy)
{
    // g has a one-pixel replicated border: pixel (y, x) sits at g[(y + 1) * gp + x + 1]
    const uint8_t *p = g + y * gp + x;
    const int a = p[0], b = p[1], c = p[2], d = p[gp], f = p[gp + 2], q = p[2 * gp], r = p[2 * gp + 1], s = p[2 * gp + 2];
    gx = (c + 2 * f + s) - (a + 2 * d + q);
    gy = (q + 2 * r + s) - (a + 2 * b + c);
}

// GWS: squares whose workspace does not fit shared memory (more than CVB_HOUGH_MAX_DIM pixels a side) run the very same
// phases on a slice of global memory per CTA -- slower votes, identical circles.
template <bool GWS>
__global__ void __launch_bounds__(HNT, 4) k_hough(const uint8_t *__restrict__ planes, size_t plane_stride, int PW,
                                              const cvb_hough_square *__restrict__ squares, int n_sq,
                                              const uint8_t *__restrict__ select, float dp, float idp, int canny_low,
                                              int canny_high, int acc_thr, HoughLayout L,
                                              cvb_hough_result *__restrict__ out, uint8_t *gws)
{
    extern __shared__ __align__(16) uint8_t smem_dyn[];
    uint8_t *smem = GWS ? gws + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)L.total : smem_dyn;
    __shared__ HoughMisc M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int sq = blockIdx.x, frame = blockIdx.y;
    cvb_hough_result *res = out + (size_t)frame * n_sq + sq;
    if (select && !select[(size_t)frame * n_sq + sq]) {
        for (int i = tid; i < (int)(sizeof(cvb_hough_result) / 4); i += HNT) reinterpret_cast<int32_t *>(res)[i] = 0;
        if (tid == 0) res->status = CVB_HOUGH_SKIPPED;
        return;
    }
    const cvb_hough_square S = squares[sq];
    const int w = S.w, h = S.h, gp = w + 2, npad = (h + 2) * gp;
    const int arows = S.acc_rows, acols = S.acc_cols, astep = acols + 2, ncells = (arows + 2) * astep;
    const unsigned inv_gp = magic_of(gp), inv_astep = magic_of(astep), inv_acols = magic_of(acols);
    const int min_r = S.min_radius, max_r = S.max_radius, nbins = S.n_bins;

    int32_t *acc = reinterpret_cast<int32_t *>(smem + L.off_acc);
    uint16_t *nz = reinterpret_cast<uint16_t *>(smem + L.off_nz);
    int *bins = reinterpret_cast<int *>(smem + L.off_acc) + warp * L.bins_pitch;  // phase D only: the accumulator is dead then
    unsigned *bmask = reinterpret_cast<unsigned *>(bins + L.bins_words_at);     // one bit per bin, after the bins
    // early view of the union: gray (u8) | map (u8) | ray steps (2 x i16)
    const int npad4 = (npad + 3) & ~3;
    uint8_t *s_g = smem + L.off_union;
    uint8_t *s_map = smem + L.off_union + npad4;
    int *s_dir = reinterpret_cast<int *>(smem + L.off_union + 2 * npad4);
    // late view: centre cells (u16) | radius (f32) | support (u16)
    const int maxc = ncells / 2 + 1;
    float *c_r = reinterpret_cast<float *>(smem + L.off_union);
    uint16_t *c_idx = reinterpret_cast<uint16_t *>(smem + L.off_union + 4 * maxc);
    uint16_t *c_sup = c_idx + maxc;

    if (tid == 0) { M.nnz = 0; M.ncent = 0; M.kept = 0; M.pad = 0; }
    for (int i = tid; i < ncells; i += HNT) acc[i] = 0;
    // ---- A: gray with a replicated border; Canny ----
    const uint8_t *img = planes + (size_t)frame * plane_stride + (size_t)S.y * PW + S.x;
    for (int i = tid; i < npad; i += HNT) {
        const int ly = div_magic(i, inv_gp), lx = i - ly * gp;
        const int y = min(max(ly - 1, 0), h - 1), x = min(max(lx - 1, 0), w - 1);
        s_g[i] = __ldg(img + (size_t)y * PW + x);
    }
    __syncthreads();
    // L1 gradient magnitude of pixel (y, x); zero outside the image, as OpenCV's padded magnitude rows
    auto mag_at = [&](int y, int x) {
        if ((unsigned)y >= (unsigned)h || (unsigned)x >= (unsigned)w) return 0;
        int gx, gy;
        sobel_at(s_g, gp, y, x, gx, gy);
        return abs(gx) + abs(gy);
    };
    // non-maximum suppression; map: 0 weak candidate, 1 no edge, 2 edge.  Few pixels pass the low threshold, so the
    // two neighbour magnitudes along the gradient are recomputed there instead of keeping a magnitude array.
    for (int i = tid; i < npad; i += HNT) {
        const int ly = div_magic(i, inv_gp), lx = i - ly * gp;
        uint8_t v = 1;
        if (ly >= 1 && ly <= h && lx >= 1 && lx <= w) {
            const int y = ly - 1, x = lx - 1;
            int xs, ys;
            sobel_at(s_g, gp, y, x, xs, ys);
            const int m = abs(xs) + abs(ys);
            if (m > canny_low) {
                const int ax = abs(xs), ay = abs(ys) << 15;            // < 2^26
                const int tg22x = ax * 13573, tg67x = tg22x + (ax << 16);   // < 2^27
                bool cand;
                if (ay < tg22x) cand = m > mag_at(y, x - 1) && m >= mag_at(y, x + 1);
                else if (ay > tg67x) cand = m > mag_at(y - 1, x) && m >= mag_at(y + 1, x);
                else {
                    const int sgn = (xs ^ ys) < 0 ? -1 : 1;
                    cand = m > mag_at(y - 1, x - sgn) && m > mag_at(y + 1, x + sgn);
                }
                if (cand) v = m > canny_high ? 2 : 0;
            }
        }
        s_map[i] = v;
    }
    __syncthreads();
    // hysteresis: grow the strong set through weak candidates until nothing changes.  Only the weak candidates can
    // change, so they are listed once (in the edge-list buffer, not yet in use) and the sweeps walk that list.
    for (int i0 = 0; i0 < npad; i0 += HNT) {
        const int i = i0 + tid;
        const bool weak = i < npad && s_map[i] == 0;
        const unsigned bal = __ballot_sync(0xffffffffu, weak);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&M.pad, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (weak) nz[base + __popc(bal & ((1u << lane) - 1))] = (uint16_t)i;
    }
    __syncthreads();
    const int nweak = M.pad;
    for (;;) {
        bool mine = false;
        for (int e = tid; e < nweak; e += HNT) {
            const int i = nz[e];
            if (s_map[i] == 0) {
                const bool hit = s_map[i - gp - 1] == 2 || s_map[i - gp] == 2 || s_map[i - gp + 1] == 2 || s_map[i - 1] == 2 ||
                                 s_map[i + 1] == 2 || s_map[i + gp - 1] == 2 || s_map[i + gp] == 2 || s_map[i + gp + 1] == 2;
                if (hit) { s_map[i] = 2; mine = true; }        // monotone 0 -> 2
            }
        }
        if (!__syncthreads_or(mine)) break;
    }
    // ---- B: edge pixels with a gradient -> list; rays -> accumulator ----
    for (int i0 = 0; i0 < npad; i0 += HNT) {
        const int i = i0 + tid;
        bool take = false;
        int ly = 0, lx = 0;
        if (i < npad && s_map[i] == 2) {
            ly = div_magic(i, inv_gp); lx = i - ly * gp;
            int gx, gy;
            sobel_at(s_g, gp, ly - 1, lx - 1, gx, gy);
            take = (gx | gy) != 0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&M.nnz, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (take) nz[base + __popc(bal & ((1u << lane) - 1))] = (uint16_t)((lx - 1) | ((ly - 1) << 8));
    }
    __syncthreads();
    const int nnz = M.nnz;
    // The unit steps (sx, sy) of a pixel's two rays are computed once, by one thread per edge pixel, when the
    // list fits the buffer behind the Canny map; a square with more edge pixels recomputes them per ray.
    const bool dir_stored = nnz <= L.dir_cap;
    auto ray_step = [&](int x, int y, int &sx, int &sy) {
        int gx, gy;
        sobel_at(s_g, gp, y, x, gx, gy);
        const float vx = (float)gx, vy = (float)gy;
        const float mag = __fsqrt_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)));
        sx = __float2int_rn(__fdiv_rn(__fmul_rn(__fmul_rn(vx, idp), 1024.f), mag));
        sy = __float2int_rn(__fdiv_rn(__fmul_rn(__fmul_rn(vy, idp), 1024.f), mag));
    };
    if (dir_stored) {
        for (int e = tid; e < nnz; e += HNT) {
            const int p = nz[e];
            int sx, sy;
            ray_step(p & 255, p >> 8, sx, sy);
            s_dir[e] = (sx & 0xffff) | (sy << 16);
        }
        __syncthreads();
    }
    // one thread per ray (pixel, direction), walking it until its first step outside the accumulator
    for (int ray = tid; ray < 2 * nnz; ray += HNT) {
        const int p = nz[ray >> 1], x = p & 255, y = p >> 8;
        int sx, sy;
        if (dir_stored) {
            const int d = s_dir[ray >> 1];
            sx = (int)(short)(d & 0xffff); sy = d >> 16;
        } else {
            ray_step(x, y, sx, sy);
        }
        if (ray & 1) { sx = -sx; sy = -sy; }
        int x1 = __float2int_rn(__fmul_rn(__fmul_rn((float)x, idp), 1024.f)) + min_r * sx;
        int y1 = __float2int_rn(__fmul_rn(__fmul_rn((float)y, idp), 1024.f)) + min_r * sy;
        for (int r = min_r; r <= max_r; ++r, x1 += sx, y1 += sy) {
            const int x2 = x1 >> 10, y2 = y1 >> 10;
            if ((unsigned)x2 >= (unsigned)acols || (unsigned)y2 >= (unsigned)arows) break;
            atomicAdd(&acc[y2 * astep + x2], 1);
        }
    }
    __syncthreads();      // gray / mag / map are dead from here on: the union switches to its late view
    // ---- C: centres ----
    if (nnz > 0 && nbins > 0) {
        const int cw = acols, ch = arows;       // candidate cells: rows 1..arows, columns 1..acols of the padded array
        for (int i0 = 0; i0 < cw * ch; i0 += HNT) {
            const int i = i0 + tid;
            bool take = false;
            int cell = 0;
            if (i < cw * ch) {
                const int ay = div_magic(i, inv_acols) + 1, ax = i - (ay - 1) * cw + 1;
                cell = ay * astep + ax;
                const int v = acc[cell];
                take = v > acc_thr && v > acc[cell - 1] && v >= acc[cell + 1] && v > acc[cell - astep] && v >= acc[cell + astep];
            }
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&M.ncent, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (take) c_idx[base + __popc(bal & ((1u << lane) - 1))] = (uint16_t)cell;
        }
    }
    __syncthreads();
    const int ncent = M.ncent;
    // ---- D: radius per centre, one warp each ----
    const float dr = dp, fmin_r = (float)min_r;
    const float minr2 = __fmul_rn(fmin_r, fmin_r), maxr2 = __fmul_rn((float)max_r, (float)max_r);
    for (int c = warp; c < ncent; c += HW_) {
        const int cell = c_idx[c], ay = div_magic(cell, inv_astep), ax = cell - ay * astep;
        const float cx = __fmul_rn((float)ax + 0.5f, dr), cy = __fmul_rn((float)ay + 0.5f, dr);
        for (int i = lane; i < nbins; i += 32) bins[i] = 0;
        __syncwarp();
        for (int k = lane; k < nnz; k += 32) {
            const int p = nz[k];
            const float ddx = __fsub_rn(cx, (float)(p & 255)), ddy = __fsub_rn(cy, (float)(p >> 8));
            const float r2 = __fadd_rn(__fmul_rn(ddx, ddx), __fmul_rn(ddy, ddy));
            if (minr2 <= r2 && r2 <= maxr2) {
                const float t = __fmul_rn(__fdiv_rn(__fsub_rn(__fsqrt_rn(r2), fmin_r), dr), 10.f);
                atomicAdd(&bins[min(max(__float2int_rn(t), 0), nbins - 1)], 1);
            }
        }
        __syncwarp();
        // hough.cpp scans the bins downwards: the next non-empty bin j > 0 opens a window of 10 bins below it, then
        // the scan resumes 11 bins lower.  A bit mask of the non-empty bins finds the window tops without the walk.
        for (int c0 = 0; c0 < nbins; c0 += 32) {
            const unsigned m = __ballot_sync(0xffffffffu, c0 + lane < nbins && bins[c0 + lane] != 0);
            if (lane == 0) bmask[c0 >> 5] = m;
        }
        __syncwarp();
        if (lane == 0) {
            int max_count = 0;
            float rbest = 0.f;
            int j = nbins - 1;
            while (j > 0) {
                int wd = j >> 5;
                unsigned m = bmask[wd] & (0xffffffffu >> (31 - (j & 31)));
                while (!m && wd > 0) m = bmask[--wd];
                if (!m) break;
                const int up = wd * 32 + 31 - __clz(m);
                if (up <= 0) break;
                int cur = 0;
                const int lo = max(up - 9, 0);
                for (int b = up; b >= lo; --b) cur += bins[b];
                j = max(up - 10, -1);
                const float rcur = __fadd_rn(__fmul_rn(__fdiv_rn(__fdiv_rn((float)(up + j), 2.f), 10.f), dr), fmin_r);
                if (__fmul_rn((float)cur, rbest) >= __fmul_rn((float)max_count, rcur) ||
                    (rbest < 1.1920929e-07f && cur >= max_count)) {
                    rbest = rcur; max_count = cur;
                }
                j--;
            }
            c_r[c] = rbest;
            c_sup[c] = (uint16_t)(max_count > acc_thr ? max_count : 0);
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- E: best live candidate first, then kill everything closer than minDist ----
    const float md2 = __fmul_rn(S.min_dist, S.min_dist);
    for (;;) {
        unsigned long long best = 0;
        for (int c = tid; c < ncent; c += HNT) {
            const unsigned sup = c_sup[c];
            if (sup) {
                const int cell = c_idx[c], ay = div_magic(cell, inv_astep), ax = cell - ay * astep;
                const unsigned long long key = ((unsigned long long)sup << 48) | ((unsigned long long)__float_as_uint(c_r[c]) << 16) |
                                               ((unsigned long long)(255 - ax) << 8) | (unsigned long long)(255 - ay);
                best = key > best ? key : best;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
        }
        if (lane == 0) M.best[warp] = best;
        __syncthreads();
        best = M.best[0];
#pragma unroll
        for (int i = 1; i < HW_; ++i) best = M.best[i] > best ? M.best[i] : best;
        if (best == 0) break;
        const int bx = 255 - (int)((best >> 8) & 255), by = 255 - (int)(best & 255);
        const float kx = __fmul_rn((float)bx + 0.5f, dr), ky = __fmul_rn((float)by + 0.5f, dr);
        if (tid == 0) {
            const int k = M.kept++;
            if (k < CVB_HOUGH_MAX_CIRCLES) {
                res->xyr[k][0] = kx; res->xyr[k][1] = ky; res->xyr[k][2] = __uint_as_float((unsigned)(best >> 16));
                res->support[k] = (int)(best >> 48);
            }
        }
        for (int c = tid; c < ncent; c += HNT) {
            if (c_sup[c]) {
                const int cell = c_idx[c], ay = div_magic(cell, inv_astep), ax = cell - ay * astep;
                const float ex = __fsub_rn(kx, __fmul_rn((float)ax + 0.5f, dr)), ey = __fsub_rn(ky, __fmul_rn((float)ay + 0.5f, dr));
                if (__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)) < md2) c_sup[c] = 0;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        const int kept = M.kept;
        res->count = kept; res->n_edges = nnz; res->n_centers = ncent; res->status = CVB_HOUGH_OK;
        for (int k = kept; k < CVB_HOUGH_MAX_CIRCLES; ++k) {
            res->xyr[k][0] = res->xyr[k][1] = res->xyr[k][2] = 0.f;
            res->support[k] = 0;
        }
    }
}

}  // namespace

// Per-square geometry of one cv2.HoughCircles call (hough.cpp: HoughCircles / HoughCirclesGradient argument handling)
int cvb_host_hough_square(const cvb_rect &r, const cvb_hough_params &p, cvb_hough_square *out)
{
    const int md = std::min(r.w, r.h);
    int min_r = p.min_radius_ratio >= 0 ? (int)((double)md * p.min_radius_ratio) : p.min_radius;      // piece_detector.py:225
    int max_r = p.max_radius_ratio >= 0 ? (int)((double)md * p.max_radius_ratio) : p.max_radius;      // piece_detector.py:226
    const float min_dist = p.min_dist_div > 0 ? (float)(md / p.min_dist_div) : p.min_dist;            // piece_detector.py:236
    if (!(min_dist > 0)) return CVB_ERR_INVALID;
    if (min_r < 0) min_r = 0;
    if (max_r <= 0) max_r = std::max(r.w, r.h);
    else if (max_r <= min_r) max_r = min_r + 2;
    const float dp = p.dp < 1.f ? 1.f : p.dp, idp = 1.f / dp;
    out->x = r.x; out->y = r.y; out->w = r.w; out->h = r.h;
    out->min_radius = min_r; out->max_radius = max_r; out->min_dist = min_dist;
    out->acc_rows = (int)std::ceil(r.h * idp);
    out->acc_cols = (int)std::ceil(r.w * idp);
    out->n_bins = (int)std::lrint((float)((float)(max_r - min_r) / dp * 10));
    return CVB_OK;
}

int launch_hough(cvb_handle *h, const uint8_t *planes, int n, size_t plane_stride, int PW, const cvb_hough_square *d_squares,
                 const cvb_hough_square *squares, int n_sq, const uint8_t *d_select, const cvb_hough_params &p,
                 cvb_hough_result *out)
{
    const float dp = p.dp < 1.f ? 1.f : p.dp, idp = 1.f / dp;
    const int canny_high = (int)std::lrint(p.param1), acc_thr = (int)std::lrint(p.param2);
    const int canny_low = std::max(1, canny_high / 2);
    // shared-memory layout for the largest square of the list
    size_t acc_b = 0, nz_b = 0, bins_i = 0, early_b = 0, late_b = 0;
    for (int i = 0; i < n_sq; ++i) {
        const cvb_hough_square &s = squares[i];
        const size_t npad = (size_t)(s.h + 2) * (s.w + 2), ncells = (size_t)(s.acc_rows + 2) * (s.acc_cols + 2);
        acc_b = std::max(acc_b, ncells * 4);
        nz_b = std::max(nz_b, (size_t)s.h * s.w * 2);
        bins_i = std::max(bins_i, (size_t)std::max(s.n_bins, 1));
        early_b = std::max(early_b, 2 * ((npad + 3) & ~(size_t)3));
        late_b = std::max(late_b, 8 * (ncells / 2 + 1));
    }
    auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    HoughLayout L;
    L.bins_words_at = (int)((bins_i + 3) & ~(size_t)3);
    L.bins_pitch = L.bins_words_at + (int)((((bins_i + 31) >> 5) + 3) & ~(size_t)3);
    L.off_acc = 0;
    L.off_nz = (int)up16(std::max(acc_b, (size_t)L.bins_pitch * 4 * HW_));
    L.off_union = L.off_nz + (int)up16(nz_b);
    // ray steps behind the early view: what the candidate list leaves free, at least 512 entries
    const size_t uni_b = std::max(late_b, early_b + 4 * 512);
    L.dir_cap = (int)((uni_b - early_b) / 4);
    L.total = L.off_union + (int)up16(uni_b);
    if (L.total <= 220 * 1024) {
        const void *fn = (const void *)k_hough<false>;
        auto it = h->squares_smem_attr.find(fn);
        if (it == h->squares_smem_attr.end() || it->second < (size_t)L.total) {
            CVB_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
            h->squares_smem_attr[fn] = (size_t)L.total;
        }
        PROF(h, "k_hough");
        k_hough<false><<<dim3(n_sq, n), HNT, L.total, h->stream>>>(planes, plane_stride, PW, d_squares, n_sq, d_select, dp, idp,
                                                                   canny_low, canny_high, acc_thr, L, out, nullptr);
        LAUNCH_CHECK(h);
        return CVB_OK;
    }
    // larger squares: the same kernel on a global-memory slice per CTA, as many frames per launch as 2 GB of slices hold
    const size_t per_frame = (size_t)L.total * n_sq;
    const int fpl = (int)std::max<size_t>(1, std::min<size_t>((size_t)n, ((size_t)2 << 30) / per_frame));
    uint8_t *gws = nullptr;
    CVB_TRY(cvb_ws(h, h->ws_hough_gws, per_frame * fpl, (void **)&gws));
    for (int f0 = 0; f0 < n; f0 += fpl) {
        const int cnt = std::min(fpl, n - f0);
        PROF(h, "k_hough_gws");
        k_hough<true><<<dim3(n_sq, cnt), HNT, 0, h->stream>>>(planes + (size_t)f0 * plane_stride, plane_stride, PW, d_squares, n_sq,
                                                              d_select ? d_select + (size_t)f0 * n_sq : nullptr, dp, idp, canny_low,
                                                              canny_high, acc_thr, L, out + (size_t)f0 * n_sq, gws);
        LAUNCH_CHECK(h);
    }
    return CVB_OK;
}

CVB_BOUNDS_TU(hough)
