// cv2.Canny (8-bit, aperture 3, L1 gradient) and the edge projections of
// SmartGridExtractor.refine_grid (grid_extractor.py:66-121) -- a calibration-time
// "next" row of the scope table, built to the same parity bar.
//
// OpenCV imgproc/src/canny.cpp: Sobel 3x3 with BORDER_REPLICATE, mag = |dx|+|dy|,
// non-maximum suppression with tan(22.5 deg) in Q15 (13573) against zero-padded
// magnitudes, hysteresis: every candidate (mag > low, NMS maximum) that is
// 8-connected to a candidate with mag > high is an edge.  The result does not
// depend on traversal order, so the hysteresis is an iterated tile-local flood.
#include "cvb_device.cuh"
#include <cmath>

namespace {

constexpr int CT = 32;   // tile edge

// map: 0 weak candidate, 1 no edge, 2 edge
__global__ void __launch_bounds__(256) k_canny_nms(const uint8_t *__restrict__ src, int H, int W, int low, int high,
                                                   uint8_t *__restrict__ map)
{
    __shared__ uint8_t s_g[CT + 4][CT + 4];
    __shared__ int s_m[CT + 2][CT + 2];
    __shared__ short s_dx[CT][CT], s_dy[CT][CT];
    const int tid = threadIdx.x, frame = blockIdx.z;
    const int x0 = blockIdx.x * CT, y0 = blockIdx.y * CT;
    const uint8_t *img = src + (size_t)frame * H * W;
    for (int i = tid; i < (CT + 4) * (CT + 4); i += 256) {
        const int ly = i / (CT + 4), lx = i - ly * (CT + 4);
        const int y = min(max(y0 - 2 + ly, 0), H - 1), x = min(max(x0 - 2 + lx, 0), W - 1);   // BORDER_REPLICATE
        s_g[ly][lx] = img[(size_t)y * W + x];
    }
    __syncthreads();
    for (int i = tid; i < (CT + 2) * (CT + 2); i += 256) {
        const int ly = i / (CT + 2), lx = i - ly * (CT + 2);
        const int y = y0 - 1 + ly, x = x0 - 1 + lx;
        int m = 0;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            // centre of the 3x3 window in s_g is (ly + 1, lx + 1)
            const int a = s_g[ly][lx], b = s_g[ly][lx + 1], c = s_g[ly][lx + 2];
            const int d = s_g[ly + 1][lx], f = s_g[ly + 1][lx + 2];
            const int g = s_g[ly + 2][lx], hh = s_g[ly + 2][lx + 1], k = s_g[ly + 2][lx + 2];
            const int gx = (c + 2 * f + k) - (a + 2 * d + g);
            const int gy = (g + 2 * hh + k) - (a + 2 * b + c);
            m = abs(gx) + abs(gy);
            if (ly >= 1 && ly <= CT && lx >= 1 && lx <= CT) { s_dx[ly - 1][lx - 1] = (short)gx; s_dy[ly - 1][lx - 1] = (short)gy; }
        }
        s_m[ly][lx] = m;      // zeros outside the image, as OpenCV's padded magnitude rows
    }
    __syncthreads();
    for (int i = tid; i < CT * CT; i += 256) {
        const int ly = i / CT, lx = i - ly * CT;
        const int y = y0 + ly, x = x0 + lx;
        if (y >= H || x >= W) continue;
        const int m = s_m[ly + 1][lx + 1];
        bool cand = false;
        if (m > low) {
            const int xs = s_dx[ly][lx], ys = s_dy[ly][lx];
            const long long ax = abs(xs), ay = (long long)abs(ys) << 15;
            const long long tg22x = ax * 13573, tg67x = tg22x + (ax << 16);
            if (ay < tg22x) cand = m > s_m[ly + 1][lx] && m >= s_m[ly + 1][lx + 2];
            else if (ay > tg67x) cand = m > s_m[ly][lx + 1] && m >= s_m[ly + 2][lx + 1];
            else {
                const int s = (xs ^ ys) < 0 ? -1 : 1;
                cand = m > s_m[ly][lx + 1 - s] && m > s_m[ly + 2][lx + 1 + s];
            }
        }
        map[(size_t)frame * H * W + (size_t)y * W + x] = cand ? (m > high ? 2 : 0) : 1;
    }
}

// one hysteresis sweep: flood inside each tile (with a 1-pixel halo of the current map) until the tile is stable
__global__ void __launch_bounds__(256) k_canny_hyst(uint8_t *__restrict__ map, int H, int W, int *__restrict__ changed)
{
    __shared__ uint8_t s[CT + 2][CT + 2];
    const int tid = threadIdx.x, frame = blockIdx.z;
    const int x0 = blockIdx.x * CT, y0 = blockIdx.y * CT;
    uint8_t *mp = map + (size_t)frame * H * W;
    for (int i = tid; i < (CT + 2) * (CT + 2); i += 256) {
        const int ly = i / (CT + 2), lx = i - ly * (CT + 2);
        const int y = y0 - 1 + ly, x = x0 - 1 + lx;
        s[ly][lx] = (y >= 0 && y < H && x >= 0 && x < W) ? mp[(size_t)y * W + x] : 1;
    }
    __syncthreads();
    bool any = false;
    for (;;) {
        bool mine = false;
        for (int i = tid; i < CT * CT; i += 256) {
            const int ly = i / CT + 1, lx = i % CT + 1;
            if (s[ly][lx] == 0) {
                const bool hit = s[ly - 1][lx - 1] == 2 || s[ly - 1][lx] == 2 || s[ly - 1][lx + 1] == 2 || s[ly][lx - 1] == 2 ||
                                 s[ly][lx + 1] == 2 || s[ly + 1][lx - 1] == 2 || s[ly + 1][lx] == 2 || s[ly + 1][lx + 1] == 2;
                if (hit) { s[ly][lx] = 2; mine = true; }     // monotone 0 -> 2: racing readers only see it earlier
            }
        }
        if (!__syncthreads_or(mine)) break;
        any = true;
    }
    if (any) {
        for (int i = tid; i < CT * CT; i += 256) {
            const int ly = i / CT, lx = i % CT;
            const int y = y0 + ly, x = x0 + lx;
            if (y < H && x < W && s[ly + 1][lx + 1] == 2) mp[(size_t)y * W + x] = 2;
        }
        if (tid == 0) atomicExch(changed, 1);
    }
}

__global__ void __launch_bounds__(256) k_canny_final(const uint8_t *__restrict__ map, size_t n, uint8_t *__restrict__ dst)
{
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) dst[i] = map[i] == 2 ? 255 : 0;
}

// row / column sums of a u8 plane (np.sum(edges, axis=1 / 0)); sparse edges -> few atomics
__global__ void __launch_bounds__(256) k_projections(const uint8_t *__restrict__ plane, int H, int W, uint32_t *__restrict__ rows,
                                                     uint32_t *__restrict__ cols)
{
    const int frame = blockIdx.z;
    const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
    if (x >= W || y >= H) return;
    const uint32_t v = plane[(size_t)frame * H * W + (size_t)y * W + x];
    if (v) {
        atomicAdd(rows + (size_t)frame * H + y, v);
        atomicAdd(cols + (size_t)frame * W + x, v);
    }
}

// cv2.dilate with a kw x kh rectangle, `iterations` times = one max filter of radius (rx, ry); positions outside the
// image do not take part (morphologyDefaultBorderValue).  64 x 16 output tile, rows then columns in shared memory.
constexpr int DT_W = 64, DT_H = 16, D_MAXR = 24;
__global__ void __launch_bounds__(256) k_dilate(const uint8_t *__restrict__ src, int H, int W, int rx, int ry, uint8_t *__restrict__ dst)
{
    __shared__ uint8_t s_in[DT_H + 2 * D_MAXR][DT_W + 2 * D_MAXR];
    __shared__ uint8_t s_row[DT_H + 2 * D_MAXR][DT_W];
    const int frame = blockIdx.z, x0 = blockIdx.x * DT_W, y0 = blockIdx.y * DT_H;
    const uint8_t *img = src + (size_t)frame * H * W;
    const int iw = DT_W + 2 * rx, ih = DT_H + 2 * ry;
    for (int i = threadIdx.x; i < iw * ih; i += 256) {
        const int ly = i / iw, lx = i - ly * iw, y = y0 - ry + ly, x = x0 - rx + lx;
        s_in[ly][lx] = (y >= 0 && y < H && x >= 0 && x < W) ? img[(size_t)y * W + x] : 0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DT_W * ih; i += 256) {
        const int ly = i / DT_W, lx = i - ly * DT_W;
        int m = 0;
        for (int k = 0; k <= 2 * rx; ++k) m = max(m, (int)s_in[ly][lx + k]);
        s_row[ly][lx] = (uint8_t)m;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DT_W * DT_H; i += 256) {
        const int ly = i / DT_W, lx = i - ly * DT_W, y = y0 + ly, x = x0 + lx;
        if (y >= H || x >= W) continue;
        int m = 0;
        for (int k = 0; k <= 2 * ry; ++k) m = max(m, (int)s_row[ly + k][lx]);
        dst[(size_t)frame * H * W + (size_t)y * W + x] = (uint8_t)m;
    }
}

}  // namespace

int launch_dilate(cvb_handle *h, const uint8_t *src, int n, int H, int W, int kw, int kh, int iterations, uint8_t *dst)
{
    const int rx = (kw - 1) * iterations / 2, ry = (kh - 1) * iterations / 2;
    if (rx > D_MAXR || ry > D_MAXR) {
        cvb_set_error("dilate: (k - 1) * iterations / 2 = %d x %d exceeds the supported radius %d", rx, ry, D_MAXR);
        return CVB_ERR_INVALID;
    }
    dim3 grid((W + DT_W - 1) / DT_W, (H + DT_H - 1) / DT_H, n);
    PROF(h, "k_dilate");
    k_dilate<<<grid, 256, 0, h->stream>>>(src, H, W, rx, ry, dst);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

int launch_canny(cvb_handle *h, const uint8_t *gray, int n, int H, int W, double low_thresh, double high_thresh, uint8_t *edges)
{
    if (low_thresh > high_thresh) std::swap(low_thresh, high_thresh);
    const int low = (int)std::floor(low_thresh), high = (int)std::floor(high_thresh);
    uint8_t *map = nullptr;
    int *flag = nullptr;
    CVB_TRY(cvb_ws(h, h->ws_plane, (size_t)n * H * W, (void **)&map));
    CVB_TRY(cvb_ws(h, h->ws_plane2, 256, (void **)&flag));
    dim3 grid((W + CT - 1) / CT, (H + CT - 1) / CT, n);
    PROF(h, "k_canny_nms");
    k_canny_nms<<<grid, 256, 0, h->stream>>>(gray, H, W, low, high, map);
    LAUNCH_CHECK(h);
    // Sweeps until no tile changes.  A sweep floods inside every 32-px tile, so a weak chain advances by at least one
    // tile border per sweep; the map only ever moves 0 -> 2, hence the loop terminates (a serpentine chain can need about
    // tiles_x * tiles_y sweeps, far more than any fixed small cap -- cv2.Canny follows it to the end, so do we).
    for (;;) {
        CVB_CHECK_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), h->stream));
        PROF(h, "k_canny_hyst");
        k_canny_hyst<<<grid, 256, 0, h->stream>>>(map, H, W, flag);
        LAUNCH_CHECK(h);
        int changed = 0;
        CVB_CHECK_CUDA(cudaMemcpyAsync(&changed, flag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
        if (!changed) break;
    }
    const size_t total = (size_t)n * H * W;
    PROF(h, "k_canny_final");
    k_canny_final<<<(unsigned)std::min<size_t>((total + 255) / 256, 4096), 256, 0, h->stream>>>(map, total, edges);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

int launch_projections(cvb_handle *h, const uint8_t *plane, int n, int H, int W, uint32_t *rows, uint32_t *cols)
{
    CVB_CHECK_CUDA(cudaMemsetAsync(rows, 0, sizeof(uint32_t) * (size_t)n * H, h->stream));
    CVB_CHECK_CUDA(cudaMemsetAsync(cols, 0, sizeof(uint32_t) * (size_t)n * W, h->stream));
    dim3 grid((W + 63) / 64, (H + 3) / 4, n);
    PROF(h, "k_projections");
    k_projections<<<grid, 256, 0, h->stream>>>(plane, H, W, rows, cols);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

CVB_BOUNDS_TU(canny)
