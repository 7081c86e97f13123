// Board overlay: the drawing calls of GameSession._draw_interface (game_session.py:293-388) as a display list applied
// on the device, pixel-exactly as OpenCV's drawing.cpp / arithm would:
//   cv2.line (axis-aligned, thickness 1), cv2.rectangle(..., -1)         -> RECT, inclusive corners, clipped
//   cv2.circle(..., -1)                                                  -> CIRCLE, the midpoint spans of Circle()
//   cv2.putText                                                          -> STAMP, a 1-bit mask rasterised by the caller
//   overlay = vis.copy(); <shapes>; cv2.addWeighted(overlay, a, vis, b, 0, vis)
//                                                                        -> the same ops with (alpha, beta) != (1, 0):
//        dst = saturate(rint(fma(colour, (float)a, dst * (float)b)))     (arithm: v_fma(src1, alpha, v_fma(src2, beta, gamma)));
//        outside the shapes addWeighted(v, a, v, b) == v for every byte (tests/test_overlay_cpu.py), so only covered pixels
//        change; ops of one `group` are the shapes drawn on ONE overlay copy: a pixel is blended once per group.
// Every op reads only the pixel it writes, so one thread carries a pixel through the whole list in order.
#include "cvb_internal.h"
#include "cvb_device.cuh"

namespace {

constexpr int OV_TW = 32, OV_TH = 8, OV_MAX_OPS = 1024;

struct OverlayArgs {
    uint8_t *img;
    int H, W;
    const cvb_overlay_op *ops;
    int n_ops;
    const uint8_t *aux;      // circle span tables (u16 half widths for dy = 0..r) and stamp bit masks
};

CVB_DEV void op_bbox(const cvb_overlay_op &o, int &x0, int &y0, int &x1, int &y1)
{
    if (o.kind == CVB_OV_RECT) { x0 = o.x0; y0 = o.y0; x1 = o.x1; y1 = o.y1; }
    else if (o.kind == CVB_OV_CIRCLE) { x0 = o.x0 - o.x1; x1 = o.x0 + o.x1; y0 = o.y0 - o.x1; y1 = o.y0 + o.x1; }
    else { x0 = o.x0; y0 = o.y0; x1 = o.x0 + o.x1 - 1; y1 = o.y0 + o.y1 - 1; }
}

CVB_DEV bool op_covers(const cvb_overlay_op &o, const uint8_t *aux, int x, int y)
{
    int x0, y0, x1, y1;
    op_bbox(o, x0, y0, x1, y1);
    if (x < x0 || x > x1 || y < y0 || y > y1) return false;
    if (o.kind == CVB_OV_RECT) return true;
    if (o.kind == CVB_OV_CIRCLE) {
        const int dy = abs(y - o.y0), dx = abs(x - o.x0);
        return dx <= (int)reinterpret_cast<const uint16_t *>(aux + o.aux_ofs)[dy];
    }
    const int bx = x - x0, by = y - y0, row_bytes = (o.x1 + 7) >> 3;
    return (aux[o.aux_ofs + (size_t)by * row_bytes + (bx >> 3)] >> (bx & 7)) & 1;
}

CVB_DEV uint32_t blend_u8(uint32_t colour, float alpha, uint32_t v, float beta)
{
    const float r = __fmaf_rn((float)colour, alpha, __fmul_rn((float)v, beta));
    const int q = __float2int_rn(r);
    return (uint32_t)min(max(q, 0), 255);
}

__global__ void __launch_bounds__(OV_TW * OV_TH) k_overlay(const OverlayArgs a)
{
    __shared__ uint32_t hit[OV_MAX_OPS / 32];
    const int tid = threadIdx.y * OV_TW + threadIdx.x;
    const int tx0 = blockIdx.x * OV_TW, ty0 = blockIdx.y * OV_TH;
    uint8_t *img = a.img + (size_t)blockIdx.z * a.H * a.W * 3;
    if (tid < OV_MAX_OPS / 32) hit[tid] = 0;
    __syncthreads();
    // which ops touch this tile (bounding boxes), kept as a bit set so that the list order survives
    for (int i = tid; i < a.n_ops; i += OV_TW * OV_TH) {
        int x0, y0, x1, y1;
        op_bbox(a.ops[i], x0, y0, x1, y1);
        if (x1 >= tx0 && x0 < tx0 + OV_TW && y1 >= ty0 && y0 < ty0 + OV_TH) atomicOr(&hit[i >> 5], 1u << (i & 31));
    }
    __syncthreads();
    const int x = tx0 + threadIdx.x, y = ty0 + threadIdx.y;
    if (x >= a.W || y >= a.H) return;
    uint8_t *p = img + ((size_t)y * a.W + x) * 3;
    CVB_BOUNDS(p + 2 < a.img + (size_t)gridDim.z * a.H * a.W * 3);
    uint32_t b = p[0], g = p[1], r = p[2];
    bool dirty = false;
    int done_group = 0;
    const int words = (a.n_ops + 31) >> 5;
    for (int w = 0; w < words; ++w) {
        uint32_t m = hit[w];
        while (m) {
            const int i = (w << 5) + __ffs(m) - 1;
            m &= m - 1;
            const cvb_overlay_op o = a.ops[i];
            if (o.group != 0 && o.group == done_group) continue;      // already blended through this overlay copy
            if (!op_covers(o, a.aux, x, y)) continue;
            if (o.alpha == 1.0f && o.beta == 0.0f) { b = o.color[0]; g = o.color[1]; r = o.color[2]; }
            else {
                b = blend_u8(o.color[0], o.alpha, b, o.beta);
                g = blend_u8(o.color[1], o.alpha, g, o.beta);
                r = blend_u8(o.color[2], o.alpha, r, o.beta);
            }
            done_group = o.group;
            dirty = true;
        }
    }
    if (dirty) { p[0] = (uint8_t)b; p[1] = (uint8_t)g; p[2] = (uint8_t)r; }
}

}   // namespace

int cvb_overlay_max_ops() { return OV_MAX_OPS; }

int launch_overlay(cvb_handle *h, uint8_t *bgr, int n, int H, int W, const cvb_overlay_op *d_ops, int n_ops,
                   const uint8_t *d_aux)
{
    const dim3 grid((W + OV_TW - 1) / OV_TW, (H + OV_TH - 1) / OV_TH, n), block(OV_TW, OV_TH);
    OverlayArgs a{bgr, H, W, d_ops, n_ops, d_aux};
    PROF(h, "k_overlay");
    k_overlay<<<grid, block, 0, h->stream>>>(a);
    LAUNCH_CHECK(h);
    return CVB_OK;
}

CVB_BOUNDS_TU(overlay)
