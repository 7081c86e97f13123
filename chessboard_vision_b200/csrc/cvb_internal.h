// Internal declarations shared by the translation units of libcvb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include <map>
#include <set>
#include "../../include/cvb200.h"

// ---- error plumbing ---------------------------------------------------------
void cvb_set_error(const char *fmt, ...);
#define CVB_CHECK_CUDA(expr)                                                           \
    do {                                                                               \
        cudaError_t _e = (expr);                                                       \
        if (_e != cudaSuccess) {                                                       \
            cvb_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                          __FILE__, __LINE__);                                         \
            return CVB_ERR_CUDA;                                                       \
        }                                                                              \
    } while (0)
#define CVB_REQUIRE(cond, ...)                                                         \
    do {                                                                               \
        if (!(cond)) { cvb_set_error(__VA_ARGS__); return CVB_ERR_INVALID; }           \
    } while (0)
#define CVB_TRY(expr)                                                                  \
    do { int _rc = (expr); if (_rc != CVB_OK) return _rc; } while (0)

// ---- tables -----------------------------------------------------------------
// Integer colour-conversion LUTs (restating OpenCV color_lab.cpp, see
// cvb_tables.cpp) in the layout the kernels stage into shared memory.
struct CvbTables {
    uint16_t gamma[256];       // sRGBGammaTab_b
    uint16_t cbrt[2048];       // LabCbrtTab_b, index <= 2040 reachable
    int32_t  lab2yf[512];      // LabToYF_b: [L] = y | ify << 16 for L < 256, rest padding
    uint8_t  invgamma[4096];   // sRGBInvGammaTab_b
    uint8_t  ltab[2048];       // L as a function of the Y index (histogram pass)
};
const CvbTables &cvb_host_tables();
void cvb_host_bilateral_tables(double sigma_color, double sigma_space, float *color768, float *space81);
int  cvb_host_gaussian_q8(int ksize, int *q);
int  cvb_host_gaussian_q8_sigma(int ksize, double sigma, int *q);
int  cvb_host_get_perspective(const float *src, const float *dst, double *M);
int  cvb_host_invert3(const double *a, double *t);
void cvb_host_square_masks(int h, int w, uint8_t *mask);

// ---- device workspace -----------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct ProfRec { const char *name; cudaEvent_t e0, e1; };

// staged copies of the host-side rectangle list / matrices of the last call (content-compared)
struct RectCache {
    std::vector<cvb_rect> rects;
    std::vector<uint8_t> select;
    bool has_select = false;
    int max_px = 0;
    size_t mask_bytes = 0;
};

constexpr unsigned CVB_TICKET_RING = 8;
struct cvb_handle {
    RectCache rect_cache;
    std::vector<double> mat_cache;
    std::set<const void *> fused_attr_done;             // kernels whose dynamic-smem attribute is set on this device
    std::map<const void *, size_t> squares_smem_attr;
    bool profiling = false;
    std::vector<ProfRec> prof;
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 0;
    int64_t launches = 0;
    CvbTables *d_tables = nullptr;
    // bilateral colour LUT cache (device) keyed by sigma_color
    float *d_color = nullptr;
    double color_sigma = -1.0, space_sigma = -1.0;
    // grow-only scratch
    DevBuf ws_lab, ws_prof, ws_in, ws_sharp, ws_enh, ws_gray, ws_blur, ws_bin, ws_warp, ws_plane, ws_plane2;
    DevBuf ws_hist, ws_lut, ws_minmax, ws_ohist, ws_otsu, ws_otsu_all, ws_stats, ws_rects, ws_select, ws_mats;
    // host-buffer pipeline: copy stream + double-buffer events, frames per chunk
    // forked tail of the analysis stage (Otsu scan + mask) for small batches: it runs on aux_stream beside the warp and
    // the square kernel, which do not depend on it (cvb_api.cu: analysis_tail / join_tail)
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool fork_tail = false, tail_pending = false;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    uint64_t tickets = 0;       // cvb_pipeline_submit: submissions so far; ev_ticket[seq % ring] marks the end of submission seq
    cudaEvent_t ev_ticket[8] = {};
    size_t stage_layout = 0;    // bytes per staging buffer of the last host-buffer pipeline call
    int chunk_frames = 0;       // frames per chunk of the host-buffer pipeline; 0 = chosen per call (cvb_pipeline_fmt)
    // mask cache for square shapes: (h<<16|w) -> offset into ws_masks
    DevBuf ws_masks;
    // Hough: staged per-square geometry (content-compared), select bytes, results
    std::vector<cvb_hough_square> hough_cache;
    DevBuf ws_hough_sq, ws_hough_sel, ws_hough_res, ws_hough_gws;
    DevBuf ws_stage;            // the two chunk staging buffers of the host-buffer pipeline (its own: see pipeline_fmt_impl)
    DevBuf ws_overlay;          // display list + circle span tables + stamp masks of the last cvb_overlay_dev call
    void *pinned = nullptr;
    size_t pinned_cap = 0;
};

struct cvb_state {
    cvb_handle *h = nullptr;
    int n_streams = 0, BH = 0, BW = 0;
    uint8_t *pd_ref = nullptr;   // n_streams * BH*BW
    uint8_t *pd_cur = nullptr;
    float *cd_mean = nullptr;
    float *cd_var = nullptr;
    uint8_t *flags = nullptr;    // per pixel: bit0 has_ref, bit1 has_cd
};

int cvb_ws(cvb_handle *h, DevBuf &b, size_t bytes, void **out);
void cvb_prof_begin(cvb_handle *h, const char *name);
void cvb_prof_end(cvb_handle *h);
// bracket a kernel launch: PROF(h, "k_name"); k<<<...>>>(...); LAUNCH_CHECK(h);
#define PROF(h, name) do { if ((h)->profiling) cvb_prof_begin((h), (name)); } while (0)
#define LAUNCH_CHECK(h)                                  \
    do {                                                 \
        if ((h)->profiling) cvb_prof_end(h);             \
        (h)->launches++;                                 \
        CVB_CHECK_CUDA(cudaGetLastError());              \
    } while (0)

// ---- kernel launchers (cvb_enhance.cu) --------------------------------------------
struct ClaheGeom {
    int tiles_x, tiles_y, tile_w, tile_h, ext_w, ext_h, clip;
    float lut_scale, inv_tw, inv_th;
};
int cvb_clahe_geom(int H, int W, double clip_limit, int tx, int ty, ClaheGeom *g);

int launch_color_profile(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_color_profile &p, uint8_t *out);
int launch_bgr2lab(cvb_handle *h, const uint8_t *bgr, long npx, uint8_t *lab);
int launch_lab2bgr(cvb_handle *h, const uint8_t *lab, long npx, uint8_t *bgr);
// histogram of L (from_bgr=1: src is BGR, L computed on the fly) or of a u8 plane
// lab_out (from_bgr only, may be null): the Lab pixels of the frames, for launch_fused(..., src_is_lab = true)
int launch_tile_hist(cvb_handle *h, const uint8_t *src, int from_bgr, int n, int H, int W,
                     const ClaheGeom &g, int32_t *hist, int32_t *minmax_init, uint8_t *lab_out);
int launch_clahe_lut(cvb_handle *h, const int32_t *hist, int n, const ClaheGeom &g, uint8_t *lut);
int launch_clahe_apply_plane(cvb_handle *h, const uint8_t *src, int n, int H, int W, const ClaheGeom &g,
                             const uint8_t *lut, uint8_t *dst);
int launch_correct_lighting(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const ClaheGeom &g,
                            const uint8_t *lut, uint8_t *out);
// the fused tile kernel: [lighting] -> [bilateral] -> [sharpen] (+ min/max)
int launch_fused(cvb_handle *h, const uint8_t *src, int n, int H, int W, bool light, bool bilateral, bool sharpen,
                 const ClaheGeom *g, const uint8_t *lut, double sigma_color, double sigma_space,
                 uint8_t *out, int32_t *minmax, bool src_is_lab = false);
// cvb_fused2.cu: the same three stages as one persistent kernel fed by tensor-map TMA loads of the Lab tiles
bool fused_tma_applicable(int H, int W, const uint8_t *lab, const uint8_t *out);
int launch_fused_tma(cvb_handle *h, const uint8_t *lab, int n, int H, int W, const ClaheGeom &g, const uint8_t *lut,
                     const float *d_wlut, const float *space81, uint8_t *out, int32_t *minmax, int variant);
int launch_minmax(cvb_handle *h, const uint8_t *src, int n, long bytes_per_frame, int32_t *minmax);
int launch_normalize(cvb_handle *h, const uint8_t *src, int n, long bytes_per_frame, const int32_t *minmax,
                     uint8_t *out);
int launch_gray(cvb_handle *h, const uint8_t *bgr, long npx, uint8_t *gray);
int launch_gaussian(cvb_handle *h, const uint8_t *src, int n, int H, int W, int ksize, double sigma, uint8_t *dst);
// normalize (optional) + gray + blur5 + 256-bin histogram
int launch_finish(cvb_handle *h, const uint8_t *src, int n, int H, int W, const int32_t *minmax,
                  uint8_t *enhanced, uint8_t *gray, uint8_t *blurred, int32_t *hist);
int launch_otsu(cvb_handle *h, const int32_t *hist, int n, long npx, int32_t *otsu_t);
int launch_threshold(cvb_handle *h, const uint8_t *src, int n, long npx, const int32_t *otsu_t, uint8_t *dst);

// per-translation-unit counters of the debug build's index checks (cvb_device.cuh)
void cvb_bounds_enhance(unsigned long long *, int *); void cvb_bounds_fused2(unsigned long long *, int *);
void cvb_bounds_grid(unsigned long long *, int *); void cvb_bounds_canny(unsigned long long *, int *);
void cvb_bounds_hough(unsigned long long *, int *); void cvb_bounds_ingest(unsigned long long *, int *);
void cvb_bounds_overlay(unsigned long long *, int *);

// ---- cvb_ingest.cu ------------------------------------------------------------------------
size_t cvb_host_frame_bytes(int format, int H, int W);
int launch_yuv_to_bgr(cvb_handle *h, const uint8_t *src, int format, int n, int H, int W, uint8_t *bgr);
// cvb_overlay.cu: one slice (<= cvb_overlay_max_ops() ops) of a display list, ops and aux resident on the device
int cvb_overlay_max_ops();
int launch_overlay(cvb_handle *h, uint8_t *bgr, int n, int H, int W, const cvb_overlay_op *d_ops, int n_ops, const uint8_t *d_aux);

// ---- cvb_canny.cu -------------------------------------------------------------------------
int launch_canny(cvb_handle *h, const uint8_t *gray, int n, int H, int W, double low_thresh, double high_thresh, uint8_t *edges);
int launch_dilate(cvb_handle *h, const uint8_t *src, int n, int H, int W, int kw, int kh, int iterations, uint8_t *dst);
int launch_projections(cvb_handle *h, const uint8_t *plane, int n, int H, int W, uint32_t *rows, uint32_t *cols);

// ---- cvb_grid.cu ----------------------------------------------------------------------
int launch_warp(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const double *d_minv, int n_mats,
                int out_h, int out_w, int rot180, uint8_t *warped);
int launch_rotate(cvb_handle *h, const uint8_t *src, int n, int H, int W, int C, int code, uint8_t *dst);
int launch_squares(cvb_handle *h, const uint8_t *boards, int n, int BH, int BW, int C,
                   const cvb_rect *d_rects, const int32_t *d_mask_ofs, const uint8_t *d_masks, int n_sq, int max_px,
                   const uint8_t *d_select, cvb_state *st, int stream0, const cvb_square_params &p,
                   const int *pd_q, const int *cd_q, cvb_square_stats *stats);
int launch_state_reset(cvb_handle *h, cvb_state *s, int stream);

// ---- cvb_hough.cu ---------------------------------------------------------------------
int cvb_host_hough_square(const cvb_rect &r, const cvb_hough_params &p, cvb_hough_square *out);
int launch_hough(cvb_handle *h, const uint8_t *planes, int n, size_t plane_stride, int PW, const cvb_hough_square *d_squares,
                 const cvb_hough_square *squares, int n_sq, const uint8_t *d_select, const cvb_hough_params &p,
                 cvb_hough_result *out);
