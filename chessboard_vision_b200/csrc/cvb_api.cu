// C ABI of libcvb200.so (see include/cvb200.h).  Handle, workspace, argument
// checking and the composition of the kernel launchers into the reference's
// methods (frame_enhancer.py:101-181, board_detection.py:61-71,
// change_detector.py:36-167, piece_detector.py:70-97).
#include "cvb_internal.h"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>

static_assert(sizeof(cvb_color_profile) == 48 && sizeof(cvb_enhance_params) == 96, "ABI struct layout (see _lib.py)");
static_assert(sizeof(cvb_overlay_op) == 40, "ABI struct layout (see _lib.py)");
static thread_local char g_err[512] = "";

void cvb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cvb_ws(cvb_handle *h, DevBuf &b, size_t bytes, void **out)
{
    if (bytes > b.cap) {
        // a grow may free a buffer that queued work still reads
        CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
        if (b.p) CVB_CHECK_CUDA(cudaFree(b.p));
        b.p = nullptr; b.cap = 0;
        size_t cap = bytes + bytes / 8 + 256;
        CVB_CHECK_CUDA(cudaMalloc(&b.p, cap));
        b.cap = cap;
    }
    *out = b.p;
    return CVB_OK;
}

void cvb_prof_begin(cvb_handle *h, const char *name)
{
    ProfRec r;
    r.name = name;
    if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
    cudaEventRecord(r.e0, h->stream);
    h->prof.push_back(r);
}
void cvb_prof_end(cvb_handle *h)
{
    if (!h->prof.empty()) cudaEventRecord(h->prof.back().e1, h->stream);
}
static void prof_clear(cvb_handle *h)
{
    for (ProfRec &r : h->prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    h->prof.clear();
}

#define WS(buf, type, count, var) \
    type *var = nullptr;          \
    CVB_TRY(cvb_ws(h, h->buf, sizeof(type) * (size_t)(count), (void **)&var))

// every public entry point runs on the handle's own device, whatever device the calling thread had current
// (several engines, one per GPU, may live in one process)
#define REQ_H(h)                                         \
    do {                                                 \
        CVB_REQUIRE((h) != nullptr, "null handle");      \
        CVB_CHECK_CUDA(cudaSetDevice((h)->device));      \
    } while (0)
#define REQ_IMG(n, H, W)                                                                                      \
    CVB_REQUIRE((n) >= 1 && (n) <= 65535 && (H) >= 1 && (W) >= 1 && (H) <= 32768 && (W) <= 32768,              \
                "bad batch/shape n=%d H=%d W=%d", (n), (H), (W))

extern "C" {

int cvb_version(void) { return CVB_VERSION; }
const char *cvb_last_error(void) { return g_err; }

int cvb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int cvb_create(int device, cvb_handle **out)
{
    CVB_REQUIRE(out != nullptr, "null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        cvb_set_error("no CUDA device available (%s); libcvb200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return CVB_ERR_NO_DEVICE;
    }
    CVB_REQUIRE(device >= 0 && device < n, "device %d out of range (0..%d)", device, n - 1);
    CVB_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    CVB_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        cvb_set_error("device %d is sm_%d%d; libcvb200 is built for sm_100a only", device, prop.major, prop.minor);
        return CVB_ERR_NO_DEVICE;
    }
    cvb_handle *h = new cvb_handle();
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    CVB_CHECK_CUDA(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    CVB_CHECK_CUDA(cudaMalloc(&h->d_tables, sizeof(CvbTables)));
    CVB_CHECK_CUDA(cudaMemcpy(h->d_tables, &cvb_host_tables(), sizeof(CvbTables), cudaMemcpyHostToDevice));
    *out = h;
    return CVB_OK;
}

void cvb_destroy(cvb_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf *bufs[] = {&h->ws_lab, &h->ws_prof, &h->ws_in, &h->ws_sharp, &h->ws_enh, &h->ws_gray, &h->ws_blur, &h->ws_bin, &h->ws_warp,
                      &h->ws_plane, &h->ws_plane2, &h->ws_hist, &h->ws_lut, &h->ws_minmax, &h->ws_ohist, &h->ws_otsu,
                      &h->ws_stats, &h->ws_otsu_all, &h->ws_rects, &h->ws_select, &h->ws_mats, &h->ws_masks,
                      &h->ws_hough_sq, &h->ws_hough_sel, &h->ws_hough_res, &h->ws_hough_gws, &h->ws_stage, &h->ws_overlay};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    for (cudaEvent_t e : h->ev_ticket) if (e) cudaEventDestroy(e);
    if (h->d_tables) cudaFree(h->d_tables);
    if (h->d_color) cudaFree(h->d_color);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->aux_stream) { cudaStreamDestroy(h->aux_stream); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join); }
    if (h->copy_stream) {
        cudaStreamDestroy(h->copy_stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(h->ev_copy[i]); cudaEventDestroy(h->ev_done[i]); }
    }
    prof_clear(h);
    delete h;
}

int cvb_set_stream(cvb_handle *h, void *cuda_stream)
{
    REQ_H(h);
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    return CVB_OK;
}
int cvb_synchronize(cvb_handle *h)
{
    REQ_H(h);
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    return CVB_OK;
}
int64_t cvb_launch_count(cvb_handle *h) { return h ? h->launches : -1; }

int cvb_profile_enable(cvb_handle *h, int on)
{
    REQ_H(h);
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    prof_clear(h);
    h->profiling = on != 0;
    return CVB_OK;
}
int cvb_profile_read(cvb_handle *h, char *names32, float *total_ms, int *counts, int max_entries, int *n_out)
{
    REQ_H(h);
    CVB_REQUIRE(names32 && total_ms && counts && n_out && max_entries > 0, "null pointer");
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    int n = 0;
    for (const ProfRec &r : h->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
        int k = 0;
        for (; k < n; ++k)
            if (strncmp(names32 + 32 * k, r.name, 31) == 0) break;
        if (k == n) {
            if (n == max_entries) continue;
            memset(names32 + 32 * k, 0, 32);
            strncpy(names32 + 32 * k, r.name, 31);
            total_ms[k] = 0.f; counts[k] = 0;
            ++n;
        }
        total_ms[k] += ms; counts[k] += 1;
    }
    *n_out = n;
    return CVB_OK;
}

int cvb_malloc(cvb_handle *h, size_t bytes, void **dptr)
{
    REQ_H(h);
    CVB_REQUIRE(dptr != nullptr, "null out pointer");
    CVB_CHECK_CUDA(cudaSetDevice(h->device));
    CVB_CHECK_CUDA(cudaMalloc(dptr, bytes ? bytes : 1));
    return CVB_OK;
}
int cvb_free(cvb_handle *h, void *dptr)
{
    REQ_H(h);
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    CVB_CHECK_CUDA(cudaFree(dptr));
    return CVB_OK;
}
int cvb_host_alloc(size_t bytes, void **hptr)
{
    CVB_REQUIRE(hptr != nullptr, "null out pointer");
    CVB_CHECK_CUDA(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return CVB_OK;
}
int cvb_host_free(void *hptr)
{
    CVB_CHECK_CUDA(cudaFreeHost(hptr));
    return CVB_OK;
}
int cvb_memcpy_h2d(cvb_handle *h, void *dst, const void *src, size_t bytes)
{
    REQ_H(h);
    CVB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    return CVB_OK;
}
int cvb_memcpy_d2h(cvb_handle *h, void *dst, const void *src, size_t bytes)
{
    REQ_H(h);
    CVB_CHECK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    return CVB_OK;
}
int cvb_memset(cvb_handle *h, void *dst, int value, size_t bytes)
{
    REQ_H(h);
    CVB_CHECK_CUDA(cudaMemsetAsync(dst, value, bytes, h->stream));
    return CVB_OK;
}
int cvb_event_create(void **ev)
{
    CVB_REQUIRE(ev != nullptr, "null out pointer");
    cudaEvent_t e;
    CVB_CHECK_CUDA(cudaEventCreate(&e));
    *ev = e;
    return CVB_OK;
}
int cvb_event_destroy(void *ev)
{
    CVB_CHECK_CUDA(cudaEventDestroy((cudaEvent_t)ev));
    return CVB_OK;
}
int cvb_event_record(cvb_handle *h, void *ev)
{
    REQ_H(h);
    CVB_CHECK_CUDA(cudaEventRecord((cudaEvent_t)ev, h->stream));
    return CVB_OK;
}
int cvb_event_elapsed_ms(void *start, void *stop, float *ms)
{
    CVB_CHECK_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
    CVB_CHECK_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return CVB_OK;
}

void cvb_color_profile_default(cvb_color_profile *p)
{
    // the defaults of frame_enhancer.py:61-68
    p->contrast = 1.0; p->brightness = 0.0;
    p->hue_shift = 0.f; p->sat_scale = 1.f; p->val_scale = 1.f;
    p->radical_mode = 0; p->target_hue = 0.f; p->hue_window = 20.f;
    p->simd_block = 32;
}
void cvb_enhance_params_default(cvb_enhance_params *p)
{
    p->clahe_clip_limit = 3.0; p->tiles_x = 8; p->tiles_y = 8;
    p->bilateral_d = 9; p->sigma_color = 75.0; p->sigma_space = 75.0;
    p->use_color_profile = 0;
    cvb_color_profile_default(&p->profile);
}
void cvb_square_params_default(cvb_square_params *p)
{
    p->ops = CVB_SQ_PD_STATS; p->pd_blur = 5; p->cd_blur = 5;
    p->z_threshold = 2.5f; p->alpha = 0.1f; p->one_minus_alpha = (float)(1 - 0.1);
    p->initial_variance = 100.0f; p->min_variance = 10.0f;
}
void cvb_hough_params_default(cvb_hough_params *p)
{
    memset(p, 0, sizeof *p);
    p->dp = 1.2f; p->param1 = 100; p->param2 = 25;
    p->min_radius_ratio = 0.20; p->max_radius_ratio = 0.55;
    p->min_radius = 0; p->max_radius = 0; p->min_dist = 0; p->min_dist_div = 3;
}
void cvb_pipeline_params_default(cvb_pipeline_params *p)
{
    cvb_enhance_params_default(&p->enhance);
    cvb_square_params_default(&p->squares);
    p->warp_enhanced = 1; p->board_size = 620; p->rotate_180 = 0; p->reserved = 0;
}

int cvb_get_tables(uint16_t *gamma256, uint16_t *cbrt2048, int32_t *lab2yf512, uint8_t *invgamma4096, uint8_t *ltab2048)
{
    const CvbTables &t = cvb_host_tables();
    if (gamma256) memcpy(gamma256, t.gamma, sizeof t.gamma);
    if (cbrt2048) memcpy(cbrt2048, t.cbrt, sizeof t.cbrt);
    if (lab2yf512)      // public layout: (y, ify) pairs as in OpenCV's LabToYF_b; the kernels keep them packed
        for (int L = 0; L < 256; ++L) {
            lab2yf512[2 * L] = t.lab2yf[L] & 0xffff;
            lab2yf512[2 * L + 1] = (int32_t)((uint32_t)t.lab2yf[L] >> 16);
        }
    if (invgamma4096) memcpy(invgamma4096, t.invgamma, sizeof t.invgamma);
    if (ltab2048) memcpy(ltab2048, t.ltab, sizeof t.ltab);
    return CVB_OK;
}
int cvb_get_bilateral_tables(double sigma_color, double sigma_space, float *color768, float *space81)
{
    cvb_host_bilateral_tables(sigma_color, sigma_space, color768, space81);
    return CVB_OK;
}
int cvb_gaussian_kernel_q8(int ksize, int *q)
{
    CVB_REQUIRE(q != nullptr, "null out pointer");
    int rc = cvb_host_gaussian_q8(ksize, q);
    if (rc != CVB_OK) cvb_set_error("GaussianBlur ksize %d unsupported (odd, 1..31)", ksize);
    return rc;
}
int cvb_get_perspective_transform(const float *src_xy4, const float *dst_xy4, double *M9)
{
    CVB_REQUIRE(src_xy4 && dst_xy4 && M9, "null pointer");
    int rc = cvb_host_get_perspective(src_xy4, dst_xy4, M9);
    if (rc != CVB_OK) cvb_set_error("degenerate quadrilateral (singular system)");
    return rc;
}

// ---- stage-isolated entry points ----------------------------------------------------------
int cvb_color_profile_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_color_profile *p, uint8_t *out)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && out && p, "null pointer");
    CVB_REQUIRE(p->simd_block >= 0, "simd_block must be >= 0");
    return launch_color_profile(h, bgr, n, H, W, *p, out);
}
int cvb_bgr2lab_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, uint8_t *lab)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && lab, "null image pointer");
    return launch_bgr2lab(h, bgr, (long)n * H * W, lab);
}
int cvb_lab2bgr_dev(cvb_handle *h, const uint8_t *lab, int n, int H, int W, uint8_t *bgr)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && lab, "null image pointer");
    return launch_lab2bgr(h, lab, (long)n * H * W, bgr);
}

static int clahe_tables(cvb_handle *h, const uint8_t *src, int from_bgr, int n, int H, int W, const ClaheGeom &g,
                        int32_t **hist_io, uint8_t **lut_io, int32_t *minmax_init, uint8_t *lab_out = nullptr)
{
    const size_t nt = (size_t)g.tiles_x * g.tiles_y * n;
    int32_t *hist = *hist_io;
    uint8_t *lut = *lut_io;
    if (!hist) CVB_TRY(cvb_ws(h, h->ws_hist, nt * 256 * sizeof(int32_t), (void **)&hist));
    if (!lut) CVB_TRY(cvb_ws(h, h->ws_lut, nt * 256, (void **)&lut));
    CVB_TRY(launch_tile_hist(h, src, from_bgr, n, H, W, g, hist, minmax_init, lab_out));
    CVB_TRY(launch_clahe_lut(h, hist, n, g, lut));
    *hist_io = hist; *lut_io = lut;
    return CVB_OK;
}

int cvb_clahe_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W, double clip_limit, int tiles_x, int tiles_y,
                  uint8_t *out, int32_t *hist_out, uint8_t *lut_out)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(plane && out, "null image pointer");
    ClaheGeom g;
    CVB_TRY(cvb_clahe_geom(H, W, clip_limit, tiles_x, tiles_y, &g));
    CVB_TRY(clahe_tables(h, plane, 0, n, H, W, g, &hist_out, &lut_out, nullptr));
    return launch_clahe_apply_plane(h, plane, n, H, W, g, lut_out, out);
}
int cvb_correct_lighting_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, double clip_limit, int tiles_x,
                             int tiles_y, uint8_t *out, int32_t *hist_out, uint8_t *lut_out)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && out, "null image pointer");
    ClaheGeom g;
    CVB_TRY(cvb_clahe_geom(H, W, clip_limit, tiles_x, tiles_y, &g));
    CVB_TRY(clahe_tables(h, bgr, 1, n, H, W, g, &hist_out, &lut_out, nullptr));
    return launch_correct_lighting(h, bgr, n, H, W, g, lut_out, out);
}
int cvb_bilateral_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, int d, double sigma_color,
                      double sigma_space, uint8_t *out)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && out && bgr != out, "null or aliased image pointer");
    CVB_REQUIRE(d == 9, "bilateral diameter %d unsupported: the reference path uses d=9 (frame_enhancer.py:131)", d);
    return launch_fused(h, bgr, n, H, W, false, true, false, nullptr, nullptr, sigma_color, sigma_space, out, nullptr);
}
int cvb_sharpen_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, uint8_t *out)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && out && bgr != out, "null or aliased image pointer");
    return launch_fused(h, bgr, n, H, W, false, false, true, nullptr, nullptr, 0, 0, out, nullptr);
}
int cvb_normalize_dev(cvb_handle *h, const uint8_t *src, int n, int H, int W, int C, uint8_t *out, int32_t *minmax_out)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(src && out && C >= 1 && C <= 4, "null image pointer or bad channel count");
    if (!minmax_out) CVB_TRY(cvb_ws(h, h->ws_minmax, sizeof(int32_t) * 2 * n, (void **)&minmax_out));
    const long bytes = (long)H * W * C;
    CVB_TRY(launch_minmax(h, src, n, bytes, minmax_out));
    return launch_normalize(h, src, n, bytes, minmax_out, out);
}
int cvb_gray_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, uint8_t *gray)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && gray, "null image pointer");
    return launch_gray(h, bgr, (long)n * H * W, gray);
}
int cvb_gaussian_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W, int ksize, uint8_t *out)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(plane && out && plane != out, "null or aliased image pointer");
    return launch_gaussian(h, plane, n, H, W, ksize, 0.0, out);
}
int cvb_gaussian_sigma_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W, int ksize, double sigma, uint8_t *out)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(plane && out && plane != out, "null or aliased image pointer");
    return launch_gaussian(h, plane, n, H, W, ksize, sigma, out);
}

static int analysis_tail(cvb_handle *h, const uint8_t *src, const int32_t *minmax, int n, int H, int W,
                         uint8_t *enhanced, uint8_t *gray, uint8_t *binary, uint8_t *blurred, int32_t *otsu_t,
                         int32_t *hist)
{
    const long npx = (long)H * W;
    const bool want_bin = binary != nullptr || otsu_t != nullptr;
    if (want_bin) {
        if (!blurred) CVB_TRY(cvb_ws(h, h->ws_blur, (size_t)n * npx, (void **)&blurred));
        if (!hist) CVB_TRY(cvb_ws(h, h->ws_ohist, sizeof(int32_t) * 256 * n, (void **)&hist));
        if (!otsu_t) CVB_TRY(cvb_ws(h, h->ws_otsu, sizeof(int32_t) * n, (void **)&otsu_t));
    }
    CVB_TRY(launch_finish(h, src, n, H, W, minmax, enhanced, gray, blurred, want_bin ? hist : nullptr));
    if (want_bin) {
        // The Otsu scan is a chain of 256 dependent f64 divisions (~40 us whatever the batch).  For batches up to a chunk the
        // whole-path entry points fork it, with the mask, onto a second stream: the warp and the square kernel that
        // follow on the main stream need neither (they read the enhanced frame), join_tail() brings the streams together.
        const bool fork = h->fork_tail && n <= 64;     // single frames up to the chunks of the host-buffer pipeline
        cudaStream_t main_stream = h->stream;
        if (fork) {
            if (!h->aux_stream) {
                CVB_CHECK_CUDA(cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
                CVB_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
                CVB_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
            }
            CVB_CHECK_CUDA(cudaEventRecord(h->ev_fork, main_stream));
            CVB_CHECK_CUDA(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
            h->stream = h->aux_stream;
        }
        int rc = launch_otsu(h, hist, n, npx, otsu_t);
        if (rc == CVB_OK && binary) rc = launch_threshold(h, blurred, n, npx, otsu_t, binary);
        h->stream = main_stream;
        CVB_TRY(rc);
        if (fork) {
            CVB_CHECK_CUDA(cudaEventRecord(h->ev_join, h->aux_stream));
            h->tail_pending = true;
        }
    }
    return CVB_OK;
}

// the main stream waits for a forked analysis tail (no-op when nothing was forked)
static int join_tail(cvb_handle *h)
{
    if (h->tail_pending) {
        h->tail_pending = false;
        CVB_CHECK_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    }
    return CVB_OK;
}

int cvb_prepare_analysis_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, uint8_t *gray, uint8_t *binary,
                             uint8_t *blurred, int32_t *otsu_t, int32_t *hist)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr != nullptr, "null image pointer");
    return analysis_tail(h, bgr, nullptr, n, H, W, nullptr, gray, binary, blurred, otsu_t, hist);
}

static int check_enhance_params(const cvb_enhance_params *p)
{
    CVB_REQUIRE(p != nullptr, "null params");
    CVB_REQUIRE(p->bilateral_d == 9, "bilateral diameter %d unsupported: the reference path uses d=9", p->bilateral_d);
    return CVB_OK;
}

int cvb_process_pipeline_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_enhance_params *p,
                             uint8_t *enhanced)
{
    return cvb_enhance_dev(h, bgr, n, H, W, p, enhanced, nullptr, nullptr, nullptr);
}

int cvb_enhance_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_enhance_params *p,
                    uint8_t *enhanced, uint8_t *gray, uint8_t *binary, int32_t *otsu_t)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr != nullptr, "null image pointer");
    CVB_TRY(check_enhance_params(p));
    ClaheGeom g;
    CVB_TRY(cvb_clahe_geom(H, W, p->clahe_clip_limit, p->tiles_x, p->tiles_y, &g));
    const size_t fb = (size_t)H * W * 3;
    WS(ws_minmax, int32_t, 2 * n, minmax);
    WS(ws_sharp, uint8_t, fb * n, sharp);
    if (p->use_color_profile) {
        // step 0 of process_pipeline (frame_enhancer.py:167): one more pointwise pass, only when a profile is loaded
        WS(ws_prof, uint8_t, fb * n, prof);
        CVB_TRY(launch_color_profile(h, bgr, n, H, W, p->profile, prof));
        bgr = prof;
    }
    int32_t *hist = nullptr;
    uint8_t *lut = nullptr;
    // pass 1: tile histograms (+ min/max reset), LUTs
    // (the pass also stores the Lab pixels: pass 2 then skips RGB2Lab for every tile + halo pixel)
    WS(ws_lab, uint8_t, fb * n, lab);
    CVB_TRY(clahe_tables(h, bgr, 1, n, H, W, g, &hist, &lut, minmax, lab));
    // pass 2: lighting -> bilateral -> sharpen (+ min/max)
    CVB_TRY(launch_fused(h, lab, n, H, W, true, true, true, &g, lut, p->sigma_color, p->sigma_space, sharp, minmax, true));
    // pass 3/4: normalize -> gray -> blur -> Otsu -> mask
    if (!enhanced && !gray && !binary && !otsu_t) return CVB_OK;
    if (!gray && !binary && !otsu_t) return launch_normalize(h, sharp, n, (long)fb, minmax, enhanced);
    return analysis_tail(h, sharp, minmax, n, H, W, enhanced, gray, binary, nullptr, otsu_t, nullptr);
}

int cvb_enhance(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_enhance_params *p, uint8_t *enhanced,
                uint8_t *gray, uint8_t *binary, int32_t *otsu_t)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr != nullptr, "null image pointer");
    const size_t npx = (size_t)H * W, fb = npx * 3;
    WS(ws_in, uint8_t, fb * n, d_in);
    uint8_t *d_enh = nullptr, *d_gray = nullptr, *d_bin = nullptr;
    int32_t *d_otsu = nullptr;
    if (enhanced) CVB_TRY(cvb_ws(h, h->ws_enh, fb * n, (void **)&d_enh));
    if (gray) CVB_TRY(cvb_ws(h, h->ws_gray, npx * n, (void **)&d_gray));
    if (binary) CVB_TRY(cvb_ws(h, h->ws_bin, npx * n, (void **)&d_bin));
    if (binary || otsu_t) CVB_TRY(cvb_ws(h, h->ws_otsu, sizeof(int32_t) * n, (void **)&d_otsu));
    CVB_CHECK_CUDA(cudaMemcpyAsync(d_in, bgr, fb * n, cudaMemcpyHostToDevice, h->stream));
    CVB_TRY(cvb_enhance_dev(h, d_in, n, H, W, p, d_enh, d_gray, d_bin, d_otsu));
    if (enhanced) CVB_CHECK_CUDA(cudaMemcpyAsync(enhanced, d_enh, fb * n, cudaMemcpyDeviceToHost, h->stream));
    if (gray) CVB_CHECK_CUDA(cudaMemcpyAsync(gray, d_gray, npx * n, cudaMemcpyDeviceToHost, h->stream));
    if (binary) CVB_CHECK_CUDA(cudaMemcpyAsync(binary, d_bin, npx * n, cudaMemcpyDeviceToHost, h->stream));
    if (otsu_t) CVB_CHECK_CUDA(cudaMemcpyAsync(otsu_t, d_otsu, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    return CVB_OK;
}

// ---- warp ------------------------------------------------------------------------------------
static int upload_inverse_mats(cvb_handle *h, const double *M9, int n_mats, double **d_out)
{
    std::vector<double> &last = h->mat_cache;
    if (h->ws_mats.p && last.size() == (size_t)n_mats * 9 && memcmp(last.data(), M9, sizeof(double) * 9 * n_mats) == 0) {
        *d_out = (double *)h->ws_mats.p;      // same matrices as the previous call: already resident
        return CVB_OK;
    }
    std::vector<double> inv((size_t)n_mats * 9);
    for (int i = 0; i < n_mats; ++i)
        if (cvb_host_invert3(M9 + 9 * i, inv.data() + 9 * i) != CVB_OK) {
            cvb_set_error("perspective matrix %d is singular", i);
            return CVB_ERR_INVALID;
        }
    WS(ws_mats, double, (size_t)n_mats * 9, d_m);
    CVB_CHECK_CUDA(cudaMemcpyAsync(d_m, inv.data(), sizeof(double) * 9 * n_mats, cudaMemcpyHostToDevice, h->stream));
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));   // `inv` dies at return
    last.assign(M9, M9 + (size_t)n_mats * 9);
    *d_out = d_m;
    return CVB_OK;
}

int cvb_warp_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const double *M9, int n_mats, int out_h,
                 int out_w, uint8_t *warped)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && warped && M9, "null pointer");
    CVB_REQUIRE(n_mats == 1 || n_mats == n, "n_mats must be 1 or n");
    CVB_REQUIRE(out_h >= 1 && out_w >= 1 && out_h <= 32768 && out_w <= 32768, "bad output size");
    double *d_m = nullptr;
    CVB_TRY(upload_inverse_mats(h, M9, n_mats, &d_m));
    return launch_warp(h, bgr, n, H, W, d_m, n_mats, out_h, out_w, 0, warped);
}
int cvb_warp_rot180_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const double *M9, int n_mats, int out_h,
                        int out_w, uint8_t *warped)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && warped && M9, "null pointer");
    CVB_REQUIRE(n_mats == 1 || n_mats == n, "n_mats must be 1 or n");
    CVB_REQUIRE(out_h >= 1 && out_w >= 1 && out_h <= 32768 && out_w <= 32768, "bad output size");
    double *d_m = nullptr;
    CVB_TRY(upload_inverse_mats(h, M9, n_mats, &d_m));
    return launch_warp(h, bgr, n, H, W, d_m, n_mats, out_h, out_w, 1, warped);
}
int cvb_rotate_dev(cvb_handle *h, const uint8_t *src, int n, int H, int W, int C, int rotate_code, uint8_t *dst)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(src && dst && src != dst, "null pointer / in-place rotation");
    CVB_REQUIRE(C == 1 || C == 3, "rotate: 1 or 3 channels, got %d", C);
    CVB_REQUIRE(rotate_code >= 0 && rotate_code <= 2, "rotate code must be 0 (90 cw), 1 (180) or 2 (90 ccw)");
    return launch_rotate(h, src, n, H, W, C, rotate_code, dst);
}

int cvb_canny_dev(cvb_handle *h, const uint8_t *gray, int n, int H, int W, double low_thresh, double high_thresh,
                  uint8_t *edges)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(gray && edges, "null image pointer");
    return launch_canny(h, gray, n, H, W, low_thresh, high_thresh, edges);
}
int cvb_dilate_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W, int kw, int kh, int iterations, uint8_t *out)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(plane && out && plane != out, "null or aliased image pointer");
    CVB_REQUIRE(kw >= 1 && kh >= 1 && (kw & 1) && (kh & 1) && iterations >= 1, "dilate: odd kernel sizes and iterations >= 1");
    return launch_dilate(h, plane, n, H, W, kw, kh, iterations, out);
}
int cvb_contour_mask_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, uint8_t *mask)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && mask, "null image pointer");
    const size_t npx = (size_t)H * W;
    WS(ws_gray, uint8_t, npx * n, gray);
    WS(ws_blur, uint8_t, npx * n, blur);
    CVB_TRY(launch_gray(h, bgr, (long)npx * n, gray));
    CVB_TRY(launch_gaussian(h, gray, n, H, W, 7, 1.0, blur));
    CVB_TRY(launch_canny(h, blur, n, H, W, 30, 100, gray));
    return launch_dilate(h, gray, n, H, W, 5, 5, 3, mask);
}
int cvb_projections_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W, uint32_t *row_sums, uint32_t *col_sums)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(plane && row_sums && col_sums, "null pointer");
    return launch_projections(h, plane, n, H, W, row_sums, col_sums);
}

// ---- per-square ---------------------------------------------------------------------------------
int cvb_state_create(cvb_handle *h, int n_streams, int BH, int BW, cvb_state **out)
{
    REQ_H(h);
    CVB_REQUIRE(out && n_streams >= 1 && BH >= 1 && BW >= 1, "bad state shape");
    *out = nullptr;
    cvb_state *s = new cvb_state();
    s->h = h; s->n_streams = n_streams; s->BH = BH; s->BW = BW;
    const size_t px = (size_t)n_streams * BH * BW;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&s->pd_ref, px);
    if (e == cudaSuccess) e = cudaMalloc(&s->pd_cur, px);
    if (e == cudaSuccess) e = cudaMalloc(&s->flags, px);
    if (e == cudaSuccess) e = cudaMalloc(&s->cd_mean, px * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&s->cd_var, px * sizeof(float));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->flags, 0, px, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->pd_ref, 0, px, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->pd_cur, 0, px, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->cd_mean, 0, px * sizeof(float), h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(s->cd_var, 0, px * sizeof(float), h->stream);
    if (e != cudaSuccess) {
        cvb_set_error("state allocation failed: %s", cudaGetErrorString(e));
        cvb_state_destroy(s);
        return CVB_ERR_CUDA;
    }
    *out = s;
    return CVB_OK;
}
void cvb_state_destroy(cvb_state *s)
{
    if (!s) return;
    if (s->h) { cudaSetDevice(s->h->device); cudaStreamSynchronize(s->h->stream); }
    cudaFree(s->pd_ref); cudaFree(s->pd_cur); cudaFree(s->flags); cudaFree(s->cd_mean); cudaFree(s->cd_var);
    delete s;
}

static int state_plane(cvb_state *s, int stream, int plane, void **p, size_t *bytes)
{
    CVB_REQUIRE(s != nullptr, "null state");
    CVB_REQUIRE(stream >= 0 && stream < s->n_streams, "stream slot %d out of range", stream);
    const size_t px = (size_t)s->BH * s->BW, o = (size_t)stream * px;
    switch (plane) {
    case 0: *p = s->pd_ref + o; *bytes = px; break;
    case 1: *p = s->cd_mean + o; *bytes = px * 4; break;
    case 2: *p = s->cd_var + o; *bytes = px * 4; break;
    case 3: *p = s->pd_cur + o; *bytes = px; break;
    case 4: *p = s->flags + o; *bytes = px; break;
    default: cvb_set_error("unknown state plane %d", plane); return CVB_ERR_INVALID;
    }
    return CVB_OK;
}
int cvb_state_get(cvb_handle *h, cvb_state *s, int stream, int plane, void *host_out)
{
    REQ_H(h);
    void *p; size_t bytes;
    CVB_TRY(state_plane(s, stream, plane, &p, &bytes));
    CVB_CHECK_CUDA(cudaMemcpyAsync(host_out, p, bytes, cudaMemcpyDeviceToHost, h->stream));
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    return CVB_OK;
}
int cvb_state_set(cvb_handle *h, cvb_state *s, int stream, int plane, const void *host_in)
{
    REQ_H(h);
    void *p; size_t bytes;
    CVB_TRY(state_plane(s, stream, plane, &p, &bytes));
    CVB_CHECK_CUDA(cudaMemcpyAsync(p, host_in, bytes, cudaMemcpyHostToDevice, h->stream));
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    return CVB_OK;
}
int cvb_state_reset(cvb_handle *h, cvb_state *s, int stream)
{
    REQ_H(h);
    CVB_REQUIRE(s != nullptr && stream < s->n_streams, "bad state / stream");
    return launch_state_reset(h, s, stream);
}

// rects/select -> device, masks per distinct shape (cached on the handle by content)
static int stage_rects(cvb_handle *h, const cvb_rect *rects, int n_sq, const uint8_t *select, int BH, int BW,
                       cvb_rect **d_rects, int32_t **d_ofs, uint8_t **d_masks, uint8_t **d_select, int *max_px)
{
    CVB_REQUIRE(rects != nullptr && n_sq >= 1 && n_sq <= 65535, "bad rect list");
    for (int i = 0; i < n_sq; ++i)
        CVB_REQUIRE(rects[i].w >= 1 && rects[i].h >= 1 && rects[i].x >= 0 && rects[i].y >= 0 &&
                        rects[i].x + rects[i].w <= BW && rects[i].y + rects[i].h <= BH,
                    "square %d (%d,%d %dx%d) outside the %dx%d board", i, rects[i].x, rects[i].y, rects[i].w,
                    rects[i].h, BW, BH);
    RectCache &c = h->rect_cache;
    const bool same_rects = (int)c.rects.size() == n_sq && memcmp(c.rects.data(), rects, sizeof(cvb_rect) * n_sq) == 0 &&
                            h->ws_rects.p != nullptr;
    const bool same_sel = same_rects && c.has_select == (select != nullptr) &&
                          (!select || memcmp(c.select.data(), select, n_sq) == 0);
    const size_t ofs_bytes = sizeof(int32_t) * n_sq, rect_bytes = sizeof(cvb_rect) * n_sq;
    if (!same_rects) {
        std::map<int, int32_t> shape_ofs;
        std::vector<int32_t> ofs(n_sq);
        std::vector<uint8_t> masks;
        int mx = 0;
        for (int i = 0; i < n_sq; ++i) {
            const int key = (rects[i].h << 16) | rects[i].w;
            auto it = shape_ofs.find(key);
            if (it == shape_ofs.end()) {
                const size_t o = masks.size();
                masks.resize(o + (size_t)rects[i].h * rects[i].w);
                cvb_host_square_masks(rects[i].h, rects[i].w, masks.data() + o);
                it = shape_ofs.emplace(key, (int32_t)o).first;
            }
            ofs[i] = it->second;
            mx = std::max(mx, rects[i].h * rects[i].w);
        }
        uint8_t *d_r = nullptr, *d_m = nullptr;
        CVB_TRY(cvb_ws(h, h->ws_rects, rect_bytes + ofs_bytes, (void **)&d_r));
        CVB_TRY(cvb_ws(h, h->ws_masks, masks.size(), (void **)&d_m));
        CVB_CHECK_CUDA(cudaMemcpyAsync(d_r, rects, rect_bytes, cudaMemcpyHostToDevice, h->stream));
        CVB_CHECK_CUDA(cudaMemcpyAsync(d_r + rect_bytes, ofs.data(), ofs_bytes, cudaMemcpyHostToDevice, h->stream));
        CVB_CHECK_CUDA(cudaMemcpyAsync(d_m, masks.data(), masks.size(), cudaMemcpyHostToDevice, h->stream));
        CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
        c.rects.assign(rects, rects + n_sq);
        c.max_px = mx;
        c.mask_bytes = masks.size();
    }
    if (!same_sel) {
        c.has_select = select != nullptr;
        if (select) {
            uint8_t *d_s = nullptr;
            CVB_TRY(cvb_ws(h, h->ws_select, n_sq, (void **)&d_s));
            CVB_CHECK_CUDA(cudaMemcpyAsync(d_s, select, n_sq, cudaMemcpyHostToDevice, h->stream));
            CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
            c.select.assign(select, select + n_sq);
        }
    }
    *d_rects = (cvb_rect *)h->ws_rects.p;
    *d_ofs = (int32_t *)((uint8_t *)h->ws_rects.p + rect_bytes);
    *d_masks = (uint8_t *)h->ws_masks.p;
    *d_select = select ? (uint8_t *)h->ws_select.p : nullptr;
    *max_px = c.max_px;
    return CVB_OK;
}

static int squares_impl(cvb_handle *h, const uint8_t *boards, int n, int BH, int BW, int C, const cvb_rect *rects,
                        int n_sq, const uint8_t *select, cvb_state *state, int stream0, const cvb_square_params *p,
                        cvb_square_stats *stats)
{
    CVB_REQUIRE(boards != nullptr && p != nullptr, "null pointer");
    CVB_REQUIRE(C == 1 || C == 3, "boards must have 1 or 3 channels, got %d", C);
    CVB_REQUIRE(n >= 1 && n <= 65535 && BH >= 1 && BW >= 1, "bad board batch");
    const int state_ops = CVB_SQ_PD_SET_REF | CVB_SQ_CD_CALIBRATE | CVB_SQ_CD_DETECT | CVB_SQ_CD_UPDATE;
    if (p->ops & state_ops) {
        if (!state) { cvb_set_error("ops 0x%x need a cvb_state", p->ops); return CVB_ERR_STATE; }
    }
    if (state) {
        if (state->BH != BH || state->BW != BW) {
            cvb_set_error("state is %dx%d but boards are %dx%d", state->BH, state->BW, BH, BW);
            return CVB_ERR_STATE;
        }
        if (stream0 < 0 || stream0 + n > state->n_streams) {
            cvb_set_error("stream slots %d..%d outside the state's %d", stream0, stream0 + n - 1, state->n_streams);
            return CVB_ERR_STATE;
        }
    }
    int pd_q[31] = {0}, cd_q[31] = {0};
    if (cvb_host_gaussian_q8(p->pd_blur, pd_q) != CVB_OK || cvb_host_gaussian_q8(p->cd_blur, cd_q) != CVB_OK) {
        cvb_set_error("blur kernels must be odd and <= 31 (pd %d, cd %d)", p->pd_blur, p->cd_blur);
        return CVB_ERR_INVALID;
    }
    cvb_rect *d_rects; int32_t *d_ofs; uint8_t *d_masks, *d_select; int max_px;
    CVB_TRY(stage_rects(h, rects, n_sq, select, BH, BW, &d_rects, &d_ofs, &d_masks, &d_select, &max_px));
    return launch_squares(h, boards, n, BH, BW, C, d_rects, d_ofs, d_masks, n_sq, max_px, d_select, state, stream0, *p,
                          pd_q, cd_q, stats);
}

int cvb_squares_dev(cvb_handle *h, const uint8_t *boards, int n, int BH, int BW, int C, const cvb_rect *rects, int n_sq,
                    const uint8_t *select, cvb_state *state, int stream0, const cvb_square_params *p,
                    cvb_square_stats *stats)
{
    REQ_H(h);
    return squares_impl(h, boards, n, BH, BW, C, rects, n_sq, select, state, stream0, p, stats);
}

// ---- Hough circles per square ---------------------------------------------------------------------
static int hough_check(const cvb_hough_params *p)
{
    CVB_REQUIRE(p != nullptr, "null Hough params");
    CVB_REQUIRE(p->dp > 0 && p->param1 > 0 && p->param2 > 0, "dp, param1 and param2 must be positive");
    CVB_REQUIRE(p->min_dist_div > 0 || p->min_dist > 0, "minDist must be positive");
    return CVB_OK;
}
int cvb_hough_geometry(const cvb_rect *rects, int n_sq, const cvb_hough_params *p, cvb_hough_square *out)
{
    CVB_TRY(hough_check(p));
    CVB_REQUIRE(rects && out && n_sq >= 1, "bad rect list");
    for (int i = 0; i < n_sq; ++i) {
        CVB_REQUIRE(rects[i].w >= 1 && rects[i].h >= 1, "empty square %d", i);
        CVB_REQUIRE(cvb_host_hough_square(rects[i], *p, out + i) == CVB_OK, "minDist of square %d is not positive", i);
    }
    return CVB_OK;
}
static int hough_impl(cvb_handle *h, const uint8_t *planes, int n, int PH, int PW, const cvb_rect *rects, int n_sq,
                      const uint8_t *select, const cvb_hough_params *p, cvb_hough_result *d_results)
{
    REQ_H(h); REQ_IMG(n, PH, PW);
    CVB_REQUIRE(planes && rects && d_results, "null pointer");
    CVB_REQUIRE(n_sq >= 1 && n_sq <= 65535, "bad rect list");
    for (int i = 0; i < n_sq; ++i) {
        CVB_REQUIRE(rects[i].w >= 1 && rects[i].h >= 1 && rects[i].x >= 0 && rects[i].y >= 0 &&
                        rects[i].x + rects[i].w <= PW && rects[i].y + rects[i].h <= PH,
                    "square %d (%d,%d %dx%d) outside the %dx%d plane", i, rects[i].x, rects[i].y, rects[i].w, rects[i].h, PW, PH);
        bool used = select == nullptr;
        for (int f = 0; f < n && !used; ++f) used = select[(size_t)f * n_sq + i] != 0;
        // the size limit only concerns squares some frame selects (an unselected large rectangle is harmless)
        CVB_REQUIRE(!used || (rects[i].w <= CVB_HOUGH_MAX_DIM_GLOBAL && rects[i].h <= CVB_HOUGH_MAX_DIM_GLOBAL),
                    "square %d is %dx%d: the Hough kernel handles squares up to %d pixels a side", i, rects[i].w, rects[i].h,
                    CVB_HOUGH_MAX_DIM_GLOBAL);
    }
    std::vector<cvb_hough_square> sq(n_sq);
    CVB_TRY(cvb_hough_geometry(rects, n_sq, p, sq.data()));
    for (int i = 0; i < n_sq; ++i)      // accumulator cells are addressed by 8-bit coordinates and 16-bit indices in the kernel
        CVB_REQUIRE(sq[i].acc_rows <= 253 && sq[i].acc_cols <= 253,
                    "square %d: a %dx%d accumulator (dp %.2f) exceeds 253 cells a side", i, sq[i].acc_cols, sq[i].acc_rows, (double)p->dp);
    WS(ws_hough_sq, cvb_hough_square, n_sq, d_sq);
    if (h->hough_cache.size() != sq.size() || memcmp(h->hough_cache.data(), sq.data(), sizeof(cvb_hough_square) * n_sq) != 0) {
        CVB_CHECK_CUDA(cudaMemcpyAsync(d_sq, sq.data(), sizeof(cvb_hough_square) * n_sq, cudaMemcpyHostToDevice, h->stream));
        CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
        h->hough_cache = sq;
    }
    uint8_t *d_sel = nullptr;
    if (select) {
        CVB_TRY(cvb_ws(h, h->ws_hough_sel, (size_t)n * n_sq, (void **)&d_sel));
        CVB_CHECK_CUDA(cudaMemcpyAsync(d_sel, select, (size_t)n * n_sq, cudaMemcpyHostToDevice, h->stream));
    }
    return launch_hough(h, planes, n, (size_t)PH * PW, PW, d_sq, sq.data(), n_sq, d_sel, *p, d_results);
}
int cvb_hough_dev(cvb_handle *h, const uint8_t *planes, int n, int PH, int PW, const cvb_rect *rects, int n_sq,
                  const uint8_t *select, const cvb_hough_params *p, cvb_hough_result *results)
{
    return hough_impl(h, planes, n, PH, PW, rects, n_sq, select, p, results);
}
int cvb_hough_state(cvb_handle *h, cvb_state *state, int stream0, int n, const cvb_rect *rects, int n_sq,
                    const uint8_t *select, const cvb_hough_params *p, cvb_hough_result *results_host)
{
    REQ_H(h);
    CVB_REQUIRE(state != nullptr && stream0 >= 0 && n >= 1 && stream0 + n <= state->n_streams, "bad state / stream range");
    CVB_REQUIRE(results_host != nullptr, "null pointer");
    WS(ws_hough_res, cvb_hough_result, (size_t)n * n_sq, d_res);
    const size_t plane = (size_t)state->BH * state->BW;
    CVB_TRY(hough_impl(h, state->pd_cur + plane * stream0, n, state->BH, state->BW, rects, n_sq, select, p, d_res));
    CVB_CHECK_CUDA(cudaMemcpyAsync(results_host, d_res, sizeof(cvb_hough_result) * (size_t)n * n_sq, cudaMemcpyDeviceToHost,
                                   h->stream));
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    return CVB_OK;
}

// ---- whole path ----------------------------------------------------------------------------------
// kernels only: matrices (inverse, device) and rectangles are already staged
static int pipeline_launch(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_pipeline_params *p,
                           const double *d_minv, int n_mats, const cvb_rect *rects, int n_sq, const uint8_t *select,
                           cvb_state *state, int stream0, uint8_t *enhanced, uint8_t *gray, uint8_t *binary,
                           int32_t *otsu_t, uint8_t *warped, cvb_square_stats *stats)
{
    const int S = p->board_size;
    const size_t npx = (size_t)H * W, fb = npx * 3;
    if (!enhanced) CVB_TRY(cvb_ws(h, h->ws_enh, fb * n, (void **)&enhanced));
    if (!gray) CVB_TRY(cvb_ws(h, h->ws_gray, npx * n, (void **)&gray));
    if (!binary) CVB_TRY(cvb_ws(h, h->ws_bin, npx * n, (void **)&binary));
    if (!otsu_t) CVB_TRY(cvb_ws(h, h->ws_otsu, sizeof(int32_t) * n, (void **)&otsu_t));
    if (!warped) CVB_TRY(cvb_ws(h, h->ws_warp, (size_t)S * S * 3 * n, (void **)&warped));
    h->fork_tail = true;
    int rc = cvb_enhance_dev(h, bgr, n, H, W, &p->enhance, enhanced, gray, binary, otsu_t);
    h->fork_tail = false;
    if (rc == CVB_OK)
        rc = launch_warp(h, p->warp_enhanced ? enhanced : bgr, n, H, W, d_minv, n_mats, S, S, p->rotate_180 != 0, warped);
    if (rc == CVB_OK) rc = squares_impl(h, warped, n, S, S, 3, rects, n_sq, select, state, stream0, &p->squares, stats);
    const int rj = join_tail(h);
    return rc != CVB_OK ? rc : rj;
}

static int pipeline_check(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_pipeline_params *p,
                          const double *M9, int n_mats, const cvb_rect *rects)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr && p && M9 && rects, "null pointer");
    CVB_REQUIRE(n_mats == 1 || n_mats == n, "n_mats must be 1 or n");
    CVB_REQUIRE(p->board_size >= 8 && p->board_size <= 8192, "bad board_size %d", p->board_size);
    return CVB_OK;
}

int cvb_pipeline_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_pipeline_params *p,
                     const double *M9, int n_mats, const cvb_rect *rects, int n_sq, const uint8_t *select,
                     cvb_state *state, int stream0, uint8_t *enhanced, uint8_t *gray, uint8_t *binary, int32_t *otsu_t,
                     uint8_t *warped, cvb_square_stats *stats)
{
    CVB_TRY(pipeline_check(h, bgr, n, H, W, p, M9, n_mats, rects));
    const int S = p->board_size;
    // small host->device staging first (it synchronises), then only kernel launches
    double *d_m = nullptr;
    CVB_TRY(upload_inverse_mats(h, M9, n_mats, &d_m));
    {
        cvb_rect *d_rects; int32_t *d_ofs; uint8_t *d_masks, *d_select; int max_px;
        CVB_TRY(stage_rects(h, rects, n_sq, select, S, S, &d_rects, &d_ofs, &d_masks, &d_select, &max_px));
    }
    return pipeline_launch(h, bgr, n, H, W, p, d_m, n_mats, rects, n_sq, select, state, stream0, enhanced, gray, binary,
                           otsu_t, warped, stats);
}

// ---- board overlay ---------------------------------------------------------------------------
// half widths of the rows dy = 0..r of a filled circle: the spans drawing.cpp's Circle() fills (midpoint algorithm)
static void circle_half_widths(int r, uint16_t *half)
{
    for (int i = 0; i <= r; ++i) half[i] = 0;
    int err = 0, dx = r, dy = 0, plus = 1, minus = (r << 1) - 1;
    while (dx >= dy) {
        if (dx > half[dy]) half[dy] = (uint16_t)dx;      // rows +-dy are filled over +-dx
        if (dy > half[dx]) half[dx] = (uint16_t)dy;      // rows +-dx over +-dy
        ++dy; err += plus; plus += 2;
        const int mask = (err <= 0) - 1;
        err -= minus & mask; dx += mask; minus -= mask & 2;
    }
}

int cvb_overlay_dev(cvb_handle *h, uint8_t *bgr, int n, int H, int W, const cvb_overlay_op *ops, int n_ops,
                    const uint8_t *masks, size_t mask_bytes)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(bgr, "null image pointer");
    CVB_REQUIRE(n_ops >= 0 && (n_ops == 0 || ops), "null display list");
    if (n_ops == 0) return CVB_OK;
    // validate, lay out [ops][circle tables][masks] in one staging vector
    std::vector<cvb_overlay_op> list(ops, ops + n_ops);
    std::map<int, uint32_t> circle_ofs;
    std::vector<uint16_t> tables;
    for (int i = 0; i < n_ops; ++i) {
        cvb_overlay_op &o = list[i];
        CVB_REQUIRE(o.kind >= CVB_OV_RECT && o.kind <= CVB_OV_STAMP, "overlay op %d: unknown kind %d", i, o.kind);
        CVB_REQUIRE(o.alpha == o.alpha && o.beta == o.beta, "overlay op %d: NaN weight", i);
        const int lim = 1 << 20;
        CVB_REQUIRE(o.x0 > -lim && o.x0 < lim && o.y0 > -lim && o.y0 < lim && o.x1 > -lim && o.x1 < lim && o.y1 > -lim && o.y1 < lim,
                    "overlay op %d: coordinate out of range", i);
        if (o.kind == CVB_OV_RECT) {                       // cv2.rectangle / cv2.line accept the corners in any order
            if (o.x0 > o.x1) std::swap(o.x0, o.x1);
            if (o.y0 > o.y1) std::swap(o.y0, o.y1);
        } else if (o.kind == CVB_OV_CIRCLE) {
            CVB_REQUIRE(o.x1 >= 0 && o.x1 <= 16384, "overlay op %d: circle radius %d", i, o.x1);
            auto it = circle_ofs.find(o.x1);
            if (it == circle_ofs.end()) {
                it = circle_ofs.emplace(o.x1, (uint32_t)(tables.size() * sizeof(uint16_t))).first;
                tables.resize(tables.size() + o.x1 + 1);
                circle_half_widths(o.x1, tables.data() + it->second / sizeof(uint16_t));
            }
            o.aux_ofs = it->second;
        } else {
            CVB_REQUIRE(o.x1 >= 1 && o.y1 >= 1 && o.x1 <= 32768 && o.y1 <= 32768, "overlay op %d: stamp size %d x %d", i, o.x1, o.y1);
            const size_t need = (size_t)o.aux_ofs + (size_t)((o.x1 + 7) >> 3) * o.y1;
            CVB_REQUIRE(masks && need <= mask_bytes, "overlay op %d: stamp mask outside the %zu mask bytes", i, mask_bytes);
        }
        if (o.group != 0 && i > 0 && list[i - 1].group == o.group)
            CVB_REQUIRE(memcmp(o.color, list[i - 1].color, 3) == 0 && o.alpha == list[i - 1].alpha && o.beta == list[i - 1].beta,
                        "overlay op %d: ops of one group share color, alpha and beta", i);
    }
    for (int i = 0; i < n_ops; ++i)                       // a group id is not reused by a later, separate run of ops
        if (list[i].group != 0 && i > 0 && list[i - 1].group != list[i].group)
            for (int j = 0; j < i - 1; ++j)
                CVB_REQUIRE(list[j].group != list[i].group, "overlay op %d: group %d is not consecutive", i, list[i].group);
    const size_t ops_bytes = sizeof(cvb_overlay_op) * (size_t)n_ops;
    const size_t tab_bytes = (tables.size() * sizeof(uint16_t) + 15) & ~(size_t)15;
    for (auto &o : list) if (o.kind == CVB_OV_STAMP) o.aux_ofs += (uint32_t)tab_bytes;
    std::vector<uint8_t> stage(ops_bytes + tab_bytes + mask_bytes);
    memcpy(stage.data(), list.data(), ops_bytes);
    if (!tables.empty()) memcpy(stage.data() + ops_bytes, tables.data(), tables.size() * sizeof(uint16_t));
    if (mask_bytes) memcpy(stage.data() + ops_bytes + tab_bytes, masks, mask_bytes);
    WS(ws_overlay, uint8_t, stage.size(), d);
    CVB_CHECK_CUDA(cudaMemcpyAsync(d, stage.data(), stage.size(), cudaMemcpyHostToDevice, h->stream));
    CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));     // `stage` dies at return
    const cvb_overlay_op *d_ops = reinterpret_cast<const cvb_overlay_op *>(d);
    const int cap = cvb_overlay_max_ops();
    for (int first = 0; first < n_ops;) {                 // longer lists: slices that do not cut a group
        int end = std::min(n_ops, first + cap);
        while (end < n_ops && end > first && list[end].group != 0 && list[end].group == list[end - 1].group) --end;
        CVB_REQUIRE(end > first, "overlay: a group of more than %d ops", cap);
        CVB_TRY(launch_overlay(h, bgr, n, H, W, d_ops + first, end - first, d + ops_bytes));
        first = end;
    }
    return CVB_OK;
}

size_t cvb_frame_bytes(int format, int H, int W) { return cvb_host_frame_bytes(format, H, W); }

int cvb_cvt_to_bgr_dev(cvb_handle *h, const uint8_t *src, int format, int n, int H, int W, uint8_t *bgr)
{
    REQ_H(h); REQ_IMG(n, H, W);
    CVB_REQUIRE(src && bgr && src != bgr, "null or aliased image pointer");
    if (format == CVB_FMT_BGR) {
        CVB_CHECK_CUDA(cudaMemcpyAsync(bgr, src, (size_t)n * H * W * 3, cudaMemcpyDeviceToDevice, h->stream));
        return CVB_OK;
    }
    return launch_yuv_to_bgr(h, src, format, n, H, W, bgr);
}

// Host-buffer variant.  The batch is cut into chunks; chunk k+1 is copied host->device on a second
// stream while chunk k is being processed (two input buffers), so the PCIe copy and the kernels
// overlap when the host memory is page-locked.  Frames arrive in `format`; YUV frames are converted
// to BGR on the device (cvb_ingest.cu), so only their 2 or 1.5 bytes per pixel cross PCIe.
// wait == false: everything is enqueued (copies in, kernels, copies out) and the call returns; cvb_pipeline_wait joins.
// The staging buffers (ws_stage) belong to this function alone, so the only hazards on them are its own earlier chunks,
// of this call or of the previous one: the per-buffer ev_done events.  The first copy of a call therefore overlaps the
// kernels of the previous call.
static int pipeline_fmt_impl(cvb_handle *h, const uint8_t *frames, int format, int n, int H, int W, const cvb_pipeline_params *p,
                             const double *M9, int n_mats, const cvb_rect *rects, int n_sq, const uint8_t *select,
                             cvb_state *state, int stream0, int32_t *otsu_t, cvb_square_stats *stats, bool wait)
{
    CVB_TRY(pipeline_check(h, frames, n, H, W, p, M9, n_mats, rects));
    CVB_REQUIRE(n_sq >= 1, "no squares");
    CVB_REQUIRE(format == CVB_FMT_BGR || format == CVB_FMT_YUY2 || format == CVB_FMT_NV12, "unknown frame format %d", format);
    CVB_REQUIRE(format == CVB_FMT_BGR || (W % 2 == 0 && (format != CVB_FMT_NV12 || H % 2 == 0)),
                "YUV frames need an even width (NV12: and height), got %dx%d", W, H);
    const int S = p->board_size;
    const size_t fb = (size_t)H * W * 3, fin = cvb_host_frame_bytes(format, H, W);
    // Frames per chunk.  Measured on B200 at 1080p (tools/chunk_sweep.py, profiles/r02_notes.md): BGR input is bound by
    // the PCIe copy and wants small chunks (the first chunk's copy is the only one not overlapped: 4..13 frames equal);
    // YUV input is bound by the kernels, which want chunks whose fused-kernel grid fills whole waves of 2 CTAs per SM
    // (25 frames at 1080p: 6800 CTAs = 22.97 waves; 8 frames = 7.35 waves lose 8 %).
    int chunk = h->chunk_frames;
    if (chunk <= 0) {
        // BGR: the copy is the bound, the first chunk's copy the only one nothing hides: about 50 MB per chunk (8 frames
        // of 1080p, 2 of 3840 x 2160: a batch of eight 4K streams in ONE chunk would not overlap copy and kernels at all)
        chunk = (int)std::max<size_t>(1, std::min<size_t>(8, ((size_t)50 << 20) / fin));
        if (format != CVB_FMT_BGR && fin * 12 <= ((size_t)100 << 20)) {
            const long tiles = (long)((W + 119) / 120) * ((H + 63) / 64), slots = 2L * h->sm_count;
            double best = 0;
            for (int c = 12; c <= 32; ++c) {
                const long ctas = tiles * c, waves = (ctas + slots - 1) / slots;
                const double eff = (double)ctas / (double)(waves * slots);
                if (eff >= best) { best = eff; chunk = c; }
            }
        }
    }
    chunk = std::max(1, std::min(n, chunk));
    if (!h->copy_stream) {
        CVB_CHECK_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CVB_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_copy[i], cudaEventDisableTiming));
            CVB_CHECK_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
        }
    }
    // two staging buffers of one chunk in the arrival format; YUV chunks are converted into one BGR buffer
    uint8_t *d_in = nullptr, *d_bgr = nullptr;
    CVB_TRY(cvb_ws(h, h->ws_stage, fin * chunk * 2, (void **)&d_in));
    if (format != CVB_FMT_BGR) CVB_TRY(cvb_ws(h, h->ws_in, fb * chunk, (void **)&d_bgr));
    WS(ws_stats, cvb_square_stats, (size_t)n * n_sq, d_stats);
    WS(ws_otsu_all, int32_t, n, d_otsu);
    double *d_m = nullptr;
    CVB_TRY(upload_inverse_mats(h, M9, n_mats, &d_m));
    {
        cvb_rect *d_rects; int32_t *d_ofs; uint8_t *d_masks, *d_select; int max_px;
        CVB_TRY(stage_rects(h, rects, n_sq, select, S, S, &d_rects, &d_ofs, &d_masks, &d_select, &max_px));
    }
    // workspaces are sized here once, so no chunk triggers a synchronising reallocation
    {
        const size_t npx = (size_t)H * W;
        void *t;
        CVB_TRY(cvb_ws(h, h->ws_enh, fb * chunk, &t)); CVB_TRY(cvb_ws(h, h->ws_gray, npx * chunk, &t));
        CVB_TRY(cvb_ws(h, h->ws_bin, npx * chunk, &t)); CVB_TRY(cvb_ws(h, h->ws_warp, (size_t)S * S * 3 * chunk, &t));
        CVB_TRY(cvb_ws(h, h->ws_sharp, fb * chunk, &t)); CVB_TRY(cvb_ws(h, h->ws_blur, npx * chunk, &t));
        CVB_TRY(cvb_ws(h, h->ws_lab, fb * chunk, &t));
    }
    if (h->stage_layout != fin * chunk) {
        // other buffer boundaries than the previous call used: both buffers are free only when all earlier kernels are
        for (int i = 0; i < 2; ++i) CVB_CHECK_CUDA(cudaEventRecord(h->ev_done[i], h->stream));
        h->stage_layout = fin * chunk;
    }
    int k = 0;
    for (int f0 = 0; f0 < n; f0 += chunk, ++k) {
        const int cnt = std::min(chunk, n - f0), b = k & 1;
        uint8_t *buf = d_in + (size_t)b * chunk * fin;
        // buffer b is free once the last kernel that read it is done: an earlier chunk of this call or of the previous
        // one (waiting on an event that was never recorded returns at once)
        CVB_CHECK_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_done[b], 0));
        CVB_CHECK_CUDA(cudaMemcpyAsync(buf, frames + (size_t)f0 * fin, fin * cnt, cudaMemcpyHostToDevice, h->copy_stream));
        CVB_CHECK_CUDA(cudaEventRecord(h->ev_copy[b], h->copy_stream));
        CVB_CHECK_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copy[b], 0));
        const uint8_t *bgr = buf;
        if (format != CVB_FMT_BGR) {
            CVB_TRY(launch_yuv_to_bgr(h, buf, format, cnt, H, W, d_bgr));
            // the staging buffer is free again as soon as the conversion has read it
            CVB_CHECK_CUDA(cudaEventRecord(h->ev_done[b], h->stream));
            bgr = d_bgr;
        }
        CVB_TRY(pipeline_launch(h, bgr, cnt, H, W, p, n_mats == 1 ? d_m : d_m + (size_t)9 * f0, n_mats == 1 ? 1 : cnt, rects,
                                n_sq, select, state, stream0 + f0, nullptr, nullptr, nullptr, d_otsu + f0, nullptr,
                                d_stats + (size_t)f0 * n_sq));
        if (format == CVB_FMT_BGR) CVB_CHECK_CUDA(cudaEventRecord(h->ev_done[b], h->stream));
    }
    if (stats)
        CVB_CHECK_CUDA(cudaMemcpyAsync(stats, d_stats, sizeof(cvb_square_stats) * (size_t)n * n_sq,
                                       cudaMemcpyDeviceToHost, h->stream));
    if (otsu_t) CVB_CHECK_CUDA(cudaMemcpyAsync(otsu_t, d_otsu, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, h->stream));
    if (wait) CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    return CVB_OK;
}

int cvb_pipeline_fmt(cvb_handle *h, const uint8_t *frames, int format, int n, int H, int W, const cvb_pipeline_params *p,
                     const double *M9, int n_mats, const cvb_rect *rects, int n_sq, const uint8_t *select, cvb_state *state,
                     int stream0, int32_t *otsu_t, cvb_square_stats *stats)
{
    return pipeline_fmt_impl(h, frames, format, n, H, W, p, M9, n_mats, rects, n_sq, select, state, stream0, otsu_t, stats, true);
}

int cvb_pipeline_submit(cvb_handle *h, const uint8_t *frames, int format, int n, int H, int W, const cvb_pipeline_params *p,
                        const double *M9, int n_mats, const cvb_rect *rects, int n_sq, const uint8_t *select, cvb_state *state,
                        int stream0, int32_t *otsu_t, cvb_square_stats *stats, uint64_t *ticket)
{
    CVB_TRY(pipeline_fmt_impl(h, frames, format, n, H, W, p, M9, n_mats, rects, n_sq, select, state, stream0, otsu_t, stats, false));
    const uint64_t seq = ++h->tickets;
    cudaEvent_t &ev = h->ev_ticket[seq % CVB_TICKET_RING];
    if (!ev) CVB_CHECK_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CVB_CHECK_CUDA(cudaEventRecord(ev, h->stream));
    if (ticket) *ticket = seq;
    return CVB_OK;
}

int cvb_pipeline_wait(cvb_handle *h, uint64_t ticket)
{
    REQ_H(h);
    CVB_REQUIRE(ticket <= h->tickets, "ticket %llu was never issued by this handle", (unsigned long long)ticket);
    // a slot of the ring that a later submission re-used marks a later point of the same stream: waiting there is enough
    if (ticket == 0 || h->tickets - ticket >= CVB_TICKET_RING) CVB_CHECK_CUDA(cudaStreamSynchronize(h->stream));
    else CVB_CHECK_CUDA(cudaEventSynchronize(h->ev_ticket[ticket % CVB_TICKET_RING]));
    return CVB_OK;
}

int cvb_pipeline(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, const cvb_pipeline_params *p, const double *M9,
                 int n_mats, const cvb_rect *rects, int n_sq, const uint8_t *select, cvb_state *state, int stream0,
                 int32_t *otsu_t, cvb_square_stats *stats)
{
    return cvb_pipeline_fmt(h, bgr, CVB_FMT_BGR, n, H, W, p, M9, n_mats, rects, n_sq, select, state, stream0, otsu_t, stats);
}

// sum of the index-check counters of the debug build (cvb_device.cuh); -1 in the release build, which has no checks
long long cvb_debug_bounds_violations(cvb_handle *h, int *first_line)
{
    if (first_line) *first_line = 0;
#ifdef CVB_DEBUG_BOUNDS
    if (!h || cudaSetDevice(h->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return -2;
    void (*tus[])(unsigned long long *, int *) = {cvb_bounds_enhance, cvb_bounds_fused2, cvb_bounds_grid, cvb_bounds_canny,
                                                  cvb_bounds_hough, cvb_bounds_ingest, cvb_bounds_overlay};
    unsigned long long total = 0;
    for (auto fn : tus) {
        unsigned long long n = 0; int line = 0;
        fn(&n, &line);
        if (n && first_line && !*first_line) *first_line = line;
        total += n;
    }
    return (long long)total;
#else
    (void)h;
    return -1;
#endif
}

int cvb_set_chunk_frames(cvb_handle *h, int frames)
{
    REQ_H(h);
    CVB_REQUIRE(frames >= 0, "chunk must be >= 1 frame, or 0 for the library's choice");
    h->chunk_frames = frames;
    return CVB_OK;
}

}  // extern "C"
