/*
 * cvb200.h -- C ABI of libcvb200.so: the B200-native (sm_100a) implementation
 * of the ChessVision per-frame hot path (hericmr/chessboard-vision).
 *
 * The reference has no FFI of its own: its "native" seam is the module-level
 * selector `from src.cython.<mod> import <Class>` (frame_enhancer.py:8-21,
 * 184-189; change_detector.py:7-19,203-208; setup.py:5-18).  The entry points
 * below are what a ctypes binding placed at that seam calls; each one cites
 * the reference method whose arithmetic it replaces.  INTEGRATION.md shows
 * the binding.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / numpy types.
 *  - every function returns 0 on success, a negative cvb_status otherwise;
 *    cvb_last_error() returns a thread-local message for the last failure.
 *  - images: uint8, interleaved HxWxC, rows contiguous, frames of a batch
 *    contiguous (frame stride = H*W*C).  `n` is the batch size.
 *  - *_dev functions take DEVICE pointers and only enqueue work on the
 *    handle's stream (call cvb_synchronize or order your own stream after
 *    it).  Functions without the suffix take HOST pointers, copy in/out and
 *    return after the results are in the host buffers.
 *  - there is no CPU fallback: without a CUDA device cvb_create fails.
 *  - a handle is single-threaded (one handle per thread / GPU).
 */
#ifndef CVB200_H
#define CVB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library itself is built -fvisibility=hidden */
#endif

#define CVB_VERSION 100

typedef struct cvb_handle cvb_handle;

typedef enum {
    CVB_OK = 0,
    CVB_ERR_INVALID = -1,     /* bad argument / unsupported shape            */
    CVB_ERR_CUDA = -2,        /* CUDA runtime error (message has the detail)  */
    CVB_ERR_NO_DEVICE = -3,   /* no usable sm_100 device                      */
    CVB_ERR_STATE = -4        /* per-stream state missing / shape mismatch    */
} cvb_status;

/* ---- lifecycle ------------------------------------------------------------ */
int  cvb_version(void);
int  cvb_device_count(void);
const char *cvb_last_error(void);
int  cvb_create(int device, cvb_handle **out);
void cvb_destroy(cvb_handle *h);
/* run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL
 * restores the handle's own stream */
int  cvb_set_stream(cvb_handle *h, void *cuda_stream);
int  cvb_synchronize(cvb_handle *h);
/* number of kernels this handle has launched since creation (bench.py's
 * gpu_launches claim is read from here) */
int64_t cvb_launch_count(cvb_handle *h);

/* per-kernel device timing: while enabled every launch is bracketed by CUDA
 * events on the handle's stream.  cvb_profile_read synchronises and returns,
 * per kernel name (32 bytes each, NUL padded), the summed milliseconds and
 * the launch count since cvb_profile_enable(h, 1).                          */
int cvb_profile_enable(cvb_handle *h, int on);
int cvb_profile_read(cvb_handle *h, char *names32, float *total_ms, int *counts, int max_entries, int *n_out);

/* ---- device / pinned memory helpers (so hosts need no torch) --------------- */
int cvb_malloc(cvb_handle *h, size_t bytes, void **dptr);
int cvb_free(cvb_handle *h, void *dptr);
int cvb_host_alloc(size_t bytes, void **hptr);            /* pinned */
int cvb_host_free(void *hptr);
int cvb_memcpy_h2d(cvb_handle *h, void *dst, const void *src, size_t bytes);  /* async on stream */
int cvb_memcpy_d2h(cvb_handle *h, void *dst, const void *src, size_t bytes);  /* async on stream */
int cvb_memset(cvb_handle *h, void *dst, int value, size_t bytes);
/* CUDA-event timers on the handle's stream */
int cvb_event_create(void **ev);
int cvb_event_destroy(void *ev);
int cvb_event_record(cvb_handle *h, void *ev);
int cvb_event_elapsed_ms(void *start, void *stop, float *ms);   /* syncs on stop */

/* ---- parameters -------------------------------------------------------------- */
/* S0 colour profile (frame_enhancer.py:56-99, the values of color_profile.json).  NumPy applies the python
 * scalars as f32, hence the float fields.  simd_block: pixels per vector block of OpenCV's HSV2BGR on the
 * reference host (32 for the AVX2 build of opencv-python 4.13): the vector body truncates, the row tail
 * rounds, and being bit-exact means reproducing that.                                                      */
typedef struct {
    double contrast, brightness;          /* cv2.convertScaleAbs(alpha, beta)    frame_enhancer.py:71 */
    float  hue_shift, sat_scale, val_scale;   /*                                  frame_enhancer.py:87-89 */
    int    radical_mode;                  /*                                      frame_enhancer.py:77-84 */
    float  target_hue, hue_window;
    int    simd_block;
} cvb_color_profile;
void cvb_color_profile_default(cvb_color_profile *p);

typedef struct {
    double clahe_clip_limit;   /* frame_enhancer.py:28  default 3.0            */
    int    tiles_x, tiles_y;   /* tileGridSize           default 8,8 (<=16)    */
    int    bilateral_d;        /* frame_enhancer.py:131  must be 9             */
    double sigma_color;        /*                         75                    */
    double sigma_space;        /*                         75                    */
    int    use_color_profile;  /* 0: step 0 of process_pipeline is the identity (no color_profile.json) */
    cvb_color_profile profile; /* frame_enhancer.py:167                          */
} cvb_enhance_params;
void cvb_enhance_params_default(cvb_enhance_params *p);

/* host-side table access (no GPU needed): the integer / f32 LUTs the kernels
 * use, for inspection and CPU-side tests.  Any pointer may be NULL. */
int cvb_get_tables(uint16_t *gamma256, uint16_t *cbrt2048, int32_t *lab2yf512,
                   uint8_t *invgamma4096, uint8_t *ltab2048);
int cvb_get_bilateral_tables(double sigma_color, double sigma_space,
                             float *color768, float *space81 /* [dy+4][dx+4], 0 outside r<=4 */);
int cvb_gaussian_kernel_q8(int ksize, int *q /* ksize entries */);

/* ---- frame_enhancer stages (stage-isolated entry points) -------------------- */
/* S0 apply_color_profile: convertScaleAbs -> BGR2HSV -> f32 hue/sat/val maths -> HSV2BGR
 *                                                frame_enhancer.py:56-99 */
int cvb_color_profile_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                          const cvb_color_profile *p, uint8_t *out);
/* S1 cv2.cvtColor(BGR2LAB)                       frame_enhancer.py:108 */
int cvb_bgr2lab_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, uint8_t *lab);
/* S3 cv2.cvtColor(LAB2BGR)                       frame_enhancer.py:120 */
int cvb_lab2bgr_dev(cvb_handle *h, const uint8_t *lab, int n, int H, int W, uint8_t *bgr);
/* S2 clahe.apply on one u8 plane                 frame_enhancer.py:36,114
 * hist_out: n*tiles*256 int32 or NULL; lut_out: n*tiles*256 u8 or NULL      */
int cvb_clahe_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W,
                  double clip_limit, int tiles_x, int tiles_y,
                  uint8_t *out, int32_t *hist_out, uint8_t *lut_out);
/* correct_lighting = S1+S2+S3 fused               frame_enhancer.py:101-120 */
int cvb_correct_lighting_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                             double clip_limit, int tiles_x, int tiles_y,
                             uint8_t *out, int32_t *hist_out, uint8_t *lut_out);
/* S4 cv2.bilateralFilter(d=9, sigma, sigma)       frame_enhancer.py:122-131 */
int cvb_bilateral_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                      int d, double sigma_color, double sigma_space, uint8_t *out);
/* S5 cv2.filter2D 3x3 sharpen, 3 channels         frame_enhancer.py:133-138 */
int cvb_sharpen_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, uint8_t *out);
/* S6 cv2.normalize(NORM_MINMAX, 0, 255); C = channels (any)
 * minmax_out: n*2 int32 DEVICE or NULL            frame_enhancer.py:140-146 */
int cvb_normalize_dev(cvb_handle *h, const uint8_t *src, int n, int H, int W, int C,
                      uint8_t *out, int32_t *minmax_out);
/* S7 cv2.cvtColor(BGR2GRAY)                       frame_enhancer.py:154 */
int cvb_gray_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, uint8_t *gray);
/* S8 cv2.GaussianBlur((k,k),0) on a u8 plane, k odd <= 31
 *                                                frame_enhancer.py:156 */
int cvb_gaussian_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W, int ksize, uint8_t *out);
/* prepare_analysis = S7+S8(5)+S9                  frame_enhancer.py:148-159
 * blurred (n*H*W) / otsu_t (n int32, DEVICE) / hist (n*256 int32) may be NULL */
int cvb_prepare_analysis_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                             uint8_t *gray, uint8_t *binary, uint8_t *blurred,
                             int32_t *otsu_t, int32_t *hist);
/* process_pipeline (colour profile off) = S1..S6    frame_enhancer.py:161-181 */
int cvb_process_pipeline_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                             const cvb_enhance_params *p, uint8_t *enhanced);
/* process_pipeline + prepare_analysis in four passes;
 * gray / binary / otsu_t (DEVICE int32[n]) may be NULL to skip the tail     */
int cvb_enhance_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                    const cvb_enhance_params *p,
                    uint8_t *enhanced, uint8_t *gray, uint8_t *binary, int32_t *otsu_t);
/* host-buffer variant: H2D, cvb_enhance_dev, D2H, synchronise.  otsu_t: HOST
 * int32[n].  Any output may be NULL.                                         */
int cvb_enhance(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                const cvb_enhance_params *p,
                uint8_t *enhanced, uint8_t *gray, uint8_t *binary, int32_t *otsu_t);

/* ---- board_detection.warp_image ---------------------------------------------- */
/* cv2.getPerspectiveTransform (host, f64 LU)       board_detection.py:65-68 */
int cvb_get_perspective_transform(const float *src_xy4, const float *dst_xy4, double *M9);
/* cv2.warpPerspective(INTER_LINEAR, BORDER_CONSTANT 0); M9: HOST, n_mats * 9
 * forward matrices (n_mats == 1: shared by the batch, else == n)
 *                                                board_detection.py:69 */
int cvb_warp_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                 const double *M9, int n_mats, int out_h, int out_w, uint8_t *warped);

/* the same followed by cv2.rotate(warped, cv2.ROTATE_180) (game_session.py:103-104,
 * 125-126: orientation_flipped), folded into the warp's stores              */
int cvb_warp_rot180_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                        const double *M9, int n_mats, int out_h, int out_w, uint8_t *warped);
/* cv2.rotate(img, code): 0 ROTATE_90_CLOCKWISE, 1 ROTATE_180,
 * 2 ROTATE_90_COUNTERCLOCKWISE; n images of H x W x C (C = 1 or 3); the 90-degree
 * codes write n images of W x H x C   (ui_renderer.py:148-149, game_session.py:126) */
int cvb_rotate_dev(cvb_handle *h, const uint8_t *src, int n, int H, int W, int C,
                   int rotate_code, uint8_t *dst);

/* ---- board_detection.find_chessboard_corners, image part (calibration time) ---- */
/* cv2.GaussianBlur(plane, (k, k), sigma) for u8; sigma <= 0: OpenCV's default for k   board_detection.py:10 */
int cvb_gaussian_sigma_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W,
                           int ksize, double sigma, uint8_t *out);
/* cv2.dilate(plane, np.ones((kh, kw), np.uint8), iterations=iterations)               board_detection.py:13-14 */
int cvb_dilate_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W,
                   int kw, int kh, int iterations, uint8_t *out);
/* gray -> GaussianBlur(7x7, 1) -> Canny(30, 100) -> dilate(5x5, 3): the mask cv2.findContours
 * is run on (host, control plane)                                                    board_detection.py:9-14 */
int cvb_contour_mask_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W, uint8_t *mask);

/* ---- SmartGridExtractor.refine_grid building blocks (calibration time) -------- */
/* cv2.Canny(gray, low, high) (aperture 3, L1 gradient)   grid_extractor.py:74 */
int cvb_canny_dev(cvb_handle *h, const uint8_t *gray, int n, int H, int W,
                  double low_thresh, double high_thresh, uint8_t *edges);
/* np.sum(plane, axis=1) / axis=0 as uint32 (DEVICE, n*H and n*W)  grid_extractor.py:78-80 */
int cvb_projections_dev(cvb_handle *h, const uint8_t *plane, int n, int H, int W,
                        uint32_t *row_sums, uint32_t *col_sums);

/* ---- per-square statistics ------------------------------------------------------ */
/* A "board" is an image (warped board or an atlas of packed squares) of
 * BH x BW x C (C = 1 or 3); squares are rectangles inside it
 * (grid_extractor.py:33-56,140-161).                                          */
typedef struct { int32_t x, y, w, h; } cvb_rect;

/* per (frame, square) result; layout is part of the ABI (128 bytes)          */
typedef struct {
    int32_t  n;                 /* pixels in the square                         */
    int32_t  has_ref;           /* 1 when a PD reference existed                */
    uint32_t sum;               /* sum g        (g = gray+blur5 square)         */
    uint32_t sad;               /* sum |g-ref|  piece_detector.py:87-93         */
    uint64_t sumsq;             /* sum g^2      piece_detector.py:305           */
    uint32_t center_sum, center_cnt;      /* piece_detector.py:184-201          */
    uint32_t border_sum, border_cnt;
    uint32_t ring_sum[4], ring_cnt[4];    /* piece_detector.py:148-166          */
    int32_t  cd_changed;        /* #(z > z_threshold) change_detector.py:129-131 */
    float    cd_zmax;           /* max z                change_detector.py:159  */
    int32_t  cd_valid;          /* 1 when CD state existed and square selected  */
    int32_t  reserved[11];
} cvb_square_stats;

/* per-stream state (ChangeDetector means/variances, PieceDetector reference):
 * planes of BH x BW living on the device, one set per stream slot            */
typedef struct cvb_state cvb_state;
int  cvb_state_create(cvb_handle *h, int n_streams, int BH, int BW, cvb_state **out);
void cvb_state_destroy(cvb_state *s);

enum {
    CVB_SQ_PD_STATS     = 1,    /* moments, centre/border, rings, SAD vs ref    */
    CVB_SQ_PD_SET_REF   = 2,    /* ref <- g for selected squares (after stats)  */
    CVB_SQ_CD_CALIBRATE = 4,    /* mean <- g, var <- initial_variance           */
    CVB_SQ_CD_DETECT    = 8,    /* z-score count / max                          */
    CVB_SQ_CD_UPDATE    = 16    /* EMA update (after detect)                    */
};
typedef struct {
    int   ops;                  /* bitmask above                                */
    int   pd_blur;              /* PieceDetector blur, 5   piece_detector.py:133 */
    int   cd_blur;              /* ChangeDetector blur_kernel|1  change_detector.py:55 */
    float z_threshold;          /* 2.5                                          */
    float alpha;                /* f32(alpha)                                   */
    float one_minus_alpha;      /* f32(1-alpha)                                 */
    float initial_variance;     /* 100                                          */
    float min_variance;         /* 10   change_detector.py:89                   */
} cvb_square_params;
void cvb_square_params_default(cvb_square_params *p);

/* boards: n x BH x BW x C (DEVICE).  rects: HOST, n_sq entries (shared by the
 * batch).  select: HOST n_sq bytes or NULL (all) -- squares with select==0 are
 * skipped for the state-changing / CD ops (focus_squares).  Frame i uses
 * state slot stream0+i.  stats: DEVICE n*n_sq cvb_square_stats or NULL.      */
int cvb_squares_dev(cvb_handle *h, const uint8_t *boards, int n, int BH, int BW, int C,
                    const cvb_rect *rects, int n_sq, const uint8_t *select,
                    cvb_state *state, int stream0, const cvb_square_params *p,
                    cvb_square_stats *stats);
/* state <-> host; plane: 0 pd_ref(u8) 1 cd_mean(f32) 2 cd_var(f32)
 * 3 pd_cur(u8: last gray+blur of pd_blur) 4 flags(u8 per pixel: bit0 has_ref,
 * bit1 has_cd).  Buffers are BH*BW elements.                                 */
int cvb_state_get(cvb_handle *h, cvb_state *s, int stream, int plane, void *host_out);
int cvb_state_set(cvb_handle *h, cvb_state *s, int stream, int plane, const void *host_in);
/* forget PD references / CD model of one stream slot (-1: all)               */
int cvb_state_reset(cvb_handle *h, cvb_state *s, int stream);

/* ---- PieceDetector._detect_circle_unified: cv2.HoughCircles per square ------------- */
/* cv2.HoughCircles(gray, HOUGH_GRADIENT, dp, minDist, param1, param2, minRadius,
 * maxRadius) on every selected square of n gray planes  piece_detector.py:210-270.
 * One CTA per square, the whole transform in shared memory for squares up to
 * CVB_HOUGH_MAX_DIM pixels a side; larger squares, up to CVB_HOUGH_MAX_DIM_GLOBAL (and
 * at most 253 accumulator cells a side, i.e. dp >= 1.01 at 254 pixels), run the same kernel
 * on a global-memory workspace.  Bit-identical to OpenCV 4.13 (circle list, order, f32
 * values).                                                                       */
#define CVB_HOUGH_MAX_CIRCLES 16   /* circles stored per square (minDist = side/3 allows <= 16) */
#define CVB_HOUGH_MAX_DIM     128
#define CVB_HOUGH_MAX_DIM_GLOBAL 254   /* (side + 2)^2 padded pixels are indexed with 16 bits */
enum { CVB_HOUGH_OK = 0, CVB_HOUGH_SKIPPED = 1 };
typedef struct {
    float  dp;                  /* 1.2 (values < 1 are raised to 1, as OpenCV does)   piece_detector.py:235 */
    float  min_dist;            /* used when min_dist_div <= 0                            */
    double param1;              /* Canny high threshold, 100   piece_detector.py:230      */
    double param2;              /* accumulator / support threshold, 25   :229             */
    double min_radius_ratio;    /* minRadius = (int)(min(h,w) * ratio), 0.20   :225; < 0: use min_radius */
    double max_radius_ratio;    /* maxRadius = (int)(min(h,w) * ratio), 0.55   :226; < 0: use max_radius */
    int    min_radius, max_radius;
    int    min_dist_div;        /* 3: minDist = min(h,w) / 3   :236                       */
    int    reserved;
} cvb_hough_params;
void cvb_hough_params_default(cvb_hough_params *p);
/* what one square's call resolves to (OpenCV's argument handling included) */
typedef struct {
    int32_t x, y, w, h;
    int32_t min_radius, max_radius;
    int32_t acc_rows, acc_cols; /* ceil(h / dp), ceil(w / dp)                             */
    int32_t n_bins;             /* radius histogram bins, 10 per dp                       */
    float   min_dist;
} cvb_hough_square;
typedef struct {
    int32_t count;              /* circles after the minDist filter (may exceed the 16 stored) */
    int32_t n_edges;            /* Canny edge pixels that voted                           */
    int32_t n_centers;          /* accumulator maxima examined                            */
    int32_t status;             /* CVB_HOUGH_OK / CVB_HOUGH_SKIPPED                       */
    float   xyr[CVB_HOUGH_MAX_CIRCLES][3];   /* (x, y, radius) in OpenCV's order          */
    int32_t support[CVB_HOUGH_MAX_CIRCLES];  /* edge pixels in the winning radius window  */
} cvb_hough_result;
/* host only: rects -> per-square geometry */
int cvb_hough_geometry(const cvb_rect *rects, int n_sq, const cvb_hough_params *p, cvb_hough_square *out);
/* planes: DEVICE n x PH x PW u8 (gray + blur, e.g. state plane 3); rects HOST
 * (shared by the batch); select: HOST n*n_sq bytes or NULL (all);
 * results: DEVICE n*n_sq.                                                        */
int cvb_hough_dev(cvb_handle *h, const uint8_t *planes, int n, int PH, int PW,
                  const cvb_rect *rects, int n_sq, const uint8_t *select,
                  const cvb_hough_params *p, cvb_hough_result *results);
/* the same on the pd_cur planes (last gray + blur of every square) of the state
 * slots stream0 .. stream0+n-1; results: HOST n*n_sq; synchronises.              */
int cvb_hough_state(cvb_handle *h, cvb_state *state, int stream0, int n,
                    const cvb_rect *rects, int n_sq, const uint8_t *select,
                    const cvb_hough_params *p, cvb_hough_result *results_host);

/* ---- the whole hot path for a batch ------------------------------------------------ */
typedef struct {
    cvb_enhance_params enhance;
    cvb_square_params  squares;
    int warp_enhanced;          /* 1: warp the enhanced frame, 0: the raw frame  */
    int board_size;             /* 620 = min(1280,720)-100  board_detection.py:65 */
    int rotate_180;             /* 1: cv2.rotate(warped, ROTATE_180) before the split, as
                                 * game_session.py:125-126 does when orientation_flipped  */
    int reserved;
} cvb_pipeline_params;
void cvb_pipeline_params_default(cvb_pipeline_params *p);

/* enhance -> prepare_analysis -> warp -> 64 squares -> PD/CD statistics.
 * All pointers DEVICE except M9 / rects / select (HOST).  enhanced, gray,
 * binary, warped may be NULL: the library then uses its own workspace.       */
int cvb_pipeline_dev(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                     const cvb_pipeline_params *p,
                     const double *M9, int n_mats,
                     const cvb_rect *rects, int n_sq, const uint8_t *select,
                     cvb_state *state, int stream0,
                     uint8_t *enhanced, uint8_t *gray, uint8_t *binary, int32_t *otsu_t,
                     uint8_t *warped, cvb_square_stats *stats);
/* host-buffer variant (the call the Python shim times as "e2e"): bgr HOST
 * (pinned or pageable) in; stats / otsu_t HOST out.  The batch is processed
 * in chunks of cvb_set_chunk_frames() frames (default 0 = the library's choice:
 * 8 for BGR input, wave-sized chunks of 12..32 for YUV input): the host->device
 * copy of chunk k+1 overlaps the kernels of chunk k when bgr is page-locked. */
int cvb_set_chunk_frames(cvb_handle *h, int frames);
int cvb_pipeline(cvb_handle *h, const uint8_t *bgr, int n, int H, int W,
                 const cvb_pipeline_params *p,
                 const double *M9, int n_mats,
                 const cvb_rect *rects, int n_sq, const uint8_t *select,
                 cvb_state *state, int stream0,
                 int32_t *otsu_t, cvb_square_stats *stats);

/* Debug build only (`make -C chessboard_vision_b200/csrc debug` -> libcvb200_dbg.so): number of shared-memory / global
 * index checks that failed since the library was loaded and the source line of the first one; -1 in the release
 * library, which carries no checks.  Stands in for compute-sanitizer, which is closed on the target pool; the
 * reference has no counterpart (its arrays are bounds-checked by NumPy, e.g. grid_extractor.py:46). */
long long cvb_debug_bounds_violations(cvb_handle *h, int *first_line);

/* ---- board overlay (second half) ---------------------------------- */
/* GameSession._draw_interface (game_session.py:293-388) draws the grid, the tints and highlights
 * and the piece letters on the warped board with cv2.line / rectangle / circle / putText and
 * `overlay = vis.copy(); <shape>; cv2.addWeighted(overlay, a, vis, b, 0, vis)`.  The same
 * drawing as a display list applied in order on a DEVICE image, pixel-exactly:
 *   RECT    inclusive corners (x0,y0)-(x1,y1), clipped: cv2.rectangle(.., -1), axis-aligned
 *           cv2.line of thickness 1, `overlay[:] = colour`       game_session.py:301-312,326-336,346
 *   CIRCLE  centre (x0,y0), radius x1: cv2.circle(.., -1)        game_session.py:356
 *   STAMP   a w x h (x1, y1) 1-bit mask with its top-left at (x0,y0): what cv2.putText sets
 *           (rasterised once per string by the caller)            game_session.py:316,375-378,382,385
 * alpha == 1 && beta == 0 stores `color`; otherwise dst = saturate(rint(fma(color, alpha,
 * dst * beta))), OpenCV's addWeighted for 8-bit images.  Ops that share a non-zero `group` are
 * shapes drawn on ONE overlay copy (game_session.py:322-338): a pixel is blended once per group;
 * they must be consecutive and share color / alpha / beta. */
enum { CVB_OV_RECT = 0, CVB_OV_CIRCLE = 1, CVB_OV_STAMP = 2 };
typedef struct {
    int32_t kind;
    int32_t x0, y0, x1, y1;
    uint8_t color[4];          /* B, G, R, 0 */
    float alpha, beta;
    int32_t group;             /* 0: on its own */
    uint32_t aux_ofs;          /* STAMP: byte offset of its mask in `masks` (rows of (w + 7) / 8 bytes, bit x & 7 of
                                  byte x >> 3); set by the library for CIRCLE */
} cvb_overlay_op;
/* bgr: DEVICE, n images of H x W x 3, drawn in place (every image gets the same list);
 * ops / masks: HOST */
int cvb_overlay_dev(cvb_handle *h, uint8_t *bgr, int n, int H, int W, const cvb_overlay_op *ops, int n_ops,
                    const uint8_t *masks, size_t mask_bytes);

/* ---- camera ingest ---------------------------------------------- */
/* play_lichess.py:16-18,45 / game_session.py:99,113 receive BGR frames from
 * cv2.VideoCapture.read(), i.e. after OpenCV has converted the camera's native YUV on the
 * host.  These entry points take the native frame instead and convert on the device,
 * bit-exactly as cv2.cvtColor(COLOR_YUV2BGR_YUY2 / COLOR_YUV2BGR_NV12) does:
 *   CVB_FMT_BGR   (H, W, 3)        3   bytes / pixel  (no conversion)
 *   CVB_FMT_YUY2  (H, W, 2)        2   bytes / pixel  Y0 U Y1 V, W even
 *   CVB_FMT_NV12  (H * 3 / 2, W)   1.5 bytes / pixel  Y plane, then interleaved U V rows; W, H even */
#define CVB_FMT_BGR  0
#define CVB_FMT_YUY2 1
#define CVB_FMT_NV12 2
size_t cvb_frame_bytes(int format, int H, int W);
/* src, bgr: DEVICE.  n frames of cvb_frame_bytes() each -> n BGR frames. */
int cvb_cvt_to_bgr_dev(cvb_handle *h, const uint8_t *src, int format, int n, int H, int W, uint8_t *bgr);
/* cvb_pipeline on frames in their native format (HOST, pinned or pageable): only
 * cvb_frame_bytes() per frame cross PCIe; cvb_pipeline(...) == cvb_pipeline_fmt(..., CVB_FMT_BGR, ...). */
int cvb_pipeline_fmt(cvb_handle *h, const uint8_t *frames, int format, int n, int H, int W,
                     const cvb_pipeline_params *p,
                     const double *M9, int n_mats,
                     const cvb_rect *rects, int n_sq, const uint8_t *select,
                     cvb_state *state, int stream0,
                     int32_t *otsu_t, cvb_square_stats *stats);
/* The same call split in two for a caller that receives batches continuously (a capture loop:
 * game_session.py:113 / play_lichess.py:45 per frame): cvb_pipeline_submit enqueues the copies
 * in, the kernels and the copies out, writes a ticket (> 0) and returns; cvb_pipeline_wait(h,
 * ticket) returns when that submission is complete (ticket 0: everything submitted so far).
 * Submitting batch k+1 before waiting for batch k lets its first host->device copy run beside
 * the kernels of batch k; batches execute in submission order, so per-stream state evolves as
 * with cvb_pipeline_fmt.  `frames`, `otsu_t` and `stats` must stay valid (and `frames`
 * unmodified) until the wait; use a different result buffer per batch in flight.  Pinned host
 * memory (cvb_host_alloc) is needed for the overlap. */
int cvb_pipeline_submit(cvb_handle *h, const uint8_t *frames, int format, int n, int H, int W,
                        const cvb_pipeline_params *p,
                        const double *M9, int n_mats,
                        const cvb_rect *rects, int n_sq, const uint8_t *select,
                        cvb_state *state, int stream0,
                        int32_t *otsu_t, cvb_square_stats *stats, uint64_t *ticket);
int cvb_pipeline_wait(cvb_handle *h, uint64_t ticket);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* CVB200_H */
