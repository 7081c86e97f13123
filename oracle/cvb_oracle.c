/*
 * cvb_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic that the ChessVision hot path
 * (hericmr/chessboard-vision) executes through OpenCV 4.13.0 / NumPy 2.3.5.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file's shared object.  The shipped
 * CUDA path never links or calls it.
 *
 * Parity pinning: every function below is checked (tests/test_oracle_*.py)
 * against golden vectors in tests/golden/ that were produced by importing the
 * UNMODIFIED reference modules from /root/reference (tools/make_golden.py),
 * and -- when cv2 is importable -- against cv2 live.  The reference's own
 * tests hold one behavioural vector for this path
 * (test_change_detector_regression.py:31-54); it is replayed too.
 *
 * The arithmetic itself lives in un-vendored third-party code
 * (opencv-python 4.13.0.92, numpy 2.3.5; requirements.txt:1-2), so each
 * function cites (a) the reference call site it stands in for and (b) the
 * upstream OpenCV source file whose published algorithm it restates.
 *
 * Conventions: u8 images are HxWxC interleaved, C-contiguous unless a stride
 * is given.  rint() is round-half-to-even (default FE mode).  This file must
 * be compiled with -ffp-contract=off so that no a*b+c is fused silently.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

typedef uint8_t u8;

#define ORC_API __attribute__((visibility("default")))

static inline int reflect101(int p, int n)
{
    /* cv::borderInterpolate(BORDER_REFLECT_101) */
    if (n == 1) return 0;
    while (p < 0 || p >= n) {
        if (p < 0) p = -p;
        else p = 2 * n - 2 - p;
    }
    return p;
}
static inline int sat_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* ------------------------------------------------------------------------ */
/* Tables                                                                   */
/* ------------------------------------------------------------------------ */
static uint16_t g_gamma[256];      /* sRGBGammaTab_b   (color_lab.cpp)      */
static uint16_t g_cbrt[3072];      /* LabCbrtTab_b                          */
static int32_t  g_lab2yf[512];     /* LabToYF_b  (y, ify) pairs             */
static u8       g_invgamma[4096];  /* sRGBInvGammaTab_b                     */
static int g_tables_ready = 0;

static double srgb_gamma(double x)
{
    return x <= 0.04045 ? x * (1.0 / 12.92) : pow((x + 0.055) * (1.0 / 1.055), 2.4);
}
static double srgb_inv_gamma(double x)
{
    return x <= 0.0031308 ? x * 12.92 : 1.055 * pow(x, 1.0 / 2.4) - 0.055;
}

ORC_API void orc_tables_init(void)
{
    if (g_tables_ready) return;
    for (int i = 0; i < 256; i++)
        g_gamma[i] = (uint16_t)rint(2040.0 * srgb_gamma(i / 255.0));
    for (int i = 0; i < 3072; i++) {
        double x = i / 2040.0;
        double f = x < 216.0 / 24389.0 ? x * (841.0 / 108.0) + 16.0 / 116.0 : cbrt(x);
        g_cbrt[i] = (uint16_t)rint(32768.0 * f);
    }
    /* OpenCV builds this table with its softfloat cbrt; two entries land on
     * the other side of a rounding boundary (probes 8-10). */
    g_cbrt[49] -= 1;
    g_cbrt[628] += 1;
    const double BASE = 16384.0;
    for (int L = 0; L < 256; L++) {
        double y, ify;
        if (L <= 20) {
            y = rint(L * BASE * 180.0 / (17.0 * 29.0 * 29.0 * 29.0));
            ify = rint(BASE * (16.0 / 116.0 + L * 5.0 / 1479.0));
        } else {
            double fy = L * 100.0 * BASE / (255.0 * 116.0) + 16.0 * BASE / 116.0;
            ify = rint(fy);
            y = rint(fy * fy * fy / (BASE * BASE));
        }
        g_lab2yf[2 * L] = (int32_t)y;
        g_lab2yf[2 * L + 1] = (int32_t)ify;
    }
    for (int i = 0; i < 4096; i++)
        g_invgamma[i] = (u8)sat_u8((int)rint(255.0 * srgb_inv_gamma(i / 4096.0)));
    g_tables_ready = 1;
}

/* table export so tests can compare the product's tables with these */
ORC_API void orc_get_tables(uint16_t *gamma256, uint16_t *cbrt3072, int32_t *lab2yf512, u8 *invgamma4096)
{
    orc_tables_init();
    if (gamma256) memcpy(gamma256, g_gamma, sizeof g_gamma);
    if (cbrt3072) memcpy(cbrt3072, g_cbrt, sizeof g_cbrt);
    if (lab2yf512) memcpy(lab2yf512, g_lab2yf, sizeof g_lab2yf);
    if (invgamma4096) memcpy(invgamma4096, g_invgamma, sizeof g_invgamma);
}

/* ------------------------------------------------------------------------ */
/* S1  BGR -> LAB (u8)        frame_enhancer.py:108                          */
/* OpenCV: imgproc/src/color_lab.cpp  RGB2Lab_b::operator()                 */
/* ------------------------------------------------------------------------ */
static inline void bgr2lab_px(const u8 *p, u8 *o)
{
    int B = g_gamma[p[0]], G = g_gamma[p[1]], R = g_gamma[p[2]];
    int fX = g_cbrt[(R * 1777 + G * 1541 + B * 778 + 2048) >> 12];
    int fY = g_cbrt[(R * 871 + G * 2929 + B * 296 + 2048) >> 12];
    int fZ = g_cbrt[(R * 73 + G * 448 + B * 3575 + 2048) >> 12];
    int L = (296 * fY - 1336934 + 16384) >> 15;
    int a = (500 * (fX - fY) + 128 * 32768 + 16384) >> 15;
    int b = (200 * (fY - fZ) + 128 * 32768 + 16384) >> 15;
    o[0] = (u8)sat_u8(L); o[1] = (u8)sat_u8(a); o[2] = (u8)sat_u8(b);
}
ORC_API void orc_bgr2lab(const u8 *bgr, long n_px, u8 *lab)
{
    orc_tables_init();
    for (long i = 0; i < n_px; i++) bgr2lab_px(bgr + 3 * i, lab + 3 * i);
}

/* ------------------------------------------------------------------------ */
/* S3  LAB -> BGR (u8)        frame_enhancer.py:120                          */
/* OpenCV: color_lab.cpp  Lab2RGBinteger::process (8-bit path)              */
/* ------------------------------------------------------------------------ */
static inline int ab_to_xz(int t)
{
    /* abToXZ_b[t - minABvalue]; C integer division truncates toward zero */
    if (t <= 3390) return t * 108 / 841 - 290;
    return t * t / 16384 * t / 16384;
}
static inline void lab2bgr_px(const u8 *p, u8 *o)
{
    int L = p[0], a = p[1], b = p[2];
    int y = g_lab2yf[2 * L], ify = g_lab2yf[2 * L + 1];
    int adiv = ((5 * a * 53687 + 128) >> 13) - 4194;
    int bdiv = ((b * 41943 + 16) >> 9) - 10485 + 1;
    int x = ab_to_xz(ify + adiv);
    int z = ab_to_xz(ify - bdiv);
    int ro = (12615 * x - 6296 * y - 2223 * z + 8192) >> 14;
    int go = (-3773 * x + 7684 * y + 185 * z + 8192) >> 14;
    int bo = (217 * x - 836 * y + 4715 * z + 8192) >> 14;
    ro = ro < 0 ? 0 : (ro > 4095 ? 4095 : ro);
    go = go < 0 ? 0 : (go > 4095 ? 4095 : go);
    bo = bo < 0 ? 0 : (bo > 4095 ? 4095 : bo);
    o[0] = g_invgamma[bo]; o[1] = g_invgamma[go]; o[2] = g_invgamma[ro];
}
ORC_API void orc_lab2bgr(const u8 *lab, long n_px, u8 *bgr)
{
    orc_tables_init();
    for (long i = 0; i < n_px; i++) lab2bgr_px(lab + 3 * i, bgr + 3 * i);
}

/* ------------------------------------------------------------------------ */
/* S2  CLAHE on a u8 plane    frame_enhancer.py:36,114                       */
/* OpenCV: imgproc/src/clahe.cpp  CLAHE_CalcLut_Body / Interpolation_Body   */
/* ------------------------------------------------------------------------ */
ORC_API void orc_clahe_geometry(int H, int W, int tiles_x, int tiles_y,
                                int *tile_w, int *tile_h, int *ext_w, int *ext_h)
{
    int ew = W, eh = H;
    if (W % tiles_x != 0 || H % tiles_y != 0) {
        /* copyMakeBorder(src, 0, ty - H%ty, 0, tx - W%tx, REFLECT_101): note
         * OpenCV pads BOTH axes by (t - dim%t) even when one divides evenly */
        ew = W + (tiles_x - W % tiles_x);
        eh = H + (tiles_y - H % tiles_y);
    }
    *ext_w = ew; *ext_h = eh; *tile_w = ew / tiles_x; *tile_h = eh / tiles_y;
}

ORC_API void orc_clahe_hist(const u8 *src, int H, int W, int tiles_x, int tiles_y, int32_t *hist)
{
    int tw, th, ew, eh;
    orc_clahe_geometry(H, W, tiles_x, tiles_y, &tw, &th, &ew, &eh);
    memset(hist, 0, sizeof(int32_t) * 256 * tiles_x * tiles_y);
    for (int y = 0; y < eh; y++) {
        int sy = reflect101(y, H), ty = y / th;
        for (int x = 0; x < ew; x++) {
            int sx = reflect101(x, W), tx = x / tw;
            hist[(ty * tiles_x + tx) * 256 + src[(long)sy * W + sx]]++;
        }
    }
}

ORC_API void orc_clahe_lut(const int32_t *hist, int n_tiles, int tile_area, double clip_limit, u8 *lut)
{
    int clip = 0;
    if (clip_limit > 0.0) {
        clip = (int)(clip_limit * tile_area / 256);
        if (clip < 1) clip = 1;
    }
    float lut_scale = (float)255 / (float)tile_area;
    for (int t = 0; t < n_tiles; t++) {
        int h[256];
        memcpy(h, hist + t * 256, sizeof h);
        if (clip > 0) {
            int clipped = 0;
            for (int i = 0; i < 256; i++)
                if (h[i] > clip) { clipped += h[i] - clip; h[i] = clip; }
            int batch = clipped / 256, resid = clipped - batch * 256;
            for (int i = 0; i < 256; i++) h[i] += batch;
            if (resid != 0) {
                int step = 256 / resid; if (step < 1) step = 1;
                for (int i = 0; i < 256 && resid > 0; i += step, resid--) h[i]++;
            }
        }
        int sum = 0;
        for (int i = 0; i < 256; i++) {
            sum += h[i];
            float v = (float)sum * lut_scale;
            lut[t * 256 + i] = (u8)sat_u8((int)rintf(v));
        }
    }
}

ORC_API void orc_clahe_apply(const u8 *src, int H, int W, int tiles_x, int tiles_y,
                             int tile_w, int tile_h, const u8 *lut, u8 *dst)
{
    float inv_tw = 1.0f / (float)tile_w, inv_th = 1.0f / (float)tile_h;
    for (int y = 0; y < H; y++) {
        float tyf = (float)y * inv_th - 0.5f;
        int ty1 = (int)floorf(tyf), ty2 = ty1 + 1;
        float ya = tyf - (float)ty1, ya1 = 1.0f - ya;
        if (ty1 < 0) ty1 = 0;
        if (ty2 > tiles_y - 1) ty2 = tiles_y - 1;
        for (int x = 0; x < W; x++) {
            float txf = (float)x * inv_tw - 0.5f;
            int tx1 = (int)floorf(txf), tx2 = tx1 + 1;
            float xa = txf - (float)tx1, xa1 = 1.0f - xa;
            if (tx1 < 0) tx1 = 0;
            if (tx2 > tiles_x - 1) tx2 = tiles_x - 1;
            int v = src[(long)y * W + x];
            float l11 = lut[(ty1 * tiles_x + tx1) * 256 + v], l12 = lut[(ty1 * tiles_x + tx2) * 256 + v];
            float l21 = lut[(ty2 * tiles_x + tx1) * 256 + v], l22 = lut[(ty2 * tiles_x + tx2) * 256 + v];
            float r0 = l11 * xa1; float r1 = l12 * xa; float top = r0 + r1; top = top * ya1;
            float r2 = l21 * xa1; float r3 = l22 * xa; float bot = r2 + r3; bot = bot * ya;
            float res = top + bot;
            dst[(long)y * W + x] = (u8)sat_u8((int)rintf(res));
        }
    }
}

/* full CLAHE; hist_out (tiles*256 int32) and lut_out (tiles*256 u8) optional */
ORC_API void orc_clahe(const u8 *src, int H, int W, double clip_limit, int tiles_x, int tiles_y,
                       u8 *dst, int32_t *hist_out, u8 *lut_out)
{
    int tw, th, ew, eh, nt = tiles_x * tiles_y;
    orc_clahe_geometry(H, W, tiles_x, tiles_y, &tw, &th, &ew, &eh);
    int32_t *hist = hist_out ? hist_out : (int32_t *)malloc(sizeof(int32_t) * 256 * nt);
    u8 *lut = lut_out ? lut_out : (u8 *)malloc(256 * nt);
    orc_clahe_hist(src, H, W, tiles_x, tiles_y, hist);
    orc_clahe_lut(hist, nt, tw * th, clip_limit, lut);
    orc_clahe_apply(src, H, W, tiles_x, tiles_y, tw, th, lut, dst);
    if (!hist_out) free(hist);
    if (!lut_out) free(lut);
}

/* correct_lighting = S1 -> S2(L) -> S3   (frame_enhancer.py:101-120) */
ORC_API void orc_correct_lighting(const u8 *bgr, int H, int W, double clip_limit, int tiles_x, int tiles_y, u8 *out)
{
    long n = (long)H * W;
    u8 *lab = (u8 *)malloc(3 * n), *l = (u8 *)malloc(n), *l2 = (u8 *)malloc(n);
    orc_bgr2lab(bgr, n, lab);
    for (long i = 0; i < n; i++) l[i] = lab[3 * i];
    orc_clahe(l, H, W, clip_limit, tiles_x, tiles_y, l2, NULL, NULL);
    for (long i = 0; i < n; i++) lab[3 * i] = l2[i];
    orc_lab2bgr(lab, n, out);
    free(lab); free(l); free(l2);
}

/* ------------------------------------------------------------------------ */
/* S4  bilateral filter 8UC3  frame_enhancer.py:131                          */
/* OpenCV: imgproc/src/bilateral_filter.dispatch.cpp (tables, radius,       */
/*         circular support, REFLECT_101) + bilateral_filter.simd.hpp       */
/*         (per-tap accumulation order, 1/wsum finalisation)                */
/* use_fma: 1 = wsum/sums accumulate with fmaf (OpenCV v_muladd on FMA3     */
/*          hosts), 0 = separate multiply and add.                          */
/* ------------------------------------------------------------------------ */
ORC_API int orc_bilateral_tables(int d, double sigma_color, double sigma_space,
                                 float *color768, float *space, int *dy, int *dx)
{
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
    int radius = d <= 0 ? (int)rint(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    for (int i = 0; i < 768; i++) color768[i] = (float)exp((double)i * i * gc);
    int maxk = 0;
    for (int i = -radius; i <= radius; i++)
        for (int j = -radius; j <= radius; j++) {
            double r = sqrt((double)i * i + (double)j * j);
            if (r > radius) continue;
            space[maxk] = (float)exp(r * r * gs);
            dy[maxk] = i; dx[maxk] = j; maxk++;
        }
    return maxk;
}

ORC_API void orc_bilateral(const u8 *src, int H, int W, int d, double sigma_color, double sigma_space,
                           int use_fma, u8 *dst)
{
    float color[768], space[512];
    int dy[512], dx[512];
    int maxk = orc_bilateral_tables(d, sigma_color, sigma_space, color, space, dy, dx);
    int radius = d <= 0 ? (int)rint(sigma_space * 1.5) : d / 2;
    if (radius < 1) radius = 1;
    int PW = W + 2 * radius, PH = H + 2 * radius;
    u8 *pad = (u8 *)malloc((long)PW * PH * 3);
    for (int y = 0; y < PH; y++) {
        int sy = reflect101(y - radius, H);
        for (int x = 0; x < PW; x++) {
            int sx = reflect101(x - radius, W);
            memcpy(pad + ((long)y * PW + x) * 3, src + ((long)sy * W + sx) * 3, 3);
        }
    }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const u8 *c = pad + ((long)(y + radius) * PW + (x + radius)) * 3;
            int b0 = c[0], g0 = c[1], r0 = c[2];
            float wsum = 0.f, sb = 0.f, sg = 0.f, sr = 0.f;
            for (int k = 0; k < maxk; k++) {
                const u8 *q = c + ((long)dy[k] * PW + dx[k]) * 3;
                int b = q[0], g = q[1], r = q[2];
                float w = space[k] * color[abs(b - b0) + abs(g - g0) + abs(r - r0)];
                if (use_fma) {
                    wsum += w;
                    sb = fmaf((float)b, w, sb); sg = fmaf((float)g, w, sg); sr = fmaf((float)r, w, sr);
                } else {
                    float tb = (float)b * w, tg = (float)g * w, tr = (float)r * w;
                    wsum += w; sb += tb; sg += tg; sr += tr;
                }
            }
            float inv = 1.0f / wsum;
            u8 *o = dst + ((long)y * W + x) * 3;
            o[0] = (u8)sat_u8((int)rintf(sb * inv));
            o[1] = (u8)sat_u8((int)rintf(sg * inv));
            o[2] = (u8)sat_u8((int)rintf(sr * inv));
        }
    free(pad);
}

/* ------------------------------------------------------------------------ */
/* S5  3x3 sharpen            frame_enhancer.py:40-42,138                    */
/* OpenCV: filter2D, kernel [[-1,-1,-1],[-1,9,-1],[-1,-1,-1]], REFLECT_101  */
/* ------------------------------------------------------------------------ */
ORC_API void orc_sharpen(const u8 *src, int H, int W, int C, u8 *dst)
{
    for (int y = 0; y < H; y++) {
        int ym = reflect101(y - 1, H), yp = reflect101(y + 1, H);
        for (int x = 0; x < W; x++) {
            int xm = reflect101(x - 1, W), xp = reflect101(x + 1, W);
            for (int c = 0; c < C; c++) {
#define PX(yy, xx) ((int)src[((long)(yy) * W + (xx)) * C + c])
                int s = PX(ym, xm) + PX(ym, x) + PX(ym, xp) + PX(y, xm) + PX(y, x) + PX(y, xp)
                      + PX(yp, xm) + PX(yp, x) + PX(yp, xp);
                dst[((long)y * W + x) * C + c] = (u8)sat_u8(10 * PX(y, x) - s);
#undef PX
            }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* S6  normalize MINMAX 0..255   frame_enhancer.py:146                       */
/* OpenCV: core/src/norm.cpp cv::normalize + convert_scale.simd.hpp (8u->8u */
/* goes through f32: v*(float)scale + (float)shift, fused on FMA3 hosts)    */
/* ------------------------------------------------------------------------ */
ORC_API void orc_minmax(const u8 *src, long n, int *mn, int *mx)
{
    int lo = 255, hi = 0;
    for (long i = 0; i < n; i++) { int v = src[i]; if (v < lo) lo = v; if (v > hi) hi = v; }
    *mn = lo; *mx = hi;
}
ORC_API void orc_normalize_lut(int smin, int smax, int use_fma, u8 *lut256)
{
    double scale = (255.0 - 0.0) * ((double)(smax - smin) > DBL_EPSILON ? 1.0 / (double)(smax - smin) : 0.0);
    double shift = 0.0 - (double)smin * scale;
    float a = (float)scale, b = (float)shift;
    for (int v = 0; v < 256; v++) {
        float r;
        if (use_fma) r = fmaf((float)v, a, b);
        else { float t = (float)v * a; r = t + b; }
        lut256[v] = (u8)sat_u8((int)rintf(r));
    }
}
ORC_API void orc_normalize(const u8 *src, long n, int use_fma, u8 *dst, int *mn_out, int *mx_out)
{
    int mn, mx; u8 lut[256];
    orc_minmax(src, n, &mn, &mx);
    orc_normalize_lut(mn, mx, use_fma, lut);
    for (long i = 0; i < n; i++) dst[i] = lut[src[i]];
    if (mn_out) *mn_out = mn;
    if (mx_out) *mx_out = mx;
}

/* ------------------------------------------------------------------------ */
/* S7  BGR -> GRAY   frame_enhancer.py:154, change_detector.py:51,          */
/*                   piece_detector.py:128                                  */
/* OpenCV: color_rgb.simd.hpp RGB2Gray<uchar> (15-bit coefficients)         */
/* ------------------------------------------------------------------------ */
ORC_API void orc_gray(const u8 *bgr, int H, int W, long stride, u8 *dst)
{
    for (int y = 0; y < H; y++) {
        const u8 *p = bgr + (long)y * stride;
        for (int x = 0; x < W; x++)
            dst[(long)y * W + x] = (u8)((3735 * p[3 * x] + 19235 * p[3 * x + 1] + 9798 * p[3 * x + 2] + 16384) >> 15);
    }
}

/* ------------------------------------------------------------------------ */
/* S8  GaussianBlur((k,k),0) on u8    frame_enhancer.py:156,                 */
/*     change_detector.py:55-56, piece_detector.py:133                      */
/* OpenCV: smooth.dispatch.cpp getGaussianKernelBitExact +                  */
/*   getGaussianKernelFixedPoint_ED (Q8, error diffusion), fixedpoint.inl,  */
/*   REFLECT_101 of the array being blurred                                 */
/* ------------------------------------------------------------------------ */
/* sigma <= 0: OpenCV's default for the size (fixed small kernels up to 7, else sigma from k) */
ORC_API int orc_gaussian_kernel_q8_sigma(int k, double sigma, int *q);
ORC_API int orc_gaussian_kernel_q8(int k, int *q) { return orc_gaussian_kernel_q8_sigma(k, 0.0, q); }
ORC_API int orc_gaussian_kernel_q8_sigma(int k, double sigma_in, int *q)
{
    if (k < 1 || k > 31 || (k & 1) == 0) return -1;
    double kern[31];
    static const double t1[] = {1.0};
    static const double t3[] = {0.25, 0.5, 0.25};
    static const double t5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625};
    static const double t7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
    if (sigma_in <= 0 && k == 1) memcpy(kern, t1, sizeof t1);
    else if (sigma_in <= 0 && k == 3) memcpy(kern, t3, sizeof t3);
    else if (sigma_in <= 0 && k == 5) memcpy(kern, t5, sizeof t5);
    else if (sigma_in <= 0 && k == 7) memcpy(kern, t7, sizeof t7);
    else {
        double sigma = sigma_in > 0 ? sigma_in : ((k - 1) * 0.5 - 1) * 0.3 + 0.8;
        double scale2x = -0.5 / (sigma * sigma), sum = 0;
        for (int i = 0; i < k; i++) {
            double x = i - (k - 1) * 0.5;
            kern[i] = exp(scale2x * x * x);
            sum += kern[i];
        }
        sum = 1.0 / sum;
        for (int i = 0; i < k; i++) kern[i] *= sum;
    }
    double err = 0; long s = 0;
    for (int i = 0; i < k / 2; i++) {
        double adj = kern[i] * 256.0 + err;
        long v = (long)rint(adj);
        err = adj - (double)v;
        q[i] = (int)v; q[k - 1 - i] = (int)v; s += v;
    }
    q[k / 2] = (int)(256 - 2 * s);
    return 0;
}

ORC_API int orc_gaussian_sigma(const u8 *src, int H, int W, long stride, int k, double sigma, u8 *dst);
ORC_API int orc_gaussian(const u8 *src, int H, int W, long stride, int k, u8 *dst) { return orc_gaussian_sigma(src, H, W, stride, k, 0.0, dst); }
ORC_API int orc_gaussian_sigma(const u8 *src, int H, int W, long stride, int k, double sigma, u8 *dst)
{
    int q[31];
    if (orc_gaussian_kernel_q8_sigma(k, sigma, q)) return -1;
    int r = k / 2;
    uint32_t *tmp = (uint32_t *)malloc(sizeof(uint32_t) * (long)H * W);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint32_t s = 0;
            for (int j = 0; j < k; j++) s += (uint32_t)q[j] * src[(long)y * stride + reflect101(x + j - r, W)];
            tmp[(long)y * W + x] = s;   /* Q8.8, <= 255*256 */
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            uint32_t s = 0;
            for (int i = 0; i < k; i++) s += (uint32_t)q[i] * tmp[(long)reflect101(y + i - r, H) * W + x];
            dst[(long)y * W + x] = (u8)((s + 32768u) >> 16);
        }
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* S9  Otsu                    frame_enhancer.py:158                         */
/* OpenCV: imgproc/src/thresh.cpp getThreshVal_Otsu_8u                      */
/* ------------------------------------------------------------------------ */
ORC_API void orc_hist256(const u8 *src, long n, int32_t *h)
{
    memset(h, 0, 256 * sizeof(int32_t));
    for (long i = 0; i < n; i++) h[src[i]]++;
}
ORC_API int orc_otsu_from_hist(const int32_t *h, long n)
{
    double mu = 0, scale = 1.0 / (double)n;
    for (int i = 0; i < 256; i++) mu += i * (double)h[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0;
    int max_val = 0;
    for (int i = 0; i < 256; i++) {
        double p_i = h[i] * scale, q2, mu2, sigma;
        mu1 *= q1;
        q1 += p_i;
        q2 = 1.0 - q1;
        double lo = q1 < q2 ? q1 : q2, hi = q1 > q2 ? q1 : q2;
        if (lo < FLT_EPSILON || hi > 1.0 - FLT_EPSILON) continue;
        mu1 = (mu1 + i * p_i) / q1;
        mu2 = (mu - q1 * mu1) / q2;
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return max_val;
}
ORC_API void orc_threshold(const u8 *src, long n, int T, u8 *dst)
{
    for (long i = 0; i < n; i++) dst[i] = src[i] > T ? 255 : 0;
}
/* prepare_analysis (frame_enhancer.py:148-159): gray, blur5, Otsu, mask */
ORC_API int orc_prepare_analysis(const u8 *bgr, int H, int W, u8 *gray, u8 *blurred, u8 *binary)
{
    long n = (long)H * W;
    int32_t h[256];
    u8 *bl = blurred ? blurred : (u8 *)malloc(n);
    orc_gray(bgr, H, W, 3L * W, gray);
    orc_gaussian(gray, H, W, W, 5, bl);
    orc_hist256(bl, n, h);
    int T = orc_otsu_from_hist(h, n);
    orc_threshold(bl, n, T, binary);
    if (!blurred) free(bl);
    return T;
}

/* process_pipeline without the colour profile (frame_enhancer.py:161-181) */
ORC_API void orc_process_pipeline(const u8 *bgr, int H, int W, int use_fma, u8 *out)
{
    long n3 = 3L * H * W;
    u8 *a = (u8 *)malloc(n3), *b = (u8 *)malloc(n3);
    orc_correct_lighting(bgr, H, W, 3.0, 8, 8, a);
    orc_bilateral(a, H, W, 9, 75.0, 75.0, use_fma, b);
    orc_sharpen(b, H, W, 3, a);
    orc_normalize(a, n3, use_fma, out, NULL, NULL);
    free(a); free(b);
}

/* ------------------------------------------------------------------------ */
/* G1  perspective warp        board_detection.py:61-71                      */
/* OpenCV: imgwarp.cpp getPerspectiveTransform (8x8 LU), cv::invert 3x3,    */
/*   WarpPerspectiveInvoker (64x16 blocks, f64 coords, Q5) + remapBilinear  */
/* ------------------------------------------------------------------------ */
static int lu_solve(double *A, int m, double *b)
{
    /* core/src/matrix_decomp.cpp LUImpl, partial pivoting */
    for (int i = 0; i < m; i++) {
        int k = i;
        for (int j = i + 1; j < m; j++)
            if (fabs(A[j * m + i]) > fabs(A[k * m + i])) k = j;
        if (fabs(A[k * m + i]) < DBL_EPSILON * 100) return 0;
        if (k != i) {
            for (int j = i; j < m; j++) { double t = A[i * m + j]; A[i * m + j] = A[k * m + j]; A[k * m + j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; j++) {
            double alpha = A[j * m + i] * d;
            for (int c = i + 1; c < m; c++) A[j * m + c] += alpha * A[i * m + c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; i--) {
        double s = b[i];
        for (int c = i + 1; c < m; c++) s -= A[i * m + c] * b[c];
        b[i] = s / A[i * m + i];
    }
    return 1;
}
/* src/dst: 4 points (x,y) as float32 values, order as given by the caller */
ORC_API int orc_get_perspective(const float *src, const float *dst, double *M)
{
    double A[64], b[8];
    memset(A, 0, sizeof A);
    for (int i = 0; i < 4; i++) {
        /* Point2f operands: the products are rounded to f32 before widening */
        float sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
        A[i * 8 + 0] = A[(i + 4) * 8 + 3] = sx;
        A[i * 8 + 1] = A[(i + 4) * 8 + 4] = sy;
        A[i * 8 + 2] = A[(i + 4) * 8 + 5] = 1;
        A[i * 8 + 6] = (float)(-sx * dx); A[i * 8 + 7] = (float)(-sy * dx);
        A[(i + 4) * 8 + 6] = (float)(-sx * dy); A[(i + 4) * 8 + 7] = (float)(-sy * dy);
        b[i] = dx; b[i + 4] = dy;
    }
    if (!lu_solve(A, 8, b)) return -1;
    for (int i = 0; i < 8; i++) M[i] = b[i];
    M[8] = 1.0;
    return 0;
}
ORC_API int orc_invert3(const double *a, double *t)
{
    double d = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) + a[2] * (a[3] * a[7] - a[4] * a[6]);
    if (d == 0.0) return -1;
    d = 1.0 / d;
    t[0] = (a[4] * a[8] - a[5] * a[7]) * d;
    t[1] = (a[2] * a[7] - a[1] * a[8]) * d;
    t[2] = (a[1] * a[5] - a[2] * a[4]) * d;
    t[3] = (a[5] * a[6] - a[3] * a[8]) * d;
    t[4] = (a[0] * a[8] - a[2] * a[6]) * d;
    t[5] = (a[2] * a[3] - a[0] * a[5]) * d;
    t[6] = (a[3] * a[7] - a[4] * a[6]) * d;
    t[7] = (a[1] * a[6] - a[0] * a[7]) * d;
    t[8] = (a[0] * a[4] - a[1] * a[3]) * d;
    return 0;
}
/* Minv maps dst -> src.  dst is S_h x S_w x 3. */
ORC_API void orc_warp(const u8 *src, int H, int W, const double *Mi, int SH, int SW, u8 *dst)
{
    for (int y = 0; y < SH; y++)
        for (int x = 0; x < SW; x++) {
            int bx = x & ~63, x1 = x & 63;    /* 64-wide column blocks of the invoker */
            double X0 = Mi[0] * bx + Mi[1] * y + Mi[2];
            double Y0 = Mi[3] * bx + Mi[4] * y + Mi[5];
            double W0 = Mi[6] * bx + Mi[7] * y + Mi[8];
            double Wd = W0 + Mi[6] * x1;
            Wd = Wd != 0.0 ? 32.0 / Wd : 0.0;
            double fX = (X0 + Mi[0] * x1) * Wd, fY = (Y0 + Mi[3] * x1) * Wd;
            if (fX < -2147483648.0) fX = -2147483648.0; if (fX > 2147483647.0) fX = 2147483647.0;
            if (fY < -2147483648.0) fY = -2147483648.0; if (fY > 2147483647.0) fY = 2147483647.0;
            int X = (int)rint(fX), Y = (int)rint(fY);
            int sx = X >> 5, sy = Y >> 5, ax = X & 31, ay = Y & 31;
            /* xy is stored as short in OpenCV */
            if (sx < -32768) sx = -32768; if (sx > 32767) sx = 32767;
            if (sy < -32768) sy = -32768; if (sy > 32767) sy = 32767;
            int w00 = (32 - ax) * (32 - ay), w01 = ax * (32 - ay), w10 = (32 - ax) * ay, w11 = ax * ay;
            for (int c = 0; c < 3; c++) {
#define P(yy, xx) (((yy) >= 0 && (yy) < H && (xx) >= 0 && (xx) < W) ? (int)src[((long)(yy) * W + (xx)) * 3 + c] : 0)
                int v = w00 * P(sy, sx) + w01 * P(sy, sx + 1) + w10 * P(sy + 1, sx) + w11 * P(sy + 1, sx + 1);
#undef P
                dst[((long)y * SW + x) * 3 + c] = (u8)((v + 512) >> 10);
            }
        }
}

/* ------------------------------------------------------------------------ */
/* Per-square statistics                                                    */
/* ------------------------------------------------------------------------ */
/* square preprocessing: gray (if 3 channels) + Gaussian k, borders         */
/* reflected inside the square (change_detector.py:49-56,                   */
/* piece_detector.py:124-135)                                               */
ORC_API int orc_square_preprocess(const u8 *sq, int h, int w, int channels, long stride, int k, u8 *out)
{
    u8 *g = (u8 *)malloc((long)h * w);
    if (channels == 3) orc_gray(sq, h, w, stride, g);
    else for (int y = 0; y < h; y++) memcpy(g + (long)y * w, sq + (long)y * stride, w);
    int rc = orc_gaussian(g, h, w, w, k, out);
    free(g);
    return rc;
}

/* mask bits for one (h,w) square shape: bit0 centre disc, bit1 corner      */
/* blocks (piece_detector.py:184-198), bits 2..5 the four rings             */
/* (piece_detector.py:148-163).                                             */
ORC_API void orc_square_masks(int h, int w, u8 *mask)
{
    int cy = h / 2, cx = w / 2, md = h < w ? h : w;
    int radius = md / 4, cs = md / 4;
    static const double rr[4] = {0.15, 0.25, 0.35, 0.45};
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            int m = 0;
            int d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
            if (d2 <= radius * radius) m |= 1;
            /* numpy slices [:cs] and [-cs:]; cs == 0 makes [-0:] the whole axis */
            int top = y < cs, left = x < cs;
            int bottom = cs == 0 ? 1 : (y >= h - cs), right = cs == 0 ? 1 : (x >= w - cs);
            if ((top && left) || (top && right) || (bottom && left) || (bottom && right)) m |= 2;
            double dist = sqrt((double)d2);
            for (int k = 0; k < 4; k++) {
                double r = md * rr[k];
                if (dist >= r - 5 && dist <= r + 5) m |= (4 << k);
            }
            mask[(long)y * w + x] = (u8)m;
        }
}

typedef struct {
    int64_t n;            /* pixels                                         */
    int64_t sum;          /* sum g                                          */
    int64_t sumsq;        /* sum g^2                                        */
    int64_t sad;          /* sum |g - ref|  (-1 when no reference)          */
    int64_t center_sum, center_cnt, border_sum, border_cnt;
    int64_t ring_sum[4], ring_cnt[4];
} orc_pd_stats;

/* g: preprocessed square (h*w), ref: same shape or NULL */
ORC_API void orc_pd_square_stats(const u8 *g, const u8 *ref, int h, int w, orc_pd_stats *st)
{
    u8 *mask = (u8 *)malloc((long)h * w);
    orc_square_masks(h, w, mask);
    memset(st, 0, sizeof *st);
    st->n = (int64_t)h * w;
    st->sad = ref ? 0 : -1;
    for (long i = 0; i < (long)h * w; i++) {
        int v = g[i], m = mask[i];
        st->sum += v; st->sumsq += v * v;
        if (ref) st->sad += abs(v - (int)ref[i]);
        if (m & 1) { st->center_sum += v; st->center_cnt++; }
        if (m & 2) { st->border_sum += v; st->border_cnt++; }
        for (int k = 0; k < 4; k++) if (m & (4 << k)) { st->ring_sum[k] += v; st->ring_cnt[k]++; }
    }
    free(mask);
}

/* ChangeDetector numerics (change_detector.py:77-92, 121-137): all f32,    */
/* every multiply/add rounded separately (NumPy semantics).                 */
ORC_API void orc_cd_calibrate(const u8 *g, long n, float initial_variance, float *mean, float *var)
{
    for (long i = 0; i < n; i++) { mean[i] = (float)g[i]; var[i] = initial_variance; }
}
ORC_API void orc_cd_update(const u8 *g, long n, float alpha, float one_minus_alpha, float *mean, float *var)
{
    for (long i = 0; i < n; i++) {
        float gr = (float)g[i];
        float t0 = one_minus_alpha * mean[i]; float t1 = alpha * gr; float nm = t0 + t1;
        float diff = gr - nm; float d2 = diff * diff;
        float t2 = one_minus_alpha * var[i]; float t3 = alpha * d2; float nv = t2 + t3;
        if (!(nv > 10.0f)) nv = nv != nv ? nv : 10.0f;  /* np.maximum */
        mean[i] = nm; var[i] = nv;
    }
}
ORC_API void orc_cd_detect(const u8 *g, long n, const float *mean, const float *var, float z_threshold,
                           int64_t *changed, float *zmax)
{
    int64_t cnt = 0; float zm = -INFINITY; int have = 0;
    for (long i = 0; i < n; i++) {
        float sd = sqrtf(var[i]);
        float diff = fabsf((float)g[i] - mean[i]);
        float z = diff / sd;
        if (z > z_threshold) cnt++;
        /* np.max propagates NaN */
        if (!have) { zm = z; have = 1; }
        else if (z != z) zm = z;
        else if (!(zm != zm) && z > zm) zm = z;
    }
    *changed = cnt; *zmax = zm;
}

/* ------------------------------------------------------------------------ */
/* S0  colour profile          frame_enhancer.py:56-99  ("next" scope row)   */
/* OpenCV: convert_scale.simd.hpp cvtabs (f32 fma, |.|, round), color_hsv   */
/* RGB2HSV_b (integer, hdiv/sdiv tables), HSV2RGB_b (f32 with fnma for the   */
/* two hue-dependent terms; the 32-pixel vector body truncates, the scalar  */
/* row tail rounds -- measured against cv2 4.13 on all 180*256*256 inputs);  */
/* NumPy f32 arithmetic in between.                                          */
/* ------------------------------------------------------------------------ */
typedef struct {
    double contrast, brightness;      /* convertScaleAbs alpha / beta          */
    float hue_shift, sat_scale, val_scale;
    int radical_mode;
    float target_hue, hue_window;
    int simd_block;                    /* pixels per vector block of HSV2BGR (32 on AVX2 builds); 0 = all scalar */
} orc_color_profile;

static int g_hsv_ready = 0;
static int32_t g_sdiv[256], g_hdiv[256];
static void hsv_tables_init(void)
{
    if (g_hsv_ready) return;
    g_sdiv[0] = g_hdiv[0] = 0;
    for (int i = 1; i < 256; i++) {
        g_sdiv[i] = (int32_t)rint((255 << 12) / (1. * i));
        g_hdiv[i] = (int32_t)rint((180 << 12) / (6. * i));
    }
    g_hsv_ready = 1;
}
ORC_API void orc_convert_scale_abs(const u8 *src, long n, double alpha, double beta, u8 *dst)
{
    const float a = (float)alpha, b = (float)beta;
    for (long i = 0; i < n; i++) dst[i] = (u8)sat_u8((int)rintf(fabsf(fmaf((float)src[i], a, b))));
}
static inline void bgr2hsv_px(const u8 *p, u8 *o)
{
    int b = p[0], g = p[1], r = p[2];
    int v = b > g ? b : g; if (r > v) v = r;
    int vmin = b < g ? b : g; if (r < vmin) vmin = r;
    int diff = v - vmin;
    int s = (diff * g_sdiv[v] + (1 << 11)) >> 12;
    int h = v == r ? g - b : (v == g ? b - r + 2 * diff : r - g + 4 * diff);
    h = (h * g_hdiv[diff] + (1 << 11)) >> 12;
    if (h < 0) h += 180;
    o[0] = (u8)sat_u8(h); o[1] = (u8)s; o[2] = (u8)v;
}
ORC_API void orc_bgr2hsv(const u8 *bgr, long n_px, u8 *hsv)
{
    hsv_tables_init();
    for (long i = 0; i < n_px; i++) bgr2hsv_px(bgr + 3 * i, hsv + 3 * i);
}
static inline void hsv2bgr_px(const u8 *p, u8 *o, int vector_body)
{
    static const int sector_data[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
    float h = (float)p[0], s = (float)p[1] * (1.f / 255.f), v = (float)p[2] * (1.f / 255.f);
    float tab[4];
    h = h * (6.f / 180.f);
    float fl = floorf(h);
    int sector = (int)fl;
    h = h - fl;
    sector %= 6; if (sector < 0) sector += 6;
    tab[0] = v;
    { float t = 1.f - s; tab[1] = v * t; }
    /* both the vector body and the (compiler-contracted) scalar tail of the AVX2 build form 1 - s*h with one rounding */
    tab[2] = v * fmaf(-s, h, 1.f);
    { float omh = 1.f - h; tab[3] = v * fmaf(-s, omh, 1.f); }
    for (int c = 0; c < 3; c++) {
        float x = tab[sector_data[sector][c]] * 255.f;
        /* the vector body truncates (v_trunc), the scalar tail rounds (saturate_cast -> cvRound) */
        int q = vector_body ? (int)x : (int)rintf(x);
        o[c] = (u8)sat_u8(q);
    }
}
/* row-wise: the first (W / block) * block pixels of a row take the vector formula */
ORC_API void orc_hsv2bgr(const u8 *hsv, int H, int W, int simd_block, u8 *bgr)
{
    const int nvec = simd_block > 0 ? (W / simd_block) * simd_block : 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
            hsv2bgr_px(hsv + ((long)y * W + x) * 3, bgr + ((long)y * W + x) * 3, x < nvec);
}
/* NumPy part of apply_color_profile (frame_enhancer.py:75-97) on one pixel, f32 throughout */
static inline void hsv_adjust_px(const u8 *p, const orc_color_profile *c, u8 *o)
{
    float h = (float)p[0], s = (float)p[1], v = (float)p[2];
    if (c->radical_mode) {
        float hd = fabsf(h - c->target_hue);
        float alt = 180.f - hd;
        if (alt < hd) hd = alt;
        s = hd < c->hue_window ? s * 2.0f : s * 0.5f;
    }
    float hs = h + c->hue_shift;
    float m = fmodf(hs, 180.f);
    if (m != 0.f) { if (m < 0.f) m += 180.f; } else m = 0.f;
    h = m;
    s = s * c->sat_scale;
    v = v * c->val_scale;
    if (h < 0.f) h = 0.f; if (h > 179.f) h = 179.f;
    if (s < 0.f) s = 0.f; if (s > 255.f) s = 255.f;
    if (v < 0.f) v = 0.f; if (v > 255.f) v = 255.f;
    o[0] = (u8)(int)h; o[1] = (u8)(int)s; o[2] = (u8)(int)v;
}
ORC_API void orc_apply_color_profile(const u8 *bgr, int H, int W, const orc_color_profile *c, u8 *out)
{
    hsv_tables_init();
    const float a = (float)c->contrast, b = (float)c->brightness;
    const int nvec = c->simd_block > 0 ? (W / c->simd_block) * c->simd_block : 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const u8 *p = bgr + ((long)y * W + x) * 3;
            u8 t[3], hsv[3], adj[3];
            for (int k = 0; k < 3; k++) t[k] = (u8)sat_u8((int)rintf(fabsf(fmaf((float)p[k], a, b))));
            bgr2hsv_px(t, hsv);
            hsv_adjust_px(hsv, c, adj);
            hsv2bgr_px(adj, out + ((long)y * W + x) * 3, x < nvec);
        }
}

/* ------------------------------------------------------------------------ */
/* Canny + grid refinement     grid_extractor.py:66-121 ("next" scope row)   */
/* OpenCV imgproc/src/canny.cpp, 8-bit input, aperture 3, L1 magnitude:     */
/* Sobel 3x3 with BORDER_REPLICATE, |dx|+|dy|, non-maximum suppression with */
/* tan(22.5deg) in Q15 (13573), zeros outside the image, hysteresis = every  */
/* candidate 8-connected to a candidate above the high threshold.           */
/* ------------------------------------------------------------------------ */
ORC_API void orc_canny(const u8 *src, int H, int W, double low_thresh, double high_thresh, u8 *dst)
{
    if (low_thresh > high_thresh) { double t = low_thresh; low_thresh = high_thresh; high_thresh = t; }
    const int low = (int)floor(low_thresh), high = (int)floor(high_thresh);
    const long n = (long)H * W;
    int16_t *dx = (int16_t *)malloc(n * 2), *dy = (int16_t *)malloc(n * 2);
    int32_t *mag = (int32_t *)calloc((long)(H + 2) * (W + 2), 4);
    u8 *map = (u8 *)malloc(n);
#define S(yy, xx) ((int)src[(long)((yy) < 0 ? 0 : (yy) >= H ? H - 1 : (yy)) * W + ((xx) < 0 ? 0 : (xx) >= W ? W - 1 : (xx))])
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int gx = (S(y - 1, x + 1) + 2 * S(y, x + 1) + S(y + 1, x + 1)) - (S(y - 1, x - 1) + 2 * S(y, x - 1) + S(y + 1, x - 1));
            int gy = (S(y + 1, x - 1) + 2 * S(y + 1, x) + S(y + 1, x + 1)) - (S(y - 1, x - 1) + 2 * S(y - 1, x) + S(y - 1, x + 1));
            dx[(long)y * W + x] = (int16_t)gx; dy[(long)y * W + x] = (int16_t)gy;
            mag[(long)(y + 1) * (W + 2) + x + 1] = abs(gx) + abs(gy);
        }
#undef S
    long *stack = (long *)malloc(n * sizeof(long)), sp = 0;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const int32_t *mp = mag + (long)(y + 1) * (W + 2) + x + 1;
            const int m = mp[0];
            int cand = 0;
            if (m > low) {
                const int xs = dx[(long)y * W + x], ys = dy[(long)y * W + x];
                const long ax = abs(xs), ay = (long)abs(ys) << 15;
                const long tg22x = ax * 13573, tg67x = tg22x + (ax << 16);
                if (ay < tg22x) cand = m > mp[-1] && m >= mp[1];
                else if (ay > tg67x) cand = m > mp[-(W + 2)] && m >= mp[W + 2];
                else { const int s = (xs ^ ys) < 0 ? -1 : 1; cand = m > mp[-(W + 2) - s] && m > mp[(W + 2) + s]; }
            }
            u8 v = 1;                      /* 1: not an edge, 0: weak candidate, 2: edge */
            if (cand) { v = m > high ? 2 : 0; if (v == 2) stack[sp++] = (long)y * W + x; }
            map[(long)y * W + x] = v;
        }
    while (sp) {
        const long i = stack[--sp];
        const int y = (int)(i / W), x = (int)(i % W);
        for (int yy = y - 1; yy <= y + 1; yy++)
            for (int xx = x - 1; xx <= x + 1; xx++)
                if (yy >= 0 && yy < H && xx >= 0 && xx < W && map[(long)yy * W + xx] == 0) {
                    map[(long)yy * W + xx] = 2; stack[sp++] = (long)yy * W + xx;
                }
    }
    for (long i = 0; i < n; i++) dst[i] = map[i] == 2 ? 255 : 0;
    free(dx); free(dy); free(mag); free(map); free(stack);
}

/* ------------------------------------------------------------------------ */
/* cv2.HoughCircles(gray, HOUGH_GRADIENT, dp, minDist, param1, param2,        */
/* minRadius, maxRadius) -- reference call piece_detector.py:232-241.         */
/* OpenCV imgproc/src/hough.cpp (HoughCirclesGradient), restated:             */
/*  1. Sobel 3x3 (replicate border) and Canny(dx, dy, max(1, param1/2),       */
/*     param1), L1 magnitude;                                                 */
/*  2. every edge pixel with a non-zero gradient votes along +-gradient for   */
/*     the radii minRadius..maxRadius in an accumulator of 1/dp resolution    */
/*     (Q10 fixed-point ray, stops at the first step outside);                */
/*  3. centres: accumulator cells (not in row 0 / column 0) > param2 that    */
/*     are local maxima (> left, >= right, > up, >= down);                    */
/*  4. per centre: distances to all edge pixels, histogram with 10 bins per   */
/*     dp, best 10-bin window scanning downwards -> (radius, support);        */
/*     kept when support > param2;                                            */
/*  5. sort by (support desc, radius desc, x asc, y asc); greedy removal of   */
/*     circles closer than minDist to a kept one.                             */
/* All float steps are f32 and unfused, as in the SSE-baseline build.         */
/* out: up to max_out rows of (x, y, r, support); returns the circle count    */
/* (may exceed max_out; only the first max_out are stored), -1 on bad input.  */
/* ------------------------------------------------------------------------ */
typedef struct { float x, y, r; int support; } orc_circle;
static int circle_before(const orc_circle *a, const orc_circle *b)
{
    if (a->support != b->support) return a->support > b->support;
    if (a->r != b->r) return a->r > b->r;
    if (a->x != b->x) return a->x < b->x;
    return a->y < b->y;
}
static int circle_cmp(const void *pa, const void *pb)
{
    const orc_circle *a = (const orc_circle *)pa, *b = (const orc_circle *)pb;
    return circle_before(a, b) ? -1 : circle_before(b, a) ? 1 : 0;
}
ORC_API int orc_hough_circles(const u8 *gray, int H, int W, long stride, double dp_d, double min_dist_d,
                              double param1, double param2, int min_radius, int max_radius,
                              float *out, int max_out, int32_t *info)
{
    if (H < 1 || W < 1 || dp_d <= 0 || min_dist_d <= 0 || param1 <= 0 || param2 <= 0) return -1;
    const float dp = (float)dp_d < 1.f ? 1.f : (float)dp_d;
    const float min_dist = (float)min_dist_d;
    const int canny_thr = (int)lrint(param1), acc_thr = (int)lrint(param2);
    if (min_radius < 0) min_radius = 0;
    if (max_radius <= 0) max_radius = H > W ? H : W;
    else if (max_radius <= min_radius) max_radius = min_radius + 2;
    const long n = (long)H * W;
    u8 *g = (u8 *)malloc(n), *edges = (u8 *)malloc(n);
    for (int y = 0; y < H; y++) memcpy(g + (long)y * W, gray + y * stride, W);
    orc_canny(g, H, W, canny_thr / 2 > 1 ? canny_thr / 2 : 1, canny_thr, edges);
    const float idp = 1.f / dp;
    const int arows = (int)ceilf(H * idp), acols = (int)ceilf(W * idp), astep = acols + 2;
    int32_t *acc = (int32_t *)calloc((long)(arows + 2) * astep, 4);
    int32_t *nz = (int32_t *)malloc(n * 4);
    long nnz = 0;
    const int ONE = 1 << 10;
#define S(yy, xx) ((int)g[(long)((yy) < 0 ? 0 : (yy) >= H ? H - 1 : (yy)) * W + ((xx) < 0 ? 0 : (xx) >= W ? W - 1 : (xx))])
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            if (!edges[(long)y * W + x]) continue;
            const int gx = (S(y - 1, x + 1) + 2 * S(y, x + 1) + S(y + 1, x + 1)) - (S(y - 1, x - 1) + 2 * S(y, x - 1) + S(y + 1, x - 1));
            const int gy = (S(y + 1, x - 1) + 2 * S(y + 1, x) + S(y + 1, x + 1)) - (S(y - 1, x - 1) + 2 * S(y - 1, x) + S(y - 1, x + 1));
            if (gx == 0 && gy == 0) continue;
            const float vx = (float)gx, vy = (float)gy;
            const float mag = sqrtf(vx * vx + vy * vy);
            if (mag < 1.0f) continue;
            nz[nnz++] = (y << 16) | x;
            int sx = (int)lrintf((vx * idp) * ONE / mag), sy = (int)lrintf((vy * idp) * ONE / mag);
            const int x0 = (int)lrintf((x * idp) * ONE), y0 = (int)lrintf((y * idp) * ONE);
            for (int k = 0; k < 2; k++) {
                int x1 = x0 + min_radius * sx, y1 = y0 + min_radius * sy;
                for (int r = min_radius; r <= max_radius; x1 += sx, y1 += sy, r++) {
                    const int x2 = x1 >> 10, y2 = y1 >> 10;
                    if ((unsigned)x2 >= (unsigned)acols || (unsigned)y2 >= (unsigned)arows) break;
                    acc[(long)(y2 + 1) * astep + x2 + 1]++;
                }
                sx = -sx; sy = -sy;
            }
        }
#undef S
    int count = 0, ncenters = 0;
    orc_circle *est = NULL;
    if (nnz) {
        const float dr = dp;
        const int per = 10;
        const int nbins = (int)lrintf((max_radius - min_radius) / dr * per);
        int *bins = (int *)malloc((nbins > 0 ? nbins : 1) * sizeof(int));
        const float minr2 = (float)min_radius * min_radius, maxr2 = (float)max_radius * max_radius;
        long cap = 64, nest = 0;
        est = (orc_circle *)malloc(cap * sizeof(orc_circle));
        /* hough.cpp votes into the padded accumulator without an offset and scans its interior, so
         * the cells of accumulator row 0 / column 0 are neighbours only, never centres */
        for (int ay = 2; ay <= arows; ay++)
            for (int ax = 2; ax <= acols; ax++) {
                const int32_t *a = acc + (long)ay * astep + ax;
                if (!(a[0] > acc_thr && a[0] > a[-1] && a[0] >= a[1] && a[0] > a[-astep] && a[0] >= a[astep])) continue;
                ncenters++;
                const float cx = ((ax - 1) + 0.5f) * dr, cy = ((ay - 1) + 0.5f) * dr;
                int max_count = 0; float rbest = 0;
                for (int i = 0; i < nbins; i++) bins[i] = 0;
                int used = 0;
                for (long k = 0; k < nnz && nbins > 0; k++) {
                    const float ddx = cx - (float)(nz[k] & 0xffff), ddy = cy - (float)(nz[k] >> 16);
                    const float a2 = ddx * ddx, b2 = ddy * ddy, r2 = a2 + b2;
                    if (minr2 <= r2 && r2 <= maxr2) {
                        int b = (int)lrintf((sqrtf(r2) - min_radius) / dr * per);
                        b = b < 0 ? 0 : b > nbins - 1 ? nbins - 1 : b;
                        bins[b]++; used++;
                    }
                }
                if (used)
                    for (int j = nbins - 1; j > 0; j--)
                        if (bins[j]) {
                            const int up = j; int cur = 0;
                            for (; j > up - per && j >= 0; j--) cur += bins[j];
                            const float rcur = (up + j) / 2.f / per * dr + min_radius;
                            if ((cur * rbest >= max_count * rcur) || (rbest < FLT_EPSILON && cur >= max_count)) { rbest = rcur; max_count = cur; }
                        }
                if (max_count > acc_thr) {
                    if (nest == cap) { cap *= 2; est = (orc_circle *)realloc(est, cap * sizeof(orc_circle)); }
                    est[nest].x = cx; est[nest].y = cy; est[nest].r = rbest; est[nest].support = max_count; nest++;
                }
            }
        qsort(est, nest, sizeof(orc_circle), circle_cmp);
        /* RemoveOverlaps: in place, keeps the first of any two closer than minDist */
        long kept = nest ? 1 : 0;
        for (long i = 1; i < nest; i++) {
            int ok = 1;
            for (long k = 0; k < kept && ok; k++) {
                const float ex = est[k].x - est[i].x, ey = est[k].y - est[i].y;
                if (ex * ex + ey * ey < min_dist * min_dist) ok = 0;
            }
            if (ok) est[kept++] = est[i];
        }
        count = (int)kept;
        for (long i = 0; i < kept && i < max_out; i++) {
            out[4 * i] = est[i].x; out[4 * i + 1] = est[i].y; out[4 * i + 2] = est[i].r; out[4 * i + 3] = (float)est[i].support;
        }
        free(bins);
    }
    if (info) { info[0] = (int32_t)nnz; info[1] = ncenters; }
    free(est); free(g); free(edges); free(acc); free(nz);
    return count;
}

/* cv2.rotate(img, code): 0 ROTATE_90_CLOCKWISE, 1 ROTATE_180, 2 ROTATE_90_COUNTERCLOCKWISE            */
/* (game_session.py:103-104,125-126 rotate the warped board by 180 degrees when orientation_flipped).  */
/* A pure permutation: clockwise dst(x, H-1-y) = src(y, x); 180 dst(H-1-y, W-1-x); ccw dst(W-1-x, y).  */
ORC_API int orc_rotate(const u8 *src, int H, int W, int C, int code, u8 *dst)
{
    if (code < 0 || code > 2) return -1;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            long o;
            if (code == 1) o = (long)(H - 1 - y) * W + (W - 1 - x);
            else if (code == 0) o = (long)x * H + (H - 1 - y);
            else o = (long)(W - 1 - x) * H + y;
            for (int c = 0; c < C; c++) dst[o * C + c] = src[((long)y * W + x) * C + c];
        }
    return 0;
}

/* cv2.dilate(src, np.ones((kh, kw)), iterations) (board_detection.py:13-14): a max filter; outside the image    */
/* counts as -inf (morphologyDefaultBorderValue), so `iterations` passes of a kw x kh rectangle equal one pass of */
/* ((kw-1)*iterations+1) x ((kh-1)*iterations+1).                                                              */
ORC_API int orc_dilate(const u8 *src, int H, int W, int kw, int kh, int iterations, u8 *dst)
{
    if (kw < 1 || kh < 1 || !(kw & 1) || !(kh & 1) || iterations < 1) return -1;
    const int rx = (kw - 1) * iterations / 2, ry = (kh - 1) * iterations / 2;
    u8 *tmp = (u8 *)malloc((long)H * W);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int m = 0;
            for (int xx = x - rx < 0 ? 0 : x - rx; xx <= x + rx && xx < W; xx++) if (src[(long)y * W + xx] > m) m = src[(long)y * W + xx];
            tmp[(long)y * W + x] = (u8)m;
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            int m = 0;
            for (int yy = y - ry < 0 ? 0 : y - ry; yy <= y + ry && yy < H; yy++) if (tmp[(long)yy * W + x] > m) m = tmp[(long)yy * W + x];
            dst[(long)y * W + x] = (u8)m;
        }
    free(tmp);
    return 0;
}
/* The image part of board_detection.find_chessboard_corners (board_detection.py:9-14):                       */
/* gray -> GaussianBlur(7x7, sigma 1) -> Canny(30, 100) -> dilate(5x5, 3 iterations)                          */
ORC_API void orc_contour_mask(const u8 *bgr, int H, int W, u8 *mask)
{
    const long n = (long)H * W;
    u8 *g = (u8 *)malloc(n), *b = (u8 *)malloc(n), *e = (u8 *)malloc(n);
    orc_gray(bgr, H, W, 3L * W, g);
    orc_gaussian_sigma(g, H, W, W, 7, 1.0, b);
    orc_canny(b, H, W, 30, 100, e);
    orc_dilate(e, H, W, 5, 5, 3, mask);
    free(g); free(b); free(e);
}

/* find_internal_lines (grid_extractor.py:83-110): border 0, seven window arg-max positions, border `length` */
static void internal_lines(const uint64_t *proj, int length, int32_t *lines)
{
    const double step = length / 8.0;
    lines[0] = 0;
    for (int i = 1; i < 8; i++) {
        const int center = (int)(i * step), radius = (int)(step * 0.3);
        int start = center - radius; if (start < 0) start = 0;
        int end = center + radius; if (end > length) end = length;
        if (end > start) {
            int best = start;
            for (int k = start + 1; k < end; k++) if (proj[k] > proj[best]) best = k;   /* np.argmax: first maximum */
            lines[i] = best;
        } else lines[i] = center;
    }
    lines[8] = length;
}
/* SmartGridExtractor.refine_grid: gray -> Canny(50,150) -> edge projections -> grid lines */
ORC_API void orc_refine_grid(const u8 *bgr, int H, int W, int32_t *grid_x9, int32_t *grid_y9, u8 *edges_out)
{
    const long n = (long)H * W;
    u8 *g = (u8 *)malloc(n), *e = edges_out ? edges_out : (u8 *)malloc(n);
    orc_gray(bgr, H, W, 3L * W, g);
    orc_canny(g, H, W, 50, 150, e);
    uint64_t *rows = (uint64_t *)calloc(H, 8), *cols = (uint64_t *)calloc(W, 8);
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) { rows[y] += e[(long)y * W + x]; cols[x] += e[(long)y * W + x]; }
    internal_lines(cols, W, grid_x9);
    internal_lines(rows, H, grid_y9);
    free(g); if (!edges_out) free(e); free(rows); free(cols);
}

/* Camera ingest (play_lichess.py:16-18,45 -- cv2.VideoCapture delivers BGR by running these   */
/* conversions on the camera's native packed / semi-planar YUV).  cv2.cvtColor(COLOR_YUV2BGR_YUY2 / _NV12),          */
/* imgproc/src/color_yuv.simd.hpp (ITU-R BT.601, 20 fractional bits):                                                */
/*   yy = max(0, y - 16) * 1220542;  r = (yy + 2^19 + 1673527 (v-128)) >> 20;                                        */
/*   g = (yy + 2^19 - 852492 (v-128) - 409993 (u-128)) >> 20;  b = (yy + 2^19 + 2116026 (u-128)) >> 20;  saturate.   */
/* fmt 1: YUY2, (H, W, 2) bytes Y0 U Y1 V per pixel pair.  fmt 2: NV12, H rows of Y then H/2 rows of interleaved U V. */
static void yuv_px(int y, int u, int v, u8 *bgr)
{
    const int yy = (y - 16 > 0 ? y - 16 : 0) * 1220542, uu = u - 128, vv = v - 128;
    const int r = (yy + (1 << 19) + 1673527 * vv) >> 20;
    const int g = (yy + (1 << 19) - 852492 * vv - 409993 * uu) >> 20;
    const int b = (yy + (1 << 19) + 2116026 * uu) >> 20;
    bgr[0] = (u8)(b < 0 ? 0 : b > 255 ? 255 : b); bgr[1] = (u8)(g < 0 ? 0 : g > 255 ? 255 : g);
    bgr[2] = (u8)(r < 0 ? 0 : r > 255 ? 255 : r);
}
ORC_API int orc_yuv_to_bgr(const u8 *src, int fmt, int H, int W, u8 *bgr)
{
    if (W < 2 || (W & 1) || H < 1 || (fmt == 2 && (H & 1)) || (fmt != 1 && fmt != 2)) return -1;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x += 2) {
            int y0, y1, u, v;
            if (fmt == 1) {
                const u8 *p = src + ((long)y * W + x) * 2;
                y0 = p[0]; u = p[1]; y1 = p[2]; v = p[3];
            } else {
                const u8 *uv = src + (long)H * W + (long)(y / 2) * W + x;
                y0 = src[(long)y * W + x]; y1 = src[(long)y * W + x + 1]; u = uv[0]; v = uv[1];
            }
            yuv_px(y0, u, v, bgr + ((long)y * W + x) * 3);
            yuv_px(y1, u, v, bgr + ((long)y * W + x + 1) * 3);
        }
    return 0;
}
