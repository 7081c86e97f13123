// Host-side tables and small host maths of libcvb200 (no CUDA needed here).
//
// The 8-bit LAB conversions of the reference (frame_enhancer.py:108,120 ->
// cv2.cvtColor) are integer algorithms driven by five small tables (OpenCV
// imgproc/src/color_lab.cpp: sRGBGammaTab_b, LabCbrtTab_b, LabToYF_b,
// sRGBInvGammaTab_b).  They are regenerated here from their closed forms in
// f64; SURVEY.md 8a (a2, a4) records that this reproduces cv2 on all 2^24
// inputs in both directions once two cbrt entries are nudged.
#include "cvb_internal.h"
#include <cmath>
#include <cfloat>
#include <cstring>
#include <mutex>

namespace {

double gamma_fwd(double x) { return x <= 0.04045 ? x * (1.0 / 12.92) : std::pow((x + 0.055) * (1.0 / 1.055), 2.4); }
double gamma_inv(double x) { return x <= 0.0031308 ? x * 12.92 : 1.055 * std::pow(x, 1.0 / 2.4) - 0.055; }

CvbTables g_tab;
std::once_flag g_once;

void build_tables()
{
    CvbTables &t = g_tab;
    for (int i = 0; i < 256; ++i) t.gamma[i] = (uint16_t)std::nearbyint(2040.0 * gamma_fwd(i / 255.0));
    const double thresh = 216.0 / 24389.0, slope = 841.0 / 108.0, ofs = 16.0 / 116.0;
    for (int i = 0; i < 2048; ++i) {
        double x = i / 2040.0;
        double f = x < thresh ? x * slope + ofs : std::cbrt(x);
        t.cbrt[i] = (uint16_t)std::nearbyint(32768.0 * f);
    }
    // OpenCV fills LabCbrtTab_b with a softfloat cbrt; these two entries round
    // the other way there (found by exhaustive comparison, SURVEY.md probe 8-10)
    t.cbrt[49] -= 1;
    t.cbrt[628] += 1;
    const double base = 16384.0;
    for (int L = 0; L < 256; ++L) {
        double y, ify;
        if (L <= 20) {
            y = std::nearbyint(L * base * 180.0 / (17.0 * 29.0 * 29.0 * 29.0));
            ify = std::nearbyint(base * (16.0 / 116.0 + L * 5.0 / 1479.0));
        } else {
            double fy = L * 100.0 * base / (255.0 * 116.0) + 16.0 * base / 116.0;
            ify = std::nearbyint(fy);
            y = std::nearbyint(fy * fy * fy / (base * base));
        }
        // kernel layout: y in the low half-word, ify in the high one (both < 2^15)
        t.lab2yf[L] = (int32_t)y | ((int32_t)ify << 16);
        t.lab2yf[256 + L] = 0;
    }
    for (int i = 0; i < 4096; ++i) {
        int v = (int)std::nearbyint(255.0 * gamma_inv(i / 4096.0));
        t.invgamma[i] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
    }
    for (int i = 0; i < 2048; ++i) {
        int L = (296 * (int)t.cbrt[i] - 1336934 + 16384) >> 15;
        t.ltab[i] = (uint8_t)(L < 0 ? 0 : L > 255 ? 255 : L);
    }
}

}  // namespace

const CvbTables &cvb_host_tables()
{
    std::call_once(g_once, build_tables);
    return g_tab;
}

// cv2.bilateralFilter tables (bilateral_filter.dispatch.cpp): colour weights
// indexed by |db|+|dg|+|dr|, spatial weights on the circular support r <= 4.
void cvb_host_bilateral_tables(double sigma_color, double sigma_space, float *color768, float *space81)
{
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    const double gc = -0.5 / (sigma_color * sigma_color), gs = -0.5 / (sigma_space * sigma_space);
    if (color768)
        for (int i = 0; i < 768; ++i) color768[i] = (float)std::exp((double)i * i * gc);
    if (space81)
        for (int dy = -4; dy <= 4; ++dy)
            for (int dx = -4; dx <= 4; ++dx) {
                double r = std::sqrt((double)dy * dy + (double)dx * dx);
                space81[(dy + 4) * 9 + dx + 4] = r > 4 ? 0.0f : (float)std::exp(r * r * gs);
            }
}

// cv2.getGaussianKernel(k, 0) quantised to Q8 with error diffusion, as the
// u8 fixed-point GaussianBlur does (smooth.dispatch.cpp).
int cvb_host_gaussian_q8(int k, int *q) { return cvb_host_gaussian_q8_sigma(k, 0.0, q); }
// sigma > 0: cv2.GaussianBlur(src, (k, k), sigma) (board_detection.py:10 uses (7, 7), 1)
int cvb_host_gaussian_q8_sigma(int k, double sigma_in, int *q)
{
    if (k < 1 || k > 31 || !(k & 1)) return CVB_ERR_INVALID;
    double kern[31];
    static const double fixed[4][7] = {{1.0},
                                       {0.25, 0.5, 0.25},
                                       {0.0625, 0.25, 0.375, 0.25, 0.0625},
                                       {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125}};
    if (k <= 7 && sigma_in <= 0) {
        for (int i = 0; i < k; ++i) kern[i] = fixed[k / 2][i];
    } else {
        double sigma = sigma_in > 0 ? sigma_in : ((k - 1) * 0.5 - 1) * 0.3 + 0.8, s2 = -0.5 / (sigma * sigma), sum = 0;
        for (int i = 0; i < k; ++i) {
            double x = i - (k - 1) * 0.5;
            kern[i] = std::exp(s2 * x * x);
            sum += kern[i];
        }
        sum = 1.0 / sum;
        for (int i = 0; i < k; ++i) kern[i] *= sum;
    }
    double carry = 0;
    long acc = 0;
    for (int i = 0; i < k / 2; ++i) {
        double want = kern[i] * 256.0 + carry;
        long v = (long)std::nearbyint(want);
        carry = want - (double)v;
        q[i] = q[k - 1 - i] = (int)v;
        acc += v;
    }
    q[k / 2] = (int)(256 - 2 * acc);
    return CVB_OK;
}

// cv2.getPerspectiveTransform: 8x8 system, LU with partial pivoting in f64;
// the Point2f products are formed in f32 first (imgwarp.cpp).
int cvb_host_get_perspective(const float *src, const float *dst, double *M)
{
    const int m = 8;
    double A[64] = {0}, b[8];
    for (int i = 0; i < 4; ++i) {
        float sx = src[2 * i], sy = src[2 * i + 1], dx = dst[2 * i], dy = dst[2 * i + 1];
        double *r0 = A + i * m, *r1 = A + (i + 4) * m;
        r0[0] = r1[3] = sx; r0[1] = r1[4] = sy; r0[2] = r1[5] = 1.0;
        r0[6] = (float)(-sx * dx); r0[7] = (float)(-sy * dx);
        r1[6] = (float)(-sx * dy); r1[7] = (float)(-sy * dy);
        b[i] = dx; b[i + 4] = dy;
    }
    for (int i = 0; i < m; ++i) {
        int piv = i;
        for (int j = i + 1; j < m; ++j)
            if (std::fabs(A[j * m + i]) > std::fabs(A[piv * m + i])) piv = j;
        if (std::fabs(A[piv * m + i]) < DBL_EPSILON * 100) return CVB_ERR_INVALID;
        if (piv != i) {
            for (int j = i; j < m; ++j) std::swap(A[i * m + j], A[piv * m + j]);
            std::swap(b[i], b[piv]);
        }
        double d = -1 / A[i * m + i];
        for (int j = i + 1; j < m; ++j) {
            double alpha = A[j * m + i] * d;
            for (int c = i + 1; c < m; ++c) A[j * m + c] += alpha * A[i * m + c];
            b[j] += alpha * b[i];
        }
    }
    for (int i = m - 1; i >= 0; --i) {
        double s = b[i];
        for (int c = i + 1; c < m; ++c) s -= A[i * m + c] * b[c];
        b[i] = s / A[i * m + i];
    }
    for (int i = 0; i < 8; ++i) M[i] = b[i];
    M[8] = 1.0;
    return CVB_OK;
}

// cv::invert for 3x3 (adjugate / determinant), as warpPerspective does.
int cvb_host_invert3(const double *a, double *t)
{
    double det = a[0] * (a[4] * a[8] - a[5] * a[7]) - a[1] * (a[3] * a[8] - a[5] * a[6]) +
                 a[2] * (a[3] * a[7] - a[4] * a[6]);
    if (det == 0.0) return CVB_ERR_INVALID;
    double d = 1.0 / det;
    t[0] = (a[4] * a[8] - a[5] * a[7]) * d; t[1] = (a[2] * a[7] - a[1] * a[8]) * d; t[2] = (a[1] * a[5] - a[2] * a[4]) * d;
    t[3] = (a[5] * a[6] - a[3] * a[8]) * d; t[4] = (a[0] * a[8] - a[2] * a[6]) * d; t[5] = (a[2] * a[3] - a[0] * a[5]) * d;
    t[6] = (a[3] * a[7] - a[4] * a[6]) * d; t[7] = (a[1] * a[6] - a[0] * a[7]) * d; t[8] = (a[0] * a[4] - a[1] * a[3]) * d;
    return CVB_OK;
}

// Region masks of one square shape (they depend on (h, w) only):
//   bit0  centre disc   (x-cx)^2+(y-cy)^2 <= (min(h,w)//4)^2   piece_detector.py:184-190
//   bit1  four corner blocks of min(h,w)//4                     piece_detector.py:192-198
//   bit2+k ring k: |dist - min(h,w)*{.15,.25,.35,.45}| <= 5     piece_detector.py:148-163
void cvb_host_square_masks(int h, int w, uint8_t *mask)
{
    const int cy = h / 2, cx = w / 2, md = h < w ? h : w, rad = md / 4, cs = md / 4;
    const double ratios[4] = {0.15, 0.25, 0.35, 0.45};
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int bits = 0;
            const int d2 = (x - cx) * (x - cx) + (y - cy) * (y - cy);
            if (d2 <= rad * rad) bits |= 1;
            // numpy [:cs] / [-cs:]; with cs == 0 the slice [-0:] is the whole axis
            const bool top = y < cs, left = x < cs;
            const bool bottom = cs == 0 || y >= h - cs, right = cs == 0 || x >= w - cs;
            if ((top || bottom) && (left || right)) bits |= 2;
            const double dist = std::sqrt((double)d2);
            for (int k = 0; k < 4; ++k) {
                double r = md * ratios[k];
                if (dist >= r - 5 && dist <= r + 5) bits |= 4 << k;
            }
            mask[(size_t)y * w + x] = (uint8_t)bits;
        }
}
