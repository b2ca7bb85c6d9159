/*
 * prl_oracle.c -- plain-C CPU restatement of PRLib's local-statistics + Otsu binarization path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may link or call it.  libprlib_cuda never does.
 *
 * It restates, from first principles, what the reference computes with
 * OpenCV calls in
 *   src/binarizations/binarizeSauvola.cpp:57-134      (Sauvola)
 *   src/binarizations/binarizeNiblack.cpp:57-127      (Niblack)
 *   src/binarizations/binarizeWolfJolion.cpp:58-147   (Wolf-Jolion)
 *   src/binarizations/binarizeNICK.cpp:58-143         (NICK)
 *   src/binarizations/binarizeFeng.cpp:59-163         (Feng)
 *   src/binarizations/binarizeLocalOtsu.cpp:138-162   (per-rectangle Otsu loop)
 * and OpenCV's getThreshVal_Otsu_8u (third-party, un-vendored; published algorithm restated in
 * SURVEY.md Appendix B.9) behind cv::threshold(THRESH_OTSU) (src/deskew/deskew.cpp:224, ...).
 *
 * Parity pinning: the reference holds no golden vectors ("parity unpinned" by reference
 * fixtures).  This file is pinned bit-for-bit against oracle/prl_oracle.py, which executes the
 * real OpenCV primitives through cv2 4.13.0 (tests/test_oracle.py), and against the SHA-1
 * digests in tests/golden/.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (no compiler FMA contraction: the reference's
 * direct-tap filter2D path rounds every product).
 *
 * Where OpenCV itself fuses: on every x86-64 build with the AVX2/FMA3 dispatch (the cv2 wheel here, any distro
 * OpenCV >= 3.x on a CPU from 2013 on) Mat::convertTo(alpha, beta), cv::scaleAdd and cv::addWeighted evaluate their
 * multiply-add with ONE rounding (v_fma / v_muladd in convert_scale.simd.hpp, matmul.simd.hpp, arithm.simd.hpp);
 * measured on the wheel by tests/test_ref.py::test_wheel_multiply_add_is_fused, and what oracle/_ref (the reference's
 * own C++ over that wheel) therefore computes.  The explicit fma() calls below restate exactly those three.
 * (addWeighted's scalar tail -- the last cols % 8 columns -- is contracted by the compiler in a different order,
 * fma(b, beta, a*alpha) + gamma; that is a <= 1 ulp difference in T on those columns and is NOT restated.)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>

enum { PRL_SAUVOLA = 0, PRL_NIBLACK = 1, PRL_WOLFJOLION = 2, PRL_NICK = 3, PRL_FENG = 4 };

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- synthpage-v2 ---------------------------------------------- */
static inline uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
static inline uint32_t key32(uint32_t s, uint32_t p, uint32_t a, uint32_t b)
{
    return s * 0x9E3779B1u + p * 0x85EBCA77u + a * 0xC2B2AE3Du + b * 0x27D4EB2Fu;
}
void oracle_synth_page(uint8_t *dst, size_t step, int rows, int cols, uint32_t seed, uint32_t page)
{
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            uint32_t noise = mix32(key32(seed, page, (uint32_t)y, (uint32_t)x)) & 15u;
            uint32_t illum = (40u * (uint32_t)x) / (uint32_t)cols + (24u * (uint32_t)y) / (uint32_t)rows;
            uint32_t stain = (mix32(key32(seed ^ 0x5BD1E995u, page, (uint32_t)y / 64u, (uint32_t)x / 64u)) >> 4) & 31u;
            uint32_t bg = 235u - illum - stain - noise;
            int band = ((y % 48) < 26) && x >= 150 && x < cols - 150 && y >= 200 && y < rows - 200;
            uint32_t b = mix32(key32(seed ^ 0xA5A5A5A5u, page, (uint32_t)y / 3u, (uint32_t)x / 3u));
            int ink = band && ((b & 0xFFu) < 56u);
            uint32_t inkv = 24u + ((b >> 8) & 127u) + (noise >> 1);
            dst[(size_t)y * step + x] = (uint8_t)(ink ? inkv : bg);
        }
}

/* ---- copyMakeBorder(REPLICATE) + integral(CV_64F) + Rect(1,1,..) crop -------------------
 * binarizeSauvola.cpp:65-77.  Inclusive prefix sums over the replicate-padded image, exact
 * 64-bit integers.  sum / sqsum: (rows+2*pad) x (cols+2*pad), row-major, contiguous. */
void oracle_integral_u8(const uint8_t *src, int rows, int cols, size_t step, int pad,
                        int64_t *sum, int64_t *sqsum)
{
    const int Hp = rows + 2 * pad, Wp = cols + 2 * pad;
    for (int Y = 0; Y < Hp; ++Y) {
        const uint8_t *row = src + (size_t)clampi(Y - pad, 0, rows - 1) * step;
        int64_t rs = 0, rq = 0;
        for (int X = 0; X < Wp; ++X) {
            int64_t p = row[clampi(X - pad, 0, cols - 1)];
            rs += p; rq += p * p;
            sum[(size_t)Y * Wp + X] = rs + (Y ? sum[(size_t)(Y - 1) * Wp + X] : 0);
            sqsum[(size_t)Y * Wp + X] = rq + (Y ? sqsum[(size_t)(Y - 1) * Wp + X] : 0);
        }
    }
}

/* geometry: SURVEY.md Appendix A.2 */
int oracle_output_shape(int method, int rows, int cols, int window, int *out_rows, int *out_cols)
{
    int w = window < rows ? window : rows; if (cols < w) w = cols;
    int h = w / 2;
    if (method == PRL_SAUVOLA || method == PRL_NIBLACK) { *out_cols = cols + 2 * h - w; *out_rows = rows + 2 * h - w; }
    else { *out_cols = cols - w; *out_rows = rows - w; }
    return w;
}

/* Mat::convertTo(CV_8UC1): saturate_cast<uchar>(cvRound(T)); NaN / +-inf -> 0 (Appendix A.6) */
static inline uint8_t to_u8(double T)
{
    /* cvRound = cvtsd2si: NaN and anything outside int32 give INT_MIN, which saturates to 0 */
    if (!(T == T) || T >= 2147483647.5 || T <= 0.0) return 0;
    if (T >= 255.5) return 255;   /* rint would be >= 256 (255.5 rounds to even 256) */
    return (uint8_t)clampi((int)nearbyint(T), 0, 255);
}

/* One FP64 threshold from the four S taps / four Q taps, the reference's operation order. */
typedef struct { double kw, p0, p1, p2, p3, imin, coeff; int method; } thr_ctx;

static inline void mean_dev(const thr_ctx *c, const int64_t *S, const int64_t *Q, size_t ia, size_t ib,
                            size_t ic, size_t id, double *m_out, double *s_out)
{
    /* filter2D direct path: taps in row-major kernel order, every product rounded
     * (binarizeSauvola.cpp:83-90, :106-107) */
    double m = c->kw * (double)S[ia] + (-c->kw) * (double)S[ib] + (-c->kw) * (double)S[ic] + c->kw * (double)S[id];
    double q = c->kw * (double)Q[ia] + (-c->kw) * (double)Q[ib] + (-c->kw) * (double)Q[ic] + c->kw * (double)Q[id];
    double msq = m * m;              /* :93  */
    *m_out = m;
    *s_out = sqrt(q - msq);          /* :109-110 */
}

static inline double threshold_value(const thr_ctx *c, double m, double s)
{
    switch (c->method) {
    case PRL_SAUVOLA: {               /* binarizeSauvola.cpp:115-118 */
        double t = fma(s, c->p1, c->p2); /* Mat::convertTo(alpha = k*(1/128), beta = 1-k): ONE rounding (header note) */
        return m * t;
    }
    case PRL_NIBLACK:                 /* binarizeNiblack.cpp:108 */
        return fma(s, c->p0, m);      /* MatExpr m + k*s -> cv::scaleAdd(s, k, m): ONE rounding (header note) */
    case PRL_WOLFJOLION: {            /* binarizeWolfJolion.cpp:128-130 */
        double d = fma(s, c->coeff, -c->p0);         /* convertTo(coeff, -k): ONE rounding */
        d = d * (m - c->imin);
        return m + d;
    }
    case PRL_NICK: {                  /* binarizeNICK.cpp:121-126 */
        double C = sqrt(m * m + s * s);
        return fma(m, 1.0, fma(C, c->p0, 0.0));      /* cv::addWeighted; == m + round(C*k) because alpha is 1 */
    }
    default: {                        /* binarizeFeng.cpp:118-142 */
        double t1 = s / s;            /* Rs aliases s (:118) -> 1, or NaN when s is 0/NaN */
        double t2 = (t1 == t1) ? 1.0 : t1;          /* cv::pow(1, gamma) == 1; NaN stays NaN */
        double alpha3 = c->p2 * t2;   /* k2 * tmpAlpha2 */
        double c2 = t2 * t1;
        double c3 = fma(alpha3, c->imin, fma(c2, -c->imin, 0.0));   /* cv::addWeighted: fma(a, alpha, fma(b, beta, gamma)) */
        double T = c2 + (1.0 - c->p0);
        T = T * m;
        return T + c3;
    }
    }
}

/*
 * Threshold map T8 and mask for one gray page.  params: Sauvola/Niblack/WJ/NICK {k};
 * Feng {alpha1,k1,k2,gamma}.  t8 and/or mask may be NULL.  aux (may be NULL) receives
 * {imin, smax}.  Returns 0, or -1 when the output rect is empty (OpenCV would throw).
 */
int oracle_binarize_local(const uint8_t *src, int rows, int cols, size_t step, int method, int window,
                          const double *params, uint8_t *t8, uint8_t *mask, double *aux)
{
    int Hout, Wout;
    const int w = oracle_output_shape(method, rows, cols, window, &Hout, &Wout);
    const int h = w / 2, d = w - 1;
    if (Hout <= 0 || Wout <= 0) return -1;
    const int Hp = rows + 2 * h, Wp = cols + 2 * h;
    int64_t *S = (int64_t *)malloc(sizeof(int64_t) * (size_t)Hp * Wp);
    int64_t *Q = (int64_t *)malloc(sizeof(int64_t) * (size_t)Hp * Wp);
    if (!S || !Q) { free(S); free(Q); return -2; }
    oracle_integral_u8(src, rows, cols, step, h, S, Q);

    thr_ctx c;
    memset(&c, 0, sizeof c);
    c.method = method;
    c.kw = 1.0 / (double)(w * w);
    c.p0 = params[0];
    if (method == PRL_SAUVOLA) { c.p1 = params[0] * (1.0 / 128.0); c.p2 = 1.0 - params[0]; }
    if (method == PRL_FENG) { c.p1 = params[1]; c.p2 = params[2]; c.p3 = params[3]; }

    double imin = 255.0;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x)
            if (src[(size_t)y * step + x] < imin) imin = src[(size_t)y * step + x];
    c.imin = imin;

    double smax = -INFINITY;    /* cv::minMaxLoc on an all-NaN map reports -inf */
    if (method == PRL_WOLFJOLION) {    /* minMaxLoc(localDevianceValues) :118-119 (NaN never wins) */
        for (int y = 0; y < Hout; ++y)
            for (int x = 0; x < Wout; ++x) {
                double m, s;
                mean_dev(&c, S, Q, (size_t)y * Wp + x, (size_t)y * Wp + x + d,
                         (size_t)(y + d) * Wp + x, (size_t)(y + d) * Wp + x + d, &m, &s);
                if (s > smax) smax = s;
            }
        c.coeff = c.p0 / smax;
    }
    if (aux) { aux[0] = imin; aux[1] = smax; }

    for (int y = 0; y < Hout; ++y)
        for (int x = 0; x < Wout; ++x) {
            double m, s;
            mean_dev(&c, S, Q, (size_t)y * Wp + x, (size_t)y * Wp + x + d,
                     (size_t)(y + d) * Wp + x, (size_t)(y + d) * Wp + x + d, &m, &s);
            uint8_t T8 = to_u8(threshold_value(&c, m, s));
            if (t8) t8[(size_t)y * Wout + x] = T8;
            if (mask) mask[(size_t)y * Wout + x] = src[(size_t)y * step + x] > T8 ? 255 : 0;
        }
    free(S); free(Q);
    return 0;
}

/* ---- morphology tail (binarizeSauvola.cpp:125-134): n x 3x3 == one (2n+1)^2, border ignored */
static void morph_pass(const uint8_t *src, uint8_t *dst, int rows, int cols, int n, int is_dilate)
{
    uint8_t *tmp = (uint8_t *)malloc((size_t)rows * cols);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            int lo = x - n < 0 ? 0 : x - n, hi = x + n >= cols ? cols - 1 : x + n;
            uint8_t v = src[(size_t)y * cols + lo];
            for (int j = lo + 1; j <= hi; ++j) {
                uint8_t u = src[(size_t)y * cols + j];
                v = is_dilate ? (u > v ? u : v) : (u < v ? u : v);
            }
            tmp[(size_t)y * cols + x] = v;
        }
    for (int y = 0; y < rows; ++y) {
        int lo = y - n < 0 ? 0 : y - n, hi = y + n >= rows ? rows - 1 : y + n;
        for (int x = 0; x < cols; ++x) {
            uint8_t v = tmp[(size_t)lo * cols + x];
            for (int j = lo + 1; j <= hi; ++j) {
                uint8_t u = tmp[(size_t)j * cols + x];
                v = is_dilate ? (u > v ? u : v) : (u < v ? u : v);
            }
            dst[(size_t)y * cols + x] = v;
        }
    }
    free(tmp);
}
void oracle_morph(uint8_t *mask, int rows, int cols, int iters)
{
    if (iters == 0) return;
    uint8_t *t = (uint8_t *)malloc((size_t)rows * cols);
    int n = iters > 0 ? iters : -iters;
    morph_pass(mask, t, rows, cols, n, iters > 0);    /* n>0: dilate then erode (closing) */
    morph_pass(t, mask, rows, cols, n, !(iters > 0)); /* n<0: erode then dilate (opening) */
    free(t);
}

/* ---- Otsu: literal getThreshVal_Otsu_8u recurrence ------------- */
int oracle_otsu_from_hist(const int32_t *hist)
{
    int64_t n = 0;
    for (int i = 0; i < 256; ++i) n += hist[i];
    if (n == 0) return 0;
    double mu = 0, scale = 1.0 / (double)n;
    for (int i = 0; i < 256; ++i) mu += i * (double)hist[i];
    mu *= scale;
    double mu1 = 0, q1 = 0, max_sigma = 0;
    int max_val = 0;
    for (int i = 0; i < 256; ++i) {
        double p_i, q2, mu2, sigma;
        p_i = hist[i] * scale;
        mu1 *= q1;
        q1 += p_i;
        q2 = 1. - q1;
        if (fmin(q1, q2) < FLT_EPSILON || fmax(q1, q2) > 1. - FLT_EPSILON) continue;
        mu1 = (mu1 + i * p_i) / q1;
        mu2 = (mu - q1 * mu1) / q2;
        sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
        if (sigma > max_sigma) { max_sigma = sigma; max_val = i; }
    }
    return max_val;
}

int oracle_otsu_threshold(const uint8_t *src, int rows, int cols, size_t step)
{
    int32_t hist[256] = {0};
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) hist[src[(size_t)y * step + x]]++;
    return oracle_otsu_from_hist(hist);
}

/* cv::threshold(src,dst,128,maxval,THRESH_BINARY|THRESH_OTSU); returns the threshold */
int oracle_otsu_global(const uint8_t *src, int rows, int cols, size_t step, double maxval,
                       uint8_t *dst, size_t dst_step)
{
    int thr = oracle_otsu_threshold(src, rows, cols, step);
    int mv = (int)nearbyint(maxval); mv = clampi(mv, 0, 255);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x)
            dst[(size_t)y * dst_step + x] = src[(size_t)y * step + x] > thr ? (uint8_t)mv : 0;
    return thr;
}

/* binarizeLocalOtsu.cpp:138-162.  xywh: n_rects x 4 int32.  dst is fully written. */
void oracle_otsu_rects(const uint8_t *src, int rows, int cols, size_t step, const int32_t *xywh,
                       int n_rects, double maxval, uint8_t *dst, size_t dst_step)
{
    int mv = clampi((int)nearbyint(maxval), 0, 255);
    for (int y = 0; y < rows; ++y) memset(dst + (size_t)y * dst_step, 255, (size_t)cols);
    for (int r = 0; r < n_rects; ++r) {
        int x0 = xywh[4 * r], y0 = xywh[4 * r + 1], w = xywh[4 * r + 2], h = xywh[4 * r + 3];
        int thr = oracle_otsu_threshold(src + (size_t)y0 * step + x0, h, w, step);
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                uint8_t t = src[(size_t)(y0 + y) * step + x0 + x] > thr ? (uint8_t)mv : 0;
                if ((uint8_t)(t ^ 255)) dst[(size_t)(y0 + y) * dst_step + x0 + x] = 0;  /* :159 */
            }
    }
}
