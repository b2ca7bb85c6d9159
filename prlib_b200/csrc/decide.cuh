// decide.cuh -- the per-pixel decision shared by kernel 2 (threshold.cu) and the fused small-window kernel
// (fused.cu): exact integer window sums -> FP32 estimate of T -> decided unless within a proven margin of the
// rounding boundary, in which case the reference's FP64 formula is evaluated literally from the int64 taps.
#pragma once
#include "common.cuh"
#include <cmath>

namespace {

struct FastArgs {
    float kwf, inv_w2f;          // 1/w^2 as float
    float c0, c1, c2;            // method constants in FP32
    float mu0, mu1;              // decision margin: mu = mu0 + mu1*|coeff|  (mu1 only for Wolf-Jolion)
    float n_floor;               // fast path only when N >= n_floor  (s* >= s_floor)
    float q_floor;               // q_win >= q_floor implies N >= n_floor (used by the variance-free pre-test)
    // second tier (fused kernel): FP64 estimate from the exact integers, margin mu2 = t2_a + t2_b / s, valid for v >= t2_vmin
    double t2_a, t2_b, t2_vmin, kw_d, c0_d, c1_d, c2_d;
    unsigned int w2;             // w*w
    int rows_per_cta;
    // fused kernel: variance-free pre-test with bounds linear in m:  T in [pa_lo*m + pb_lo, pa_hi*m + pb_hi] for s in [0,128]
    float pa_hi, pa_lo, pb_hi, pb_lo;
    unsigned int qfloor_u;       // integer form of q_floor
    double dv_ref;               // bound on |v_ref - v| of the reference's FP64 variance
    int n32;                     // w^2 * (w-1)^2 * 255^2 < 2^32: N fits 32 bits
    int dbg_skip_exact;          // DIAGNOSTIC ONLY (wrong masks): undecided pixels are not evaluated -- measures what the exact path costs
};

__device__ __forceinline__ int to_u8(double T)
{
    // cvRound == cvtsd2si: NaN and anything outside int32 become INT_MIN -> saturates to 0
    if (!(T > 0.0) || T >= 2147483647.5) return 0;
    if (T >= 255.5) return 255;
    return __double2int_rn(T);
}

// exact int64 -> double for 0 <= v < 2^52 without the (quarter-rate) I2F.F64.S64
__device__ __forceinline__ double i2d(long long v)
{
    return __dadd_rn(__longlong_as_double(v | 0x4330000000000000LL), -4503599627370496.0);
}

__device__ __forceinline__ double tap4(double kw, double nkw, long long a, long long b, long long c, long long d)
{
    double r = __dmul_rn(kw, i2d(a));
    r = __dadd_rn(r, __dmul_rn(nkw, i2d(b)));
    r = __dadd_rn(r, __dmul_rn(nkw, i2d(c)));
    r = __dadd_rn(r, __dmul_rn(kw, i2d(d)));
    return r;
}

// The reference's FP64 threshold formulas (binarizeSauvola.cpp:115-118, binarizeNiblack.cpp:108,
// binarizeWolfJolion.cpp:128-130, binarizeNICK.cpp:121-126, binarizeFeng.cpp:118-142), operation order kept.
// Roundings follow what OpenCV executes on an FMA-capable host (pinned by the compiled reference of the test tree, the reference's own C++ over the
// cv2 wheel): Mat::convertTo(alpha, beta) and cv::scaleAdd fuse their multiply-add (one rounding), cv::addWeighted is
// fma(a, alpha, fma(b, beta, gamma)); filter2D's taps, Mat::mul, add and subtract round every operation.
template <int METHOD>
__device__ __forceinline__ double thr_value_p(double m, double s, double p0, double p1, double p2, double imin, double coeff)
{
    if (METHOD == PRL_SAUVOLA) {
        return __dmul_rn(m, __fma_rn(s, p1, p2));                      // convertTo(k/128, 1-k), then mul
    } else if (METHOD == PRL_NIBLACK) {
        return __fma_rn(s, p0, m);                                     // scaleAdd(s, k, m)
    } else if (METHOD == PRL_WOLFJOLION) {
        double dd = __fma_rn(s, coeff, -p0);                           // convertTo(coeff, -k)
        dd = __dmul_rn(dd, __dadd_rn(m, -imin));
        return __dadd_rn(m, dd);
    } else if (METHOD == PRL_NICK) {
        double C = __dsqrt_rn(__dadd_rn(__dmul_rn(m, m), __dmul_rn(s, s)));
        return __dadd_rn(m, __dmul_rn(C, p0));
    } else {
        if (!(s == s) || s == 0.0) return __longlong_as_double(0x7ff8000000000000LL);
        double c3 = __fma_rn(p2, imin, -imin);                         // addWeighted(alpha3, imin, c2, -imin, 0), c2 == 1
        return __dadd_rn(__dmul_rn(p1, m), c3);
    }
}

// (scalars by value: a reference to the kernel-parameter struct would force a per-thread stack copy)
template <int METHOD>
__device__ __noinline__ int exact_t8_at(const long long* __restrict__ s0, const long long* __restrict__ q0, size_t drow,
                                        int d, double kw, double p0, double p1, double p2, double imin, double coeff)
{
    const double nkw = -kw;
    const double m = tap4(kw, nkw, __ldg(s0), __ldg(s0 + d), __ldg(s0 + drow), __ldg(s0 + drow + d));
    const double q = tap4(kw, nkw, __ldg(q0), __ldg(q0 + d), __ldg(q0 + drow), __ldg(q0 + drow + d));
    const double s = __dsqrt_rn(__dadd_rn(q, -__dmul_rn(m, m)));
    return to_u8(thr_value_p<METHOD>(m, s, p0, p1, p2, imin, coeff));
}

// the reference arithmetic from the eight integral taps
template <int METHOD>
__device__ __forceinline__ int exact_t8_from_taps(long long sa, long long sb, long long sc, long long sd, long long qa,
                                                  long long qb, long long qc, long long qd, double kw, double p0,
                                                  double p1, double p2, double imin, double coeff)
{
    const double nkw = -kw;
    const double m = tap4(kw, nkw, sa, sb, sc, sd);
    const double q = tap4(kw, nkw, qa, qb, qc, qd);
    const double s = __dsqrt_rn(__dadd_rn(q, -__dmul_rn(m, m)));
    return to_u8(thr_value_p<METHOD>(m, s, p0, p1, p2, imin, coeff));
}

// (float)N for the exact variance numerator N = w^2 Q - S^2 without the slow 64-bit I2F (XU pipe, ~20 % of kernel 2's
// stall samples when it was used): 32-bit arithmetic when N provably fits (w <= 15), else high and low words converted
// separately (I2FP) and joined by one FMA -- one more rounding than a direct conversion, which the margin accounts for.
__device__ __forceinline__ float n_to_float(const FastArgs& F, unsigned int sw, unsigned int qw)
{
    if (F.n32) return (float)(F.w2 * qw - sw * sw);
    const unsigned long long N = (unsigned long long)F.w2 * qw - (unsigned long long)sw * sw;
    return fmaf((float)(unsigned int)(N >> 32), 4294967296.0f, (float)(unsigned int)N);
}

// FP32 estimate of T from the exact window sums; returns false when the pixel must take the exact path
// PRE: also try the variance-free bounds first (pays off where the decision arithmetic dominates: the fused
// kernel and Sauvola in kernel 2; measured slower for Niblack / NICK in kernel 2)
template <int METHOD, bool PRE>
__device__ __forceinline__ bool fast_decide(unsigned int sw, unsigned int qw, unsigned int p, const FastArgs& F,
                                            float iminf, float coefff, float mu, int& out)
{
    if (qw == 0u) { out = 0; return true; }            // all-zero window => p == 0 => (0 > T8) is false
    const float m = (float)sw * F.kwf;
    const float pm = (float)p - 0.5f;
    if (PRE && (METHOD == PRL_SAUVOLA || METHOD == PRL_NIBLACK || METHOD == PRL_NICK)) {
        // T is monotone in s and 0 <= s <= 128: with T(m, 0) and T(m, 128) as bounds most pixels (flat background,
        // solid ink) are settled without the variance.  Sound only where the full test would not send the pixel
        // to the exact path for conditioning (s >= s_floor), which q_win >= q_floor guarantees cheaply:
        // N = w^2 Q - S^2 >= w^2 Q - (w-1)^2 Q = (2w-1) Q  (Cauchy-Schwarz over the (w-1)^2 window).
        float t0, t1;
        if (METHOD == PRL_SAUVOLA) { t0 = m * F.c2; t1 = m * fmaf(128.0f, F.c1, F.c2); }
        else if (METHOD == PRL_NIBLACK) { t0 = m; t1 = fmaf(F.c0, 128.0f, m); }
        else { t0 = fmaf(F.c0, m, m); t1 = fmaf(F.c0, sqrtf(fmaf(m, m, 16384.0f)), m); }   // NICK: sqrt(m^2+s^2) in [m, sqrt(m^2+128^2)]
        const float tlo = fmaxf(fminf(t0, t1), 0.0f), thi = fmaxf(fmaxf(t0, t1), 0.0f);
        if ((float)qw >= F.q_floor) {
            if (pm - thi > mu) { out = 255; return true; }
            if (pm - tlo < -mu) { out = 0; return true; }
        }
    }
    const float fn = n_to_float(F, sw, qw);            // N = w^2 Q - S^2, exact and >= 0 before the conversion
    const float s = (fn * rsqrtf(fn)) * F.inv_w2f;     // sqrt via MUFU.RSQ (2 ulp, covered by the margin); fn == 0 gives NaN -> exact path
    float T;
    if (METHOD == PRL_SAUVOLA) T = m * fmaf(s, F.c1, F.c2);
    else if (METHOD == PRL_NIBLACK) T = fmaf(F.c0, s, m);
    else if (METHOD == PRL_WOLFJOLION) T = fmaf(fmaf(s, coefff, -F.c0), m - iminf, m);
    else if (METHOD == PRL_NICK) T = fmaf(F.c0, sqrtf(fmaf(m, m, s * s)), m);
    else T = fmaf(F.c1, m, fmaf(F.c2, iminf, -iminf));
    const float g = pm - fmaxf(T, 0.0f);
    const bool ok = fn >= F.n_floor;
    if (ok && g > mu) { out = 255; return true; }
    if (ok && g < -mu) { out = 0; return true; }
    return false;                                       // near the rounding boundary, ill-conditioned, or NaN
}

// Second tier: FP64 estimate of T from the exact window sums.  Decides the pixel unless it lies within
// mu2 = t2_a + t2_b / s of the rounding boundary (mu2 bounds the reference's own FP64 rounding, which depends
// on the absolute integral values this caller does not have) or the window is too dark for the bound to hold.
// Returns 0 / 255, or -1 when the pixel stays undecided.
template <int METHOD>
__device__ __noinline__ int tier2_decide(unsigned int sw, unsigned int qw, unsigned int p, double kw, double c0, double c1,
                                         double c2, double t2_a, double t2_b, double t2_vmin, unsigned int w2, double imin)
{
    const unsigned long long N = (unsigned long long)w2 * qw - (unsigned long long)sw * sw;
    const double v = (double)N * kw * kw;
    if (!(v >= t2_vmin)) return -1;
    const double m = (double)sw * kw, s = sqrt(v);
    double T;
    if (METHOD == PRL_SAUVOLA) T = m * (s * c1 + c2);
    else if (METHOD == PRL_NIBLACK) T = m + c0 * s;
    else if (METHOD == PRL_NICK) T = m + c0 * sqrt(m * m + s * s);
    else T = c1 * m + (c2 * imin - imin);
    const double g = ((double)p - 0.5) - fmax(T, 0.0);
    const double mu2 = t2_a + t2_b / s;
    if (g > mu2) return 255;
    if (g < -mu2) return 0;
    return -1;
}

// Host-side error analysis for the fast path: returns false when the margin is too large to be useful.
//   reference FP64 error (vs exact real arithmetic), u = 2^-53:
//     dm_ref <= 16 u kw Smax,  dq_ref <= 16 u kw Qmax,  dv_ref <= dq_ref + 2*255*dm_ref + u*2*255^2
//     ds_ref <= dv_ref / s_floor + u*128          (fast path requires s* >= s_floor)
//   FP32 estimate error (exact integer inputs), e = 2^-24:
//     dm_est <= 3 e 255,  ds_est <= 10 e 128  (N -> float in up to two roundings, MUFU.RSQ at 2 ulp, two products)
//   |dT| <= A dm + B ds + 8 e (Tmax + 512), A/B = sup |dT/dm|, |dT/ds| over m in [0,255], s in [0,128]
inline bool fast_margins(int method, const double* params, const prl_geom& g, FastArgs* F)
{
    const double u = 1.1102230246251565e-16, e = 5.9604644775390625e-08;
    const double w2 = (double)g.w * g.w, kw = 1.0 / w2;
    const double area = (double)g.Hp * (double)g.Wp;
    const double Smax = 255.0 * area, Qmax = 65025.0 * area;
    const double s_floor = 0.25;
    const double dm_ref = 16 * u * kw * Smax, dq_ref = 16 * u * kw * Qmax;
    const double dv_ref = dq_ref + 510.0 * dm_ref + u * 2 * 65025.0;
    if (!(dv_ref < 0.25 * s_floor * s_floor)) return false;
    const double ds_ref = dv_ref / s_floor + u * 128;
    const double dm = dm_ref + 3 * e * 255, ds = ds_ref + 10 * e * 128;  // ds: int->float (two-word: 2 roundings), rsqrt (2 ulp), two products
    double Acoef, Bcoef, Tmax, mu1 = 0.0, cd0 = 0.0, cd1 = 0.0, cd2 = 0.0;
    const double k = params[0];
    switch (method) {
    case PRL_SAUVOLA: {
        const double c1 = k * (1.0 / 128.0), c2 = 1.0 - k;
        Acoef = fabs(c2) + 128 * fabs(c1); Bcoef = 255 * fabs(c1); Tmax = 255 * Acoef;
        F->c0 = (float)k; F->c1 = (float)c1; F->c2 = (float)c2; cd0 = k; cd1 = c1; cd2 = c2; break;
    }
    case PRL_NIBLACK:
        Acoef = 1; Bcoef = fabs(k); Tmax = 255 + 128 * fabs(k);
        F->c0 = (float)k; F->c1 = F->c2 = 0; cd0 = k; break;
    case PRL_NICK:
        Acoef = 1 + fabs(k); Bcoef = fabs(k); Tmax = 255 + fabs(k) * 286;
        F->c0 = (float)k; F->c1 = F->c2 = 0; cd0 = k; break;
    case PRL_WOLFJOLION:
        // T = m + (s*coeff - k)(m - imin), |coeff| = |k|/smax known only on the device:
        // dT/dm = 1 + s*coeff - k -> |.| <= 1 + |k| + 128|coeff|;  dT/ds = coeff (m - imin) -> <= 255 |coeff|
        // the FP32 rounding of coeff itself adds e*|coeff|*128*255
        Acoef = 1 + fabs(k); Bcoef = 0; Tmax = 255 * (1 + fabs(k));
        mu1 = 4 * (128 * dm + 255 * ds + e * 128 * 255 + 8 * e * 128 * 255);
        F->c0 = (float)k; F->c1 = F->c2 = 0; break;
    default: {   // Feng: T = p1*m + (k2*imin - imin)
        const double p1 = 1.0 + (1.0 - params[0]), k2 = params[2];
        Acoef = fabs(p1); Bcoef = 0; Tmax = 255 * fabs(p1) + 255 * (fabs(k2) + 1);
        F->c0 = 0; F->c1 = (float)p1; F->c2 = (float)k2; cd1 = p1; cd2 = k2; break;
    }
    }
    if (!(Tmax < 1e6)) return false;
    const double mu0 = 4 * (Acoef * dm + Bcoef * ds + 8 * e * (Tmax + 512));
    if (!(mu0 < 0.2)) return false;
    F->mu0 = (float)(mu0 < 2e-3 ? 2e-3 : mu0);
    F->mu1 = (float)mu1;
    F->kwf = (float)kw; F->inv_w2f = (float)kw;
    F->w2 = (unsigned int)(g.w * g.w);
    const double nf = s_floor * w2;
    F->n_floor = (float)(nf * nf * 1.0001);
    F->q_floor = (float)(nf * nf * 1.001 / (2.0 * g.w - 1.0));
    // tier 2: |T_fp64 - T_ref| <= A dm_ref + B dv_ref / s + FP64 rounding of the estimate (<= 64 u Tmax), times 4
    F->t2_a = 4 * (Acoef * dm_ref + Bcoef * u * 128 + 64 * u * (Tmax + 512));
    F->t2_b = 4 * Bcoef * dv_ref;
    F->t2_vmin = 16 * dv_ref;
    F->kw_d = kw; F->c0_d = cd0; F->c1_d = cd1; F->c2_d = cd2;
    F->dv_ref = dv_ref;
    // linear bounds of T over s in [0,128] (m >= 0); sqrt(m^2+s^2) in [m, m+128] for NICK
    double a0 = 1, a1 = 1, b0 = 0, b1 = 0;          // T(s=0) = a0*m + b0,  T(s=128) (or its bound) = a1*m + b1
    switch (method) {
    case PRL_SAUVOLA: a0 = cd2; a1 = cd2 + 128.0 * cd1; break;
    case PRL_NIBLACK: b1 = 128.0 * k; break;
    case PRL_NICK:    a0 = a1 = 1.0 + k; b1 = 128.0 * k; break;
    case PRL_FENG:    a0 = a1 = cd1; break;         // + (k2*imin - imin), added per page on the device
    default: break;
    }
    F->pa_hi = (float)(a0 > a1 ? a0 : a1); F->pa_lo = (float)(a0 > a1 ? a1 : a0);
    F->pb_hi = (float)(b0 > b1 ? b0 : b1); F->pb_lo = (float)(b0 > b1 ? b1 : b0);
    F->qfloor_u = (unsigned int)(nf * nf * 1.001 / (2.0 * g.w - 1.0)) + 2u;
    F->n32 = ((double)g.w * g.w * (double)g.d * g.d * 65025.0 < 4294967296.0) ? 1 : 0;
    F->dbg_skip_exact = 0;
    return true;
}


}  // namespace
