// threshold.cu -- kernel 2: fused 4-tap window sums + per-method threshold formula + u8 cast +
// compare, straight from the int64 integral planes to the 0/255 mask.
//
// Replaces, per reference function, everything between cv::filter2D and operator> :
//   Sauvola      binarizeSauvola.cpp:83-122       T = m * (s*(k/128) + (1-k))
//   Niblack      binarizeNiblack.cpp:83-112       T = m + k*s
//   Wolf-Jolion  binarizeWolfJolion.cpp:91-135    T = m + (s*(k/smax) - k) * (m - Imin)
//   NICK         binarizeNICK.cpp:91-131          T = m + k*sqrt(m*m + s*s)
//   Feng         binarizeFeng.cpp:87-148          T = (1 + (1-a1))*m + (k2*Imin - Imin)   (as written)
// with m = kw*a + (-kw)*b + (-kw)*c + kw*d over the four integral taps (every product rounded,
// no FMA: the reference's direct-tap filter2D path), s = sqrt(q - m*m), T8 = saturate(cvRound(T))
// with NaN / out-of-int32 -> 0, mask = p > T8 ? 255 : 0.
//
// Two kernels:
//  * threshold_exact_kernel -- the reference arithmetic literally, FP64 with __d*_rn intrinsics (the
//    compiler can never contract a multiply-add the CPU path does not contract).  Used for the T8
//    parity hook, the Wolf-Jolion s_max pass, odd tap distances and unaligned buffers.
//  * threshold_fast_kernel  -- the throughput path.  mask = (p > rint(T)) only asks on which side of
//    p - 1/2 the threshold lies, so:
//      1. window sums are formed in EXACT integer arithmetic (S_win, Q_win < 2^32, so the low 32 bits
//         of the int64 taps suffice; N = w^2*Q_win - S_win^2 is the exact variance numerator);
//      2. T is estimated in FP32 from those exact integers;
//      3. the pixel is decided from the estimate only if |p - 1/2 - T~| exceeds a margin `mu` that
//         bounds |T~ - T_ref| (host-derived from the FP64 rounding of the reference formula on this
//         page size + the FP32 rounding of the estimate, times 4); otherwise -- and for near-black
//         windows where sqrt is ill-conditioned -- the pixel runs the literal FP64 reference path.
//    The result is bit-identical to the exact kernel; FP64 is touched by ~1e-3 of the pixels.
//    Memory: each thread owns 4 adjacent columns (one 32-byte LDG.256 per plane and row side); the
//    vertical differences D = S[y+d] - S[y] go through shared memory so the x+d taps are never
//    re-read from L1/L2: 2 global loads per plane and pixel-quad instead of 4.
#include "common.cuh"

namespace {

struct ThrArgs {
    const uint8_t* src; size_t src_step, src_page_stride;
    const int64_t* S; const int64_t* Q; size_t pitch, plane_page_stride;
    uint8_t* dst; size_t dst_step, dst_page_stride;
    const uint32_t* imin; long long* smax;
    int out_rows, out_cols, d;
    double kw, nkw, p0, p1, p2;
};

struct FastArgs {
    float kwf, inv_w2f;          // 1/w^2 as float
    float c0, c1, c2;            // method constants in FP32
    float mu0, mu1;              // decision margin: mu = mu0 + mu1*|coeff|  (mu1 only for Wolf-Jolion)
    float n_floor;               // fast path only when N >= n_floor  (s* >= s_floor)
    unsigned int w2;             // w*w
    int rows_per_cta;
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

template <int METHOD>
__device__ __forceinline__ double thr_value(double m, double s, const ThrArgs& A, double imin, double coeff)
{
    if (METHOD == PRL_SAUVOLA) {
        return __dmul_rn(m, __dadd_rn(__dmul_rn(s, A.p1), A.p2));
    } else if (METHOD == PRL_NIBLACK) {
        return __dadd_rn(m, __dmul_rn(A.p0, s));
    } else if (METHOD == PRL_WOLFJOLION) {
        double dd = __dadd_rn(__dmul_rn(s, coeff), -A.p0);
        dd = __dmul_rn(dd, __dadd_rn(m, -imin));
        return __dadd_rn(m, dd);
    } else if (METHOD == PRL_NICK) {
        double C = __dsqrt_rn(__dadd_rn(__dmul_rn(m, m), __dmul_rn(s, s)));
        return __dadd_rn(m, __dmul_rn(C, A.p0));
    } else {
        // Feng as written: s/Rs with Rs aliasing s is 1, or NaN when s is 0 or NaN
        if (!(s == s) || s == 0.0) return __longlong_as_double(0x7ff8000000000000LL);
        double c3 = __dadd_rn(__dmul_rn(A.p2, imin), -imin);
        return __dadd_rn(__dmul_rn(A.p1, m), c3);   // p1 = 1 + (1 - alpha1), p2 = k2
    }
}

// The reference arithmetic for one output pixel, taps read as scalars (any alignment).
template <int METHOD>
__device__ __forceinline__ double thr_value_p(double m, double s, double p0, double p1, double p2, double imin, double coeff)
{
    if (METHOD == PRL_SAUVOLA) {
        return __dmul_rn(m, __dadd_rn(__dmul_rn(s, p1), p2));
    } else if (METHOD == PRL_NIBLACK) {
        return __dadd_rn(m, __dmul_rn(p0, s));
    } else if (METHOD == PRL_WOLFJOLION) {
        double dd = __dadd_rn(__dmul_rn(s, coeff), -p0);
        dd = __dmul_rn(dd, __dadd_rn(m, -imin));
        return __dadd_rn(m, dd);
    } else if (METHOD == PRL_NICK) {
        double C = __dsqrt_rn(__dadd_rn(__dmul_rn(m, m), __dmul_rn(s, s)));
        return __dadd_rn(m, __dmul_rn(C, p0));
    } else {
        if (!(s == s) || s == 0.0) return __longlong_as_double(0x7ff8000000000000LL);
        double c3 = __dadd_rn(__dmul_rn(p2, imin), -imin);
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

constexpr int kTR = 4;   // output rows per CTA (exact kernel)

__device__ __forceinline__ longlong2 ld2(const int64_t* p, bool aligned16)
{
    if (aligned16) return __ldg(reinterpret_cast<const longlong2*>(p));
    longlong2 r;
    r.x = __ldg(reinterpret_cast<const long long*>(p));
    r.y = __ldg(reinterpret_cast<const long long*>(p) + 1);
    return r;
}

// MODE 0: mask, 1: T8 map, 2: s_max reduction (Wolf-Jolion pass 1)
template <int METHOD, int MODE>
__global__ void __launch_bounds__(256)
threshold_exact_kernel(const ThrArgs A)
{
    const int page = blockIdx.z;
    const int x = (blockIdx.x * 256 + threadIdx.x) * 2;
    const int y0 = blockIdx.y * kTR;
    const int64_t* S = A.S + (size_t)page * A.plane_page_stride;
    const int64_t* Q = A.Q + (size_t)page * A.plane_page_stride;
    const uint8_t* src = A.src + (size_t)page * A.src_page_stride;

    double imin = 0.0, coeff = 0.0;
    if (METHOD == PRL_WOLFJOLION || METHOD == PRL_FENG) imin = (double)A.imin[page];
    if (METHOD == PRL_WOLFJOLION && MODE != 2)
        coeff = __ddiv_rn(A.p0, __longlong_as_double(A.smax[page]));   // coeff = k / devianceMax

    double smax_local = __longlong_as_double(0xfff0000000000000LL);   // -inf
    if (x < A.out_cols) {
#pragma unroll
        for (int r = 0; r < kTR; ++r) {
            const int y = y0 + r;
            if (y >= A.out_rows) break;
            const int64_t* s0 = S + (size_t)y * A.pitch + x;
            const int64_t* s1 = S + (size_t)(y + A.d) * A.pitch + x;
            const int64_t* q0 = Q + (size_t)y * A.pitch + x;
            const int64_t* q1 = Q + (size_t)(y + A.d) * A.pitch + x;
            const bool al = (A.d & 1) == 0;   // even tap distance (odd window): 16-byte aligned pairs
            const longlong2 sa = ld2(s0, true), sb = ld2(s0 + A.d, al), sc = ld2(s1, true), sd = ld2(s1 + A.d, al);
            const longlong2 qa = ld2(q0, true), qb = ld2(q0 + A.d, al), qc = ld2(q1, true), qd = ld2(q1 + A.d, al);
            const bool two = (x + 1) < A.out_cols;

            double m[2], s[2];
            m[0] = tap4(A.kw, A.nkw, sa.x, sb.x, sc.x, sd.x);
            m[1] = tap4(A.kw, A.nkw, sa.y, sb.y, sc.y, sd.y);
            s[0] = __dsqrt_rn(__dadd_rn(tap4(A.kw, A.nkw, qa.x, qb.x, qc.x, qd.x), -__dmul_rn(m[0], m[0])));
            s[1] = __dsqrt_rn(__dadd_rn(tap4(A.kw, A.nkw, qa.y, qb.y, qc.y, qd.y), -__dmul_rn(m[1], m[1])));

            if (MODE == 2) {
                if (s[0] > smax_local) smax_local = s[0];            // NaN never wins (cv::minMaxLoc)
                if (two && s[1] > smax_local) smax_local = s[1];
            } else {
                const int t0 = to_u8(thr_value<METHOD>(m[0], s[0], A, imin, coeff));
                const int t1 = to_u8(thr_value<METHOD>(m[1], s[1], A, imin, coeff));
                uint8_t* o = A.dst + (size_t)page * A.dst_page_stride + (size_t)y * A.dst_step + x;
                int o0, o1;
                if (MODE == 1) { o0 = t0; o1 = t1; }
                else {
                    const uint8_t* p = src + (size_t)y * A.src_step + x;
                    o0 = (int)p[0] > t0 ? 255 : 0;
                    o1 = two ? ((int)p[1] > t1 ? 255 : 0) : 0;
                }
                o[0] = (uint8_t)o0;
                if (two) o[1] = (uint8_t)o1;
            }
        }
    }
    if (MODE == 2) {
        // -inf has the most negative signed bit pattern of all candidates; non-negative doubles
        // order like signed integers -> deterministic integer atomicMax
        long long v = __double_as_longlong(smax_local);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            long long t = __shfl_xor_sync(0xffffffffu, v, o);
            v = t > v ? t : v;
        }
        __shared__ long long wmax[8];
        if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < 8; ++k) v = wmax[k] > v ? wmax[k] : v;
            atomicMax(A.smax + page, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// fast kernel
// ------------------------------------------------------------------------------------------------
constexpr int kFT = 128;              // threads per CTA
constexpr int kFC = kFT * 4;          // columns loaded per CTA row (512)
constexpr int kFR = 2;                // rows per iteration

__device__ __forceinline__ void ldg256(const int64_t* p, long long& a, long long& b, long long& c, long long& d)
{
    asm volatile("ld.global.nc.v4.s64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

// FP32 estimate of T from the exact window sums; returns false when the pixel must take the exact path
template <int METHOD>
__device__ __forceinline__ bool fast_decide(unsigned int sw, unsigned int qw, unsigned int p, const FastArgs& F,
                                            float iminf, float coefff, float mu, int& out)
{
    if (qw == 0u) { out = 0; return true; }            // all-zero window => p == 0 => (0 > T8) is false
    const unsigned long long N = (unsigned long long)F.w2 * qw - (unsigned long long)sw * sw;   // exact, >= 0
    const float fn = (float)N;
    const float m = (float)sw * F.kwf;
    const float s = sqrtf(fn) * F.inv_w2f;
    float T;
    if (METHOD == PRL_SAUVOLA) T = m * fmaf(s, F.c1, F.c2);
    else if (METHOD == PRL_NIBLACK) T = fmaf(F.c0, s, m);
    else if (METHOD == PRL_WOLFJOLION) T = fmaf(fmaf(s, coefff, -F.c0), m - iminf, m);
    else if (METHOD == PRL_NICK) T = fmaf(F.c0, sqrtf(fmaf(m, m, s * s)), m);
    else T = fmaf(F.c1, m, fmaf(F.c2, iminf, -iminf));
    const float g = ((float)p - 0.5f) - fmaxf(T, 0.0f);
    const bool ok = fn >= F.n_floor;
    if (ok && g > mu) { out = 255; return true; }
    if (ok && g < -mu) { out = 0; return true; }
    return false;                                       // near the rounding boundary, ill-conditioned, or NaN
}

template <int METHOD>
__global__ void __launch_bounds__(kFT)
threshold_fast_kernel(const ThrArgs A, const FastArgs F)
{
    __shared__ __align__(16) unsigned int sD[2][kFR][2][kFC];   // [buffer][row][plane S/Q][column]

    const int page = blockIdx.z;
    const int oc = (kFC - A.d) & ~3;                 // output columns per CTA
    const int X0 = blockIdx.x * oc;
    const int x = X0 + 4 * threadIdx.x;              // this thread's first column (32-byte aligned in the planes)
    const int y_begin = blockIdx.y * F.rows_per_cta;
    const int y_end = min(y_begin + F.rows_per_cta, A.out_rows);
    const int64_t* S = A.S + (size_t)page * A.plane_page_stride;
    const int64_t* Q = A.Q + (size_t)page * A.plane_page_stride;
    const uint8_t* src = A.src + (size_t)page * A.src_page_stride;
    uint8_t* dst = A.dst + (size_t)page * A.dst_page_stride;

    double imin = 0.0, coeff = 0.0;
    float iminf = 0.f, coefff = 0.f, mu = F.mu0;
    if (METHOD == PRL_WOLFJOLION || METHOD == PRL_FENG) { imin = (double)A.imin[page]; iminf = (float)imin; }
    if (METHOD == PRL_WOLFJOLION) {
        coeff = __ddiv_rn(A.p0, __longlong_as_double(A.smax[page]));
        coefff = (float)coeff;
        mu = F.mu0 + F.mu1 * fabsf(coefff);
        if (!(mu < 0.25f)) mu = 1e30f;               // degenerate s_max: every pixel takes the exact path
    }

    const bool in_plane = x < (int)A.pitch;          // pitch is a multiple of 16 columns
    const bool has_out = (4 * threadIdx.x + 3 + A.d < kFC) && (4 * (int)threadIdx.x < oc) && x < A.out_cols;
    const bool full4 = x + 3 < A.out_cols;

    int buf = 0;
    for (int y = y_begin; y < y_end; y += kFR, buf ^= 1) {
        // ---- vertical differences of the low words -> shared memory (own copy stays in registers)
        unsigned int dsr[kFR][4], dqr[kFR][4];
#pragma unroll
        for (int r = 0; r < kFR; ++r) {
            unsigned int (&ds)[4] = dsr[r];
            unsigned int (&dq)[4] = dqr[r];
            ds[0] = ds[1] = ds[2] = ds[3] = 0; dq[0] = dq[1] = dq[2] = dq[3] = 0;
            if (in_plane && y + r < y_end) {
                long long a0, a1, a2, a3, b0, b1, b2, b3;
                ldg256(S + (size_t)(y + r) * A.pitch + x, a0, a1, a2, a3);
                ldg256(S + (size_t)(y + r + A.d) * A.pitch + x, b0, b1, b2, b3);
                ds[0] = (unsigned int)b0 - (unsigned int)a0; ds[1] = (unsigned int)b1 - (unsigned int)a1;
                ds[2] = (unsigned int)b2 - (unsigned int)a2; ds[3] = (unsigned int)b3 - (unsigned int)a3;
                ldg256(Q + (size_t)(y + r) * A.pitch + x, a0, a1, a2, a3);
                ldg256(Q + (size_t)(y + r + A.d) * A.pitch + x, b0, b1, b2, b3);
                dq[0] = (unsigned int)b0 - (unsigned int)a0; dq[1] = (unsigned int)b1 - (unsigned int)a1;
                dq[2] = (unsigned int)b2 - (unsigned int)a2; dq[3] = (unsigned int)b3 - (unsigned int)a3;
            }
            *reinterpret_cast<uint4*>(&sD[buf][r][0][4 * threadIdx.x]) = make_uint4(ds[0], ds[1], ds[2], ds[3]);
            *reinterpret_cast<uint4*>(&sD[buf][r][1][4 * threadIdx.x]) = make_uint4(dq[0], dq[1], dq[2], dq[3]);
        }
        __syncthreads();
        // ---- horizontal differences, decision, store
        if (has_out) {
#pragma unroll
            for (int r = 0; r < kFR; ++r) {
                const int yy = y + r;
                if (yy >= y_end) break;
                const unsigned int* ls = &sD[buf][r][0][4 * threadIdx.x];
                const unsigned int* lq = &sD[buf][r][1][4 * threadIdx.x];
                const uint4 s_l = make_uint4(dsr[r][0], dsr[r][1], dsr[r][2], dsr[r][3]);
                const uint4 q_l = make_uint4(dqr[r][0], dqr[r][1], dqr[r][2], dqr[r][3]);
                const uint2 s_r0 = *reinterpret_cast<const uint2*>(ls + A.d);       // d even -> 8-byte aligned
                const uint2 s_r1 = *reinterpret_cast<const uint2*>(ls + A.d + 2);
                const uint2 q_r0 = *reinterpret_cast<const uint2*>(lq + A.d);
                const uint2 q_r1 = *reinterpret_cast<const uint2*>(lq + A.d + 2);
                const unsigned int sw[4] = {s_r0.x - s_l.x, s_r0.y - s_l.y, s_r1.x - s_l.z, s_r1.y - s_l.w};
                const unsigned int qw[4] = {q_r0.x - q_l.x, q_r0.y - q_l.y, q_r1.x - q_l.z, q_r1.y - q_l.w};
                const uint8_t* prow = src + (size_t)yy * A.src_step + x;
                unsigned int p4;
                if (full4) p4 = __ldg(reinterpret_cast<const unsigned int*>(prow));
                else {
                    p4 = 0;
                    for (int i = 0; i < 4; ++i) if (x + i < A.out_cols) p4 |= (unsigned int)prow[i] << (8 * i);
                }
                unsigned int o4 = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const unsigned int p = (p4 >> (8 * i)) & 0xffu;
                    int o;
                    if (!fast_decide<METHOD>(sw[i], qw[i], p, F, iminf, coefff, mu, o)) {
                        if (x + i < A.out_cols) {
                            const size_t e0 = (size_t)yy * A.pitch + x + i;
                            const int t8 = exact_t8_at<METHOD>(reinterpret_cast<const long long*>(S) + e0,
                                                               reinterpret_cast<const long long*>(Q) + e0, (size_t)A.d * A.pitch,
                                                               A.d, A.kw, A.p0, A.p1, A.p2, imin, coeff);
                            o = (int)p > t8 ? 255 : 0;
                        } else o = 0;
                    }
                    o4 |= (unsigned int)o << (8 * i);
                }
                // dst may be dense (pitch == out_cols, odd): store with whatever alignment the row has
                uint8_t* orow = dst + (size_t)yy * A.dst_step + x;
                const unsigned int al = (unsigned int)(uintptr_t)orow & 3u;
                if (!full4) {
                    for (int i = 0; i < 4; ++i) if (x + i < A.out_cols) orow[i] = (uint8_t)(o4 >> (8 * i));
                } else if (al == 0) {
                    *reinterpret_cast<unsigned int*>(orow) = o4;
                } else if (al == 2) {
                    *reinterpret_cast<unsigned short*>(orow) = (unsigned short)o4;
                    *reinterpret_cast<unsigned short*>(orow + 2) = (unsigned short)(o4 >> 16);
                } else {
                    orow[0] = (uint8_t)o4;
                    *reinterpret_cast<unsigned short*>(orow + 1) = (unsigned short)(o4 >> 8);
                    orow[3] = (uint8_t)(o4 >> 24);
                }
            }
        }
    }
}

template <int METHOD>
void launch_exact(prl_cuda_ctx* ctx, int mode, const ThrArgs& A, dim3 grid)
{
    switch (mode) {
    case 0: threshold_exact_kernel<METHOD, 0><<<grid, 256, 0, ctx->stream>>>(A); break;
    case 1: threshold_exact_kernel<METHOD, 1><<<grid, 256, 0, ctx->stream>>>(A); break;
    default: threshold_exact_kernel<METHOD, 2><<<grid, 256, 0, ctx->stream>>>(A); break;
    }
}

__global__ void init_smax_kernel(long long* smax, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) smax[i] = (long long)0xfff0000000000000LL;   // -inf
}

// Host-side error analysis for the fast path: returns false when the margin is too large to be useful.
//   reference FP64 error (vs exact real arithmetic), u = 2^-53:
//     dm_ref <= 16 u kw Smax,  dq_ref <= 16 u kw Qmax,  dv_ref <= dq_ref + 2*255*dm_ref + u*2*255^2
//     ds_ref <= dv_ref / s_floor + u*128          (fast path requires s* >= s_floor)
//   FP32 estimate error (exact integer inputs), e = 2^-24:
//     dm_est <= 3 e 255,  ds_est <= 4 e 128
//   |dT| <= A dm + B ds + 8 e (Tmax + 512), A/B = sup |dT/dm|, |dT/ds| over m in [0,255], s in [0,128]
bool fast_margins(int method, const double* params, const prl_geom& g, FastArgs* F)
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
    const double dm = dm_ref + 3 * e * 255, ds = ds_ref + 4 * e * 128;
    double Acoef, Bcoef, Tmax, mu1 = 0.0;
    const double k = params[0];
    switch (method) {
    case PRL_SAUVOLA: {
        const double c1 = k * (1.0 / 128.0), c2 = 1.0 - k;
        Acoef = fabs(c2) + 128 * fabs(c1); Bcoef = 255 * fabs(c1); Tmax = 255 * Acoef;
        F->c0 = (float)k; F->c1 = (float)c1; F->c2 = (float)c2; break;
    }
    case PRL_NIBLACK:
        Acoef = 1; Bcoef = fabs(k); Tmax = 255 + 128 * fabs(k);
        F->c0 = (float)k; F->c1 = F->c2 = 0; break;
    case PRL_NICK:
        Acoef = 1 + fabs(k); Bcoef = fabs(k); Tmax = 255 + fabs(k) * 286;
        F->c0 = (float)k; F->c1 = F->c2 = 0; break;
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
        F->c0 = 0; F->c1 = (float)p1; F->c2 = (float)k2; break;
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
    return true;
}

}  // namespace

int prl_k_threshold(prl_cuda_ctx* ctx, int method, int mode, const uint8_t* d_src, int n_pages,
                    const prl_geom& g, size_t src_step, size_t src_page_stride, const int64_t* d_S,
                    const int64_t* d_Q, size_t plane_page_stride, const double* params,
                    const uint32_t* d_imin, long long* d_smax, uint8_t* d_dst, size_t dst_step,
                    size_t dst_page_stride)
{
    ThrArgs A;
    A.src = d_src; A.src_step = src_step; A.src_page_stride = src_page_stride;
    A.S = d_S; A.Q = d_Q; A.pitch = g.pitch; A.plane_page_stride = plane_page_stride;
    A.dst = d_dst; A.dst_step = dst_step; A.dst_page_stride = dst_page_stride;
    A.imin = d_imin; A.smax = d_smax;
    A.out_rows = g.out_rows; A.out_cols = g.out_cols; A.d = g.d;
    A.kw = 1.0 / (double)(g.w * g.w);     // wSqrBack, binarizeSauvola.cpp:58-59
    A.nkw = -A.kw;
    A.p0 = params[0]; A.p1 = 0; A.p2 = 0;
    if (method == PRL_SAUVOLA) { A.p1 = params[0] * (1.0 / 128.0); A.p2 = 1.0 - params[0]; }   // (k*RBack), (1-k) :115-117
    if (method == PRL_FENG) { A.p1 = 1.0 + (1.0 - params[0]); A.p2 = params[2]; }              // c2 + c1, k2
    if (method < PRL_SAUVOLA || method > PRL_FENG) return prl_set_err(ctx, PRL_E_INVALID, "unknown method");

    dim3 grid((g.out_cols + 511) / 512, (g.out_rows + kTR - 1) / kTR, n_pages);
    if (n_pages > 65535 || grid.y > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");

    if (method == PRL_WOLFJOLION) {
        {
            prl_launch_scope ls(ctx, FAM_SMAX);
            init_smax_kernel<<<(n_pages + 255) / 256, 256, 0, ctx->stream>>>(d_smax, n_pages);
        }
        prl_launch_scope ls(ctx, FAM_SMAX);
        launch_exact<PRL_WOLFJOLION>(ctx, 2, A, grid);
    }

    // fast path eligibility: mask output, even tap distance < 256, window sums < 2^32, 4/32-byte aligned buffers
    FastArgs F;
    const bool aligned = ((src_step | src_page_stride | (uintptr_t)d_src) & 3) == 0 &&
                         ((((uintptr_t)d_S) | ((uintptr_t)d_Q)) & 31) == 0 && (plane_page_stride & 3) == 0 && (g.pitch & 3) == 0;
    const bool fast = mode == 0 && !ctx->force_exact && (g.d & 1) == 0 && g.d <= 254 && aligned &&
                      fast_margins(method, params, g, &F);
    prl_launch_scope ls(ctx, FAM_THRESHOLD);
    if (fast) {
        const int oc = (kFC - g.d) & ~3;
        // Rows per CTA.  Each S/Q row is fetched twice (as the bottom row of output row y-d, then as the
        // top row of y); the second fetch must hit L2, so the CTAs in flight have to cover a compact
        // set of rows: short tiles issued in raster order keep the live set at ~(in-flight rows + d)
        // rows of one or two pages (tens of MB), tall tiles thrash the 126 MB L2 (measured: 2x DRAM reads).
        int rpc = 4;
        F.rows_per_cta = rpc;
        dim3 fg((g.out_cols + oc - 1) / oc, (g.out_rows + rpc - 1) / rpc, n_pages);
        switch (method) {
        case PRL_SAUVOLA:    threshold_fast_kernel<PRL_SAUVOLA><<<fg, kFT, 0, ctx->stream>>>(A, F); break;
        case PRL_NIBLACK:    threshold_fast_kernel<PRL_NIBLACK><<<fg, kFT, 0, ctx->stream>>>(A, F); break;
        case PRL_WOLFJOLION: threshold_fast_kernel<PRL_WOLFJOLION><<<fg, kFT, 0, ctx->stream>>>(A, F); break;
        case PRL_NICK:       threshold_fast_kernel<PRL_NICK><<<fg, kFT, 0, ctx->stream>>>(A, F); break;
        default:             threshold_fast_kernel<PRL_FENG><<<fg, kFT, 0, ctx->stream>>>(A, F); break;
        }
    } else {
        switch (method) {
        case PRL_SAUVOLA:    launch_exact<PRL_SAUVOLA>(ctx, mode, A, grid); break;
        case PRL_NIBLACK:    launch_exact<PRL_NIBLACK>(ctx, mode, A, grid); break;
        case PRL_WOLFJOLION: launch_exact<PRL_WOLFJOLION>(ctx, mode, A, grid); break;
        case PRL_NICK:       launch_exact<PRL_NICK>(ctx, mode, A, grid); break;
        default:             launch_exact<PRL_FENG>(ctx, mode, A, grid); break;
        }
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
