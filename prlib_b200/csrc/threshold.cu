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
#include "decide.cuh"

namespace {

struct ThrArgs {
    const uint8_t* src; size_t src_step, src_page_stride;
    const int64_t* S; const int64_t* Q; size_t pitch, plane_page_stride;   // compact layout: really uint32_t* (low words)
    const uint32_t* AS; const uint32_t* AQ; size_t a_page_stride; int ashift; // compact layout: high words of the anchor rows
    uint8_t* dst; size_t dst_step, dst_page_stride;
    const uint32_t* imin; long long* smax;
    int out_rows, out_cols, d;
    double kw, nkw, p0, p1, p2;
};

// one spelling of the reference's formulas for every kernel: decide.cuh:thr_value_p
template <int METHOD>
__device__ __forceinline__ double thr_value(double m, double s, const ThrArgs& A, double imin, double coeff)
{
    return thr_value_p<METHOD>(m, s, A.p0, A.p1, A.p2, imin, coeff);
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
__device__ __forceinline__ uint4 ldg128u(const uint32_t* p)
{
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// low words of 4 adjacent plane elements at element offset `e` of a page's plane, either layout
template <bool COMPACT>
__device__ __forceinline__ uint4 lo4(const int64_t* plane, size_t e)
{
    if (COMPACT) return ldg128u(reinterpret_cast<const uint32_t*>(plane) + e);
    long long a, b, c, d;
    ldg256(plane + e, a, b, c, d);
    return make_uint4((unsigned int)a, (unsigned int)b, (unsigned int)c, (unsigned int)d);
}

// compact layout: the full int64 value at (Y, X) from the low-word plane L and the anchor high words H (common.cuh: prl_planes)
__device__ __forceinline__ long long full_tap(const uint32_t* __restrict__ L, const uint32_t* __restrict__ H, size_t pitch,
                                              int ashift, int Y, int X)
{
    const int ya = Y >> ashift;
    const unsigned int lo_a = __ldg(L + ((size_t)ya << ashift) * pitch + X);
    const unsigned int hi = __ldg(H + (size_t)ya * pitch + X);
    const unsigned int lo = __ldg(L + (size_t)Y * pitch + X);
    return (long long)((((unsigned long long)hi << 32) | lo_a) + (unsigned long long)(unsigned int)(lo - lo_a));
}

// the literal reference arithmetic for output pixel (y, x) from compact planes of one page
template <int METHOD>
__device__ __noinline__ int exact_t8_compact(const uint32_t* __restrict__ S, const uint32_t* __restrict__ Q,
                                             const uint32_t* __restrict__ AS, const uint32_t* __restrict__ AQ, size_t pitch,
                                             int ashift, int y, int x, int d, double kw, double p0, double p1, double p2,
                                             double imin, double coeff)
{
    const long long sa = full_tap(S, AS, pitch, ashift, y, x), sb = full_tap(S, AS, pitch, ashift, y, x + d);
    const long long sc = full_tap(S, AS, pitch, ashift, y + d, x), sd = full_tap(S, AS, pitch, ashift, y + d, x + d);
    const long long qa = full_tap(Q, AQ, pitch, ashift, y, x), qb = full_tap(Q, AQ, pitch, ashift, y, x + d);
    const long long qc = full_tap(Q, AQ, pitch, ashift, y + d, x), qd = full_tap(Q, AQ, pitch, ashift, y + d, x + d);
    return exact_t8_from_taps<METHOD>(sa, sb, sc, sd, qa, qb, qc, qd, kw, p0, p1, p2, imin, coeff);
}

// NT threads per CTA: 128 (512 columns loaded per row) for small windows, 256 (1024 columns) when the tap
// distance d would otherwise waste a large share of each CTA's loads on the column halo
template <int METHOD, int NT, bool COMPACT>
__global__ void __launch_bounds__(NT, 1024 / NT)     // <= 64 registers: 32 resident warps per SM
threshold_fast_kernel(const ThrArgs A, const FastArgs F)
{
    constexpr int FC = NT * 4;                        // columns loaded per CTA row
    __shared__ __align__(16) unsigned int sD[2][kFR][2][FC];   // [buffer][row][plane S/Q][column]

    const int page = blockIdx.z;
    const int oc = (FC - A.d) & ~3;                  // output columns per CTA
    const int X0 = blockIdx.x * oc;
    const int x = X0 + 4 * threadIdx.x;              // this thread's first column (32-byte aligned in the planes)
    const int y_begin = blockIdx.y * F.rows_per_cta;
    const int y_end = min(y_begin + F.rows_per_cta, A.out_rows);
    // (compact layout: the same expressions step through u32 elements, see lo4)
    const int64_t* S = COMPACT ? reinterpret_cast<const int64_t*>(reinterpret_cast<const uint32_t*>(A.S) + (size_t)page * A.plane_page_stride)
                               : A.S + (size_t)page * A.plane_page_stride;
    const int64_t* Q = COMPACT ? reinterpret_cast<const int64_t*>(reinterpret_cast<const uint32_t*>(A.Q) + (size_t)page * A.plane_page_stride)
                               : A.Q + (size_t)page * A.plane_page_stride;
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
    const bool has_out = (4 * threadIdx.x + 3 + A.d < FC) && (4 * (int)threadIdx.x < oc) && x < A.out_cols;
    const bool full4 = x + 3 < A.out_cols;

    int buf = 0;
    for (int y = y_begin; y < y_end; y += kFR, buf ^= 1) {
        // ---- vertical differences of the low words -> shared memory (own copy stays in registers)
        unsigned int dsr[kFR][4], dqr[kFR][4];
        if (COMPACT) {
            // 16-byte loads: all eight of an iteration are in flight before the first difference is formed
            uint4 ta[kFR], tb[kFR], ua[kFR], ub[kFR];
#pragma unroll
            for (int r = 0; r < kFR; ++r) {
                ta[r] = tb[r] = ua[r] = ub[r] = make_uint4(0u, 0u, 0u, 0u);
                if (in_plane && y + r < y_end) {
                    const size_t e = (size_t)(y + r) * A.pitch + x, eb = e + (size_t)A.d * A.pitch;
                    ta[r] = lo4<true>(S, e); tb[r] = lo4<true>(S, eb);
                    ua[r] = lo4<true>(Q, e); ub[r] = lo4<true>(Q, eb);
                }
            }
#pragma unroll
            for (int r = 0; r < kFR; ++r) {
                dsr[r][0] = tb[r].x - ta[r].x; dsr[r][1] = tb[r].y - ta[r].y; dsr[r][2] = tb[r].z - ta[r].z; dsr[r][3] = tb[r].w - ta[r].w;
                dqr[r][0] = ub[r].x - ua[r].x; dqr[r][1] = ub[r].y - ua[r].y; dqr[r][2] = ub[r].z - ua[r].z; dqr[r][3] = ub[r].w - ua[r].w;
                *reinterpret_cast<uint4*>(&sD[buf][r][0][4 * threadIdx.x]) = make_uint4(dsr[r][0], dsr[r][1], dsr[r][2], dsr[r][3]);
                *reinterpret_cast<uint4*>(&sD[buf][r][1][4 * threadIdx.x]) = make_uint4(dqr[r][0], dqr[r][1], dqr[r][2], dqr[r][3]);
            }
        } else {
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
                    if (!fast_decide<METHOD, METHOD == PRL_SAUVOLA>(sw[i], qw[i], p, F, iminf, coefff, mu, o)) {
                        if (x + i < A.out_cols) {
                            int t8;
                            if (COMPACT) {
                                t8 = exact_t8_compact<METHOD>(reinterpret_cast<const uint32_t*>(S), reinterpret_cast<const uint32_t*>(Q),
                                                              A.AS + (size_t)page * A.a_page_stride, A.AQ + (size_t)page * A.a_page_stride,
                                                              A.pitch, A.ashift, yy, x + i, A.d, A.kw, A.p0, A.p1, A.p2, imin, coeff);
                            } else {
                                const size_t e0 = (size_t)yy * A.pitch + x + i;
                                t8 = exact_t8_at<METHOD>(reinterpret_cast<const long long*>(S) + e0,
                                                         reinterpret_cast<const long long*>(Q) + e0, (size_t)A.d * A.pitch,
                                                         A.d, A.kw, A.p0, A.p1, A.p2, imin, coeff);
                            }
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

// ------------------------------------------------------------------------------------------------
// Wolf-Jolion s_max (cv::minMaxLoc(localDevianceValues), binarizeWolfJolion.cpp:118-119) without running
// the FP64 formula on every pixel.  s_ref = sqrt(fl(q_ref - m_ref^2)) and |fl(q_ref - m_ref^2) - N/w^4| <= dv
// (the reference's own rounding, bounded on the host), with N = w^2 Q_win - S_win^2 the EXACT integer
// variance numerator.  sqrt is monotone, so the pixel that attains s_max has N >= N_max - 2 dv w^4:
//   pass A (nmax_fast_kernel): exact N for every pixel in integer arithmetic -> N_max per page and per tile;
//   pass B (smax_candidates_kernel): only tiles whose maximum reaches N_max - margin are revisited, and only
//   their pixels within the margin evaluate the literal FP64 s; the maximum over those is s_max, bit-exact.
// ------------------------------------------------------------------------------------------------
template <bool COMPACT>
__global__ void __launch_bounds__(kFT)
nmax_fast_kernel(const ThrArgs A, const FastArgs F, unsigned long long* __restrict__ nmax_page,
                 unsigned long long* __restrict__ nmax_tile)
{
    __shared__ __align__(16) unsigned int sD[2][kFR][2][kFC];
    __shared__ unsigned long long wmax[kFT / 32];
    const int page = blockIdx.z;
    const int oc = (kFC - A.d) & ~3;
    const int X0 = blockIdx.x * oc;
    const int x = X0 + 4 * threadIdx.x;
    const int y_begin = blockIdx.y * F.rows_per_cta;
    const int y_end = min(y_begin + F.rows_per_cta, A.out_rows);
    const int64_t* S = COMPACT ? reinterpret_cast<const int64_t*>(reinterpret_cast<const uint32_t*>(A.S) + (size_t)page * A.plane_page_stride)
                               : A.S + (size_t)page * A.plane_page_stride;
    const int64_t* Q = COMPACT ? reinterpret_cast<const int64_t*>(reinterpret_cast<const uint32_t*>(A.Q) + (size_t)page * A.plane_page_stride)
                               : A.Q + (size_t)page * A.plane_page_stride;
    const bool in_plane = x < (int)A.pitch;
    const bool has_out = (4 * threadIdx.x + 3 + A.d < kFC) && (4 * (int)threadIdx.x < oc) && x < A.out_cols;
    unsigned long long best = 0;
    int buf = 0;
    for (int y = y_begin; y < y_end; y += kFR, buf ^= 1) {
        unsigned int dsr[kFR][4], dqr[kFR][4];
#pragma unroll
        for (int r = 0; r < kFR; ++r) {
            unsigned int (&ds)[4] = dsr[r];
            unsigned int (&dq)[4] = dqr[r];
            ds[0] = ds[1] = ds[2] = ds[3] = 0; dq[0] = dq[1] = dq[2] = dq[3] = 0;
            if (in_plane && y + r < y_end) {
                const size_t e = (size_t)(y + r) * A.pitch + x, eb = e + (size_t)A.d * A.pitch;
                const uint4 sa = lo4<COMPACT>(S, e), sb = lo4<COMPACT>(S, eb), qa = lo4<COMPACT>(Q, e), qb = lo4<COMPACT>(Q, eb);
                ds[0] = sb.x - sa.x; ds[1] = sb.y - sa.y; ds[2] = sb.z - sa.z; ds[3] = sb.w - sa.w;
                dq[0] = qb.x - qa.x; dq[1] = qb.y - qa.y; dq[2] = qb.z - qa.z; dq[3] = qb.w - qa.w;
            }
            *reinterpret_cast<uint4*>(&sD[buf][r][0][4 * threadIdx.x]) = make_uint4(ds[0], ds[1], ds[2], ds[3]);
            *reinterpret_cast<uint4*>(&sD[buf][r][1][4 * threadIdx.x]) = make_uint4(dq[0], dq[1], dq[2], dq[3]);
        }
        __syncthreads();
        if (has_out) {
#pragma unroll
            for (int r = 0; r < kFR; ++r) {
                if (y + r >= y_end) break;
                const unsigned int* ls = &sD[buf][r][0][4 * threadIdx.x];
                const unsigned int* lq = &sD[buf][r][1][4 * threadIdx.x];
                const uint2 s_r0 = *reinterpret_cast<const uint2*>(ls + A.d), s_r1 = *reinterpret_cast<const uint2*>(ls + A.d + 2);
                const uint2 q_r0 = *reinterpret_cast<const uint2*>(lq + A.d), q_r1 = *reinterpret_cast<const uint2*>(lq + A.d + 2);
                const unsigned int sw[4] = {s_r0.x - dsr[r][0], s_r0.y - dsr[r][1], s_r1.x - dsr[r][2], s_r1.y - dsr[r][3]};
                const unsigned int qw[4] = {q_r0.x - dqr[r][0], q_r0.y - dqr[r][1], q_r1.x - dqr[r][2], q_r1.y - dqr[r][3]};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (x + i < A.out_cols) {
                        const unsigned long long N = (unsigned long long)F.w2 * qw[i] - (unsigned long long)sw[i] * sw[i];
                        best = N > best ? N : best;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, best, o);
        best = t > best ? t : best;
    }
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < kFT / 32; ++k) best = wmax[k] > best ? wmax[k] : best;
        nmax_tile[((size_t)page * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = best;
        atomicMax(nmax_page + page, best);
    }
}

// One CTA revisits up to 16 consecutive tiles of pass A (same tile geometry: oc columns x rows_per_cta rows).
template <bool COMPACT>
__global__ void __launch_bounds__(128)
smax_candidates_kernel(const ThrArgs A, const FastArgs F, const unsigned long long* __restrict__ nmax_page,
                       const unsigned long long* __restrict__ nmax_tile, int tiles_x, int tiles_y,
                       unsigned long long margin)
{
    __shared__ long long wmax[4];
    const int page = blockIdx.y;
    const unsigned long long nmax = nmax_page[page];
    const unsigned long long lo = nmax > margin ? nmax - margin : 0ull;
    const int oc = (kFC - A.d) & ~3;
    const long long* S = reinterpret_cast<const long long*>(A.S) + (size_t)page * A.plane_page_stride;
    const long long* Q = reinterpret_cast<const long long*>(A.Q) + (size_t)page * A.plane_page_stride;
    const uint32_t* S32 = reinterpret_cast<const uint32_t*>(A.S) + (size_t)page * A.plane_page_stride;
    const uint32_t* Q32 = reinterpret_cast<const uint32_t*>(A.Q) + (size_t)page * A.plane_page_stride;
    const uint32_t* AS = COMPACT ? A.AS + (size_t)page * A.a_page_stride : nullptr;
    const uint32_t* AQ = COMPACT ? A.AQ + (size_t)page * A.a_page_stride : nullptr;
    long long best = (long long)0xfff0000000000000LL;   // -inf
    const int tiles = tiles_x * tiles_y;
    for (int t = blockIdx.x * 16; t < min(tiles, blockIdx.x * 16 + 16); ++t) {
        if (nmax_tile[(size_t)page * tiles + t] < lo) continue;             // uniform across the CTA
        const int tx = t % tiles_x, ty = t / tiles_x;
        const int x_begin = tx * oc, x_end = min(x_begin + oc, A.out_cols);
        const int y_begin = ty * F.rows_per_cta, y_end = min(y_begin + F.rows_per_cta, A.out_rows);
        const int w = x_end - x_begin, n = w * (y_end - y_begin);
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int y = y_begin + i / w, x = x_begin + i % w;
            long long sa, sb, sc, sd, qa, qb, qc, qd;
            const size_t dr = (size_t)A.d * A.pitch;
            if (COMPACT) {
                // the integer screen needs the low words only; the full taps are rebuilt for the few candidates
                const uint32_t* s0 = S32 + (size_t)y * A.pitch + x;
                const uint32_t* q0 = Q32 + (size_t)y * A.pitch + x;
                const unsigned int swl = (__ldg(s0 + dr + A.d) - __ldg(s0 + dr)) - (__ldg(s0 + A.d) - __ldg(s0));
                const unsigned int qwl = (__ldg(q0 + dr + A.d) - __ldg(q0 + dr)) - (__ldg(q0 + A.d) - __ldg(q0));
                const unsigned long long Nl = (unsigned long long)F.w2 * qwl - (unsigned long long)swl * swl;
                if (Nl < lo) continue;
                sa = full_tap(S32, AS, A.pitch, A.ashift, y, x); sb = full_tap(S32, AS, A.pitch, A.ashift, y, x + A.d);
                sc = full_tap(S32, AS, A.pitch, A.ashift, y + A.d, x); sd = full_tap(S32, AS, A.pitch, A.ashift, y + A.d, x + A.d);
                qa = full_tap(Q32, AQ, A.pitch, A.ashift, y, x); qb = full_tap(Q32, AQ, A.pitch, A.ashift, y, x + A.d);
                qc = full_tap(Q32, AQ, A.pitch, A.ashift, y + A.d, x); qd = full_tap(Q32, AQ, A.pitch, A.ashift, y + A.d, x + A.d);
            } else {
                const long long* s0 = S + (size_t)y * A.pitch + x;
                const long long* q0 = Q + (size_t)y * A.pitch + x;
                sa = __ldg(s0); sb = __ldg(s0 + A.d); sc = __ldg(s0 + dr); sd = __ldg(s0 + dr + A.d);
                qa = __ldg(q0); qb = __ldg(q0 + A.d); qc = __ldg(q0 + dr); qd = __ldg(q0 + dr + A.d);
            }
            const unsigned long long sw = (unsigned long long)((sd - sc) - (sb - sa));
            const unsigned long long qw = (unsigned long long)((qd - qc) - (qb - qa));
            const unsigned long long N = (unsigned long long)F.w2 * qw - sw * sw;
            if (N < lo) continue;
            const double m = tap4(A.kw, A.nkw, sa, sb, sc, sd);
            const double q = tap4(A.kw, A.nkw, qa, qb, qc, qd);
            const double sdev = __dsqrt_rn(__dadd_rn(q, -__dmul_rn(m, m)));
            if (sdev == sdev) {                                              // NaN never wins (cv::minMaxLoc)
                const long long bits = __double_as_longlong(sdev);
                best = bits > best ? bits : best;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long t = __shfl_xor_sync(0xffffffffu, best, o);
        best = t > best ? t : best;
    }
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 4; ++k) best = wmax[k] > best ? wmax[k] : best;
        if (best != (long long)0xfff0000000000000LL) atomicMax(A.smax + page, best);
    }
}

__global__ void init_smax_kernel(long long* smax, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) smax[i] = (long long)0xfff0000000000000LL;   // -inf
}

}  // namespace

// What kernel 2's fast path needs besides suitable planes: an even tap distance below 256, aligned pages, and margins the
// host error analysis accepts.  The compact plane layout exists only for this path, so its callers ask first.
bool prl_threshold_fast_ok(const prl_cuda_ctx* ctx, int method, const double* params, const prl_geom& g,
                           const uint8_t* d_src, size_t src_step, size_t src_page_stride)
{
    FastArgs F;
    return !ctx->force_exact && (g.d & 1) == 0 && g.d <= 254 && ((src_step | src_page_stride | (uintptr_t)d_src) & 3) == 0 &&
           fast_margins(method, params, g, &F);
}

int prl_k_threshold(prl_cuda_ctx* ctx, int method, int mode, const uint8_t* d_src, int n_pages,
                    const prl_geom& g, size_t src_step, size_t src_page_stride, const prl_planes& P, const double* params,
                    const uint32_t* d_imin, long long* d_smax, uint8_t* d_dst, size_t dst_step,
                    size_t dst_page_stride)
{
    const int64_t* d_S = (const int64_t*)P.S;
    const int64_t* d_Q = (const int64_t*)P.Q;
    const size_t plane_page_stride = P.page_stride;
    ThrArgs A;
    A.src = d_src; A.src_step = src_step; A.src_page_stride = src_page_stride;
    A.S = d_S; A.Q = d_Q; A.pitch = P.pitch; A.plane_page_stride = plane_page_stride;
    A.AS = P.AS; A.AQ = P.AQ; A.a_page_stride = P.a_page_stride; A.ashift = P.ashift;
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

    // fast path eligibility: mask output, even tap distance < 256, window sums < 2^32, 4/32-byte aligned buffers
    FastArgs F;
    const bool aligned = ((src_step | src_page_stride | (uintptr_t)d_src) & 3) == 0 &&
                         ((((uintptr_t)d_S) | ((uintptr_t)d_Q)) & (P.compact ? 15 : 31)) == 0 && (plane_page_stride & 3) == 0 && (P.pitch & 3) == 0;
    const bool fast_ok = !ctx->force_exact && (g.d & 1) == 0 && g.d <= 254 && aligned && fast_margins(method, params, g, &F);
    if (P.compact && !(fast_ok && mode == 0))
        return prl_set_err(ctx, PRL_E_INVALID, "compact planes serve the fast mask path only");
    const int rpc = ctx->thr_rows > 0 ? ctx->thr_rows : (g.d > 64 ? 8 : 4);   // rows per CTA of the fast kernels, see below (8 measured 4 % faster for wide windows)
    F.rows_per_cta = rpc;

    if (method == PRL_WOLFJOLION) {
        {
            prl_launch_scope ls(ctx, FAM_SMAX);
            init_smax_kernel<<<(n_pages + 255) / 256, 256, 0, ctx->stream>>>(d_smax, n_pages);
        }
        if (fast_ok) {
            const int oc = (kFC - g.d) & ~3;
            const int tiles_x = (g.out_cols + oc - 1) / oc, tiles_y = (g.out_rows + rpc - 1) / rpc;
            const size_t need = ((size_t)n_pages + (size_t)n_pages * tiles_x * tiles_y) * sizeof(unsigned long long);
            int rc = prl_ensure(ctx, &ctx->d_misc, &ctx->d_misc_bytes, need); if (rc) return rc;
            unsigned long long* nmax_page = (unsigned long long*)ctx->d_misc;
            unsigned long long* nmax_tile = nmax_page + n_pages;
            PRL_CUDA_TRY(ctx, cudaMemsetAsync(nmax_page, 0, sizeof(unsigned long long) * n_pages, ctx->stream));
            // |fl(q_ref - m_ref^2) - N/w^4| <= dv (same bound as fast_margins) -> candidates have N >= N_max - 2 dv w^4
            const double u = 1.1102230246251565e-16, w2 = (double)g.w * g.w, area = (double)g.Hp * (double)g.Wp;
            const double dv = 16 * u / w2 * 65025.0 * area + 510.0 * (16 * u / w2 * 255.0 * area) + u * 2 * 65025.0;
            const unsigned long long margin = (unsigned long long)(2.0 * dv * w2 * w2) + 2;
            {
                prl_launch_scope ls(ctx, FAM_SMAX);
                if (P.compact) nmax_fast_kernel<true><<<dim3(tiles_x, tiles_y, n_pages), kFT, 0, ctx->stream>>>(A, F, nmax_page, nmax_tile);
                else nmax_fast_kernel<false><<<dim3(tiles_x, tiles_y, n_pages), kFT, 0, ctx->stream>>>(A, F, nmax_page, nmax_tile);
            }
            {
                prl_launch_scope ls(ctx, FAM_SMAX);
                if (P.compact)
                    smax_candidates_kernel<true><<<dim3((tiles_x * tiles_y + 15) / 16, n_pages), 128, 0, ctx->stream>>>(
                        A, F, nmax_page, nmax_tile, tiles_x, tiles_y, margin);
                else
                    smax_candidates_kernel<false><<<dim3((tiles_x * tiles_y + 15) / 16, n_pages), 128, 0, ctx->stream>>>(
                        A, F, nmax_page, nmax_tile, tiles_x, tiles_y, margin);
            }
        } else {
            prl_launch_scope ls(ctx, FAM_SMAX);
            launch_exact<PRL_WOLFJOLION>(ctx, 2, A, grid);
        }
    }

    const bool fast = mode == 0 && fast_ok;
    prl_launch_scope ls(ctx, FAM_THRESHOLD);
    if (fast) {
        // Rows per CTA (rpc = 4).  Each S/Q row is fetched twice (as the bottom row of output row y-d, then as the
        // top row of y); the second fetch must hit L2, so the CTAs in flight have to cover a compact
        // set of rows: short tiles issued in raster order keep the live set at ~(in-flight rows + d)
        // rows of one or two pages (tens of MB), tall tiles thrash the 126 MB L2 (measured: 2x DRAM reads).
        const bool wide = g.d > 64;
        const int nt = wide ? 256 : 128;
        const int oc = (nt * 4 - g.d) & ~3;
        dim3 fg((g.out_cols + oc - 1) / oc, (g.out_rows + rpc - 1) / rpc, n_pages);
#define PRL_LAUNCH_FAST(M)                                                                            \
        do { if (P.compact) { if (wide) threshold_fast_kernel<M, 256, true><<<fg, 256, 0, ctx->stream>>>(A, F);     \
                              else threshold_fast_kernel<M, 128, true><<<fg, 128, 0, ctx->stream>>>(A, F); }          \
             else { if (wide) threshold_fast_kernel<M, 256, false><<<fg, 256, 0, ctx->stream>>>(A, F);               \
                    else threshold_fast_kernel<M, 128, false><<<fg, 128, 0, ctx->stream>>>(A, F); } } while (0)
        switch (method) {
        case PRL_SAUVOLA:    PRL_LAUNCH_FAST(PRL_SAUVOLA); break;
        case PRL_NIBLACK:    PRL_LAUNCH_FAST(PRL_NIBLACK); break;
        case PRL_WOLFJOLION: PRL_LAUNCH_FAST(PRL_WOLFJOLION); break;
        case PRL_NICK:       PRL_LAUNCH_FAST(PRL_NICK); break;
        default:             PRL_LAUNCH_FAST(PRL_FENG); break;
        }
#undef PRL_LAUNCH_FAST
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
