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
#include "tma.cuh"
#include <algorithm>

namespace {

struct ThrArgs {
    const uint8_t* src; size_t src_step, src_page_stride;
    const int64_t* S; const int64_t* Q; size_t pitch, plane_page_stride;   // compact layout: S is the uint2 {S lo, Q lo} plane, Q unused
    const uint2* AH; size_t a_pitch, a_page_stride; int ashift;            // compact layout: high words at the anchors (prl_planes)
    uint8_t* dst; size_t dst_step, dst_page_stride;
    const uint32_t* imin; long long* smax;
    int out_rows, out_cols, d;
    double kw, nkw, p0, p1, p2;
    // indirect launch of the exact kernel (hand-back of the fused path): plane slot z holds page page_map[slot_base + z]
    const int* page_map = nullptr; const int* page_count = nullptr; int slot_base = 0;
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
    int page = blockIdx.z;
    const int slot = blockIdx.z;
    if (A.page_map != nullptr) {
        if (A.slot_base + slot >= *A.page_count) return;
        page = A.page_map[A.slot_base + slot];
    }
    const int x = (blockIdx.x * 256 + threadIdx.x) * 2;
    const int64_t* S = A.S + (size_t)slot * A.plane_page_stride;
    const int64_t* Q = A.Q + (size_t)slot * A.plane_page_stride;
    const uint8_t* src = A.src + (size_t)page * A.src_page_stride;

    double imin = 0.0, coeff = 0.0;
    if (METHOD == PRL_WOLFJOLION || METHOD == PRL_FENG) imin = (double)A.imin[page];
    if (METHOD == PRL_WOLFJOLION && MODE != 2)
        coeff = __ddiv_rn(A.p0, __longlong_as_double(A.smax[page]));   // coeff = k / devianceMax

    double smax_local = __longlong_as_double(0xfff0000000000000LL);   // -inf
    // (direct launches have one CTA per block of kTR rows; the indirect hand-back launches a short grid that strides, so
    // that a launch over an empty page list costs a few thousand CTAs, not a million)
    if (x < A.out_cols)
    for (int y0 = blockIdx.y * kTR; y0 < A.out_rows; y0 += gridDim.y * kTR) {
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
// low words of S and of Q at 4 adjacent plane elements starting at element offset e (e % 4 == 0), either layout
template <bool COMPACT>
__device__ __forceinline__ void lo4(const ThrArgs& A, size_t e, uint4& s, uint4& q)
{
    long long a, b, c, d;
    if (COMPACT) {                                   // one 32-byte load: 4 x {S lo, Q lo}
        ldg256(A.S + e, a, b, c, d);
        s = make_uint4((unsigned int)a, (unsigned int)b, (unsigned int)c, (unsigned int)d);
        q = make_uint4((unsigned int)(a >> 32), (unsigned int)(b >> 32), (unsigned int)(c >> 32), (unsigned int)(d >> 32));
    } else {
        ldg256(A.S + e, a, b, c, d);
        s = make_uint4((unsigned int)a, (unsigned int)b, (unsigned int)c, (unsigned int)d);
        ldg256(A.Q + e, a, b, c, d);
        q = make_uint4((unsigned int)a, (unsigned int)b, (unsigned int)c, (unsigned int)d);
    }
}

// compact layout: the full int64 S and Q at (Y, X) of one page, rebuilt from the anchor that serves it (common.cuh: prl_planes)
__device__ __forceinline__ void full_taps(const uint2* __restrict__ L, const uint2* __restrict__ H, size_t pitch, size_t a_pitch,
                                          int ashift, int Y, int X, long long& s, long long& q)
{
    const int ya = Y >> ashift, xa = X & ~3;
    const uint2 lo_a = __ldg(L + ((size_t)ya << ashift) * pitch + xa);
    const uint2 hi = __ldg(H + (size_t)ya * a_pitch + (xa >> 2));
    const uint2 lo = __ldg(L + (size_t)Y * pitch + X);
    s = (long long)((((unsigned long long)hi.x << 32) | lo_a.x) + (unsigned long long)(unsigned int)(lo.x - lo_a.x));
    q = (long long)((((unsigned long long)hi.y << 32) | lo_a.y) + (unsigned long long)(unsigned int)(lo.y - lo_a.y));
}

// the literal reference arithmetic for output pixel (y, x) from the compact planes of one page
template <int METHOD>
__device__ __noinline__ int exact_t8_compact(const uint2* __restrict__ L, const uint2* __restrict__ H, size_t pitch, size_t a_pitch,
                                             int ashift, int y, int x, int d, double kw, double p0, double p1, double p2,
                                             double imin, double coeff)
{
    long long sa, sb, sc, sd, qa, qb, qc, qd;
    full_taps(L, H, pitch, a_pitch, ashift, y, x, sa, qa);
    full_taps(L, H, pitch, a_pitch, ashift, y, x + d, sb, qb);
    full_taps(L, H, pitch, a_pitch, ashift, y + d, x, sc, qc);
    full_taps(L, H, pitch, a_pitch, ashift, y + d, x + d, sd, qd);
    return exact_t8_from_taps<METHOD>(sa, sb, sc, sd, qa, qb, qc, qd, kw, p0, p1, p2, imin, coeff);
}

template <int METHOD, int NT>
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
    const bool has_out = (4 * threadIdx.x + 3 + A.d < FC) && (4 * (int)threadIdx.x < oc) && x < A.out_cols;
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
                    if (!fast_decide<METHOD, METHOD == PRL_SAUVOLA>(sw[i], qw[i], p, F, iminf, coefff, mu, o)) {
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

// ------------------------------------------------------------------------------------------------
// streaming kernel (the default mask path).  Same data path and the same decision rule as threshold_fast_kernel above
// -- exact integer window sums, FP32 estimate T~, decided unless within the proven margin mu, else the reference's FP64
// formula -- restructured for instruction count, which is what bounded that kernel once the planes went compact
// (ncu: 65 % issue-active, 76 thread-instructions per pixel, 45 % of them in the branchy two-tier decision):
//   * persistent CTAs walk the tiles in raster order (grid = resident CTAs), so the per-tile prologue is a handful of
//     integer updates and the second fetch of every plane row still finds it in L2;
//   * the FP32 estimate is evaluated for all four pixels of a thread unconditionally and branch-free (no variance-free
//     pre-test, no early-outs, hence no divergence); only pixels it cannot settle leave the straight line;
//   * u8 -> float and S_win -> float go through the 2^23 magic-number trick on the ALU pipe instead of I2F on the XU
//     pipe (exact for values < 2^23: S_win <= 255 (w-1)^2 needs w <= 181; wider windows convert with I2F).
// The estimate is the expression fast_decide evaluates (same operations, same order), so fast_margins covers it as is.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float u8_to_float(unsigned int p4, int i)
{
    // byte i of p4 under the exponent of 2^23: 0x4B0000pp = 8388608 + pp exactly
    return __uint_as_float(__byte_perm(p4, 0x4B000000u, 0x7540u | (unsigned int)i)) - 8388608.0f;
}

template <int METHOD, int NT, bool COMPACT>
__global__ void __launch_bounds__(NT, 1024 / NT)
threshold_stream_kernel(const ThrArgs A, const FastArgs F, int tiles_x, int tiles_y, int n_pages)
{
    constexpr int FC = NT * 4;                        // columns loaded per CTA row
    __shared__ __align__(16) unsigned int sD[2][kFR][2][FC];   // [buffer][row][plane S/Q][column]

    const int oc = (FC - A.d) & ~3;                  // output columns per tile
    const int tid = threadIdx.x;
    const bool col_out = (4 * tid + 3 + A.d < FC) && (4 * tid < oc);
    const bool small_s = F.w2 <= 181u * 181u;        // S_win < 2^23: magic-number conversion is exact

    // tile = (page, ty, tx), x fastest; this CTA takes tiles blockIdx.x, blockIdx.x + gridDim.x, ...
    int tx = blockIdx.x % tiles_x, t1 = blockIdx.x / tiles_x;
    int ty = t1 % tiles_y, page = t1 / tiles_y;
    const int sx = gridDim.x % tiles_x, s1 = gridDim.x / tiles_x;
    const int sy = s1 % tiles_y, sp = s1 / tiles_y;

    int buf = 0, cur_page = -1;
    double imin = 0.0, coeff = 0.0;
    float iminf = 0.f, coefff = 0.f, mu = F.mu0;
    for (; page < n_pages;) {
        if ((METHOD == PRL_WOLFJOLION || METHOD == PRL_FENG) && page != cur_page) {
            cur_page = page;
            imin = (double)A.imin[page]; iminf = (float)imin;
            if (METHOD == PRL_WOLFJOLION) {
                coeff = __ddiv_rn(A.p0, __longlong_as_double(A.smax[page]));
                coefff = (float)coeff;
                mu = F.mu0 + F.mu1 * fabsf(coefff);
                if (!(mu < 0.25f)) mu = 1e30f;       // degenerate s_max: every pixel takes the exact path
            }
        }
        const int x = tx * oc + 4 * tid;             // this thread's first column (16/32-byte aligned in the planes)
        const int y_begin = ty * F.rows_per_cta;
        const int y_end = min(y_begin + F.rows_per_cta, A.out_rows);
        const bool in_plane = x < (int)A.pitch;
        const bool has_out = col_out && x < A.out_cols;
        const bool full4 = x + 3 < A.out_cols;
        const size_t pbase = (size_t)page * A.plane_page_stride + x;
        const uint8_t* srow = A.src + (size_t)page * A.src_page_stride + (size_t)y_begin * A.src_step + x;
        uint8_t* orow = A.dst + (size_t)page * A.dst_page_stride + (size_t)y_begin * A.dst_step + x;

        for (int y = y_begin; y < y_end; y += kFR, buf ^= 1) {
            // ---- vertical differences of the low words -> shared memory (own copy stays in registers)
            unsigned int dsr[kFR][4], dqr[kFR][4], pix[kFR];
            {
                uint4 ta[kFR], tb[kFR], ua[kFR], ub[kFR];
#pragma unroll
                for (int r = 0; r < kFR; ++r) {
                    // the page pixels travel with the plane loads: fetched in the decision phase they were its longest stall
                    pix[r] = 0u;
                    if (has_out && y + r < y_end) {
                        const uint8_t* prow = srow + (size_t)(y + r - y_begin) * A.src_step;
                        if (full4) pix[r] = __ldg(reinterpret_cast<const unsigned int*>(prow));
                        else for (int i = 0; i < 4; ++i) if (x + i < A.out_cols) pix[r] |= (unsigned int)prow[i] << (8 * i);
                    }
                }
#pragma unroll
                for (int r = 0; r < kFR; ++r) {
                    ta[r] = tb[r] = ua[r] = ub[r] = make_uint4(0u, 0u, 0u, 0u);
                    if (in_plane && y + r < y_end) {
                        const size_t e = pbase + (size_t)(y + r) * A.pitch, eb = e + (size_t)A.d * A.pitch;
                        lo4<COMPACT>(A, e, ta[r], ua[r]);
                        lo4<COMPACT>(A, eb, tb[r], ub[r]);
                    }
                }
#pragma unroll
                for (int r = 0; r < kFR; ++r) {
                    dsr[r][0] = tb[r].x - ta[r].x; dsr[r][1] = tb[r].y - ta[r].y; dsr[r][2] = tb[r].z - ta[r].z; dsr[r][3] = tb[r].w - ta[r].w;
                    dqr[r][0] = ub[r].x - ua[r].x; dqr[r][1] = ub[r].y - ua[r].y; dqr[r][2] = ub[r].z - ua[r].z; dqr[r][3] = ub[r].w - ua[r].w;
                    *reinterpret_cast<uint4*>(&sD[buf][r][0][4 * tid]) = make_uint4(dsr[r][0], dsr[r][1], dsr[r][2], dsr[r][3]);
                    *reinterpret_cast<uint4*>(&sD[buf][r][1][4 * tid]) = make_uint4(dqr[r][0], dqr[r][1], dqr[r][2], dqr[r][3]);
                }
            }
            __syncthreads();
            if (has_out) {
#pragma unroll
                for (int r = 0; r < kFR; ++r) {
                    const int yy = y + r;
                    if (yy >= y_end) break;
                    const unsigned int* ls = &sD[buf][r][0][4 * tid + A.d];          // d even -> 8-byte aligned
                    const unsigned int* lq = &sD[buf][r][1][4 * tid + A.d];
                    const uint2 s_r0 = *reinterpret_cast<const uint2*>(ls), s_r1 = *reinterpret_cast<const uint2*>(ls + 2);
                    const uint2 q_r0 = *reinterpret_cast<const uint2*>(lq), q_r1 = *reinterpret_cast<const uint2*>(lq + 2);
                    const unsigned int sw[4] = {s_r0.x - dsr[r][0], s_r0.y - dsr[r][1], s_r1.x - dsr[r][2], s_r1.y - dsr[r][3]};
                    const unsigned int qw[4] = {q_r0.x - dqr[r][0], q_r0.y - dqr[r][1], q_r1.x - dqr[r][2], q_r1.y - dqr[r][3]};
                    const unsigned int p4 = pix[r];
                    unsigned int o4 = 0, und = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float pm = u8_to_float(p4, i) - 0.5f;
                        const float mf = small_s ? __uint_as_float(0x4B000000u + sw[i]) - 8388608.0f : (float)sw[i];
                        const float m = mf * F.kwf;
                        const float fn = n_to_float(F, sw[i], qw[i]);          // N = w^2 Q - S^2, exact and >= 0 before the conversion
                        const float sd = (fn * rsqrtf(fn)) * F.inv_w2f;        // fn == 0 gives NaN -> undecided (or the all-zero window below)
                        float T;
                        if (METHOD == PRL_SAUVOLA) T = m * fmaf(sd, F.c1, F.c2);
                        else if (METHOD == PRL_NIBLACK) T = fmaf(F.c0, sd, m);
                        else if (METHOD == PRL_WOLFJOLION) T = fmaf(fmaf(sd, coefff, -F.c0), m - iminf, m);
                        else if (METHOD == PRL_NICK) T = fmaf(F.c0, sqrtf(fmaf(m, m, sd * sd)), m);
                        else T = fmaf(F.c1, m, fmaf(F.c2, iminf, -iminf));
                        const float g = pm - fmaxf(T, 0.0f);
                        const bool ok = fn >= F.n_floor;
                        const bool white = ok && g > mu;
                        const bool black = (ok && g < -mu) || qw[i] == 0u;     // all-zero window => p == 0 => (0 > T8) is false
                        if (white) o4 |= 0xffu << (8 * i);
                        if (!(white || black)) und |= 1u << i;
                    }
                    if (und) {                                                  // ~1e-3 of the pixels: the reference arithmetic, literally
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (((und >> i) & 1u) && x + i < A.out_cols) {
                                const unsigned int p = (p4 >> (8 * i)) & 0xffu;
                                int t8;
                                if (COMPACT) {
                                    t8 = exact_t8_compact<METHOD>(reinterpret_cast<const uint2*>(A.S) + (size_t)page * A.plane_page_stride,
                                                                  A.AH + (size_t)page * A.a_page_stride, A.pitch, A.a_pitch, A.ashift, yy, x + i,
                                                                  A.d, A.kw, A.p0, A.p1, A.p2, imin, coeff);
                                } else {
                                    const size_t e0 = pbase + (size_t)yy * A.pitch + i;
                                    t8 = exact_t8_at<METHOD>(reinterpret_cast<const long long*>(A.S) + e0, reinterpret_cast<const long long*>(A.Q) + e0,
                                                             (size_t)A.d * A.pitch, A.d, A.kw, A.p0, A.p1, A.p2, imin, coeff);
                                }
                                if ((int)p > t8) o4 |= 0xffu << (8 * i);
                            }
                        }
                    }
                    // dst may be dense (pitch == out_cols, odd): store with whatever alignment the row has
                    uint8_t* op = orow + (size_t)(yy - y_begin) * A.dst_step;
                    const unsigned int al = (unsigned int)(uintptr_t)op & 3u;
                    if (!full4) {
                        for (int i = 0; i < 4; ++i) if (x + i < A.out_cols) op[i] = (uint8_t)(o4 >> (8 * i));
                    } else if (al == 0) {
                        *reinterpret_cast<unsigned int*>(op) = o4;
                    } else if (al == 2) {
                        *reinterpret_cast<unsigned short*>(op) = (unsigned short)o4;
                        *reinterpret_cast<unsigned short*>(op + 2) = (unsigned short)(o4 >> 16);
                    } else {
                        op[0] = (uint8_t)o4;
                        *reinterpret_cast<unsigned short*>(op + 1) = (unsigned short)(o4 >> 8);
                        op[3] = (uint8_t)(o4 >> 24);
                    }
                }
            }
        }
        // next tile of this CTA
        tx += sx; ty += sy; page += sp;
        if (tx >= tiles_x) { tx -= tiles_x; ++ty; }
        if (ty >= tiles_y) { ty -= tiles_y; ++page; }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA kernel (the default mask path on compact planes).  The streaming kernel above is latency-bound: its loads live in
// registers, a CTA alternates "issue 8 loads - wait - decide", and 8 CTAs per SM keep only ~60 KB in flight against a
// loaded-DRAM latency of ~2 us (measured: the same 7.1 ms for every method, 43 % of DRAM peak, long_scoreboard on top).
// Here a PRODUCER warp pulls the plane rows and the page pixels of one work unit after the other into a shared-memory
// ring with TMA (cp.async.bulk.tensor, mbarrier completion) and eight CONSUMER warps decide the pixels; full / empty
// mbarriers per stage, no block-wide barrier, no register held across the DRAM latency.
//   unit   = R rows x FC columns (4 x 512, or 2 x 1024 for tap distances > 64).  Per unit and side (top taps = plane rows
//            y.., bottom taps = rows y + d..) FC / 256 boxes of 256 pixels x R rows of uint2 {S lo, Q lo}, plus one box
//            of page pixels: 34 KB per stage.
//   order  = raster (x fastest), handed out through a global counter: CTAs that run ahead or behind still work on
//            neighbouring units, so the rows a unit needs as top taps were pulled through L2 as bottom taps d rows
//            earlier.  (Static striding let the CTAs drift apart and cost 35 % extra DRAM reads.)
//   thread = two PAIRS of adjacent pixels 64 columns apart in two rows of the unit: every shared-memory access of a
//            warp is 32 consecutive 16-byte words (conflict-free LDS.128).
// Decision rule and arithmetic are those of the streaming kernel (fast_margins covers both).
// ------------------------------------------------------------------------------------------------
constexpr int kTmaConsumers = 256;
constexpr int kTmaThreads = kTmaConsumers + 32;
constexpr int kTmaPix = 2048;                                     // pixels per unit = R * FC
constexpr int kTmaStageBytes = 2 * kTmaPix * 8 + kTmaPix;         // two sides of uint2 + the page pixels = 34816 (272 * 128)

__device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned int lds16(uint32_t addr)
{
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(prl_tma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float rsq_approx(float x)
{
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));     // MUFU.RSQ, 2 ulp (what the margin assumes); x == 0 -> +inf
    return r;
}

// N32: the variance numerator N = w^2 Q - S^2 fits 32 bits (w <= 15); SMALLS: S_win < 2^23 (w <= 181, magic-number int -> float)
template <int METHOD, bool WIDE, int NS, bool N32>
__global__ void __launch_bounds__(kTmaThreads, (NS == 2) ? 3 : 2)
threshold_tma_kernel(const __grid_constant__ CUtensorMap tmSQ, const __grid_constant__ CUtensorMap tmP, const ThrArgs A,
                     const FastArgs F, int tiles_x, int yblocks, int n_units, int* __restrict__ counter)
{
    constexpr int FC = WIDE ? 1024 : 512;
    constexpr int R = kTmaPix / FC;                               // rows per unit (4 / 2)
    constexpr int NBOX = FC / 256;                                // plane boxes per side
    constexpr int RPP = kTmaConsumers / (FC / 4);                 // rows the consumers cover in one pass (2 / 1); two passes per unit
    extern __shared__ __align__(128) uint8_t ring[];
    __shared__ uint64_t full[NS], empty[NS];
    __shared__ int4 info[NS];                                     // {page, yb, tx, -} of the unit in each stage; page < 0: no more units

    const int tid = threadIdx.x;
    const int oc = (FC - A.d) & ~15;                              // output columns per unit (16-byte aligned pixel boxes)

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) { prl_tma::mbar_init(&full[s], 1); prl_tma::mbar_init(&empty[s], kTmaConsumers / 32); }
        prl_tma::mbar_fence_init();
    }
    __syncthreads();

    if (tid >= kTmaConsumers) {
        // ---------------- producer: one lane ----------------
        if (tid == kTmaConsumers) {
            for (int k = 0;; ++k) {
                const int stage = k % NS;
                if (k >= NS) prl_tma::mbar_wait(&empty[stage], (uint32_t)(((k / NS) - 1) & 1));
                const int u = atomicAdd(counter, 1);
                if (u >= n_units) {
                    info[stage] = make_int4(-1, 0, 0, 0);
                    mbar_arrive(&full[stage]);
                    break;
                }
                const int tx = u % tiles_x, t1 = u / tiles_x;
                const int yb = t1 % yblocks, page = t1 / yblocks;
                info[stage] = make_int4(page, yb, tx, 0);
                uint8_t* dst = ring + (size_t)stage * kTmaStageBytes;
                prl_tma::mbar_expect_tx(&full[stage], kTmaStageBytes);
                const int x0 = tx * oc, y0 = yb * R;
#pragma unroll
                for (int b = 0; b < NBOX; ++b) {
                    prl_tma::tma_load_3d(dst + (size_t)b * (R * 2048), &tmSQ, x0 + 256 * b, y0, page, &full[stage]);
                    prl_tma::tma_load_3d(dst + kTmaPix * 8 + (size_t)b * (R * 2048), &tmSQ, x0 + 256 * b, y0 + A.d, page, &full[stage]);
                }
                prl_tma::tma_load_3d(dst + 2 * kTmaPix * 8, &tmP, x0 / (WIDE ? 4 : 2), y0, page, &full[stage]);
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const bool small_s = F.w2 <= 181u * 181u;
    const int rs = tid / (FC / 4);                                // row slot of the pass
    const int j = tid % (FC / 4);
    const int c0 = (j >> 5) * 128 + 2 * (j & 31);                 // first pixel of pair 0; pair 1 starts 64 columns further
    // loop-invariant per pair: shared-memory offsets (own columns, columns + d, pixels) and the output columns left of it
    uint32_t off_a[2], off_b[2];
    int cmax[2];                                                  // the pair produces output iff x0 < cmax (x0 = first column of the unit)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = c0 + 64 * h, cd = c + A.d;
        off_a[h] = (uint32_t)((c >> 8) * (R << 11) + ((c & 255) << 3));
        off_b[h] = (uint32_t)((cd >> 8) * (R << 11) + ((cd & 255) << 3));
        cmax[h] = (c + 1 + A.d < FC && c < oc) ? A.out_cols - c : -0x40000000;
    }
    const uint32_t ring_u32 = prl_tma::smem_u32(ring), full_u32 = prl_tma::smem_u32(full), empty_u32 = prl_tma::smem_u32(empty);
    const uint32_t info_u32 = prl_tma::smem_u32(info);
    const unsigned int dst_step = (unsigned int)A.dst_step;
    // method constants with 1/w^2 folded in where the formula allows: sd always appears as (fn * rsqrt(fn)) * inv_w2
    const float kwf = F.kwf, inv_w2f = F.inv_w2f, n_floor = F.n_floor;
    const float c0f = F.c0, c1f = F.c1, c2f = F.c2;
    const unsigned int w2 = F.w2;
    // Sauvola and Niblack with 1/w^2 folded into the coefficients: T = S (r C1 + C2) and T = r C0 + S/w^2 with r = N rsqrt(N) --
    // two multiplies per pixel fewer, and fewer roundings than the expression fast_margins budgets for (decide.cuh)
    const float fc1 = (float)(F.kw_d * F.kw_d * F.c1_d), fc2 = (float)(F.kw_d * F.c2_d), fc0 = (float)(F.kw_d * F.c0_d);
    // a pair of pixels leaves as one 16-bit store when every row of dst starts on an even address (x0 and c are even)
    const bool even_dst = ((((uintptr_t)A.dst) | (uintptr_t)A.dst_step | (uintptr_t)A.dst_page_stride) & 1u) == 0;

    int cur_page = -1;
    double imin = 0.0, coeff = 0.0;
    float iminf = 0.f, coefff = 0.f, mu = F.mu0;
    for (int k = 0;; ++k) {
        const int stage = k % NS;
        {   // wait for the stage (suspending try_wait: the hardware parks the warp instead of spinning)
            const uint32_t bar = full_u32 + 8u * (uint32_t)stage, parity = (uint32_t)((k / NS) & 1);
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "WAITF_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n\t"
                "@p bra DONEF_%=;\n\t"
                "bra WAITF_%=;\n\t"
                "DONEF_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
        }
        int4 ui;
        asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(ui.x), "=r"(ui.y), "=r"(ui.z), "=r"(ui.w) : "r"(info_u32 + 16u * (uint32_t)stage));
        const int page = ui.x;
        if (page < 0) break;
        if ((METHOD == PRL_WOLFJOLION || METHOD == PRL_FENG) && page != cur_page) {
            cur_page = page;
            imin = (double)A.imin[page]; iminf = (float)imin;
            if (METHOD == PRL_WOLFJOLION) {
                coeff = __ddiv_rn(A.p0, __longlong_as_double(A.smax[page]));
                coefff = (float)coeff;
                mu = F.mu0 + F.mu1 * fabsf(coefff);
                if (!(mu < 0.25f)) mu = 1e30f;       // degenerate s_max: every pixel takes the exact path
            }
        }
        const float g_white = 0.5f + mu, g_black = 0.5f - mu;     // the -0.5 of (p - 0.5) - T rides in the comparison
        const uint32_t st = ring_u32 + (uint32_t)stage * kTmaStageBytes;
        const int x0 = ui.z * oc, yu = ui.y * R;
        uint8_t* const obase = A.dst + (size_t)(unsigned int)page * A.dst_page_stride + (size_t)((unsigned int)yu * dst_step + (unsigned int)x0);
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const int rr = rs + pass * RPP;                       // row of the unit
            const int y = yu + rr;
            if (y < A.out_rows) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (x0 < cmax[h]) {
                        const int c = c0 + 64 * h;
                        // taps: a = top[c], b = top[c + d], c = bottom[c], d = bottom[c + d]; each load = 2 pixels x {S lo, Q lo}
                        const uint32_t ra = st + (uint32_t)(rr << 11) + off_a[h], rb = st + (uint32_t)(rr << 11) + off_b[h];
                        const uint4 ta = lds128(ra), tb = lds128(rb), ba = lds128(ra + kTmaPix * 8), bb = lds128(rb + kTmaPix * 8);
                        const unsigned int p2 = lds16(st + 2 * kTmaPix * 8 + (uint32_t)(rr * FC + c));
                        const unsigned int sw[2] = {(bb.x - ba.x) - (tb.x - ta.x), (bb.z - ba.z) - (tb.z - ta.z)};
                        const unsigned int qw[2] = {(bb.y - ba.y) - (tb.y - ta.y), (bb.w - ba.w) - (tb.w - ta.w)};
                        unsigned int o2 = 0, und = 0;
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const float pf = u8_to_float(p2, i);
                            const float mf = small_s ? __uint_as_float(0x4B000000u + sw[i]) - 8388608.0f : (float)sw[i];
                            float fn;                                              // N = w^2 Q - S^2, exact and >= 0 before the conversion
                            if (N32) fn = (float)(w2 * qw[i] - sw[i] * sw[i]);
                            else fn = n_to_float(F, sw[i], qw[i]);
                            const float r = fn * rsq_approx(fn);                   // fn == 0 gives NaN -> undecided (or the all-zero window below)
                            float T;
                            if (METHOD == PRL_SAUVOLA) T = mf * fmaf(r, fc1, fc2);
                            else if (METHOD == PRL_NIBLACK) T = fmaf(r, fc0, mf * kwf);
                            else {
                                const float m = mf * kwf, sd = r * inv_w2f;
                                if (METHOD == PRL_WOLFJOLION) T = fmaf(fmaf(sd, coefff, -c0f), m - iminf, m);
                                else if (METHOD == PRL_NICK) T = fmaf(c0f, sqrtf(fmaf(m, m, sd * sd)), m);
                                else T = fmaf(c1f, m, fmaf(c2f, iminf, -iminf));
                            }
                            const float g = pf - fmaxf(T, 0.0f);                   // (p - 0.5) - T  =  g - 0.5
                            const bool ok = fn >= n_floor;
                            const bool white = ok && g > g_white;
                            const bool black = (ok && g < g_black) || qw[i] == 0u; // all-zero window => p == 0 => (0 > T8) is false
                            o2 |= white ? (0xffu << (8 * i)) : 0u;
                            und |= (white || black) ? 0u : (1u << i);
                        }
                        const int x = x0 + c;
                        if (und && !F.dbg_skip_exact) {                             // ~1e-3 of the pixels: the reference arithmetic, literally
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                if (((und >> i) & 1u) && x + i < A.out_cols) {
                                    const unsigned int p = (p2 >> (8 * i)) & 0xffu;
                                    const int t8 = exact_t8_compact<METHOD>(reinterpret_cast<const uint2*>(A.S) + (size_t)page * A.plane_page_stride,
                                                                            A.AH + (size_t)page * A.a_page_stride, A.pitch, A.a_pitch, A.ashift, y, x + i,
                                                                            A.d, A.kw, A.p0, A.p1, A.p2, imin, coeff);
                                    if ((int)p > t8) o2 |= 0xffu << (8 * i);
                                }
                            }
                        }
                        // dst may be dense (pitch == out_cols, odd): a pair goes out as one 16-bit store where its address allows
                        uint8_t* op = obase + ((unsigned int)rr * dst_step + (unsigned int)c);
                        if (x0 + 1 < cmax[h] && even_dst) {
                            *reinterpret_cast<unsigned short*>(op) = (unsigned short)o2;
                        } else {
                            op[0] = (uint8_t)o2;
                            if (x0 + 1 < cmax[h]) op[1] = (uint8_t)(o2 >> 8);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if ((tid & 31) == 0)                                      // this warp is done with the stage
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(empty_u32 + 8u * (uint32_t)stage) : "memory");
    }
}

template <int METHOD>
int launch_tma_threshold(prl_cuda_ctx* ctx, const ThrArgs& A, const FastArgs& F, const prl_geom& g, const prl_planes& P,
                         const uint8_t* d_src, size_t src_step, size_t src_page_stride, int n_pages, bool* launched)
{
    *launched = false;
    const bool wide = g.d > 64;
    const int FC = wide ? 1024 : 512, R = kTmaPix / FC;
    const int oc = (FC - g.d) & ~15;
    if (oc <= 0 || A.dst_step >= ((size_t)1 << 32)) return PRL_OK;
    const int tiles_x = (g.out_cols + oc - 1) / oc, yblocks = (g.out_rows + R - 1) / R;
    const long long n_units = (long long)tiles_x * yblocks * n_pages;
    if (n_units >= (1ll << 30)) return PRL_OK;
    CUtensorMap tmSQ, tmP;
    // planes: {pitch, Hp, pages} of 8-byte elements, boxes of 256 elements x R rows
    if (!prl_tma::encode_3d(&tmSQ, CU_TENSOR_MAP_DATA_TYPE_UINT64, P.S, P.pitch, (uint64_t)g.Hp, (uint64_t)n_pages, P.pitch * 8,
                            P.page_stride * 8, 256, R))
        return PRL_OK;
    // page pixels: rows of FC bytes described as 256 elements of 2 (4) bytes
    const int eb = wide ? 4 : 2;
    if (!prl_tma::encode_3d(&tmP, wide ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, d_src, src_step / eb,
                            (uint64_t)g.rows, (uint64_t)n_pages, src_step, n_pages > 1 ? src_page_stride : src_step * g.rows, 256, R))
        return PRL_OK;
    int rc = prl_ensure(ctx, &ctx->sched, &ctx->sched_bytes, 256); if (rc) return rc;
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(ctx->sched, 0, sizeof(int), ctx->stream));
    const int ns = ctx->thr_stages == 3 ? 3 : 2;
    const size_t smem = (size_t)ns * kTmaStageBytes;
    const int grid = (int)std::min<long long>(n_units, (long long)ctx->num_sms * (ns == 2 ? 3 : 2));
#define PRL_LAUNCH_TMA(W, N, B)                                                                                       \
    do { auto kfn = threshold_tma_kernel<METHOD, W, N, B>;                                                            \
         PRL_CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
         kfn<<<grid, kTmaThreads, smem, ctx->stream>>>(tmSQ, tmP, A, F, tiles_x, yblocks, (int)n_units, (int*)ctx->sched); } while (0)
    if (wide) { if (ns == 2) PRL_LAUNCH_TMA(true, 2, false); else PRL_LAUNCH_TMA(true, 3, false); }   // d > 64: N never fits 32 bits
    else if (F.n32) { if (ns == 2) PRL_LAUNCH_TMA(false, 2, true); else PRL_LAUNCH_TMA(false, 3, true); }
    else { if (ns == 2) PRL_LAUNCH_TMA(false, 2, false); else PRL_LAUNCH_TMA(false, 3, false); }
#undef PRL_LAUNCH_TMA
    *launched = true;
    return PRL_OK;
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
    const size_t pbase = (size_t)page * A.plane_page_stride + x;
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
                const size_t e = pbase + (size_t)(y + r) * A.pitch, eb = e + (size_t)A.d * A.pitch;
                uint4 sa, sb, qa, qb;
                lo4<COMPACT>(A, e, sa, qa);
                lo4<COMPACT>(A, eb, sb, qb);
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
    const uint2* L = reinterpret_cast<const uint2*>(A.S) + (size_t)page * A.plane_page_stride;
    const uint2* H = COMPACT ? A.AH + (size_t)page * A.a_page_stride : nullptr;
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
                const uint2* l0 = L + (size_t)y * A.pitch + x;
                const uint2 ta = __ldg(l0), tb = __ldg(l0 + A.d), tc = __ldg(l0 + dr), td = __ldg(l0 + dr + A.d);
                const unsigned int swl = (td.x - tc.x) - (tb.x - ta.x), qwl = (td.y - tc.y) - (tb.y - ta.y);
                const unsigned long long Nl = (unsigned long long)F.w2 * qwl - (unsigned long long)swl * swl;
                if (Nl < lo) continue;
                full_taps(L, H, A.pitch, A.a_pitch, A.ashift, y, x, sa, qa);
                full_taps(L, H, A.pitch, A.a_pitch, A.ashift, y, x + A.d, sb, qb);
                full_taps(L, H, A.pitch, A.a_pitch, A.ashift, y + A.d, x, sc, qc);
                full_taps(L, H, A.pitch, A.a_pitch, A.ashift, y + A.d, x + A.d, sd, qd);
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

static ThrArgs make_thr_args(int method, const uint8_t* d_src, const prl_geom& g, size_t src_step, size_t src_page_stride,
                             const prl_planes& P, const double* params, const uint32_t* d_imin, long long* d_smax,
                             uint8_t* d_dst, size_t dst_step, size_t dst_page_stride)
{
    ThrArgs A;
    A.src = d_src; A.src_step = src_step; A.src_page_stride = src_page_stride;
    A.S = (const int64_t*)P.S; A.Q = (const int64_t*)P.Q; A.pitch = P.pitch; A.plane_page_stride = P.page_stride;
    A.AH = (const uint2*)P.AS; A.a_pitch = P.a_pitch; A.a_page_stride = P.a_page_stride; A.ashift = P.ashift;
    A.dst = d_dst; A.dst_step = dst_step; A.dst_page_stride = dst_page_stride;
    A.imin = d_imin; A.smax = d_smax;
    A.out_rows = g.out_rows; A.out_cols = g.out_cols; A.d = g.d;
    A.kw = 1.0 / (double)(g.w * g.w);     // wSqrBack, binarizeSauvola.cpp:58-59
    A.nkw = -A.kw;
    A.p0 = params[0]; A.p1 = 0; A.p2 = 0;
    if (method == PRL_SAUVOLA) { A.p1 = params[0] * (1.0 / 128.0); A.p2 = 1.0 - params[0]; }   // (k*RBack), (1-k) :115-117
    if (method == PRL_FENG) { A.p1 = 1.0 + (1.0 - params[0]); A.p2 = params[2]; }              // c2 + c1, k2
    return A;
}

// Masks of the pages page_map[slot_base .. slot_base + slots) (as many as *page_count says exist) from the int64 planes in
// slots 0 .. slots-1, every pixel through the literal FP64 path.  Serves the device-side hand-back of the fused path
// (never Wolf-Jolion: that method does not take the fused path).
int prl_k_threshold_exact_indirect(prl_cuda_ctx* ctx, int method, const uint8_t* d_src, int slots, const prl_geom& g, size_t src_step,
                                   size_t src_page_stride, const prl_planes& P, const double* params, const uint32_t* d_imin,
                                   uint8_t* d_dst, size_t dst_step, size_t dst_page_stride, const int* d_map, const int* d_count,
                                   int slot_base)
{
    if (P.compact || method == PRL_WOLFJOLION || method < PRL_SAUVOLA || method > PRL_FENG)
        return prl_set_err(ctx, PRL_E_INVALID, "indirect exact threshold: int64 planes, not Wolf-Jolion");
    ThrArgs A = make_thr_args(method, d_src, g, src_step, src_page_stride, P, params, d_imin, nullptr, d_dst, dst_step, dst_page_stride);
    A.page_map = d_map; A.page_count = d_count; A.slot_base = slot_base;
    dim3 grid((g.out_cols + 511) / 512, std::min((g.out_rows + kTR - 1) / kTR, 64), slots);
    if (slots > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    prl_launch_scope ls(ctx, FAM_THRESHOLD);
    switch (method) {
    case PRL_SAUVOLA: launch_exact<PRL_SAUVOLA>(ctx, 0, A, grid); break;
    case PRL_NIBLACK: launch_exact<PRL_NIBLACK>(ctx, 0, A, grid); break;
    case PRL_NICK:    launch_exact<PRL_NICK>(ctx, 0, A, grid); break;
    default:          launch_exact<PRL_FENG>(ctx, 0, A, grid); break;
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

int prl_k_threshold(prl_cuda_ctx* ctx, int method, int mode, const uint8_t* d_src, int n_pages,
                    const prl_geom& g, size_t src_step, size_t src_page_stride, const prl_planes& P, const double* params,
                    const uint32_t* d_imin, long long* d_smax, uint8_t* d_dst, size_t dst_step,
                    size_t dst_page_stride)
{
    const int64_t* d_S = (const int64_t*)P.S;
    const int64_t* d_Q = (const int64_t*)P.Q;
    const size_t plane_page_stride = P.page_stride;
    ThrArgs A = make_thr_args(method, d_src, g, src_step, src_page_stride, P, params, d_imin, d_smax, d_dst, dst_step, dst_page_stride);
    if (method < PRL_SAUVOLA || method > PRL_FENG) return prl_set_err(ctx, PRL_E_INVALID, "unknown method");

    dim3 grid((g.out_cols + 511) / 512, (g.out_rows + kTR - 1) / kTR, n_pages);
    if (n_pages > 65535 || grid.y > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");

    // fast path eligibility: mask output, even tap distance < 256, window sums < 2^32, 4/32-byte aligned buffers
    FastArgs F;
    const bool aligned = ((src_step | src_page_stride | (uintptr_t)d_src) & 3) == 0 &&
                         ((((uintptr_t)d_S) | ((uintptr_t)d_Q)) & 31) == 0 && (plane_page_stride & 3) == 0 && (P.pitch & 3) == 0;
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
    F.dbg_skip_exact = ctx->dbg_skip_exact ? 1 : 0;
    prl_launch_scope ls(ctx, FAM_THRESHOLD);
    if (fast && P.compact && !ctx->thr_no_tma && ((src_step | src_page_stride | (uintptr_t)d_src) & 15) == 0) {
        bool launched = false;
        int rc;
        switch (method) {
        case PRL_SAUVOLA:    rc = launch_tma_threshold<PRL_SAUVOLA>(ctx, A, F, g, P, d_src, src_step, src_page_stride, n_pages, &launched); break;
        case PRL_NIBLACK:    rc = launch_tma_threshold<PRL_NIBLACK>(ctx, A, F, g, P, d_src, src_step, src_page_stride, n_pages, &launched); break;
        case PRL_WOLFJOLION: rc = launch_tma_threshold<PRL_WOLFJOLION>(ctx, A, F, g, P, d_src, src_step, src_page_stride, n_pages, &launched); break;
        case PRL_NICK:       rc = launch_tma_threshold<PRL_NICK>(ctx, A, F, g, P, d_src, src_step, src_page_stride, n_pages, &launched); break;
        default:             rc = launch_tma_threshold<PRL_FENG>(ctx, A, F, g, P, d_src, src_step, src_page_stride, n_pages, &launched); break;
        }
        if (rc) return rc;
        if (launched) {
            PRL_CUDA_TRY(ctx, cudaGetLastError());
            return PRL_OK;
        }
    }
    if (fast) {
        // Rows per CTA (rpc = 4).  Each S/Q row is fetched twice (as the bottom row of output row y-d, then as the
        // top row of y); the second fetch must hit L2, so the CTAs in flight have to cover a compact
        // set of rows: short tiles issued in raster order keep the live set at ~(in-flight rows + d)
        // rows of one or two pages (tens of MB), tall tiles thrash the 126 MB L2 (measured: 2x DRAM reads).
        const bool wide = g.d > 64;
        const int nt = wide ? 256 : 128;
        const int oc = (nt * 4 - g.d) & ~3;
        dim3 fg((g.out_cols + oc - 1) / oc, (g.out_rows + rpc - 1) / rpc, n_pages);
        // streaming kernel: persistent CTAs (every resident slot of the device) over the same tiles in the same order
        const long long n_tiles = (long long)fg.x * fg.y * n_pages;
        const int slots = ctx->num_sms * (1024 / nt);
        const int pg = (int)std::min<long long>(n_tiles, (long long)slots);
        const bool stream_ok = (P.compact || !ctx->thr_legacy) && n_tiles < (1ll << 31);   // the round-1 kernel knows int64 planes only
#define PRL_LAUNCH_FAST(M)                                                                            \
        do { if (stream_ok) {                                                                         \
                 if (P.compact) { if (wide) threshold_stream_kernel<M, 256, true><<<pg, 256, 0, ctx->stream>>>(A, F, fg.x, fg.y, n_pages);    \
                                  else threshold_stream_kernel<M, 128, true><<<pg, 128, 0, ctx->stream>>>(A, F, fg.x, fg.y, n_pages); }        \
                 else { if (wide) threshold_stream_kernel<M, 256, false><<<pg, 256, 0, ctx->stream>>>(A, F, fg.x, fg.y, n_pages);             \
                        else threshold_stream_kernel<M, 128, false><<<pg, 128, 0, ctx->stream>>>(A, F, fg.x, fg.y, n_pages); }                 \
             } else { if (wide) threshold_fast_kernel<M, 256><<<fg, 256, 0, ctx->stream>>>(A, F);                    \
                      else threshold_fast_kernel<M, 128><<<fg, 128, 0, ctx->stream>>>(A, F); } } while (0)
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
