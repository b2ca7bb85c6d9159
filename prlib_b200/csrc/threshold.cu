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
// with NaN / out-of-int32 -> 0, mask = p > T8 ? 255 : 0.  All FP64, __d*_rn intrinsics so the
// compiler can never contract a multiply-add the CPU path does not contract.
//
// HBM-bound: per output pixel 16 B of S/Q are compulsory; each S/Q element is used as 4 different
// taps.  Horizontal reuse (taps x and x+d) is served by L1 within the CTA, vertical reuse (rows y
// and y+d) by L2: CTAs are issued row-band-major so the live working set is a band of ~d rows.
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

__device__ __forceinline__ int to_u8(double T)
{
    // cvRound == cvtsd2si: NaN and anything outside int32 become INT_MIN -> saturates to 0
    if (!(T > 0.0) || T >= 2147483647.5) return 0;
    if (T >= 255.5) return 255;
    return __double2int_rn(T);
}

__device__ __forceinline__ double tap4(double kw, double nkw, long long a, long long b, long long c, long long d)
{
    double r = __dmul_rn(kw, (double)a);
    r = __dadd_rn(r, __dmul_rn(nkw, (double)b));
    r = __dadd_rn(r, __dmul_rn(nkw, (double)c));
    r = __dadd_rn(r, __dmul_rn(kw, (double)d));
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

constexpr int kTR = 4;   // output rows per CTA

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
threshold_kernel(const ThrArgs A)
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

template <int METHOD>
int launch_method(prl_cuda_ctx* ctx, int mode, const ThrArgs& A, dim3 grid)
{
    switch (mode) {
    case 0: threshold_kernel<METHOD, 0><<<grid, 256, 0, ctx->stream>>>(A); break;
    case 1: threshold_kernel<METHOD, 1><<<grid, 256, 0, ctx->stream>>>(A); break;
    default: threshold_kernel<METHOD, 2><<<grid, 256, 0, ctx->stream>>>(A); break;
    }
    return 0;
}

__global__ void init_smax_kernel(long long* smax, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) smax[i] = (long long)0xfff0000000000000LL;   // -inf
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

    dim3 grid((g.out_cols + 511) / 512, (g.out_rows + kTR - 1) / kTR, n_pages);
    if (n_pages > 65535 || grid.y > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");

    if (method == PRL_WOLFJOLION) {
        {
            prl_launch_scope ls(ctx, FAM_SMAX);
            init_smax_kernel<<<(n_pages + 255) / 256, 256, 0, ctx->stream>>>(d_smax, n_pages);
        }
        prl_launch_scope ls(ctx, FAM_SMAX);
        launch_method<PRL_WOLFJOLION>(ctx, 2, A, grid);
    }
    {
        prl_launch_scope ls(ctx, FAM_THRESHOLD);
        switch (method) {
        case PRL_SAUVOLA:    launch_method<PRL_SAUVOLA>(ctx, mode, A, grid); break;
        case PRL_NIBLACK:    launch_method<PRL_NIBLACK>(ctx, mode, A, grid); break;
        case PRL_WOLFJOLION: launch_method<PRL_WOLFJOLION>(ctx, mode, A, grid); break;
        case PRL_NICK:       launch_method<PRL_NICK>(ctx, mode, A, grid); break;
        case PRL_FENG:       launch_method<PRL_FENG>(ctx, mode, A, grid); break;
        default: return prl_set_err(ctx, PRL_E_INVALID, "unknown method");
        }
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
