// fused.cu -- small-window path: the integral planes never reach HBM.
//
// Same results as kernel 1 + kernel 2 (integral.cu, threshold.cu) for Sauvola / Niblack / NICK / Feng with
// windows up to 31, at 1 byte read + 1 byte written per pixel instead of ~35:
//   * a page is cut into strips of 128 padded columns; ONE WARP owns a strip and walks down all rows, so
//     there is no block barrier and no cross-warp row offset.  Window sums are differences of four integral
//     taps, and differences do not care where the integral's origin is: the warp keeps a STRIP-LOCAL integral
//     (row prefix starting at the strip's first column), running column sums in registers, the last d+1 rows
//     of their low 32 bits in a shared-memory ring (S_win, Q_win < 2^32);
//   * per row: dp4a lane prefixes + two 5-step shuffle scans (as kernel 1), ring write, then the output row
//     d rows up: taps from the ring, exact integer window sums, FP32 decision (decide.cuh);
//   * the rare pixel the FP32 estimate cannot settle needs the reference's FP64 arithmetic on the TRUE int64
//     integral taps.  true = local + L, where L[Y] = S[Y][strip_start - 1] comes from a cheap pre-pass over
//     the u8 page (strip_rowsum_kernel + strip_colscan_kernel, which also yields the page minimum), and the
//     64-bit local Q is rebuilt from the lane's own 64-bit column sum plus low-word differences (each < 2^32).
// Strips overlap by d columns (a strip emits 128 - d outputs), i.e. ~12 % redundant scan work at w = 15.
#include "common.cuh"
#include "decide.cuh"

namespace {

constexpr int kFW = 128;          // scanned padded columns per warp strip

struct FusedArgs {
    const uint8_t* src; size_t src_step, src_page_stride;
    uint8_t* dst; size_t dst_step, dst_page_stride;
    const uint32_t* imin;            // per page
    const longlong2* L;              // [page][Hp][ns] true {S, Q} integral just left of each strip
    int rows, cols, pad, d, Hp, Wp, out_rows, out_cols, ow, ns, n_pages;
    double kw, p0, p1, p2;
};

// 4 bytes at `p` (any alignment)
__device__ __forceinline__ uint32_t ldg_u32_unaligned(const uint8_t* p)
{
    const uint32_t sh = (uint32_t)(uintptr_t)p & 3u;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(p - sh);
    uint32_t w = __ldg(a);
    if (sh) w = __funnelshift_r(w, __ldg(a + 1), 8 * sh);
    return w;
}

__device__ __forceinline__ uint32_t warp_scan_u32(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// ---- pre-pass 1: rowpre[page][y][b] = {sum, sum of squares} of the padded columns X < b*ow of source row y
__global__ void __launch_bounds__(256)
strip_rowsum_kernel(const uint8_t* __restrict__ src, size_t step, size_t page_stride, int rows, int cols, int pad,
                    int ow, int ns, uint2* __restrict__ rowpre, uint32_t* __restrict__ imin)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int page = blockIdx.y;
    const int y = blockIdx.x * 8 + wid;
    if (y >= rows) return;
    const int Wp = cols + 2 * pad;
    const uint8_t* row = src + (size_t)page * page_stride + (size_t)y * step;
    uint2* out = rowpre + ((size_t)page * rows + y) * ns;
    uint32_t run_s = 0, run_q = 0, mn = 255u;
    for (int b = 0; b < ns; ++b) {
        if (lane == 0) out[b] = make_uint2(run_s, run_q);
        const int X1 = min((b + 1) * ow, Wp);
        uint32_t s = 0, q = 0;
        for (int X = b * ow + lane; X < X1; X += 32) {
            const uint32_t p = __ldg(row + min(max(X - pad, 0), cols - 1));
            s += p; q += p * p; mn = min(mn, p);
        }
        run_s += __reduce_add_sync(0xffffffffu, s);
        run_q += __reduce_add_sync(0xffffffffu, q);
    }
    // columns beyond the last strip start never enter an L value, but they do enter the page minimum
    for (int X = ns * ow + lane; X < Wp; X += 32) mn = min(mn, (uint32_t)__ldg(row + min(max(X - pad, 0), cols - 1)));
    mn = __reduce_min_sync(0xffffffffu, mn);
    if (lane == 0 && imin) atomicMin(imin + page, mn);
}

// ---- pre-pass 2: L[page][Y][b] = sum over padded rows Y' <= Y of rowpre[page][clamp(Y' - pad)][b]
__global__ void __launch_bounds__(128)
strip_colscan_kernel(const uint2* __restrict__ rowpre, longlong2* __restrict__ L, int rows, int pad, int Hp, int ns,
                     int n_pages)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= ns * n_pages) return;
    const int page = idx / ns, b = idx - page * ns;
    long long as = 0, aq = 0;
    for (int Y = 0; Y < Hp; ++Y) {
        const int y = min(max(Y - pad, 0), rows - 1);
        const uint2 v = __ldg(rowpre + ((size_t)page * rows + y) * ns + b);
        as += v.x; aq += v.y;
        L[((size_t)page * Hp + Y) * ns + b] = make_longlong2(as, aq);
    }
}

// ---- the fused kernel: one warp = one strip of one page
template <int METHOD, int RR, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
local_fused_kernel(const FusedArgs A, const FastArgs F)
{
    extern __shared__ __align__(16) uint32_t fsm[];
    constexpr int WARP_WORDS = RR * kFW * 2 + 2 * RR * (kFW / 4) + 16;   // S ring, Q ring, pixel ring, Q-high-byte ring (+ slack)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * WARPS + wid;
    if (gw >= A.ns * A.n_pages) return;                 // warps are independent: no block-level barrier below
    const int page = gw / A.ns, strip = gw - page * A.ns;
    const int x0 = strip * A.ow;                        // first scanned padded column == first output column
    uint32_t* ringS = fsm + wid * WARP_WORDS;
    uint32_t* ringQ = ringS + RR * kFW;
    uint32_t* pixr = ringQ + RR * kFW;
    uint32_t* ringH = pixr + RR * (kFW / 4);           // bits 32..39 of the local Q sums, one byte per column
    const uint8_t* src = A.src + (size_t)page * A.src_page_stride;
    uint8_t* dst = A.dst + (size_t)page * A.dst_page_stride;
    const int pad = A.pad, d = A.d;

    double imin = 0.0;
    float iminf = 0.f;
    if (METHOD == PRL_FENG) { imin = (double)A.imin[page]; iminf = (float)imin; }
    const float mu = F.mu0;

    // lane's padded columns X = x0 + 4*lane + i  <->  source columns xs0 + i
    const int xs0 = x0 + 4 * lane - pad;
    const bool interior = (x0 >= pad) && (x0 + kFW - pad + 8 < A.cols);
    // raw fetch of one source row: two aligned words (interior) or four clamped bytes packed (edge strips);
    // the funnel shift is applied only when the row is consumed, PF rows later, so the loads stay in flight
    auto fetch = [&](int y, uint32_t& lo, uint32_t& hi) {
        const uint8_t* row = src + (size_t)y * A.src_step;
        if (interior) {
            const uint8_t* p = row + xs0;
            const uint32_t* a = reinterpret_cast<const uint32_t*>(p - ((uintptr_t)p & 3u));
            lo = __ldg(a); hi = __ldg(a + 1);
        } else {
            uint32_t w = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) w |= (uint32_t)__ldg(row + min(max(xs0 + i, 0), A.cols - 1)) << (8 * i);
            lo = w; hi = 0;
        }
    };
    auto combine = [&](int y, uint32_t lo, uint32_t hi) -> uint32_t {
        if (!interior) return lo;
        const uint32_t sh = (uint32_t)(uintptr_t)(src + (size_t)y * A.src_step + xs0) & 3u;
        return sh ? __funnelshift_r(lo, hi, 8 * sh) : lo;
    };

    // which of this lane's 4 outputs exist
    const int xo = x0 + 4 * lane;                       // output column of byte 0
    const bool lane_out = (4 * lane < A.ow) && (4 * lane + 3 + d < kFW) && xo < A.out_cols;
    const bool full4 = xo + 3 < A.out_cols;

    uint32_t accS[4] = {0, 0, 0, 0}, accQ[4] = {0, 0, 0, 0};
    uint32_t qhi = 0;                                   // 4 packed 8-bit carry counters: high words of the local Q sums
    int Y = 0;                                          // next padded row to emit
    constexpr int PF = 4;                               // rows of prefetch distance
    uint32_t plo[PF], phi[PF];
#pragma unroll
    for (int k = 0; k < PF; ++k) { plo[k] = phi[k] = 0; if (k < A.rows) fetch(k, plo[k], phi[k]); }
    for (int y = 0; y < A.rows; ++y) {
        const uint32_t w = combine(y, plo[0], phi[0]);
#pragma unroll
        for (int k = 0; k + 1 < PF; ++k) { plo[k] = plo[k + 1]; phi[k] = phi[k + 1]; }
        if (y + PF < A.rows) fetch(y + PF, plo[PF - 1], phi[PF - 1]);
        // strip-local row prefix of this source row (u32)
        const uint32_t a0 = w & 0xffu, a1 = __dp4a(w, 0x00000101u, 0u), a2 = __dp4a(w, 0x00010101u, 0u), a3 = __dp4a(w, 0x01010101u, 0u);
        const uint32_t b0 = a0 * a0, b1 = __dp4a(w, w & 0x0000ffffu, 0u), b2 = __dp4a(w, w & 0x00ffffffu, 0u), b3 = __dp4a(w, w, 0u);
        const uint32_t es = warp_scan_u32(a3, lane) - a3, eq = warp_scan_u32(b3, lane) - b3;
        const uint32_t rs[4] = {es + a0, es + a1, es + a2, es + a3};
        const uint32_t rq[4] = {eq + b0, eq + b1, eq + b2, eq + b3};
        int rep = 1;                                    // row 0 / row rows-1 also feed the replicated border rows
        if (y == 0) rep += pad;
        if (y == A.rows - 1) rep += pad;
        for (int k = 0; k < rep; ++k, ++Y) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                accS[i] += rs[i];
                const uint32_t old = accQ[i];
                accQ[i] += rq[i];
                qhi += (accQ[i] < old ? 1u : 0u) << (8 * i);
            }
            const int slot = Y & (RR - 1);
            __syncwarp();                               // every lane is done reading the slot this row overwrites
            *reinterpret_cast<uint4*>(ringS + slot * kFW + 4 * lane) = make_uint4(accS[0], accS[1], accS[2], accS[3]);
            *reinterpret_cast<uint4*>(ringQ + slot * kFW + 4 * lane) = make_uint4(accQ[0], accQ[1], accQ[2], accQ[3]);
            pixr[slot * (kFW / 4) + lane] = w;
            ringH[slot * (kFW / 4) + lane] = qhi;
            __syncwarp();
            const int yo = Y - d;                       // output row whose bottom taps are this row
            if (yo < 0 || yo >= A.out_rows || !lane_out) continue;
            const int top = yo & (RR - 1);
            const uint32_t* tS = ringS + top * kFW + 4 * lane;
            const uint32_t* tQ = ringQ + top * kFW + 4 * lane;
            const uint32_t* bS = ringS + slot * kFW + 4 * lane;
            const uint32_t* bQ = ringQ + slot * kFW + 4 * lane;
            const uint4 sa = *reinterpret_cast<const uint4*>(tS), qa = *reinterpret_cast<const uint4*>(tQ);
            const uint2 sb0 = *reinterpret_cast<const uint2*>(tS + d), sb1 = *reinterpret_cast<const uint2*>(tS + d + 2);
            const uint2 qb0 = *reinterpret_cast<const uint2*>(tQ + d), qb1 = *reinterpret_cast<const uint2*>(tQ + d + 2);
            const uint2 sd0 = *reinterpret_cast<const uint2*>(bS + d), sd1 = *reinterpret_cast<const uint2*>(bS + d + 2);
            const uint2 qd0 = *reinterpret_cast<const uint2*>(bQ + d), qd1 = *reinterpret_cast<const uint2*>(bQ + d + 2);
            const uint32_t sA[4] = {sa.x, sa.y, sa.z, sa.w}, qA[4] = {qa.x, qa.y, qa.z, qa.w};
            const uint32_t sB[4] = {sb0.x, sb0.y, sb1.x, sb1.y}, qB[4] = {qb0.x, qb0.y, qb1.x, qb1.y};
            const uint32_t sD[4] = {sd0.x, sd0.y, sd1.x, sd1.y}, qD[4] = {qd0.x, qd0.y, qd1.x, qd1.y};
            // pixels p(yo, xo..xo+3): padded row yo+pad, strip byte offset 4*lane + pad
            uint32_t p4;
            {
                const uint32_t* pr = pixr + ((yo + pad) & (RR - 1)) * (kFW / 4);
                const int bo = 4 * lane + pad;
                p4 = pr[bo >> 2];
                if (bo & 3) p4 = __funnelshift_r(p4, pr[(bo >> 2) + 1], 8 * (bo & 3));
            }
            uint32_t o4 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t sw = (sD[i] - accS[i]) - (sB[i] - sA[i]);
                const uint32_t qw = (qD[i] - accQ[i]) - (qB[i] - qA[i]);
                const uint32_t p = (p4 >> (8 * i)) & 0xffu;
                int o;
                if (!fast_decide<METHOD, true>(sw, qw, p, F, iminf, 0.f, mu, o)) {
                    o = 0;
                    if (xo + i < A.out_cols) {
                        // true int64 taps = strip-local + L.  The local S fits 32 bits; the local Q is 40 bits: low
                        // word from the Q ring, bits 32..39 from the high-byte ring.
                        const longlong2 Lt = __ldg(A.L + ((size_t)page * A.Hp + yo) * A.ns + strip);
                        const longlong2 Lb = __ldg(A.L + ((size_t)page * A.Hp + Y) * A.ns + strip);
                        const uint8_t* hT = reinterpret_cast<const uint8_t*>(ringH + top * (kFW / 4));
                        const uint8_t* hB = reinterpret_cast<const uint8_t*>(ringH + slot * (kFW / 4));
                        const int cl = 4 * lane + i, cr = cl + d;
                        const long long qaa = (long long)(((unsigned long long)hT[cl] << 32) | qA[i]);
                        const long long qbb = (long long)(((unsigned long long)hT[cr] << 32) | qB[i]);
                        const long long qc = (long long)(((unsigned long long)hB[cl] << 32) | accQ[i]);
                        const long long qdd = (long long)(((unsigned long long)hB[cr] << 32) | qD[i]);
                        const int t8 = exact_t8_from_taps<METHOD>((long long)sA[i] + Lt.x, (long long)sB[i] + Lt.x,
                                                                  (long long)accS[i] + Lb.x, (long long)sD[i] + Lb.x,
                                                                  qaa + Lt.y, qbb + Lt.y, qc + Lb.y, qdd + Lb.y,
                                                                  A.kw, A.p0, A.p1, A.p2, imin, 0.0);
                        o = (int)p > t8 ? 255 : 0;
                    }
                }
                o4 |= (uint32_t)o << (8 * i);
            }
            uint8_t* orow = dst + (size_t)yo * A.dst_step + xo;
            const unsigned int al = (unsigned int)(uintptr_t)orow & 3u;
            if (!full4) {
                for (int i = 0; i < 4; ++i) if (xo + i < A.out_cols) orow[i] = (uint8_t)(o4 >> (8 * i));
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

template <int METHOD, int RR, int WARPS>
int launch_fused_t(prl_cuda_ctx* ctx, const FusedArgs& A, const FastArgs& F)
{
    const size_t smem = (size_t)WARPS * (RR * kFW * 2 + 2 * RR * (kFW / 4) + 16) * sizeof(uint32_t);
    auto kfn = local_fused_kernel<METHOD, RR, WARPS>;
    static bool configured = false;
    if (!configured) {
        PRL_CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int warps = A.ns * A.n_pages;
    kfn<<<(warps + WARPS - 1) / WARPS, WARPS * 32, smem, ctx->stream>>>(A, F);
    return PRL_OK;
}

template <int METHOD>
int launch_fused_m(prl_cuda_ctx* ctx, const FusedArgs& A, const FastArgs& F)
{
    if (A.d + 1 <= 16) return launch_fused_t<METHOD, 16, 2>(ctx, A, F);
    return launch_fused_t<METHOD, 32, 2>(ctx, A, F);
}

}  // namespace

// Can this call take the fused path?  (mask output only; Wolf-Jolion needs s_max first -> planes path)
bool prl_fused_eligible(const prl_cuda_ctx* ctx, int method, int n_pages, const prl_geom& g, const double* params)
{
    if (!ctx->use_fused || ctx->force_exact) return false;   // opt-in: see DESIGN.md section 4 (F2)
    if (method == PRL_WOLFJOLION) return false;
    if ((g.d & 1) || g.d + 1 > 32 || g.d < 2) return false;
    if (g.Hp > 100000) return false;                       // 8-bit carry counters of the local Q high words
    const int ow = (kFW - g.d) & ~3;
    const int ns = (g.out_cols + ow - 1) / ow;
    (void)n_pages;
    FastArgs F;
    return fast_margins(method, params, g, &F);
}

int prl_k_fused(prl_cuda_ctx* ctx, int method, const uint8_t* d_src, int n_pages, const prl_geom& g, size_t src_step,
                size_t src_page_stride, const double* params, uint32_t* d_imin, uint8_t* d_dst, size_t dst_step,
                size_t dst_page_stride)
{
    FastArgs F;
    if (!fast_margins(method, params, g, &F)) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "fused path not eligible");
    FusedArgs A;
    A.src = d_src; A.src_step = src_step; A.src_page_stride = src_page_stride;
    A.dst = d_dst; A.dst_step = dst_step; A.dst_page_stride = dst_page_stride;
    A.rows = g.rows; A.cols = g.cols; A.pad = g.h; A.d = g.d; A.Hp = g.Hp; A.Wp = g.Wp;
    A.out_rows = g.out_rows; A.out_cols = g.out_cols;
    A.ow = (kFW - g.d) & ~3;
    A.ns = (g.out_cols + A.ow - 1) / A.ow;
    A.n_pages = n_pages;
    A.kw = 1.0 / (double)(g.w * g.w);
    A.p0 = params[0]; A.p1 = 0; A.p2 = 0;
    if (method == PRL_SAUVOLA) { A.p1 = params[0] * (1.0 / 128.0); A.p2 = 1.0 - params[0]; }
    if (method == PRL_FENG) { A.p1 = 1.0 + (1.0 - params[0]); A.p2 = params[2]; }
    A.imin = d_imin;

    // scratch: rowpre (u32 pairs) in ctx->colsum, L (int64 pairs) in ctx->carry
    const size_t rowpre_bytes = (size_t)n_pages * g.rows * A.ns * sizeof(uint2);
    const size_t L_bytes = (size_t)n_pages * g.Hp * A.ns * sizeof(longlong2);
    int rc = prl_ensure(ctx, &ctx->colsum, &ctx->colsum_bytes, rowpre_bytes); if (rc) return rc;
    rc = prl_ensure(ctx, &ctx->carry, &ctx->carry_bytes, L_bytes); if (rc) return rc;
    A.L = (const longlong2*)ctx->carry;
    if (n_pages > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "more than 65535 pages per launch");
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(d_imin, 0xff, sizeof(uint32_t) * n_pages, ctx->stream));
    {
        prl_launch_scope ls(ctx, FAM_FUSED_PRE);
        strip_rowsum_kernel<<<dim3((g.rows + 7) / 8, n_pages), 256, 0, ctx->stream>>>(
            d_src, src_step, src_page_stride, g.rows, g.cols, g.h, A.ow, A.ns, (uint2*)ctx->colsum, d_imin);
    }
    {
        prl_launch_scope ls(ctx, FAM_FUSED_PRE);
        const int total = A.ns * n_pages;
        strip_colscan_kernel<<<(total + 127) / 128, 128, 0, ctx->stream>>>((const uint2*)ctx->colsum, (longlong2*)ctx->carry,
                                                                         g.rows, g.h, g.Hp, A.ns, n_pages);
    }
    {
        prl_launch_scope ls(ctx, FAM_FUSED);
        switch (method) {
        case PRL_SAUVOLA: rc = launch_fused_m<PRL_SAUVOLA>(ctx, A, F); break;
        case PRL_NIBLACK: rc = launch_fused_m<PRL_NIBLACK>(ctx, A, F); break;
        case PRL_NICK:    rc = launch_fused_m<PRL_NICK>(ctx, A, F); break;
        case PRL_FENG:    rc = launch_fused_m<PRL_FENG>(ctx, A, F); break;
        default: return prl_set_err(ctx, PRL_E_INVALID, "method not served by the fused path");
        }
        if (rc) return rc;
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
