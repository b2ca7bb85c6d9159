// integral_sq.cu -- kernel 1 for the COMPACT plane layout (common.cuh: prl_planes): fused replicate-pad + row/column
// inclusive scan of a u8 page into ONE plane of uint2 {S mod 2^32, Q mod 2^32} per padded pixel, plus the high words at
// the anchor positions (every A-th padded row, every 4th column).
//
// Replaces the same reference lines as integral.cu (cv::copyMakeBorder + cv::integral(CV_64F) + the Rect(1,1,..) crop,
// binarizeSauvola.cpp:65-77 and the identical blocks of the other four binarizers; cv::minMaxLoc(image),
// binarizeWolfJolion.cpp:115-116).  The scan itself is exact 64-bit arithmetic: what is dropped is only where the high
// words are STORED (kernel 2 rebuilds any int64 tap it needs from an anchor, see threshold.cu:full_tap).
//
// Why a second kernel: with 8.5 instead of 16 bytes written per pixel the int64 kernel (integral.cu) stopped being
// bound by its store stream and turned out to execute ~70 thread-instructions per pixel (ncu: 72 % issue-active).  Here:
//   * a lane owns 8 adjacent padded columns (a warp 256): one 5-step shuffle scan per plane serves 8 pixels;
//   * the u8 rows are staged by TMA exactly like in integral.cu (272-byte x 8-row boxes, 3-stage mbarrier ring per warp);
//     sweep 1 undoes the residual byte shift / substitutes the replicated border ONCE and writes the 8 bytes back
//     aligned, so sweep 2 is a single 8-byte shared load per lane and row;
//   * lane-local prefixes come out of dp4a with the lane's exclusive base as the accumulator input (one instruction per
//     output value), column sums are 32-bit adds except for the two anchor columns of a lane (64-bit);
//   * one 32-byte store per 4 pixels (st.global.v4.b64 of {S,Q} pairs), one 16-byte anchor store per lane on anchor rows.
#include "common.cuh"
#include "tma.cuh"
#include <algorithm>

using namespace prl_tma;

namespace {

constexpr int kWC = 256;                 // padded columns per warp
constexpr int kBox = kWC + 16;           // box bytes per row (16-byte aligned origin + residual shift)
constexpr int kR = 8;                    // source rows per chunk
constexpr int kNS = 3;                   // TMA stages per warp
constexpr int kStage = kR * kBox;        // 2176 = 17 * 128

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ unsigned long long pack2(uint32_t lo, uint32_t hi)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void st_v4_b64(void* p, unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d)
{
    asm volatile("st.global.v4.b64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void st_v4_u32(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// The replicated border, for the one or two warps of a CTA whose strip touches it.  Which of a lane's 8 bytes lie left of
// source column 0 / right of column cols-1 never changes from row to row, so the byte masks are built once per kernel
// (edge_masks) and a row costs two multiplies and four LOP3 -- the first version rebuilt them byte by byte in every row,
// which made the two edge warps ~30 % slower than the other eight and everybody wait for them at the chunk barrier
// (ncu round 2: 29 % of the warp samples on that barrier).
// edge_counts packs nl = bytes [0, nl) that replicate column 0 and nr = bytes [8 - nr, 8) that replicate column cols - 1.
__device__ __forceinline__ uint32_t edge_counts(int x_first, int cols)
{
    const int nl = min(max(-x_first, 0), 8), nr = min(max(x_first + 8 - cols, 0), 8);
    return (uint32_t)(64 - 8 * nl) | ((uint32_t)(64 - 8 * nr) << 8);          // as shift counts; shr/shl.b64 clamp at 64 -> 0
}
__device__ __forceinline__ uint2 fix_edge(uint2 w, uint32_t counts, uint32_t lv, uint32_t rv)
{
    unsigned long long ml, mr;
    asm("shr.b64 %0, %1, %2;" : "=l"(ml) : "l"(~0ull), "r"(counts & 0xffu));
    asm("shl.b64 %0, %1, %2;" : "=l"(mr) : "l"(~0ull), "r"(counts >> 8));
    mr &= ~ml;
    const uint32_t lx = (uint32_t)ml, ly = (uint32_t)(ml >> 32), rx = (uint32_t)mr, ry = (uint32_t)(mr >> 32);
    const uint32_t l4 = lv * 0x01010101u, r4 = rv * 0x01010101u;
    w.x = (w.x & ~(lx | rx)) | (l4 & lx) | (r4 & rx);
    w.y = (w.y & ~(ly | ry)) | (l4 & ly) | (r4 & ry);
    return w;
}

template <int MAXW, int MINB, bool WITH_MIN>
__global__ void __launch_bounds__(MAXW * 32, MINB)
integral_sq_kernel(const __grid_constant__ CUtensorMap tmap, int rows, int cols, int pad, uint2* __restrict__ SQ, size_t pitch,
                   size_t page_stride, int rows_per_band, const int64_t* __restrict__ carry, size_t carry_pitch,
                   uint32_t* __restrict__ imin, uint2* __restrict__ ASQ, int ashift, size_t a_pitch, size_t a_page_stride)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint2 tot[2][kR][MAXW + 1];         // entry w: row total of warp w (entry MAXW unused; keeps the stride odd)
    __shared__ uint64_t bars[MAXW][kNS];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int page = blockIdx.y, band = blockIdx.x, bands = gridDim.x;
    const int y0 = band * rows_per_band;
    const int y1 = min(y0 + rows_per_band, rows);

    SQ += (size_t)page * page_stride;
    ASQ += (size_t)page * a_page_stride;
    const int amask = (1 << ashift) - 1;

    uint8_t* slot = smem_raw + (size_t)wid * (kNS * kStage);             // this warp's ring
    const int X0 = wid * kWC;
    const int Xl = X0 + 8 * lane;
    // Box origin (source byte column, multiple of 16).  Normally the aligned byte below the strip's first source column;
    // a strip lying entirely in the right replicate border is moved left so that it still contains the last image column.
    int xb = (X0 - pad) & ~15;
    const bool all_right = X0 - pad >= cols;                             // every column replicates column cols-1
    const bool all_left = X0 + kWC <= pad;                               // every column replicates column 0
    if (all_right) xb = (cols - 1) & ~15;
    if (all_left) xb = 0;
    const int sh = X0 - pad - xb;                                        // strip column 0 sits `sh` bytes into the box
    const bool edge = (X0 < pad) || (X0 + kWC > pad + cols);             // strip touches a replicated column
    const int lcol = -xb;                                                // box offset of source column 0 (valid when X0 < pad)
    const int rcol = cols - 1 - xb;                                      // box offset of source column cols-1
    const uint32_t em = edge_counts(Xl - pad, cols);

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kNS; ++s) mbar_init(&bars[wid][s], 1);
        mbar_fence_init();
    }
    __syncwarp();

    const bool st = Xl < (int)pitch;                                     // pitch % 16 == 0: a lane is in or out as a whole
    // running column sums: 32-bit, except the two anchor columns (0 and 4 of the lane), whose high words are stored
    uint32_t aS[8], aQ[8];
    unsigned long long a64S[2] = {0ull, 0ull}, a64Q[2] = {0ull, 0ull};
#pragma unroll
    for (int i = 0; i < 8; ++i) aS[i] = aQ[i] = 0u;
    if (carry != nullptr && band > 0 && st) {
        const long long* c = reinterpret_cast<const long long*>(carry) + ((size_t)page * bands + band) * 2 * carry_pitch + Xl;
#pragma unroll
        for (int i = 0; i < 8; ++i) { aS[i] = (uint32_t)c[i]; aQ[i] = (uint32_t)c[carry_pitch + i]; }
        a64S[0] = (unsigned long long)c[0]; a64S[1] = (unsigned long long)c[4];
        a64Q[0] = (unsigned long long)c[carry_pitch]; a64Q[1] = (unsigned long long)c[carry_pitch + 4];
    }

    const int n_chunks = (y1 - y0 + kR - 1) / kR;
    // (the descriptor address must be the kernel-parameter address itself: no lambda / local copy)
#define PRL_ISSUE_TMA(c_)                                                                               \
    do {                                                                                                \
        const int s_ = (c_) % kNS;                                                                      \
        mbar_expect_tx(&bars[wid][s_], kR * kBox);                                                      \
        tma_load_3d(slot + (size_t)s_ * kStage, &tmap, xb / 2, y0 + (c_) * kR, page, &bars[wid][s_]);   \
    } while (0)
    if (lane == 0)
        for (int c = 0; c < kNS - 1 && c < n_chunks; ++c) PRL_ISSUE_TMA(c);

    // (Measured and dropped, round 2: software-pipelining the two sweeps -- sweep 1 of chunk c + 1 before sweep 2 of chunk c,
    // hand-over through a split mbarrier, four tot buffers -- removes the block-wide barrier but ran 29 % SLOWER (5.2 vs
    // 4.03 ms per 256 A4 pages): the longer live ranges spill at 64 registers and the waits turn into mbarrier polling.)
    uint32_t mn4 = 0xffffffffu;
    int buf_sel = 0;
    for (int c = 0; c < n_chunks; ++c, buf_sel ^= 1) {
        __syncwarp();                                         // every lane is done with slot (c-1) % kNS
        if (lane == 0 && c + kNS - 1 < n_chunks) PRL_ISSUE_TMA(c + kNS - 1);
        const int yc = y0 + c * kR;
        mbar_wait(&bars[wid][c % kNS], (uint32_t)((c / kNS) & 1));
        uint8_t* buf = slot + (size_t)(c % kNS) * kStage;

        // ---- sweep 1: fetch the lane's 8 pixels of every row (shift / border undone), row totals, aligned write-back
#pragma unroll 2
        for (int r = 0; r < kR; ++r) {
            uint8_t* row = buf + r * kBox;
            uint2 w;
            if (all_right) {
                w.x = w.y = row[rcol] * 0x01010101u;
            } else if (all_left) {
                w.x = w.y = row[0] * 0x01010101u;
            } else {
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(row + ((sh + 8 * lane) & ~3));
                const uint32_t w0 = wp[0], w1 = wp[1];
                if (sh & 3) {
                    const uint32_t w2 = wp[2];
                    w.x = __funnelshift_r(w0, w1, 8 * (sh & 3));
                    w.y = __funnelshift_r(w1, w2, 8 * (sh & 3));
                } else {
                    w.x = w0; w.y = w1;
                }
                if (edge) {
                    const uint32_t lv = (X0 < pad) ? row[lcol] : 0u, rv = (rcol >= 0 && rcol < kBox) ? row[rcol] : 0u;
                    w = fix_edge(w, em, lv, rv);
                }
            }
            if (WITH_MIN && yc + r < y1) mn4 = __vminu4(__vminu4(mn4, w.x), w.y);
            uint32_t s = __dp4a(w.y, 0x01010101u, __dp4a(w.x, 0x01010101u, 0u));
            uint32_t q = __dp4a(w.y, w.y, __dp4a(w.x, w.x, 0u));
            s = __reduce_add_sync(0xffffffffu, s);
            q = __reduce_add_sync(0xffffffffu, q);
            if (lane == 0) tot[buf_sel][r][wid] = make_uint2(s, q);
            __syncwarp();                                     // all lanes have read this row before it is overwritten in place
            reinterpret_cast<uint2*>(row)[lane] = w;
        }
        __syncthreads();

        // ---- sweep 2: scans, column accumulation, stores
#pragma unroll 1
        for (int r = 0; r < kR; ++r) {
            const int y = yc + r;
            if (y >= y1) break;
            const uint2 w = reinterpret_cast<const uint2*>(buf + r * kBox)[lane];
            const uint2 t = (lane < wid) ? tot[buf_sel][r][lane] : make_uint2(0u, 0u);
            const uint32_t off_s = __reduce_add_sync(0xffffffffu, t.x);
            const uint32_t off_q = __reduce_add_sync(0xffffffffu, t.y);
            const uint32_t ts = __dp4a(w.y, 0x01010101u, __dp4a(w.x, 0x01010101u, 0u));
            const uint32_t tq = __dp4a(w.y, w.y, __dp4a(w.x, w.x, 0u));
            const uint32_t es = off_s + warp_incl_scan(ts, lane) - ts;       // row prefix left of this lane
            const uint32_t eq = off_q + warp_incl_scan(tq, lane) - tq;
            // inclusive row prefixes of the lane's 8 columns: the base rides in as the dp4a accumulator
            uint32_t rs[8], rq[8];
            rs[0] = __dp4a(w.x, 0x00000001u, es); rs[1] = __dp4a(w.x, 0x00000101u, es);
            rs[2] = __dp4a(w.x, 0x00010101u, es); rs[3] = __dp4a(w.x, 0x01010101u, es);
            rs[4] = __dp4a(w.y, 0x00000001u, rs[3]); rs[5] = __dp4a(w.y, 0x00000101u, rs[3]);
            rs[6] = __dp4a(w.y, 0x00010101u, rs[3]); rs[7] = __dp4a(w.y, 0x01010101u, rs[3]);
            rq[0] = __dp4a(w.x, w.x & 0x000000ffu, eq); rq[1] = __dp4a(w.x, w.x & 0x0000ffffu, eq);
            rq[2] = __dp4a(w.x, w.x & 0x00ffffffu, eq); rq[3] = __dp4a(w.x, w.x, eq);
            rq[4] = __dp4a(w.y, w.y & 0x000000ffu, rq[3]); rq[5] = __dp4a(w.y, w.y & 0x0000ffffu, rq[3]);
            rq[6] = __dp4a(w.y, w.y & 0x00ffffffu, rq[3]); rq[7] = __dp4a(w.y, w.y, rq[3]);
            // source row y -> padded rows: row 0 also feeds the `pad` rows above it, row rows-1 the rows below
            int Y = y + pad, rep = 1;
            if (y == 0) { Y = 0; rep += pad; }
            if (y == rows - 1) rep += pad;
            for (int k = 0; k < rep; ++k, ++Y) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { aS[i] += rs[i]; aQ[i] += rq[i]; }
                a64S[0] += rs[0]; a64S[1] += rs[4]; a64Q[0] += rq[0]; a64Q[1] += rq[4];
                if (st) {
                    uint2* o = SQ + (size_t)Y * pitch + Xl;
                    st_v4_b64(o, pack2(aS[0], aQ[0]), pack2(aS[1], aQ[1]), pack2(aS[2], aQ[2]), pack2(aS[3], aQ[3]));
                    st_v4_b64(o + 4, pack2(aS[4], aQ[4]), pack2(aS[5], aQ[5]), pack2(aS[6], aQ[6]), pack2(aS[7], aQ[7]));
                    if ((Y & amask) == 0)
                        st_v4_u32(ASQ + (size_t)(Y >> ashift) * a_pitch + (Xl >> 2), (uint32_t)(a64S[0] >> 32), (uint32_t)(a64Q[0] >> 32),
                                  (uint32_t)(a64S[1] >> 32), (uint32_t)(a64Q[1] >> 32));
                }
            }
        }
    }

    if (WITH_MIN) {
        uint32_t m = min(min(mn4 & 0xff, (mn4 >> 8) & 0xff), min((mn4 >> 16) & 0xff, mn4 >> 24));
        m = __reduce_min_sync(0xffffffffu, m);
        if (lane == 0 && y1 > y0) atomicMin(imin + page, m);
    }
#undef PRL_ISSUE_TMA
}

template <int MAXW, int MINB>
int launch_sq(prl_cuda_ctx* ctx, dim3 grid, int nwarps, const CUtensorMap& tmap, int rows, int cols, int pad, const prl_planes& P,
              int rpb, const int64_t* d_carry, uint32_t* d_imin)
{
    const size_t smem = (size_t)nwarps * kNS * kStage;
    if (d_imin) {
        auto kfn = integral_sq_kernel<MAXW, MINB, true>;
        PRL_CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kfn<<<grid, nwarps * 32, smem, ctx->stream>>>(tmap, rows, cols, pad, (uint2*)P.S, P.pitch, P.page_stride, rpb, d_carry, P.pitch,
                                                      d_imin, (uint2*)P.AS, P.ashift, P.a_pitch, P.a_page_stride);
    } else {
        auto kfn = integral_sq_kernel<MAXW, MINB, false>;
        PRL_CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kfn<<<grid, nwarps * 32, smem, ctx->stream>>>(tmap, rows, cols, pad, (uint2*)P.S, P.pitch, P.page_stride, rpb, d_carry, P.pitch,
                                                      d_imin, (uint2*)P.AS, P.ashift, P.a_pitch, P.a_page_stride);
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

}  // namespace

int prl_k_integral_sq(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                      size_t src_page_stride, int pad, const prl_planes& P, uint32_t* d_imin)
{
    const int Wp = cols + 2 * pad;
    if (n_pages > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "more than 65535 pages per launch");
    if (!P.compact || !prl_integral_compact_ok(ctx, d_src, src_step, src_page_stride, rows, cols, pad) ||
        (((uintptr_t)P.S) & 31) != 0 || (((uintptr_t)P.AS) & 15) != 0 || (P.pitch & 15) != 0 || (P.page_stride & 3) != 0 ||
        (P.a_page_stride & 1) != 0 || P.a_pitch * 4 != P.pitch)
        return prl_set_err(ctx, PRL_E_INVALID, "compact planes: unaligned buffers or a page the compact kernel does not take");

    const int nwarps = (Wp + kWC - 1) / kWC;
    const int ctas_per_sm = nwarps <= 10 ? 3 : (nwarps <= 16 ? 2 : 1);
    int bands = prl_choose_bands(ctx, n_pages, rows, ctas_per_sm);
    int rpb = (rows + bands - 1) / bands;
    rpb = (rpb + kR - 1) / kR * kR;
    bands = (rows + rpb - 1) / rpb;

    if (d_imin) PRL_CUDA_TRY(ctx, cudaMemsetAsync(d_imin, 0xff, sizeof(uint32_t) * n_pages, ctx->stream));
    const int64_t* d_carry = nullptr;
    if (bands > 1) {
        int rc = prl_band_carries(ctx, d_src, n_pages, rows, cols, src_step, src_page_stride, pad, bands, rpb, P.pitch, &d_carry);
        if (rc) return rc;
    }

    CUtensorMap tmap;
    // rows described as src_step/2 u16 elements: a 272-byte box row is 136 elements (<= 256 allowed)
    if (!encode_3d(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, d_src, src_step / 2, (uint64_t)rows, (uint64_t)n_pages, src_step,
                   n_pages > 1 ? src_page_stride : src_step * rows, kBox / 2, kR))
        return prl_set_err(ctx, PRL_E_UNSUPPORTED, "tensor map rejected for the page batch");

    prl_launch_scope ls(ctx, FAM_INTEGRAL);
    dim3 grid(bands, n_pages);
    if (nwarps <= 10) return launch_sq<10, 3>(ctx, grid, nwarps, tmap, rows, cols, pad, P, rpb, d_carry, d_imin);
    if (nwarps <= 16) return launch_sq<16, 2>(ctx, grid, nwarps, tmap, rows, cols, pad, P, rpb, d_carry, d_imin);
    return launch_sq<32, 1>(ctx, grid, nwarps, tmap, rows, cols, pad, P, rpb, d_carry, d_imin);
}
