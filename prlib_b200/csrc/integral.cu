// integral.cu -- kernel 1: fused replicate-pad + row/column inclusive scan of a u8 page into two
// exact int64 integral planes S (sum) and Q (sum of squares).
//
// Replaces cv::copyMakeBorder(BORDER_REPLICATE) + cv::integral(CV_64F) + the Rect(1,1,..) crop of
// the reference (binarizeSauvola.cpp:65-77 and the identical blocks in Niblack/WolfJolion/NICK/
// Feng), plus cv::minMaxLoc(image) of binarizeWolfJolion.cpp:115-116 / binarizeFeng.cpp:111-112
// (page minimum, fused: the kernel sees every pixel anyway).
//
// Decomposition (HBM-bound: 1 byte read, 16 bytes written per padded pixel):
//   * one CTA owns one page (or one row band of it) over its FULL padded width; warp k owns the
//     256 padded columns [256k, 256k+256) as 4 sub-segments of 64 columns, lane l holding the
//     column pair (64j + 2l, 64j + 2l + 1) -> every S/Q store is one 16-byte word per lane and
//     512 contiguous bytes per warp instruction (full 128-B lines, no partial sectors);
//   * row direction: lane-local pair sums -> 5-step __shfl_up warp scan per sub-segment ->
//     sub-segment carries by shuffle broadcast -> cross-warp row offsets through a double-
//     buffered shared-memory table, ONE __syncthreads per chunk of R rows (row prefixes fit u32:
//     255^2 * 65536 < 2^32);
//   * column direction: each lane keeps the running int64 column sums of its 8 columns x 2
//     planes in registers while the CTA walks down the rows -- the row prefix never touches
//     memory, each S/Q element is written exactly once;
//   * the next chunk's pixels are loaded (register prefetch) before the current chunk's scan.
// Latency mode (few pages): the page is cut into row bands; two small kernels produce each
// band's top carry (column sums of the bands above, row-scanned) so the bands run concurrently.
#include "common.cuh"

namespace {

constexpr int kWarpCols = 256;   // padded columns per warp
constexpr int kSub = 4;          // sub-segments of 64 columns

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ void st_v2_s64(int64_t* p, int64_t a, int64_t b)
{
    asm volatile("st.global.v2.s64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}

template <int MAXW, int R>
__global__ void __launch_bounds__(MAXW * 32)
integral_scan_kernel(const uint8_t* __restrict__ src, size_t src_step, size_t src_page_stride, int rows, int cols,
                     int pad, int64_t* __restrict__ S, int64_t* __restrict__ Q, size_t pitch, size_t page_stride,
                     int rows_per_band, const int64_t* __restrict__ carry, uint32_t* __restrict__ imin)
{
    __shared__ uint2 tot[2][R][MAXW];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int page = blockIdx.y, band = blockIdx.x, bands = gridDim.x;
    const int Hp = rows + 2 * pad, Wp = cols + 2 * pad;
    const int Y0 = band * rows_per_band;
    const int Y1 = min(Y0 + rows_per_band, Hp);

    src += (size_t)page * src_page_stride;
    S += (size_t)page * page_stride;
    Q += (size_t)page * page_stride;

    const int Xb = wid * kWarpCols + 2 * lane;
    int xs[kSub][2];
    bool ok[kSub][2];
#pragma unroll
    for (int j = 0; j < kSub; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            int X = Xb + 64 * j + e;
            ok[j][e] = X < Wp;
            xs[j][e] = min(max(X - pad, 0), cols - 1);
        }

    long long accS[kSub][2], accQ[kSub][2];
#pragma unroll
    for (int j = 0; j < kSub; ++j) {
        accS[j][0] = accS[j][1] = accQ[j][0] = accQ[j][1] = 0;
        if (carry != nullptr && band > 0 && ok[j][0]) {
            const int64_t* c = carry + ((size_t)page * bands + band) * 2 * pitch + Xb + 64 * j;
            longlong2 cs = *reinterpret_cast<const longlong2*>(c);
            longlong2 cq = *reinterpret_cast<const longlong2*>(c + pitch);
            accS[j][0] = cs.x; accS[j][1] = cs.y; accQ[j][0] = cq.x; accQ[j][1] = cq.y;
        }
    }

    uint32_t mn = 255u;
    uint32_t cur[R][kSub], nxt[R][kSub];

    auto load_chunk = [&](uint32_t (&px)[R][kSub], int Yc) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int Y = Yc + r;
            const uint8_t* rowp = src + (size_t)min(max(Y - pad, 0), rows - 1) * src_step;
            const bool rv = Y < Y1;
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
                uint32_t v0 = (rv && ok[j][0]) ? (uint32_t)__ldg(rowp + xs[j][0]) : 0u;
                uint32_t v1 = (rv && ok[j][1]) ? (uint32_t)__ldg(rowp + xs[j][1]) : 0u;
                px[r][j] = v0 | (v1 << 16);
            }
        }
    };

    load_chunk(cur, Y0);
    int buf = 0;
    for (int Yc = Y0; Yc < Y1; Yc += R, buf ^= 1) {
        // sweep 1: this warp's row totals for the chunk
#pragma unroll
        for (int r = 0; r < R; ++r) {
            uint32_t s = 0, q = 0;
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
                uint32_t v0 = cur[r][j] & 0xffffu, v1 = cur[r][j] >> 16;
                s += v0 + v1;
                q += v0 * v0 + v1 * v1;
                if (ok[j][0] && Yc + r < Y1) mn = min(mn, v0);
                if (ok[j][1] && Yc + r < Y1) mn = min(mn, v1);
            }
            s = __reduce_add_sync(0xffffffffu, s);
            q = __reduce_add_sync(0xffffffffu, q);
            if (lane == 0) tot[buf][r][wid] = make_uint2(s, q);
        }
        // prefetch the next chunk while this one is scanned
        if (Yc + R < Y1) load_chunk(nxt, Yc + R);
        __syncthreads();

        // sweep 2: scans, column accumulation, stores
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int Y = Yc + r;
            if (Y < Y1) {
                uint2 t = (lane < wid) ? tot[buf][r][lane] : make_uint2(0u, 0u);
                uint32_t off_s = __reduce_add_sync(0xffffffffu, t.x);
                uint32_t off_q = __reduce_add_sync(0xffffffffu, t.y);
                int64_t* Srow = S + (size_t)Y * pitch + Xb;
                int64_t* Qrow = Q + (size_t)Y * pitch + Xb;
#pragma unroll
                for (int j = 0; j < kSub; ++j) {
                    const uint32_t v0 = cur[r][j] & 0xffffu, v1 = cur[r][j] >> 16;
                    const uint32_t v1q = v1 * v1;
                    const uint32_t is = warp_incl_scan(v0 + v1, lane);
                    const uint32_t iq = warp_incl_scan(v0 * v0 + v1q, lane);
                    const uint32_t r1s = off_s + is, r1q = off_q + iq;
                    accS[j][0] += (long long)(r1s - v1);
                    accS[j][1] += (long long)r1s;
                    accQ[j][0] += (long long)(r1q - v1q);
                    accQ[j][1] += (long long)r1q;
                    if (ok[j][0]) {
                        st_v2_s64(Srow + 64 * j, accS[j][0], accS[j][1]);
                        st_v2_s64(Qrow + 64 * j, accQ[j][0], accQ[j][1]);
                    }
                    off_s += __shfl_sync(0xffffffffu, is, 31);
                    off_q += __shfl_sync(0xffffffffu, iq, 31);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < kSub; ++j) cur[r][j] = nxt[r][j];
    }

    if (imin != nullptr) {
        mn = __reduce_min_sync(0xffffffffu, mn);
        if (lane == 0) atomicMin(imin + page, mn);
    }
}

// ---- latency mode: per-band column sums, then row-scanned carries --------------------------
// colsum[page][band][plane][X] = sum over the band's padded rows of P[Y][X] (plane 0) / P^2 (plane 1)
__global__ void __launch_bounds__(256)
band_colsum_kernel(const uint8_t* __restrict__ src, size_t src_step, size_t src_page_stride, int rows, int cols,
                   int pad, int rows_per_band, unsigned long long* __restrict__ colsum, size_t pitch)
{
    const int page = blockIdx.y, band = blockIdx.x, bands = gridDim.x;
    const int Hp = rows + 2 * pad, Wp = cols + 2 * pad;
    const int Y0 = band * rows_per_band, Y1 = min(Y0 + rows_per_band, Hp);
    src += (size_t)page * src_page_stride;
    unsigned long long* out = colsum + ((size_t)page * bands + band) * 2 * pitch;
    for (int X = threadIdx.x; X < Wp; X += blockDim.x) {
        const int x = min(max(X - pad, 0), cols - 1);
        unsigned long long s = 0, q = 0;
        for (int Y = Y0; Y < Y1; ++Y) {
            unsigned int p = __ldg(src + (size_t)min(max(Y - pad, 0), rows - 1) * src_step + x);
            s += p; q += p * p;
        }
        out[X] = s;
        out[pitch + X] = q;
    }
}

// carry[page][band][plane][X] = sum_{b' < band} sum_{X' <= X} colsum[page][b'][plane][X']
//                             = S[Y0(band) - 1][X]   (the integral row just above the band)
__global__ void __launch_bounds__(1024)
band_carry_kernel(const unsigned long long* __restrict__ colsum, long long* __restrict__ carry, int Wp, size_t pitch)
{
    __shared__ unsigned long long wtot[2][32];
    const int page = blockIdx.y, band = blockIdx.x, bands = gridDim.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long* in = colsum + (size_t)page * bands * 2 * pitch;
    long long* out = carry + ((size_t)page * bands + band) * 2 * pitch;
    unsigned long long run_s = 0, run_q = 0;
    for (int base = 0; base < Wp; base += 1024) {
        const int X = base + threadIdx.x;
        unsigned long long vs = 0, vq = 0;
        if (X < Wp)
            for (int b = 0; b < band; ++b) {
                vs += in[(size_t)b * 2 * pitch + X];
                vq += in[(size_t)b * 2 * pitch + pitch + X];
            }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long ts = __shfl_up_sync(0xffffffffu, vs, d);
            unsigned long long tq = __shfl_up_sync(0xffffffffu, vq, d);
            if (lane >= d) { vs += ts; vq += tq; }
        }
        if (lane == 31) { wtot[0][wid] = vs; wtot[1][wid] = vq; }
        __syncthreads();
        unsigned long long ps = 0, pq = 0, all_s = 0, all_q = 0;
        for (int k = 0; k < 32; ++k) {
            unsigned long long a = wtot[0][k], b = wtot[1][k];
            if (k < wid) { ps += a; pq += b; }
            all_s += a; all_q += b;
        }
        if (X < Wp) {
            out[X] = (long long)(run_s + ps + vs);
            out[pitch + X] = (long long)(run_q + pq + vq);
        }
        run_s += all_s; run_q += all_q;
        __syncthreads();
    }
}

}  // namespace

// Number of row bands a page is cut into: 1 when the batch alone fills the machine.
static int choose_bands(const prl_cuda_ctx* ctx, int n_pages, int Hp)
{
    const int want = 2 * ctx->num_sms;
    if (n_pages >= want / 2 + want / 4) return 1;
    int bands = (want + n_pages - 1) / n_pages;
    int max_bands = Hp / 32; if (max_bands < 1) max_bands = 1;
    if (bands > max_bands) bands = max_bands;
    if (bands > 64) bands = 64;
    return bands < 1 ? 1 : bands;
}

int prl_k_integral(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                   size_t src_page_stride, int pad, int64_t* d_S, int64_t* d_Q, size_t pitch,
                   size_t plane_page_stride, uint32_t* d_imin)
{
    const int Hp = rows + 2 * pad, Wp = cols + 2 * pad;
    const int nwarps = (Wp + kWarpCols - 1) / kWarpCols;
    if (nwarps > 32) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "padded width > 8192 columns is not supported yet");
    if (n_pages > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "more than 65535 pages per launch");
    constexpr int R = 4;
    int bands = choose_bands(ctx, n_pages, Hp);
    int rpb = (Hp + bands - 1) / bands;
    rpb = (rpb + R - 1) / R * R;
    bands = (Hp + rpb - 1) / rpb;

    if (d_imin) PRL_CUDA_TRY(ctx, cudaMemsetAsync(d_imin, 0xff, sizeof(uint32_t) * n_pages, ctx->stream));

    const int64_t* d_carry = nullptr;
    if (bands > 1) {
        size_t need = (size_t)n_pages * bands * 2 * pitch * sizeof(int64_t);
        int rc = prl_ensure(ctx, &ctx->colsum, &ctx->colsum_bytes, need); if (rc) return rc;
        rc = prl_ensure(ctx, &ctx->carry, &ctx->carry_bytes, need); if (rc) return rc;
        {
            prl_launch_scope ls(ctx, FAM_BAND_CARRY);
            band_colsum_kernel<<<dim3(bands, n_pages), 256, 0, ctx->stream>>>(
                d_src, src_step, src_page_stride, rows, cols, pad, rpb, (unsigned long long*)ctx->colsum, pitch);
        }
        {
            prl_launch_scope ls(ctx, FAM_BAND_CARRY);
            band_carry_kernel<<<dim3(bands, n_pages), 1024, 0, ctx->stream>>>(
                (const unsigned long long*)ctx->colsum, (long long*)ctx->carry, Wp, pitch);
        }
        d_carry = (const int64_t*)ctx->carry;
    }
    {
        prl_launch_scope ls(ctx, FAM_INTEGRAL);
        dim3 grid(bands, n_pages);
        if (nwarps <= 16)
            integral_scan_kernel<16, R><<<grid, nwarps * 32, 0, ctx->stream>>>(
                d_src, src_step, src_page_stride, rows, cols, pad, d_S, d_Q, pitch, plane_page_stride, rpb, d_carry, d_imin);
        else
            integral_scan_kernel<32, R><<<grid, nwarps * 32, 0, ctx->stream>>>(
                d_src, src_step, src_page_stride, rows, cols, pad, d_S, d_Q, pitch, plane_page_stride, rpb, d_carry, d_imin);
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
