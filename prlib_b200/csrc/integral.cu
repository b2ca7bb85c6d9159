// integral.cu -- kernel 1: fused replicate-pad + row/column inclusive scan of a u8 page into two
// exact int64 integral planes S (sum) and Q (sum of squares).
//
// Replaces cv::copyMakeBorder(BORDER_REPLICATE) + cv::integral(CV_64F) + the Rect(1,1,..) crop of
// the reference (binarizeSauvola.cpp:65-77 and the identical blocks in Niblack/WolfJolion/NICK/
// Feng), plus cv::minMaxLoc(image) of binarizeWolfJolion.cpp:115-116 / binarizeFeng.cpp:111-112
// (page minimum, fused: the kernel sees every pixel anyway).
//
// HBM-bound: 1 byte read, 16 bytes written per padded pixel.  Decomposition:
//   * one CTA owns one page (or one band of source rows) over its FULL padded width; warp k owns
//     the 256 padded columns [256k, 256k+256) as 2 sub-segments of 128 columns, lane l holding the
//     4 adjacent columns 128j + 4l .. +3 -> every S/Q store is one 32-byte STG.256 per lane, 1 KB
//     contiguous per warp instruction (full lines, no partial sectors);
//   * the u8 rows are staged by TMA (cp.async.bulk.tensor, one 272-byte x R box per warp and chunk
//     of R source rows, 3-stage mbarrier ring per warp, no registers held across the latency).  TMA
//     needs a 16-byte aligned box origin, so the box starts at the aligned byte below the strip's
//     first padded column (272 = 256 + 16 bytes, described as 136 u16 elements because a box
//     dimension is limited to 256 elements) and lanes undo the residual shift with one funnel shift
//     of two aligned shared-memory words; TMA zero-fills outside the image and the two edge warps
//     substitute the replicated border pixel;
//   * row direction: lane-local prefix of 4 pixels -> 5-step __shfl_up warp scan per sub-segment ->
//     cross-warp row offsets through a double-buffered shared table, ONE __syncthreads per chunk
//     (row prefixes fit u32: 255^2 * 65536 < 2^32);
//   * column direction: each lane keeps the running int64 column sums of its 8 columns x 2 planes
//     in registers while the CTA walks down the rows; the replicated top/bottom border rows are
//     emitted by re-accumulating the same row prefix, so each S/Q element is written exactly once
//     and the row prefix never touches memory.
// Latency mode (few pages): the page is cut into bands of source rows; two small kernels produce
// each band's top carry (weighted column sums of the rows above, row-scanned) so bands run
// concurrently.  integral_generic_kernel is the any-alignment / any-pitch fallback (no TMA).
#include "common.cuh"
#include "tma.cuh"
#include <algorithm>

namespace {

constexpr int kWarpCols = 256;   // padded columns per warp

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
__device__ __forceinline__ void st_v4_s64(int64_t* p, long long a, long long b, long long c, long long d)
{
    asm volatile("st.global.v4.s64 [%0], {%1, %2, %3, %4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

using namespace prl_tma;

// ------------------------------------------------------------------------------------------------
// TMA kernel.  Band b covers source rows [b*rows_per_band, ...); it emits padded row y+pad for each
// of its source rows, plus the `pad` replicated rows above row 0 / below row rows-1.
// J = sub-segments of 128 columns per warp (1: 128 columns/warp, 16 accumulator registers/plane-pair
// fewer -> twice the resident warps for pages up to 4096 columns; 2: 256 columns/warp for wide pages).
// ------------------------------------------------------------------------------------------------
template <int J> __host__ __device__ constexpr int box_bytes() { return 128 * J + 16; }   // 16-byte aligned origin
template <int J, int R> __host__ __device__ constexpr int stage_bytes() { return (R * box_bytes<J>() + 127) / 128 * 128; }

template <int MAXW, int MINB, int J, int R, int NS, bool MULTI>
__global__ void __launch_bounds__(MAXW * 32, MINB)
integral_tma_kernel(const __grid_constant__ CUtensorMap tmap, int rows, int cols, int pad, int64_t* __restrict__ S,
                    int64_t* __restrict__ Q, size_t pitch, size_t page_stride, int rows_per_band,
                    const int64_t* __restrict__ carry, uint32_t* __restrict__ imin, int col0,
                    uint2* __restrict__ rowoff, int has_in, int has_out)
{
    // Wide pages are covered by several launches ("column passes") of at most MAXW strips each: pass p starts at
    // padded column col0 and takes, per source row, the row prefix accumulated by the passes to its left from
    // rowoff[page][y] (has_in), and leaves its own running total there for the next pass (has_out).
    constexpr int WC = 128 * J;                  // padded columns per warp
    constexpr int BOX = box_bytes<J>();
    constexpr int STAGE = stage_bytes<J, R>();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint2 tot[2][R][MAXW + 1];          // entry 0: incoming row offset, entry 1 + w: total of warp w
    __shared__ uint64_t bars[MAXW][NS];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int page = blockIdx.y, band = blockIdx.x, bands = gridDim.x;
    const int y0 = band * rows_per_band;
    const int y1 = min(y0 + rows_per_band, rows);

    S += (size_t)page * page_stride;
    Q += (size_t)page * page_stride;

    uint8_t* slot = smem_raw + (size_t)wid * (NS * STAGE);               // this warp's ring
    const int X0 = (MULTI ? col0 : 0) + wid * WC;
    const int Xl = X0 + 4 * lane;                                        // + 128 j
    // Box origin (source byte column, multiple of 16).  Normally the aligned byte below the strip's
    // first source column; a strip lying entirely in the right replicate border is moved left so
    // that it still contains the last image column.
    int xb = (X0 - pad) & ~15;
    const bool all_right = X0 - pad >= cols;                             // every column replicates column cols-1
    const bool all_left = X0 + WC <= pad;                                // every column replicates column 0
    if (all_right) xb = (cols - 1) & ~15;
    if (all_left) xb = 0;
    const int sh = X0 - pad - xb;                                        // strip column 0 sits `sh` bytes into the box
    const bool edge = (X0 < pad) || (X0 + WC > pad + cols);              // strip touches a replicated column
    const int lcol = -xb;                                                // box offset of source column 0 (valid when X0 < pad)
    const int rcol = cols - 1 - xb;                                      // box offset of source column cols-1

    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) mbar_init(&bars[wid][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    bool st[J];
    long long accS[J][4], accQ[J][4];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        st[j] = Xl + 128 * j < (int)pitch;                               // pitch % 4 == 0
#pragma unroll
        for (int i = 0; i < 4; ++i) accS[j][i] = accQ[j][i] = 0;
        if (carry != nullptr && band > 0 && st[j]) {
            const long long* c = reinterpret_cast<const long long*>(carry) + ((size_t)page * bands + band) * 2 * pitch + Xl + 128 * j;
#pragma unroll
            for (int i = 0; i < 4; ++i) { accS[j][i] = c[i]; accQ[j][i] = c[pitch + i]; }
        }
    }

    const int n_chunks = (y1 - y0 + R - 1) / R;
    // (the descriptor address must be the kernel-parameter address itself: no lambda / local copy)
#define PRL_ISSUE_TMA(c_)                                                                               \
    do {                                                                                                \
        const int s_ = (c_) % NS;                                                                       \
        mbar_expect_tx(&bars[wid][s_], R * BOX);                                                        \
        tma_load_3d(slot + (size_t)s_ * STAGE, &tmap, xb / 2, y0 + (c_) * R, page, &bars[wid][s_]);     \
    } while (0)
    if (lane == 0)
        for (int c = 0; c < NS - 1 && c < n_chunks; ++c) PRL_ISSUE_TMA(c);

    // 4 pixels of chunk row r, sub-segment j, with the replicated border substituted
    auto fetch = [&](const uint8_t* row, int j) -> uint32_t {
        uint32_t w;
        if (all_right) {
            w = row[rcol] * 0x01010101u;
        } else if (all_left) {
            w = row[0] * 0x01010101u;
        } else {
            const int b = sh + 128 * j + 4 * lane;
            const uint32_t* wp = reinterpret_cast<const uint32_t*>(row + (b & ~3));
            w = wp[0];
            if (sh & 3) w = __funnelshift_r(w, wp[1], 8 * (sh & 3));
            if (edge) {
                const int x = Xl + 128 * j - pad;      // source column of byte 0
                const uint32_t lv = (X0 < pad) ? row[lcol] : 0u, rv = (rcol >= 0 && rcol < BOX) ? row[rcol] : 0u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int xi = x + i;
                    if (xi < 0) w = (w & ~(0xffu << (8 * i))) | (lv << (8 * i));
                    else if (xi >= cols) w = (w & ~(0xffu << (8 * i))) | (rv << (8 * i));
                }
            }
        }
        return w;
    };

    uint32_t mn4 = 0xffffffffu;
    int buf_sel = 0;
    // incoming row offsets: lane r of warp 0 carries row (chunk start + r), fetched one chunk ahead
    uint2 roff = make_uint2(0u, 0u);
    if (MULTI && has_in && wid == 0 && lane < R && y0 + lane < y1) roff = rowoff[(size_t)page * rows + y0 + lane];
    for (int c = 0; c < n_chunks; ++c, buf_sel ^= 1) {
        __syncwarp();                                         // every lane is done with slot (c-1) % NS
        if (lane == 0 && c + NS - 1 < n_chunks) PRL_ISSUE_TMA(c + NS - 1);
        const int yc = y0 + c * R;
        if (wid == 0 && lane < R) {
            tot[buf_sel][lane][0] = roff;
            roff = make_uint2(0u, 0u);
            if (MULTI && has_in && yc + R + lane < y1) roff = rowoff[(size_t)page * rows + yc + R + lane];
        }
        mbar_wait(&bars[wid][c % NS], (uint32_t)((c / NS) & 1));
        const uint8_t* buf = slot + (size_t)(c % NS) * STAGE;

        // sweep 1: this warp's row totals
#pragma unroll 2
        for (int r = 0; r < R; ++r) {
            uint32_t s = 0, q = 0;
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const uint32_t w = fetch(buf + r * BOX, j);
                if (imin != nullptr && yc + r < y1) mn4 = __vminu4(mn4, w);   // (byte-wise minimum is emulated: ~15 instructions)
                s = __dp4a(w, 0x01010101u, s);       // sum of the 4 bytes
                q = __dp4a(w, w, q);                 // sum of their squares
            }
            s = __reduce_add_sync(0xffffffffu, s);
            q = __reduce_add_sync(0xffffffffu, q);
            if (lane == 0) tot[buf_sel][r][wid + 1] = make_uint2(s, q);
        }
        __syncthreads();

        // sweep 2: scans, column accumulation, stores
#pragma unroll 2
        for (int r = 0; r < R; ++r) {
            const int y = yc + r;
            if (y < y1) {
                const uint2 t = (lane <= wid) ? tot[buf_sel][r][lane] : make_uint2(0u, 0u);
                uint32_t off_s = __reduce_add_sync(0xffffffffu, t.x);
                uint32_t off_q = __reduce_add_sync(0xffffffffu, t.y);
                if (MULTI && has_out && wid == nwarps - 1 && lane == 0) {       // running row total for the next column pass
                    const uint2 own = tot[buf_sel][r][wid + 1];
                    rowoff[(size_t)page * rows + y] = make_uint2(off_s + own.x, off_q + own.y);
                }
                uint32_t rs[J][4], rq[J][4];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const uint32_t w = fetch(buf + r * BOX, j);
                    // lane-local inclusive prefixes of the 4 bytes / their squares: one dp4a each
                    const uint32_t a0 = w & 0xffu, a1 = __dp4a(w, 0x00000101u, 0u), a2 = __dp4a(w, 0x00010101u, 0u),
                                   a3 = __dp4a(w, 0x01010101u, 0u);
                    const uint32_t b0 = a0 * a0, b1 = __dp4a(w, w & 0x0000ffffu, 0u), b2 = __dp4a(w, w & 0x00ffffffu, 0u),
                                   b3 = __dp4a(w, w, 0u);
                    const uint32_t is = warp_incl_scan(a3, lane), iq = warp_incl_scan(b3, lane);
                    const uint32_t es = off_s + is - a3, eq = off_q + iq - b3;   // exclusive base of this lane
                    rs[j][0] = es + a0; rs[j][1] = es + a1; rs[j][2] = es + a2; rs[j][3] = es + a3;
                    rq[j][0] = eq + b0; rq[j][1] = eq + b1; rq[j][2] = eq + b2; rq[j][3] = eq + b3;
                    if (J > 1) {
                        off_s += __shfl_sync(0xffffffffu, is, 31);
                        off_q += __shfl_sync(0xffffffffu, iq, 31);
                    }
                }
                // source row y -> padded rows: row 0 also feeds the `pad` rows above it, row rows-1 the rows below
                int Y = y + pad, rep = 1;
                if (y == 0) { Y = 0; rep += pad; }
                if (y == rows - 1) rep += pad;
                for (int k = 0; k < rep; ++k, ++Y) {
                    int64_t* Srow = S + (size_t)Y * pitch + Xl;
                    int64_t* Qrow = Q + (size_t)Y * pitch + Xl;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { accS[j][i] += (long long)rs[j][i]; accQ[j][i] += (long long)rq[j][i]; }
                        if (st[j]) {
                            st_v4_s64(Srow + 128 * j, accS[j][0], accS[j][1], accS[j][2], accS[j][3]);
                            st_v4_s64(Qrow + 128 * j, accQ[j][0], accQ[j][1], accQ[j][2], accQ[j][3]);
                        }
                    }
                }
            }
        }
    }

    if (imin != nullptr) {
        uint32_t m = min(min(mn4 & 0xff, (mn4 >> 8) & 0xff), min((mn4 >> 16) & 0xff, mn4 >> 24));
        m = __reduce_min_sync(0xffffffffu, m);
        if (lane == 0 && y1 > y0) atomicMin(imin + page, m);
    }
#undef PRL_ISSUE_TMA
}

// ------------------------------------------------------------------------------------------------
// generic kernel (no TMA, any alignment / pitch): 2 columns per lane x 4 sub-segments of 64 columns,
// 16-byte stores, pixels fetched with clamped byte loads.  Same band convention as the TMA kernel.
// ------------------------------------------------------------------------------------------------
template <int MAXW, int R>
__global__ void __launch_bounds__(MAXW * 32)
integral_generic_kernel(const uint8_t* __restrict__ src, size_t src_step, size_t src_page_stride, int rows, int cols,
                        int pad, int64_t* __restrict__ S, int64_t* __restrict__ Q, size_t pitch, size_t page_stride,
                        int rows_per_band, const int64_t* __restrict__ carry, uint32_t* __restrict__ imin, int vec_ok,
                        const int* __restrict__ page_map = nullptr, const int* __restrict__ page_count = nullptr, int slot_base = 0)
{
    constexpr int kSub = 4;
    __shared__ uint2 tot[2][R][MAXW];

    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int band = blockIdx.x, bands = gridDim.x;
    // indirect launch (hand-back of the fused path): plane slot blockIdx.y holds page page_map[slot_base + blockIdx.y] of the
    // batch; the list and its length live in device memory, CTAs beyond it leave at once
    int page = blockIdx.y;
    const int slot = blockIdx.y;
    if (page_map != nullptr) {
        if (slot_base + slot >= *page_count) return;
        page = page_map[slot_base + slot];
    }
    const int Wp = cols + 2 * pad;
    const int y0 = band * rows_per_band;
    const int y1 = min(y0 + rows_per_band, rows);

    src += (size_t)page * src_page_stride;
    S += (size_t)slot * page_stride;
    Q += (size_t)slot * page_stride;

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
        if (carry != nullptr && band > 0) {
            const long long* c = reinterpret_cast<const long long*>(carry) + ((size_t)page * bands + band) * 2 * pitch + Xb + 64 * j;
            if (ok[j][0]) { accS[j][0] = c[0]; accQ[j][0] = c[pitch]; }
            if (ok[j][1]) { accS[j][1] = c[1]; accQ[j][1] = c[pitch + 1]; }
        }
    }

    uint32_t mn = 255u;
    int buf = 0;
    for (int yc = y0; yc < y1; yc += R, buf ^= 1) {
        uint32_t px[R][kSub];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int y = min(yc + r, rows - 1);
            const uint8_t* rowp = src + (size_t)y * src_step;
            const bool rvld = yc + r < y1;
            uint32_t s = 0, q = 0;
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
                const uint32_t v0 = (rvld && ok[j][0]) ? (uint32_t)__ldg(rowp + xs[j][0]) : 0u;
                const uint32_t v1 = (rvld && ok[j][1]) ? (uint32_t)__ldg(rowp + xs[j][1]) : 0u;
                px[r][j] = v0 | (v1 << 16);
                s += v0 + v1;
                q += v0 * v0 + v1 * v1;
                if (rvld && ok[j][0]) mn = min(mn, v0);
                if (rvld && ok[j][1]) mn = min(mn, v1);
            }
            s = __reduce_add_sync(0xffffffffu, s);
            q = __reduce_add_sync(0xffffffffu, q);
            if (lane == 0) tot[buf][r][wid] = make_uint2(s, q);
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int y = yc + r;
            if (y < y1) {
                const uint2 t = (lane < wid) ? tot[buf][r][lane] : make_uint2(0u, 0u);
                uint32_t off_s = __reduce_add_sync(0xffffffffu, t.x);
                uint32_t off_q = __reduce_add_sync(0xffffffffu, t.y);
                uint32_t r0s[kSub], r1s[kSub], r0q[kSub], r1q[kSub];
#pragma unroll
                for (int j = 0; j < kSub; ++j) {
                    const uint32_t v0 = px[r][j] & 0xffffu, v1 = px[r][j] >> 16;
                    const uint32_t v1q = v1 * v1;
                    const uint32_t is = warp_incl_scan(v0 + v1, lane);
                    const uint32_t iq = warp_incl_scan(v0 * v0 + v1q, lane);
                    r1s[j] = off_s + is; r0s[j] = r1s[j] - v1;
                    r1q[j] = off_q + iq; r0q[j] = r1q[j] - v1q;
                    off_s += __shfl_sync(0xffffffffu, is, 31);
                    off_q += __shfl_sync(0xffffffffu, iq, 31);
                }
                int Y = y + pad, rep = 1;
                if (y == 0) { Y = 0; rep += pad; }
                if (y == rows - 1) rep += pad;
                for (int k = 0; k < rep; ++k, ++Y) {
                    int64_t* Srow = S + (size_t)Y * pitch + Xb;
                    int64_t* Qrow = Q + (size_t)Y * pitch + Xb;
#pragma unroll
                    for (int j = 0; j < kSub; ++j) {
                        accS[j][0] += (long long)r0s[j]; accS[j][1] += (long long)r1s[j];
                        accQ[j][0] += (long long)r0q[j]; accQ[j][1] += (long long)r1q[j];
                        if (vec_ok && ok[j][0] && Xb + 64 * j + 1 < (int)pitch) {
                            st_v2_s64(Srow + 64 * j, accS[j][0], accS[j][1]);
                            st_v2_s64(Qrow + 64 * j, accQ[j][0], accQ[j][1]);
                        } else {
                            if (ok[j][0]) { Srow[64 * j] = accS[j][0]; Qrow[64 * j] = accQ[j][0]; }
                            if (ok[j][1]) { Srow[64 * j + 1] = accS[j][1]; Qrow[64 * j + 1] = accQ[j][1]; }
                        }
                    }
                }
            }
        }
    }

    if (imin != nullptr) {
        mn = __reduce_min_sync(0xffffffffu, mn);
        if (lane == 0 && y1 > y0) atomicMin(imin + page, mn);
    }
}

// ---- latency mode: weighted per-band column sums, then row-scanned carries -------------------
// colsum[page][band][plane][X] = sum over the band's source rows y of mult(y) * P(y, X), where
// mult counts the replicated copies (row 0 and row rows-1 appear pad+1 times in the padded image).
__global__ void __launch_bounds__(256)
band_colsum_kernel(const uint8_t* __restrict__ src, size_t src_step, size_t src_page_stride, int rows, int cols,
                   int pad, int rows_per_band, unsigned long long* __restrict__ colsum, size_t pitch)
{
    // grid (column tiles of 1024, bands, pages); each thread owns 4 adjacent padded columns
    const int page = blockIdx.z, band = blockIdx.y, bands = gridDim.y;
    const int Wp = cols + 2 * pad;
    const int y0 = band * rows_per_band, y1 = min(y0 + rows_per_band, rows);
    const int X = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (X >= Wp) return;
    src += (size_t)page * src_page_stride;
    int xs[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xs[i] = min(max(X + i - pad, 0), cols - 1);
    unsigned int s[4] = {0, 0, 0, 0};
    unsigned long long q[4] = {0, 0, 0, 0};
    unsigned long long sb[4] = {0, 0, 0, 0};      // replicated border rows carry a multiplicity > 1
    int y = y0;
    for (; y + 4 <= y1; y += 4) {
        unsigned int p[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 4; ++i) p[r][i] = __ldg(src + (size_t)(y + r) * src_step + xs[i]);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int yy = y + r;
            unsigned int mult = 1;
            if (yy == 0) mult += pad;
            if (yy == rows - 1) mult += pad;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (mult == 1) { s[i] += p[r][i]; q[i] += p[r][i] * p[r][i]; }
                else { sb[i] += (unsigned long long)mult * p[r][i]; q[i] += (unsigned long long)mult * p[r][i] * p[r][i]; }
            }
        }
    }
    for (; y < y1; ++y) {
        unsigned long long mult = 1;
        if (y == 0) mult += pad;
        if (y == rows - 1) mult += pad;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const unsigned long long p = __ldg(src + (size_t)y * src_step + xs[i]);
            sb[i] += mult * p; q[i] += mult * p * p;
        }
    }
    unsigned long long* out = colsum + ((size_t)page * bands + band) * 2 * pitch;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (X + i < Wp) { out[X + i] = sb[i] + s[i]; out[pitch + X + i] = q[i]; }
}

// Same sums per SOURCE column (the carry kernel maps padded columns onto them), for 16-byte aligned pages: a thread owns
// 16 adjacent columns (one 16-byte load per row, 8 rows in flight), a CTA 2048 columns of one slice of 128 rows of a
// band; 32-bit partial sums (255^2 * 128 * (pad + 1) < 2^32 for pad < 515) are added to the band's 64-bit sums with atomics.
__global__ void __launch_bounds__(128)
band_colsum_vec_kernel(const uint8_t* __restrict__ src, size_t src_step, size_t src_page_stride, int rows, int cols,
                       int pad, int rows_per_band, int slices, unsigned long long* __restrict__ colsum, size_t pitch)
{
    const int page = blockIdx.z, band = blockIdx.y / slices, slice = blockIdx.y - band * slices, bands = gridDim.y / slices;
    const int x = (blockIdx.x * 128 + threadIdx.x) * 16;
    if (x >= cols) return;
    const int yb0 = band * rows_per_band, yb1 = min(yb0 + rows_per_band, rows);
    const int y0 = yb0 + slice * 128, y1 = min(y0 + 128, yb1);
    if (y0 >= y1) return;
    src += (size_t)page * src_page_stride + x;
    uint32_t s[16], q[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) s[i] = q[i] = 0;
    auto add = [&](const uint4& v, uint32_t mult) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t p = (w[k] >> (8 * i)) & 0xffu;
                s[4 * k + i] += mult * p; q[4 * k + i] += mult * p * p;
            }
    };
    int y = y0;
    for (; y + 8 <= y1; y += 8) {
        uint4 v[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) v[r] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(y + r) * src_step));
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int yy = y + r;
            add(v[r], 1u + ((yy == 0) ? pad : 0) + ((yy == rows - 1) ? pad : 0));      // replicated border rows count pad more times
        }
    }
    for (; y < y1; ++y)
        add(__ldg(reinterpret_cast<const uint4*>(src + (size_t)y * src_step)), 1u + ((y == 0) ? pad : 0) + ((y == rows - 1) ? pad : 0));
    unsigned long long* out = colsum + ((size_t)page * bands + band) * 2 * pitch + x;
#pragma unroll
    for (int i = 0; i < 16; ++i)
        if (x + i < cols) { atomicAdd(out + i, (unsigned long long)s[i]); atomicAdd(out + pitch + i, (unsigned long long)q[i]); }
}

// carry[page][band][plane][X] = sum_{b' < band} sum_{X' <= X} colsum[page][b'][plane][X']
//                             = the integral row just above the band's first emitted padded row
__global__ void __launch_bounds__(1024)
band_carry_kernel(const unsigned long long* __restrict__ colsum, long long* __restrict__ carry, int Wp, size_t pitch,
                  int src_cols, int pad)
{
    // src_cols > 0: colsum is indexed by source column; padded column X reads source column clamp(X - pad)
    __shared__ unsigned long long wtot[2][32];
    const int page = blockIdx.y, band = blockIdx.x, bands = gridDim.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned long long* in = colsum + (size_t)page * bands * 2 * pitch;
    long long* out = carry + ((size_t)page * bands + band) * 2 * pitch;
    unsigned long long run_s = 0, run_q = 0;
    for (int base = 0; base < Wp; base += 1024) {
        const int X = base + threadIdx.x;
        unsigned long long vs = 0, vq = 0;
        if (X < Wp) {
            const int xi = src_cols > 0 ? min(max(X - pad, 0), src_cols - 1) : X;
            for (int b = 0; b < band; ++b) {
                vs += in[(size_t)b * 2 * pitch + xi];
                vq += in[(size_t)b * 2 * pitch + pitch + xi];
            }
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

// ---- host side ----------------------------------------------------------------------------------
// Number of row bands a page is cut into: 1 when the batch alone fills the machine.
}  // namespace

int prl_choose_bands(const prl_cuda_ctx* ctx, int n_pages, int rows, int ctas_per_sm)
{
    if (ctx->k1_bands > 0) return std::max(1, std::min(ctx->k1_bands, rows / 32));   // tuning knob (set_option "k1_bands")
    const int want = ctas_per_sm * ctx->num_sms;
    if (n_pages >= want - want / 4) return 1;   // >= 1.5 pages per SM: the batch alone fills the machine
    int bands = want / n_pages;                 // floor: one full wave of CTAs, never a small second wave
    int max_bands = rows / 32; if (max_bands < 1) max_bands = 1;
    if (bands > max_bands) bands = max_bands;
    if (bands > 64) bands = 64;
    return bands < 1 ? 1 : bands;
}

namespace {

template <int MAXW, int MINB, int J, int R, int NS, bool MULTI>
int launch_tma(prl_cuda_ctx* ctx, encode_tiled_fn enc, dim3 grid, int nwarps, const uint8_t* d_src, int n_pages,
               size_t src_step, size_t src_page_stride, int rows, int cols, int pad, int64_t* d_S, int64_t* d_Q,
               size_t pitch, size_t plane_page_stride, int rpb, const int64_t* d_carry, uint32_t* d_imin, bool* launched,
               int col0 = 0, uint2* rowoff = nullptr, int has_in = 0, int has_out = 0)
{
    *launched = false;
    CUtensorMap tmap;
    // rows described as src_step/2 u16 elements: a (128J+16)-byte box row is <= 136 elements (<= 256 allowed)
    const cuuint64_t gdim[3] = {(cuuint64_t)(src_step / 2), (cuuint64_t)rows, (cuuint64_t)n_pages};
    const cuuint64_t gstr[2] = {(cuuint64_t)src_step, (cuuint64_t)(n_pages > 1 ? src_page_stride : src_step * rows)};
    const cuuint32_t box[3] = {(cuuint32_t)(box_bytes<J>() / 2), R, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, (void*)d_src, gdim, gstr, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return PRL_OK;     // caller falls back to the generic kernel
    const size_t smem = (size_t)nwarps * NS * stage_bytes<J, R>();
    auto kfn = integral_tma_kernel<MAXW, MINB, J, R, NS, MULTI>;
    // (per device: the attribute lives in the current device's copy of the function, so no process-wide cache)
    PRL_CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kfn<<<grid, nwarps * 32, smem, ctx->stream>>>(tmap, rows, cols, pad, d_S, d_Q, pitch, plane_page_stride, rpb, d_carry, d_imin,
                                                  col0, rowoff, has_in, has_out);
    *launched = true;
    return PRL_OK;
}

}  // namespace

// Latency mode pre-pass: the top carry of every band (the integral row just above its first emitted padded row),
// carry[page][band][plane S/Q][pitch] int64.
int prl_band_carries(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                     size_t src_page_stride, int pad, int bands, int rpb, size_t pitch, const int64_t** d_carry)
{
    const int Wp = cols + 2 * pad;
    size_t need = (size_t)n_pages * bands * 2 * pitch * sizeof(int64_t);
    int rc = prl_ensure(ctx, &ctx->colsum, &ctx->colsum_bytes, need); if (rc) return rc;
    rc = prl_ensure(ctx, &ctx->carry, &ctx->carry_bytes, need); if (rc) return rc;
    const bool vec_src = (((uintptr_t)d_src | src_step | src_page_stride) & 15) == 0 && pad < 512 &&
                         src_step >= (size_t)((cols + 15) & ~15) && (size_t)bands * ((rpb + 127) / 128) <= 65535;
    if (vec_src) {
        const int slices = (rpb + 127) / 128;
        PRL_CUDA_TRY(ctx, cudaMemsetAsync(ctx->colsum, 0, need, ctx->stream));
        prl_launch_scope ls(ctx, FAM_BAND_CARRY);
        band_colsum_vec_kernel<<<dim3((cols + 2047) / 2048, bands * slices, n_pages), 128, 0, ctx->stream>>>(
            d_src, src_step, src_page_stride, rows, cols, pad, rpb, slices, (unsigned long long*)ctx->colsum, pitch);
    } else {
        prl_launch_scope ls(ctx, FAM_BAND_CARRY);
        band_colsum_kernel<<<dim3((Wp + 1023) / 1024, bands, n_pages), 256, 0, ctx->stream>>>(
            d_src, src_step, src_page_stride, rows, cols, pad, rpb, (unsigned long long*)ctx->colsum, pitch);
    }
    {
        prl_launch_scope ls(ctx, FAM_BAND_CARRY);
        band_carry_kernel<<<dim3(bands, n_pages), 1024, 0, ctx->stream>>>(
            (const unsigned long long*)ctx->colsum, (long long*)ctx->carry, Wp, pitch, vec_src ? cols : 0, pad);
    }
    *d_carry = (const int64_t*)ctx->carry;
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

int prl_k_integral(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                   size_t src_page_stride, int pad, int64_t* d_S, int64_t* d_Q, size_t pitch,
                   size_t plane_page_stride, uint32_t* d_imin)
{
    const int Wp = cols + 2 * pad;
    if (n_pages > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "more than 65535 pages per launch");

    // TMA path: 16-byte aligned source rows, 32-byte aligned planes, pitch % 4 == 0
    encode_tiled_fn enc = get_encode_tiled();
    const bool tma_ok = enc != nullptr && !ctx->no_tma && (((uintptr_t)d_src | src_step | src_page_stride) & 15) == 0 &&
                        ((((uintptr_t)d_S) | ((uintptr_t)d_Q)) & 31) == 0 && (pitch & 3) == 0 && (plane_page_stride & 3) == 0;
    const bool vec_ok = ((((uintptr_t)d_S) | ((uintptr_t)d_Q)) & 15) == 0 && (pitch & 1) == 0 && (plane_page_stride & 1) == 0;
    const bool narrow = tma_ok && Wp <= 24 * 128;        // 128 columns per warp, up to 24 warps, 2 CTAs per SM
    // wider pages run as chained column passes on the TMA path (any width); the generic kernel covers one pass only
    if (Wp > 8192 && !tma_ok)
        return prl_set_err(ctx, PRL_E_UNSUPPORTED, "padded width > 8192 columns needs 16-byte aligned pages and 32-byte aligned planes");

    constexpr int R = 8;
    int bands = prl_choose_bands(ctx, n_pages, rows, 2);
    int rpb = (rows + bands - 1) / bands;
    rpb = (rpb + R - 1) / R * R;
    bands = (rows + rpb - 1) / rpb;

    if (d_imin) PRL_CUDA_TRY(ctx, cudaMemsetAsync(d_imin, 0xff, sizeof(uint32_t) * n_pages, ctx->stream));

    const int64_t* d_carry = nullptr;
    if (bands > 1) {
        int rc = prl_band_carries(ctx, d_src, n_pages, rows, cols, src_step, src_page_stride, pad, bands, rpb, pitch, &d_carry);
        if (rc) return rc;
    }

    prl_launch_scope ls(ctx, FAM_INTEGRAL);
    dim3 grid(bands, n_pages);
    if (tma_ok) {
        bool launched = false;
        int rc;
        if (narrow && Wp <= 20 * 128) {
            const int nw = (Wp + 127) / 128;          // A4-class widths: 51 registers available at 2 CTAs/SM
            rc = launch_tma<20, 2, 1, R, 3, false>(ctx, enc, grid, nw, d_src, n_pages, src_step, src_page_stride, rows, cols, pad, d_S, d_Q,
                                                   pitch, plane_page_stride, rpb, d_carry, d_imin, &launched);
        } else if (narrow) {
            const int nw = (Wp + 127) / 128;          // up to 3072 columns: still one pass (42 registers, small spills)
            rc = launch_tma<24, 2, 1, R, 3, false>(ctx, enc, grid, nw, d_src, n_pages, src_step, src_page_stride, rows, cols, pad, d_S, d_Q,
                                                   pitch, plane_page_stride, rpb, d_carry, d_imin, &launched);
        } else {
            // wide page: column passes of <= 20 strips of 128 columns each, chained through rowoff
            const int nw_total = (Wp + 127) / 128;
            const int npass = (nw_total + 19) / 20, wp = (nw_total + npass - 1) / npass;
            rc = prl_ensure(ctx, &ctx->d_misc, &ctx->d_misc_bytes, (size_t)n_pages * rows * sizeof(uint2)); if (rc) return rc;
            for (int p = 0; p < npass; ++p) {
                const int nw = std::min(wp, nw_total - p * wp);
                rc = launch_tma<20, 2, 1, R, 3, true>(ctx, enc, grid, nw, d_src, n_pages, src_step, src_page_stride, rows, cols, pad, d_S, d_Q,
                                                pitch, plane_page_stride, rpb, d_carry, d_imin, &launched, p * wp * 128,
                                                (uint2*)ctx->d_misc, p > 0, p + 1 < npass);
                if (rc || !launched) break;
            }
        }
        if (rc) return rc;
        if (launched) {
            PRL_CUDA_TRY(ctx, cudaGetLastError());
            return PRL_OK;
        }
        // the driver rejected the tensor map (e.g. stride limits): fall through to the generic kernel
    }
    const int nwarps = (Wp + kWarpCols - 1) / kWarpCols;
    if (nwarps > 32) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "padded width > 8192 columns: tensor map rejected, no generic path");
    if (nwarps <= 16)
        integral_generic_kernel<16, 4><<<grid, nwarps * 32, 0, ctx->stream>>>(
            d_src, src_step, src_page_stride, rows, cols, pad, d_S, d_Q, pitch, plane_page_stride, rpb, d_carry, d_imin, vec_ok);
    else
        integral_generic_kernel<32, 4><<<grid, nwarps * 32, 0, ctx->stream>>>(
            d_src, src_step, src_page_stride, rows, cols, pad, d_S, d_Q, pitch, plane_page_stride, rpb, d_carry, d_imin, vec_ok);
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

// int64 planes of the pages page_map[slot_base .. slot_base + slots) (as many as *page_count says exist) into plane
// slots 0 .. slots-1: the generic kernel, one CTA per page.  Serves the device-side hand-back of the fused path.
int prl_k_integral_indirect(prl_cuda_ctx* ctx, const uint8_t* d_src, int slots, int rows, int cols, size_t src_step,
                            size_t src_page_stride, int pad, int64_t* d_S, int64_t* d_Q, size_t pitch, size_t plane_page_stride,
                            const int* d_map, const int* d_count, int slot_base)
{
    const int Wp = cols + 2 * pad;
    const int nwarps = (Wp + kWarpCols - 1) / kWarpCols;
    if (nwarps > 32 || slots > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "indirect integral: page too wide or too many slots");
    const bool vec_ok = ((((uintptr_t)d_S) | ((uintptr_t)d_Q)) & 15) == 0 && (pitch & 1) == 0 && (plane_page_stride & 1) == 0;
    int rpb = (rows + 3) / 4 * 4;
    prl_launch_scope ls(ctx, FAM_INTEGRAL);
    dim3 grid(1, slots);
    if (nwarps <= 16)
        integral_generic_kernel<16, 4><<<grid, nwarps * 32, 0, ctx->stream>>>(
            d_src, src_step, src_page_stride, rows, cols, pad, d_S, d_Q, pitch, plane_page_stride, rpb, nullptr, nullptr, vec_ok,
            d_map, d_count, slot_base);
    else
        integral_generic_kernel<32, 4><<<grid, nwarps * 32, 0, ctx->stream>>>(
            d_src, src_step, src_page_stride, rows, cols, pad, d_S, d_Q, pitch, plane_page_stride, rpb, nullptr, nullptr, vec_ok,
            d_map, d_count, slot_base);
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

bool prl_integral_compact_ok(const prl_cuda_ctx* ctx, const uint8_t* d_src, size_t src_step, size_t src_page_stride, int rows, int cols, int pad)
{
    return get_encode_tiled() != nullptr && !ctx->no_tma && (((uintptr_t)d_src | src_step | src_page_stride) & 15) == 0 &&
           cols + 2 * pad <= 8192 && prl_anchor_shift(rows + 2 * pad, cols + 2 * pad) >= 0;
}

int prl_k_integral_planes(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                          size_t src_page_stride, int pad, const prl_planes& P, uint32_t* d_imin)
{
    if (P.compact) return prl_k_integral_sq(ctx, d_src, n_pages, rows, cols, src_step, src_page_stride, pad, P, d_imin);
    return prl_k_integral(ctx, d_src, n_pages, rows, cols, src_step, src_page_stride, pad, (int64_t*)P.S, (int64_t*)P.Q, P.pitch,
                          P.page_stride, d_imin);
}
