// morph.cu -- morphology tail of the five local-statistics binarizers.
//
// Replaces cv::dilate / cv::erode with the default 3x3 element and `n` iterations
// (binarizeSauvola.cpp:125-134, binarizeNiblack.cpp:115-127, binarizeWolfJolion.cpp:138-147,
// binarizeNICK.cpp:134-143, binarizeFeng.cpp:151-163).  n iterations of a 3x3 rectangle are one
// (2n+1)x(2n+1) rectangle; pixels outside the image are ignored (morphologyDefaultBorderValue), i.e.
// they act as 0 for a dilation and as 255 for an erosion.
//
// morph_fused_kernel does the whole closing (or opening) in ONE pass over HBM (1 byte read + 1 byte
// written per pixel): a CTA stages its tile plus a 2n halo in shared memory and runs the four
// separable passes (row max, column max, row min, column min) there, 4 pixels per thread with the
// byte-SIMD __vmaxu4 / __vminu4 and funnel shifts for the horizontal taps.  morph_pass_kernel is the
// generic one-pass-per-launch fallback for n > 8.
#include "common.cuh"
#include <algorithm>

namespace {

constexpr int kSWw = 64;                      // staged tile: 64 words (256 bytes) wide ...
constexpr int kSW = kSWw * 4;
constexpr int kSH = 64;                       // ... and 64 rows high, halo included
constexpr int kMaxN = 8;                      // fused kernel handles n <= 8

template <bool IS_MAX> __device__ __forceinline__ uint32_t vop(uint32_t a, uint32_t b) { return IS_MAX ? __vmaxu4(a, b) : __vminu4(a, b); }

// horizontal (2n+1) max/min, rows [r0, r1), word columns [c0, c1): warp w takes rows w, w+8, ..; lanes take words
template <bool IS_MAX, int NMAX>
__device__ __forceinline__ void hpass(const uint32_t* in, uint32_t* out, int n, int r0, int r1, int c0, int c1)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int r = r0 + wid; r < r1; r += 8) {
        const uint32_t* row = in + r * kSWw;
        for (int c = c0 + lane; c < c1; c += 32) {
            // neighbours outside the staged row only feed bytes nobody consumes (see the halo bookkeeping below)
            const uint32_t w0 = row[c];
            const uint32_t wm1 = c > 0 ? row[c - 1] : w0, wp1 = c < kSWw - 1 ? row[c + 1] : w0;
            uint32_t wm2 = w0, wp2 = w0;
            if (NMAX > 4) { if (c > 1) wm2 = row[c - 2]; if (c < kSWw - 2) wp2 = row[c + 2]; }
            uint32_t acc = w0;
#pragma unroll
            for (int k = 1; k <= NMAX; ++k) {
                if (k <= n) {
                    uint32_t left, right;     // pixels x-k and x+k
                    if (k < 4) { left = __funnelshift_r(wm1, w0, 32 - 8 * k); right = __funnelshift_r(w0, wp1, 8 * k); }
                    else if (k == 4) { left = wm1; right = wp1; }
                    else if (k < 8) { left = __funnelshift_r(wm2, wm1, 32 - 8 * (k - 4)); right = __funnelshift_r(wp1, wp2, 8 * (k - 4)); }
                    else { left = wm2; right = wp2; }
                    acc = vop<IS_MAX>(acc, vop<IS_MAX>(left, right));
                }
            }
            out[r * kSWw + c] = acc;
        }
    }
}

// vertical (2n+1) max/min, rows [r0, r1), word columns [c0, c1)
template <bool IS_MAX>
__device__ __forceinline__ void vpass(const uint32_t* in, uint32_t* out, int n, int r0, int r1, int c0, int c1)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int r = r0 + wid; r < r1; r += 8)
        for (int c = c0 + lane; c < c1; c += 32) {
            uint32_t acc = in[r * kSWw + c];
            for (int k = 1; k <= n; ++k) acc = vop<IS_MAX>(acc, vop<IS_MAX>(in[(r - k) * kSWw + c], in[(r + k) * kSWw + c]));
            out[r * kSWw + c] = acc;
        }
}

// 4 bytes at `p` (any alignment, all four inside the row)
__device__ __forceinline__ uint32_t ld_u32_unaligned(const uint8_t* p)
{
    const uint32_t sh = (uint32_t)(uintptr_t)p & 3u;
    const uint32_t* a = reinterpret_cast<const uint32_t*>(p - sh);
    uint32_t w = __ldg(a);
    if (sh) w = __funnelshift_r(w, __ldg(a + 1), 8 * sh);
    return w;
}

// CLOSE = true: dilate then erode (morph_iters > 0); false: erode then dilate (morph_iters < 0).
// HALO = 2 * (largest n served) rounded to a multiple of 4: 4 (n <= 2), 8 (n <= 4), 16 (n <= 8).
// Output tile: (256 - 2 HALO) x (64 - 2 HALO) pixels.
template <bool CLOSE, int HALO>
__global__ void __launch_bounds__(256)
morph_fused_kernel(const uint8_t* __restrict__ src, size_t src_step, size_t src_page_stride, uint8_t* __restrict__ dst,
                   size_t dst_step, size_t dst_page_stride, int rows, int cols, int n)
{
    constexpr int TW = kSW - 2 * HALO, TH = kSH - 2 * HALO, NMAX = HALO / 2, HW = HALO / 4;
    __shared__ uint32_t bufA[kSH * kSWw], bufB[kSH * kSWw];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    src += (size_t)blockIdx.z * src_page_stride;
    dst += (size_t)blockIdx.z * dst_page_stride;
    const uint32_t first_fill = CLOSE ? 0u : 255u, second_fill = CLOSE ? 255u : 0u;

    // stage tile + halo; pixels outside the image take the neutral value of the first operator
    for (int r = wid; r < kSH; r += 8) {
        const int y = ty0 - HALO + r;
        for (int c = lane; c < kSWw; c += 32) {
            const int x = tx0 - HALO + 4 * c;
            uint32_t w = first_fill * 0x01010101u;
            if (y >= 0 && y < rows && x + 3 >= 0 && x < cols) {
                const uint8_t* p = src + (size_t)y * src_step + x;
                if (x >= 0 && x + 7 < cols) w = ld_u32_unaligned(p);     // (+7: the aligned pair may reach 3 bytes further)
                else {
                    w = 0;
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t v = (x + k >= 0 && x + k < cols) ? (uint32_t)p[k] : first_fill;
                        w |= v << (8 * k);
                    }
                }
            }
            bufA[r * kSWw + c] = w;
        }
    }
    __syncthreads();
    // Halo bookkeeping (bytes in x, rows in y; HALO >= 2n): outputs [HALO, S-HALO) need the first operator's
    // result on [HALO-n, S-HALO+n), which needs the input on [HALO-2n, S-HALO+2n) -- all staged.  Every pass is
    // run over whole words / all rows it can; results outside those ranges are garbage that nothing reads.
    // pass 1 (rows)
    hpass<CLOSE, NMAX>(bufA, bufB, n, 0, kSH, 0, kSWw);
    __syncthreads();
    // pass 2 (columns)
    vpass<CLOSE>(bufB, bufA, n, n, kSH - n, 0, kSWw);
    __syncthreads();
    // the second operator ignores what lies outside the image, whatever the first one produced there
    for (int r = wid; r < kSH; r += 8) {
        const int y = ty0 - HALO + r;
        for (int c = lane; c < kSWw; c += 32) {
            const int x = tx0 - HALO + 4 * c;
            if (y < 0 || y >= rows || x + 3 < 0 || x >= cols) bufA[r * kSWw + c] = second_fill * 0x01010101u;
            else if (x < 0 || x + 3 >= cols) {
                uint32_t w = bufA[r * kSWw + c];
                for (int k = 0; k < 4; ++k)
                    if (x + k < 0 || x + k >= cols) w = (w & ~(0xffu << (8 * k))) | (second_fill << (8 * k));
                bufA[r * kSWw + c] = w;
            }
        }
    }
    __syncthreads();
    // pass 3 (rows)
    hpass<!CLOSE, NMAX>(bufA, bufB, n, n, kSH - n, 0, kSWw);
    __syncthreads();
    // pass 4 (columns) straight to global: staged rows [HALO, HALO + TH), words [HW, HW + TW/4)
    for (int r = HALO + wid; r < HALO + TH; r += 8) {
        const int y = ty0 + (r - HALO);
        if (y >= rows) break;
        for (int c = HW + lane; c < HW + TW / 4; c += 32) {
            const int x = tx0 + 4 * (c - HW);
            if (x >= cols) break;
            uint32_t acc = bufB[r * kSWw + c];
            for (int k = 1; k <= n; ++k) acc = vop<!CLOSE>(acc, vop<!CLOSE>(bufB[(r - k) * kSWw + c], bufB[(r + k) * kSWw + c]));
            uint8_t* o = dst + (size_t)y * dst_step + x;
            if (x + 3 < cols && ((uintptr_t)o & 3) == 0) *reinterpret_cast<uint32_t*>(o) = acc;
            else for (int k = 0; k < 4; ++k) if (x + k < cols) o[k] = (uint8_t)(acc >> (8 * k));
        }
    }
}

template <bool CLOSE, int HALO>
void launch_fused(prl_cuda_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_pages, int rows, int cols, size_t in_step,
                  size_t in_page_stride, size_t out_step, size_t out_page_stride, int n)
{
    constexpr int TW = kSW - 2 * HALO, TH = kSH - 2 * HALO;
    dim3 grid((cols + TW - 1) / TW, (rows + TH - 1) / TH, n_pages);
    morph_fused_kernel<CLOSE, HALO><<<grid, 256, 0, ctx->stream>>>(d_in, in_step, in_page_stride, d_out, out_step, out_page_stride, rows, cols, n);
}

// ---- binary masks: the whole closing / opening on bit-packed rows --------------------------------------
// The masks this tail runs on hold only 0 and 255 (binarizeSauvola.cpp:122), so max is OR and min is AND on
// one bit per pixel.  One warp owns a strip of 30 x 32 = 960 output columns (lanes 0 and 31 carry the 32-pixel
// halos, enough for n <= 15) of one band of rows and streams down the rows:
//   two 16-byte loads per lane and row -> 32 bits (signed x unsigned dp4a gathers the 0/-1 bytes),
//   horizontal pass with funnel shifts against the neighbour lanes' words,
//   vertical pass over a lane-private ring of the last 2n+1 rows in shared memory (no barrier, no conflicts),
//   the same twice (second operator, with everything outside the image forced to its neutral value),
//   32 bits -> 32 bytes, re-aligned by bit shifts across lanes so that every full store is a 16-byte store
//   whatever the destination pitch (the dense masks of the batch path have an odd pitch).
// ~3 instructions per pixel instead of ~25 for the byte version; the input must be 16-byte aligned (it is the
// library's own scratch); anything else takes the byte kernels above.
constexpr int kBStrip = 960;

__device__ __forceinline__ int dp4a_su(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// 32 mask bytes (0 / 255) -> 32 bits, bit i = pixel i
__device__ __forceinline__ uint32_t pack32(const uint4& a, const uint4& b)
{
    const int v0 = dp4a_su(a.y, 0x80402010u, dp4a_su(a.x, 0x08040201u, 0));      // = -(bits 0..7)
    const int v1 = dp4a_su(a.w, 0x80402010u, dp4a_su(a.z, 0x08040201u, 0));
    const int v2 = dp4a_su(b.y, 0x80402010u, dp4a_su(b.x, 0x08040201u, 0));
    const int v3 = dp4a_su(b.w, 0x80402010u, dp4a_su(b.z, 0x08040201u, 0));
    return 0u - ((uint32_t)v0 + ((uint32_t)v1 << 8) + ((uint32_t)v2 << 16) + ((uint32_t)v3 << 24));
}

// 4 bits (nibble k of `bits`) -> 4 bytes 0 / 255: the multiply parks bit i on the sign bit of byte i, prmt replicates it
__device__ __forceinline__ uint32_t expand4(uint32_t bits, int k)
{
    const uint32_t t = ((bits >> (4 * k)) & 0xfu) * 0x10204080u;
    uint32_t r;
    asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(r) : "r"(t));
    return r;
}

// horizontal (2n+1) OR / AND of a bit row; L / R are the neighbour lanes' words
template <bool IS_OR>
__device__ __forceinline__ uint32_t hbits(uint32_t b, int n, int lane)
{
    uint32_t L = __shfl_up_sync(0xffffffffu, b, 1), R = __shfl_down_sync(0xffffffffu, b, 1);
    if (lane == 0) L = IS_OR ? 0u : 0xffffffffu;
    if (lane == 31) R = IS_OR ? 0u : 0xffffffffu;
    uint32_t acc = b;
    for (int k = 1; k <= n; ++k) {
        const uint32_t a = __funnelshift_l(L, b, k), c = __funnelshift_r(b, R, k);   // pixels x-k and x+k
        acc = IS_OR ? (acc | a | c) : (acc & a & c);
    }
    return acc;
}

template <bool CLOSE, int RR>
__global__ void __launch_bounds__(128)
morph_bits_kernel(const uint8_t* __restrict__ src, size_t src_step, size_t src_page_stride, uint8_t* __restrict__ dst,
                  size_t dst_step, size_t dst_page_stride, int rows, int cols, int n1, int n2, int ns, int nb, int band_rows, int n_pages)
{
    __shared__ uint32_t ring[4][2][RR][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int gw = blockIdx.x * 4 + wid;
    const int per_page = ns * nb;
    if (gw >= per_page * n_pages) return;                 // warps are independent
    const int page = gw / per_page, rem = gw - page * per_page;
    const int band = rem / ns, strip = rem - band * ns;
    src += (size_t)page * src_page_stride;
    dst += (size_t)page * dst_page_stride;
    uint32_t (*ringA)[32] = ring[wid][0];
    uint32_t (*ringB)[32] = ring[wid][1];
    const int xs = strip * kBStrip, xe = min(xs + kBStrip, cols);
    const int X0 = xs + 32 * (lane - 1);                  // first pixel column of this lane's word
    uint32_t colmask = 0;                                 // bits of in-image columns
    if (X0 + 31 >= 0 && X0 < cols) {
        colmask = 0xffffffffu;
        if (X0 < 0) colmask = 0;                          // (X0 is a multiple of 32: a word is never cut on the left)
        if (X0 + 32 > cols) colmask &= 0xffffffffu >> (X0 + 32 - cols);
    }
    const bool ld0 = X0 >= 0 && X0 < cols, ld1 = X0 >= 0 && X0 + 16 < cols;
    const int y0 = band * band_rows, y1 = min(y0 + band_rows, rows);
    const uint32_t ones = 0xffffffffu;

    auto fetch = [&](int Y, uint4& a, uint4& b) {
        a = make_uint4(0, 0, 0, 0); b = a;
        if (Y >= 0 && Y < rows) {
            const uint8_t* p = src + (size_t)Y * src_step + X0;
            if (ld0) a = __ldg(reinterpret_cast<const uint4*>(p));
            if (ld1) b = __ldg(reinterpret_cast<const uint4*>(p + 16));
        }
    };
    // first operator over (2 n1 + 1)^2, second over (2 n2 + 1)^2 (n2 = 0: the first operator alone)
    uint4 na, nbv;
    fetch(y0 - n1 - n2, na, nbv);
    for (int Y = y0 - n1 - n2; Y < y1 + n1 + n2; ++Y) {
        uint32_t b = pack32(na, nbv);
        fetch(Y + 1, na, nbv);
        // first operator; what lies outside the image is ignored = its neutral value
        const bool in1 = Y >= 0 && Y < rows;
        b = CLOSE ? (in1 ? (b & colmask) : 0u) : (in1 ? (b | ~colmask) : ones);
        ringA[Y & (RR - 1)][lane] = hbits<CLOSE>(b, n1, lane);
        if (Y < y0 + n1 - n2) continue;
        const int yd = Y - n1;                            // row whose first-operator window is complete
        uint32_t D = ringA[Y & (RR - 1)][lane];
        for (int j = 1; j <= 2 * n1; ++j) { const uint32_t v = ringA[(Y - j) & (RR - 1)][lane]; D = CLOSE ? (D | v) : (D & v); }
        // second operator
        const bool in2 = yd >= 0 && yd < rows;
        D = CLOSE ? (in2 ? (D | ~colmask) : ones) : (in2 ? (D & colmask) : 0u);
        ringB[yd & (RR - 1)][lane] = hbits<!CLOSE>(D, n2, lane);
        const int ye = yd - n2;
        if (ye < y0 || ye >= y1) continue;
        uint32_t E = ringB[yd & (RR - 1)][lane];
        for (int j = 1; j <= 2 * n2; ++j) { const uint32_t v = ringB[(yd - j) & (RR - 1)][lane]; E = CLOSE ? (E & v) : (E | v); }
        // store row ye: shift the bit row so that this lane's 32 bytes start on a 16-byte boundary
        uint8_t* drow = dst + (size_t)ye * dst_step;
        const int m = (int)((uintptr_t)(drow + xs) & 15);
        const uint32_t El = __shfl_up_sync(0xffffffffu, E, 1);
        const uint32_t Ea = __funnelshift_l(El, E, m);    // byte j of the chunk <-> pixel X0 - m + j
        const int lo = max(0, xs - X0 + m), hi = min(32, xe - X0 + m);
        if (lo >= hi) continue;
        uint8_t* o = drow + X0 - m;
        if (lo == 0 && hi == 32) {
            *reinterpret_cast<uint4*>(o) = make_uint4(expand4(Ea, 0), expand4(Ea, 1), expand4(Ea, 2), expand4(Ea, 3));
            *reinterpret_cast<uint4*>(o + 16) = make_uint4(expand4(Ea, 4), expand4(Ea, 5), expand4(Ea, 6), expand4(Ea, 7));
        } else {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t w = expand4(Ea, k);
                if (4 * k >= lo && 4 * k + 4 <= hi) *reinterpret_cast<uint32_t*>(o + 4 * k) = w;
                else
                    for (int i = 0; i < 4; ++i)
                        if (4 * k + i >= lo && 4 * k + i < hi) o[4 * k + i] = (uint8_t)(w >> (8 * i));
            }
        }
    }
}

template <bool CLOSE>
void launch_bits(prl_cuda_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_pages, int rows, int cols, size_t in_step,
                 size_t in_page_stride, size_t out_step, size_t out_page_stride, int n1, int n2)
{
    const int n = std::max(n1, n2);
    const int ns = (cols + kBStrip - 1) / kBStrip;
    // row bands: about one wave of 48 resident warps per SM; a band redoes 2 (n1 + n2) rows of its neighbours
    const long long want = (long long)ctx->num_sms * 48;
    int nb = (int)((want + (long long)ns * n_pages - 1) / ((long long)ns * n_pages));
    nb = std::max(1, std::min(nb, (rows + 63) / 64));
    const int band_rows = (rows + nb - 1) / nb;
    nb = (rows + band_rows - 1) / band_rows;
    const long long warps = (long long)ns * nb * n_pages;
    const unsigned grid = (unsigned)((warps + 3) / 4);
    if (2 * n + 1 <= 8)
        morph_bits_kernel<CLOSE, 8><<<grid, 128, 0, ctx->stream>>>(d_in, in_step, in_page_stride, d_out, out_step, out_page_stride,
                                                                   rows, cols, n1, n2, ns, nb, band_rows, n_pages);
    else
        morph_bits_kernel<CLOSE, 32><<<grid, 128, 0, ctx->stream>>>(d_in, in_step, in_page_stride, d_out, out_step, out_page_stride,
                                                                    rows, cols, n1, n2, ns, nb, band_rows, n_pages);
}

// ---- generic fallback: one separable pass per launch -----------------------------------------
template <int DIR, bool IS_MAX>
__global__ void __launch_bounds__(256)
morph_pass_kernel(const uint8_t* __restrict__ src, size_t src_step, size_t src_page_stride,
                  uint8_t* __restrict__ dst, size_t dst_step, size_t dst_page_stride, int rows, int cols, int n)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= cols) return;
    src += (size_t)blockIdx.z * src_page_stride;
    dst += (size_t)blockIdx.z * dst_page_stride;
    int v = IS_MAX ? 0 : 255;
    if (DIR == 0) {
        const int lo = max(x - n, 0), hi = min(x + n, cols - 1);
        const uint8_t* row = src + (size_t)y * src_step;
        for (int j = lo; j <= hi; ++j) { int u = row[j]; v = IS_MAX ? max(v, u) : min(v, u); }
    } else {
        const int lo = max(y - n, 0), hi = min(y + n, rows - 1);
        for (int j = lo; j <= hi; ++j) { int u = src[(size_t)j * src_step + x]; v = IS_MAX ? max(v, u) : min(v, u); }
    }
    dst[(size_t)y * dst_step + x] = (uint8_t)v;
}

template <bool IS_MAX>
void morph_op(prl_cuda_ctx* ctx, uint8_t* d_mask, uint8_t* d_tmp, int n_pages, int rows, int cols, size_t step,
              size_t page_stride, size_t tmp_step, size_t tmp_page_stride, int n)
{
    dim3 grid((cols + 255) / 256, rows, n_pages);
    {
        prl_launch_scope ls(ctx, FAM_MORPH);
        morph_pass_kernel<0, IS_MAX><<<grid, 256, 0, ctx->stream>>>(d_mask, step, page_stride, d_tmp, tmp_step,
                                                                   tmp_page_stride, rows, cols, n);
    }
    {
        prl_launch_scope ls(ctx, FAM_MORPH);
        morph_pass_kernel<1, IS_MAX><<<grid, 256, 0, ctx->stream>>>(d_tmp, tmp_step, tmp_page_stride, d_mask, step,
                                                                   page_stride, rows, cols, n);
    }
}

}  // namespace

// *flag |= 1 when some pixel is neither 0 nor 255 (then the bit-packed kernels do not apply)
__global__ void __launch_bounds__(256)
not_binary_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, int* __restrict__ flag)
{
    int bad = 0;
    for (int y = blockIdx.x; y < rows; y += gridDim.x) {
        const uint8_t* row = src + (size_t)y * step;
        for (int x = threadIdx.x; x < cols; x += 256) { const int v = row[x]; bad |= (v != 0 && v != 255); }
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

int prl_k_not_binary(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int* d_flag)
{
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(d_flag, 0, sizeof(int), ctx->stream));
    prl_launch_scope ls(ctx, FAM_MORPH);
    not_binary_kernel<<<std::min(rows, 4 * ctx->num_sms), 256, 0, ctx->stream>>>(d_src, step, rows, cols, d_flag);
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

// Out of place: reads d_in, writes d_out (d_in is clobbered by the n > 8 fallback).
int prl_k_morph(prl_cuda_ctx* ctx, uint8_t* d_in, uint8_t* d_out, int n_pages, int rows, int cols, size_t in_step,
                size_t in_page_stride, size_t out_step, size_t out_page_stride, int iters, bool binary)
{
    if (n_pages > 65535 || rows > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    if (iters == 0) {
        for (int p = 0; p < n_pages; ++p)
            PRL_CUDA_TRY(ctx, cudaMemcpy2DAsync(d_out + (size_t)p * out_page_stride, out_step, d_in + (size_t)p * in_page_stride, in_step,
                                                cols, rows, cudaMemcpyDeviceToDevice, ctx->stream));
        return PRL_OK;
    }
    const int n = iters > 0 ? iters : -iters;
    if (binary && !ctx->morph_bytes && n <= 15 && ((((uintptr_t)d_in) | in_step | in_page_stride) & 15) == 0) {
        prl_launch_scope ls(ctx, FAM_MORPH);
        if (iters > 0) launch_bits<true>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n, n);
        else launch_bits<false>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n, n);
        PRL_CUDA_TRY(ctx, cudaGetLastError());
        return PRL_OK;
    }
    if (n <= kMaxN) {
        prl_launch_scope ls(ctx, FAM_MORPH);
        const bool close = iters > 0;
        if (n <= 2) { if (close) launch_fused<true, 4>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n);
                      else launch_fused<false, 4>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n); }
        else if (n <= 4) { if (close) launch_fused<true, 8>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n);
                           else launch_fused<false, 8>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n); }
        else { if (close) launch_fused<true, 16>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n);
               else launch_fused<false, 16>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n); }
        PRL_CUDA_TRY(ctx, cudaGetLastError());
        return PRL_OK;
    }
    // fallback: four single passes ping-ponging d_in <-> d_out, then one copy
    if (iters > 0) {
        morph_op<true>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n);
        morph_op<false>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n);
    } else {
        morph_op<false>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n);
        morph_op<true>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n);
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    for (int p = 0; p < n_pages; ++p)
        PRL_CUDA_TRY(ctx, cudaMemcpy2DAsync(d_out + (size_t)p * out_page_stride, out_step, d_in + (size_t)p * in_page_stride, in_step,
                                            cols, rows, cudaMemcpyDeviceToDevice, ctx->stream));
    return PRL_OK;
}

// One operator alone on binary masks: cv::dilate (dilate = true) or cv::erode with the 3x3 element, n iterations
// (binarizeLocalOtsu.cpp:92).  d_in 16-byte aligned (library scratch), 1 <= n <= 15.
int prl_k_morph_single(prl_cuda_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_pages, int rows, int cols, size_t in_step,
                       size_t in_page_stride, size_t out_step, size_t out_page_stride, int n, bool dilate)
{
    if (n_pages > 65535 || rows > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    if (n < 1 || n > 15 || ((((uintptr_t)d_in) | in_step | in_page_stride) & 15) != 0)
        return prl_set_err(ctx, PRL_E_UNSUPPORTED, "single binary operator: 1 <= n <= 15 and 16-byte aligned input");
    prl_launch_scope ls(ctx, FAM_MORPH);
    if (dilate) launch_bits<true>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n, 0);
    else launch_bits<false>(ctx, d_in, d_out, n_pages, rows, cols, in_step, in_page_stride, out_step, out_page_stride, n, 0);
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
