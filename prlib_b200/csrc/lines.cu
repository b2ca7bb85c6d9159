// lines.cu -- prl::removeLines (src/removeLines.cpp:30-77, a Global-Otsu caller: SURVEY.md section 8, row F4) on the
// device:  bw = threshold(~gray, OTSU) ; horizontal = open(bw, 1 x cols/50) ; vertical = open(bw, rows/50 x 1) ;
// out = ~(bw - horizontal - vertical).
// After the Otsu threshold everything is binary, so the image is packed to one bit per pixel (1.1 MB for an A4 page)
// and the long one-dimensional erosions / dilations become AND / OR of shifted copies with window doubling:
// a one-sided window of 2^k pixels costs k passes, any length one more; cv::erode and cv::dilate use the same offsets
// [-L/2, L-1-L/2] (border ignored), i.e. a window of L/2 + 1 pixels towards the origin combined with one of
// L - L/2 pixels away from it.
#include "common.cuh"

namespace {

// bits: rows x wpr 32-bit words, bit i of word w = pixel 32 w + i, padding bits undefined (masked on load)
struct BitImg { const uint32_t* p; int rows, cols, wpr; };

__device__ __forceinline__ uint32_t bit_word(const BitImg& B, int y, int j, bool neutral_one)
{
    const uint32_t neutral = neutral_one ? 0xffffffffu : 0u;
    if (y < 0 || y >= B.rows || j < 0 || j >= B.wpr) return neutral;
    uint32_t v = B.p[(size_t)y * B.wpr + j];
    const int valid = B.cols - 32 * j;                      // valid bits of this word
    if (valid < 32) { const uint32_t m = (1u << valid) - 1u; v = neutral_one ? (v | ~m) : (v & m); }
    return v;
}

// word w of row y of the image moved by (sx, sy): result pixel (x, y) = source pixel (x + sx, y + sy), outside = neutral
__device__ __forceinline__ uint32_t shifted_word(const BitImg& B, int y, int w, int sx, int sy, bool neutral_one)
{
    const int bitpos = 32 * w + sx;
    const int q = bitpos >= 0 ? bitpos >> 5 : -((-bitpos + 31) >> 5);
    const int r = bitpos - 32 * q;
    const uint32_t lo = bit_word(B, y + sy, q, neutral_one);
    if (r == 0) return lo;
    return __funnelshift_r(lo, bit_word(B, y + sy, q + 1, neutral_one), r);
}

// MODE 0: out = in OP shift(in, sx, sy)   (window doubling step);   MODE 1: out = shift(in, sx, sy)   (anchor)
template <bool IS_AND, int MODE>
__global__ void __launch_bounds__(256)
bits_pass_kernel(BitImg B, int sx, int sy, uint32_t* __restrict__ out)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (w >= B.wpr) return;
    const uint32_t s = shifted_word(B, y, w, sx, sy, IS_AND);
    uint32_t v;
    if (MODE == 0) { const uint32_t a = bit_word(B, y, w, IS_AND); v = IS_AND ? (a & s) : (a | s); }
    else v = s;
    out[(size_t)y * B.wpr + w] = v;
}

// bits of (255 - gray > thr): the Otsu threshold of the inverted image, cv::threshold(~gray, bw, 255, 255, BINARY | OTSU)
__global__ void __launch_bounds__(256)
inv_threshold_bits_kernel(const uint8_t* __restrict__ gray, size_t step, int rows, int cols, int wpr, const int32_t* __restrict__ thr,
                          uint32_t* __restrict__ bits)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (w >= wpr) return;
    const int t = thr[0];
    const uint8_t* p = gray + (size_t)y * step + 32 * w;
    const int n = min(32, cols - 32 * w);
    uint32_t v = 0;
    if (n == 32 && (((uintptr_t)p) & 15) == 0) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p + 16));
        const uint32_t ws[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) v |= (uint32_t)((255 - (int)((ws[k] >> (8 * i)) & 0xffu)) > t) << (4 * k + i);
    } else {
        for (int i = 0; i < n; ++i) v |= (uint32_t)((255 - (int)p[i]) > t) << i;
    }
    bits[(size_t)y * wpr + w] = v;
}

__global__ void __launch_bounds__(256)
invert_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, uint8_t* __restrict__ dst, size_t dstep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < cols) dst[(size_t)y * dstep + x] = (uint8_t)(255 - src[(size_t)y * step + x]);
}

// out = ~(bw - horizontal - vertical): 0 where bw & ~h & ~v, else 255
__global__ void __launch_bounds__(256)
lines_combine_kernel(const uint32_t* __restrict__ bw, const uint32_t* __restrict__ hz, const uint32_t* __restrict__ vt, int rows, int cols,
                     int wpr, uint8_t* __restrict__ dst, size_t dstep)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (w >= wpr) return;
    const size_t i = (size_t)y * wpr + w;
    const uint32_t keep = bw[i] & ~hz[i] & ~vt[i];
    uint8_t* o = dst + (size_t)y * dstep + 32 * w;
    const int n = min(32, cols - 32 * w);
    for (int k = 0; k < n; ++k) o[k] = ((keep >> k) & 1u) ? 0 : 255;
}

template <bool IS_AND>
__global__ void __launch_bounds__(256)
bits_combine_kernel(const uint32_t* __restrict__ a, const uint32_t* __restrict__ b, uint32_t* __restrict__ out, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = IS_AND ? (a[i] & b[i]) : (a[i] | b[i]);
}

// one-sided window of n pixels starting at the pixel itself and running in direction dir (+1 / -1) along x or y:
// doubling passes ping-pong between bufX and bufY; returns where the result lives (src itself for n == 1)
template <bool IS_AND>
const uint32_t* dir_window(prl_cuda_ctx* ctx, const uint32_t* src, uint32_t* bufX, uint32_t* bufY, int rows, int cols, int wpr,
                           int n, int dir, bool along_x)
{
    dim3 grid((wpr + 255) / 256, rows);
    const uint32_t* cur = src;
    uint32_t* nxt = bufX;
    auto pass = [&](int s) {
        BitImg B{cur, rows, cols, wpr};
        prl_launch_scope ls(ctx, FAM_LINES);
        bits_pass_kernel<IS_AND, 0><<<grid, 256, 0, ctx->stream>>>(B, along_x ? dir * s : 0, along_x ? 0 : dir * s, nxt);
        cur = nxt; nxt = (nxt == bufX) ? bufY : bufX;
    };
    int len = 1;
    while (2 * len <= n) { pass(len); len *= 2; }
    if (len < n) pass(n - len);
    return cur;
}

// cv::erode (IS_AND) / cv::dilate with a 1 x L or L x 1 rectangle anchored at L / 2: offsets [-a, L - 1 - a], a = L / 2,
// pixels outside the image ignored.  = (window of a + 1 pixels towards the origin) OP (window of L - a pixels away from it),
// each one-sided so that the border rule holds on both sides.  src -> dst; four scratch images.
template <bool IS_AND>
void line_op(prl_cuda_ctx* ctx, const uint32_t* src, uint32_t* dst, uint32_t* s0, uint32_t* s1, uint32_t* s2, uint32_t* s3,
             int rows, int cols, int wpr, int L, bool along_x)
{
    const int a = L / 2;
    const uint32_t* back = dir_window<IS_AND>(ctx, src, s0, s1, rows, cols, wpr, a + 1, -1, along_x);
    const uint32_t* fwd = dir_window<IS_AND>(ctx, src, s2, s3, rows, cols, wpr, L - a, +1, along_x);
    const size_t n = (size_t)rows * wpr;
    prl_launch_scope ls(ctx, FAM_LINES);
    bits_combine_kernel<IS_AND><<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(back, fwd, dst, n);
}

}  // namespace

size_t prl_lines_scratch_bytes(int rows, int cols)
{
    const size_t img = ((size_t)rows * ((cols + 31) / 32) * 4 + 255) & ~(size_t)255;
    return 8 * img + 256 + (((size_t)rows * ((cols + 15) & ~15) + 255) & ~(size_t)255);
}

// d_gray: rows x cols u8 in HBM -> d_dst; scratch of prl_lines_scratch_bytes
int prl_k_remove_lines(prl_cuda_ctx* ctx, const uint8_t* d_gray, int rows, int cols, size_t step, uint8_t* d_dst, size_t dst_step,
                       void* scratch)
{
    if (rows > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    const int Lh = cols / 50, Lv = rows / 50;                // removeLines.cpp:51,61
    if (Lh < 1 || Lv < 1)
        return prl_set_err(ctx, PRL_E_EMPTY_ROI, "removeLines: rows / 50 or cols / 50 is 0 -- the reference's zero-sized structuring "
                                                 "element makes cv::erode assert (cv::Exception)");
    const int wpr = (cols + 31) / 32;
    const size_t img = ((size_t)rows * wpr * 4 + 255) & ~(size_t)255;
    uint8_t* b = (uint8_t*)scratch;
    uint32_t* bw = (uint32_t*)b; uint32_t* hz = (uint32_t*)(b + img); uint32_t* vt = (uint32_t*)(b + 2 * img);
    uint32_t* er = (uint32_t*)(b + 3 * img);
    uint32_t* s0 = (uint32_t*)(b + 4 * img); uint32_t* s1 = (uint32_t*)(b + 5 * img);
    uint32_t* s2 = (uint32_t*)(b + 6 * img); uint32_t* s3 = (uint32_t*)(b + 7 * img);
    int32_t* d_thr = (int32_t*)(b + 8 * img);
    uint8_t* inv = b + 8 * img + 256;
    const size_t istep = ((size_t)cols + 15) & ~(size_t)15;
    {
        prl_launch_scope ls(ctx, FAM_LINES);
        invert_kernel<<<dim3((cols + 255) / 256, rows), 256, 0, ctx->stream>>>(d_gray, step, rows, cols, inv, istep);
    }
    int rc = prl_k_otsu_global(ctx, inv, 1, rows, cols, istep, istep * rows, 255.0, nullptr, 0, 0, d_thr, false); if (rc) return rc;
    dim3 grid((wpr + 255) / 256, rows);
    {
        prl_launch_scope ls(ctx, FAM_LINES);
        inv_threshold_bits_kernel<<<grid, 256, 0, ctx->stream>>>(d_gray, step, rows, cols, wpr, d_thr, bw);
    }
    // horizontal = dilate(erode(bw, 1 x Lh))  (:54-58), vertical = dilate(erode(bw, Lv x 1))  (:64-68)
    line_op<true>(ctx, bw, er, s0, s1, s2, s3, rows, cols, wpr, Lh, true);
    line_op<false>(ctx, er, hz, s0, s1, s2, s3, rows, cols, wpr, Lh, true);
    line_op<true>(ctx, bw, er, s0, s1, s2, s3, rows, cols, wpr, Lv, false);
    line_op<false>(ctx, er, vt, s0, s1, s2, s3, rows, cols, wpr, Lv, false);
    {
        prl_launch_scope ls(ctx, FAM_LINES);
        lines_combine_kernel<<<grid, 256, 0, ctx->stream>>>(bw, hz, vt, rows, cols, wpr, d_dst, dst_step);
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
