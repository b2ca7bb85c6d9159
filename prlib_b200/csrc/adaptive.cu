// adaptive.cu -- the adaptive-mean family of src/binarizations: cv::medianBlur and
// cv::adaptiveThreshold (MEAN_C / GAUSSIAN_C) on the device, bit-identical to OpenCV 4.x, behind
//   prl::binarizeNativeAdaptive   binarizeNativeAdaptive.cpp:34-140   (median / Gaussian blur -> adaptiveThreshold(BINARY_INV)
//                                                                       -> invert when the mean is below 128)
//   prl::binarizeAT / AGT          binarizeAT.cpp:33-67, binarizeAGT.cpp:33-58  (medianBlur on the colour image -> BGR2GRAY ->
//                                                                       adaptiveThreshold(MEAN_C / GAUSSIAN_C, BINARY))
//   prl::binarizePureAdaptiveGaussian  binarizePureAdaptiveGaussian.cpp:31-71   (BGR2GRAY -> adaptiveThreshold(GAUSSIAN_C))
// Third-party semantics restated here (pinned against the real cv2 calls by tests/test_adaptive.py):
//   * medianBlur (u8, 1/3/4 channels, odd ksize): the exact median of the ksize x ksize window per channel, BORDER_REPLICATE.
//     Whatever algorithm OpenCV picks (sorting network, O(1) histogram) the median is unique; here: an 8-step radix select.
//   * adaptiveThreshold: mean = boxFilter(normalize, BORDER_REPLICATE) or GaussianBlur on CV_32F (sigma = 0 -> derived from
//     the block size) converted back to u8; dst = tab[src - mean + 255], tab = (i - 255 > -ceil(delta)) ? maxval : 0 for
//     THRESH_BINARY and (i - 255 <= -floor(delta)) ? maxval : 0 for THRESH_BINARY_INV, maxval = saturate_cast<uchar>.
//       - box mean of u8: cvRound(sum / bs^2).  bs is odd, so sum / bs^2 is never a tie and round-half-even, the float32
//         product of OpenCV's SIMD body and the double product of its scalar tail all give floor((2 sum + bs^2) / (2 bs^2))
//         (proof for bs <= 69: the nearest tie is 1 / (2 bs^2) away, float32 rounding moves the product by < 255 * 2^-23).
//       - GaussianBlur on CV_32F is sepFilter2D with float32 coefficients (getGaussianKernel: the tables for 1..9 taps,
//         else exp() normalised in double and rounded to float).  The summation order and the fusing are those of OpenCV's
//         AVX2 build, measured column by column against the cv2 wheel (scripts in DESIGN.md):
//           row pass    s = x[0] k[0];  s = fma(x[i], k[i], s) for columns below (cols & ~3); the scalar tail adds the
//                       first 4 * floor((n - 1) / 4) products unfused (gcc vectorises them) and fuses the rest;
//           column pass s = r[c] k[c];  s = fma(r[c + j] + r[c - j], k[c + j], s) for columns below (cols & ~7), unfused
//                       multiply-then-add in the scalar tail.
//   * bilateralFilter (u8, 1 channel; the optional last step of binarizeNativeAdaptive.cpp:116-134, applied to the 0 / maxval
//     mask): OpenCV 4.13's own C++ path (the wheel's closed IPP routine switched off, like for Otsu), measured against the cv2 wheel:
//       radius = d / 2 (d <= 0: cvRound(1.5 sigmaSpace)), at least 1; BORDER_REFLECT_101; taps (dy, dx) with dy^2 + dx^2 <= radius^2
//       in raster order, the centre among them;
//       colour weight  cw[i] = v_exp(float(i * i) * float(-0.5 / sigmaColor^2)), OpenCV's float32 SIMD exponential (Cephes expf:
//                      range reduction by ln 2 in two pieces, degree-5 polynomial) with unfused multiply-adds (baseline SSE build);
//       space weight   sw = float(exp(double(dy^2 + dx^2) * double(float(-0.5 / sigmaSpace^2))));
//       per tap        w = sw * cw[|v - centre|];  wsum += w;  sum = fma(v, w, sum);      dst = cvRound(sum / wsum), all float32.
//     Identical to the wheel on every input tried for radius != 2, and for radius 2 (d = 4, 5) on every two-level image -- all
//     2^13 neighbourhoods checked, which is everything the reference can feed it; on full-range gray input the wheel's radius-2
//     special case (ring sums, and itself dependent on the CPU dispatch: AVX-512 and baseline builds differ) rounds ~1e-5 of
//     the pixels the other way.
#include "common.cuh"
#include <algorithm>
#include <cmath>
#include <vector>

namespace {

inline size_t r16(size_t v) { return (v + 15) & ~(size_t)15; }
inline size_t r256(size_t v) { return (v + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------------------------------
// medianBlur
// ------------------------------------------------------------------------------------------------
constexpr int kMedTW = 32, kMedTH = 8;

template <int CH>
__global__ void __launch_bounds__(kMedTW * kMedTH)
median_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, int ksize, uint8_t* __restrict__ dst, size_t dstep)
{
    extern __shared__ uint8_t tile[];                       // (kMedTH + ksize - 1) x (kMedTW + ksize - 1) x CH
    const int h = ksize / 2, tw = kMedTW + ksize - 1, th = kMedTH + ksize - 1;
    const int x0 = blockIdx.x * kMedTW - h, y0 = blockIdx.y * kMedTH - h;
    for (int i = threadIdx.x; i < tw * th; i += kMedTW * kMedTH) {
        const int ty = i / tw, tx = i - ty * tw;
        const int sy = min(max(y0 + ty, 0), rows - 1), sx = min(max(x0 + tx, 0), cols - 1);     // BORDER_REPLICATE
        const uint8_t* p = src + (size_t)sy * step + (size_t)sx * CH;
#pragma unroll
        for (int c = 0; c < CH; ++c) tile[i * CH + c] = p[c];
    }
    __syncthreads();
    const int lx = threadIdx.x % kMedTW, ly = threadIdx.x / kMedTW;
    const int x = blockIdx.x * kMedTW + lx, y = blockIdx.y * kMedTH + ly;
    if (x >= cols || y >= rows) return;
    const int r = (ksize * ksize) / 2;                      // 0-based rank of the median
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        // the greatest t with #{v < t} <= r is the r-th smallest value: fix its bits from the top
        int res = 0;
        for (int bit = 7; bit >= 0; --bit) {
            const int cand = res | (1 << bit);
            int cnt = 0;
            for (int dy = 0; dy < ksize; ++dy) {
                const uint8_t* row = tile + ((ly + dy) * tw + lx) * CH + c;
                for (int dx = 0; dx < ksize; ++dx) cnt += row[dx * CH] < cand ? 1 : 0;
            }
            if (cnt <= r) res = cand;
        }
        dst[(size_t)y * dstep + (size_t)x * CH + c] = (uint8_t)res;
    }
}

// ---- kernel sizes 3 and 5: a median selection network instead of the radix select.  Two adjacent output BYTES per thread ride in
// the two 16-bit lanes of a register (VIMNMX.U16x2: one instruction per min or max of both); with interleaved channels the
// neighbour of byte q along x is byte q +- CH, so one kernel serves 1, 3 and 4 channels.  The networks are N. Devillard's
// opt_med9 / opt_med25 (19 and 99 compare-exchanges, median left in wire 4 / 12); tests/test_adaptive.py checks them by the
// 0-1 principle over all 2^9 / 2^25 inputs.  [B200] 5 x 5 on an A4 page: 0.55 ms with the radix select.
#define PRL_MED9_NETWORK \
    CE(1, 2) CE(4, 5) CE(7, 8) CE(0, 1) CE(3, 4) CE(6, 7) CE(1, 2) CE(4, 5) CE(7, 8) CE(0, 3) CE(5, 8) CE(4, 7) \
    CE(3, 6) CE(1, 4) CE(2, 5) CE(4, 7) CE(4, 2) CE(6, 4) CE(4, 2)
#define PRL_MED25_NETWORK \
    CE(0, 1) CE(3, 4) CE(2, 4) CE(2, 3) CE(6, 7) CE(5, 7) CE(5, 6) CE(9, 10) CE(8, 10) CE(8, 9) CE(12, 13) CE(11, 13) \
    CE(11, 12) CE(15, 16) CE(14, 16) CE(14, 15) CE(18, 19) CE(17, 19) CE(17, 18) CE(21, 22) CE(20, 22) CE(20, 21) \
    CE(23, 24) CE(2, 5) CE(3, 6) CE(0, 6) CE(0, 3) CE(4, 7) CE(1, 7) CE(1, 4) CE(11, 14) CE(8, 14) CE(8, 11) \
    CE(12, 15) CE(9, 15) CE(9, 12) CE(13, 16) CE(10, 16) CE(10, 13) CE(20, 23) CE(17, 23) CE(17, 20) CE(21, 24) \
    CE(18, 24) CE(18, 21) CE(19, 22) CE(8, 17) CE(9, 18) CE(0, 18) CE(0, 9) CE(10, 19) CE(1, 19) CE(1, 10) CE(11, 20) \
    CE(2, 20) CE(2, 11) CE(12, 21) CE(3, 21) CE(3, 12) CE(13, 22) CE(4, 22) CE(4, 13) CE(14, 23) CE(5, 23) CE(5, 14) \
    CE(15, 24) CE(6, 24) CE(6, 15) CE(7, 16) CE(7, 19) CE(13, 21) CE(15, 23) CE(7, 13) CE(7, 15) CE(1, 9) CE(3, 11) \
    CE(5, 17) CE(11, 17) CE(9, 17) CE(4, 10) CE(6, 12) CE(7, 14) CE(4, 6) CE(4, 7) CE(12, 14) CE(10, 14) CE(6, 7) \
    CE(10, 12) CE(6, 10) CE(6, 17) CE(12, 17) CE(7, 17) CE(7, 10) CE(12, 18) CE(7, 12) CE(10, 18) CE(12, 20) \
    CE(10, 20) CE(10, 12)

constexpr int kMsTW = 64, kMsTH = 8;          // output bytes per CTA row (two per thread), rows per CTA

template <int K, int CH>
__global__ void __launch_bounds__(kMsTW / 2 * kMsTH)
median_small_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, uint8_t* __restrict__ dst, size_t dstep)
{
    constexpr int H = K / 2, TWB = kMsTW + 2 * H * CH, TH = kMsTH + 2 * H;      // tile: TH rows of TWB bytes
    __shared__ uint8_t tile[TH * TWB];
    const int wbytes = cols * CH;
    const int q0 = blockIdx.x * kMsTW, y0 = blockIdx.y * kMsTH;                 // first output byte / row of the CTA
    for (int i = threadIdx.x; i < TH * TWB; i += kMsTW / 2 * kMsTH) {
        const int ty = i / TWB, tb = i - ty * TWB;
        const int sy = min(max(y0 - H + ty, 0), rows - 1);                      // BORDER_REPLICATE, per pixel
        const int q = q0 - H * CH + tb;                                         // byte position in the row, may be outside
        const int px = q >= 0 ? q / CH : -((-q + CH - 1) / CH);                 // floor(q / CH)
        const int ch = q - px * CH;
        tile[i] = src[(size_t)sy * step + (size_t)min(max(px, 0), cols - 1) * CH + ch];
    }
    __syncthreads();
    const int lx = threadIdx.x % (kMsTW / 2), ly = threadIdx.x / (kMsTW / 2);
    const int q = q0 + 2 * lx, y = y0 + ly;
    if (q >= wbytes || y >= rows) return;
    uint32_t p[K * K];
#pragma unroll
    for (int dy = 0; dy < K; ++dy) {
        const uint8_t* row = tile + (ly + dy) * TWB + 2 * lx;                   // byte q - H*CH of tile row ly + dy
#pragma unroll
        for (int dx = 0; dx < K; ++dx) p[dy * K + dx] = (uint32_t)row[dx * CH] | ((uint32_t)row[dx * CH + 1] << 16);
    }
#define CE(a, b) { const uint32_t lo = __vminu2(p[a], p[b]); p[b] = __vmaxu2(p[a], p[b]); p[a] = lo; }
    uint32_t m;
    if (K == 3) { PRL_MED9_NETWORK m = p[4]; }
    else { PRL_MED25_NETWORK m = p[12]; }
#undef CE
    uint8_t* o = dst + (size_t)y * dstep + q;
    o[0] = (uint8_t)m;
    if (q + 1 < wbytes) o[1] = (uint8_t)(m >> 16);
}

// ------------------------------------------------------------------------------------------------
// adaptiveThreshold
// ------------------------------------------------------------------------------------------------
struct TabArgs { int inv; int idelta; int maxv; };

// dst = tab[src - mean + 255]; counts the pixels set (for the mean < 128 test of binarizeNativeAdaptive.cpp:108)
__device__ __forceinline__ int tab_value(const TabArgs& T, int s, int m)
{
    const int i = s - m;
    return T.inv ? (i <= -T.idelta ? T.maxv : 0) : (i > -T.idelta ? T.maxv : 0);
}

// MEAN_C: box sums by two sliding passes in shared memory, one CTA per tile of TW x TH output pixels
__global__ void __launch_bounds__(256)
box_threshold_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, int bs, int TW, int TH, TabArgs T,
                     uint8_t* __restrict__ dst, size_t dstep, unsigned long long* __restrict__ nset)
{
    extern __shared__ __align__(16) uint8_t smem[];
    const int h = bs / 2, tw = TW + bs - 1, th = TH + bs - 1;
    uint8_t* tile = smem;                                               // th x tw
    unsigned short* hs = reinterpret_cast<unsigned short*>(smem + (((size_t)tw * th + 15) & ~(size_t)15));   // th x TW
    const int x0 = blockIdx.x * TW - h, y0 = blockIdx.y * TH - h;
    for (int i = threadIdx.x; i < tw * th; i += 256) {
        const int ty = i / tw, tx = i - ty * tw;
        tile[i] = src[(size_t)min(max(y0 + ty, 0), rows - 1) * step + min(max(x0 + tx, 0), cols - 1)];
    }
    __syncthreads();
    // horizontal window sums: units of (row, 16 columns)
    const int segs = TW / 16;
    for (int u = threadIdx.x; u < th * segs; u += 256) {
        const int ty = u / segs, sx = (u - ty * segs) * 16;
        const uint8_t* row = tile + ty * tw + sx;
        int s = 0;
        for (int k = 0; k < bs; ++k) s += row[k];
        unsigned short* o = hs + ty * TW + sx;
        o[0] = (unsigned short)s;
        for (int j = 1; j < 16; ++j) { s += row[j + bs - 1] - row[j - 1]; o[j] = (unsigned short)s; }
    }
    __syncthreads();
    // vertical window sums: units of (column, 8 rows)
    const int vsegs = TH / 8;
    const unsigned int area = (unsigned int)bs * bs;
    unsigned int set = 0;
    for (int u = threadIdx.x; u < TW * vsegs; u += 256) {
        const int lx = u % TW, sy = (u / TW) * 8;
        const unsigned short* col = hs + sy * TW + lx;
        unsigned int s = 0;
        for (int k = 0; k < bs; ++k) s += col[k * TW];
        const int x = blockIdx.x * TW + lx;
        for (int j = 0; j < 8; ++j) {
            if (j) s += col[(j + bs - 1) * TW] - col[(j - 1) * TW];
            const int y = blockIdx.y * TH + sy + j;
            if (x < cols && y < rows) {
                const int mean = (int)((2u * s + area) / (2u * area));
                const int v = tab_value(T, tile[(sy + j + h) * tw + lx + h], mean);
                dst[(size_t)y * dstep + x] = (uint8_t)v;
                set += v ? 1u : 0u;
            }
        }
    }
    if (nset != nullptr) {
        set = __reduce_add_sync(0xffffffffu, set);
        if ((threadIdx.x & 31) == 0 && set) atomicAdd(nset, (unsigned long long)set);
    }
}

struct GaussF { int n; int n_unfused; float k[256]; };      // n_unfused: leading terms of the row pass's scalar tail that are not fused

// GAUSSIAN_C, row pass: u8 -> float32 rows
__global__ void __launch_bounds__(256)
gaussf_rows_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, const __grid_constant__ GaussF K,
                   float* __restrict__ tmp, size_t tstep)
{
    __shared__ float kk[256];
    __shared__ uint8_t line[256 + 256];
    const int n = K.n, h = n / 2;
    for (int i = threadIdx.x; i < n; i += 256) kk[i] = K.k[i];
    const int y = blockIdx.y, xb = blockIdx.x * 256;
    const uint8_t* row = src + (size_t)y * step;
    for (int i = threadIdx.x; i < 256 + n - 1; i += 256) line[i] = row[min(max(xb + i - h, 0), cols - 1)];
    __syncthreads();
    const int x = xb + threadIdx.x;
    if (x >= cols) return;
    const uint8_t* p = line + threadIdx.x;
    float s = __fmul_rn((float)p[0], kk[0]);
    if (x < (cols & ~3)) {
        for (int i = 1; i < n; ++i) s = __fmaf_rn((float)p[i], kk[i], s);
    } else {
        const int nu = K.n_unfused;
        for (int i = 1; i <= nu; ++i) s = __fadd_rn(s, __fmul_rn((float)p[i], kk[i]));
        for (int i = nu + 1; i < n; ++i) s = __fmaf_rn((float)p[i], kk[i], s);
    }
    tmp[(size_t)y * tstep + x] = s;
}

// GAUSSIAN_C, column pass + convertTo(u8) + table
__global__ void __launch_bounds__(256)
gaussf_cols_threshold_kernel(const float* __restrict__ tmp, size_t tstep, const uint8_t* __restrict__ src, size_t step, int rows, int cols,
                             const __grid_constant__ GaussF K, TabArgs T, uint8_t* __restrict__ dst, size_t dstep,
                             unsigned long long* __restrict__ nset)
{
    __shared__ float kk[256];
    const int n = K.n, h = n / 2;
    for (int i = threadIdx.x; i < n; i += 256) kk[i] = K.k[i];
    __syncthreads();
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    unsigned int set = 0;
    if (x < cols) {
        const float* c = tmp + x;
        float s = __fmul_rn(c[(size_t)y * tstep], kk[h]);
        const bool fused = x < (cols & ~7);
        for (int j = 1; j <= h; ++j) {
            const float a = c[(size_t)min(y + j, rows - 1) * tstep], b = c[(size_t)max(y - j, 0) * tstep];
            const float pr = __fadd_rn(a, b);
            s = fused ? __fmaf_rn(pr, kk[h + j], s) : __fadd_rn(s, __fmul_rn(pr, kk[h + j]));
        }
        int mean = __float2int_rn(s);                                   // saturate_cast<uchar>(float): cvRound, then clamp
        mean = min(max(mean, 0), 255);
        const int v = tab_value(T, src[(size_t)y * step + x], mean);
        dst[(size_t)y * dstep + x] = (uint8_t)v;
        set = v ? 1u : 0u;
    }
    if (nset != nullptr) {
        set = __reduce_add_sync(0xffffffffu, set);
        if ((threadIdx.x & 31) == 0 && set) atomicAdd(nset, (unsigned long long)set);
    }
}

// GAUSSIAN_C in one kernel for blocks up to 63 taps: a CTA owns 64 x 64 outputs, keeps the u8 tile and its row-pass results in
// shared memory and runs the column pass from there -- the float32 plane of the two kernels above (4 B per pixel written, n times
// that re-read through L2) never exists.  Same operations in the same order per pixel, so the same bits.
constexpr int kGfTW = 64, kGfTH = 64;

__global__ void __launch_bounds__(256)
gaussf_fused_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, const __grid_constant__ GaussF K, TabArgs T,
                    uint8_t* __restrict__ dst, size_t dstep, unsigned long long* __restrict__ nset)
{
    extern __shared__ __align__(16) uint8_t gf_sm[];
    const int n = K.n, h = n / 2;
    const int tw = kGfTW + n - 1, th = kGfTH + n - 1;
    float* kk = reinterpret_cast<float*>(gf_sm);                         // 64 coefficients
    float* rsum = kk + 64;                                               // th x kGfTW row-pass results
    float* tile = rsum + (size_t)th * kGfTW;                             // th x tw pixels as floats (converted once, not per tap), BORDER_REPLICATE
    const int x0 = blockIdx.x * kGfTW, y0 = blockIdx.y * kGfTH;
    for (int i = threadIdx.x; i < n; i += 256) kk[i] = K.k[i];
    for (int ty = threadIdx.x >> 5; ty < th; ty += 8) {                 // a warp per tile row: no division by the run-time tile width
        const uint8_t* srow = src + (size_t)min(max(y0 - h + ty, 0), rows - 1) * step;
        float* trow = tile + ty * tw;
        for (int tx = threadIdx.x & 31; tx < tw; tx += 32) trow[tx] = (float)srow[min(max(x0 - h + tx, 0), cols - 1)];
    }
    __syncthreads();
    // row pass, four adjacent outputs per thread: one pixel load feeds four running sums, the coefficient window slides through
    // registers.  Output j meets tap i = k - j at step k; outside [0, n) its coefficient is 0 and fma(v, 0, s) == s exactly, and the
    // first real step fma(v, k0, 0) is the rounded product OpenCV starts from -- the same bits as one output per thread.
    const int nu = K.n_unfused, cols4 = cols & ~3;
    for (int i = threadIdx.x; i < th * (kGfTW / 4); i += 256) {
        const int ty = i / (kGfTW / 4), tx = (i - ty * (kGfTW / 4)) * 4;
        const float* p = tile + ty * tw + tx;
        float* out = rsum + ty * kGfTW + tx;
        if (x0 + tx < cols4) {                               // cols4 and tx are multiples of 4: all four columns are on this side
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
            for (int k = 0; k < n + 3; ++k) {
                const float v = p[k], c0 = k < n ? kk[k] : 0.f;
                s0 = __fmaf_rn(v, c0, s0); s1 = __fmaf_rn(v, c1, s1); s2 = __fmaf_rn(v, c2, s2); s3 = __fmaf_rn(v, c3, s3);
                c3 = c2; c2 = c1; c1 = c0;
            }
            *reinterpret_cast<float4*>(out) = make_float4(s0, s1, s2, s3);
        } else {
            for (int j = 0; j < 4; ++j) {                    // the scalar tail of OpenCV's row filter (and columns past the image)
                float s = __fmul_rn(p[j], kk[0]);
                for (int k = 1; k <= nu; ++k) s = __fadd_rn(s, __fmul_rn(p[j + k], kk[k]));
                for (int k = nu + 1; k < n; ++k) s = __fmaf_rn(p[j + k], kk[k], s);
                out[j] = s;
            }
        }
    }
    __syncthreads();
    const int lx = threadIdx.x % kGfTW, ly0 = threadIdx.x / kGfTW;       // 4 rows of threads, 16 output rows each
    const int x = x0 + lx;
    const bool fused = x < (cols & ~7);
    unsigned int set = 0;
    for (int ly = ly0; ly < kGfTH; ly += 256 / kGfTW) {
        const int y = y0 + ly;
        if (x < cols && y < rows) {
            const float* c = rsum + (ly + h) * kGfTW + lx;
            float s = __fmul_rn(c[0], kk[h]);
            if (fused) {
#pragma unroll 3
                for (int j = 1; j <= h; ++j) s = __fmaf_rn(__fadd_rn(c[j * kGfTW], c[-j * kGfTW]), kk[h + j], s);
            } else {
                for (int j = 1; j <= h; ++j) s = __fadd_rn(s, __fmul_rn(__fadd_rn(c[j * kGfTW], c[-j * kGfTW]), kk[h + j]));
            }
            int mean = __float2int_rn(s);                                // saturate_cast<uchar>(float): cvRound, then clamp
            mean = min(max(mean, 0), 255);
            const int v = tab_value(T, (int)tile[(ly + h) * tw + lx + h], mean);
            dst[(size_t)y * dstep + x] = (uint8_t)v;
            set += v ? 1u : 0u;
        }
    }
    if (nset != nullptr) {
        set = __reduce_add_sync(0xffffffffu, set);
        if ((threadIdx.x & 31) == 0 && set) atomicAdd(nset, (unsigned long long)set);
    }
}

// inputImageChannels[c] = 255 - inputImageChannels[c] when cv::mean(...)[0] < 128 (binarizeNativeAdaptive.cpp:108-111)
__global__ void __launch_bounds__(256)
invert_if_dark_kernel(uint8_t* __restrict__ img, size_t step, int rows, int cols, int maxv, const unsigned long long* __restrict__ nset)
{
    // mean = maxv * nset / (rows * cols) < 128  <=>  maxv * nset < 128 * rows * cols   (exact in 64-bit integers)
    if ((unsigned long long)maxv * *nset >= 128ull * (unsigned long long)rows * (unsigned long long)cols) return;
    // 16 pixels per thread where the rows are 16-byte aligned
    const int y = blockIdx.y;
    uint8_t* row = img + (size_t)y * step;
    if (((((uintptr_t)img) | step) & 15u) == 0) {
        const int x = (blockIdx.x * 256 + threadIdx.x) * 16;
        if (x + 16 <= cols) {
            uint4 v = *reinterpret_cast<uint4*>(row + x);
            v.x = ~v.x; v.y = ~v.y; v.z = ~v.z; v.w = ~v.w;              // 255 - b == ~b for bytes
            *reinterpret_cast<uint4*>(row + x) = v;
        } else {
            for (int i = x; i < cols; ++i) row[i] = (uint8_t)(255 - row[i]);
        }
    } else {
        for (int x = (blockIdx.x * 256 + threadIdx.x) * 16, e = min(x + 16, cols); x < e; ++x) row[x] = (uint8_t)(255 - row[x]);
    }
}


// ------------------------------------------------------------------------------------------------
// bilateralFilter (u8, one channel)
// ------------------------------------------------------------------------------------------------
constexpr int kBilTW = 32, kBilTH = 8;
struct BilTap { short dy, dx; float sw; };

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while ((unsigned)p >= (unsigned)len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

__global__ void __launch_bounds__(kBilTW * kBilTH)
bilateral_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, int radius, const BilTap* __restrict__ taps, int n_taps,
                 const float* __restrict__ color_weight, uint8_t* __restrict__ dst, size_t dstep)
{
    extern __shared__ __align__(16) uint8_t bil_sm[];
    float* cw = reinterpret_cast<float*>(bil_sm);                       // 256 colour weights
    BilTap* stap = reinterpret_cast<BilTap*>(bil_sm + 256 * sizeof(float));   // the taps (every thread walks the same list)
    uint8_t* tile = bil_sm + 256 * sizeof(float) + (size_t)n_taps * sizeof(BilTap);   // (kBilTH + 2 radius) x (kBilTW + 2 radius) pixels
    const int tw = kBilTW + 2 * radius, th = kBilTH + 2 * radius;
    const int tid = threadIdx.y * kBilTW + threadIdx.x;
    const int x0 = blockIdx.x * kBilTW - radius, y0 = blockIdx.y * kBilTH - radius;
    for (int i = tid; i < 256; i += kBilTW * kBilTH) cw[i] = color_weight[i];
    for (int i = tid; i < n_taps; i += kBilTW * kBilTH) stap[i] = taps[i];
    for (int ty = tid >> 5; ty < th; ty += kBilTW * kBilTH / 32) {      // a warp per tile row
        const uint8_t* srow = src + (size_t)reflect101(y0 + ty, rows) * step;
        for (int tx = tid & 31; tx < tw; tx += 32) tile[ty * tw + tx] = srow[reflect101(x0 + tx, cols)];
    }
    __syncthreads();
    const int x = blockIdx.x * kBilTW + threadIdx.x, y = blockIdx.y * kBilTH + threadIdx.y;
    if (x >= cols || y >= rows) return;
    const uint8_t* c = tile + (threadIdx.y + radius) * tw + threadIdx.x + radius;
    const int v0 = *c;
    float sum = 0.f, wsum = 0.f;
#pragma unroll 4
    for (int k = 0; k < n_taps; ++k) {
        const BilTap t = stap[k];
        const int v = c[t.dy * tw + t.dx];
        const float w = __fmul_rn(t.sw, cw[abs(v - v0)]);
        wsum = __fadd_rn(wsum, w);
        sum = __fmaf_rn((float)v, w, sum);
    }
    dst[(size_t)y * dstep + x] = (uint8_t)__float2int_rn(__fdiv_rn(sum, wsum));
}

// OpenCV's v_exp for float32 (Cephes expf), every multiply-add unfused as in the baseline build that fills the table.
// `volatile` keeps the host compiler from contracting a product into the following sum whatever -march it is given.
float vexp_unfused(float x)
{
    auto madd = [](float a, float b, float c) { volatile float p = a * b; return p + c; };
    x = std::min(std::max(x, -88.3762626647949f), 89.f);
    float fx = madd(x, 1.44269504088896341f, 0.5f);
    const int mm = (int)std::floor(fx);
    fx = (float)mm;
    x = madd(fx, -6.93359375E-1f, x);
    x = madd(fx, 2.12194440E-4f, x);
    volatile float xx = x * x;
    float y = madd(x, 1.9875691500E-4f, 1.3981999507E-3f);
    y = madd(y, x, 8.3334519073E-3f);
    y = madd(y, x, 4.1665795894E-2f);
    y = madd(y, x, 1.6666665459E-1f);
    y = madd(y, x, 5.0000001201E-1f);
    y = madd(y, xx, x);
    volatile float y1 = y + 1.f;
    return y1 * std::ldexp(1.f, mm);
}

}  // namespace

// float32 Gaussian coefficients as cv::getGaussianKernel(n, sigma <= 0, CV_32F) returns them (tables for n <= 9)
int prl_gauss_kernel_float(int n, float* k)
{
    if (n < 1 || n > 255 || (n & 1) == 0) return PRL_E_INVALID;
    static const double t1[] = {1.0}, t3[] = {0.25, 0.5, 0.25}, t5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625},
                        t7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125},
                        t9[] = {4.0 / 256, 13.0 / 256, 30.0 / 256, 51.0 / 256, 60.0 / 256, 51.0 / 256, 30.0 / 256, 13.0 / 256, 4.0 / 256};
    if (n <= 9) {
        const double* t = n == 1 ? t1 : n == 3 ? t3 : n == 5 ? t5 : n == 7 ? t7 : t9;
        for (int i = 0; i < n; ++i) k[i] = (float)t[i];
        return PRL_OK;
    }
    const double s = ((n - 1) * 0.5 - 1) * 0.3 + 0.8, scale2 = -0.5 / (s * s);
    double c[255], sum = 0;
    for (int i = 0; i < n; ++i) { const double x = i - (n - 1) * 0.5; c[i] = exp(scale2 * x * x); sum += c[i]; }
    sum = 1.0 / sum;
    for (int i = 0; i < n; ++i) k[i] = (float)(c[i] * sum);
    return PRL_OK;
}

size_t prl_adaptive_scratch_bytes(int rows, int cols)
{
    return r256(sizeof(unsigned long long) * 4) + r256((size_t)rows * r16((size_t)cols) * sizeof(float));
}

// cv::medianBlur on the device (d_src != d_dst)
int prl_k_median_blur(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int channels, int ksize,
                      uint8_t* d_dst, size_t dst_step)
{
    if (ksize < 3 || (ksize & 1) == 0 || ksize > 63) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "median kernel size must be odd and in [3, 63]");
    if (ksize <= 5 && !ctx->median_legacy) {
        dim3 grid((cols * channels + kMsTW - 1) / kMsTW, (rows + kMsTH - 1) / kMsTH);
        if (grid.y > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "image too tall");
        prl_launch_scope ls(ctx, FAM_ADAPTIVE);
#define PRL_MEDIAN_SMALL(K, C) median_small_kernel<K, C><<<grid, kMsTW / 2 * kMsTH, 0, ctx->stream>>>(d_src, step, rows, cols, d_dst, dst_step)
        if (ksize == 3) { if (channels == 1) PRL_MEDIAN_SMALL(3, 1); else if (channels == 3) PRL_MEDIAN_SMALL(3, 3); else if (channels == 4) PRL_MEDIAN_SMALL(3, 4); else return prl_set_err(ctx, PRL_E_INVALID, "channels must be 1, 3 or 4"); }
        else            { if (channels == 1) PRL_MEDIAN_SMALL(5, 1); else if (channels == 3) PRL_MEDIAN_SMALL(5, 3); else if (channels == 4) PRL_MEDIAN_SMALL(5, 4); else return prl_set_err(ctx, PRL_E_INVALID, "channels must be 1, 3 or 4"); }
#undef PRL_MEDIAN_SMALL
        PRL_CUDA_TRY(ctx, cudaGetLastError());
        return PRL_OK;
    }
    const size_t smem = (size_t)(kMedTW + ksize - 1) * (kMedTH + ksize - 1) * channels;
    dim3 grid((cols + kMedTW - 1) / kMedTW, (rows + kMedTH - 1) / kMedTH);
    if (grid.y > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "image too tall");
    prl_launch_scope ls(ctx, FAM_ADAPTIVE);
#define PRL_MEDIAN(C)                                                                                                              \
    do { auto kfn = median_kernel<C>;                                                                                              \
         PRL_CUDA_TRY(ctx, cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));  \
         kfn<<<grid, kMedTW * kMedTH, smem, ctx->stream>>>(d_src, step, rows, cols, ksize, d_dst, dst_step); } while (0)
    if (channels == 1) PRL_MEDIAN(1);
    else if (channels == 3) PRL_MEDIAN(3);
    else if (channels == 4) PRL_MEDIAN(4);
    else return prl_set_err(ctx, PRL_E_INVALID, "channels must be 1, 3 or 4");
#undef PRL_MEDIAN
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

// cv::adaptiveThreshold on the device (d_src != d_dst).  method 0 = MEAN_C, 1 = GAUSSIAN_C;
// type 0 = THRESH_BINARY, 1 = THRESH_BINARY_INV.  scratch: prl_adaptive_scratch_bytes.  invert_if_dark: the
// `mean < 128 -> 255 - image` step of prl::binarizeNativeAdaptive, decided on the device.
int prl_k_adaptive_threshold(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, double maxval, int method, int type,
                             int block_size, double delta, uint8_t* d_dst, size_t dst_step, void* scratch, bool invert_if_dark)
{
    if (block_size <= 1 || (block_size & 1) == 0 || block_size > 255)
        return prl_set_err(ctx, PRL_E_UNSUPPORTED, "adaptive block size must be odd and in [3, 255]");
    if (rows > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "image too tall");
    unsigned long long* nset = (unsigned long long*)scratch;
    float* tmp = (float*)((uint8_t*)scratch + r256(sizeof(unsigned long long) * 4));
    const size_t tstep = r16((size_t)cols);
    TabArgs T;
    T.inv = type != 0;
    T.idelta = (int)(type == 0 ? ceil(delta) : floor(delta));
    double mv = nearbyint(maxval);                                  // saturate_cast<uchar>(double) = cvRound + clamp
    T.maxv = (int)std::min(std::max(mv, 0.0), 255.0);
    if (maxval < 0) {                                               // "if( maxValue < 0 ) { dst = Scalar(0); return; }"
        PRL_CUDA_TRY(ctx, cudaMemset2DAsync(d_dst, dst_step, 0, cols, rows, ctx->stream));
        T.maxv = 0;
        PRL_CUDA_TRY(ctx, cudaMemsetAsync(nset, 0, sizeof(unsigned long long), ctx->stream));
    } else {
        PRL_CUDA_TRY(ctx, cudaMemsetAsync(nset, 0, sizeof(unsigned long long), ctx->stream));
        if (method == 0) {
            const int TW = block_size <= 63 ? 64 : 32, TH = 32;
            const size_t tile = ((size_t)(TW + block_size - 1) * (TH + block_size - 1) + 15) & ~(size_t)15;
            const size_t smem = tile + (size_t)(TH + block_size - 1) * TW * sizeof(unsigned short);
            dim3 grid((cols + TW - 1) / TW, (rows + TH - 1) / TH);
            if (grid.y > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "image too tall");
            PRL_CUDA_TRY(ctx, cudaFuncSetAttribute(box_threshold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            prl_launch_scope ls(ctx, FAM_ADAPTIVE);
            box_threshold_kernel<<<grid, 256, smem, ctx->stream>>>(d_src, step, rows, cols, block_size, TW, TH, T, d_dst, dst_step, nset);
        } else if (method == 1) {
            // cv::GaussianBlur: a one-pixel-wide / -high image is not filtered along that axis (ksize.width / .height := 1)
            GaussF K, K1;
            K.n = block_size;
            K.n_unfused = 4 * ((block_size - 1) / 4);
            if (prl_gauss_kernel_float(block_size, K.k) != PRL_OK) return prl_set_err(ctx, PRL_E_INVALID, "bad Gaussian block size");
            for (int i = block_size; i < 256; ++i) K.k[i] = 0.f;     // the coefficients travel as a kernel parameter (1 KB)
            K1 = K; K1.n = 1; K1.n_unfused = 0; K1.k[0] = 1.f;
            const GaussF& Kx = cols > 1 ? K : K1;
            const GaussF& Ky = rows > 1 ? K : K1;
            if (block_size <= 63 && rows > 1 && cols > 1 && !ctx->gauss_legacy) {
                const size_t smem = 64 * sizeof(float) + (size_t)(kGfTH + block_size - 1) * kGfTW * sizeof(float) +
                                    (size_t)(kGfTH + block_size - 1) * (kGfTW + block_size - 1) * sizeof(float);
                dim3 grid((cols + kGfTW - 1) / kGfTW, (rows + kGfTH - 1) / kGfTH);
                PRL_CUDA_TRY(ctx, cudaFuncSetAttribute(gaussf_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                prl_launch_scope ls(ctx, FAM_ADAPTIVE);
                gaussf_fused_kernel<<<grid, 256, smem, ctx->stream>>>(d_src, step, rows, cols, K, T, d_dst, dst_step, nset);
            } else {
                dim3 grid((cols + 255) / 256, rows);
                {
                    prl_launch_scope ls(ctx, FAM_ADAPTIVE);
                    gaussf_rows_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, step, rows, cols, Kx, tmp, tstep);
                }
                {
                    prl_launch_scope ls(ctx, FAM_ADAPTIVE);
                    gaussf_cols_threshold_kernel<<<grid, 256, 0, ctx->stream>>>(tmp, tstep, d_src, step, rows, cols, Ky, T, d_dst, dst_step, nset);
                }
            }
        } else {
            return prl_set_err(ctx, PRL_E_INVALID, "Unknown/unsupported adaptive threshold method");
        }
    }
    if (invert_if_dark) {
        prl_launch_scope ls(ctx, FAM_ADAPTIVE);
        invert_if_dark_kernel<<<dim3((cols + 4095) / 4096, rows), 256, 0, ctx->stream>>>(d_dst, dst_step, rows, cols, T.maxv, nset);
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

size_t prl_bilateral_scratch_bytes(int d, double sigma_space)
{
    int radius = d > 0 ? d / 2 : (int)std::nearbyint(sigma_space * 1.5);
    radius = std::max(radius, 1);
    return r256(256 * sizeof(float)) + r256((size_t)(2 * radius + 1) * (2 * radius + 1) * sizeof(BilTap));
}

// cv::bilateralFilter(src, dst, d, sigmaColor, sigmaSpace) on the device, u8, one channel (d_src != d_dst).
// scratch: prl_bilateral_scratch_bytes (receives the two weight tables).
int prl_k_bilateral(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int d, double sigma_color, double sigma_space,
                    uint8_t* d_dst, size_t dst_step, void* scratch)
{
    if (sigma_color <= 0) sigma_color = 1;
    if (sigma_space <= 0) sigma_space = 1;
    int radius = d > 0 ? d / 2 : (int)std::nearbyint(sigma_space * 1.5);             // cvRound
    radius = std::max(radius, 1);
    if (radius > 32) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "bilateral filter diameters above 65 are not supported");
    if ((rows + kBilTH - 1) / kBilTH > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "image too tall");
    const float gc = (float)(-0.5 / (sigma_color * sigma_color)), gs = (float)(-0.5 / (sigma_space * sigma_space));
    float cw[256];
    for (int i = 0; i < 256; ++i) { volatile float a = (float)(i * i) * gc; cw[i] = vexp_unfused(a); }
    std::vector<BilTap> taps;
    for (int dy = -radius; dy <= radius; ++dy)
        for (int dx = -radius; dx <= radius; ++dx) {
            const int r2 = dy * dy + dx * dx;
            if (r2 > radius * radius) continue;
            taps.push_back(BilTap{(short)dy, (short)dx, (float)std::exp((double)r2 * (double)gs)});
        }
    float* d_cw = (float*)scratch;
    BilTap* d_taps = (BilTap*)((char*)scratch + r256(256 * sizeof(float)));
    // the tables are a few KB of pageable memory: the copies return once staged, so the locals may go out of scope
    PRL_CUDA_TRY(ctx, cudaMemcpyAsync(d_cw, cw, sizeof(cw), cudaMemcpyHostToDevice, ctx->stream));
    PRL_CUDA_TRY(ctx, cudaMemcpyAsync(d_taps, taps.data(), taps.size() * sizeof(BilTap), cudaMemcpyHostToDevice, ctx->stream));
    const size_t smem = 256 * sizeof(float) + taps.size() * sizeof(BilTap) + (size_t)(kBilTW + 2 * radius) * (kBilTH + 2 * radius);
    PRL_CUDA_TRY(ctx, cudaFuncSetAttribute(bilateral_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    prl_launch_scope ls(ctx, FAM_ADAPTIVE);
    bilateral_kernel<<<dim3((cols + kBilTW - 1) / kBilTW, (rows + kBilTH - 1) / kBilTH), dim3(kBilTW, kBilTH), smem, ctx->stream>>>(
        d_src, step, rows, cols, radius, d_taps, (int)taps.size(), d_cw, d_dst, dst_step);
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
