// edges.cu -- the edge-detection front-end of prl::binarizeLocalOtsu:
// CannyEdgeDetection (src/imageLibCommon.cpp:244-324) followed by the dilation of binarizeLocalOtsu.cpp:92, i.e.
//   cv::GaussianBlur(gray, k x k, sigma 0) -> Otsu threshold VALUE of the blurred image -> cv::Canny(blurred,
//   lowerCoeff * upper, upper = upperCoeff * otsu) -> closing / opening by CannyMorphIters -> dilate x 3,
// on the device, bit-identical to OpenCV 4.x's 8-bit paths (third-party; semantics pinned against cv2 4.13 by
// tests/test_edges.py):
//  * GaussianBlur on CV_8U is fixed point: 8.8 coefficients (error-diffusion rounding from the tails to the centre,
//    which takes the remainder to 256), row pass exact in 16 bits, column pass exact in 32 bits, one rounding
//    (+ 0x8000) >> 16; BORDER_REFLECT_101.
//  * Canny (aperture 3, L1 norm): Sobel with BORDER_REPLICATE, magnitude |dx| + |dy| with a zero frame, low / high =
//    floor of the thresholds, direction test on |dy| << 15 against tan(22.5) = 13573 and tan(67.5), neighbours
//    compared with > on one side and >= on the other (horizontal, vertical) or > on both (diagonals: up-left /
//    down-right when dx and dy have the same sign), hysteresis = the 8-connected components of the surviving
//    pixels (m > low) that contain a pixel with m > high.
// The hysteresis is a union-find labelling (atomicMin label equivalence) instead of OpenCV's stack walk: same set.
#include "common.cuh"
#include <cmath>
#include <algorithm>

namespace {

constexpr int kMaxGauss = 63;
struct GaussK { int n; int k[kMaxGauss]; };

__device__ __forceinline__ int reflect101(int i, int n)
{
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

// row pass: u8 -> 8.8 fixed point (exact: the coefficients sum to 256)
__global__ void __launch_bounds__(256)
gauss_rows_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, GaussK K, uint16_t* __restrict__ tmp, size_t tstep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const uint8_t* row = src + (size_t)y * step;
    const int r = K.n >> 1;
    uint32_t acc = 0;
    if (x >= r && x + r < cols) {
        for (int t = 0; t < K.n; ++t) acc += (uint32_t)row[x + t - r] * (uint32_t)K.k[t];
    } else {
        for (int t = 0; t < K.n; ++t) acc += (uint32_t)row[reflect101(x + t - r, cols)] * (uint32_t)K.k[t];
    }
    tmp[(size_t)y * tstep + x] = (uint16_t)acc;
}

// column pass: 8.8 -> 16.16 -> u8
__global__ void __launch_bounds__(256)
gauss_cols_kernel(const uint16_t* __restrict__ tmp, size_t tstep, int rows, int cols, GaussK K, uint8_t* __restrict__ dst, size_t dstep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int r = K.n >> 1;
    uint32_t acc = 0;
    if (y >= r && y + r < rows) {
        for (int t = 0; t < K.n; ++t) acc += (uint32_t)tmp[(size_t)(y + t - r) * tstep + x] * (uint32_t)K.k[t];
    } else {
        for (int t = 0; t < K.n; ++t) acc += (uint32_t)tmp[(size_t)reflect101(y + t - r, rows) * tstep + x] * (uint32_t)K.k[t];
    }
    const uint32_t v = (acc + 0x8000u) >> 16;
    dst[(size_t)y * dstep + x] = (uint8_t)min(v, 255u);
}

// ---- Canny: Sobel + magnitude + non-maximum suppression -> class map (0 none, 1 weak, 2 strong) --------
constexpr int kCT = 32, kCH = 16;         // output tile

__global__ void __launch_bounds__(kCT * kCH)
canny_nms_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, const int32_t* __restrict__ otsu_thr,
                 double upper_coeff, double lower_coeff, double fixed_low, double fixed_high, uint8_t* __restrict__ cls, size_t cstep)
{
    __shared__ uint8_t sp[kCH + 4][kCT + 4];          // pixels, halo 2 (replicated at the image border)
    __shared__ int smag[kCH + 2][kCT + 2];            // magnitudes, halo 1 (zero outside the image)
    __shared__ short sdx[kCH][kCT], sdy[kCH][kCT];
    const int tx = threadIdx.x % kCT, ty = threadIdx.x / kCT;
    const int x0 = blockIdx.x * kCT, y0 = blockIdx.y * kCH;
    // thresholds: upper = upperCoeff * otsu, lower = lowerCoeff * upper (imageLibCommon.cpp:299-303), or given
    double lo_t = fixed_low, hi_t = fixed_high;
    if (otsu_thr) { hi_t = __dmul_rn(upper_coeff, (double)otsu_thr[0]); lo_t = __dmul_rn(lower_coeff, hi_t); }
    if (lo_t > hi_t) { const double t = lo_t; lo_t = hi_t; hi_t = t; }
    const int low = (int)floor(lo_t), high = (int)floor(hi_t);

    for (int i = threadIdx.x; i < (kCH + 4) * (kCT + 4); i += kCT * kCH) {
        const int ly = i / (kCT + 4), lx = i - ly * (kCT + 4);
        const int gy = min(max(y0 + ly - 2, 0), rows - 1), gx = min(max(x0 + lx - 2, 0), cols - 1);
        sp[ly][lx] = src[(size_t)gy * step + gx];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < (kCH + 2) * (kCT + 2); i += kCT * kCH) {
        const int ly = i / (kCT + 2), lx = i - ly * (kCT + 2);
        const int gy = y0 + ly - 1, gx = x0 + lx - 1;
        int m = 0;
        if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) {
            // 3x3 around sp[ly + 1][lx + 1]
            const int a = sp[ly][lx], b = sp[ly][lx + 1], c = sp[ly][lx + 2];
            const int d = sp[ly + 1][lx], f = sp[ly + 1][lx + 2];
            const int g = sp[ly + 2][lx], h = sp[ly + 2][lx + 1], k = sp[ly + 2][lx + 2];
            const int dx = (c + 2 * f + k) - (a + 2 * d + g);
            const int dy = (g + 2 * h + k) - (a + 2 * b + c);
            m = abs(dx) + abs(dy);
            if (ly >= 1 && ly <= kCH && lx >= 1 && lx <= kCT) { sdx[ly - 1][lx - 1] = (short)dx; sdy[ly - 1][lx - 1] = (short)dy; }
        }
        smag[ly][lx] = m;
    }
    __syncthreads();
    const int gx = x0 + tx, gy = y0 + ty;
    if (gx >= cols || gy >= rows) return;
    const int m = smag[ty + 1][tx + 1];
    uint8_t out = 0;
    if (m > low) {
        const int xs = sdx[ty][tx], ys = sdy[ty][tx];
        const int x = abs(xs), y = abs(ys) << 15;
        const int tg22x = x * 13573;
        bool is_max;
        if (y < tg22x) is_max = m > smag[ty + 1][tx] && m >= smag[ty + 1][tx + 2];
        else {
            const int tg67x = tg22x + (x << 16);
            if (y > tg67x) is_max = m > smag[ty][tx + 1] && m >= smag[ty + 2][tx + 1];
            else if ((xs ^ ys) < 0) is_max = m > smag[ty][tx + 2] && m > smag[ty + 2][tx];      // up-right / down-left
            else is_max = m > smag[ty][tx] && m > smag[ty + 2][tx + 2];                          // up-left / down-right
        }
        if (is_max) out = m > high ? 2 : 1;
    }
    cls[(size_t)gy * cstep + gx] = out;
}

// ---- hysteresis: union-find over the surviving pixels, 8-connectivity --------------------------------
__device__ __forceinline__ int uf_find(const int* __restrict__ L, int x)
{
    int p = L[x];
    while (p != x) { x = p; p = L[x]; }
    return x;
}
__device__ __forceinline__ void uf_union(int* L, int a, int b)
{
    for (;;) {
        a = uf_find(L, a); b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }      // a > b: hang a under b
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;                                             // somebody re-rooted a meanwhile: merge that root too
    }
}

__global__ void __launch_bounds__(256)
ccl_init_kernel(const uint8_t* __restrict__ cls, size_t cstep, int rows, int cols, int* __restrict__ L)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int p = y * cols + x;
    L[p] = cls[(size_t)y * cstep + x] ? p : -1;
}

__global__ void __launch_bounds__(256)
ccl_merge_kernel(const uint8_t* __restrict__ cls, size_t cstep, int rows, int cols, int* __restrict__ L)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const uint8_t* row = cls + (size_t)y * cstep;
    if (!row[x]) return;
    const int p = y * cols + x;
    if (x > 0 && row[x - 1]) uf_union(L, p, p - 1);
    if (y > 0) {
        const uint8_t* up = row - cstep;
        if (up[x]) uf_union(L, p, p - cols);
        if (x > 0 && up[x - 1]) uf_union(L, p, p - cols - 1);
        if (x + 1 < cols && up[x + 1]) uf_union(L, p, p - cols + 1);
    }
}

// flag[root] = 1 for every component that holds a strong pixel
__global__ void __launch_bounds__(256)
ccl_mark_kernel(const uint8_t* __restrict__ cls, size_t cstep, int rows, int cols, const int* __restrict__ L, uint8_t* __restrict__ flag)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    if (cls[(size_t)y * cstep + x] == 2) flag[uf_find(L, y * cols + x)] = 1;
}

__global__ void __launch_bounds__(256)
ccl_emit_kernel(const uint8_t* __restrict__ cls, size_t cstep, int rows, int cols, const int* __restrict__ L,
                const uint8_t* __restrict__ flag, uint8_t* __restrict__ dst, size_t dstep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    uint8_t out = 0;
    if (cls[(size_t)y * cstep + x]) out = flag[uf_find(L, y * cols + x)] ? 255 : 0;
    dst[(size_t)y * dstep + x] = out;
}

// ---- contours -> rectangles: cv::findContours(RETR_EXTERNAL) + cv::boundingRect without the contours ------------
// RETR_EXTERNAL keeps one outer border per 8-connected component of non-zero pixels that does not lie inside a hole
// of another component; the rectangle loop only needs its bounding box.  Topologically (Suzuki-Abe, 8-connected
// 1-pixels / 4-connected 0-pixels): a component is top-level iff one of its pixels has, in its 4-neighbourhood, a
// 0-pixel of the background component that reaches the image frame (or the frame itself).  So: label 1-pixels with
// 8-connectivity and 0-pixels with 4-connectivity in ONE union-find forest (plus a node for the frame), then every
// border pixel of a component votes "top-level" and stretches its root's bounding box.
// Rows are labelled by runs first (no atomics: every pixel points at the start of its horizontal run), so the
// atomicMin unions only happen once per vertically overlapping pair of runs.
__device__ __forceinline__ int uf_find_c(int* L, int x)
{
    for (;;) {
        const int p = L[x];
        if (p == x) return x;
        const int gp = L[p];
        if (gp != p) atomicMin(&L[x], gp);                   // path halving; labels only ever decrease
        x = p;
    }
}
__device__ __forceinline__ void uf_union_c(int* L, int a, int b)
{
    for (;;) {
        a = uf_find_c(L, a); b = uf_find_c(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

// one CTA per row: L[y * cols + x] = y * cols + (start of the horizontal run of equal class that holds x)
__global__ void __launch_bounds__(256)
ccl_runs_kernel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, int* __restrict__ L)
{
    __shared__ int wmax[8];
    const int y = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint8_t* row = E + (size_t)y * estep;
    const int seg = (cols + 255) / 256, x0 = tid * seg, x1 = min(x0 + seg, cols);
    int last = -1;                                            // last run start inside this thread's segment
    for (int x = x0; x < x1; ++x)
        if (x == 0 || (row[x] != 0) != (row[x - 1] != 0)) last = x;
    // exclusive running maximum over the threads to the left
    int v = last;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v = max(v, t); }
    if (lane == 31) wmax[wid] = v;
    __syncthreads();
    int carry = -1;
    for (int k = 0; k < wid; ++k) carry = max(carry, wmax[k]);
    int prev = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) prev = -1;
    int cur = max(carry, prev);
    for (int x = x0; x < x1; ++x) {
        if (x == 0 || (row[x] != 0) != (row[x - 1] != 0)) cur = x;
        L[y * cols + x] = y * cols + cur;
    }
    if (y == 0 && tid == 0) L[rows * cols] = rows * cols;     // the frame node
}

__global__ void __launch_bounds__(256)
ccl_vmerge_kernel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, int* __restrict__ L)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const uint8_t* row = E + (size_t)y * estep;
    const bool c = row[x] != 0;
    const bool cw = x > 0 ? row[x - 1] != 0 : !c;             // "different class" at the row start
    const int p = y * cols + x;
    if (y > 0) {
        const uint8_t* up = row - estep;
        const bool cn = up[x] != 0;
        const bool cnw = x > 0 ? up[x - 1] != 0 : !cn;
        if (cn == c) {
            if (cw != c || cnw != cn) uf_union_c(L, p, p - cols);       // first column of this overlap of two runs
        } else if (c) {
            // 8-connectivity of the 1-pixels: diagonal neighbours that no 4-neighbour already bridges
            if (x > 0 && cnw && !cw) uf_union_c(L, p, p - cols - 1);
            if (x + 1 < cols && up[x + 1] != 0) uf_union_c(L, p, p - cols + 1);
        }
    }
    if (!c) {
        // 0-pixels on the image border belong to the outer background (findContours works on a zero-framed copy)
        const bool run_start = cw != c;
        if (((y == 0 || y == rows - 1) && run_start) || x == 0 || x == cols - 1) uf_union_c(L, p, rows * cols);
    }
}

__global__ void __launch_bounds__(256)
ccl_border_kernel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, int* __restrict__ L, uint8_t* __restrict__ top,
                  int* __restrict__ bx0, int* __restrict__ by0, int* __restrict__ bx1, int* __restrict__ by1)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const uint8_t* row = E + (size_t)y * estep;
    if (!row[x]) return;
    // 4-neighbours: outside the image (= frame), 0-pixel, or 1-pixel
    const bool oW = x == 0, oE = x == cols - 1, oN = y == 0, oS = y == rows - 1;
    const uint8_t* up = row - estep;
    const uint8_t* dn = row + estep;
    const bool zW = !oW && !row[x - 1], zE = !oE && !row[x + 1], zN = !oN && !up[x], zS = !oS && !dn[x];
    if (!(oW | oE | oN | oS | zW | zE | zN | zS)) return;     // interior pixel: neither a vote nor a bounding-box extreme
    const int p = y * cols + x, root = uf_find_c(L, p);
    bool outer = oW | oE | oN | oS;
    if (!outer) {
        const int fr = uf_find_c(L, rows * cols);
        outer = (zW && uf_find_c(L, p - 1) == fr) || (zE && uf_find_c(L, p + 1) == fr) ||
                (zN && uf_find_c(L, p - cols) == fr) || (zS && uf_find_c(L, p + cols) == fr);
    }
    if (outer && !top[root]) top[root] = 1;
    // look before the atomic: after the first few pixels of a component almost none still stretches its box
    if (x < bx0[root]) atomicMin(&bx0[root], x);
    if (y < by0[root]) atomicMin(&by0[root], y);
    if (x > bx1[root]) atomicMax(&bx1[root], x);
    if (y > by1[root]) atomicMax(&by1[root], y);
}

__global__ void __launch_bounds__(256)
ccl_rects_kernel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, const int* __restrict__ L, const uint8_t* __restrict__ top,
                 const int* __restrict__ bx0, const int* __restrict__ by0, const int* __restrict__ bx1, const int* __restrict__ by1,
                 int* __restrict__ count, int32_t* __restrict__ xywh, int cap)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const int p = y * cols + x;
    if (!E[(size_t)y * estep + x] || L[p] != p || !top[p]) return;
    const int slot = atomicAdd(count, 1);
    if (slot < cap) {
        xywh[4 * slot] = bx0[p]; xywh[4 * slot + 1] = by0[p];
        xywh[4 * slot + 2] = bx1[p] - bx0[p] + 1; xywh[4 * slot + 3] = by1[p] - by0[p] + 1;
    }
}

inline size_t r16(size_t v) { return (v + 15) & ~(size_t)15; }
inline size_t r256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

// The 8.8 fixed-point kernel OpenCV builds for CV_8U (getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED,
// third-party): exact small tables for n <= 7 with sigma <= 0, else exp(-x^2 / (2 sigma^2)) normalised; rounding
// by error diffusion from the tails inwards, the centre takes the remainder.
int prl_gauss_kernel_fixed(int n, double sigma, int* k)
{
    if (n < 1 || n > kMaxGauss || (n & 1) == 0) return PRL_E_INVALID;
    double ker[kMaxGauss];
    static const double s1[] = {1.0}, s3[] = {0.25, 0.5, 0.25}, s5[] = {0.0625, 0.25, 0.375, 0.25, 0.0625},
                        s7[] = {0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125};
    if (n <= 7 && sigma <= 0) {
        const double* t = n == 1 ? s1 : n == 3 ? s3 : n == 5 ? s5 : s7;
        for (int i = 0; i < n; ++i) ker[i] = t[i];
    } else {
        const double s = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
        const double scale2 = -0.5 / (s * s);
        double sum = 0;
        for (int i = 0; i < n; ++i) { const double x = i - (n - 1) * 0.5; ker[i] = exp(scale2 * x * x); sum += ker[i]; }
        sum = 1.0 / sum;
        for (int i = 0; i < n; ++i) ker[i] *= sum;
    }
    const int n2 = n / 2;
    double err = 0;
    long long sum = 0;
    for (int i = 0; i < n2; ++i) {
        const double adj = ker[i] * 256.0 + err;
        const long long v0 = (long long)nearbyint(adj);      // cvRound: half to even
        err = adj - (double)v0;
        k[i] = k[n - 1 - i] = (int)v0;
        sum += v0;
    }
    k[n2] = (int)(256 - 2 * sum);
    return PRL_OK;
}

int prl_k_gaussian_blur(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int ksize, double sigma,
                        uint8_t* d_dst, size_t dst_step, uint16_t* d_tmp /* rows x r16(cols) */)
{
    if (rows > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    GaussK K; K.n = ksize;
    if (prl_gauss_kernel_fixed(ksize, sigma, K.k) != PRL_OK)
        return prl_set_err(ctx, PRL_E_INVALID, "Gaussian kernel size must be odd and in [1, 63]");
    const size_t tstep = r16((size_t)cols);
    dim3 grid((cols + 255) / 256, rows);
    {
        prl_launch_scope ls(ctx, FAM_EDGES);
        gauss_rows_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, step, rows, cols, K, d_tmp, tstep);
    }
    {
        prl_launch_scope ls(ctx, FAM_EDGES);
        gauss_cols_kernel<<<grid, 256, 0, ctx->stream>>>(d_tmp, tstep, rows, cols, K, d_dst, dst_step);
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

// cv::Canny(src, edges, low, high) (aperture 3, L2gradient = false).  d_otsu != nullptr: thresholds derived on the
// device from the Otsu value there (upper = upper_coeff * otsu, lower = lower_coeff * upper).
// scratch: cls rows x r16(cols) bytes | labels rows*cols int32 | flags rows*cols bytes
int prl_k_canny(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, const int32_t* d_otsu,
                double upper_coeff, double lower_coeff, double low, double high, uint8_t* d_dst, size_t dst_step, void* scratch)
{
    if (rows > 65535 || (long long)rows * cols > 0x7fffffffLL) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "image too large");
    const size_t cstep = r16((size_t)cols);
    uint8_t* cls = (uint8_t*)scratch;
    int* L = (int*)((uint8_t*)scratch + r256(cstep * rows));
    uint8_t* flag = (uint8_t*)L + r256((size_t)rows * cols * sizeof(int));
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, (size_t)rows * cols, ctx->stream));
    dim3 grid((cols + 255) / 256, rows);
    {
        prl_launch_scope ls(ctx, FAM_EDGES);
        canny_nms_kernel<<<dim3((cols + kCT - 1) / kCT, (rows + kCH - 1) / kCH), kCT * kCH, 0, ctx->stream>>>(
            d_src, step, rows, cols, d_otsu, upper_coeff, lower_coeff, low, high, cls, cstep);
    }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_init_kernel<<<grid, 256, 0, ctx->stream>>>(cls, cstep, rows, cols, L); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_merge_kernel<<<grid, 256, 0, ctx->stream>>>(cls, cstep, rows, cols, L); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_mark_kernel<<<grid, 256, 0, ctx->stream>>>(cls, cstep, rows, cols, L, flag); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_emit_kernel<<<grid, 256, 0, ctx->stream>>>(cls, cstep, rows, cols, L, flag, d_dst, dst_step); }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

size_t prl_canny_scratch_bytes(int rows, int cols)
{
    return r256(r16((size_t)cols) * rows) + r256((size_t)rows * cols * sizeof(int)) + r256((size_t)rows * cols);
}

// Bounding rectangles of the top-level components of a 0/255 map (what cv::findContours(RETR_EXTERNAL) +
// cv::boundingRect give, binarizeLocalOtsu.cpp:104-105,150).  d_count: one int; d_xywh: cap x 4 ints.
// scratch: labels (rows*cols + 1) int32 | top rows*cols bytes | 4 x rows*cols int32 boxes
size_t prl_rects_scratch_bytes(int rows, int cols)
{
    const size_t n = (size_t)rows * cols;
    return r256((n + 1) * sizeof(int)) + r256(n) + 4 * r256(n * sizeof(int));
}

int prl_k_external_rects(prl_cuda_ctx* ctx, const uint8_t* d_edges, int rows, int cols, size_t step, int* d_count,
                         int32_t* d_xywh, int cap, void* scratch)
{
    if (rows > 65535 || (long long)rows * cols >= 0x7fffffffLL) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "image too large");
    const size_t n = (size_t)rows * cols;
    uint8_t* b = (uint8_t*)scratch;
    int* L = (int*)b; b += r256((n + 1) * sizeof(int));
    uint8_t* top = b; b += r256(n);
    int* bx0 = (int*)b; b += r256(n * sizeof(int));
    int* by0 = (int*)b; b += r256(n * sizeof(int));
    int* bx1 = (int*)b; b += r256(n * sizeof(int));
    int* by1 = (int*)b;
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(top, 0, n, ctx->stream));
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(bx0, 0x7f, 2 * r256(n * sizeof(int)), ctx->stream));      // x0, y0 = 0x7f7f7f7f
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(bx1, 0xff, 2 * r256(n * sizeof(int)), ctx->stream));      // x1, y1 = -1
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
    dim3 grid((cols + 255) / 256, rows);
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_runs_kernel<<<rows, 256, 0, ctx->stream>>>(d_edges, step, rows, cols, L); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_vmerge_kernel<<<grid, 256, 0, ctx->stream>>>(d_edges, step, rows, cols, L); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_border_kernel<<<grid, 256, 0, ctx->stream>>>(d_edges, step, rows, cols, L, top, bx0, by0, bx1, by1); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_rects_kernel<<<grid, 256, 0, ctx->stream>>>(d_edges, step, rows, cols, L, top, bx0, by0, bx1, by1, d_count, d_xywh, cap); }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
