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

// row pass: u8 -> 8.8 fixed point (exact: the coefficients sum to 256).  A CTA stages 1024 + n - 1 pixels of one row in
// shared memory (reflected at the image border); a thread produces 4 adjacent outputs, so a staged pixel is read once per
// thread instead of once per output.
constexpr int kGRowOut = 4;

__global__ void __launch_bounds__(256)
gauss_rows_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, const __grid_constant__ GaussK K, uint16_t* __restrict__ tmp,
                  size_t tstep)
{
    __shared__ uint8_t line[256 * kGRowOut + kMaxGauss + 1];
    __shared__ int kk[kMaxGauss + 1];
    const int y = blockIdx.y, xb = blockIdx.x * 256 * kGRowOut, r = K.n >> 1;
    const uint8_t* row = src + (size_t)y * step;
    for (int i = threadIdx.x; i < K.n; i += 256) kk[i] = K.k[i];
    const int span = min(256 * kGRowOut, cols - xb) + K.n - 1;
    for (int i = threadIdx.x; i < span; i += 256) line[i] = row[reflect101(xb + i - r, cols)];
    __syncthreads();
    const int x = xb + threadIdx.x * kGRowOut;
    if (x >= cols) return;
    uint32_t acc[kGRowOut] = {0, 0, 0, 0};
    const uint8_t* p = line + threadIdx.x * kGRowOut;
    for (int j = 0; j < K.n + kGRowOut - 1; ++j) {
        const uint32_t v = p[j];
#pragma unroll
        for (int o = 0; o < kGRowOut; ++o) {
            const int t = j - o;
            if (t >= 0 && t < K.n) acc[o] += v * (uint32_t)kk[t];
        }
    }
    uint16_t* out = tmp + (size_t)y * tstep + x;
#pragma unroll
    for (int o = 0; o < kGRowOut; ++o) if (x + o < cols) out[o] = (uint16_t)acc[o];
}

// column pass: 8.8 -> 16.16 -> u8.  A thread produces 8 vertically adjacent outputs of one column: n + 7 loads instead of 8 n.
constexpr int kGColOut = 8;

__global__ void __launch_bounds__(256)
gauss_cols_kernel(const uint16_t* __restrict__ tmp, size_t tstep, int rows, int cols, const __grid_constant__ GaussK K, uint8_t* __restrict__ dst,
                  size_t dstep)
{
    __shared__ int kk[kMaxGauss + 1];
    for (int i = threadIdx.x; i < K.n; i += 256) kk[i] = K.k[i];
    __syncthreads();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y0 = blockIdx.y * kGColOut;
    if (x >= cols) return;
    const int r = K.n >> 1;
    uint32_t acc[kGColOut];
#pragma unroll
    for (int o = 0; o < kGColOut; ++o) acc[o] = 0;
    const bool inner = y0 >= r && y0 + kGColOut - 1 + r < rows;
    for (int j = 0; j < K.n + kGColOut - 1; ++j) {
        const int yy = y0 + j - r;
        const uint32_t v = tmp[(size_t)(inner ? yy : reflect101(yy, rows)) * tstep + x];
#pragma unroll
        for (int o = 0; o < kGColOut; ++o) {
            const int t = j - o;
            if (t >= 0 && t < K.n) acc[o] += v * (uint32_t)kk[t];
        }
    }
#pragma unroll
    for (int o = 0; o < kGColOut; ++o)
        if (y0 + o < rows) {
            const uint32_t v = (acc[o] + 0x8000u) >> 16;
            dst[(size_t)(y0 + o) * dstep + x] = (uint8_t)min(v, 255u);
        }
}

// ---- Canny: Sobel + magnitude + non-maximum suppression -> class map (0 none, 1 weak, 2 strong) --------
constexpr int kCT = 32, kCH = 16;         // output tile

__global__ void __launch_bounds__(kCT * kCH)
canny_nms_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, const int32_t* __restrict__ otsu_thr,
                 double upper_coeff, double lower_coeff, double fixed_low, double fixed_high, uint8_t* __restrict__ cls, size_t cstep,
                 int* __restrict__ L, uint8_t* __restrict__ flag, int* __restrict__ list, int* __restrict__ count)
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
    const bool inside = gx < cols && gy < rows;
    const int m = inside ? smag[ty + 1][tx + 1] : 0;
    uint8_t out = 0;
    if (inside && m > low) {
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
    if (inside) cls[(size_t)gy * cstep + gx] = out;
    // Labels and component flags exist only for surviving pixels, and those pixels (~2 % of a page) are LISTED: the
    // hysteresis passes run one thread per listed pixel instead of one per pixel of the page (24 + 149 + 49 + 77 us per
    // A4 page before).  One atomic per warp (ballot + prefix).
    __shared__ int wcount[kCT * kCH / 32 + 1];
    const uint32_t alive = __ballot_sync(0xffffffffu, out != 0);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) wcount[wid] = __popc(alive);
    __syncthreads();
    if (threadIdx.x == 0) {                                   // one atomic per CTA
        int tot = 0;
        for (int k = 0; k < kCT * kCH / 32; ++k) { const int c = wcount[k]; wcount[k] = tot; tot += c; }
        wcount[kCT * kCH / 32] = tot ? atomicAdd(count, tot) : 0;
    }
    __syncthreads();
    if (out) {
        const int p = gy * cols + gx;
        L[p] = p; flag[p] = 0;
        list[wcount[kCT * kCH / 32] + wcount[wid] + __popc(alive & ((1u << lane) - 1u))] = p;
    }
}

// ---- hysteresis: union-find over the surviving pixels, 8-connectivity --------------------------------
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

// The hysteresis passes: one thread per LISTED (surviving) pixel, grid-stride over the list whose length lives in device memory
__global__ void __launch_bounds__(256)
ccl_merge_kernel(const uint8_t* __restrict__ cls, size_t cstep, int rows, int cols, int* __restrict__ L, const int* __restrict__ list,
                 const int* __restrict__ count)
{
    const int n = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = list[i], y = p / cols, x = p - y * cols;
        const uint8_t* row = cls + (size_t)y * cstep;
        // 8-connectivity with the fewest unions: the diagonal ones only where no 4-neighbour already bridges them
        // (NW is joined through N or W, NE through N or E, by those pixels' own unions)
        const bool w = x > 0 && row[x - 1];
        if (w) uf_union_c(L, p, p - 1);
        if (y > 0) {
            const uint8_t* up = row - cstep;
            const bool n = up[x] != 0;
            if (n) uf_union_c(L, p, p - cols);
            else {
                if (!w && x > 0 && up[x - 1]) uf_union_c(L, p, p - cols - 1);
                if (x + 1 < cols && up[x + 1] && !row[x + 1]) uf_union_c(L, p, p - cols + 1);
            }
        }
    }
}

// flag[root] = 1 for every component that holds a strong pixel
__global__ void __launch_bounds__(256)
ccl_mark_kernel(const uint8_t* __restrict__ cls, size_t cstep, int rows, int cols, int* __restrict__ L, uint8_t* __restrict__ flag,
                const int* __restrict__ list, const int* __restrict__ count)
{
    const int n = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = list[i], y = p / cols, x = p - y * cols;
        if (cls[(size_t)y * cstep + x] == 2) flag[uf_find_c(L, p)] = 1;
    }
}

// dst was cleared; the pixels of the kept components are set
__global__ void __launch_bounds__(256)
ccl_emit_kernel(int rows, int cols, int* __restrict__ L, const uint8_t* __restrict__ flag, uint8_t* __restrict__ dst, size_t dstep,
                const int* __restrict__ list, const int* __restrict__ count)
{
    const int n = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = list[i], y = p / cols, x = p - y * cols;
        if (flag[uf_find_c(L, p)]) dst[(size_t)y * dstep + x] = 255;
    }
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
// one CTA per row: L[y * cols + x] = y * cols + (start of the horizontal run of equal class that holds x).
// A component's root is its smallest label, i.e. always the start of a run: the per-root state (top-level vote, bounding
// box) is initialised here, at the starts of the 1-runs only -- no full-page memsets (17 bytes per pixel before).
__global__ void __launch_bounds__(256)
ccl_runs_kernel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, int* __restrict__ L, uint8_t* __restrict__ top,
                int* __restrict__ bx0, int* __restrict__ by0, int* __restrict__ bx1, int* __restrict__ by1)
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
        if (x == 0 || (row[x] != 0) != (row[x - 1] != 0)) {
            cur = x;
            if (row[x]) {
                const int p = y * cols + x;
                top[p] = 0; bx0[p] = 0x7f7f7f7f; by0[p] = 0x7f7f7f7f; bx1[p] = -1; by1[p] = -1;
            }
        }
        L[y * cols + x] = y * cols + cur;
    }
    if (y == 0 && tid == 0) L[rows * cols] = rows * cols;     // the frame node
}

// 16 pixels per thread: bit i of the result = pixel x0 + i is non-zero (bytes at or beyond `cols` read as 0; vec = the row
// is 16-byte aligned)
__device__ __forceinline__ uint32_t nz16(const uint8_t* __restrict__ row, int x0, int cols, bool vec)
{
    uint32_t m = 0;
    if (vec && x0 + 16 <= cols) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + x0));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t t = ((((w[k] & 0x7f7f7f7fu) + 0x7f7f7f7fu) | w[k]) & 0x80808080u) >> 7;      // 1 per non-zero byte
            t = (t | (t >> 7) | (t >> 14) | (t >> 21)) & 0xfu;
            m |= t << (4 * k);
        }
    } else {
        for (int i = 0; i < 16 && x0 + i < cols; ++i) m |= (row[x0 + i] != 0 ? 1u : 0u) << i;
    }
    return m;
}

// the unions of one pixel (what every pixel did in the one-thread-per-pixel version; now the image frame and ragged ends)
__device__ __forceinline__ void vmerge_pixel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, int* __restrict__ L, int x, int y)
{
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

// Vertical (and diagonal) unions, 16 pixels per thread.  Interior groups work on bit masks of the two rows: the positions
// that need a union are the set bits of three expressions, everything else costs two 16-byte loads.
__global__ void __launch_bounds__(256)
ccl_vmerge_kernel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, int* __restrict__ L, int vec)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16, y = blockIdx.y;
    if (x0 >= cols) return;
    if (y == 0 || y == rows - 1 || x0 == 0 || x0 + 17 > cols) {         // touches the frame, or has no right neighbour column
        for (int i = 0; i < 16 && x0 + i < cols; ++i) vmerge_pixel(E, estep, rows, cols, L, x0 + i, y);
        return;
    }
    const uint8_t* row = E + (size_t)y * estep;
    const uint8_t* up = row - estep;
    const uint32_t C = nz16(row, x0, cols, vec != 0), CN = nz16(up, x0, cols, vec != 0);
    const uint32_t cl = row[x0 - 1] != 0, ul = up[x0 - 1] != 0, ur = up[x0 + 16] != 0;
    const uint32_t CW = ((C << 1) | cl) & 0xffffu, CNW = ((CN << 1) | ul) & 0xffffu, CNE = (CN >> 1) | (ur << 15);
    uint32_t A = ~(C ^ CN) & ((CW ^ C) | (CNW ^ CN)) & 0xffffu;         // same class above, first column of the overlap
    uint32_t B = C & ~CN & CNW & ~CW & 0xffffu;                        // 1-pixel, 0 above, 1 at north-west, 0 at west
    uint32_t D = C & ~CN & CNE & 0xffffu;                              // 1-pixel, 0 above, 1 at north-east
    const int p0 = y * cols + x0;
    while (A) { const int i = __ffs(A) - 1; A &= A - 1; uf_union_c(L, p0 + i, p0 + i - cols); }
    while (B) { const int i = __ffs(B) - 1; B &= B - 1; uf_union_c(L, p0 + i, p0 + i - cols - 1); }
    while (D) { const int i = __ffs(D) - 1; D &= D - 1; uf_union_c(L, p0 + i, p0 + i - cols + 1); }
}

__device__ __forceinline__ void border_pixel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, int* __restrict__ L,
                                             uint8_t* __restrict__ top, int* __restrict__ bx0, int* __restrict__ by0, int* __restrict__ bx1,
                                             int* __restrict__ by1, int x, int y)
{
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

// Every 1-pixel with a 0 (or the frame) in its 4-neighbourhood votes and stretches its root's box.  Pass 1 (16 pixels per
// thread, bit masks of three rows) LISTS those pixels -- the candidates of an interior group are the set bits of
// C & ~(W & E & N & S) -- with one atomic per CTA; pass 2 runs one thread per listed pixel.
__global__ void __launch_bounds__(256)
ccl_border_list_kernel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, int vec, int* __restrict__ list, int* __restrict__ count)
{
    __shared__ int wcount[9];
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16, y = blockIdx.y;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t cand = 0;
    if (x0 < cols) {
        const uint8_t* row = E + (size_t)y * estep;
        const uint32_t C = nz16(row, x0, cols, vec != 0);
        cand = C;
        if (C && y > 0 && y < rows - 1 && x0 > 0 && x0 + 17 <= cols) {
            const uint32_t N = nz16(row - estep, x0, cols, vec != 0), S = nz16(row + estep, x0, cols, vec != 0);
            const uint32_t cl = row[x0 - 1] != 0, cr = row[x0 + 16] != 0;
            const uint32_t W = ((C << 1) | cl) & 0xffffu, Eb = (C >> 1) | (cr << 15);
            cand = C & ~(W & Eb & N & S);
        }
    }
    // exclusive prefix of the candidate counts over the CTA
    const int mine = __popc(cand);
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) wcount[wid] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int k = 0; k < 8; ++k) { const int c = wcount[k]; wcount[k] = tot; tot += c; }
        wcount[8] = tot ? atomicAdd(count, tot) : 0;
    }
    __syncthreads();
    int o = wcount[8] + wcount[wid] + incl - mine;
    const int p0 = y * cols + x0;
    while (cand) { const int i = __ffs(cand) - 1; cand &= cand - 1; list[o++] = p0 + i; }
}

__global__ void __launch_bounds__(256)
ccl_border_kernel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, int* __restrict__ L, uint8_t* __restrict__ top,
                  int* __restrict__ bx0, int* __restrict__ by0, int* __restrict__ bx1, int* __restrict__ by1, const int* __restrict__ list,
                  const int* __restrict__ count)
{
    const int n = *count;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int p = list[i], y = p / cols;
        border_pixel(E, estep, rows, cols, L, top, bx0, by0, bx1, by1, p - y * cols, y);
    }
}

// roots are starts of 1-runs: only those positions are looked at
__global__ void __launch_bounds__(256)
ccl_rects_kernel(const uint8_t* __restrict__ E, size_t estep, int rows, int cols, const int* __restrict__ L, const uint8_t* __restrict__ top,
                 const int* __restrict__ bx0, const int* __restrict__ by0, const int* __restrict__ bx1, const int* __restrict__ by1,
                 int* __restrict__ count, int32_t* __restrict__ xywh, int cap, int vec)
{
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 16, y = blockIdx.y;
    if (x0 >= cols) return;
    const uint8_t* row = E + (size_t)y * estep;
    const uint32_t C = nz16(row, x0, cols, vec != 0);
    if (!C) return;
    const uint32_t cl = x0 > 0 ? (row[x0 - 1] != 0) : 0u;
    uint32_t starts = C & ~(((C << 1) | cl) & 0xffffu);
    while (starts) {
        const int i = __ffs(starts) - 1; starts &= starts - 1;
        const int p = y * cols + x0 + i;
        if (L[p] != p || !top[p]) continue;
        const int slot = atomicAdd(count, 1);
        if (slot < cap) {
            xywh[4 * slot] = bx0[p]; xywh[4 * slot + 1] = by0[p];
            xywh[4 * slot + 2] = bx1[p] - bx0[p] + 1; xywh[4 * slot + 3] = by1[p] - by0[p] + 1;
        }
    }
}

// ---- CLAHE + equalizeHist: EnhanceLocalContrastByCLAHE (imageLibCommon.cpp:326-346), the optional first step of
// prl::binarizeLocalOtsu (binarizeLocalOtsu.cpp:79-82).  cv::createCLAHE() defaults: 8 x 8 tiles; the image is extended
// to a multiple of the tile grid with BORDER_REFLECT_101 on the bottom / right; per tile: histogram, clip at
// max(int(clipLimit * area / 256), 1) with OpenCV's redistribution (equal share + one more in every (256 / residual)-th
// bin), LUT = round(cumsum * (255.f / area)); per pixel: bilinear blend of the four neighbouring tiles' LUTs in FP32
// with OpenCV's operation order (no FMA).  equalizeHist: LUT = round(cumsum over bins above the first non-empty one
// * (255.f / (total - hist[first]))).
constexpr int kClaheTiles = 8;

__global__ void __launch_bounds__(256)
clahe_lut_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, int tw, int th, int clip, float lut_scale,
                 uint8_t* __restrict__ luts)
{
    __shared__ uint32_t hist[256];
    __shared__ uint32_t part[8];
    const int tx = blockIdx.x % kClaheTiles, ty = blockIdx.x / kClaheTiles, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < tw * th; i += 256) {
        const int y = ty * th + i / tw, x = tx * tw + i % tw;
        atomicAdd(&hist[src[(size_t)reflect101(y, rows) * step + reflect101(x, cols)]], 1u);
    }
    __syncthreads();
    uint32_t h = hist[tid];
    if (clip > 0) {
        uint32_t ex = h > (uint32_t)clip ? h - clip : 0u;
        h = min(h, (uint32_t)clip);
        ex = __reduce_add_sync(0xffffffffu, ex);
        if (lane == 0) part[wid] = ex;
        __syncthreads();
        uint32_t clipped = 0;
        for (int k = 0; k < 8; ++k) clipped += part[k];
        const uint32_t batch = clipped / 256u, residual = clipped - batch * 256u;
        h += batch;
        if (residual != 0) {
            const uint32_t st = max(256u / residual, 1u);
            if (tid % st == 0 && tid / st < residual) h += 1;
        }
        __syncthreads();
    }
    // inclusive prefix sum over the 256 bins
    uint32_t v = h;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
    if (lane == 31) part[wid] = v;
    __syncthreads();
    for (int k = 0; k < wid; ++k) v += part[k];
    const int r = __float2int_rn(__fmul_rn((float)v, lut_scale));
    luts[(size_t)blockIdx.x * 256 + tid] = (uint8_t)min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256)
clahe_interp_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, float inv_tw, float inv_th,
                    const uint8_t* __restrict__ luts, uint8_t* __restrict__ dst, size_t dstep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= cols) return;
    const float txf = __fadd_rn(__fmul_rn((float)x, inv_tw), -0.5f), tyf = __fadd_rn(__fmul_rn((float)y, inv_th), -0.5f);
    int tx1 = (int)floorf(txf), ty1 = (int)floorf(tyf);
    int tx2 = tx1 + 1, ty2 = ty1 + 1;
    const float xa = __fadd_rn(txf, -(float)tx1), xa1 = __fadd_rn(1.0f, -xa);
    const float ya = __fadd_rn(tyf, -(float)ty1), ya1 = __fadd_rn(1.0f, -ya);
    tx1 = max(tx1, 0); tx2 = min(tx2, kClaheTiles - 1); ty1 = max(ty1, 0); ty2 = min(ty2, kClaheTiles - 1);
    const int v = src[(size_t)y * step + x];
    const float a = (float)luts[(ty1 * kClaheTiles + tx1) * 256 + v], b = (float)luts[(ty1 * kClaheTiles + tx2) * 256 + v];
    const float c = (float)luts[(ty2 * kClaheTiles + tx1) * 256 + v], d = (float)luts[(ty2 * kClaheTiles + tx2) * 256 + v];
    const float top = __fadd_rn(__fmul_rn(a, xa1), __fmul_rn(b, xa)), bot = __fadd_rn(__fmul_rn(c, xa1), __fmul_rn(d, xa));
    const int r = __float2int_rn(__fadd_rn(__fmul_rn(top, ya1), __fmul_rn(bot, ya)));
    dst[(size_t)y * dstep + x] = (uint8_t)min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256)
hist256_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, uint32_t* __restrict__ hist)
{
    __shared__ uint32_t sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    for (int y = blockIdx.x; y < rows; y += gridDim.x)
        for (int x = threadIdx.x; x < cols; x += 256) atomicAdd(&sh[src[(size_t)y * step + x]], 1u);
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// cv::equalizeHist's LUT from the page histogram (one CTA of 256 threads)
__global__ void __launch_bounds__(256)
equalize_lut_kernel(const uint32_t* __restrict__ hist, uint32_t total, uint8_t* __restrict__ lut)
{
    __shared__ uint32_t part[8];
    __shared__ int first_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const uint32_t h = hist[tid];
    if (tid == 0) first_s = 256;
    __syncthreads();
    if (h) atomicMin(&first_s, tid);
    __syncthreads();
    const int first = first_s;
    const uint32_t hf = hist[first];
    if (hf == total) { lut[tid] = (uint8_t)first; return; }            // one grey level: dst.setTo(first)
    const float scale = __fdiv_rn(255.0f, (float)(total - hf));
    uint32_t v = tid > first ? h : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += t; }
    if (lane == 31) part[wid] = v;
    __syncthreads();
    for (int k = 0; k < wid; ++k) v += part[k];
    const int r = tid > first ? __float2int_rn(__fmul_rn((float)v, scale)) : 0;
    lut[tid] = (uint8_t)min(max(r, 0), 255);
}

__global__ void __launch_bounds__(256)
apply_lut_kernel(const uint8_t* __restrict__ src, size_t step, int rows, int cols, const uint8_t* __restrict__ lut, uint8_t* __restrict__ dst,
                 size_t dstep)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x < cols) dst[(size_t)y * dstep + x] = lut[src[(size_t)y * step + x]];
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
    {
        prl_launch_scope ls(ctx, FAM_EDGES);
        gauss_rows_kernel<<<dim3((cols + 256 * kGRowOut - 1) / (256 * kGRowOut), rows), 256, 0, ctx->stream>>>(d_src, step, rows, cols, K, d_tmp, tstep);
    }
    {
        prl_launch_scope ls(ctx, FAM_EDGES);
        gauss_cols_kernel<<<dim3((cols + 255) / 256, (rows + kGColOut - 1) / kGColOut), 256, 0, ctx->stream>>>(d_tmp, tstep, rows, cols, K, d_dst, dst_step);
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

// cv::Canny(src, edges, low, high) (aperture 3, L2gradient = false).  d_otsu != nullptr: thresholds derived on the
// device from the Otsu value there (upper = upper_coeff * otsu, lower = lower_coeff * upper).
// scratch: cls rows x r16(cols) bytes | labels rows*cols int32 | flags rows*cols bytes | list rows*cols + 1 int32
int prl_k_canny(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, const int32_t* d_otsu,
                double upper_coeff, double lower_coeff, double low, double high, uint8_t* d_dst, size_t dst_step, void* scratch)
{
    if (rows > 65535 || (long long)rows * cols > 0x7fffffffLL) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "image too large");
    const size_t cstep = r16((size_t)cols);
    uint8_t* cls = (uint8_t*)scratch;
    int* L = (int*)((uint8_t*)scratch + r256(cstep * rows));
    uint8_t* flag = (uint8_t*)L + r256((size_t)rows * cols * sizeof(int));
    int* list = (int*)(flag + r256((size_t)rows * cols));
    int* count = list + (size_t)rows * cols;
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(count, 0, sizeof(int), ctx->stream));
    PRL_CUDA_TRY(ctx, cudaMemset2DAsync(d_dst, dst_step, 0, cols, rows, ctx->stream));
    const int lgrid = ctx->num_sms * 8;
    {
        prl_launch_scope ls(ctx, FAM_EDGES);
        canny_nms_kernel<<<dim3((cols + kCT - 1) / kCT, (rows + kCH - 1) / kCH), kCT * kCH, 0, ctx->stream>>>(
            d_src, step, rows, cols, d_otsu, upper_coeff, lower_coeff, low, high, cls, cstep, L, flag, list, count);
    }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_merge_kernel<<<lgrid, 256, 0, ctx->stream>>>(cls, cstep, rows, cols, L, list, count); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_mark_kernel<<<lgrid, 256, 0, ctx->stream>>>(cls, cstep, rows, cols, L, flag, list, count); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_emit_kernel<<<lgrid, 256, 0, ctx->stream>>>(rows, cols, L, flag, d_dst, dst_step, list, count); }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

size_t prl_canny_scratch_bytes(int rows, int cols)
{
    // class map | labels | flags | list of surviving pixels + its length
    return r256(r16((size_t)cols) * rows) + r256((size_t)rows * cols * sizeof(int)) + r256((size_t)rows * cols) +
           r256(((size_t)rows * cols + 1) * sizeof(int));
}

// Bounding rectangles of the top-level components of a 0/255 map (what cv::findContours(RETR_EXTERNAL) +
// cv::boundingRect give, binarizeLocalOtsu.cpp:104-105,150).  d_count: one int; d_xywh: cap x 4 ints.
// scratch: labels (rows*cols + 1) int32 | top rows*cols bytes | 4 x rows*cols int32 boxes | border-pixel list (rows*cols + 1) int32
size_t prl_rects_scratch_bytes(int rows, int cols)
{
    const size_t n = (size_t)rows * cols;
    return r256((n + 1) * sizeof(int)) + r256(n) + 4 * r256(n * sizeof(int)) + r256((n + 1) * sizeof(int));     // + list of border pixels
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
    int* blist = (int*)((uint8_t*)by1 + r256(n * sizeof(int)));
    int* bcount = blist + n;
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(d_count, 0, sizeof(int), ctx->stream));
    PRL_CUDA_TRY(ctx, cudaMemsetAsync(bcount, 0, sizeof(int), ctx->stream));
    const int vec = ((((uintptr_t)d_edges) | step) & 15) == 0;
    dim3 grid16((cols + 16 * 256 - 1) / (16 * 256), rows);            // 16 pixels per thread
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_runs_kernel<<<rows, 256, 0, ctx->stream>>>(d_edges, step, rows, cols, L, top, bx0, by0, bx1, by1); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_vmerge_kernel<<<grid16, 256, 0, ctx->stream>>>(d_edges, step, rows, cols, L, vec); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_border_list_kernel<<<grid16, 256, 0, ctx->stream>>>(d_edges, step, rows, cols, vec, blist, bcount); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_border_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(d_edges, step, rows, cols, L, top, bx0, by0, bx1, by1, blist, bcount); }
    { prl_launch_scope ls(ctx, FAM_EDGES); ccl_rects_kernel<<<grid16, 256, 0, ctx->stream>>>(d_edges, step, rows, cols, L, top, bx0, by0, bx1, by1, d_count, d_xywh, cap, vec); }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

// EnhanceLocalContrastByCLAHE(src, dst, clip_limit, equalize) for one channel (imageLibCommon.cpp:326-346).
// d_tmp: rows x dst_step bytes (the CLAHE output when equalize is set); scratch: 64 x 256 LUT bytes + 256 hist words + 256 LUT bytes
int prl_k_clahe(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, double clip_limit, bool equalize,
                uint8_t* d_dst, size_t dst_step, uint8_t* d_tmp, void* scratch)
{
    if (rows > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    uint8_t* luts = (uint8_t*)scratch;
    uint32_t* hist = (uint32_t*)(luts + kClaheTiles * kClaheTiles * 256);
    uint8_t* eq_lut = (uint8_t*)(hist + 256);
    // tile grid over the image extended to a multiple of 8 (copyMakeBorder bottom / right, BORDER_REFLECT_101).  As
    // OpenCV writes it: when EITHER dimension is not a multiple, BOTH are extended by 8 - (size % 8), i.e. a dimension
    // that already is a multiple grows by a whole 8.
    const bool exact = cols % kClaheTiles == 0 && rows % kClaheTiles == 0;
    const int ext_w = exact ? cols : cols + kClaheTiles - cols % kClaheTiles;
    const int ext_h = exact ? rows : rows + kClaheTiles - rows % kClaheTiles;
    const int tw = ext_w / kClaheTiles, th = ext_h / kClaheTiles, area = tw * th;
    int clip = 0;
    if (clip_limit > 0.0) { clip = (int)(clip_limit * area / 256); clip = std::max(clip, 1); }
    const float lut_scale = 255.0f / (float)area;
    const float inv_tw = 1.0f / (float)tw, inv_th = 1.0f / (float)th;
    dim3 grid((cols + 255) / 256, rows);
    uint8_t* clahe_out = equalize ? d_tmp : d_dst;
    { prl_launch_scope ls(ctx, FAM_EDGES); clahe_lut_kernel<<<kClaheTiles * kClaheTiles, 256, 0, ctx->stream>>>(d_src, step, rows, cols, tw, th, clip, lut_scale, luts); }
    { prl_launch_scope ls(ctx, FAM_EDGES); clahe_interp_kernel<<<grid, 256, 0, ctx->stream>>>(d_src, step, rows, cols, inv_tw, inv_th, luts, clahe_out, dst_step); }
    if (equalize) {
        PRL_CUDA_TRY(ctx, cudaMemsetAsync(hist, 0, 256 * sizeof(uint32_t), ctx->stream));
        { prl_launch_scope ls(ctx, FAM_EDGES); hist256_kernel<<<std::min(rows, 8 * ctx->num_sms), 256, 0, ctx->stream>>>(d_tmp, dst_step, rows, cols, hist); }
        { prl_launch_scope ls(ctx, FAM_EDGES); equalize_lut_kernel<<<1, 256, 0, ctx->stream>>>(hist, (uint32_t)((size_t)rows * cols), eq_lut); }
        { prl_launch_scope ls(ctx, FAM_EDGES); apply_lut_kernel<<<grid, 256, 0, ctx->stream>>>(d_tmp, dst_step, rows, cols, eq_lut, d_dst, dst_step); }
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
