// common.cuh -- context, error plumbing and launch instrumentation shared by libprlib_cuda.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include <map>
#include "../../include/prlib_cuda.h"

#define PRL_NUM_SMS_FALLBACK 148

// Layout of the integral planes in HBM: rows of `pitch` int64 elements, pitch a multiple of 16
// elements (128 B) so every row starts on a cache-line boundary and 16-byte vector accesses at
// even columns are aligned.
static inline size_t prl_plane_pitch(int padded_cols) { return ((size_t)padded_cols + 15) & ~(size_t)15; }

// The integral planes of a chunk of pages, in one of two layouts:
//   compact == 0  S, Q: int64 (what cv::integral(CV_64F) holds, exactly) -- the exported bit-exactness hook
//                 (prl_cuda_integral_u8*), the literal-FP64 kernels, and any buffer the fast path cannot take;
//   compact == 1  ONE plane of uint2 {S mod 2^32, Q mod 2^32} per padded pixel (S points to it, Q is unused): window sums
//                 are < 2^32, so differences of the low words ARE the exact window sums.  The HIGH words are kept only at
//                 anchor positions (Y % A == 0, X % 4 == 0), A = 1 << ashift, as uint2 {S >> 32, Q >> 32} in AS.  A full
//                 int64 tap -- needed only by the ~1e-3 of the pixels that evaluate the reference's FP64 formula -- is
//                     full(Y, X) = (hi(Ya, Xa) << 32 | lo(Ya, Xa)) + u32(lo(Y, X) - lo(Ya, Xa)),   Ya = Y & ~(A-1), Xa = X & ~3,
//                 exact because an integral grows by less than 2^32 between an anchor and the (A-1) rows / 3 columns it
//                 serves (prl_anchor_shift).  8.5 bytes per padded pixel on both sides of the kernel-1 / kernel-2 hand-over
//                 instead of 16, and 32-byte vectors of 4 pixels for stores, loads and TMA boxes.
struct prl_planes {
    int compact = 0;
    void* S = nullptr;          // int64_t* (compact: uint2*, low words of S and Q interleaved)
    void* Q = nullptr;          // int64_t* (compact: unused)
    size_t pitch = 0;           // elements per row (rows are 128-byte multiples)
    size_t page_stride = 0;     // elements per page
    void* AS = nullptr;         // compact: uint2 high words at the anchors, ceil(Hp / A) rows of a_pitch = pitch / 4 elements
    int ashift = 0;
    size_t a_pitch = 0;
    size_t a_page_stride = 0;   // elements per page in AS
};
// largest A = 1 << shift <= 8 with ((A-1) Wp + 3 Hp) * 255^2 < 2^32; -1: no anchor spacing works (Hp >= 22016): keep int64 planes
static inline int prl_anchor_shift(int padded_rows, int padded_cols)
{
    for (int sh = 3; sh >= 0; --sh)
        if (((double)((1 << sh) - 1) * padded_cols + 3.0 * padded_rows) * 65025.0 < 4294967296.0) return sh;
    return -1;
}

struct prl_timing_rec { cudaEvent_t a, b; int family; };

struct prl_cuda_ctx {
    int device = 0;
    int num_sms = PRL_NUM_SMS_FALLBACK;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;      // the stream every launch goes to (own or borrowed)
    std::string err;

    // scratch: integral planes S and Q for the in-flight chunk of pages
    int64_t* planes = nullptr;  size_t planes_bytes = 0;
    size_t workspace_limit = (size_t)48 << 30;
    // band carries (latency mode), per-page scalars {imin (u32), smax (i64 bit pattern)}
    void* carry = nullptr;      size_t carry_bytes = 0;
    void* colsum = nullptr;     size_t colsum_bytes = 0;
    void* scalars = nullptr;    size_t scalars_bytes = 0;
    void* sched = nullptr;      size_t sched_bytes = 0;    // work counters of the persistent kernels
    void* fused_ws = nullptr;   size_t fused_ws_bytes = 0; // fused path: row sums per strip, fixup counters and lists
    unsigned long long* d_redo_total = nullptr;            // pages the fused path handed back so far (device counter)
    int fused_page_cap = 128;                              // undecided pixels per page the fused path finishes itself
    bool fused_no_tier2 = false;                           // validation: fused path without its FP64 estimate tier
    // device staging for host-pointer entry points
    uint8_t* d_in = nullptr;    size_t d_in_bytes = 0;
    uint8_t* d_out = nullptr;   size_t d_out_bytes = 0;
    uint8_t* d_tmp = nullptr;   size_t d_tmp_bytes = 0;
    uint8_t* d_bgr = nullptr;   size_t d_bgr_bytes = 0;    // 3/4-channel staging of the cvtColor front step
    void* d_misc = nullptr;     size_t d_misc_bytes = 0;   // histograms, thresholds, rect lists
    void* clahe_ws = nullptr;   size_t clahe_ws_bytes = 0; // CLAHE: enhanced image, intermediate, LUTs
    void* rects_ws = nullptr;   size_t rects_ws_bytes = 0; // contour rectangles: count, list, thresholds, labels, boxes
    uint8_t* d_res = nullptr;   size_t d_res_bytes = 0;    // adaptive family, single-image call: the result image
    void* adaptive_ws = nullptr; size_t adaptive_ws_bytes = 0; // adaptive family: pixel counter, float32 rows of the Gaussian mean
    void* edges_ws = nullptr;   size_t edges_ws_bytes = 0; // edge front-end: blurred image, 8.8 rows, class map, labels, flags
    // pinned host staging

    bool force_exact = false;   // validation: kernel 2 runs the literal FP64 path for every pixel
    bool gauss_legacy = false;   // adaptiveThreshold GAUSSIAN_C: the two-kernel form (float32 plane in HBM) for every block size (A/B and tests)
    bool median_legacy = false;  // medianBlur 3 / 5: the radix-select kernel instead of the selection network (A/B and tests)
    bool dbg_skip_exact = false; // DIAGNOSTIC ONLY: kernel 2 (TMA) leaves undecided pixels black; timing experiments, never a result
    int k1_bands = 0;           // kernel 1: row bands per page (0 = automatic)
    int thr_stages = 2;         // kernel 2 (TMA): ring stages per CTA (2: three CTAs per SM -- measured 7 % faster; 3: two CTAs per SM)
    bool thr_no_tma = false;    // validation: kernel 2 on compact planes runs the register-staged streaming kernel instead of the TMA ring
    bool thr_legacy = false;    // validation: kernel 2's mask path runs the round-1 two-tier kernel instead of the streaming kernel
    bool no_compact = false;    // validation: the two-kernel path keeps full int64 planes (16 B per padded pixel) instead of the compact layout
    bool no_tma = false;        // validation: kernel 1 uses the generic (non-TMA) kernel
    int thr_rows = 0;           // kernel 2: output rows per CTA (0 = automatic: 4, or 8 when the tap distance exceeds 64)
    int tiles_legacy = 0;       // validation: tile Otsu runs 1 = the round-1 warp-batched kernel, 2 = the lane-per-tile kernel with register-staged
                                // global loads, 3 = the ring-fed lane-per-tile kernel for every tile width up to 64 (0: the ring-fed kernel for
                                // 64-wide tiles at least 48 high, kernel 2 for the other aligned shapes up to 128 wide, kernel 1 for the rest)
    int tile_prefetch = 0;      // tile Otsu: how many tiles ahead a warp pulls into L2 (0 = off)
    bool morph_bytes = false;   // validation: the morphology tail runs the byte kernels even on binary masks
    bool use_fused = true;      // windows <= 31 take the fused small-window strip kernel (integral planes never reach HBM);
                                // set_option("enable_fused", 0) forces kernel 1 + kernel 2

    // page lanes of the batched F3 / F4 entry points: sub-contexts (own stream, own scratch) that run the single-image
    // kernel sequences of several pages side by side
    std::vector<prl_cuda_ctx*> lanes;
    int* h_lane_counts = nullptr;                          // pinned: contour counts of the pages in flight
    cudaEvent_t lane_ev[3] = {nullptr, nullptr, nullptr};  // lane 0 / lane 1 done, context stream reached the call
    int otsu_group = -1;                                   // Global Otsu batches: pages per group of the two-lane overlap (-1 = n/4 in [16, 256], 0 = off)

    // instrumentation
    bool timing = false;
    long long launches = 0;
    std::vector<prl_timing_rec> recs;
    std::vector<cudaEvent_t> event_pool;
    std::map<int, std::pair<double, long long>> totals;
};

enum prl_family {
    FAM_INTEGRAL = 0, FAM_THRESHOLD, FAM_SMAX, FAM_MORPH, FAM_OTSU_HIST, FAM_OTSU_SEARCH,
    FAM_OTSU_APPLY, FAM_OTSU_TILES, FAM_SYNTH, FAM_BGR2GRAY, FAM_BAND_CARRY, FAM_FUSED, FAM_FUSED_PRE, FAM_FUSED_FIX, FAM_PACK, FAM_EDGES, FAM_LINES, FAM_ADAPTIVE, FAM_COUNT
};

int  prl_set_err(prl_cuda_ctx* ctx, int code, const char* what, cudaError_t ce = cudaSuccess);
int  prl_ensure(prl_cuda_ctx* ctx, void** ptr, size_t* have, size_t need);          // device scratch
void prl_launch_begin(prl_cuda_ctx* ctx, int family);
void prl_launch_end(prl_cuda_ctx* ctx);

inline size_t round16(size_t v) { return (v + 15) & ~(size_t)15; }

// 2-D copy that degenerates to ONE linear DMA when both pitches equal the row width: the copy
// engines move 2.4 KB rows at ~15 GB/s but a linear range at ~55 GB/s (measured, PCIe Gen5 x16).
inline cudaError_t copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                          cudaMemcpyKind kind, cudaStream_t s)
{
    if (dpitch == width && spitch == width) return cudaMemcpyAsync(dst, src, width * height, kind, s);
    return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, s);
}

#define PRL_CUDA_TRY(ctx, call)                                                        \
    do { cudaError_t _e = (call);                                                      \
         if (_e != cudaSuccess) return prl_set_err((ctx), PRL_E_CUDA, #call, _e); } while (0)

// RAII bracket: counts the launch and, when timing is on, records events around it.
struct prl_launch_scope {
    prl_cuda_ctx* c;
    prl_launch_scope(prl_cuda_ctx* ctx, int family) : c(ctx) { prl_launch_begin(c, family); }
    ~prl_launch_scope() { prl_launch_end(c); }
};

// ---- kernel launchers (defined in the .cu files; all asynchronous on ctx->stream) --------
struct prl_geom {
    int rows, cols;        // source page
    int w, h, d;           // clamped window, w/2, w-1
    int Hp, Wp;            // padded size
    int out_rows, out_cols;
    size_t pitch;          // plane pitch in int64 elements
};
int prl_make_geom(int method, int rows, int cols, int window, prl_geom* g);

int prl_k_integral(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                   size_t src_page_stride, int pad, int64_t* d_S, int64_t* d_Q, size_t pitch,
                   size_t plane_page_stride, uint32_t* d_imin /*per page or null*/);
// same, into either plane layout; the compact layout needs the TMA kernel (prl_integral_compact_ok)
int prl_k_integral_planes(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                          size_t src_page_stride, int pad, const prl_planes& P, uint32_t* d_imin);
bool prl_integral_compact_ok(const prl_cuda_ctx* ctx, const uint8_t* d_src, size_t src_step, size_t src_page_stride, int rows, int cols, int pad);
// the compact-layout kernel (integral_sq.cu) and the band pre-pass it shares with the int64 kernel (integral.cu)
int prl_k_integral_sq(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                      size_t src_page_stride, int pad, const prl_planes& P, uint32_t* d_imin);
int prl_choose_bands(const prl_cuda_ctx* ctx, int n_pages, int rows, int ctas_per_sm);
int prl_band_carries(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                     size_t src_page_stride, int pad, int bands, int rpb, size_t pitch, const int64_t** d_carry);
bool prl_threshold_fast_ok(const prl_cuda_ctx* ctx, int method, const double* params, const prl_geom& g,
                           const uint8_t* d_src, size_t src_step, size_t src_page_stride);
struct prl_thr_params { double kw, nkw, p0, p1, p2; };
int prl_k_threshold(prl_cuda_ctx* ctx, int method, int mode /*0 mask, 1 T8*/, const uint8_t* d_src, int n_pages,
                    const prl_geom& g, size_t src_step, size_t src_page_stride, const prl_planes& P, const double* params,
                    const uint32_t* d_imin, long long* d_smax, uint8_t* d_dst, size_t dst_step,
                    size_t dst_page_stride);
bool prl_fused_eligible(const prl_cuda_ctx* ctx, int method, int n_pages, const prl_geom& g, const double* params);
int prl_k_fused(prl_cuda_ctx* ctx, int method, const uint8_t* d_src, int n_pages, const prl_geom& g, size_t src_step,
                size_t src_page_stride, const double* params, uint32_t* d_imin, uint8_t* d_dst, size_t dst_step,
                size_t dst_page_stride, const int** d_redo_map, const int** d_redo_count);
// indirect launches over a device-resident page list (hand-back of the fused path)
int prl_k_integral_indirect(prl_cuda_ctx* ctx, const uint8_t* d_src, int slots, int rows, int cols, size_t src_step,
                            size_t src_page_stride, int pad, int64_t* d_S, int64_t* d_Q, size_t pitch, size_t plane_page_stride,
                            const int* d_map, const int* d_count, int slot_base);
int prl_k_threshold_exact_indirect(prl_cuda_ctx* ctx, int method, const uint8_t* d_src, int slots, const prl_geom& g, size_t src_step,
                                   size_t src_page_stride, const prl_planes& P, const double* params, const uint32_t* d_imin,
                                   uint8_t* d_dst, size_t dst_step, size_t dst_page_stride, const int* d_map, const int* d_count,
                                   int slot_base);
int prl_k_morph(prl_cuda_ctx* ctx, uint8_t* d_in, uint8_t* d_out, int n_pages, int rows, int cols, size_t in_step,
                size_t in_page_stride, size_t out_step, size_t out_page_stride, int iters, bool binary);
int prl_k_morph_single(prl_cuda_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, int n_pages, int rows, int cols, size_t in_step,
                       size_t in_page_stride, size_t out_step, size_t out_page_stride, int n, bool dilate);
int prl_gauss_kernel_fixed(int n, double sigma, int* k);
// the adaptive-mean family (adaptive.cu)
int prl_gauss_kernel_float(int n, float* k);
size_t prl_adaptive_scratch_bytes(int rows, int cols);
int prl_k_median_blur(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int channels, int ksize,
                      uint8_t* d_dst, size_t dst_step);
size_t prl_bilateral_scratch_bytes(int d, double sigma_space);
int prl_k_bilateral(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int d, double sigma_color, double sigma_space,
                    uint8_t* d_dst, size_t dst_step, void* scratch);
int prl_k_adaptive_threshold(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, double maxval, int method, int type,
                             int block_size, double delta, uint8_t* d_dst, size_t dst_step, void* scratch, bool invert_if_dark);
int prl_k_gaussian_blur(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int ksize, double sigma,
                        uint8_t* d_dst, size_t dst_step, uint16_t* d_tmp);
int prl_k_canny(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, const int32_t* d_otsu,
                double upper_coeff, double lower_coeff, double low, double high, uint8_t* d_dst, size_t dst_step, void* scratch);
int prl_k_clahe(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, double clip_limit, bool equalize,
                uint8_t* d_dst, size_t dst_step, uint8_t* d_tmp, void* scratch);
size_t prl_canny_scratch_bytes(int rows, int cols);
size_t prl_rects_scratch_bytes(int rows, int cols);
size_t prl_lines_scratch_bytes(int rows, int cols);
int prl_k_remove_lines(prl_cuda_ctx* ctx, const uint8_t* d_gray, int rows, int cols, size_t step, uint8_t* d_dst, size_t dst_step,
                       void* scratch);
int prl_k_external_rects(prl_cuda_ctx* ctx, const uint8_t* d_edges, int rows, int cols, size_t step, int* d_count,
                         int32_t* d_xywh, int cap, void* scratch);
int prl_k_not_binary(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int* d_flag);
int prl_k_pack_mask(prl_cuda_ctx* ctx, const uint8_t* d_mask, int n_pages, int rows, int cols, size_t step,
                    size_t page_stride, uint32_t* d_bits);
int prl_k_bgr2gray(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int channels,
                   uint8_t* d_dst, size_t dst_step, bool rgb = false);
int prl_k_synth(prl_cuda_ctx* ctx, uint8_t* d_dst, int n_pages, int rows, int cols, size_t step,
                size_t page_stride, uint32_t seed, uint32_t first_page);
int prl_k_otsu_global(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                      size_t src_page_stride, double maxval, uint8_t* d_dst, size_t dst_step,
                      size_t dst_page_stride, int32_t* d_thr /*n_pages*/, bool apply);
int prl_k_otsu_rects(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t src_step,
                     const int32_t* d_xywh, int n_rects, double maxval, uint8_t* d_dst, size_t dst_step,
                     int32_t* d_thr);
int prl_k_otsu_tiles(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                     size_t src_page_stride, int tile_w, int tile_h, double maxval, uint8_t* d_dst,
                     size_t dst_step, size_t dst_page_stride);
