/*
 * prlib_cuda.h -- C-ABI of libprlib_cuda, the B200 (sm_100a) implementation of PRLib's
 * local-statistics + Otsu binarization path.
 *
 * The reference has no FFI layer: its "operator API" for this path is six C++ free functions
 * (namespace prl, cv::Mat in / cv::Mat out).  Each entry point below names the reference
 * interface (file:line under the PRLib tree) whose body it replaces.  The cv::Mat-facing shim
 * that keeps the reference signatures lives in prlib_b200/shim/ and calls only this header.
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types; every function returns int:
 *     PRL_OK (0) or a negative PRL_E_* code; nothing throws, nothing calls back.
 *   - `step` / `*_step` are row pitches in BYTES (cv::Mat::step).  Host-pointer entry points
 *     accept pageable or pinned memory; `*_dev` entry points take device pointers that are
 *     already resident in HBM and run asynchronously on the context's stream.
 *   - there is NO CPU fallback: without a usable CUDA device every call fails with PRL_E_CUDA.
 *   - a context owns one device, one stream (or borrows one), scratch planes and staging
 *     buffers; a context must not be used from two host threads at once (use one per thread).
 */
#ifndef PRLIB_CUDA_H
#define PRLIB_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRL_OK              0
#define PRL_E_INVALID      (-1)  /* empty image, NULL pointer, bad window (reference: std::invalid_argument) */
#define PRL_E_EMPTY_ROI    (-2)  /* min(rows, cols) <= window for WJ/NICK/Feng (reference: cv::Exception) */
#define PRL_E_CUDA         (-3)  /* CUDA runtime / launch failure, or no device */
#define PRL_E_NOMEM        (-4)  /* device or pinned-host allocation failed */
#define PRL_E_UNSUPPORTED  (-5)  /* shape outside what the kernels handle (e.g. padded width > 65536) */

/* local-statistics methods */
enum {
    PRL_SAUVOLA    = 0,  /* prl::binarizeSauvola     binarizeSauvola.h:43-47,    binarizeSauvola.cpp:32-135    */
    PRL_NIBLACK    = 1,  /* prl::binarizeNiblack     binarizeNiblack.h:43-47,    binarizeNiblack.cpp:32-127    */
    PRL_WOLFJOLION = 2,  /* prl::binarizeWolfJolion  binarizeWolfJolion.h:43-47, binarizeWolfJolion.cpp:33-148 */
    PRL_NICK       = 3,  /* prl::binarizeNICK        binarizeNICK.h:43-47,       binarizeNICK.cpp:33-144       */
    PRL_FENG       = 4   /* prl::binarizeFeng        binarizeFeng.h:46-53,       binarizeFeng.cpp:31-164       */
};

typedef struct prl_cuda_ctx prl_cuda_ctx;

/* ---- context ------------------------------------------------------------------------- */
int         prl_cuda_device_count(void);
int         prl_cuda_create(int device, prl_cuda_ctx** out);
void        prl_cuda_destroy(prl_cuda_ctx* ctx);
const char* prl_cuda_last_error(const prl_cuda_ctx* ctx);   /* ctx may be NULL: last create() error */
/* Borrow an externally owned cudaStream_t (e.g. torch's current stream) for all later calls;
 * NULL restores the context's own stream; pass cudaStreamLegacy ((void*)0x1) for the legacy
 * default stream. */
int         prl_cuda_set_stream(prl_cuda_ctx* ctx, void* cuda_stream);
int         prl_cuda_synchronize(prl_cuda_ctx* ctx);
/* Upper bound (bytes) for the S/Q scratch planes of one in-flight chunk of pages (default 48 GiB,
 * clipped to 70 % of free HBM at first use). */
int         prl_cuda_set_workspace_limit(prl_cuda_ctx* ctx, size_t bytes);
/* Path selection and validation switches (masks and planes are bit-identical either way; only the speed differs):
 *   "enable_fused"    = 0  : DEFAULT 1.  With 1, windows <= 31 (Sauvola/Niblack/NICK/Feng, padded width <= 8192) run the
 *                            fused strip kernel that never materialises the integral planes in HBM; the rare page it cannot
 *                            finish is handed to kernel 1 + kernel 2 through a page list kept in device memory (no host
 *                            read-back: the *_dev calls stay asynchronous).  0 forces kernel 1 + kernel 2 for every window;
 *   "exact_threshold" != 0 : kernel 2 evaluates the reference's FP64 formula for EVERY pixel instead
 *                            of only for the pixels its exact-integer/FP32 decision cannot settle;
 *   "disable_tma"     != 0 : kernel 1 uses its generic byte-load kernel instead of the TMA-staged one;
 *   "disable_compact" != 0 : kernel 1 writes / kernel 2 reads full int64 planes (16 B per padded pixel) instead of the
 *                            compact layout (low words + sparse high-word anchors, 8.25 B);
 *   "fused_page_cap"  = n  : undecided pixels per page (0..128, default 128) the fused path finishes itself by
 *                            brute force; a page with more is redone by kernel 1 + kernel 2 (test hook: 0);
 *   "fused_no_tier2"  != 0 : the fused path skips its FP64 estimate, so every pixel its FP32 estimate cannot settle
 *                            goes to the brute-force list (test hook for the list and the hand-back);
 *   "median_legacy"   != 0 : cv::medianBlur with kernel sizes 3 and 5 runs the radix-select kernel of the larger sizes
 *                            instead of the selection network (same medians; A/B and test hook);
 *   "gauss_legacy"    != 0 : cv::adaptiveThreshold(GAUSSIAN_C) runs its row and column passes as two kernels over a float32
 *                            plane in HBM for every block size, not only above 63 (same bytes; A/B and test hook). */
int         prl_cuda_set_option(prl_cuda_ctx* ctx, const char* name, long long value);

/* Geometry of the reference's processingRect (binarizeSauvola.cpp:57,66; binarizeWolfJolion.cpp:58,69):
 * Sauvola/Niblack -> (rows+2h-w) x (cols+2h-w); WJ/NICK/Feng -> (rows-w) x (cols-w), with
 * w = min(window, rows, cols), h = w/2.  Returns PRL_E_INVALID for a bad window, PRL_E_EMPTY_ROI
 * when the rect is empty. */
int prl_cuda_output_shape(int method, int rows, int cols, int window, int* out_rows, int* out_cols);

/* ---- kernel 1: replicate-pad + integral images ----------------------------------------
 * Replaces cv::copyMakeBorder + cv::integral(CV_64F) + Rect(1,1,..) crop,
 * binarizeSauvola.cpp:65-77 (same block in Niblack:65-77, WolfJolion:71-85, NICK:71-85, Feng:68-82).
 * sum / sqsum: (rows+2*pad) x (cols+2*pad) INCLUSIVE prefix sums, contiguous int64 (exact). */
int prl_cuda_integral_u8(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step,
                         int pad, int64_t* sum, int64_t* sqsum);

/* ---- kernels 1+2: the five local-statistics binarizers --------------------------------
 * params: Sauvola/Niblack/WJ/NICK {k}; Feng {alpha1, k1, k2, gamma} (always pass 4 doubles).
 * src is single-channel u8 (see prl_cuda_bgr2gray for the cvtColor front step).
 * dst receives out_rows x out_cols u8 0/255 (cv::compare CMP_GT), after the optional
 * morphology tail (morph_iters > 0: dilate xN then erode xN; < 0: erode then dilate;
 * binarizeSauvola.cpp:125-134). */
int prl_cuda_binarize_local(prl_cuda_ctx* ctx, int method, const uint8_t* src, int rows, int cols,
                            size_t step, int window, const double* params, int morph_iters,
                            uint8_t* dst, size_t dst_step, int* out_rows, int* out_cols);

/* Same, for the image exactly as the reference receives it (binarizeSauvola.cpp:49-52): channels = 1, 3 (BGR) or
 * 4 (BGRA).  Multi-channel input is converted on the device (cv::cvtColor BGR2GRAY), so it crosses PCIe once.
 * gray_out (may be NULL) receives the rows x cols gray image -- what the shim needs to reproduce the reference's
 * side effect of leaving the padded gray image in imageInput. */
int prl_cuda_binarize_local_image(prl_cuda_ctx* ctx, int method, const uint8_t* src, int rows, int cols,
                                  size_t step, int channels, int window, const double* params, int morph_iters,
                                  uint8_t* dst, size_t dst_step, int* out_rows, int* out_cols,
                                  uint8_t* gray_out, size_t gray_step);

/* Parity hook: the u8 threshold surface T8 = saturate_cast<uchar>(cvRound(T))
 * (thresholdsValues.convertTo(CV_8UC1), binarizeSauvola.cpp:119).  aux (may be NULL) receives
 * {I_min, s_max} as used by Wolf-Jolion / Feng (binarizeWolfJolion.cpp:115-119). */
int prl_cuda_threshold_map(prl_cuda_ctx* ctx, int method, const uint8_t* src, int rows, int cols,
                           size_t step, int window, const double* params, uint8_t* t8,
                           size_t t8_step, int* out_rows, int* out_cols, double* aux);

/* cv::cvtColor(BGR2GRAY / BGRA2GRAY) front step (binarizeSauvola.cpp:49-52), OpenCV 4.x 8-bit
 * fixed point: (B*3735 + G*19235 + R*9798 + 16384) >> 15.  channels = 3 or 4. */
int prl_cuda_bgr2gray(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step,
                      int channels, uint8_t* dst, size_t dst_step);

/* Morphology tail alone (cv::dilate / cv::erode with the default 3x3 element, n iterations,
 * border ignored), binarizeSauvola.cpp:125-134. */
int prl_cuda_morph(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step,
                   int morph_iters, uint8_t* dst, size_t dst_step);

/* ---- kernel 3: Otsu -------------------------------------------------------------------
 * cv::threshold(src, dst, 128, maxval, THRESH_BINARY|THRESH_OTSU): call sites
 * src/deskew/deskew.cpp:224, src/removeLines.cpp:45, src/imageLibCommon.cpp:295-296. */
int prl_cuda_otsu_threshold(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, int* thr);
int prl_cuda_otsu_global(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step,
                         double maxval, uint8_t* dst, size_t dst_step, int* thr);
/* Per-rectangle Otsu loop of prl::binarizeLocalOtsu, binarizeLocalOtsu.cpp:138-162:
 * dst = 255; for each rect (x,y,w,h): dst(rect) = 0 where (src > otsu(rect) ? maxval : 0) ^ 255 != 0.
 * thr_out (may be NULL) receives the n_rects thresholds. */
int prl_cuda_otsu_rects(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step,
                        const int32_t* xywh, int n_rects, double maxval, uint8_t* dst, size_t dst_step,
                        int32_t* thr_out);
/* Same loop with rects := the regular tile_w x tile_h grid (edge tiles clipped): BASELINE config 4. */
int prl_cuda_otsu_tiles(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step,
                        int tile_w, int tile_h, double maxval, uint8_t* dst, size_t dst_step);

/* ---- device-resident batch entry points (inputs already in HBM; asynchronous) ----------
 * Pages are n_pages images of rows x cols u8 at d_src + p*src_page_stride, row pitch src_step
 * (bytes).  Masks go to d_dst + p*dst_page_stride, row pitch dst_step; they are out_rows x
 * out_cols (prl_cuda_output_shape).  Pitches that are multiples of 16 bytes get the vector path. */
int prl_cuda_binarize_local_batch_dev(prl_cuda_ctx* ctx, int method, const uint8_t* d_src, int n_pages,
                                      int rows, int cols, size_t src_step, size_t src_page_stride,
                                      int window, const double* params, int morph_iters,
                                      uint8_t* d_dst, size_t dst_step, size_t dst_page_stride);
/* Kernel 1 alone over a batch; planes are (rows+2*pad) rows of plane_pitch int64 ELEMENTS. */
int prl_cuda_integral_u8_batch_dev(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols,
                                   size_t src_step, size_t src_page_stride, int pad,
                                   int64_t* d_sum, int64_t* d_sqsum, size_t plane_pitch, size_t plane_page_stride);
/* (batches of 64 pages or more run in groups that alternate between two internal streams, so that one group's histogram
 * pass -- bound by shared-memory atomics -- overlaps the other's apply pass -- bound by HBM; stream-ordered like every *_dev call) */
int prl_cuda_otsu_global_batch_dev(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols,
                                   size_t src_step, size_t src_page_stride, double maxval,
                                   uint8_t* d_dst, size_t dst_step, size_t dst_page_stride, int32_t* d_thr);
int prl_cuda_otsu_tiles_batch_dev(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols,
                                  size_t src_step, size_t src_page_stride, int tile_w, int tile_h, double maxval,
                                  uint8_t* d_dst, size_t dst_step, size_t dst_page_stride);
/* synthpage-v2 generator (bit-identical to oracle/prl_oracle.py:synth_page). */
int prl_cuda_synth_pages_dev(prl_cuda_ctx* ctx, uint8_t* d_dst, int n_pages, int rows, int cols,
                             size_t step, size_t page_stride, uint32_t seed, uint32_t first_page);

/* ---- host batch + page dispatcher -------------------------------------------------------
 * n_pages contiguous rows x cols u8 pages in host memory -> n_pages contiguous out_rows x
 * out_cols masks.  Pages are sharded in contiguous ranges over `devices` (one host thread and one
 * cached worker per device: a context, three streams and a 3-slot ring of DEVICE buffers so that the
 * H2D copy of chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap; no collective).
 * devices == NULL or n_dev <= 0 means "every visible device".
 * Host memory: page-locked buffers (prl_cuda_host_alloc, prl_cuda_host_register, cudaMallocHost, a torch
 * pin_memory() tensor) go to the copy engines directly and are what the quoted end-to-end throughput needs.
 * Pageable buffers are detected (cudaPointerGetAttributes) and staged through library-owned pinned bounce
 * buffers with one memcpy per direction on the worker thread: correct overlap, bound by the host memcpy.
 * Threading: the call takes no context and is re-entrant; concurrent callers that name the same device take
 * turns on that device's worker (a mutex held for the caller's whole shard).  After an error nothing is left in
 * flight on the caller's buffers and the worker is rebuilt on the next call. */
int prl_cuda_binarize_batch(const int* devices, int n_dev, int method, const uint8_t* pages, int n_pages,
                            int rows, int cols, int window, const double* params, int morph_iters,
                            uint8_t* masks);
/* Page-locked host memory for the batch loader ("pinned-memory batch loader" of the north star): alloc/free a
 * portable pinned range, or pin/unpin a range the caller already owns (e.g. the data of a cv::Mat or an mmap). */
int prl_cuda_host_alloc(size_t bytes, void** out);
int prl_cuda_host_free(void* p);
int prl_cuda_host_register(void* p, size_t bytes);
int prl_cuda_host_unregister(void* p);
/* Process-wide options of the ctx-less entry points: "batch_chunk_pages" (pages per ring slot, 0 = automatic,
 * about 72 MiB of input), "batch_stage_pageable" (1 = bounce pageable memory through pinned buffers, default;
 * 0 = hand it to the driver as it is), "batch_unpack_threads" (n > 0: the byte masks of prl_cuda_binarize_batch cross
 * PCIe as 1 bit per pixel into library-owned pinned buffers and n host threads per device expand them into `masks`
 * while later chunks are in flight -- the link carries 1.125 instead of 2 bytes per pixel and `masks` need not be
 * page-locked; 0: the bytes themselves cross; -1, the default: on boxes of 4 or more GPUs always bits with host cores per
 * GPU - 1 threads (2..8), on 1 or 2 GPUs min(8, cores per GPU - 2) threads where that is at least 6, else bytes --
 * [1 B200, 16 cores] 4.8 k A4 pages/s as bytes, 5.9 k as bits with 8 threads, 3.2 k with 4; [8 B200, 32 cores] 7.6 k as
 * bytes, 9.3 k as bits with 3 threads per GPU),
 * "batch_unpack_nt" (1, default: the expansion uses non-temporal stores; 0: ordinary stores -- A/B switch),
 * "batch_unpack_lag" (0, default: the expansion jobs wait for their chunk's copy themselves; 1: the submitting thread
 * waits and hands them out two chunks later -- A/B switch).
 * PRL_E_INVALID for an unknown name. */
int prl_cuda_set_global_option(const char* name, long long value);
/* What "batch_unpack_threads" resolves to right now: host threads per device expanding 1-bit masks, 0 = the bytes cross PCIe. */
int prl_cuda_batch_unpack_threads(void);
/* The host half of that return path alone (test hook): PIX words (prl_cuda_pack_mask_dev layout, wpl = (cols + 31) / 32) to a
 * dense rows x cols 0/255 mask; force_scalar != 0 takes the portable loop instead of the AVX2 one. */
int prl_cuda_unpack_mask_host(const uint32_t* bits, int rows, int cols, uint8_t* mask, int force_scalar);

/* ---- edge front-end of prl::binarizeLocalOtsu -----------------------
 * prl_cuda_canny_edge_detection = CannyEdgeDetection (src/imageLibCommon.cpp:244-324: GaussianBlur k x k sigma 0,
 * Otsu value of the blurred image, cv::Canny(lowerCoeff * upper, upper = upperCoeff * otsu), closing (morph_iters > 0)
 * or opening (< 0)) followed by `post_dilate` dilations (binarizeLocalOtsu.cpp:92 uses 3); single-channel u8 in,
 * 0/255 edge map out, bit-identical to OpenCV 4.x.  PRL_E_INVALID mirrors the reference's std::invalid_argument
 * checks (:248-272), PRL_E_EMPTY_ROI OpenCV's odd-kernel assertion (a cv::Exception); sizes above 63: PRL_E_UNSUPPORTED.  The contour step (cv::findContours, :104-110) stays with the
 * caller; its rectangles go to prl_cuda_otsu_rects.  prl_cuda_gaussian_blur / prl_cuda_canny are the two building
 * blocks alone (cv::GaussianBlur on CV_8U with BORDER_DEFAULT; cv::Canny with aperture 3, L1 gradient). */
/* the 8.8 fixed-point Gaussian coefficients cv::GaussianBlur uses on CV_8U (host-only helper; n odd, <= 63; k[n]) */
int prl_cuda_gauss_kernel_fixed(int n, double sigma, int* k);
int prl_cuda_gaussian_blur(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, int ksize,
                           double sigma, uint8_t* dst, size_t dst_step);
int prl_cuda_canny(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, double low, double high,
                   uint8_t* dst, size_t dst_step);
int prl_cuda_canny_edge_detection(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step,
                                  int gauss_ksize, double upper_coeff, double lower_coeff, int morph_iters,
                                  int post_dilate, uint8_t* dst, size_t dst_step);
int prl_cuda_canny_edge_detection_dev(prl_cuda_ctx* ctx, const uint8_t* d_gray, int rows, int cols, size_t step,
                                      int gauss_ksize, double upper_coeff, double lower_coeff, int morph_iters,
                                      int post_dilate, uint8_t* d_dst, size_t dst_step);

/* EnhanceLocalContrastByCLAHE for one channel (imageLibCommon.cpp:326-346): cv::createCLAHE() (8 x 8 tiles) with the given
 * clip limit, then cv::equalizeHist when `equalize` is non-zero.  Bit-identical to OpenCV 4.x. */
int prl_cuda_clahe(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, double clip_limit, int equalize,
                   uint8_t* dst, size_t dst_step);

/* prl::binarizeLocalOtsu (binarizeLocalOtsu.h:50-57) in one call: clahe_clip_limit > 0 first enhances the gray image
 * (CLAHE + equalizeHist, binarizeLocalOtsu.cpp:79-82); then the edge map above with
 * post_dilate = 3, the bounding rectangles of its top-level contours -- what cv::findContours(RETR_EXTERNAL) +
 * cv::boundingRect return (binarizeLocalOtsu.cpp:104-105,150), found by a union-find labelling on the device -- and
 * the per-rectangle Otsu loop (:138-162).  *n_rects receives the number of rectangles; rects_out (optional, rects_cap
 * x,y,w,h quadruples) a copy of them in no particular order.  PRL_E_INVALID when there is no contour
 * (RemoveChildrenContours throws std::invalid_argument, imageLibCommon.cpp:643-646).  channels 3 or 4: interleaved
 * 8-bit, converted like cv::cvtColor(COLOR_RGB2GRAY) -- the code the reference uses here (binarizeLocalOtsu.cpp:63). */
int prl_cuda_binarize_local_otsu(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, int channels,
                                 double maxval, double clahe_clip_limit, int gauss_ksize, double upper_coeff, double lower_coeff, int morph_iters,
                                 uint8_t* dst, size_t dst_step, int* n_rects, int32_t* rects_out, int rects_cap);

/* The contour step alone: bounding rectangles (x, y, w, h) of the top-level contours of a binary image (non-zero =
 * foreground), the set cv::findContours(RETR_EXTERNAL, ...) + cv::boundingRect give (binarizeLocalOtsu.cpp:104-105,150),
 * in no particular order.  *n_rects = how many exist (may exceed rects_cap; only rects_cap are written). */
int prl_cuda_external_rects(prl_cuda_ctx* ctx, const uint8_t* mask, int rows, int cols, size_t step, int32_t* rects_out,
                            int rects_cap, int* n_rects);

/* ---- prl::removeLines (src/removeLines.cpp:30-77; a Global-Otsu caller) -------------
 * bw = cv::threshold(~gray, OTSU); horizontal / vertical = opening of bw by a 1 x cols/50 / rows/50 x 1 element;
 * out = ~(bw - horizontal - vertical).  channels 1, or 3 = BGR (cvtColor BGR2GRAY, :33-36).  rows < 50 or cols < 50:
 * PRL_E_EMPTY_ROI, the cv::Exception of the reference (a zero-sized structuring element fails cv::erode's anchor assertion). */
int prl_cuda_remove_lines(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, int channels,
                          uint8_t* dst, size_t dst_step);

/* Batched forms over gray pages resident in HBM.  The per-page kernel sequences are
 * dozens of small launches; the library runs 16 pages side by side on internal streams.  Both calls are SYNCHRONOUS (they
 * first wait for the context's stream, and return when every page is done).  status[p] / n_rects[p] (host arrays of n_pages
 * ints, optional): the page's own outcome -- PRL_OK, PRL_E_INVALID = "Contours array is empty" (the std::invalid_argument the
 * reference throws for that image, imageLibCommon.cpp:643-646; its output page is left untouched), PRL_E_UNSUPPORTED = more than
 * 65535 contours -- and its number of rectangles.  prl_cuda_remove_lines_batch_dev fails as a whole with PRL_E_EMPTY_ROI
 * when rows < 50 or cols < 50 (the cv::Exception of the reference). */
int prl_cuda_binarize_local_otsu_batch_dev(prl_cuda_ctx* ctx, const uint8_t* d_gray, int n_pages, int rows, int cols, size_t step,
                                           size_t page_stride, double maxval, double clahe_clip_limit, int gauss_ksize,
                                           double upper_coeff, double lower_coeff, int morph_iters, uint8_t* d_dst, size_t dst_step,
                                           size_t dst_page_stride, int32_t* n_rects, int32_t* status);
int prl_cuda_remove_lines_batch_dev(prl_cuda_ctx* ctx, const uint8_t* d_gray, int n_pages, int rows, int cols, size_t step,
                                    size_t page_stride, uint8_t* d_dst, size_t dst_step, size_t dst_page_stride);

/* ---- the adaptive-mean family: binarizeNativeAdaptive.cpp:34-140, binarizeAT.cpp:33-67,
 * binarizeAGT.cpp:33-58, binarizePureAdaptiveGaussian.cpp:31-71 ------------------------------------------------
 * cv::medianBlur (u8; 1, 3 or 4 interleaved channels; odd ksize 3..63; ksize <= 1 copies; even ksize: PRL_E_EMPTY_ROI, the
 * cv::Exception of OpenCV) and cv::adaptiveThreshold (method 0 = ADAPTIVE_THRESH_MEAN_C, 1 = ADAPTIVE_THRESH_GAUSSIAN_C;
 * type 0 = THRESH_BINARY, 1 = THRESH_BINARY_INV; odd block_size 3..255, else PRL_E_EMPTY_ROI) on one image, byte-identical
 * to OpenCV 4.x; prl_cuda_gauss_kernel_float = cv::getGaussianKernel(n, 0, CV_32F), the coefficients of GAUSSIAN_C. */
int prl_cuda_median_blur(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, int channels, int ksize,
                         uint8_t* dst, size_t dst_step);
int prl_cuda_adaptive_threshold(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, double maxval,
                                int method, int type, int block_size, double delta, uint8_t* dst, size_t dst_step);
int prl_cuda_gauss_kernel_float(int n, float* k);
/* The whole call sequence of one family member in one call (the image crosses PCIe once each way).  The C++ shim fills
 * the structure per reference function; 1 channel in = used as it is, 3 / 4 = BGR / BGRA. */
typedef struct prl_adaptive_params {
    int gray_first;      /* 1: cvtColor(BGR2GRAY), then the blur (NativeAdaptive :58-61); 0: blur the colour image, then gray (AT, AGT) */
    int blur;            /* 0 none, 1 cv::medianBlur(blur_ksize), 2 cv::GaussianBlur(blur_ksize x blur_ksize, blur_sigma) */
    int blur_ksize;
    double blur_sigma;
    int assert_ksize;    /* 1: CV_Assert(medianBlurKernelSize >= 3) of NativeAdaptive :65 */
    int method, type;    /* as prl_cuda_adaptive_threshold */
    double maxval;
    int check_maxval;    /* 1: maxval outside [0, 255] is PRL_E_INVALID (NativeAdaptive :53-56) */
    int block_size;
    int auto_block;      /* 1: block_size < 3 means int(diagonal / 333 + 7) (NativeAdaptive :86-93) */
    double delta;
    int invert_if_dark;  /* 1: 255 - image when cv::mean(image)[0] < 128 (NativeAdaptive :108-111) */
    int bilateral_d;     /* >= 3: cv::bilateralFilter(result, d, sigma_color, sigma_space) as the last step (NativeAdaptive :116-134);
                            a sigma <= 0 is then PRL_E_INVALID, raised after the threshold's own checks like the reference does */
    double bilateral_sigma_color, bilateral_sigma_space;
} prl_adaptive_params;
int prl_cuda_binarize_adaptive(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, int channels,
                               const prl_adaptive_params* params, uint8_t* dst, size_t dst_step);
/* The same over n_pages images of one size resident in HBM (image p at d_src + p * src_page_stride, `channels` interleaved bytes
 * per pixel; result p at d_dst + p * dst_page_stride, one byte per pixel): the per-page kernel sequences run concurrently on the
 * context's page lanes.  Synchronous.  Argument errors as prl_cuda_binarize_adaptive. */
int prl_cuda_binarize_adaptive_batch_dev(prl_cuda_ctx* ctx, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                                         size_t src_page_stride, int channels, const prl_adaptive_params* params, uint8_t* d_dst,
                                         size_t dst_step, size_t dst_page_stride);
/* cv::bilateralFilter(src, dst, d, sigmaColor, sigmaSpace) for CV_8UC1, BORDER_DEFAULT -- OpenCV's own C++ arithmetic (float32
 * weights from its SIMD exponential, taps in raster order; csrc/adaptive.cu spells it out).  The reference applies it to the
 * 0 / maxval mask of binarizeNativeAdaptive (binarizeNativeAdaptive.cpp:129-133).  d <= 0 derives the radius from sigmaSpace. */
int prl_cuda_bilateral_filter(prl_cuda_ctx* ctx, const uint8_t* src, int rows, int cols, size_t step, int d, double sigma_color,
                              double sigma_space, uint8_t* dst, size_t dst_step);

/* ---- 1 bit per pixel ---------------------------------------------
 * The masks in Leptonica's PIX layout, the reference's second image container (src/formatConvert.cpp:39-69
 * writes PIX words with SET_DATA_BIT): rows of wpl = (cols + 31) / 32 32-bit words, pixel x of a row in word
 * x >> 5 at bit 31 - (x & 31), 1 = black (mask byte 0), padding bits 0.  An eighth of the bytes over PCIe.
 * prl_cuda_pack_mask_dev packs masks resident in HBM (d_bits: n_pages * rows * wpl words, dense);
 * prl_cuda_binarize_batch_packed is prl_cuda_binarize_batch with `bits` receiving n_pages * out_rows * wpl words. */
int prl_cuda_pack_mask_dev(prl_cuda_ctx* ctx, const uint8_t* d_mask, int n_pages, int rows, int cols, size_t step,
                           size_t page_stride, uint32_t* d_bits);
int prl_cuda_binarize_batch_packed(const int* devices, int n_dev, int method, const uint8_t* pages, int n_pages,
                                   int rows, int cols, int window, const double* params, int morph_iters,
                                   uint32_t* bits);

/* ---- instrumentation ---------------------------------------------------------------------
 * With timing enabled every kernel launch is bracketed by CUDA events on the launching stream.
 * prl_cuda_timing_get sums them per kernel family ("integral", "threshold", "smax", "morph",
 * "otsu_hist", "otsu_search", "otsu_apply", "otsu_tiles", "synth", "bgr2gray", "band_carry", "fused", "fused_pre", "fused_fix", "pack", "edges", "lines", "adaptive");
 * it synchronizes the stream. */
int  prl_cuda_timing_enable(prl_cuda_ctx* ctx, int on);
int  prl_cuda_timing_reset(prl_cuda_ctx* ctx);
int  prl_cuda_timing_get(prl_cuda_ctx* ctx, const char* family, double* total_ms, long long* launches);
long long prl_cuda_launch_count(const prl_cuda_ctx* ctx);   /* kernels launched since create/reset */
/* pages the fused small-window path handed back to the two-kernel path since create (a device-side counter:
 * this call synchronises the context's stream) */
long long prl_cuda_fused_redo_count(prl_cuda_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* PRLIB_CUDA_H */
