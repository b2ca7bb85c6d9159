// api.cu -- the C-ABI of libprlib_cuda (include/prlib_cuda.h): context, host-pointer drop-in entry
// points and device-resident batch entry points.  The ctx-less host batch loader + page dispatcher lives in batch.cu.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <algorithm>

static thread_local std::string g_create_err;

// ------------------------------------------------------------------------------------------------
// plumbing
// ------------------------------------------------------------------------------------------------
int prl_set_err(prl_cuda_ctx* ctx, int code, const char* what, cudaError_t ce)
{
    std::string m = what ? what : "";
    if (ce != cudaSuccess) { m += ": "; m += cudaGetErrorString(ce); }
    if (ctx) ctx->err = m; else g_create_err = m;
    return code;
}

int prl_ensure(prl_cuda_ctx* ctx, void** ptr, size_t* have, size_t need)
{
    if (need <= *have && *ptr) return PRL_OK;
    if (*ptr) {
        // earlier launches on this stream may still read the old buffer
        PRL_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        cudaFree(*ptr); *ptr = nullptr; *have = 0;
    }
    if (need == 0) need = 256;
    cudaError_t e = cudaMalloc(ptr, need);
    if (e != cudaSuccess) { *ptr = nullptr; return prl_set_err(ctx, PRL_E_NOMEM, "cudaMalloc", e); }
    *have = need;
    return PRL_OK;
}

void prl_launch_begin(prl_cuda_ctx* ctx, int family)
{
    ctx->launches++;
    if (!ctx->timing) return;
    prl_timing_rec r; r.family = family;
    cudaEvent_t ev[2];
    for (int i = 0; i < 2; ++i) {
        if (!ctx->event_pool.empty()) { ev[i] = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
        else cudaEventCreate(&ev[i]);
    }
    r.a = ev[0]; r.b = ev[1];
    cudaEventRecord(r.a, ctx->stream);
    ctx->recs.push_back(r);
}

void prl_launch_end(prl_cuda_ctx* ctx)
{
    if (!ctx->timing || ctx->recs.empty()) return;
    cudaEventRecord(ctx->recs.back().b, ctx->stream);
}

static void timing_collect(prl_cuda_ctx* ctx)
{
    if (ctx->recs.empty()) return;
    cudaStreamSynchronize(ctx->stream);
    for (auto& r : ctx->recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto& t = ctx->totals[r.family];
            t.first += ms; t.second += 1;
        }
        ctx->event_pool.push_back(r.a); ctx->event_pool.push_back(r.b);
    }
    ctx->recs.clear();
}

static const char* kFamilyNames[FAM_COUNT] = {"integral", "threshold", "smax", "morph", "otsu_hist", "otsu_search",
                                              "otsu_apply", "otsu_tiles", "synth", "bgr2gray", "band_carry", "fused", "fused_pre", "fused_fix", "pack", "edges", "lines", "adaptive"};

int prl_make_geom(int method, int rows, int cols, int window, prl_geom* g)
{
    if (rows <= 0 || cols <= 0) return PRL_E_INVALID;                       // imageInput.empty()  binarizeSauvola.cpp:38-41
    if (!((window > 1) && ((window % 2) == 1))) return PRL_E_INVALID;       // binarizeSauvola.cpp:43-47
    if (method < PRL_SAUVOLA || method > PRL_FENG) return PRL_E_INVALID;
    g->rows = rows; g->cols = cols;
    g->w = std::min(window, std::min(cols, rows));                          // :57
    g->h = g->w / 2; g->d = g->w - 1;
    g->Hp = rows + 2 * g->h; g->Wp = cols + 2 * g->h;                       // copyMakeBorder :65
    if (method == PRL_SAUVOLA || method == PRL_NIBLACK) {                   // rect after padding :66
        g->out_cols = g->Wp - g->w; g->out_rows = g->Hp - g->w;
    } else {                                                                // rect before padding (WolfJolion.cpp:69)
        g->out_cols = cols - g->w; g->out_rows = rows - g->w;
    }
    g->pitch = prl_plane_pitch(g->Wp);
    if (g->out_cols <= 0 || g->out_rows <= 0) return PRL_E_EMPTY_ROI;
    return PRL_OK;
}


// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int prl_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int prl_cuda_create(int device, prl_cuda_ctx** out)
{
    if (!out) return PRL_E_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return prl_set_err(nullptr, PRL_E_CUDA, "no CUDA device available (libprlib_cuda has no CPU fallback)", e);
    }
    if (device < 0 || device >= n) return prl_set_err(nullptr, PRL_E_INVALID, "bad device index");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return prl_set_err(nullptr, PRL_E_CUDA, "cudaSetDevice", e);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return prl_set_err(nullptr, PRL_E_CUDA, "cudaGetDeviceProperties", e);
    if (prop.major != 10) {
        char buf[160];
        snprintf(buf, sizeof buf, "device %d is sm_%d%d; libprlib_cuda is built for sm_100a only", device, prop.major, prop.minor);
        return prl_set_err(nullptr, PRL_E_CUDA, buf);
    }
    prl_cuda_ctx* c = new prl_cuda_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return prl_set_err(nullptr, PRL_E_CUDA, "cudaStreamCreate", e); }
    c->stream = c->own_stream;
    *out = c;
    return PRL_OK;
}

extern "C" void prl_cuda_destroy(prl_cuda_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (prl_cuda_ctx* l : c->lanes) prl_cuda_destroy(l);
    c->lanes.clear();
    for (int i = 0; i < 3; ++i) if (c->lane_ev[i]) cudaEventDestroy(c->lane_ev[i]);
    if (c->h_lane_counts) cudaFreeHost(c->h_lane_counts);
    for (auto& r : c->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto& ev : c->event_pool) cudaEventDestroy(ev);
    cudaFree(c->sched); cudaFree(c->planes); cudaFree(c->carry); cudaFree(c->colsum); cudaFree(c->scalars); cudaFree(c->fused_ws);
    cudaFree(c->d_redo_total);
    cudaFree(c->d_in); cudaFree(c->d_out); cudaFree(c->d_res); cudaFree(c->d_tmp); cudaFree(c->d_misc); cudaFree(c->d_bgr); cudaFree(c->edges_ws); cudaFree(c->adaptive_ws); cudaFree(c->rects_ws); cudaFree(c->clahe_ws);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

extern "C" const char* prl_cuda_last_error(const prl_cuda_ctx* c) { return c ? c->err.c_str() : g_create_err.c_str(); }

extern "C" int prl_cuda_set_stream(prl_cuda_ctx* c, void* s)
{
    if (!c) return PRL_E_INVALID;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return PRL_OK;
}

extern "C" int prl_cuda_synchronize(prl_cuda_ctx* c)
{
    if (!c) return PRL_E_INVALID;
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

extern "C" int prl_cuda_set_workspace_limit(prl_cuda_ctx* c, size_t bytes)
{
    if (!c || bytes == 0) return PRL_E_INVALID;
    c->workspace_limit = bytes;
    return PRL_OK;
}

extern "C" int prl_cuda_set_option(prl_cuda_ctx* c, const char* name, long long value)
{
    if (!c || !name) return PRL_E_INVALID;
    if (strcmp(name, "exact_threshold") == 0) c->force_exact = value != 0;
    else if (strcmp(name, "disable_tma") == 0) c->no_tma = value != 0;
    else if (strcmp(name, "disable_compact") == 0) c->no_compact = value != 0;
    else if (strcmp(name, "thr_legacy") == 0) c->thr_legacy = value != 0;
    else if (strcmp(name, "thr_no_tma") == 0) c->thr_no_tma = value != 0;
    else if (strcmp(name, "dbg_skip_exact") == 0) c->dbg_skip_exact = value != 0;
    else if (strcmp(name, "thr_stages") == 0) c->thr_stages = value == 3 ? 3 : 2;
    else if (strcmp(name, "k1_bands") == 0) c->k1_bands = (int)std::min<long long>(std::max<long long>(value, 0), 64);
    else if (strcmp(name, "enable_fused") == 0) c->use_fused = value != 0;
    else if (strcmp(name, "morph_bytes") == 0) c->morph_bytes = value != 0;
    else if (strcmp(name, "thr_rows") == 0) c->thr_rows = value <= 0 ? 0 : ((int)std::min<long long>(std::max<long long>(value, 2), 64) & ~1);
    else if (strcmp(name, "tiles_legacy") == 0) c->tiles_legacy = (value == 2 || value == 3) ? (int)value : (value != 0);
    else if (strcmp(name, "median_legacy") == 0) c->median_legacy = value != 0;
    else if (strcmp(name, "gauss_legacy") == 0) c->gauss_legacy = value != 0;
    else if (strcmp(name, "otsu_group") == 0) c->otsu_group = (int)std::min<long long>(std::max<long long>(value, -1), 65535);
    else if (strcmp(name, "tile_prefetch") == 0) c->tile_prefetch = (int)std::min<long long>(std::max<long long>(value, 0), 31);
    else if (strcmp(name, "fused_no_tier2") == 0) c->fused_no_tier2 = value != 0;
    else if (strcmp(name, "fused_page_cap") == 0) c->fused_page_cap = (int)std::min<long long>(std::max<long long>(value, 0), 128);
    else return prl_set_err(c, PRL_E_INVALID, "unknown option");
    return PRL_OK;
}

extern "C" int prl_cuda_output_shape(int method, int rows, int cols, int window, int* out_rows, int* out_cols)
{
    prl_geom g;
    int rc = prl_make_geom(method, rows, cols, window, &g);
    if (rc == PRL_E_INVALID) return rc;
    if (out_rows) *out_rows = g.out_rows;
    if (out_cols) *out_cols = g.out_cols;
    return rc;
}

// (the page lanes of a context are timed and counted with it)
extern "C" int prl_cuda_timing_enable(prl_cuda_ctx* c, int on)
{
    if (!c) return PRL_E_INVALID;
    timing_collect(c); c->timing = on != 0;
    for (prl_cuda_ctx* l : c->lanes) { timing_collect(l); l->timing = on != 0; }
    return PRL_OK;
}
extern "C" int prl_cuda_timing_reset(prl_cuda_ctx* c)
{
    if (!c) return PRL_E_INVALID;
    timing_collect(c); c->totals.clear(); c->launches = 0;
    for (prl_cuda_ctx* l : c->lanes) { timing_collect(l); l->totals.clear(); l->launches = 0; }
    return PRL_OK;
}
extern "C" int prl_cuda_timing_get(prl_cuda_ctx* c, const char* family, double* total_ms, long long* launches)
{
    if (!c || !family) return PRL_E_INVALID;
    cudaSetDevice(c->device);
    timing_collect(c);
    for (prl_cuda_ctx* l : c->lanes) timing_collect(l);
    for (int f = 0; f < FAM_COUNT; ++f)
        if (strcmp(family, kFamilyNames[f]) == 0) {
            double ms = 0.0; long long n = 0;
            auto it = c->totals.find(f);
            if (it != c->totals.end()) { ms += it->second.first; n += it->second.second; }
            for (prl_cuda_ctx* l : c->lanes) {
                auto jt = l->totals.find(f);
                if (jt != l->totals.end()) { ms += jt->second.first; n += jt->second.second; }
            }
            if (total_ms) *total_ms = ms;
            if (launches) *launches = n;
            return PRL_OK;
        }
    return prl_set_err(c, PRL_E_INVALID, "unknown kernel family");
}
extern "C" long long prl_cuda_launch_count(const prl_cuda_ctx* c)
{
    if (!c) return 0;
    long long n = c->launches;
    for (const prl_cuda_ctx* l : c->lanes) n += l->launches;
    return n;
}
// pages the fused path handed back to the two-kernel path so far: a device counter, so this call synchronises the stream
extern "C" long long prl_cuda_fused_redo_count(prl_cuda_ctx* c)
{
    if (!c || !c->d_redo_total) return 0;
    unsigned long long v = 0;
    if (cudaSetDevice(c->device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess ||
        cudaMemcpy(&v, c->d_redo_total, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (long long)v;
}

// ------------------------------------------------------------------------------------------------
// device-resident batch entry points
// ------------------------------------------------------------------------------------------------
// Bytes the S/Q planes of one in-flight chunk may take.  cudaMemGetInfo is only consulted when the
// planes have to grow (it takes the driver's allocator lock and was seen to stall for tens of ms).
static size_t planes_budget(prl_cuda_ctx* c, size_t want)
{
    if (want <= c->planes_bytes && c->planes) return c->planes_bytes;
    size_t lim = c->workspace_limit;
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
        size_t avail = fr + c->planes_bytes;          // what we already hold counts as available to us
        lim = std::min(lim, (size_t)(avail * 0.7));
    }
    return std::max(lim, c->planes_bytes);
}

// The two-kernel path over a batch: kernel 1 (integral planes) -> kernel 2 (threshold) [-> morphology], in chunks of
// as many pages as the planes scratch holds.  mode 0: masks (+morph), mode 1: T8 maps
static int planes_pages(prl_cuda_ctx* c, int method, int mode, const uint8_t* d_src, int n_pages, const prl_geom& g,
                        size_t src_step, size_t src_page_stride, const double* params, int morph_iters,
                        uint8_t* d_dst, size_t dst_step, size_t dst_page_stride)
{
    // Plane layout (common.cuh: prl_planes).  The mask path of aligned batches takes the compact one: interleaved low words
    // + sparse anchor high words, 8.5 instead of 16 bytes per padded pixel written by kernel 1 and read by kernel 2.
    const bool compact = mode == 0 && !c->no_compact && prl_threshold_fast_ok(c, method, params, g, d_src, src_step, src_page_stride) &&
                         prl_integral_compact_ok(c, d_src, src_step, src_page_stride, g.rows, g.cols, g.h);
    prl_planes P;
    P.compact = compact ? 1 : 0;
    P.pitch = g.pitch;                                          // elements per row: int64, or uint2 {S lo, Q lo}
    P.ashift = compact ? prl_anchor_shift(g.Hp, g.Wp) : 0;
    const size_t plane_elems = (size_t)g.Hp * P.pitch;          // one plane of one page
    const size_t anchor_rows = compact ? (((size_t)g.Hp + ((size_t)1 << P.ashift) - 1) >> P.ashift) : 0;
    P.a_pitch = compact ? P.pitch / 4 : 0;
    const size_t anchor_elems = anchor_rows * P.a_pitch;
    const size_t per_page = compact ? (plane_elems + anchor_elems) * sizeof(uint2) : 2 * plane_elems * sizeof(int64_t);
    const size_t budget = std::min(planes_budget(c, (size_t)n_pages * per_page), c->workspace_limit);
    int chunk = (int)std::min<size_t>((size_t)n_pages, std::max<size_t>(1, budget / per_page));
    int rc;
    if ((size_t)chunk * per_page > c->planes_bytes || !c->planes) {
        rc = prl_ensure(c, (void**)&c->planes, &c->planes_bytes, (size_t)chunk * per_page);
        if (rc) return rc;
    }
    rc = prl_ensure(c, &c->scalars, &c->scalars_bytes, (size_t)n_pages * 16);
    if (rc) return rc;
    uint32_t* d_imin = (uint32_t*)c->scalars;
    long long* d_smax = (long long*)((uint8_t*)c->scalars + (size_t)n_pages * 8);

    size_t tmp_step = round16((size_t)g.out_cols);
    if (morph_iters != 0 && mode == 0) {
        rc = prl_ensure(c, (void**)&c->d_tmp, &c->d_tmp_bytes, (size_t)chunk * g.out_rows * tmp_step);
        if (rc) return rc;
    }

    P.page_stride = plane_elems;
    P.a_page_stride = anchor_elems;
    if (compact) {
        P.S = c->planes;
        P.AS = (uint2*)c->planes + (size_t)chunk * plane_elems;
    } else {
        P.S = c->planes;
        P.Q = c->planes + (size_t)chunk * plane_elems;
    }
    for (int p0 = 0; p0 < n_pages; p0 += chunk) {
        const int np = std::min(chunk, n_pages - p0);
        const uint8_t* src = d_src + (size_t)p0 * src_page_stride;
        uint8_t* dst = d_dst + (size_t)p0 * dst_page_stride;
        // the page minimum (cv::minMaxLoc, binarizeWolfJolion.cpp:115-116 / binarizeFeng.cpp) is fused into kernel 1 only when needed
        const bool need_min = method == PRL_WOLFJOLION || method == PRL_FENG;
        rc = prl_k_integral_planes(c, src, np, g.rows, g.cols, src_step, src_page_stride, g.h, P, need_min ? d_imin + p0 : nullptr);
        if (rc) return rc;
        const bool with_morph = morph_iters != 0 && mode == 0;
        // with a morphology tail kernel 2 writes the raw mask into the scratch and the tail writes the final one
        rc = prl_k_threshold(c, method, mode, src, np, g, src_step, src_page_stride, P, params,
                             d_imin + p0, d_smax + p0, with_morph ? c->d_tmp : dst, with_morph ? tmp_step : dst_step,
                             with_morph ? (size_t)g.out_rows * tmp_step : dst_page_stride);
        if (rc) return rc;
        if (with_morph) {
            rc = prl_k_morph(c, c->d_tmp, dst, np, g.out_rows, g.out_cols, tmp_step, (size_t)g.out_rows * tmp_step,
                             dst_step, dst_page_stride, morph_iters, true);
            if (rc) return rc;
        }
    }
    return PRL_OK;
}

// Hand-back of the fused path, without the host: the pages whose undecided-pixel list overflowed (rare: large exactly
// flat areas whose threshold sits on a rounding boundary) are listed in device memory (fused.cu: overflow_list_kernel);
// kernel 1 (generic, int64 planes) and kernel 2 (literal FP64 for every pixel) are launched OVER THAT LIST -- grid slot z
// works on page map[z], slots beyond the list's length leave at once -- in rounds of as many plane slots as the scratch
// holds (at most 32: 4.5 GB for A4), enough rounds to cover the case that every page overflowed.  Nothing is read back,
// so the *_dev entry points stay asynchronous and the batch loader keeps its three streams overlapped; the price is two
// near-empty launches per 32 pages (~3 us each).
static int fused_hand_back(prl_cuda_ctx* c, int method, const uint8_t* d_src, int n_pages, const prl_geom& g, size_t src_step,
                           size_t src_page_stride, const double* params, const uint32_t* d_imin, uint8_t* d_dst, size_t dst_step,
                           size_t dst_page_stride, const int* d_map, const int* d_cnt)
{
    const size_t plane_elems = (size_t)g.Hp * g.pitch;
    const size_t per_page = 2 * plane_elems * sizeof(int64_t);
    const size_t want = (size_t)std::min(n_pages, 32) * per_page;
    const size_t budget = std::min(planes_budget(c, want), c->workspace_limit);
    const int slots = (int)std::min<size_t>((size_t)std::min(n_pages, 32), std::max<size_t>(1, budget / per_page));
    if ((size_t)slots * per_page > c->planes_bytes || !c->planes) {
        int rc = prl_ensure(c, (void**)&c->planes, &c->planes_bytes, (size_t)slots * per_page);
        if (rc) return rc;
    }
    prl_planes P;
    P.compact = 0; P.pitch = g.pitch; P.page_stride = plane_elems;
    P.S = c->planes; P.Q = c->planes + (size_t)slots * plane_elems;
    for (int base = 0; base < n_pages; base += slots) {
        const int ns = std::min(slots, n_pages - base);
        int rc = prl_k_integral_indirect(c, d_src, ns, g.rows, g.cols, src_step, src_page_stride, g.h, (int64_t*)P.S, (int64_t*)P.Q, P.pitch,
                                         P.page_stride, d_map, d_cnt, base);
        if (rc) return rc;
        rc = prl_k_threshold_exact_indirect(c, method, d_src, ns, g, src_step, src_page_stride, P, params, d_imin, d_dst, dst_step,
                                            dst_page_stride, d_map, d_cnt, base);
        if (rc) return rc;
    }
    return PRL_OK;
}

static int local_batch_dev(prl_cuda_ctx* c, int method, int mode, const uint8_t* d_src, int n_pages, int rows, int cols,
                           size_t src_step, size_t src_page_stride, int window, const double* params, int morph_iters,
                           uint8_t* d_dst, size_t dst_step, size_t dst_page_stride)
{
    if (!c || !d_src || !d_dst || !params || n_pages <= 0) return prl_set_err(c, PRL_E_INVALID, "null pointer or empty batch");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    prl_geom g;
    int rc = prl_make_geom(method, rows, cols, window, &g);
    if (rc) return prl_set_err(c, rc, rc == PRL_E_EMPTY_ROI ? "empty processingRect: min(rows, cols) <= windowSize"
                                                             : "empty image or window not (>1 and odd)");
    if ((size_t)g.out_cols > dst_step) return prl_set_err(c, PRL_E_INVALID, "dst_step smaller than out_cols");

    // small windows: the fused path (planes never reach HBM).  It is optimistic: pages on which too many pixels need
    // the reference's exact FP64 arithmetic (large exactly-flat areas under Niblack / Feng) are reported back and redone
    // by the two-kernel path below.
    if (mode == 0 && prl_fused_eligible(c, method, n_pages, g, params)) {
        rc = prl_ensure(c, &c->scalars, &c->scalars_bytes, (size_t)n_pages * 16);
        if (rc) return rc;
        uint32_t* imin = (uint32_t*)c->scalars;
        const bool with_morph = morph_iters != 0;
        const size_t t_step = round16((size_t)g.out_cols), t_page = (size_t)g.out_rows * t_step;
        if (with_morph) {
            rc = prl_ensure(c, (void**)&c->d_tmp, &c->d_tmp_bytes, (size_t)n_pages * t_page);
            if (rc) return rc;
        }
        uint8_t* raw = with_morph ? c->d_tmp : d_dst;
        const size_t raw_step = with_morph ? t_step : dst_step, raw_page = with_morph ? t_page : dst_page_stride;
        const int* d_map = nullptr; const int* d_cnt = nullptr;
        rc = prl_k_fused(c, method, d_src, n_pages, g, src_step, src_page_stride, params, imin, raw, raw_step, raw_page, &d_map, &d_cnt);
        if (rc) return rc;
        rc = fused_hand_back(c, method, d_src, n_pages, g, src_step, src_page_stride, params, imin, raw, raw_step, raw_page, d_map, d_cnt);
        if (rc) return rc;
        if (with_morph)
            rc = prl_k_morph(c, c->d_tmp, d_dst, n_pages, g.out_rows, g.out_cols, t_step, t_page, dst_step, dst_page_stride,
                             morph_iters, true);
        return rc;
    }
    return planes_pages(c, method, mode, d_src, n_pages, g, src_step, src_page_stride, params, morph_iters, d_dst, dst_step,
                        dst_page_stride);
}

extern "C" int prl_cuda_binarize_local_batch_dev(prl_cuda_ctx* c, int method, const uint8_t* d_src, int n_pages,
                                                 int rows, int cols, size_t src_step, size_t src_page_stride,
                                                 int window, const double* params, int morph_iters,
                                                 uint8_t* d_dst, size_t dst_step, size_t dst_page_stride)
{
    return local_batch_dev(c, method, 0, d_src, n_pages, rows, cols, src_step, src_page_stride, window, params,
                           morph_iters, d_dst, dst_step, dst_page_stride);
}

extern "C" int prl_cuda_integral_u8_batch_dev(prl_cuda_ctx* c, const uint8_t* d_src, int n_pages, int rows, int cols,
                                              size_t src_step, size_t src_page_stride, int pad,
                                              int64_t* d_sum, int64_t* d_sqsum, size_t plane_pitch, size_t plane_page_stride)
{
    if (!c || !d_src || !d_sum || !d_sqsum || n_pages <= 0 || rows <= 0 || cols <= 0 || pad < 0)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    if (plane_pitch < (size_t)(cols + 2 * pad) + ((cols + 2 * pad) & 1) || (plane_pitch & 1))
        return prl_set_err(c, PRL_E_INVALID, "plane_pitch must be even and >= padded width rounded up to even");
    if ((((uintptr_t)d_sum) | ((uintptr_t)d_sqsum)) & 15 || (plane_page_stride & 1))
        return prl_set_err(c, PRL_E_INVALID, "planes must be 16-byte aligned");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    return prl_k_integral(c, d_src, n_pages, rows, cols, src_step, src_page_stride, pad, d_sum, d_sqsum, plane_pitch,
                          plane_page_stride, nullptr);
}

static int ensure_lanes(prl_cuda_ctx* c, int n);

// Global Otsu over a batch: the histogram pass is bound by the shared-memory atomic unit (3.4 TB/s of HBM traffic), the
// apply pass by HBM itself (5.7 TB/s) -- back to back they leave each other's resource idle.  Large batches are therefore
// cut into groups that alternate between two page lanes (own streams), staggered by a short first group so that one
// lane's histogram pass runs beside the other lane's apply pass.  Stream-ordered like every *_dev call: the lanes wait
// for the context's stream at entry, the context's stream waits for the lanes at exit.
extern "C" int prl_cuda_otsu_global_batch_dev(prl_cuda_ctx* c, const uint8_t* d_src, int n_pages, int rows, int cols,
                                              size_t src_step, size_t src_page_stride, double maxval,
                                              uint8_t* d_dst, size_t dst_step, size_t dst_page_stride, int32_t* d_thr)
{
    if (!c || !d_src || !d_dst || !d_thr || n_pages <= 0 || rows <= 0 || cols <= 0) return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    // pages per group; 0 = one launch sequence for the whole batch.  [B200] 1024 A4 pages: 5.77 ms in one sequence, 5.16 ms in
    // groups of 256 (0.71 -> 0.79 of the HBM peak on the 3 H W figure; 256 pages 0.67 -> 0.75 in groups of 64); limiting the histogram pass to 2-4 CTAs per SM so
    // that the other lane's CTAs find room was measured too and is slower
    const int group = c->otsu_group < 0 ? std::min(256, std::max(16, n_pages / 4)) : c->otsu_group;   // measured best: n / 4
    if (group <= 0 || n_pages < 4 * group)
        return prl_k_otsu_global(c, d_src, n_pages, rows, cols, src_step, src_page_stride, maxval, d_dst, dst_step,
                                 dst_page_stride, d_thr, true);
    int rc = ensure_lanes(c, 2); if (rc) return rc;
    if (!c->lane_ev[0])
        for (int i = 0; i < 3; ++i) PRL_CUDA_TRY(c, cudaEventCreateWithFlags(&c->lane_ev[i], cudaEventDisableTiming));
    PRL_CUDA_TRY(c, cudaEventRecord(c->lane_ev[2], c->stream));
    for (int i = 0; i < 2; ++i) PRL_CUDA_TRY(c, cudaStreamWaitEvent(c->lanes[i]->stream, c->lane_ev[2], 0));
    int p = 0, turn = 0;
    while (p < n_pages) {
        const int np = std::min(n_pages - p, (p == 0) ? std::max(1, group / 2) : group);      // short first group = the stagger
        prl_cuda_ctx* l = c->lanes[turn & 1];
        rc = prl_k_otsu_global(l, d_src + (size_t)p * src_page_stride, np, rows, cols, src_step, src_page_stride, maxval,
                               d_dst + (size_t)p * dst_page_stride, dst_step, dst_page_stride, d_thr + p, true);
        if (rc) { c->err = l->err; break; }
        p += np; ++turn;
    }
    for (int i = 0; i < 2; ++i) {
        cudaEventRecord(c->lane_ev[i], c->lanes[i]->stream);
        cudaStreamWaitEvent(c->stream, c->lane_ev[i], 0);
    }
    return rc;
}

extern "C" int prl_cuda_otsu_tiles_batch_dev(prl_cuda_ctx* c, const uint8_t* d_src, int n_pages, int rows, int cols,
                                             size_t src_step, size_t src_page_stride, int tile_w, int tile_h, double maxval,
                                             uint8_t* d_dst, size_t dst_step, size_t dst_page_stride)
{
    if (!c || !d_src || !d_dst || n_pages <= 0 || rows <= 0 || cols <= 0 || tile_w <= 0 || tile_h <= 0)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    if ((long long)tile_w * tile_h > 0x7fffffffLL) return prl_set_err(c, PRL_E_UNSUPPORTED, "tile too large");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    return prl_k_otsu_tiles(c, d_src, n_pages, rows, cols, src_step, src_page_stride, tile_w, tile_h, maxval, d_dst,
                            dst_step, dst_page_stride);
}

extern "C" int prl_cuda_synth_pages_dev(prl_cuda_ctx* c, uint8_t* d_dst, int n_pages, int rows, int cols,
                                        size_t step, size_t page_stride, uint32_t seed, uint32_t first_page)
{
    if (!c || !d_dst || n_pages <= 0 || rows <= 0 || cols <= 0) return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    return prl_k_synth(c, d_dst, n_pages, rows, cols, step, page_stride, seed, first_page);
}

// ------------------------------------------------------------------------------------------------
// host-pointer (drop-in) entry points: stage -> kernels -> stage back, synchronous
// ------------------------------------------------------------------------------------------------
static int stage_in(prl_cuda_ctx* c, const uint8_t* src, int rows, size_t width_bytes, size_t step, size_t* dstep)
{
    *dstep = round16(width_bytes);
    int rc = prl_ensure(c, (void**)&c->d_in, &c->d_in_bytes, *dstep * rows);
    if (rc) return rc;
    PRL_CUDA_TRY(c, copy2d(c->d_in, *dstep, src, step, width_bytes, rows, cudaMemcpyHostToDevice, c->stream));
    return PRL_OK;
}

static int local_host(prl_cuda_ctx* c, int method, int mode, const uint8_t* src, int rows, int cols, size_t step,
                      int window, const double* params, int morph_iters, uint8_t* dst, size_t dst_step,
                      int* out_rows, int* out_cols, double* aux, int channels = 1, uint8_t* gray_out = nullptr,
                      size_t gray_step = 0)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || !params) return prl_set_err(c, PRL_E_INVALID, "null pointer");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    prl_geom g;
    int rc = prl_make_geom(method, rows, cols, window, &g);
    if (rc) return prl_set_err(c, rc, rc == PRL_E_EMPTY_ROI ? "empty processingRect: min(rows, cols) <= windowSize"
                                                             : "empty image or window not (>1 and odd)");
    if (channels != 1 && channels != 3 && channels != 4) return prl_set_err(c, PRL_E_INVALID, "channels must be 1, 3 or 4");
    if (step < (size_t)cols * channels || dst_step < (size_t)g.out_cols || (gray_out && gray_step < (size_t)cols))
        return prl_set_err(c, PRL_E_INVALID, "step smaller than row width");
    size_t in_step;
    if (channels == 1) {
        rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    } else {
        // cvtColor(BGR2GRAY) on the device: the 3/4-channel image crosses PCIe once, the gray one never goes back
        // unless the caller asks for it (binarizeSauvola.cpp:49-52)
        const size_t bstep = round16((size_t)cols * channels);
        in_step = round16(cols);
        rc = prl_ensure(c, (void**)&c->d_bgr, &c->d_bgr_bytes, bstep * rows); if (rc) return rc;
        rc = prl_ensure(c, (void**)&c->d_in, &c->d_in_bytes, in_step * rows); if (rc) return rc;
        PRL_CUDA_TRY(c, copy2d(c->d_bgr, bstep, src, step, (size_t)cols * channels, rows, cudaMemcpyHostToDevice, c->stream));
        rc = prl_k_bgr2gray(c, c->d_bgr, rows, cols, bstep, channels, c->d_in, in_step); if (rc) return rc;
    }
    if (gray_out)
        PRL_CUDA_TRY(c, copy2d(gray_out, gray_step, c->d_in, in_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    // dense device mask when the caller's rows are dense too (a continuous cv::Mat): one linear D2H
    // (library-side staging -- the image copied into pinned memory in bands by a pool of host threads, the mask back as bits and
    // expanded by them -- was measured for this call and is slower than the driver's own staged copies at every thread count:
    // A4 gray 0.99 ms with the driver, 1.24 ms with 8 threads, 2.1 ms with 2; the pool's wake-ups cost more than they save)
    const size_t o_step = (dst_step == (size_t)g.out_cols) ? (size_t)g.out_cols : round16(g.out_cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * g.out_rows + 16); if (rc) return rc;
    rc = local_batch_dev(c, method, mode, c->d_in, 1, rows, cols, in_step, in_step * rows, window, params, morph_iters,
                         c->d_out, o_step, o_step * g.out_rows);
    if (rc) return rc;
    PRL_CUDA_TRY(c, copy2d(dst, dst_step, c->d_out, o_step, g.out_cols, g.out_rows, cudaMemcpyDeviceToHost, c->stream));
    if (aux) {
        uint32_t imin = 0; long long smax = 0;
        PRL_CUDA_TRY(c, cudaMemcpyAsync(&imin, c->scalars, 4, cudaMemcpyDeviceToHost, c->stream));
        PRL_CUDA_TRY(c, cudaMemcpyAsync(&smax, (uint8_t*)c->scalars + 8, 8, cudaMemcpyDeviceToHost, c->stream));
        PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
        aux[0] = (double)imin;
        double s; memcpy(&s, &smax, 8); aux[1] = (method == PRL_WOLFJOLION) ? s : NAN;
    }
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (out_rows) *out_rows = g.out_rows;
    if (out_cols) *out_cols = g.out_cols;
    return PRL_OK;
}

extern "C" int prl_cuda_binarize_local(prl_cuda_ctx* c, int method, const uint8_t* src, int rows, int cols,
                                       size_t step, int window, const double* params, int morph_iters,
                                       uint8_t* dst, size_t dst_step, int* out_rows, int* out_cols)
{
    return local_host(c, method, 0, src, rows, cols, step, window, params, morph_iters, dst, dst_step, out_rows, out_cols, nullptr);
}

extern "C" int prl_cuda_binarize_local_image(prl_cuda_ctx* c, int method, const uint8_t* src, int rows, int cols,
                                             size_t step, int channels, int window, const double* params,
                                             int morph_iters, uint8_t* dst, size_t dst_step, int* out_rows,
                                             int* out_cols, uint8_t* gray_out, size_t gray_step)
{
    return local_host(c, method, 0, src, rows, cols, step, window, params, morph_iters, dst, dst_step, out_rows, out_cols,
                      nullptr, channels, gray_out, gray_step);
}

extern "C" int prl_cuda_threshold_map(prl_cuda_ctx* c, int method, const uint8_t* src, int rows, int cols,
                                      size_t step, int window, const double* params, uint8_t* t8,
                                      size_t t8_step, int* out_rows, int* out_cols, double* aux)
{
    return local_host(c, method, 1, src, rows, cols, step, window, params, 0, t8, t8_step, out_rows, out_cols, aux);
}

extern "C" int prl_cuda_integral_u8(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step,
                                    int pad, int64_t* sum, int64_t* sqsum)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !sum || !sqsum || rows <= 0 || cols <= 0 || pad < 0 || step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    const int Hp = rows + 2 * pad, Wp = cols + 2 * pad;
    const size_t pitch = prl_plane_pitch(Wp), plane = (size_t)Hp * pitch;
    size_t in_step;
    int rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    rc = prl_ensure(c, (void**)&c->planes, &c->planes_bytes, 2 * plane * sizeof(int64_t)); if (rc) return rc;
    rc = prl_k_integral(c, c->d_in, 1, rows, cols, in_step, in_step * rows, pad, c->planes, c->planes + plane, pitch, plane, nullptr);
    if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(sum, (size_t)Wp * 8, c->planes, pitch * 8, (size_t)Wp * 8, Hp, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(sqsum, (size_t)Wp * 8, c->planes + plane, pitch * 8, (size_t)Wp * 8, Hp, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

extern "C" int prl_cuda_bgr2gray(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step,
                                 int channels, uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || (channels != 3 && channels != 4) || step < (size_t)cols * channels ||
        dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, (size_t)cols * channels, step, &in_step); if (rc) return rc;
    const size_t o_step = round16(cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows); if (rc) return rc;
    rc = prl_k_bgr2gray(c, c->d_in, rows, cols, in_step, channels, c->d_out, o_step); if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, c->d_out, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

extern "C" int prl_cuda_pack_mask_dev(prl_cuda_ctx* c, const uint8_t* d_mask, int n_pages, int rows, int cols, size_t step,
                                      size_t page_stride, uint32_t* d_bits)
{
    if (!c) return PRL_E_INVALID;
    if (!d_mask || !d_bits || n_pages <= 0 || rows <= 0 || cols <= 0 || step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    return prl_k_pack_mask(c, d_mask, n_pages, rows, cols, step, page_stride, d_bits);
}

extern "C" int prl_cuda_morph(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step,
                              int morph_iters, uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    rc = prl_ensure(c, (void**)&c->d_tmp, &c->d_tmp_bytes, in_step * rows); if (rc) return rc;
    // a mask (only 0 and 255, what the reference's tail runs on) takes the bit-packed kernels; anything else the byte kernels
    rc = prl_ensure(c, &c->d_misc, &c->d_misc_bytes, 256); if (rc) return rc;
    rc = prl_k_not_binary(c, c->d_in, rows, cols, in_step, (int*)c->d_misc); if (rc) return rc;
    int not_binary = 1;
    PRL_CUDA_TRY(c, cudaMemcpyAsync(&not_binary, c->d_misc, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    rc = prl_k_morph(c, c->d_in, c->d_tmp, 1, rows, cols, in_step, in_step * rows, in_step, in_step * rows, morph_iters, !not_binary);
    if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, c->d_tmp, in_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

static int otsu_host(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, double maxval,
                     uint8_t* dst, size_t dst_step, int* thr, bool apply)
{
    if (!c) return PRL_E_INVALID;
    if (!src || rows <= 0 || cols <= 0 || step < (size_t)cols || (apply && (!dst || dst_step < (size_t)cols)))
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    const size_t o_step = round16(cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows); if (rc) return rc;
    rc = prl_ensure(c, &c->scalars, &c->scalars_bytes, 64); if (rc) return rc;
    int32_t* d_thr = (int32_t*)c->scalars;
    rc = prl_k_otsu_global(c, c->d_in, 1, rows, cols, in_step, in_step * rows, maxval, c->d_out, o_step, o_step * rows, d_thr, apply);
    if (rc) return rc;
    int32_t t = 0;
    PRL_CUDA_TRY(c, cudaMemcpyAsync(&t, d_thr, 4, cudaMemcpyDeviceToHost, c->stream));
    if (apply)
        PRL_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, c->d_out, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (thr) *thr = t;
    return PRL_OK;
}

// ---- edge front-end of prl::binarizeLocalOtsu ---------------------------------
// workspace: [blur u8][rows 8.8 u16][edge map A][edge map B][Otsu value][Canny scratch]
struct EdgesWs { uint8_t* blur; uint16_t* tmp16; uint8_t* ea; uint8_t* eb; int32_t* thr; void* canny; size_t p16; };

static int edges_ws(prl_cuda_ctx* c, int rows, int cols, EdgesWs* w)
{
    const size_t p16 = round16((size_t)cols), img = (p16 * rows + 255) & ~(size_t)255;
    const size_t need = img * 5 + 256 + prl_canny_scratch_bytes(rows, cols);
    int rc = prl_ensure(c, &c->edges_ws, &c->edges_ws_bytes, need); if (rc) return rc;
    uint8_t* b = (uint8_t*)c->edges_ws;
    w->p16 = p16; w->blur = b; w->tmp16 = (uint16_t*)(b + img); w->ea = b + 3 * img; w->eb = b + 4 * img;
    w->thr = (int32_t*)(b + 5 * img); w->canny = b + 5 * img + 256;
    return PRL_OK;
}

// CannyEdgeDetection (imageLibCommon.cpp:244-324) + the dilation of binarizeLocalOtsu.cpp:92, gray image in HBM
extern "C" int prl_cuda_canny_edge_detection_dev(prl_cuda_ctx* c, const uint8_t* d_gray, int rows, int cols, size_t step,
                                                 int gauss_ksize, double upper_coeff, double lower_coeff, int morph_iters,
                                                 int post_dilate, uint8_t* d_dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!d_gray || !d_dst || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    // the reference's own argument checks (imageLibCommon.cpp:248-272); cv::GaussianBlur needs an odd size
    if (gauss_ksize < 3) return prl_set_err(c, PRL_E_INVALID, "Gaussian blur kernel size is lesser than 3");     // :253-256
    if ((gauss_ksize & 1) == 0)                                       // cv::GaussianBlur asserts an odd size: cv::Exception
        return prl_set_err(c, PRL_E_EMPTY_ROI, "GaussianBlur: the kernel size must be odd");
    if (gauss_ksize > 63) return prl_set_err(c, PRL_E_UNSUPPORTED, "Gaussian blur kernel sizes above 63 are not supported");
    if (!(upper_coeff >= 0 && upper_coeff <= 1) || !(lower_coeff >= 0 && lower_coeff <= 1) || lower_coeff > upper_coeff)
        return prl_set_err(c, PRL_E_INVALID, "Canny threshold coefficients must satisfy 0 <= lower <= upper <= 1");
    if (post_dilate < 0 || post_dilate > 15 || morph_iters > 15 || morph_iters < -15)
        return prl_set_err(c, PRL_E_INVALID, "morphology iteration counts must be within [-15, 15] / [0, 15]");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    EdgesWs w;
    int rc = edges_ws(c, rows, cols, &w); if (rc) return rc;
    const size_t page = w.p16 * rows;
    rc = prl_k_gaussian_blur(c, d_gray, rows, cols, step, gauss_ksize, 0.0, w.blur, w.p16, w.tmp16); if (rc) return rc;
    rc = prl_k_otsu_global(c, w.blur, 1, rows, cols, w.p16, page, 255.0, nullptr, 0, 0, w.thr, false); if (rc) return rc;
    const bool last_is_canny = morph_iters == 0 && post_dilate == 0;
    rc = prl_k_canny(c, w.blur, rows, cols, w.p16, w.thr, upper_coeff, lower_coeff, 0, 0, last_is_canny ? d_dst : w.ea,
                     last_is_canny ? dst_step : w.p16, w.canny);
    if (rc) return rc;
    if (morph_iters != 0) {
        const bool last = post_dilate == 0;
        rc = prl_k_morph(c, w.ea, last ? d_dst : w.eb, 1, rows, cols, w.p16, page, last ? dst_step : w.p16, last ? dst_step * rows : page,
                         morph_iters, true);
        if (rc) return rc;
    }
    if (post_dilate > 0)
        rc = prl_k_morph_single(c, morph_iters != 0 ? w.eb : w.ea, d_dst, 1, rows, cols, w.p16, page, dst_step, dst_step * rows, post_dilate, true);
    return rc;
}

static int edges_host(prl_cuda_ctx* c, int what, const uint8_t* src, int rows, int cols, size_t step, int ksize, double a, double b,
                      int morph_iters, int post_dilate, uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    const size_t o_step = round16(cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows); if (rc) return rc;
    if (what == 0) {
        EdgesWs w; rc = edges_ws(c, rows, cols, &w); if (rc) return rc;
        rc = prl_k_gaussian_blur(c, c->d_in, rows, cols, in_step, ksize, a, c->d_out, o_step, w.tmp16);
    } else if (what == 1) {
        EdgesWs w; rc = edges_ws(c, rows, cols, &w); if (rc) return rc;
        rc = prl_k_canny(c, c->d_in, rows, cols, in_step, nullptr, 0, 0, a, b, c->d_out, o_step, w.canny);
    } else {
        rc = prl_cuda_canny_edge_detection_dev(c, c->d_in, rows, cols, in_step, ksize, a, b, morph_iters, post_dilate, c->d_out, o_step);
    }
    if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, c->d_out, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

// prl::binarizeLocalOtsu (binarizeLocalOtsu.cpp:38-163, CLAHE off) in one call: edge map, bounding rectangles of the
// top-level contours and the per-rectangle Otsu loop all on the device; the image crosses PCIe once each way.
// EnhanceLocalContrastByCLAHE on d_gray (rows x cols, pitch step) -> returns the enhanced image inside c->clahe_ws (pitch round16(cols))
static int clahe_dev(prl_cuda_ctx* c, const uint8_t* d_gray, int rows, int cols, size_t step, double clip_limit, bool equalize,
                     uint8_t** d_enh)
{
    const size_t p16 = round16((size_t)cols), img = (p16 * rows + 255) & ~(size_t)255;
    int rc = prl_ensure(c, &c->clahe_ws, &c->clahe_ws_bytes, 2 * img + 64 * 256 + 256 * 4 + 256 + 256); if (rc) return rc;
    uint8_t* b = (uint8_t*)c->clahe_ws;
    rc = prl_k_clahe(c, d_gray, rows, cols, step, clip_limit, equalize, b, p16, b + img, b + 2 * img); if (rc) return rc;
    *d_enh = b;
    return PRL_OK;
}

// EnhanceLocalContrastByCLAHE (imageLibCommon.cpp:326-346,378-396) for a single-channel image
extern "C" int prl_cuda_clahe(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, double clip_limit, int equalize,
                              uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    uint8_t* enh = nullptr;
    rc = clahe_dev(c, c->d_in, rows, cols, in_step, clip_limit, equalize != 0, &enh); if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, enh, round16((size_t)cols), cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

// ---- prl::binarizeLocalOtsu on a gray image resident in HBM, in two asynchronous halves around the one number the host
// needs (how many contour rectangles there are: it sizes the rectangle loop and decides the reference's error cases).
struct LocalOtsuJob {
    const uint8_t* proc = nullptr;      // imageToProc: the gray image, or its CLAHE + equalizeHist enhancement (binarizeLocalOtsu.cpp:79-82)
    size_t proc_step = 0;
    int* d_count = nullptr; int32_t* d_xywh = nullptr; int32_t* d_thr = nullptr;
};
constexpr int kRectCap = 65535;

// phase A: edge map (binarizeLocalOtsu.cpp:85-92) -> c->d_tmp, bounding rectangles of the top-level contours (:104-110,150);
// the count is copied to *h_count (pinned, or any host word for a synchronous caller) on the context's stream
static int local_otsu_phase_a(prl_cuda_ctx* c, const uint8_t* d_gray, int rows, int cols, size_t in_step, double clahe_clip_limit,
                              int gauss_ksize, double upper_coeff, double lower_coeff, int morph_iters, LocalOtsuJob* J, int* h_count)
{
    const size_t o_step = round16(cols);
    int rc = prl_ensure(c, (void**)&c->d_tmp, &c->d_tmp_bytes, o_step * rows); if (rc) return rc;
    J->proc = d_gray; J->proc_step = in_step;
    if (clahe_clip_limit > 0) {
        uint8_t* enh = nullptr;
        rc = clahe_dev(c, d_gray, rows, cols, in_step, clahe_clip_limit, true, &enh); if (rc) return rc;
        J->proc = enh; J->proc_step = round16((size_t)cols);
    }
    rc = prl_cuda_canny_edge_detection_dev(c, J->proc, rows, cols, J->proc_step, gauss_ksize, upper_coeff, lower_coeff, morph_iters, 3,
                                           c->d_tmp, o_step);
    if (rc) return rc;
    // the list never leaves the device on its way to the Otsu loop
    const size_t list_bytes = (256 + (size_t)kRectCap * 16 + (size_t)kRectCap * 4 + 255) & ~(size_t)255;
    rc = prl_ensure(c, &c->rects_ws, &c->rects_ws_bytes, list_bytes + prl_rects_scratch_bytes(rows, cols)); if (rc) return rc;
    J->d_count = (int*)c->rects_ws;
    J->d_xywh = (int32_t*)((uint8_t*)c->rects_ws + 256);
    J->d_thr = J->d_xywh + (size_t)kRectCap * 4;
    rc = prl_k_external_rects(c, c->d_tmp, rows, cols, o_step, J->d_count, J->d_xywh, kRectCap, (uint8_t*)c->rects_ws + list_bytes);
    if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpyAsync(h_count, J->d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    return PRL_OK;
}

// phase B: the rectangle loop (:138-162) into d_dst, or the reference's outcome for the degenerate counts.
// returns PRL_OK, PRL_E_INVALID ("Contours array is empty", imageLibCommon.cpp:643-646) or PRL_E_UNSUPPORTED (> 65535 contours)
static int local_otsu_phase_b(prl_cuda_ctx* c, const LocalOtsuJob& J, int rows, int cols, int count, double maxval, uint8_t* d_dst,
                              size_t dst_step, int* n_rects)
{
    if (n_rects) *n_rects = count;
    if (count == 0) return prl_set_err(c, PRL_E_INVALID, "Contours array is empty");
    if (count > kRectCap) return prl_set_err(c, PRL_E_UNSUPPORTED, "more than 65535 contours");
    if (rows == 1 || cols == 1) {
        // Every component of a one-pixel-wide image is a straight run: its CHAIN_APPROX_SIMPLE contour has 1 or 2
        // points and CheckHierarhyLevelRecursively ignores contours with fewer than 3 (imageLibCommon.cpp:729-736), so
        // no rectangle is thresholded and the result stays 255 (binarizeLocalOtsu.cpp:142).  With 2 or more rows and
        // columns a component of the 3x-dilated edge map always holds a 2 x 2 block, i.e. at least 4 contour points.
        if (n_rects) *n_rects = 0;
        PRL_CUDA_TRY(c, cudaMemset2DAsync(d_dst, dst_step, 255, cols, rows, c->stream));
        return PRL_OK;
    }
    return prl_k_otsu_rects(c, J.proc, rows, cols, J.proc_step, J.d_xywh, count, maxval, d_dst, dst_step, J.d_thr);
}

extern "C" int prl_cuda_binarize_local_otsu(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, int channels,
                                            double maxval, double clahe_clip_limit, int gauss_ksize, double upper_coeff, double lower_coeff, int morph_iters,
                                            uint8_t* dst, size_t dst_step, int* n_rects, int32_t* rects_out, int rects_cap)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3 && channels != 4) ||
        step < (size_t)cols * channels || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    if (!(maxval >= 0 && maxval <= 255)) return prl_set_err(c, PRL_E_INVALID, "Max value must be in range [0; 255]");   // :52-55
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc;
    if (channels == 1) {
        rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    } else {
        // cv::cvtColor(inputImage, ..., COLOR_RGB2GRAY) (binarizeLocalOtsu.cpp:63) on the device
        const size_t bstep = round16((size_t)cols * channels);
        in_step = round16(cols);
        rc = prl_ensure(c, (void**)&c->d_bgr, &c->d_bgr_bytes, bstep * rows); if (rc) return rc;
        rc = prl_ensure(c, (void**)&c->d_in, &c->d_in_bytes, in_step * rows); if (rc) return rc;
        PRL_CUDA_TRY(c, copy2d(c->d_bgr, bstep, src, step, (size_t)cols * channels, rows, cudaMemcpyHostToDevice, c->stream));
        rc = prl_k_bgr2gray(c, c->d_bgr, rows, cols, bstep, channels, c->d_in, in_step, true); if (rc) return rc;
    }
    const size_t o_step = round16(cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows); if (rc) return rc;
    LocalOtsuJob J;
    int count = 0;
    rc = local_otsu_phase_a(c, c->d_in, rows, cols, in_step, clahe_clip_limit, gauss_ksize, upper_coeff, lower_coeff, morph_iters, &J, &count);
    if (rc) return rc;
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    if (rects_out && rects_cap > 0 && count > 0 && count <= kRectCap && rows > 1 && cols > 1)
        PRL_CUDA_TRY(c, cudaMemcpyAsync(rects_out, J.d_xywh, (size_t)std::min(count, rects_cap) * 16, cudaMemcpyDeviceToHost, c->stream));
    rc = local_otsu_phase_b(c, J, rows, cols, count, maxval, c->d_out, o_step, n_rects);
    if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, c->d_out, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

// ---- batched F3 / F4 rows: pages resident in HBM, several pages side by side --------------------------------------------
// The kernel sequences of these two functions are ~25 / ~50 small launches per page, each far too small to fill the
// device.  Instead of re-writing every kernel over a page index, the batch runs kLanes pages CONCURRENTLY: each lane is a
// sub-context with its own stream and scratch, so the launches of different pages overlap on the SMs.
constexpr int kLanes = 16;

static int ensure_lanes(prl_cuda_ctx* c, int n)
{
    while ((int)c->lanes.size() < n) {
        prl_cuda_ctx* l = nullptr;
        int rc = prl_cuda_create(c->device, &l);
        if (rc) return prl_set_err(c, rc, prl_cuda_last_error(nullptr));
        c->lanes.push_back(l);
    }
    if (!c->h_lane_counts) PRL_CUDA_TRY(c, cudaMallocHost((void**)&c->h_lane_counts, sizeof(int) * kLanes));
    return PRL_OK;
}

static int sync_lanes(prl_cuda_ctx* c, int n)
{
    for (int i = 0; i < n; ++i) PRL_CUDA_TRY(c, cudaStreamSynchronize(c->lanes[i]->stream));
    return PRL_OK;
}

// prl::binarizeLocalOtsu (binarizeLocalOtsu.cpp:38-163) over n_pages gray pages in HBM.  Synchronous.  status[p] (host,
// optional) receives the page's own outcome: PRL_OK, PRL_E_INVALID = "Contours array is empty" (the std::invalid_argument of
// the reference for that image; its output page is left untouched), PRL_E_UNSUPPORTED = more than 65535 contours.
extern "C" int prl_cuda_binarize_local_otsu_batch_dev(prl_cuda_ctx* c, const uint8_t* d_gray, int n_pages, int rows, int cols, size_t step,
                                                      size_t page_stride, double maxval, double clahe_clip_limit, int gauss_ksize,
                                                      double upper_coeff, double lower_coeff, int morph_iters, uint8_t* d_dst, size_t dst_step,
                                                      size_t dst_page_stride, int32_t* n_rects, int32_t* status)
{
    if (!c) return PRL_E_INVALID;
    if (!d_gray || !d_dst || n_pages <= 0 || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    if (!(maxval >= 0 && maxval <= 255)) return prl_set_err(c, PRL_E_INVALID, "Max value must be in range [0; 255]");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    const int nl = std::min(kLanes, n_pages);
    int rc = ensure_lanes(c, nl); if (rc) return rc;
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));               // the pages may have been produced on the caller's stream
    LocalOtsuJob jobs[kLanes];
    for (int p0 = 0; p0 < n_pages; p0 += nl) {
        const int np = std::min(nl, n_pages - p0);
        for (int i = 0; i < np; ++i) {
            prl_cuda_ctx* l = c->lanes[i];
            rc = local_otsu_phase_a(l, d_gray + (size_t)(p0 + i) * page_stride, rows, cols, step, clahe_clip_limit, gauss_ksize, upper_coeff,
                                    lower_coeff, morph_iters, &jobs[i], c->h_lane_counts + i);
            if (rc) { c->err = l->err; sync_lanes(c, nl); return rc; }
        }
        rc = sync_lanes(c, np); if (rc) return rc;
        for (int i = 0; i < np; ++i) {
            prl_cuda_ctx* l = c->lanes[i];
            int nr = 0;
            const int prc = local_otsu_phase_b(l, jobs[i], rows, cols, c->h_lane_counts[i], maxval, d_dst + (size_t)(p0 + i) * dst_page_stride,
                                               dst_step, &nr);
            if (prc == PRL_E_CUDA || prc == PRL_E_NOMEM) { c->err = l->err; sync_lanes(c, nl); return prc; }
            if (n_rects) n_rects[p0 + i] = nr;
            if (status) status[p0 + i] = prc;
        }
    }
    return sync_lanes(c, nl);
}

// prl::removeLines (removeLines.cpp:30-77) over n_pages gray pages in HBM.  Synchronous.
extern "C" int prl_cuda_remove_lines_batch_dev(prl_cuda_ctx* c, const uint8_t* d_gray, int n_pages, int rows, int cols, size_t step,
                                               size_t page_stride, uint8_t* d_dst, size_t dst_step, size_t dst_page_stride)
{
    if (!c) return PRL_E_INVALID;
    if (!d_gray || !d_dst || n_pages <= 0 || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    const int nl = std::min(kLanes, n_pages);
    int rc = ensure_lanes(c, nl); if (rc) return rc;
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    for (int p = 0; p < n_pages; ++p) {
        prl_cuda_ctx* l = c->lanes[p % nl];
        rc = prl_ensure(l, &l->edges_ws, &l->edges_ws_bytes, prl_lines_scratch_bytes(rows, cols));
        if (!rc) rc = prl_k_remove_lines(l, d_gray + (size_t)p * page_stride, rows, cols, step, d_dst + (size_t)p * dst_page_stride, dst_step, l->edges_ws);
        if (rc) { c->err = l->err; sync_lanes(c, nl); return rc; }
    }
    return sync_lanes(c, nl);
}

// bounding rectangles of the top-level contours of a binary image (non-zero = foreground): the set
// cv::findContours(RETR_EXTERNAL) + cv::boundingRect returns (binarizeLocalOtsu.cpp:104-105,150)
extern "C" int prl_cuda_external_rects(prl_cuda_ctx* c, const uint8_t* mask, int rows, int cols, size_t step, int32_t* rects_out,
                                       int rects_cap, int* n_rects)
{
    if (!c) return PRL_E_INVALID;
    if (!mask || rows <= 0 || cols <= 0 || step < (size_t)cols || rects_cap < 0 || (rects_cap > 0 && !rects_out) || !n_rects)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, mask, rows, cols, step, &in_step); if (rc) return rc;
    const int cap = std::max(rects_cap, 1);
    const size_t list_bytes = (256 + (size_t)cap * 16 + 255) & ~(size_t)255;
    rc = prl_ensure(c, &c->rects_ws, &c->rects_ws_bytes, list_bytes + prl_rects_scratch_bytes(rows, cols)); if (rc) return rc;
    int* d_count = (int*)c->rects_ws;
    int32_t* d_xywh = (int32_t*)((uint8_t*)c->rects_ws + 256);
    rc = prl_k_external_rects(c, c->d_in, rows, cols, in_step, d_count, d_xywh, rects_cap, (uint8_t*)c->rects_ws + list_bytes);
    if (rc) return rc;
    int count = 0;
    PRL_CUDA_TRY(c, cudaMemcpyAsync(&count, d_count, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    *n_rects = count;
    if (rects_cap > 0 && count > 0)
        PRL_CUDA_TRY(c, cudaMemcpy(rects_out, d_xywh, (size_t)std::min(count, rects_cap) * 16, cudaMemcpyDeviceToHost));
    return PRL_OK;
}

// prl::removeLines (src/removeLines.cpp:30-77): 1 or 3 (BGR) channels in, 0/255 image out
extern "C" int prl_cuda_remove_lines(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, int channels,
                                     uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3) || step < (size_t)cols * channels ||
        dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc;
    if (channels == 1) {
        rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    } else {
        const size_t bstep = round16((size_t)cols * channels);
        in_step = round16(cols);
        rc = prl_ensure(c, (void**)&c->d_bgr, &c->d_bgr_bytes, bstep * rows); if (rc) return rc;
        rc = prl_ensure(c, (void**)&c->d_in, &c->d_in_bytes, in_step * rows); if (rc) return rc;
        PRL_CUDA_TRY(c, copy2d(c->d_bgr, bstep, src, step, (size_t)cols * channels, rows, cudaMemcpyHostToDevice, c->stream));
        rc = prl_k_bgr2gray(c, c->d_bgr, rows, cols, bstep, channels, c->d_in, in_step, false); if (rc) return rc;   // COLOR_BGR2GRAY :35
    }
    const size_t o_step = round16(cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows); if (rc) return rc;
    rc = prl_ensure(c, &c->edges_ws, &c->edges_ws_bytes, prl_lines_scratch_bytes(rows, cols)); if (rc) return rc;
    rc = prl_k_remove_lines(c, c->d_in, rows, cols, in_step, c->d_out, o_step, c->edges_ws); if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, c->d_out, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

// ---- the adaptive-mean family ---------------------------------------------------
extern "C" int prl_cuda_gauss_kernel_float(int n, float* k)
{
    if (!k) return PRL_E_INVALID;
    return prl_gauss_kernel_float(n, k);
}

extern "C" int prl_cuda_median_blur(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, int channels, int ksize,
                                    uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3 && channels != 4) ||
        step < (size_t)cols * channels || dst_step < (size_t)cols * channels)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    if ((ksize & 1) == 0) return prl_set_err(c, PRL_E_EMPTY_ROI, "medianBlur: the kernel size must be odd");     // cv::Exception
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, (size_t)cols * channels, step, &in_step); if (rc) return rc;
    const size_t o_step = round16((size_t)cols * channels);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows); if (rc) return rc;
    if (ksize <= 1) {                                                                                             // src.copyTo(dst)
        PRL_CUDA_TRY(c, cudaMemcpy2DAsync(c->d_out, o_step, c->d_in, in_step, (size_t)cols * channels, rows, cudaMemcpyDeviceToDevice, c->stream));
    } else {
        rc = prl_k_median_blur(c, c->d_in, rows, cols, in_step, channels, ksize, c->d_out, o_step); if (rc) return rc;
    }
    PRL_CUDA_TRY(c, copy2d(dst, dst_step, c->d_out, o_step, (size_t)cols * channels, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

extern "C" int prl_cuda_adaptive_threshold(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, double maxval,
                                           int method, int type, int block_size, double delta, uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    // cv::adaptiveThreshold's own assertions -> cv::Exception
    if (block_size <= 1 || (block_size & 1) == 0) return prl_set_err(c, PRL_E_EMPTY_ROI, "adaptiveThreshold: blockSize % 2 == 1 && blockSize > 1");
    if ((method != 0 && method != 1) || (type != 0 && type != 1)) return prl_set_err(c, PRL_E_EMPTY_ROI, "adaptiveThreshold: unknown method or threshold type");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    const size_t o_step = (dst_step == (size_t)cols) ? (size_t)cols : round16(cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows + 16); if (rc) return rc;
    rc = prl_ensure(c, &c->adaptive_ws, &c->adaptive_ws_bytes, prl_adaptive_scratch_bytes(rows, cols)); if (rc) return rc;
    rc = prl_k_adaptive_threshold(c, c->d_in, rows, cols, in_step, maxval, method, type, block_size, delta, c->d_out, o_step,
                                  c->adaptive_ws, false);
    if (rc) return rc;
    PRL_CUDA_TRY(c, copy2d(dst, dst_step, c->d_out, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

// prl_adaptive_params checked the way the reference's functions and cv::adaptiveThreshold / medianBlur / GaussianBlur check their
// arguments; *bs receives the block size (the automatic one when asked for)
static int adaptive_check(prl_cuda_ctx* c, int rows, int cols, int channels, const prl_adaptive_params* p, int* bs_out)
{
    if (!(p->maxval >= 0 && p->maxval <= 255) && p->check_maxval)
        return prl_set_err(c, PRL_E_INVALID, "Max value must be in range [0; 255]");                // binarizeNativeAdaptive.cpp:53-56
    int bs = p->block_size;
    if (p->auto_block && bs < 3) {                                                                  // :86-93
        const double diagonal = std::sqrt((double)(rows * rows + cols * cols));
        bs = (int)(diagonal / 333 + 7);
    }
    if (bs <= 1 || (bs & 1) == 0) return prl_set_err(c, PRL_E_EMPTY_ROI, "adaptiveThreshold: blockSize % 2 == 1 && blockSize > 1");
    if (p->blur == 1 && p->blur_ksize < 3 && p->assert_ksize) return prl_set_err(c, PRL_E_EMPTY_ROI, "medianBlurKernelSize >= 3");   // CV_Assert :65
    if (p->blur == 1 && (p->blur_ksize & 1) == 0) return prl_set_err(c, PRL_E_EMPTY_ROI, "medianBlur: the kernel size must be odd");
    if (p->blur == 2) {
        if (p->blur_ksize < 3 || !(p->blur_sigma > 0)) return prl_set_err(c, PRL_E_EMPTY_ROI, "GaussianBlurKernelSize >= 3 && GaussianBlurSigma > 0");   // :70-71
        if ((p->blur_ksize & 1) == 0) return prl_set_err(c, PRL_E_EMPTY_ROI, "GaussianBlur: the kernel size must be odd");
        if (p->blur_ksize > 63) return prl_set_err(c, PRL_E_UNSUPPORTED, "Gaussian blur kernel sizes above 63 are not supported");
        if (channels != 1 && !p->gray_first) return prl_set_err(c, PRL_E_UNSUPPORTED, "Gaussian blur of a colour image is not supported");
    }
    if ((p->method != 0 && p->method != 1) || (p->type != 0 && p->type != 1)) return prl_set_err(c, PRL_E_INVALID, "unknown method or threshold type");
    *bs_out = bs;
    return PRL_OK;
}

// The kernel sequence of the adaptive family for one image resident in HBM, on c's stream with c's scratch: colour conversion,
// blur, threshold, the mean test and the optional bilateral filter.  d_src: `channels` interleaved bytes per pixel, pitch
// src_step; d_dst: one byte per pixel.  Asynchronous.
static int adaptive_dev(prl_cuda_ctx* c, const uint8_t* d_src, int rows, int cols, size_t src_step, int channels, const prl_adaptive_params* p,
                        int bs, uint8_t* d_dst, size_t dst_step)
{
    const size_t g_step = round16((size_t)cols), g_img = g_step * rows;
    const size_t c_step = round16((size_t)cols * channels);
    const bool bilateral = p->bilateral_d >= 3;
    int rc = prl_ensure(c, (void**)&c->d_tmp, &c->d_tmp_bytes, g_img); if (rc) return rc;
    rc = prl_ensure(c, &c->adaptive_ws, &c->adaptive_ws_bytes,
                    prl_adaptive_scratch_bytes(rows, cols) + (bilateral ? prl_bilateral_scratch_bytes(p->bilateral_d, p->bilateral_sigma_space) : 0));
    if (rc) return rc;
    const uint8_t* gray = d_src;            // what adaptiveThreshold reads
    size_t gray_step = src_step;
    if (channels != 1) {
        rc = prl_ensure(c, (void**)&c->d_in, &c->d_in_bytes, g_img); if (rc) return rc;
        const uint8_t* colour = d_src;
        size_t colour_step = src_step;
        if (!p->gray_first && p->blur == 1 && p->blur_ksize > 1) {                  // cv::medianBlur on the colour image (binarizeAT.cpp:53)
            rc = prl_ensure(c, &c->d_misc, &c->d_misc_bytes, c_step * rows); if (rc) return rc;
            rc = prl_k_median_blur(c, d_src, rows, cols, src_step, channels, p->blur_ksize, (uint8_t*)c->d_misc, c_step); if (rc) return rc;
            colour = (const uint8_t*)c->d_misc; colour_step = c_step;
        }
        rc = prl_k_bgr2gray(c, colour, rows, cols, colour_step, channels, c->d_in, g_step, false); if (rc) return rc;   // COLOR_BGR2GRAY
        gray = c->d_in; gray_step = g_step;
    }
    if (p->blur != 0 && (channels == 1 || p->gray_first)) {
        if (p->blur == 1 && p->blur_ksize > 1) {
            rc = prl_k_median_blur(c, gray, rows, cols, gray_step, 1, p->blur_ksize, c->d_tmp, g_step); if (rc) return rc;
            gray = c->d_tmp; gray_step = g_step;
        } else if (p->blur == 2) {
            rc = prl_ensure(c, &c->edges_ws, &c->edges_ws_bytes, g_img * 2 + 256); if (rc) return rc;
            rc = prl_k_gaussian_blur(c, gray, rows, cols, gray_step, p->blur_ksize, p->blur_sigma, c->d_tmp, g_step, (uint16_t*)c->edges_ws);
            if (rc) return rc;
            gray = c->d_tmp; gray_step = g_step;
        }
    }
    uint8_t* thr = d_dst;                   // the threshold writes the result itself unless the bilateral filter follows
    size_t thr_step = dst_step;
    if (bilateral) {
        rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, g_img + 16); if (rc) return rc;
        thr = c->d_out; thr_step = g_step;
    }
    rc = prl_k_adaptive_threshold(c, gray, rows, cols, gray_step, p->maxval, p->method, p->type, bs, p->delta, thr, thr_step,
                                  c->adaptive_ws, p->invert_if_dark != 0);
    if (rc) return rc;
    if (bilateral) {                                                                                // binarizeNativeAdaptive.cpp:116-134
        // the reference reaches these checks after the threshold, so its cv::Exceptions above come first
        if (p->bilateral_sigma_color <= 0) return prl_set_err(c, PRL_E_INVALID, "Color sigma for bilateral filtration must be greater than 0");
        if (p->bilateral_sigma_space <= 0) return prl_set_err(c, PRL_E_INVALID, "Space sigma for bilateral filtration must be greater than 0");
        rc = prl_k_bilateral(c, thr, rows, cols, thr_step, p->bilateral_d, p->bilateral_sigma_color, p->bilateral_sigma_space,
                             d_dst, dst_step, (char*)c->adaptive_ws + prl_adaptive_scratch_bytes(rows, cols));
        if (rc) return rc;
    }
    return PRL_OK;
}

// One call for prl::binarizeNativeAdaptive / binarizeAT / binarizeAGT / binarizePureAdaptiveGaussian: the image crosses
// PCIe once each way; colour conversion, blur, threshold, the mean test and the bilateral filter all run on the device.
extern "C" int prl_cuda_binarize_adaptive(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, int channels,
                                          const prl_adaptive_params* p, uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || !p || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3 && channels != 4) ||
        step < (size_t)cols * channels || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    int bs = 0;
    int rc = adaptive_check(c, rows, cols, channels, p, &bs); if (rc) return rc;
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    const size_t c_step = round16((size_t)cols * channels);
    rc = prl_ensure(c, (void**)&c->d_bgr, &c->d_bgr_bytes, c_step * rows); if (rc) return rc;
    PRL_CUDA_TRY(c, copy2d(c->d_bgr, c_step, src, step, (size_t)cols * channels, rows, cudaMemcpyHostToDevice, c->stream));
    const size_t o_step = (dst_step == (size_t)cols) ? (size_t)cols : round16((size_t)cols);
    rc = prl_ensure(c, (void**)&c->d_res, &c->d_res_bytes, o_step * rows + 16); if (rc) return rc;  // d_out may hold the mask the bilateral filter reads
    uint8_t* d_res = c->d_res;
    rc = adaptive_dev(c, c->d_bgr, rows, cols, c_step, channels, p, bs, d_res, o_step); if (rc) return rc;
    PRL_CUDA_TRY(c, copy2d(dst, dst_step, d_res, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

// The same over n_pages images of one size resident in HBM (page p at d_src + p * src_page_stride): the per-page kernel sequences
// run on the context's page lanes (own streams and scratch), up to 16 pages at a time.  Synchronous.
extern "C" int prl_cuda_binarize_adaptive_batch_dev(prl_cuda_ctx* c, const uint8_t* d_src, int n_pages, int rows, int cols, size_t src_step,
                                                    size_t src_page_stride, int channels, const prl_adaptive_params* p, uint8_t* d_dst,
                                                    size_t dst_step, size_t dst_page_stride)
{
    if (!c) return PRL_E_INVALID;
    if (!d_src || !d_dst || !p || n_pages <= 0 || rows <= 0 || cols <= 0 || (channels != 1 && channels != 3 && channels != 4) ||
        src_step < (size_t)cols * channels || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    int bs = 0;
    int rc = adaptive_check(c, rows, cols, channels, p, &bs); if (rc) return rc;
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    const int nl = std::min(kLanes, n_pages);
    rc = ensure_lanes(c, nl); if (rc) return rc;
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));               // the pages may have been produced on the caller's stream
    for (int pg = 0; pg < n_pages; ++pg) {
        prl_cuda_ctx* l = c->lanes[pg % nl];
        rc = adaptive_dev(l, d_src + (size_t)pg * src_page_stride, rows, cols, src_step, channels, p, bs, d_dst + (size_t)pg * dst_page_stride, dst_step);
        if (rc) { c->err = l->err; sync_lanes(c, nl); return rc; }
    }
    return sync_lanes(c, nl);
}

extern "C" int prl_cuda_bilateral_filter(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, int d, double sigma_color,
                                         double sigma_space, uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    const size_t o_step = round16(cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows); if (rc) return rc;
    rc = prl_ensure(c, &c->adaptive_ws, &c->adaptive_ws_bytes, prl_bilateral_scratch_bytes(d, sigma_space)); if (rc) return rc;
    rc = prl_k_bilateral(c, c->d_in, rows, cols, in_step, d, sigma_color, sigma_space, c->d_out, o_step, c->adaptive_ws); if (rc) return rc;
    PRL_CUDA_TRY(c, copy2d(dst, dst_step, c->d_out, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

extern "C" int prl_cuda_gauss_kernel_fixed(int n, double sigma, int* k)
{
    if (!k) return PRL_E_INVALID;
    return prl_gauss_kernel_fixed(n, sigma, k);
}

extern "C" int prl_cuda_gaussian_blur(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, int ksize,
                                      double sigma, uint8_t* dst, size_t dst_step)
{
    return edges_host(c, 0, src, rows, cols, step, ksize, sigma, 0, 0, 0, dst, dst_step);
}

extern "C" int prl_cuda_canny(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, double low, double high,
                              uint8_t* dst, size_t dst_step)
{
    return edges_host(c, 1, src, rows, cols, step, 0, low, high, 0, 0, dst, dst_step);
}

extern "C" int prl_cuda_canny_edge_detection(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step,
                                             int gauss_ksize, double upper_coeff, double lower_coeff, int morph_iters,
                                             int post_dilate, uint8_t* dst, size_t dst_step)
{
    return edges_host(c, 2, src, rows, cols, step, gauss_ksize, upper_coeff, lower_coeff, morph_iters, post_dilate, dst, dst_step);
}

extern "C" int prl_cuda_otsu_threshold(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step, int* thr)
{
    return otsu_host(c, src, rows, cols, step, 255.0, nullptr, 0, thr, false);
}

extern "C" int prl_cuda_otsu_global(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step,
                                    double maxval, uint8_t* dst, size_t dst_step, int* thr)
{
    return otsu_host(c, src, rows, cols, step, maxval, dst, dst_step, thr, true);
}

extern "C" int prl_cuda_otsu_rects(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step,
                                   const int32_t* xywh, int n_rects, double maxval, uint8_t* dst, size_t dst_step,
                                   int32_t* thr_out)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols || n_rects < 0 ||
        (n_rects > 0 && !xywh))
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    for (int r = 0; r < n_rects; ++r) {   // cv::Mat::operator()(Rect) asserts the ROI is inside the image
        const int32_t* q = xywh + 4 * r;
        if (q[2] <= 0 || q[3] <= 0 || q[0] < 0 || q[1] < 0 || (long long)q[0] + q[2] > cols || (long long)q[1] + q[3] > rows)
            return prl_set_err(c, PRL_E_INVALID, "rectangle outside the image");
    }
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    const size_t o_step = round16(cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows); if (rc) return rc;
    const size_t rl = (size_t)std::max(1, n_rects) * 16;
    rc = prl_ensure(c, (void**)&c->d_tmp, &c->d_tmp_bytes, rl + (size_t)std::max(1, n_rects) * 4); if (rc) return rc;
    int32_t* d_rects = (int32_t*)c->d_tmp;
    int32_t* d_thr = (int32_t*)(c->d_tmp + rl);
    if (n_rects) PRL_CUDA_TRY(c, cudaMemcpyAsync(d_rects, xywh, (size_t)n_rects * 16, cudaMemcpyHostToDevice, c->stream));
    rc = prl_k_otsu_rects(c, c->d_in, rows, cols, in_step, d_rects, n_rects, maxval, c->d_out, o_step, d_thr);
    if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, c->d_out, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    if (thr_out && n_rects)
        PRL_CUDA_TRY(c, cudaMemcpyAsync(thr_out, d_thr, (size_t)n_rects * 4, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}

extern "C" int prl_cuda_otsu_tiles(prl_cuda_ctx* c, const uint8_t* src, int rows, int cols, size_t step,
                                   int tile_w, int tile_h, double maxval, uint8_t* dst, size_t dst_step)
{
    if (!c) return PRL_E_INVALID;
    if (!src || !dst || rows <= 0 || cols <= 0 || step < (size_t)cols || dst_step < (size_t)cols || tile_w <= 0 || tile_h <= 0)
        return prl_set_err(c, PRL_E_INVALID, "bad argument");
    PRL_CUDA_TRY(c, cudaSetDevice(c->device));
    size_t in_step;
    int rc = stage_in(c, src, rows, cols, step, &in_step); if (rc) return rc;
    const size_t o_step = round16(cols);
    rc = prl_ensure(c, (void**)&c->d_out, &c->d_out_bytes, o_step * rows); if (rc) return rc;
    rc = prl_cuda_otsu_tiles_batch_dev(c, c->d_in, 1, rows, cols, in_step, in_step * rows, tile_w, tile_h, maxval,
                                       c->d_out, o_step, o_step * rows);
    if (rc) return rc;
    PRL_CUDA_TRY(c, cudaMemcpy2DAsync(dst, dst_step, c->d_out, o_step, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    PRL_CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return PRL_OK;
}
