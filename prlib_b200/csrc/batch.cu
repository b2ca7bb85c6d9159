// batch.cu -- the ctx-less host batch entry points: pinned-memory loader + page dispatcher (one host thread per device, no
// collective), the 1-bit return path of the masks with its host-side expansion, page-locked memory helpers and the
// process-wide options.  Replaces nothing in the reference (its API is one cv::Mat per call); it is what BASELINE
// configs[4] and the north star's "pinned-memory batch loader" ask for.
#include "common.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <map>
#include <thread>
#include <mutex>
#include <memory>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <functional>
#if defined(__linux__)
#include <sched.h>
#endif
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#endif

// process-wide knobs of the ctx-less batch entry points (prl_cuda_set_global_option)
static std::atomic<long long> g_batch_chunk_pages{0};     // pages per ring slot, 0 = automatic (~72 MiB of input)
static std::atomic<long long> g_batch_unpack_threads{-1}; // > 0: byte masks cross PCIe as 1 bit per pixel and this many host threads per
                                                          //    device expand them into the caller's buffer; 0: the bytes themselves cross;
                                                          //    -1 (default): decided from the host cores per GPU, see unpack_threads_auto
static std::atomic<long long> g_batch_unpack_lag{0};      // 0: the expansion jobs wait for their chunk's D2H themselves; 1: the submitting thread does
static std::atomic<long long> g_batch_unpack_nt{1};       // AVX2 expansion: non-temporal stores (1) or ordinary ones (0)
static std::atomic<long long> g_batch_pageable{1};        // 1: pageable host buffers are staged through library-owned pinned
                                                          //    bounce buffers (correct overlap, host-memcpy bound); 0: handed to the
                                                          //    driver as they are (synchronous staged copies, no overlap)

// ---- masks over PCIe at 1 bit per pixel.  The end-to-end rate of prl_cuda_binarize_batch is the PCIe link's: 8.7 MB in and 8.7 MB
// out per A4 page, both directions busy ([B200 box] 49.6 GB/s each way together, 55.6 GB/s for H2D alone).  The result is a 0/255
// mask, so it crosses as PIX words (prl_k_pack_mask: bit 31 - (x & 31) of word x >> 5, 1 = black) into library-owned pinned
// buffers and a few host threads per device expand it into the caller's buffer while later chunks are in flight: the link then
// carries 1.125 bytes per pixel instead of 2, and the caller's mask buffer no longer has to be page-locked.
static void unpack_rows_scalar(const uint32_t* bits, size_t wpl, uint8_t* dst, size_t dpitch, int rows, int cols)
{
    static uint64_t lut[256];
    static std::once_flag once;
    std::call_once(once, [] {
        for (int b = 0; b < 256; ++b) {
            uint64_t v = 0;
            for (int i = 0; i < 8; ++i) if (!((b >> (7 - i)) & 1)) v |= 0xffull << (8 * i);    // first pixel = top bit; 0 = white = 255
            lut[b] = v;
        }
    });
    for (int y = 0; y < rows; ++y) {
        const uint32_t* w = bits + (size_t)y * wpl;
        uint8_t* o = dst + (size_t)y * dpitch;
        int x = 0;
        for (; x + 32 <= cols; x += 32) {
            const uint32_t v = w[x >> 5];
            for (int k = 0; k < 4; ++k) memcpy(o + x + 8 * k, &lut[(v >> (24 - 8 * k)) & 0xffu], 8);
        }
        for (; x < cols; ++x) o[x] = ((w[x >> 5] >> (31 - (x & 31))) & 1u) ? 0 : 255;
    }
}

#if defined(__x86_64__) && defined(__GNUC__)
__attribute__((target("avx2"))) static void unpack_rows_avx2(const uint32_t* bits, size_t wpl, uint8_t* dst, size_t dpitch, int rows, int cols)
{
    // byte j of the vector takes source byte 3 - j / 8 of the word (PIX words are most-significant-bit first), bit 7 - j % 8.
    // The destination is written once and not read back here: aligned non-temporal stores (no read-for-ownership traffic; an
    // ordinary store loop tops out near 7 GB/s per thread); rows start at any alignment, so the bit stream is re-cut at the
    // first 32-byte boundary of each row.
    const __m256i pick = _mm256_setr_epi8(3, 3, 3, 3, 3, 3, 3, 3, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i bit = _mm256_set1_epi64x((long long)0x0102040810204080ull);
    const bool nt = g_batch_unpack_nt.load() != 0;
    for (int y = 0; y < rows; ++y) {
        const uint32_t* w = bits + (size_t)y * wpl;
        uint8_t* o = dst + (size_t)y * dpitch;
        const int head = std::min(cols, (int)((32 - ((uintptr_t)o & 31)) & 31));
        int x = 0;
        for (; x < head; ++x) o[x] = ((w[x >> 5] >> (31 - (x & 31))) & 1u) ? 0 : 255;
        const int sh = head & 31;
        for (int i = 0; x + 32 <= cols; x += 32, ++i) {
            const uint32_t v = sh ? (w[i] << sh) | (w[i + 1] >> (32 - sh)) : w[i];            // pixels x .. x + 31, first pixel in the top bit
            const __m256i e = _mm256_shuffle_epi8(_mm256_set1_epi32((int)v), pick);
            const __m256i white = _mm256_cmpeq_epi8(_mm256_and_si256(e, bit), _mm256_setzero_si256());
            if (nt) _mm256_stream_si256(reinterpret_cast<__m256i*>(o + x), white); else _mm256_store_si256(reinterpret_cast<__m256i*>(o + x), white);
        }
        for (; x < cols; ++x) o[x] = ((w[x >> 5] >> (31 - (x & 31))) & 1u) ? 0 : 255;
    }
    _mm_sfence();
}
static void unpack_rows(const uint32_t* bits, size_t wpl, uint8_t* dst, size_t dpitch, int rows, int cols)
{
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2) unpack_rows_avx2(bits, wpl, dst, dpitch, rows, cols); else unpack_rows_scalar(bits, wpl, dst, dpitch, rows, cols);
}
#else
static void unpack_rows(const uint32_t* bits, size_t wpl, uint8_t* dst, size_t dpitch, int rows, int cols) { unpack_rows_scalar(bits, wpl, dst, dpitch, rows, cols); }
#endif

// How many host threads per device expand masks when the caller did not say.  The expansion writes 8.7 MB per A4 page, about
// 0.75 k pages/s per thread on the measured hosts, and it competes with what the same call achieves sending bytes:
//   [1 B200, 16 cores]   bytes 4.8 k pages/s; bits with 4 threads 3.2 k, 8 threads 5.9 k, 12 threads 5.8 k
//   [2 B200]             bytes 9.0 k; bits with 8 threads per GPU 9.3 k
//   [8 B200, 32 cores]   bytes 7.6 k (D2H into host memory is that box's weak direction: 91 GB/s alone, 63 GB/s beside H2D, against
//                        187 GB/s for H2D alone); bits with 1 thread per GPU 5.9 k, 2: 8.6 k, 3: 9.3 k, 4: 9.4 k, 6: 9.4 k
// So: on boxes of four or more GPUs, where the link is shared and bytes are the expensive direction, always bits with
// cores per GPU - 1 threads (2 to 8); on one or two GPUs only where 6 or more threads can be spared, else bytes.
static int unpack_threads_auto()
{
    static const int n = [] {
        int cores = (int)std::thread::hardware_concurrency();
#if defined(__linux__)
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
#endif
        const int gpus = std::max(1, prl_cuda_device_count());
        const int per_gpu = cores / gpus;
        if (gpus >= 4) return std::min(8, std::max(2, per_gpu - 1));
        const int t = std::min(8, per_gpu - 2);                   // leave the submitting thread and the caller some room
        return t >= 6 ? t : 0;
    }();
    return n;
}

extern "C" int prl_cuda_batch_unpack_threads(void)
{
    const long long v = g_batch_unpack_threads.load();
    return v < 0 ? unpack_threads_auto() : (int)v;
}

// test hook: the host-side expansion alone
extern "C" int prl_cuda_unpack_mask_host(const uint32_t* bits, int rows, int cols, uint8_t* mask, int force_scalar)
{
    if (!bits || !mask || rows <= 0 || cols <= 0) return PRL_E_INVALID;
    const size_t wpl = ((size_t)cols + 31) / 32;
    if (force_scalar) unpack_rows_scalar(bits, wpl, mask, (size_t)cols, rows, cols); else unpack_rows(bits, wpl, mask, (size_t)cols, rows, cols);
    return PRL_OK;
}

namespace {
// a few host threads that expand packed masks (and stage pageable images); a job is a band of rows
struct UnpackPool {
    struct Job { std::function<void()> fn; std::atomic<int>* left; };
    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    std::deque<Job> jobs;
    bool stop = false;
    void start(int n)
    {
        std::lock_guard<std::mutex> lk(mu);
        while ((int)threads.size() < n)
            threads.emplace_back([this] {
                for (;;) {
                    Job j;
                    {
                        std::unique_lock<std::mutex> lk2(mu);
                        cv.wait(lk2, [this] { return stop || !jobs.empty(); });
                        if (jobs.empty()) return;
                        j = std::move(jobs.front()); jobs.pop_front();
                    }
                    j.fn();
                    if (j.left->fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk2(mu); cv_done.notify_all(); }
                }
            });
    }
    void submit(std::function<void()> fn, std::atomic<int>* left) { { std::lock_guard<std::mutex> lk(mu); jobs.push_back(Job{std::move(fn), left}); } cv.notify_one(); }
    void wait(std::atomic<int>& left) { std::unique_lock<std::mutex> lk(mu); cv_done.wait(lk, [&] { return left.load() == 0; }); }
    ~UnpackPool()
    {
        { std::lock_guard<std::mutex> lk(mu); stop = true; }
        cv.notify_all();
        for (auto& t : threads) t.join();
    }
};
}  // namespace

// ------------------------------------------------------------------------------------------------
// host batch: pinned-memory loader + page dispatcher (one host thread per device, no collective)
// ------------------------------------------------------------------------------------------------

namespace {

struct DeviceWorker {
    prl_cuda_ctx* ctx = nullptr;
    static constexpr int HB = 6;                              // host slots of the 1-bit return path (twice the device ring: the
    uint32_t* h_bits[HB] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // expansion of a chunk trails its D2H)
    size_t h_bits_bytes = 0;
    cudaEvent_t ev_hb[HB];
    int n_hb_events = 0;
    std::atomic<int> hb_left[HB];
    UnpackPool pool;
    std::mutex busy;                                          // held for a whole shard: concurrent callers on one device take turns
    cudaStream_t s_in = nullptr, s_out = nullptr;
    static constexpr int NBUF = 3;
    uint8_t* d_in[NBUF] = {nullptr, nullptr, nullptr};
    uint8_t* d_out[NBUF] = {nullptr, nullptr, nullptr};
    uint32_t* d_bits[NBUF] = {nullptr, nullptr, nullptr};     // packed variant: 1 bit per pixel leaves the device
    uint8_t* h_in[NBUF] = {nullptr, nullptr, nullptr};        // pinned bounce buffers, only for pageable caller memory
    uint8_t* h_out[NBUF] = {nullptr, nullptr, nullptr};
    size_t in_bytes = 0, out_bytes = 0, bits_bytes = 0, h_in_bytes = 0, h_out_bytes = 0;
    cudaEvent_t ev_in[NBUF], ev_comp[NBUF], ev_out[NBUF];
    int n_events = 0;
    bool poisoned = false;                                    // a CUDA error occurred: the worker is rebuilt on the next call
    ~DeviceWorker()
    {
        if (!ctx) return;
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        for (int i = 0; i < NBUF; ++i) {
            cudaFree(d_in[i]); cudaFree(d_out[i]); cudaFree(d_bits[i]);
            if (h_in[i]) cudaFreeHost(h_in[i]);
            if (h_out[i]) cudaFreeHost(h_out[i]);
        }
        for (int i = 0; i < HB; ++i) if (h_bits[i]) cudaFreeHost(h_bits[i]);
        for (int i = 0; i < n_hb_events; ++i) cudaEventDestroy(ev_hb[i]);
        for (int i = 0; i < n_events; ++i) { cudaEventDestroy(ev_in[i]); cudaEventDestroy(ev_comp[i]); cudaEventDestroy(ev_out[i]); }
        if (s_in) cudaStreamDestroy(s_in);
        if (s_out) cudaStreamDestroy(s_out);
        prl_cuda_destroy(ctx);
        cudaGetLastError();
    }
};

std::mutex g_workers_mu;
// cached per device across calls; leaked on purpose (no CUDA calls from static destructors at exit)
auto& g_workers = *new std::map<int, std::shared_ptr<DeviceWorker>>();

std::shared_ptr<DeviceWorker> get_worker(int device, std::string* err)
{
    std::lock_guard<std::mutex> lk(g_workers_mu);
    auto it = g_workers.find(device);
    if (it != g_workers.end()) {
        if (!it->second->poisoned) return it->second;
        g_workers.erase(it);                                  // the last user still holds a reference; it is destroyed when that ends
    }
    std::shared_ptr<DeviceWorker> w(new DeviceWorker());
    int rc = prl_cuda_create(device, &w->ctx);
    if (rc) { *err = prl_cuda_last_error(nullptr); w->ctx = nullptr; return nullptr; }
    cudaError_t e = cudaStreamCreateWithFlags(&w->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&w->s_out, cudaStreamNonBlocking);
    for (int i = 0; e == cudaSuccess && i < DeviceWorker::NBUF; ++i) {
        e = cudaEventCreateWithFlags(&w->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->ev_comp[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&w->ev_out[i], cudaEventDisableTiming);
        if (e == cudaSuccess) w->n_events = i + 1;
    }
    for (int i = 0; e == cudaSuccess && i < DeviceWorker::HB; ++i) {
        e = cudaEventCreateWithFlags(&w->ev_hb[i], cudaEventDisableTiming);
        if (e == cudaSuccess) w->n_hb_events = i + 1;
        w->hb_left[i].store(0);
    }
    if (e != cudaSuccess) { *err = std::string("batch worker: ") + cudaGetErrorString(e); cudaGetLastError(); return nullptr; }
    g_workers[device] = w;
    return w;
}

// is this host range page-locked (cudaMallocHost / cudaHostRegister / prl_cuda_host_alloc)?
bool host_range_pinned(const void* p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// pages [p0, p1) of the batch on one device: 3-slot ring, H2D / kernels / D2H on three streams
int run_shard_locked(DeviceWorker* w, int method, const uint8_t* pages, int p0, int p1, int rows, int cols, int window,
                     const double* params, int morph_iters, uint8_t* masks, const prl_geom& g, std::string* err,
                     uint32_t* packed)
{
    prl_cuda_ctx* c = w->ctx;
#define SHARD_TRY(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { *err = std::string(#call) + ": " + cudaGetErrorString(_e); return PRL_E_CUDA; } } while (0)
    SHARD_TRY(cudaSetDevice(c->device));
    // byte masks returned as bits and expanded on the host (see UnpackPool) unless switched off
    const int unpack_threads = prl_cuda_batch_unpack_threads();
    const bool via_bits = !packed && unpack_threads > 0;
    // masks dense on the device: linear D2H.  Packed variants: aligned mask rows feed the pack kernel, the bits are dense.
    const size_t in_step = round16(cols), o_step = (packed || via_bits) ? round16((size_t)g.out_cols) : (size_t)g.out_cols;
    const size_t in_page = in_step * rows, out_page = o_step * g.out_rows;
    const size_t wpl = ((size_t)g.out_cols + 31) / 32, bits_page = wpl * g.out_rows * sizeof(uint32_t);
    const size_t host_in_page = (size_t)rows * cols, host_out_page = packed ? bits_page : (size_t)g.out_rows * g.out_cols;
    // chunk: about 64 MiB of input per slot (8 A4 pages; measured best of 4..64), at least 1 page
    int chunk = (int)std::max<size_t>(1, ((size_t)72 << 20) / in_page);
    const long long forced = g_batch_chunk_pages.load();
    if (forced > 0) chunk = (int)std::min<long long>(forced, 1 << 20);
    chunk = std::min(chunk, std::max(1, (p1 - p0 + 2) / 3));
    if (in_page * chunk > w->in_bytes || out_page * chunk > w->out_bytes) {
        SHARD_TRY(cudaDeviceSynchronize());
        for (int i = 0; i < DeviceWorker::NBUF; ++i) {
            cudaFree(w->d_in[i]); cudaFree(w->d_out[i]); w->d_in[i] = w->d_out[i] = nullptr;
        }
        w->in_bytes = w->out_bytes = 0;
        for (int i = 0; i < DeviceWorker::NBUF; ++i) {
            SHARD_TRY(cudaMalloc((void**)&w->d_in[i], in_page * chunk));
            SHARD_TRY(cudaMalloc((void**)&w->d_out[i], out_page * chunk + 16));
        }
        w->in_bytes = in_page * chunk; w->out_bytes = out_page * chunk;
    }
    if ((packed || via_bits) && bits_page * chunk > w->bits_bytes) {
        SHARD_TRY(cudaDeviceSynchronize());
        for (int i = 0; i < DeviceWorker::NBUF; ++i) { cudaFree(w->d_bits[i]); w->d_bits[i] = nullptr; }
        w->bits_bytes = 0;
        for (int i = 0; i < DeviceWorker::NBUF; ++i) SHARD_TRY(cudaMalloc((void**)&w->d_bits[i], bits_page * chunk));
        w->bits_bytes = bits_page * chunk;
    }
    // Pageable caller memory: cudaMemcpyAsync on it is a synchronous, driver-staged copy and the three-stream overlap is
    // lost.  Stage through pinned bounce buffers instead (one memcpy per direction on this thread: host-memcpy bound,
    // but the DMA and the kernels overlap it).  Page-locked memory (prl_cuda_host_alloc / prl_cuda_host_register /
    // cudaMallocHost) goes to the copy engines directly and is what the quoted end-to-end throughput needs.
    const bool stage_pageable = g_batch_pageable.load() != 0;
    const bool bounce_in = stage_pageable && !host_range_pinned(pages);
    const bool bounce_out = !via_bits && stage_pageable && !host_range_pinned(packed ? (const void*)packed : (const void*)masks);
    if (via_bits) {
        if (bits_page * chunk > w->h_bits_bytes) {
            SHARD_TRY(cudaDeviceSynchronize());
            for (int i = 0; i < DeviceWorker::HB; ++i) { if (w->h_bits[i]) cudaFreeHost(w->h_bits[i]); w->h_bits[i] = nullptr; }
            w->h_bits_bytes = 0;
            for (int i = 0; i < DeviceWorker::HB; ++i) SHARD_TRY(cudaMallocHost((void**)&w->h_bits[i], bits_page * chunk));
            w->h_bits_bytes = bits_page * chunk;
        }
        w->pool.start(unpack_threads);
    }
    // via_bits: chunk j sits in host slot j % HB; its expansion is handed to the pool two iterations after its D2H was queued
    // and must be over before the slot is reused HB iterations later
    struct HostChunk { int p, np; };
    HostChunk hchunk[DeviceWorker::HB] = {};
    const bool pool_waits = g_batch_unpack_lag.load() == 0;                   // the pool's threads wait for the chunk themselves
    auto expand = [&](int j) -> int {                                         // chunk j is on its way: hand its pages to the pool
        const int hs = j % DeviceWorker::HB;
        cudaEvent_t arrived = pool_waits ? w->ev_hb[hs] : nullptr;
        if (!pool_waits) {
            cudaError_t e = cudaEventSynchronize(w->ev_hb[hs]);
            if (e != cudaSuccess) { *err = std::string("cudaEventSynchronize: ") + cudaGetErrorString(e); return PRL_E_CUDA; }
        }
        const int bands = std::max(1, std::min(unpack_threads, 4));           // a page in a few bands: short jobs, even load
        w->hb_left[hs].store(hchunk[hs].np * bands);
        for (int i = 0; i < hchunk[hs].np; ++i)
            for (int b = 0; b < bands; ++b) {
                const int r0 = (int)((long long)g.out_rows * b / bands), r1 = (int)((long long)g.out_rows * (b + 1) / bands);
                const uint32_t* jb = w->h_bits[hs] + ((size_t)i * g.out_rows + r0) * wpl;
                uint8_t* jd = masks + ((size_t)(hchunk[hs].p + i) * g.out_rows + r0) * g.out_cols;
                const int jr = r1 - r0, jc = g.out_cols;
                const int dev = c->device;
                w->pool.submit([arrived, dev, jb, wpl, jd, jr, jc] {
                    // (the pool's threads start on device 0: waiting there would open a context on it from every process of a box)
                    if (arrived) { cudaSetDevice(dev); cudaEventSynchronize(arrived); }   // a failed copy surfaces at the stream synchronisation below
                    unpack_rows(jb, wpl, jd, (size_t)jc, jr, jc);
                }, &w->hb_left[hs]);
            }
        return PRL_OK;
    };
    if (bounce_in && host_in_page * chunk > w->h_in_bytes) {
        SHARD_TRY(cudaDeviceSynchronize());
        for (int i = 0; i < DeviceWorker::NBUF; ++i) { if (w->h_in[i]) cudaFreeHost(w->h_in[i]); w->h_in[i] = nullptr; }
        w->h_in_bytes = 0;
        for (int i = 0; i < DeviceWorker::NBUF; ++i) SHARD_TRY(cudaMallocHost((void**)&w->h_in[i], host_in_page * chunk));
        w->h_in_bytes = host_in_page * chunk;
    }
    if (bounce_out && host_out_page * chunk > w->h_out_bytes) {
        SHARD_TRY(cudaDeviceSynchronize());
        for (int i = 0; i < DeviceWorker::NBUF; ++i) { if (w->h_out[i]) cudaFreeHost(w->h_out[i]); w->h_out[i] = nullptr; }
        w->h_out_bytes = 0;
        for (int i = 0; i < DeviceWorker::NBUF; ++i) SHARD_TRY(cudaMallocHost((void**)&w->h_out[i], host_out_page * chunk));
        w->h_out_bytes = host_out_page * chunk;
    }
    uint8_t* host_out = packed ? reinterpret_cast<uint8_t*>(packed) : masks;
    struct Pending { int p, np; };
    Pending pend[DeviceWorker::NBUF] = {{0, 0}, {0, 0}, {0, 0}};           // bounce_out: chunks whose results sit in h_out[slot]
    int it = 0;
    for (int p = p0; p < p1; p += chunk, ++it) {
        const int np = std::min(chunk, p1 - p);
        const int slot = it % DeviceWorker::NBUF;
        if (it >= DeviceWorker::NBUF) {
            if (bounce_in) SHARD_TRY(cudaEventSynchronize(w->ev_in[slot]));   // h_in[slot] has left the host
            if (bounce_out && pend[slot].np) {                                // h_out[slot] arrived: hand it to the caller
                SHARD_TRY(cudaEventSynchronize(w->ev_out[slot]));
                memcpy(host_out + (size_t)pend[slot].p * host_out_page, w->h_out[slot], host_out_page * pend[slot].np);
                pend[slot].np = 0;
            }
            SHARD_TRY(cudaStreamWaitEvent(w->s_in, w->ev_comp[slot], 0));     // d_in[slot] consumed
            SHARD_TRY(cudaStreamWaitEvent(c->stream, w->ev_out[slot], 0));    // d_out[slot] drained
        }
        const int hs = it % DeviceWorker::HB;
        if (via_bits && it >= DeviceWorker::HB) w->pool.wait(w->hb_left[hs]);   // the chunk that used this host slot is expanded
        const uint8_t* hsrc = pages + (size_t)p * host_in_page;
        if (bounce_in) { memcpy(w->h_in[slot], hsrc, host_in_page * np); hsrc = w->h_in[slot]; }
        SHARD_TRY(copy2d(w->d_in[slot], in_step, hsrc, cols, cols, (size_t)rows * np, cudaMemcpyHostToDevice, w->s_in));
        SHARD_TRY(cudaEventRecord(w->ev_in[slot], w->s_in));
        SHARD_TRY(cudaStreamWaitEvent(c->stream, w->ev_in[slot], 0));
        int rc = prl_cuda_binarize_local_batch_dev(c, method, w->d_in[slot], np, rows, cols, in_step, in_page, window,
                                                   params, morph_iters, w->d_out[slot], o_step, out_page);
        if (rc) { *err = c->err; return rc; }
        if (packed || via_bits) {
            rc = prl_k_pack_mask(c, w->d_out[slot], np, g.out_rows, g.out_cols, o_step, out_page, w->d_bits[slot]);
            if (rc) { *err = c->err; return rc; }
        }
        SHARD_TRY(cudaEventRecord(w->ev_comp[slot], c->stream));
        SHARD_TRY(cudaStreamWaitEvent(w->s_out, w->ev_comp[slot], 0));
        uint8_t* hdst = bounce_out ? w->h_out[slot] : host_out + (size_t)p * host_out_page;
        if (via_bits) {
            SHARD_TRY(cudaMemcpyAsync(w->h_bits[hs], w->d_bits[slot], bits_page * np, cudaMemcpyDeviceToHost, w->s_out));
            SHARD_TRY(cudaEventRecord(w->ev_hb[hs], w->s_out));
            hchunk[hs] = HostChunk{p, np};
        } else if (packed)
            SHARD_TRY(cudaMemcpyAsync(hdst, w->d_bits[slot], bits_page * np, cudaMemcpyDeviceToHost, w->s_out));
        else
            SHARD_TRY(copy2d(hdst, g.out_cols, w->d_out[slot], o_step, g.out_cols, (size_t)g.out_rows * np, cudaMemcpyDeviceToHost, w->s_out));
        SHARD_TRY(cudaEventRecord(w->ev_out[slot], w->s_out));
        if (bounce_out) pend[slot] = Pending{p, np};
        if (via_bits && pool_waits) { rc = expand(it); if (rc) return rc; }
        else if (via_bits && it >= 2) { rc = expand(it - 2); if (rc) return rc; }
    }
    if (via_bits) {
        if (!pool_waits) for (int j = std::max(0, it - 2); j < it; ++j) { int rc = expand(j); if (rc) return rc; }
        for (int j = std::max(0, it - DeviceWorker::HB); j < it; ++j) w->pool.wait(w->hb_left[j % DeviceWorker::HB]);
    }
    SHARD_TRY(cudaStreamSynchronize(w->s_out));
    SHARD_TRY(cudaStreamSynchronize(c->stream));
    SHARD_TRY(cudaStreamSynchronize(w->s_in));
    if (bounce_out)
        for (int s2 = 0; s2 < DeviceWorker::NBUF; ++s2)
            if (pend[s2].np) memcpy(host_out + (size_t)pend[s2].p * host_out_page, w->h_out[s2], host_out_page * pend[s2].np);
#undef SHARD_TRY
    return PRL_OK;
}

// Serialises callers per device; after ANY failure nothing is left in flight on the caller's buffers (the copies are
// drained before returning) and the worker is rebuilt on the next call.
int run_shard(DeviceWorker* w, int method, const uint8_t* pages, int p0, int p1, int rows, int cols, int window,
              const double* params, int morph_iters, uint8_t* masks, const prl_geom& g, std::string* err,
              uint32_t* packed = nullptr)
{
    std::lock_guard<std::mutex> lk(w->busy);
    const int rc = run_shard_locked(w, method, pages, p0, p1, rows, cols, window, params, morph_iters, masks, g, err, packed);
    if (rc) {
        cudaSetDevice(w->ctx->device);
        for (int i = 0; i < DeviceWorker::HB; ++i) w->pool.wait(w->hb_left[i]);      // expansions already handed out write into `masks`
        cudaStreamSynchronize(w->s_in); cudaStreamSynchronize(w->ctx->stream); cudaStreamSynchronize(w->s_out);
        cudaGetLastError();
        w->poisoned = true;
    }
    return rc;
}

}  // namespace

static int binarize_batch_impl(const int* devices, int n_dev, int method, const uint8_t* pages, int n_pages,
                               int rows, int cols, int window, const double* params, int morph_iters,
                               uint8_t* masks, uint32_t* packed)
{
    if (!pages || (!masks && !packed) || !params || n_pages <= 0) return prl_set_err(nullptr, PRL_E_INVALID, "null pointer or empty batch");
    prl_geom g;
    int rc = prl_make_geom(method, rows, cols, window, &g);
    if (rc) return prl_set_err(nullptr, rc, "bad geometry / window");
    std::vector<int> devs;
    if (!devices || n_dev <= 0) { int n = prl_cuda_device_count(); for (int i = 0; i < n; ++i) devs.push_back(i); }
    else devs.assign(devices, devices + n_dev);
    if (devs.empty()) return prl_set_err(nullptr, PRL_E_CUDA, "no CUDA device available (libprlib_cuda has no CPU fallback)");
    const int G = (int)devs.size();
    std::vector<int> rcs(G, PRL_OK);
    std::vector<std::string> errs(G);
    std::vector<std::thread> threads;
    for (int gi = 0; gi < G; ++gi) {
        // contiguous page ranges: device gi <- pages [gi*N/G, (gi+1)*N/G)
        const int p0 = (int)((long long)gi * n_pages / G), p1 = (int)((long long)(gi + 1) * n_pages / G);
        if (p1 <= p0) continue;
        auto job = [&, gi, p0, p1]() {
            std::shared_ptr<DeviceWorker> w = get_worker(devs[gi], &errs[gi]);
            if (!w) { rcs[gi] = PRL_E_CUDA; return; }
            rcs[gi] = run_shard(w.get(), method, pages, p0, p1, rows, cols, window, params, morph_iters, masks, g, &errs[gi], packed);
        };
        if (G == 1) job(); else threads.emplace_back(job);
    }
    for (auto& t : threads) t.join();
    for (int gi = 0; gi < G; ++gi)
        if (rcs[gi]) {
            char buf[64]; snprintf(buf, sizeof buf, "device %d: ", devs[gi]);
            return prl_set_err(nullptr, rcs[gi], (std::string(buf) + errs[gi]).c_str());
        }
    return PRL_OK;
}

extern "C" int prl_cuda_binarize_batch(const int* devices, int n_dev, int method, const uint8_t* pages, int n_pages,
                                       int rows, int cols, int window, const double* params, int morph_iters,
                                       uint8_t* masks)
{
    return binarize_batch_impl(devices, n_dev, method, pages, n_pages, rows, cols, window, params, morph_iters, masks, nullptr);
}

extern "C" int prl_cuda_binarize_batch_packed(const int* devices, int n_dev, int method, const uint8_t* pages, int n_pages,
                                              int rows, int cols, int window, const double* params, int morph_iters,
                                              uint32_t* bits)
{
    return binarize_batch_impl(devices, n_dev, method, pages, n_pages, rows, cols, window, params, morph_iters, nullptr, bits);
}

// ------------------------------------------------------------------------------------------------
// page-locked host memory for the batch loader, process-wide options
// ------------------------------------------------------------------------------------------------
extern "C" int prl_cuda_host_alloc(size_t bytes, void** out)
{
    if (!out || bytes == 0) return prl_set_err(nullptr, PRL_E_INVALID, "bad argument");
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; return prl_set_err(nullptr, PRL_E_NOMEM, "cudaHostAlloc", e); }
    return PRL_OK;
}

extern "C" int prl_cuda_host_free(void* p)
{
    if (!p) return PRL_OK;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) { cudaGetLastError(); return prl_set_err(nullptr, PRL_E_CUDA, "cudaFreeHost", e); }
    return PRL_OK;
}

extern "C" int prl_cuda_host_register(void* p, size_t bytes)
{
    if (!p || bytes == 0) return prl_set_err(nullptr, PRL_E_INVALID, "bad argument");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return prl_set_err(nullptr, PRL_E_CUDA, "cudaHostRegister", e); }
    return PRL_OK;
}

extern "C" int prl_cuda_host_unregister(void* p)
{
    if (!p) return PRL_OK;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return prl_set_err(nullptr, PRL_E_CUDA, "cudaHostUnregister", e); }
    return PRL_OK;
}

extern "C" int prl_cuda_set_global_option(const char* name, long long value)
{
    if (!name) return PRL_E_INVALID;
    if (strcmp(name, "batch_chunk_pages") == 0) g_batch_chunk_pages.store(value < 0 ? 0 : value);
    else if (strcmp(name, "batch_stage_pageable") == 0) g_batch_pageable.store(value != 0);
    else if (strcmp(name, "batch_unpack_nt") == 0) g_batch_unpack_nt.store(value != 0);
    else if (strcmp(name, "batch_unpack_lag") == 0) g_batch_unpack_lag.store(value != 0);
    else if (strcmp(name, "batch_unpack_threads") == 0) g_batch_unpack_threads.store(std::min<long long>(std::max<long long>(value, -1), 64));
    else return prl_set_err(nullptr, PRL_E_INVALID, "unknown global option");
    return PRL_OK;
}
