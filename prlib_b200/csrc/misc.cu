// misc.cu -- synthpage-v2 device generator and the BGR->gray front step.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t mix32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t key32(uint32_t s, uint32_t p, uint32_t a, uint32_t b)
{
    return s * 0x9E3779B1u + p * 0x85EBCA77u + a * 0xC2B2AE3Du + b * 0x27D4EB2Fu;
}

// synthpage-v2: bit-identical to oracle/prl_oracle.py:synth_page.
__global__ void __launch_bounds__(256)
synth_kernel(uint8_t* __restrict__ dst, int rows, int cols, size_t step, size_t page_stride, uint32_t seed,
             uint32_t first_page)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= cols) return;
    const uint32_t page = first_page + blockIdx.z;
    const uint32_t ux = (uint32_t)x, uy = (uint32_t)y;
    const uint32_t noise = mix32(key32(seed, page, uy, ux)) & 15u;
    const uint32_t illum = (40u * ux) / (uint32_t)cols + (24u * uy) / (uint32_t)rows;
    const uint32_t stain = (mix32(key32(seed ^ 0x5BD1E995u, page, uy / 64u, ux / 64u)) >> 4) & 31u;
    const uint32_t bg = 235u - illum - stain - noise;
    const bool band = ((y % 48) < 26) && x >= 150 && x < cols - 150 && y >= 200 && y < rows - 200;
    const uint32_t b = mix32(key32(seed ^ 0xA5A5A5A5u, page, uy / 3u, ux / 3u));
    const bool ink = band && ((b & 0xFFu) < 56u);
    const uint32_t inkv = 24u + ((b >> 8) & 127u) + (noise >> 1);
    dst[(size_t)blockIdx.z * page_stride + (size_t)y * step + x] = (uint8_t)(ink ? inkv : bg);
}

// cv::cvtColor(BGR2GRAY) for 8U, OpenCV 4.x: (B*3735 + G*19235 + R*9798 + (1<<14)) >> 15
// (binarizeSauvola.cpp:49-52).
__global__ void __launch_bounds__(256)
bgr2gray_kernel(const uint8_t* __restrict__ src, int rows, int cols, size_t step, int channels,
                uint8_t* __restrict__ dst, size_t dst_step, int rgb)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= cols) return;
    const uint8_t* p = src + (size_t)y * step + (size_t)x * channels;
    // rgb: COLOR_RGB2GRAY (channel 0 is red), as binarizeLocalOtsu.cpp:63 calls it
    const uint32_t v = (uint32_t)p[rgb ? 2 : 0] * 3735u + (uint32_t)p[1] * 19235u + (uint32_t)p[rgb ? 0 : 2] * 9798u + 16384u;
    dst[(size_t)y * dst_step + x] = (uint8_t)(v >> 15);
}

// mask (0 / 255 bytes) -> 1 bit per pixel in Leptonica's PIX layout (the reference's other image container,
// src/formatConvert.cpp:57-69): rows of wpl 32-bit words, pixel x in word x >> 5 at bit 31 - (x & 31), 1 = black.
// One thread per output word: 32 mask bytes (two 16-byte loads when the row is aligned), gathered by dp4a.
__device__ __forceinline__ int dp4a_su_m(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__global__ void __launch_bounds__(256)
pack_mask_kernel(const uint8_t* __restrict__ mask, int rows, int cols, size_t step, size_t page_stride, int wpl,
                 uint32_t* __restrict__ bits, int aligned)
{
    const int wx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (wx >= wpl) return;
    const uint8_t* p = mask + (size_t)blockIdx.z * page_stride + (size_t)y * step + 32 * wx;
    const int nvalid = min(32, cols - 32 * wx);
    uint32_t white = 0;                                   // bit i = pixel 32 wx + i is white (255)
    if (aligned && nvalid > 16) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p)), b = __ldg(reinterpret_cast<const uint4*>(p + 16));
        const int v0 = dp4a_su_m(a.y, 0x80402010u, dp4a_su_m(a.x, 0x08040201u, 0));      // bytes are 0 / -1: v = -(8 bits)
        const int v1 = dp4a_su_m(a.w, 0x80402010u, dp4a_su_m(a.z, 0x08040201u, 0));
        const int v2 = dp4a_su_m(b.y, 0x80402010u, dp4a_su_m(b.x, 0x08040201u, 0));
        const int v3 = dp4a_su_m(b.w, 0x80402010u, dp4a_su_m(b.z, 0x08040201u, 0));
        white = 0u - ((uint32_t)v0 + ((uint32_t)v1 << 8) + ((uint32_t)v2 << 16) + ((uint32_t)v3 << 24));
    } else {
        for (int i = 0; i < nvalid; ++i) white |= (uint32_t)(p[i] != 0) << i;
    }
    const uint32_t valid = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
    bits[((size_t)blockIdx.z * rows + y) * wpl + wx] = __brev(~white & valid);
}

}  // namespace

int prl_k_pack_mask(prl_cuda_ctx* ctx, const uint8_t* d_mask, int n_pages, int rows, int cols, size_t step,
                    size_t page_stride, uint32_t* d_bits)
{
    if (rows > 65535 || n_pages > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    const int wpl = (cols + 31) / 32;
    // 16-byte loads need aligned rows and a row pitch that covers the last, partly valid 32-byte group
    const int aligned = ((((uintptr_t)d_mask) | step | page_stride) & 15) == 0 && step >= (size_t)(((cols + 15) / 16) * 16);
    prl_launch_scope ls(ctx, FAM_PACK);
    pack_mask_kernel<<<dim3((wpl + 255) / 256, rows, n_pages), 256, 0, ctx->stream>>>(d_mask, rows, cols, step, page_stride,
                                                                                    wpl, d_bits, aligned);
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

int prl_k_synth(prl_cuda_ctx* ctx, uint8_t* d_dst, int n_pages, int rows, int cols, size_t step,
                size_t page_stride, uint32_t seed, uint32_t first_page)
{
    if (rows > 65535 || n_pages > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    prl_launch_scope ls(ctx, FAM_SYNTH);
    synth_kernel<<<dim3((cols + 255) / 256, rows, n_pages), 256, 0, ctx->stream>>>(d_dst, rows, cols, step,
                                                                                  page_stride, seed, first_page);
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}

int prl_k_bgr2gray(prl_cuda_ctx* ctx, const uint8_t* d_src, int rows, int cols, size_t step, int channels,
                   uint8_t* d_dst, size_t dst_step, bool rgb)
{
    if (rows > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    prl_launch_scope ls(ctx, FAM_BGR2GRAY);
    bgr2gray_kernel<<<dim3((cols + 255) / 256, rows), 256, 0, ctx->stream>>>(d_src, rows, cols, step, channels,
                                                                           d_dst, dst_step, rgb ? 1 : 0);
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
