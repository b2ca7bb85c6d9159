// morph.cu -- morphology tail of the five local-statistics binarizers.
//
// Replaces cv::dilate / cv::erode with the default 3x3 element and `n` iterations
// (binarizeSauvola.cpp:125-134, binarizeNiblack.cpp:115-127, binarizeWolfJolion.cpp:138-147,
// binarizeNICK.cpp:134-143, binarizeFeng.cpp:151-163).  n iterations of a 3x3 rectangle are one
// (2n+1)x(2n+1) rectangle; pixels outside the image are ignored (morphologyDefaultBorderValue),
// so each pass is a separable running max / min clipped to the image.
#include "common.cuh"

namespace {

// one separable pass: horizontal (DIR 0) or vertical (DIR 1) max (IS_MAX) / min over [-n, n]
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

// In place on d_mask; d_tmp is a same-shaped scratch.
int prl_k_morph(prl_cuda_ctx* ctx, uint8_t* d_mask, uint8_t* d_tmp, int n_pages, int rows, int cols, size_t step,
                size_t page_stride, size_t tmp_step, size_t tmp_page_stride, int iters)
{
    if (iters == 0) return PRL_OK;
    if (rows > 65535 || n_pages > 65535) return prl_set_err(ctx, PRL_E_UNSUPPORTED, "grid too large");
    const int n = iters > 0 ? iters : -iters;
    if (iters > 0) {   // closing: dilate then erode
        morph_op<true>(ctx, d_mask, d_tmp, n_pages, rows, cols, step, page_stride, tmp_step, tmp_page_stride, n);
        morph_op<false>(ctx, d_mask, d_tmp, n_pages, rows, cols, step, page_stride, tmp_step, tmp_page_stride, n);
    } else {           // opening: erode then dilate
        morph_op<false>(ctx, d_mask, d_tmp, n_pages, rows, cols, step, page_stride, tmp_step, tmp_page_stride, n);
        morph_op<true>(ctx, d_mask, d_tmp, n_pages, rows, cols, step, page_stride, tmp_step, tmp_page_stride, n);
    }
    PRL_CUDA_TRY(ctx, cudaGetLastError());
    return PRL_OK;
}
