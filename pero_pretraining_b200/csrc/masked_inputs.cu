// The two callers' sides of the masked-label head that SURVEY 8f lists next:
//   * pero_mask_pixels: the backbone's input-pixel masking (models/transformers.py:53-68, TransformerEncoder.mask) on the
//     device, driven by the SAME masked-frame list the masked cross-entropy uses, so the numpy mask makes one trip
//     to the GPU per step instead of three (transformers.py:57, masked_pretraining/model.py:44-45).
//   * pero_head_argmax_prepare: the head as a "codebook" of the distance kernel, so that the label prediction of every
//     frame (masked_pretraining/visualizer.py:32, torch.argmax(output['output'], dim=-1)) is one run of the fused
//     GEMM + arg-min kernel and the [N, V] logits never exist:  argmax_v (h.W_v + b_v) = argmin_v (-2 b_v - 2 h.W_v).
#include <cuda_bf16.h>
#include <math_constants.h>
#include "../../include/pero_b200.h"
#include "layout.h"

namespace pero {

// One thread per (masked frame, channel, pixel row): the frame's 8-pixel column segment of that row is overwritten with
// the tile's row (32 contiguous bytes; x[mask == 1] = pattern[mask == 1] of the reference, the pattern being the tile
// repeated every `pw` pixels).
__global__ void __launch_bounds__(256)
mask_pixels_kernel(float* __restrict__ x, const int* __restrict__ rows, long long M, int C, int H, int W, int frames_per_line,
                   int pw, const float* __restrict__ tile) {
    const long long total = M * C * H;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long m = i / (C * H);
        const int ch = (int)(i - m * (C * H));           // c * H + h
        const int r = __ldg(rows + m);
        const int n = r / frames_per_line, t = r - n * frames_per_line;
        float* dst = x + ((long long)n * C * H + ch) * W + (long long)t * pw;
        const float* src = tile + (long long)ch * pw;
        const int w0 = t * pw;
        for (int j = 0; j < pw; ++j)
            if (w0 + j < W) dst[j] = __ldg(src + j);
    }
}

// bf16 copy of W (zero-padded to Dp columns) and the column vector -2 b (+inf beyond V) in the prepared-codebook layout.
__global__ void head_argmax_prepare_kernel(const float* __restrict__ w, const float* __restrict__ bias, int V, int Dh, int Dp,
                                           int Vp, __nv_bfloat16* __restrict__ cb, float* __restrict__ cvec) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= Vp) return;
    if (k >= V) { if (lane == 0) cvec[k] = CUDART_INF_F; return; }
    const float* row = w + (size_t)k * Dh;
    __nv_bfloat16* dst = cb + (size_t)k * Dp;
    for (int d = lane; d < Dp; d += 32) dst[d] = __float2bfloat16_rn(d < Dh ? row[d] : 0.f);
    if (lane == 0) cvec[k] = bias ? -2.0f * __ldg(bias + k) : 0.f;
}

}  // namespace pero

using namespace pero;

extern "C" {

int pero_mask_pixels(float* x, int64_t n_lines, int64_t C, int64_t H, int64_t W, const int32_t* rows, int64_t M,
                     int64_t frames_per_line, int64_t patch_width, const float* tile, pero_stream_t stream) {
    if (M == 0) return PERO_OK;
    if (!x || !rows || !tile) return PERO_ERR_NULL;
    if (n_lines <= 0 || C <= 0 || H <= 0 || W <= 0 || M < 0 || frames_per_line <= 0 || patch_width <= 0 ||
        frames_per_line * patch_width < W - patch_width + 1 || n_lines * frames_per_line > (1ll << 31) - 1)
        return PERO_ERR_BAD_SHAPE;
    long long blocks = (M * C * H + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    mask_pixels_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, rows, M, (int)C, (int)H, (int)W, (int)frames_per_line,
                                                            (int)patch_width, tile);
    return (int)cudaGetLastError();
}

int pero_head_argmax_prepare(const float* W, const float* bias, int64_t V, int64_t Dh, void* codebook, size_t codebook_bytes,
                             pero_stream_t stream) {
    if (!W || !codebook) return PERO_ERR_NULL;
    if (V <= 0 || Dh <= 0 || V > (1ll << 30) || Dh > 65536) return PERO_ERR_BAD_SHAPE;
    const CodebookLayout cl = codebook_layout(V, Dh);
    if (codebook_bytes < cl.total) return PERO_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(codebook) & 255) return PERO_ERR_BAD_ALIGN;
    char* base = static_cast<char*>(codebook);
    head_argmax_prepare_kernel<<<(unsigned)((cl.Kp + 7) / 8), 256, 0, stream>>>(
        W, bias, (int)V, (int)Dh, (int)cl.Dp, (int)cl.Kp, reinterpret_cast<__nv_bfloat16*>(base + cl.cb_off),
        reinterpret_cast<float*>(base + cl.cnorm_off));
    return (int)cudaGetLastError();
}

}  // extern "C"
