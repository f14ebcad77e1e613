// HBM-bound row kernels of the quantizer:
//   gather + straight-through   models/autoencoders.py:218-222, 239-241
//   commitment / latent MSE     models/autoencoders.py:193-202 (forward and backward)
//   bincount                    models/autoencoders.py:165
// All of them stream every byte once with 8/16-byte accesses; reductions are fixed-order (deterministic).
#include <cuda_runtime.h>
#include "../../include/pero_b200.h"
#include "layout.h"

namespace pero {

constexpr int kMseBlocks = 592;   // 4 x 148 SMs; fixed so the summation tree never changes

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;   // valid in thread 0
}

// out[nl, d, hw] = x + (w[idx] - x): rows are read coalesced along d, transposed through shared
// memory in 64(d) x 32(hw) tiles and written coalesced along hw (channels-first, what the decoder
// projection conv expects).  The fp32 expression is the reference's, so the forward value is bit-equal.
// partial != NULL: the block also leaves the sum of (out - x)^2 over its elements in partial[linear block index]
// (the commitment / latent loss of models/autoencoders.py:198-200 needs exactly this mean; computing it here saves
// a second pass over out and x).
__global__ void __launch_bounds__(256)
gather_st_cf_kernel(const float* __restrict__ xr, const long long* __restrict__ idx, const float* __restrict__ w,
                    int D, int HW, float* __restrict__ out, float* __restrict__ partial) {
    __shared__ float tile[64][33];
    __shared__ float red[8];
    float sq = 0.f;
    const int nl = blockIdx.z, hw0 = blockIdx.x * 32, d0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int d = d0 + 2 * tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int hwl = ty + i * 8, hw = hw0 + hwl;
        float q0 = 0.f, q1 = 0.f;
        if (hw < HW) {
            const size_t n = (size_t)nl * HW + hw;
            const float* xrow = xr + n * D;
            const float* wrow = w + (size_t)__ldg(idx + n) * D;
            if (d < D) { const float x0 = __ldg(xrow + d); q0 = x0 + (__ldg(wrow + d) - x0); const float e = q0 - x0; sq = fmaf(e, e, sq); }
            if (d + 1 < D) { const float x1 = __ldg(xrow + d + 1); q1 = x1 + (__ldg(wrow + d + 1) - x1); const float e = q1 - x1; sq = fmaf(e, e, sq); }
        }
        tile[2 * tx][hwl] = q0;
        tile[2 * tx + 1][hwl] = q1;
    }
    __syncthreads();
    float* ol = out + (size_t)nl * D * HW;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int dd = d0 + ty + i * 8, hw = hw0 + tx;
        if (dd < D && hw < HW) ol[(size_t)dd * HW + hw] = tile[ty + i * 8][tx];
    }
    if (partial) {
        const float t = block_sum_256(sq, red);
        if (threadIdx.x == 0) partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
gather_st_rows_kernel(const float* __restrict__ xr, const long long* __restrict__ idx, const float* __restrict__ w,
                      int D, long long N, float* __restrict__ out, float* __restrict__ partial) {
    __shared__ float red[8];
    float sq = 0.f;
    const long long total = N * D;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long n = i / D;
        const int d = (int)(i - n * D);
        const float x = __ldg(xr + i);
        const float q = x + (__ldg(w + (size_t)__ldg(idx + n) * D + d) - x);
        out[i] = q;
        const float e = q - x;
        sq = fmaf(e, e, sq);
    }
    if (partial) {
        const float t = block_sum_256(sq, red);
        if (threadIdx.x == 0) partial[blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------- MSE
__global__ void __launch_bounds__(256)
mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, long long numel, float* __restrict__ partial) {
    __shared__ float sh[8];
    float s = 0.f;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    if (vec) {
        const long long n4 = numel >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        // 4 grid strides per iteration: 8 independent 16-byte loads in flight per thread (the kernel often runs with a
        // single CTA per SM beside a GEMM); the terms are added in the same order as a plain grid-stride loop would
        for (long long i = t0; i < n4; i += 4 * stride) {
            float4 u[4], v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long j = i + k * stride;
                u[k] = j < n4 ? __ldg(a4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[k] = j < n4 ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i + k * stride < n4) {
                    const float d0 = u[k].x - v[k].x, d1 = u[k].y - v[k].y, d2 = u[k].z - v[k].z, d3 = u[k].w - v[k].w;
                    s += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
                }
            }
        }
        for (long long i = (n4 << 2) + t0; i < numel; i += stride) { const float d = a[i] - b[i]; s += d * d; }
    } else {
        for (long long i = t0; i < numel; i += stride) { const float d = a[i] - b[i]; s += d * d; }
    }
    const float t = block_sum_256(s, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void __launch_bounds__(256)
mse_final_kernel(const float* __restrict__ partial, int nblocks, float numel, float scale_a, float scale_b,
                 float* __restrict__ out) {
    __shared__ float sh[8];
    float s = 0.f;
    for (int i = threadIdx.x; i < nblocks; i += blockDim.x) s += partial[i];
    const float t = block_sum_256(s, sh);
    if (threadIdx.x == 0) {
        const float mean = __fdiv_rn(t, numel);
        out[0] = __fadd_rn(__fmul_rn(scale_a, mean), __fmul_rn(scale_b, mean));   // each product rounded, as torch does
    }
}

__global__ void __launch_bounds__(256)
mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long numel, float coef,
               const float* __restrict__ grad_out, float* __restrict__ g_a, float* __restrict__ g_b) {
    const float c = coef * (grad_out ? __ldg(grad_out) : 1.0f);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += stride) {
        const float g = c * (__ldg(b + i) - __ldg(a + i));
        if (g_b) g_b[i] = g;
        if (g_a) g_a[i] = -g;
    }
}

// Backward of quantize + commitment loss in one pass: the straight-through estimator passes
// g_quantized unchanged and the commitment term adds coef * g_loss * (inputs - quantized).
__global__ void __launch_bounds__(256, 4)      // <= 64 registers: one CTA fits beside a resident GEMM CTA
st_commit_bwd_kernel(const float* __restrict__ gq, const float* __restrict__ q, const float* __restrict__ x, long long numel,
                     float coef, const float* __restrict__ grad_loss, float* __restrict__ gx) {
    const float c = coef * (grad_loss ? __ldg(grad_loss) : 1.0f);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(gq) | reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(x) |
                       reinterpret_cast<uintptr_t>(gx)) & 15) == 0;
    if (vec) {
        const long long n4 = numel >> 2;
        // 3 grid strides per iteration: 9 independent 16-byte loads in flight per thread (see mse_partial_kernel)
        for (long long i = t0; i < n4; i += 3 * stride) {
            float4 g[3], a[3], b[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const long long j = i + k * stride;
                if (j < n4) {
                    g[k] = __ldg(reinterpret_cast<const float4*>(gq) + j);
                    a[k] = __ldg(reinterpret_cast<const float4*>(q) + j);
                    b[k] = __ldg(reinterpret_cast<const float4*>(x) + j);
                }
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const long long j = i + k * stride;
                if (j < n4)
                    reinterpret_cast<float4*>(gx)[j] = make_float4(g[k].x + c * (b[k].x - a[k].x), g[k].y + c * (b[k].y - a[k].y),
                                                                  g[k].z + c * (b[k].z - a[k].z), g[k].w + c * (b[k].w - a[k].w));
            }
        }
        for (long long i = (n4 << 2) + t0; i < numel; i += stride) gx[i] = gq[i] + c * (x[i] - q[i]);
    } else {
        for (long long i = t0; i < numel; i += stride) gx[i] = gq[i] + c * (x[i] - q[i]);
    }
}

__global__ void counts_kernel(const long long* __restrict__ idx, long long N, long long K, unsigned long long* __restrict__ counts) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const long long k = idx[i];
    if (k >= 0 && k < K) atomicAdd(counts + k, 1ull);   // integer atomics: order-independent result
}

inline unsigned grid_for(long long work, int per_block, long long cap) {
    long long g = (work + per_block - 1) / per_block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace pero

using namespace pero;

extern "C" {

// Shared by pero_vq_gather_st (partial = NULL) and pero_vq_gather_st_mse; returns the number of blocks (= partial sums).
static int launch_gather_st(const float* x_rows, const int64_t* idx, const float* weight, int64_t n_lines, int64_t frames_per_line,
                            int channels_first, int64_t D, float* out, float* partial, cudaStream_t stream, long long* nblocks) {
    const int64_t N = n_lines * frames_per_line;
    if (channels_first) {
        if (n_lines > 65535) return PERO_ERR_BAD_SHAPE;
        dim3 grid((unsigned)((frames_per_line + 31) / 32), (unsigned)((D + 63) / 64), (unsigned)n_lines);
        *nblocks = (long long)grid.x * grid.y * grid.z;
        gather_st_cf_kernel<<<grid, 256, 0, stream>>>(x_rows, reinterpret_cast<const long long*>(idx), weight, (int)D,
                                                      (int)frames_per_line, out, partial);
    } else {
        const unsigned g = grid_for(N * D, 256, 148 * 16);
        *nblocks = g;
        gather_st_rows_kernel<<<g, 256, 0, stream>>>(x_rows, reinterpret_cast<const long long*>(idx), weight, (int)D, N, out,
                                                     partial);
    }
    return (int)cudaGetLastError();
}

static long long gather_st_blocks(int64_t n_lines, int64_t frames_per_line, int channels_first, int64_t D) {
    if (channels_first) return ((frames_per_line + 31) / 32) * ((D + 63) / 64) * n_lines;
    return grid_for(n_lines * frames_per_line * D, 256, 148 * 16);
}

int pero_vq_gather_st(const float* x_rows, const int64_t* idx, const float* weight, int64_t n_lines,
                      int64_t frames_per_line, int channels_first, int64_t K, int64_t D, float* out,
                      pero_stream_t stream) {
    if (n_lines < 0 || frames_per_line < 0) return PERO_ERR_BAD_SHAPE;
    const int64_t N = n_lines * frames_per_line;
    if (N == 0) return PERO_OK;
    if (!x_rows || !idx || !weight || !out) return PERO_ERR_NULL;
    if (K <= 0 || D <= 0 || D > 65536) return PERO_ERR_BAD_SHAPE;
    long long nb = 0;
    return launch_gather_st(x_rows, idx, weight, n_lines, frames_per_line, channels_first, D, out, nullptr,
                            reinterpret_cast<cudaStream_t>(stream), &nb);
}

size_t pero_vq_gather_st_mse_workspace_bytes(int64_t n_lines, int64_t frames_per_line, int channels_first, int64_t D) {
    if (n_lines <= 0 || frames_per_line <= 0 || D <= 0) return 0;
    return align256((size_t)gather_st_blocks(n_lines, frames_per_line, channels_first, D) * sizeof(float));
}

int pero_vq_gather_st_mse(const float* x_rows, const int64_t* idx, const float* weight, int64_t n_lines,
                          int64_t frames_per_line, int channels_first, int64_t K, int64_t D, float* out, float scale_a,
                          float scale_b, float* loss_out, void* workspace, size_t workspace_bytes, pero_stream_t stream) {
    if (n_lines <= 0 || frames_per_line <= 0) return PERO_ERR_BAD_SHAPE;
    const int64_t N = n_lines * frames_per_line;
    if (!x_rows || !idx || !weight || !out || !loss_out || !workspace) return PERO_ERR_NULL;
    if (K <= 0 || D <= 0 || D > 65536) return PERO_ERR_BAD_SHAPE;
    if (workspace_bytes < pero_vq_gather_st_mse_workspace_bytes(n_lines, frames_per_line, channels_first, D)) return PERO_ERR_WORKSPACE;
    float* partial = static_cast<float*>(workspace);
    long long nb = 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int rc = launch_gather_st(x_rows, idx, weight, n_lines, frames_per_line, channels_first, D, out, partial, st, &nb);
    if (rc) return rc;
    mse_final_kernel<<<1, 256, 0, st>>>(partial, (int)nb, (float)(N * D), scale_a, scale_b, loss_out);
    return (int)cudaGetLastError();
}

size_t pero_mse_workspace_bytes(int64_t numel) {
    (void)numel;
    return align256(kMseBlocks * sizeof(float));
}

int pero_mse_fwd(const float* a, const float* b, int64_t numel, float scale_a, float scale_b, float* out,
                 void* workspace, size_t workspace_bytes, pero_stream_t stream) {
    if (!a || !b || !out || !workspace) return PERO_ERR_NULL;
    if (numel <= 0) return PERO_ERR_BAD_SHAPE;
    if (workspace_bytes < pero_mse_workspace_bytes(numel)) return PERO_ERR_WORKSPACE;
    float* partial = static_cast<float*>(workspace);
    const int blocks = (int)grid_for(numel, 256 * 8, kMseBlocks);
    mse_partial_kernel<<<blocks, 256, 0, stream>>>(a, b, numel, partial);
    mse_final_kernel<<<1, 256, 0, stream>>>(partial, blocks, (float)numel, scale_a, scale_b, out);
    return (int)cudaGetLastError();
}

int pero_mse_bwd(const float* a, const float* b, int64_t numel, float coef, const float* grad_out, float* g_a,
                 float* g_b, pero_stream_t stream) {
    if (!a || !b || (!g_a && !g_b)) return PERO_ERR_NULL;
    if (numel <= 0) return PERO_ERR_BAD_SHAPE;
    mse_bwd_kernel<<<grid_for(numel, 256 * 4, 148 * 16), 256, 0, stream>>>(a, b, numel, coef, grad_out, g_a, g_b);
    return (int)cudaGetLastError();
}

int pero_vq_st_commit_bwd(const float* g_quantized, const float* quantized, const float* inputs, int64_t numel, float coef,
                          const float* grad_loss, float* g_inputs, pero_stream_t stream) {
    if (!g_quantized || !quantized || !inputs || !g_inputs) return PERO_ERR_NULL;
    if (numel <= 0) return PERO_ERR_BAD_SHAPE;
    // 2 CTAs per SM at most, ~8 float4 triples per thread: enough loads in flight alone and beside a resident GEMM CTA
    st_commit_bwd_kernel<<<grid_for(numel, 256 * 16, 148 * 2), 256, 0, stream>>>(g_quantized, quantized, inputs, numel, coef,
                                                                               grad_loss, g_inputs);
    return (int)cudaGetLastError();
}

int pero_vq_counts(const int64_t* idx, int64_t N, int64_t K, int64_t* counts, pero_stream_t stream) {
    if (!counts || (N > 0 && !idx)) return PERO_ERR_NULL;
    if (K <= 0 || N < 0) return PERO_ERR_BAD_SHAPE;
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)K * 8, stream);
    if (e != cudaSuccess) return (int)e;
    if (N > 0)
        counts_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const long long*>(idx), N, K,
                                                                       reinterpret_cast<unsigned long long*>(counts));
    return (int)cudaGetLastError();
}

}  // extern "C"
