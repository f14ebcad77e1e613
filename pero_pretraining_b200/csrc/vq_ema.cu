// EMA codebook update (models/autoencoders.py:225-237), deterministic:
//   1. sort (codeword, frame) pairs by codeword  — radix sort is stable, so frames stay ascending
//   2. segment the sorted list (binary search per codeword) -> counts
//   3. chunked segmented sum of the fp32 frame rows in sorted order: every 32 sorted positions form one
//      chunk; runs that lie inside a chunk are summed and stored directly, runs that cross chunk borders
//      leave a head/tail partial that a second kernel adds up in chunk order.  No atomics, so the sums
//      are bit-identical run to run, and a collapsed codebook (one codeword owning every frame, as in
//      the reference's cold start) still spreads over all SMs.
//   4. apply: cluster-size EMA + Laplace smoothing, ema_w EMA, weight = ema_w / size, and refresh of the
//      bf16 operand + |c|^2 used by the next assign.
// Row reads/writes are 16-byte vectors, coalesced along D.
#include <cub/device/device_radix_sort.cuh>
#include <cuda_bf16.h>
#include <math_constants.h>
#include "../../include/pero_b200.h"
#include "layout.h"

namespace pero {

constexpr int kChunk = 32;   // sorted positions per chunk

struct EmaWsLayout {
    size_t keys_in, keys_out, vals_in, vals_out, seg, partial, scalar, cub, total;
    size_t cub_bytes;
};

inline int key_bits(int64_t K) {
    int b = 1;
    while ((1ll << b) < K) ++b;
    return b;
}

inline EmaWsLayout ema_ws_layout(int64_t N, int64_t K, int64_t D) {
    EmaWsLayout l;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    l.keys_in = take((size_t)N * 4);
    l.keys_out = take((size_t)N * 4);
    l.vals_in = take((size_t)N * 4);
    l.vals_out = take((size_t)N * 4);
    l.seg = take((size_t)(K + 1) * 4);
    const int64_t chunks = (N + kChunk - 1) / kChunk;
    l.partial = take((size_t)chunks * 2 * D * 4);
    l.scalar = take(256);
    size_t cub_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                    (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)N, 0, key_bits(K));
    if (e != cudaSuccess || cub_bytes == 0) {   // no device to query (CPU-only host): conservative bound
        (void)cudaGetLastError();
        cub_bytes = (size_t)N * 16 + (1u << 20);
    }
    l.cub_bytes = cub_bytes;
    l.cub = take(cub_bytes);
    l.total = off;
    return l;
}

__global__ void ema_keys_kernel(const long long* __restrict__ idx, long long N, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ vals) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    keys[i] = (uint32_t)idx[i];
    vals[i] = (uint32_t)i;
}

// seg[k] = first sorted position whose key >= k; seg[K] = N.
__global__ void ema_segments_kernel(const uint32_t* __restrict__ keys, int N, int K, int* __restrict__ seg) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > K) return;
    int lo = 0, hi = N;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < (uint32_t)k) lo = mid + 1; else hi = mid;
    }
    seg[k] = lo;
}

// One CTA per chunk; thread t owns float4 column groups t, t + blockDim, ...
template <int VEC>
__global__ void __launch_bounds__(128)
ema_chunk_sum_kernel(const float* __restrict__ xr, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rows,
                     const int* __restrict__ seg, int N, int D, float* __restrict__ sums, float* __restrict__ partial) {
    __shared__ uint32_t s_key[kChunk], s_row[kChunk];
    const int c = blockIdx.x;
    const int pos0 = c * kChunk, pos1 = min(N, pos0 + kChunk), len = pos1 - pos0;
    if (threadIdx.x < len) { s_key[threadIdx.x] = keys[pos0 + threadIdx.x]; s_row[threadIdx.x] = rows[pos0 + threadIdx.x]; }
    __syncthreads();
    const int groups = D / VEC;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        uint32_t cur = s_key[0];
        auto flush = [&](uint32_t k) {
            const int s0 = __ldg(seg + k), s1 = __ldg(seg + k + 1);
            float* dst;
            if (s0 >= pos0 && s1 <= pos1) dst = sums + (size_t)k * D;                 // whole run in this chunk
            else if (s0 < pos0) dst = partial + ((size_t)c * 2 + 0) * D;                // head: began earlier
            else dst = partial + ((size_t)c * 2 + 1) * D;                               // tail: continues later
            if constexpr (VEC == 4) reinterpret_cast<float4*>(dst)[g] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            else dst[g] = acc[0];
        };
        for (int p = 0; p < len; ++p) {
            const uint32_t k = s_key[p];
            if (k != cur) {
                flush(cur);
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
                cur = k;
            }
            const float* src = xr + (size_t)s_row[p] * D;
            if constexpr (VEC == 4) {
                const float4 x = __ldg(reinterpret_cast<const float4*>(src) + g);
                acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
            } else {
                acc[0] += __ldg(src + g);
            }
        }
        flush(cur);
    }
}

// One warp per codeword: counts, zero rows for unused codewords, and the ordered sum of the chunk
// partials of runs that crossed chunk borders.
template <int VEC>
__global__ void __launch_bounds__(256)
ema_finalize_kernel(const int* __restrict__ seg, int K, int D, const float* __restrict__ partial,
                    float* __restrict__ sums, float* __restrict__ counts) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= K) return;
    const int s0 = seg[k], s1 = seg[k + 1];
    if (lane == 0) counts[k] = (float)(s1 - s0);
    const int groups = D / VEC;
    float* dst = sums + (size_t)k * D;
    if (s1 == s0) {
        for (int g = lane; g < groups; g += 32) {
            if constexpr (VEC == 4) reinterpret_cast<float4*>(dst)[g] = make_float4(0.f, 0.f, 0.f, 0.f);
            else dst[g] = 0.f;
        }
        return;
    }
    const int cf = s0 / kChunk, cl = (s1 - 1) / kChunk;
    if (cf == cl) return;   // summed and stored by the chunk kernel
    for (int g = lane; g < groups; g += 32) {
        float acc[VEC];
        const float* p0 = partial + ((size_t)cf * 2 + 1) * D;
        if constexpr (VEC == 4) {
            const float4 x = reinterpret_cast<const float4*>(p0)[g];
            acc[0] = x.x; acc[1] = x.y; acc[2] = x.z; acc[3] = x.w;
        } else {
            acc[0] = p0[g];
        }
        for (int c = cf + 1; c <= cl; ++c) {
            const float* pc = partial + ((size_t)c * 2 + 0) * D;
            if constexpr (VEC == 4) {
                const float4 x = reinterpret_cast<const float4*>(pc)[g];
                acc[0] += x.x; acc[1] += x.y; acc[2] += x.z; acc[3] += x.w;
            } else {
                acc[0] += pc[g];
            }
        }
        if constexpr (VEC == 4) reinterpret_cast<float4*>(dst)[g] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else dst[g] = acc[0];
    }
}

// cluster size: cs <- cs*decay + (1-decay)*counts; n = sum(cs); cs <- (cs + eps) / (n + K*eps) * n.
// Single CTA, fixed-order tree: K <= 2^20 values.
__global__ void __launch_bounds__(1024)
ema_cluster_size_kernel(const float* __restrict__ counts, int K, float decay, float one_minus_decay, float eps,
                        float k_eps, float* __restrict__ cs) {
    __shared__ float sh[32];
    __shared__ float s_n;
    float part = 0.f;
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        const float v = __fadd_rn(__fmul_rn(cs[k], decay), __fmul_rn(one_minus_decay, counts[k]));   // no FMA: torch rounds each op
        cs[k] = v;
        part += v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = sh[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) s_n = t;
    }
    __syncthreads();
    const float n = s_n;
    const float denom = n + k_eps;
    for (int k = threadIdx.x; k < K; k += blockDim.x) cs[k] = __fmul_rn(__fdiv_rn(__fadd_rn(cs[k], eps), denom), n);
}

// One warp per codeword: ema_w and weight update plus the refreshed GEMM operand and |c|^2.
__global__ void __launch_bounds__(256)
ema_apply_rows_kernel(const float* __restrict__ sums, const float* __restrict__ cs, int K, int D, int Dp, int Kp, float decay,
                      float one_minus_decay, float* __restrict__ ema_w, float* __restrict__ weight, __nv_bfloat16* __restrict__ cb,
                      float* __restrict__ cnorm) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= Kp) return;
    if (k >= K) { if (cnorm && lane == 0) cnorm[k] = CUDART_INF_F; return; }
    const float size = cs[k];
    float s = 0.f;
    for (int d = lane; d < Dp; d += 32) {
        float wv = 0.f;
        if (d < D) {
            const size_t o = (size_t)k * D + d;
            const float e = __fadd_rn(__fmul_rn(ema_w[o], decay), __fmul_rn(one_minus_decay, sums[o]));
            ema_w[o] = e;
            wv = __fdiv_rn(e, size);
            weight[o] = wv;
        }
        s = fmaf(wv, wv, s);
        if (cb) cb[(size_t)k * Dp + d] = __float2bfloat16_rn(wv);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (cnorm && lane == 0) cnorm[k] = s;
}

}  // namespace pero

using namespace pero;

extern "C" {

size_t pero_vq_ema_workspace_bytes(int64_t N, int64_t K, int64_t D) {
    if (N <= 0 || K <= 0 || D <= 0) return 0;
    return ema_ws_layout(N, K, D).total;
}

int pero_vq_ema_accumulate(const float* x_rows, const int64_t* idx, int64_t N, int64_t K, int64_t D,
                           float* sums_counts, void* workspace, size_t workspace_bytes, pero_stream_t stream) {
    if (!x_rows || !idx || !sums_counts || !workspace) return PERO_ERR_NULL;
    if (N <= 0 || K <= 0 || D <= 0 || N > (1ll << 31) - 64 || K > (1ll << 30)) return PERO_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return PERO_ERR_BAD_ALIGN;
    const EmaWsLayout l = ema_ws_layout(N, K, D);
    if (workspace_bytes < l.total) return PERO_ERR_WORKSPACE;
    char* ws = static_cast<char*>(workspace);
    uint32_t* keys_in = reinterpret_cast<uint32_t*>(ws + l.keys_in);
    uint32_t* keys_out = reinterpret_cast<uint32_t*>(ws + l.keys_out);
    uint32_t* vals_in = reinterpret_cast<uint32_t*>(ws + l.vals_in);
    uint32_t* vals_out = reinterpret_cast<uint32_t*>(ws + l.vals_out);
    int* seg = reinterpret_cast<int*>(ws + l.seg);
    float* partial = reinterpret_cast<float*>(ws + l.partial);
    float* sums = sums_counts;
    float* counts = sums_counts + (size_t)K * D;

    ema_keys_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const long long*>(idx), N, keys_in, vals_in);
    size_t cub_bytes = l.cub_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(ws + l.cub, cub_bytes, keys_in, keys_out, vals_in, vals_out, (int)N, 0,
                                                    key_bits(K), stream);
    if (e != cudaSuccess) return (int)e;
    ema_segments_kernel<<<(unsigned)((K + 1 + 255) / 256), 256, 0, stream>>>(keys_out, (int)N, (int)K, seg);
    const unsigned chunks = (unsigned)((N + kChunk - 1) / kChunk);
    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(x_rows) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(sums_counts) & 15) == 0);
    if (vec) {
        const int threads = (int)std::min<int64_t>(128, std::max<int64_t>(32, round_up(D / 4, 32)));
        ema_chunk_sum_kernel<4><<<chunks, threads, 0, stream>>>(x_rows, keys_out, vals_out, seg, (int)N, (int)D, sums, partial);
        ema_finalize_kernel<4><<<(unsigned)((K + 7) / 8), 256, 0, stream>>>(seg, (int)K, (int)D, partial, sums, counts);
    } else {
        ema_chunk_sum_kernel<1><<<chunks, 128, 0, stream>>>(x_rows, keys_out, vals_out, seg, (int)N, (int)D, sums, partial);
        ema_finalize_kernel<1><<<(unsigned)((K + 7) / 8), 256, 0, stream>>>(seg, (int)K, (int)D, partial, sums, counts);
    }
    return (int)cudaGetLastError();
}

int pero_vq_ema_apply(const float* sums_counts, int64_t K, int64_t D, double decay, double epsilon, float* ema_w,
                      float* ema_cluster_size, float* weight, void* codebook, size_t codebook_bytes,
                      void* workspace, size_t workspace_bytes, pero_stream_t stream) {
    (void)workspace; (void)workspace_bytes;
    if (!sums_counts || !ema_w || !ema_cluster_size || !weight) return PERO_ERR_NULL;
    if (K <= 0 || D <= 0 || K > (1ll << 20)) return PERO_ERR_BAD_SHAPE;
    const CodebookLayout cl = codebook_layout(K, D);
    __nv_bfloat16* cb = nullptr;
    float* cnorm = nullptr;
    if (codebook) {
        if (codebook_bytes < cl.total) return PERO_ERR_WORKSPACE;
        if (reinterpret_cast<uintptr_t>(codebook) & 255) return PERO_ERR_BAD_ALIGN;
        cb = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(codebook) + cl.cb_off);
        cnorm = reinterpret_cast<float*>(static_cast<char*>(codebook) + cl.cnorm_off);
    }
    const float* sums = sums_counts;
    const float* counts = sums_counts + (size_t)K * D;
    // Python scalars of the reference are doubles that torch casts to fp32 per operand:
    // decay, (1 - decay), epsilon and K * epsilon are each rounded once, here on the host.
    const float decay_f = (float)decay, omd_f = (float)(1.0 - decay), eps_f = (float)epsilon,
                keps_f = (float)((double)K * epsilon);
    ema_cluster_size_kernel<<<1, 1024, 0, stream>>>(counts, (int)K, decay_f, omd_f, eps_f, keps_f, ema_cluster_size);
    const int rows = (int)(codebook ? cl.Kp : K);
    ema_apply_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(sums, ema_cluster_size, (int)K, (int)D, (int)cl.Dp,
                                                                         rows, decay_f, omd_f, ema_w, weight, cb, cnorm);
    return (int)cudaGetLastError();
}

}  // extern "C"
