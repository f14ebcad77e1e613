// EMA codebook update (models/autoencoders.py:225-237), deterministic:
//   1. sort (codeword, frame) pairs by codeword, frames ascending inside a codeword.  Up to 8192 frames
//      one CTA does it in a single launch with a stable block radix sort (and emits the segment
//      boundaries); larger batches use CUB's device radix sort (stable as well).
//   2. chunked segmented sum of the fp32 frame rows in sorted order: every 16 sorted positions form one
//      chunk; runs that lie inside a chunk are summed and stored directly, runs that cross chunk borders
//      leave a head/tail partial that a second kernel adds up in chunk order.  No atomics, so the sums
//      are bit-identical run to run, and a collapsed codebook (one codeword owning every frame, as in
//      the reference's cold start) still spreads over all SMs.
//   3. apply: cluster-size EMA + Laplace smoothing, ema_w EMA, weight = ema_w / size, and refresh of the
//      bf16 operand + |c|^2 used by the next assign.
// Row reads/writes are 16-byte vectors, coalesced along D.
#include <stdlib.h>
#include <cub/block/block_radix_sort.cuh>
#include <cub/device/device_radix_sort.cuh>
#include <cuda_bf16.h>
#include <math_constants.h>
#include "../../include/pero_b200.h"
#include "layout.h"
#include "knobs.h"
#include <atomic>

namespace pero {

constexpr int kChunk = 16;            // sorted positions per chunk
constexpr int kSmallSortMax = 8192;   // frames handled by the single-CTA sort (8 per thread x 1024 threads)
constexpr int kClusterBlocks = 64;    // partial sums of the cluster-size reduction

struct EmaWsLayout {
    size_t keys_in, keys_out, vals_in, vals_out, seg, partial, cluster_partial, cub, total;
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
    l.cluster_partial = take(kClusterBlocks * 4);
    size_t cub_bytes = 0;
    if (N > kSmallSortMax) {
        cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                                        (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)N, 0, key_bits(K));
        if (e != cudaSuccess || cub_bytes == 0) {   // no device to query (CPU-only host): conservative bound
            (void)cudaGetLastError();
            cub_bytes = (size_t)N * 16 + (1u << 20);
        }
    }
    l.cub_bytes = cub_bytes;
    l.cub = take(cub_bytes);
    l.total = off;
    return l;
}

// ------------------------------------------------------------------------------------------------ sort
// Single CTA: stable block radix sort (cub::BlockRadixSort) of (codeword, frame) pairs held 8 per
// thread in blocked order -- frames therefore stay ascending inside a codeword -- then the keys, the frame
// ids and the segment table seg[k] = first sorted position with codeword >= k (seg[K] = N).
template <int ITEMS>
__global__ void __launch_bounds__(1024)
ema_sort_small_kernel(const long long* __restrict__ idx, int N, int K, int end_bit, uint32_t* __restrict__ keys_out,
                      uint32_t* __restrict__ vals_out, int* __restrict__ seg) {
    using Sort = cub::BlockRadixSort<uint32_t, 1024, ITEMS, uint32_t>;
    __shared__ typename Sort::TempStorage temp;
    uint32_t keys[ITEMS], vals[ITEMS];
    const uint32_t sentinel = 1u << (end_bit - 1);          // above every codeword: padding sorts last
#pragma unroll
    for (int e = 0; e < ITEMS; ++e) {
        const int i = threadIdx.x * ITEMS + e;
        keys[e] = i < N ? (uint32_t)idx[i] : sentinel;
        vals[e] = (uint32_t)i;
    }
    Sort(temp).Sort(keys, vals, 0, end_bit);
#pragma unroll
    for (int e = 0; e < ITEMS; ++e) {
        const int p = threadIdx.x * ITEMS + e;
        if (p < N) { keys_out[p] = keys[e]; vals_out[p] = vals[e]; }
    }
    __syncthreads();
    // boundaries from the sorted keys just written (same CTA: visible after the barrier)
    for (int p = threadIdx.x; p <= N; p += blockDim.x) {
        const int k = p < N ? (int)keys_out[p] : K;
        const int kprev = p == 0 ? -1 : (int)keys_out[p - 1];
        for (int kk = kprev + 1; kk <= k; ++kk) seg[kk] = p;
    }
}

// Single CTA, N <= 8192 frames, K <= kBinSortMaxK codewords: counting sort in shared memory instead of the radix
// sort above (a third of its time at the bench shape).
//   1. histogram of the codewords (shared-memory atomics: integer counts, order-free)
//   2. exclusive scan -> seg[k]
//   3. every frame takes the next free slot of its codeword's segment (atomic cursor: any order inside a segment)
//   4. the order inside a segment is then made ascending in the frame index, which is all the stable sort was for:
//      segments of 2..32 frames by an insertion sort of their owner thread, longer ones (collapsed codebooks: a few
//      codewords own most frames) by a block-wide ordered compaction of the frames that carry that codeword.
// Output identical to ema_sort_small_kernel: keys_out / vals_out sorted by (codeword, frame), seg[0..K].
constexpr int kBinSortMaxK = 16384;
constexpr int kBinSortThreads = 1024;
constexpr int kBinSortItems = 8;
constexpr int kBinLongMin = 33;

__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    if (w == 0) {
        int ws = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += n;
        }
        warp_sums[lane] = ws;               // inclusive over warps
    }
    __syncthreads();
    total = warp_sums[31];
    const int base = w == 0 ? 0 : warp_sums[w - 1];
    __syncthreads();
    return base + incl - v;
}

__global__ void __launch_bounds__(kBinSortThreads)
ema_bin_small_kernel(const long long* __restrict__ idx, int N, int K, uint32_t* __restrict__ keys_out,
                     uint32_t* __restrict__ vals_out, int* __restrict__ seg) {
    extern __shared__ uint32_t bsm[];
    uint32_t* cursor = bsm;                         // [K]    counts, then next free slot of every segment
    uint32_t* vals = bsm + K;                       // [kBinSortThreads * kBinSortItems]
    uint32_t* start = vals + kBinSortThreads * kBinSortItems;   // [K] first slot of every segment
    __shared__ int warp_sums[32];
    __shared__ int long_list[256];
    __shared__ int n_long;
    const int t = threadIdx.x;
    for (int k = t; k < K; k += kBinSortThreads) cursor[k] = 0u;
    if (t == 0) n_long = 0;
    __syncthreads();
    int key[kBinSortItems];
#pragma unroll
    for (int e = 0; e < kBinSortItems; ++e) {
        const int i = e * kBinSortThreads + t;      // coalesced; frame index ascending in (e, t)
        key[e] = i < N ? (int)idx[i] : -1;
        if (key[e] >= 0) atomicAdd(&cursor[key[e]], 1u);
    }
    __syncthreads();
    // exclusive scan over the K bins: thread t owns bins [t * per, (t + 1) * per)
    const int per = (K + kBinSortThreads - 1) / kBinSortThreads;
    const int k0 = t * per, k1 = min(K, k0 + per);
    int local = 0;
    for (int k = k0; k < k1; ++k) local += (int)cursor[k];
    int total;
    int run = block_exclusive_scan(local, warp_sums, total);
    for (int k = k0; k < k1; ++k) {
        const int c = (int)cursor[k];
        seg[k] = run;
        start[k] = (uint32_t)run;
        cursor[k] = (uint32_t)run;
        if (c >= kBinLongMin) { const int s = atomicAdd(&n_long, 1); long_list[s] = k; }    // at most N / 33 < 256 of them
        run += c;
    }
    if (t == 0) seg[K] = N;
    __syncthreads();
    // placement (order inside a segment: whatever the atomics give; fixed below)
#pragma unroll
    for (int e = 0; e < kBinSortItems; ++e) {
        if (key[e] >= 0) {
            const uint32_t pos = atomicAdd(&cursor[key[e]], 1u);
            vals[pos] = (uint32_t)(e * kBinSortThreads + t);
            keys_out[pos] = (uint32_t)key[e];
        }
    }
    __syncthreads();
    // short segments: insertion sort by the bin's owner thread (cursor[k] is now the END of segment k)
    for (int k = k0; k < k1; ++k) {
        const int s0 = (int)start[k], s1 = (int)cursor[k], c = s1 - s0;
        if (c >= 2 && c < kBinLongMin) {
            for (int a = s0 + 1; a < s1; ++a) {
                const uint32_t v = vals[a];
                int b = a - 1;
                while (b >= s0 && vals[b] > v) { vals[b + 1] = vals[b]; --b; }
                vals[b + 1] = v;
            }
        }
    }
    __syncthreads();
    // long segments: ordered compaction of the frames that carry codeword k, 1024 frames per round
    const int nl = n_long;
    for (int j = 0; j < nl; ++j) {
        const int k = long_list[j];
        int base = (int)start[k];
#pragma unroll 1
        for (int e = 0; e < kBinSortItems; ++e) {
            const int flag = (key[e] == k) ? 1 : 0;
            int round_total;
            const int pre = block_exclusive_scan(flag, warp_sums, round_total);
            if (flag) vals[base + pre] = (uint32_t)(e * kBinSortThreads + t);
            base += round_total;
        }
    }
    __syncthreads();
    for (int p = t; p < N; p += kBinSortThreads) vals_out[p] = vals[p];
}

__global__ void ema_keys_kernel(const long long* __restrict__ idx, long long N, uint32_t* __restrict__ keys,
                                uint32_t* __restrict__ vals) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    keys[i] = (uint32_t)idx[i];
    vals[i] = (uint32_t)i;
}

// seg[k] = first sorted position whose key >= k; seg[K] = N.  One thread per sorted position writes the
// (usually zero or one) boundaries that fall between its predecessor's key and its own.
__global__ void ema_boundaries_kernel(const uint32_t* __restrict__ keys, int N, int K, int* __restrict__ seg) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > N) return;
    const int k = p < N ? (int)keys[p] : K;
    const int kprev = p == 0 ? -1 : (int)keys[p - 1];
    for (int kk = kprev + 1; kk <= k; ++kk) seg[kk] = p;
}

// ------------------------------------------------------------------------------------------------ segmented sum
// One CTA per chunk of 16 sorted positions; thread t owns 16-byte column groups t, t + blockDim, ...
// All 16 row loads of a column group are issued before the running sums are formed.
template <int VEC>
__global__ void __launch_bounds__(128)
ema_chunk_sum_kernel(const float* __restrict__ xr, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ rows,
                     const int* __restrict__ seg, int N, int D, float* __restrict__ sums, float* __restrict__ partial) {
    __shared__ uint32_t s_key[kChunk], s_row[kChunk];
    __shared__ int s_dst[kChunk];      // where the run ending at position p goes: -1 none, 0 sums, 1 head, 2 tail
    const int c = blockIdx.x;
    const int pos0 = c * kChunk, pos1 = min(N, pos0 + kChunk), len = pos1 - pos0;
    if (threadIdx.x < kChunk) {
        const int p = threadIdx.x;
        int dst = -1;
        uint32_t k = 0, r = 0;
        if (p < len) {
            k = keys[pos0 + p]; r = rows[pos0 + p];
            const bool run_end = (p + 1 == len) || (keys[pos0 + p + 1] != k);
            if (run_end) {
                const int s0 = __ldg(seg + k), s1 = __ldg(seg + k + 1);
                dst = (s0 >= pos0 && s1 <= pos1) ? 0 : (s0 < pos0 ? 1 : 2);
            }
        }
        s_key[p] = k; s_row[p] = r; s_dst[p] = dst;
    }
    __syncthreads();
    const int groups = D / VEC;
    for (int g = threadIdx.x; g < groups; g += blockDim.x) {
        float v[kChunk][VEC];
#pragma unroll
        for (int p = 0; p < kChunk; ++p) {
            if (p < len) {
                const float* src = xr + (size_t)s_row[p] * D;
                if constexpr (VEC == 4) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(src) + g);
                    v[p][0] = x.x; v[p][1] = x.y; v[p][2] = x.z; v[p][3] = x.w;
                } else {
                    v[p][0] = __ldg(src + g);
                }
            }
        }
        float acc[VEC];
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
        for (int p = 0; p < kChunk; ++p) {
            if (p < len) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) acc[e] += v[p][e];
                const int dst = s_dst[p];
                if (dst >= 0) {
                    float* out = dst == 0 ? sums + (size_t)s_key[p] * D : partial + ((size_t)c * 2 + (dst - 1)) * D;
                    if constexpr (VEC == 4) reinterpret_cast<float4*>(out)[g] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    else out[g] = acc[0];
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
                }
            }
        }
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

// ------------------------------------------------------------------------------------------------ apply
// cs' = cs*decay + (1-decay)*counts (not stored); block partial sums of cs' in a fixed order.
__global__ void __launch_bounds__(256)
ema_cluster_partial_kernel(const float* __restrict__ counts, const float* __restrict__ cs, int K, float decay,
                           float one_minus_decay, float* __restrict__ partial) {
    __shared__ float sh[8];
    float part = 0.f;
    const int per = (K + gridDim.x - 1) / gridDim.x;
    const int k0 = blockIdx.x * per, k1 = min(K, k0 + per);
    for (int k = k0 + threadIdx.x; k < k1; k += blockDim.x)
        part += __fadd_rn(__fmul_rn(cs[k], decay), __fmul_rn(one_minus_decay, counts[k]));   // no FMA: torch rounds each op
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < 8 ? sh[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = t;
    }
}

// One warp per codeword: n = sum of the block partials (same fixed order in every warp),
// cs <- (cs' + eps) / (n + K*eps) * n, ema_w and weight update, refreshed GEMM operand and |c|^2.
__global__ void __launch_bounds__(256)
ema_apply_rows_kernel(const float* __restrict__ sums, const float* __restrict__ counts, const float* __restrict__ cluster_partial,
                      int nblocks, int K, int D, int Dp, int Kp, float decay, float one_minus_decay, float eps, float k_eps,
                      float* __restrict__ cs, float* __restrict__ ema_w, float* __restrict__ weight,
                      __nv_bfloat16* __restrict__ cb, float* __restrict__ cnorm) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= Kp) return;
    if (k >= K) { if (cnorm && lane == 0) cnorm[k] = CUDART_INF_F; return; }
    float n = 0.f;
    for (int b = 0; b < nblocks; ++b) n += __ldg(cluster_partial + b);
    const float csp = __fadd_rn(__fmul_rn(cs[k], decay), __fmul_rn(one_minus_decay, counts[k]));
    const float size = __fmul_rn(__fdiv_rn(__fadd_rn(csp, eps), n + k_eps), n);
    __syncwarp();                     // every lane has read cs[k] before lane 0 overwrites it
    if (lane == 0) cs[k] = size;
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

// Mini-batch k-means centre update (scikit-learn MiniBatchKMeans, _k_means_minibatch.pyx update_center_dense, the
// fitter behind scripts/fit_kmeans.py:20-32): for every centre with members in the batch
//     c <- (c * w + sum of members) * (1 / (w + n)),   w <- w + n
// and untouched otherwise.  One warp per centre; also refreshes the bf16 GEMM operand and |c|^2 for the next assign.
__global__ void __launch_bounds__(256)
kmeans_update_rows_kernel(const float* __restrict__ sums, const float* __restrict__ counts, int K, int D, int Dp, int Kp,
                          float* __restrict__ centers, float* __restrict__ weight_sums, __nv_bfloat16* __restrict__ cb,
                          float* __restrict__ cnorm) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= Kp) return;
    if (k >= K) { if (cnorm && lane == 0) cnorm[k] = CUDART_INF_F; return; }
    const float n = counts[k], w = weight_sums[k];
    const float w_new = __fadd_rn(w, n);
    const float alpha = __fdiv_rn(1.0f, w_new);
    __syncwarp();
    if (lane == 0 && n > 0.f) weight_sums[k] = w_new;
    float s = 0.f;
    for (int d = lane; d < Dp; d += 32) {
        float c = 0.f;
        if (d < D) {
            const size_t o = (size_t)k * D + d;
            c = centers[o];
            if (n > 0.f) {
                c = __fmul_rn(__fadd_rn(__fmul_rn(c, w), sums[o]), alpha);
                centers[o] = c;
            }
        }
        s = fmaf(c, c, s);
        if (cb) cb[(size_t)k * Dp + d] = __float2bfloat16_rn(c);
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

    const int bin_sort = PERO_KNOB("PERO_EMA_BIN_SORT", 1);     // dev build, 0: radix sort also for the small case
    if (bin_sort && N <= kBinSortThreads * kBinSortItems && K <= kBinSortMaxK) {
        const size_t smem = ((size_t)2 * K + kBinSortThreads * kBinSortItems) * 4;
        {
            static std::atomic<bool> attr_done[64];
            int dev = 0;
            cudaGetDevice(&dev);
            dev = (dev >= 0 && dev < 64) ? dev : 0;
            if (!attr_done[dev].load(std::memory_order_acquire)) {
                cudaError_t e = cudaFuncSetAttribute(ema_bin_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)(((size_t)2 * kBinSortMaxK + kBinSortThreads * kBinSortItems) * 4));
                if (e != cudaSuccess) return (int)e;
                attr_done[dev].store(true, std::memory_order_release);
            }
        }
        ema_bin_small_kernel<<<1, kBinSortThreads, smem, stream>>>(reinterpret_cast<const long long*>(idx), (int)N, (int)K, keys_out,
                                                                  vals_out, seg);
    } else if (N <= kSmallSortMax) {
        const int end_bit = key_bits(K) + 1;
        ema_sort_small_kernel<8><<<1, 1024, 0, stream>>>(reinterpret_cast<const long long*>(idx), (int)N, (int)K, end_bit, keys_out,
                                                        vals_out, seg);
    } else {
        ema_keys_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const long long*>(idx), N, keys_in, vals_in);
        size_t cub_bytes = l.cub_bytes;
        cudaError_t e = cub::DeviceRadixSort::SortPairs(ws + l.cub, cub_bytes, keys_in, keys_out, vals_in, vals_out, (int)N, 0,
                                                        key_bits(K), stream);
        if (e != cudaSuccess) return (int)e;
        ema_boundaries_kernel<<<(unsigned)((N + 1 + 255) / 256), 256, 0, stream>>>(keys_out, (int)N, (int)K, seg);
    }
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
    if (!sums_counts || !ema_w || !ema_cluster_size || !weight || !workspace) return PERO_ERR_NULL;
    if (K <= 0 || D <= 0 || K > (1ll << 24)) return PERO_ERR_BAD_SHAPE;
    if (workspace_bytes < align256(kClusterBlocks * 4)) return PERO_ERR_WORKSPACE;
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
    float* cluster_partial = static_cast<float*>(workspace);
    // Python scalars of the reference are doubles that torch casts to fp32 per operand:
    // decay, (1 - decay), epsilon and K * epsilon are each rounded once, here on the host.
    const float decay_f = (float)decay, omd_f = (float)(1.0 - decay), eps_f = (float)epsilon,
                keps_f = (float)((double)K * epsilon);
    const int nblocks = (int)std::min<int64_t>(kClusterBlocks, (K + 255) / 256);
    ema_cluster_partial_kernel<<<nblocks, 256, 0, stream>>>(counts, ema_cluster_size, (int)K, decay_f, omd_f, cluster_partial);
    const int rows = (int)(codebook ? cl.Kp : K);
    ema_apply_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(sums, counts, cluster_partial, nblocks, (int)K, (int)D,
                                                                         (int)cl.Dp, rows, decay_f, omd_f, eps_f, keps_f,
                                                                         ema_cluster_size, ema_w, weight, cb, cnorm);
    return (int)cudaGetLastError();
}

int pero_kmeans_update(const float* sums_counts, int64_t K, int64_t D, float* centers, float* weight_sums, void* codebook,
                       size_t codebook_bytes, pero_stream_t stream) {
    if (!sums_counts || !centers || !weight_sums) return PERO_ERR_NULL;
    if (K <= 0 || D <= 0 || K > (1ll << 24)) return PERO_ERR_BAD_SHAPE;
    const CodebookLayout cl = codebook_layout(K, D);
    __nv_bfloat16* cb = nullptr;
    float* cnorm = nullptr;
    if (codebook) {
        if (codebook_bytes < cl.total) return PERO_ERR_WORKSPACE;
        if (reinterpret_cast<uintptr_t>(codebook) & 255) return PERO_ERR_BAD_ALIGN;
        cb = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(codebook) + cl.cb_off);
        cnorm = reinterpret_cast<float*>(static_cast<char*>(codebook) + cl.cnorm_off);
    }
    const int rows = (int)(codebook ? cl.Kp : K);
    kmeans_update_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(sums_counts, sums_counts + (size_t)K * D, (int)K,
                                                                             (int)D, (int)cl.Dp, rows, centers, weight_sums, cb,
                                                                             cnorm);
    return (int)cudaGetLastError();
}

}  // extern "C"
