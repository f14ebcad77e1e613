// EMA codebook update (models/autoencoders.py:225-237), deterministic:
//   1. sort (codeword, frame) pairs by codeword, frames ascending inside a codeword.  Up to 8192 frames and
//      16384 codewords: a stable multi-CTA counting sort in three short launches (per-CTA stable ranks +
//      histograms, one-CTA scan, placement); otherwise a one-CTA block radix sort (<= 8192 frames) or CUB's
//      device radix sort (stable as well).
//   2. segmented sum of the fp32 frame rows in sorted order, ONE launch: codewords that own up to kLongSeg frames (all
//      of them once the codebook is warm) are summed by one warp each, which also writes the zero rows of unused
//      codewords and the counts; a longer segment (a popular codeword, or the collapsed codebook of the reference's
//      cold start, where one codeword owns every frame) is summed by the 8 warps of its CTA, an eighth each, and the 8
//      partial rows are added in warp order.  No float atomics anywhere: the sums are bit-identical run to run.
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

constexpr int kSmallSortMax = 8192;   // frames handled by the single-CTA sort (8 per thread x 1024 threads)
constexpr int kClusterBlocks = 64;    // partial sums of the cluster-size reduction
constexpr int kLongSeg = 64;          // segments longer than this are summed by the whole CTA
constexpr int kRankThreads = 1024;    // frames per CTA of the counting sort
constexpr int kRankMaxBlocks = kSmallSortMax / kRankThreads;

struct EmaWsLayout {
    size_t keys_in, keys_out, vals_in, vals_out, seg, cluster_partial, hist, excl, btot, cub, total;
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
    l.cluster_partial = take(kClusterBlocks * 4);
    l.hist = take(N <= kSmallSortMax ? (size_t)kRankMaxBlocks * K * 4 : 0);    // counting sort: per-CTA histograms / bases
    l.excl = take(N <= kSmallSortMax ? (size_t)K * 4 : 0);                      // counting sort: scan inside 256-codeword blocks
    l.btot = take(N <= kSmallSortMax ? (size_t)((K + 255) / 256) * 4 : 0);      // ... and the block sums
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

// The counting sort below handles up to kBinSortMaxK codewords (its per-CTA counters live in shared memory).
constexpr int kBinSortMaxK = 16384;

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

// ---- stable multi-CTA counting sort (N <= 8192 frames, K <= kBinSortMaxK codewords) -----------------------------------
// Launch 1, one CTA of 256 threads per 1024 frames (thread t owns frames base + r * 256 + t, r = 0..3): the stable rank
// of every frame among the frames of ITS CTA that carry the same codeword, and the CTA's histogram.  Lanes of a warp
// that share a codeword are found with match.any; the (round, warp) pairs then take turns in ascending frame order,
// the lowest lane of each group reserving `group size` slots of the codeword's 16-bit counter in shared memory: ranks
// ascend with the frame index, no atomics.  256 threads, 32 registers and 2 K bytes of shared memory: the CTA fits
// beside a resident GEMM CTA instead of waiting for a free SM.
constexpr int kRankCtaThreads = 256;
constexpr int kRankRounds = kRankThreads / kRankCtaThreads;
__global__ void __launch_bounds__(kRankCtaThreads)
ema_rank_kernel(const long long* __restrict__ idx, int N, int K, uint32_t* __restrict__ lrank, uint32_t* __restrict__ hist) {
    extern __shared__ unsigned short cnt[];           // [K], at most 1024 per codeword
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    for (int k = t; k < K; k += kRankCtaThreads) cnt[k] = 0;
    int key[kRankRounds], leader[kRankRounds], rank_in_warp[kRankRounds], gsize[kRankRounds];
    uint32_t base[kRankRounds];
#pragma unroll
    for (int r = 0; r < kRankRounds; ++r) {
        const int i = blockIdx.x * kRankThreads + r * kRankCtaThreads + t;
        key[r] = i < N ? (int)idx[i] : -1 - lane;     // padding lanes: distinct negative keys, groups of one
        const unsigned grp = __match_any_sync(0xffffffffu, key[r]);
        leader[r] = __ffs(grp) - 1;
        rank_in_warp[r] = __popc(grp & ((1u << lane) - 1u));
        gsize[r] = __popc(grp);
        base[r] = 0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kRankRounds; ++r) {
        for (int w = 0; w < kRankCtaThreads / 32; ++w) {
            if (warp == w && lane == leader[r] && key[r] >= 0) {
                base[r] = cnt[key[r]];
                cnt[key[r]] = (unsigned short)(base[r] + gsize[r]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int r = 0; r < kRankRounds; ++r) {
        const uint32_t b = __shfl_sync(0xffffffffu, base[r], leader[r]);
        const int i = blockIdx.x * kRankThreads + r * kRankCtaThreads + t;
        if (i < N) lrank[i] = b + (uint32_t)rank_in_warp[r];
    }
    uint32_t* h = hist + (size_t)blockIdx.x * K;
    for (int k = t; k < K; k += kRankCtaThreads) h[k] = cnt[k];
}

// Launch 2, one CTA per 256 codewords: total[k] = frames of codeword k over all rank CTAs, its exclusive scan INSIDE the
// CTA's 256 codewords (excl_local) and the CTA's sum (block_total).  The scan across CTAs (at most 64 values) is
// finished by whoever needs seg[k]: seg[k] = sum(block_total[0 .. k/256)) + excl_local[k].
__global__ void __launch_bounds__(256)
ema_scan_local_kernel(const uint32_t* __restrict__ hist, int nblocks, int K, int* __restrict__ excl_local,
                      int* __restrict__ block_total) {
    __shared__ int wsum[8];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int k = blockIdx.x * 256 + t;
    int total = 0;
    if (k < K)
        for (int b = 0; b < nblocks; ++b) total += (int)hist[(size_t)b * K + k];
    int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    int base = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) if (j < w) base += wsum[j];
    if (k < K) excl_local[k] = base + incl - total;
    if (t == 255) block_total[blockIdx.x] = base + incl;
}

__device__ __forceinline__ int ema_block_offset(const int* __restrict__ block_total, int j) {
    int off = 0;
    for (int i = 0; i < j; ++i) off += __ldg(block_total + i);
    return off;
}

// Launch 3: every frame goes to seg[its codeword] + (frames of that codeword in earlier rank CTAs) + its rank; the same
// threads also publish seg[0 .. K] for the kernels behind.
__global__ void __launch_bounds__(256)
ema_place_kernel(const long long* __restrict__ idx, const uint32_t* __restrict__ lrank, const uint32_t* __restrict__ hist,
                 const int* __restrict__ excl_local, const int* __restrict__ block_total, int N, int K,
                 uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int* __restrict__ seg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) seg[i] = ema_block_offset(block_total, i >> 8) + __ldg(excl_local + i);
    if (i == K) seg[K] = N;
    if (i >= N) return;
    const int key = (int)idx[i];
    int pos = ema_block_offset(block_total, key >> 8) + __ldg(excl_local + key) + (int)lrank[i];
    const int blk = i / kRankThreads;
    for (int b = 0; b < blk; ++b) pos += (int)__ldg(hist + (size_t)b * K + key);
    keys_out[pos] = (uint32_t)key;
    vals_out[pos] = (uint32_t)i;
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
// One CTA per 8 codewords.  Pass 1, one warp per codeword: counts, the zero row of an unused codeword, and -- for
// segments of up to kLongSeg frames, i.e. all of them once the codebook is warm -- the sum of the segment's frame rows in
// ascending frame order (lane <-> 16-byte column groups, 4 row loads in flight).  Pass 2, rare: a longer segment (a
// popular codeword; or the collapsed codebook of the reference's cold start, where one codeword owns every frame) is
// summed by all 8 warps of the CTA, warp w taking the w-th eighth of the segment in ascending order, and the 8
// partial rows are added in warp order: a fixed order again, so the sums stay bit-identical run to run.  No float
// atomics, no second launch.
template <int VEC>
__device__ __forceinline__ void ema_accumulate_rows(const float* __restrict__ xr, const uint32_t* __restrict__ rows, int p0, int p1,
                                                    int D, int ga, int gb, int groups, float (&a)[VEC], float (&b)[VEC]) {
#pragma unroll
    for (int e = 0; e < VEC; ++e) { a[e] = 0.f; b[e] = 0.f; }
    int p = p0;
    for (; p + 2 <= p1; p += 2) {                      // two frames per iteration: four independent loads
        const float* r0 = xr + (size_t)__ldg(rows + p) * D;
        const float* r1 = xr + (size_t)__ldg(rows + p + 1) * D;
        if constexpr (VEC == 4) {
            float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0, y0 = x0, y1 = x0;
            if (ga < groups) { x0 = __ldg(reinterpret_cast<const float4*>(r0) + ga); x1 = __ldg(reinterpret_cast<const float4*>(r1) + ga); }
            if (gb < groups) { y0 = __ldg(reinterpret_cast<const float4*>(r0) + gb); y1 = __ldg(reinterpret_cast<const float4*>(r1) + gb); }
            a[0] += x0.x; a[1] += x0.y; a[2] += x0.z; a[3] += x0.w;
            a[0] += x1.x; a[1] += x1.y; a[2] += x1.z; a[3] += x1.w;
            b[0] += y0.x; b[1] += y0.y; b[2] += y0.z; b[3] += y0.w;
            b[0] += y1.x; b[1] += y1.y; b[2] += y1.z; b[3] += y1.w;
        } else {
            const float x0 = ga < groups ? __ldg(r0 + ga) : 0.f, x1 = ga < groups ? __ldg(r1 + ga) : 0.f;
            const float y0 = gb < groups ? __ldg(r0 + gb) : 0.f, y1 = gb < groups ? __ldg(r1 + gb) : 0.f;
            a[0] += x0; a[0] += x1; b[0] += y0; b[0] += y1;
        }
    }
    if (p < p1) {
        const float* r0 = xr + (size_t)__ldg(rows + p) * D;
        if constexpr (VEC == 4) {
            if (ga < groups) { const float4 x0 = __ldg(reinterpret_cast<const float4*>(r0) + ga); a[0] += x0.x; a[1] += x0.y; a[2] += x0.z; a[3] += x0.w; }
            if (gb < groups) { const float4 y0 = __ldg(reinterpret_cast<const float4*>(r0) + gb); b[0] += y0.x; b[1] += y0.y; b[2] += y0.z; b[3] += y0.w; }
        } else {
            if (ga < groups) a[0] += __ldg(r0 + ga);
            if (gb < groups) b[0] += __ldg(r0 + gb);
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(256)
ema_rowsum_kernel(const float* __restrict__ xr, const uint32_t* __restrict__ rows, const int* __restrict__ seg, int K, int D,
                  float* __restrict__ sums, float* __restrict__ counts) {
    __shared__ float part[8][64 * VEC];               // pass 2: one partial row slab (64 column groups) per warp
    __shared__ int long_k[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + warp;
    const int groups = D / VEC;
    if (threadIdx.x < 8) long_k[threadIdx.x] = -1;
    __syncthreads();
    if (k < K) {
        const int s0 = __ldg(seg + k), s1 = __ldg(seg + k + 1), c = s1 - s0;
        if (lane == 0) counts[k] = (float)c;
        if (c > kLongSeg) {
            if (lane == 0) long_k[warp] = k;
        } else {
            float* dst = sums + (size_t)k * D;
            for (int g0 = 0; g0 < groups; g0 += 64) {          // two column groups per lane and pass
                const int ga = g0 + lane, gb = g0 + 32 + lane;
                float a[VEC], b[VEC];
                ema_accumulate_rows<VEC>(xr, rows, s0, s1, D, ga, gb, groups, a, b);
                if constexpr (VEC == 4) {
                    if (ga < groups) reinterpret_cast<float4*>(dst)[ga] = make_float4(a[0], a[1], a[2], a[3]);
                    if (gb < groups) reinterpret_cast<float4*>(dst)[gb] = make_float4(b[0], b[1], b[2], b[3]);
                } else {
                    if (ga < groups) dst[ga] = a[0];
                    if (gb < groups) dst[gb] = b[0];
                }
            }
        }
    }
    __syncthreads();
    for (int j = 0; j < 8; ++j) {
        const int kk = long_k[j];                              // the same for every thread of the CTA
        if (kk < 0) continue;
        const int s0 = __ldg(seg + kk), s1 = __ldg(seg + kk + 1);
        const int per = (s1 - s0 + 7) / 8;
        const int p0 = min(s1, s0 + warp * per), p1 = min(s1, p0 + per);
        float* dst = sums + (size_t)kk * D;
        for (int g0 = 0; g0 < groups; g0 += 64) {
            const int ga = g0 + lane, gb = g0 + 32 + lane;
            float a[VEC], b[VEC];
            ema_accumulate_rows<VEC>(xr, rows, p0, p1, D, ga, gb, groups, a, b);
#pragma unroll
            for (int e = 0; e < VEC; ++e) { part[warp][lane * VEC + e] = a[e]; part[warp][(32 + lane) * VEC + e] = b[e]; }
            __syncthreads();
            for (int t = threadIdx.x; t < 64 * VEC; t += 256) {          // column t of the slab: the 8 partials in warp order
                const int col = g0 * VEC + t;
                if (col < D) {
                    float acc = part[0][t];
#pragma unroll
                    for (int w = 1; w < 8; ++w) acc += part[w][t];
                    dst[col] = acc;
                }
            }
            __syncthreads();
        }
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

// One warp per codeword: n = sum of the block partials (the same fixed shuffle tree in every warp, so every codeword
// sees the same n), cs <- (cs' + eps) / (n + K*eps) * n, ema_w and weight update, refreshed GEMM operand and |c|^2.
// kVec4 (D % 4 == 0, 16-byte aligned arrays): 16-byte row accesses.
template <bool kVec4>
__global__ void __launch_bounds__(256)
ema_apply_rows_kernel(const float* __restrict__ sums, const float* __restrict__ counts, const float* __restrict__ cluster_partial,
                      int nblocks, int K, int D, int Dp, int Kp, float decay, float one_minus_decay, float eps, float k_eps,
                      float* __restrict__ cs, float* __restrict__ ema_w, float* __restrict__ weight,
                      __nv_bfloat16* __restrict__ cb, float* __restrict__ cnorm) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= Kp) return;
    if (k >= K) { if (cnorm && lane == 0) cnorm[k] = CUDART_INF_F; return; }
    float n = (lane < nblocks ? __ldg(cluster_partial + lane) : 0.f) + (lane + 32 < nblocks ? __ldg(cluster_partial + lane + 32) : 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    const float csp = __fadd_rn(__fmul_rn(cs[k], decay), __fmul_rn(one_minus_decay, counts[k]));
    const float size = __fmul_rn(__fdiv_rn(__fadd_rn(csp, eps), n + k_eps), n);
    __syncwarp();                     // every lane has read cs[k] before lane 0 overwrites it
    if (lane == 0) cs[k] = size;
    float s = 0.f;
    if constexpr (kVec4) {
        const size_t row = (size_t)k * D;
        for (int g = lane; g < Dp / 4; g += 32) {
            float4 wv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (4 * g < D) {
                const float4 e0 = *reinterpret_cast<const float4*>(ema_w + row + 4 * g);
                const float4 sm = __ldg(reinterpret_cast<const float4*>(sums + row + 4 * g));
                float4 e;
                e.x = __fadd_rn(__fmul_rn(e0.x, decay), __fmul_rn(one_minus_decay, sm.x));
                e.y = __fadd_rn(__fmul_rn(e0.y, decay), __fmul_rn(one_minus_decay, sm.y));
                e.z = __fadd_rn(__fmul_rn(e0.z, decay), __fmul_rn(one_minus_decay, sm.z));
                e.w = __fadd_rn(__fmul_rn(e0.w, decay), __fmul_rn(one_minus_decay, sm.w));
                *reinterpret_cast<float4*>(ema_w + row + 4 * g) = e;
                wv = make_float4(__fdiv_rn(e.x, size), __fdiv_rn(e.y, size), __fdiv_rn(e.z, size), __fdiv_rn(e.w, size));
                *reinterpret_cast<float4*>(weight + row + 4 * g) = wv;
            }
            s = fmaf(wv.x, wv.x, s); s = fmaf(wv.y, wv.y, s); s = fmaf(wv.z, wv.z, s); s = fmaf(wv.w, wv.w, s);
            if (cb) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(wv.x, wv.y), hi = __floats2bfloat162_rn(wv.z, wv.w);
                uint2 o;
                o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(cb + (size_t)k * Dp + 4 * g) = o;
            }
        }
    } else {
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
    float* sums = sums_counts;
    float* counts = sums_counts + (size_t)K * D;

    const int bin_sort = PERO_KNOB("PERO_EMA_BIN_SORT", 1);     // dev build, 0: radix sort also for the small case
    if (bin_sort && N <= kSmallSortMax && K <= kBinSortMaxK) {
        // stable counting sort in three short launches (ranks + histograms, scan, placement); the per-frame ranks
        // borrow keys_in, which only the CUB path uses
        uint32_t* hist = reinterpret_cast<uint32_t*>(ws + l.hist);
        uint32_t* lrank = keys_in;
        const int nb = (int)((N + kRankThreads - 1) / kRankThreads);
        {
            static std::atomic<bool> attr_done[64];
            int dev = 0;
            cudaGetDevice(&dev);
            dev = (dev >= 0 && dev < 64) ? dev : 0;
            if (!attr_done[dev].load(std::memory_order_acquire)) {
                cudaError_t e = cudaFuncSetAttribute(ema_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBinSortMaxK * 2);
                if (e != cudaSuccess) return (int)e;
                attr_done[dev].store(true, std::memory_order_release);
            }
        }
        ema_rank_kernel<<<nb, kRankCtaThreads, (size_t)K * 2, stream>>>(reinterpret_cast<const long long*>(idx), (int)N, (int)K, lrank,
                                                                     hist);
        int* excl = reinterpret_cast<int*>(ws + l.excl);
        int* btot = reinterpret_cast<int*>(ws + l.btot);
        ema_scan_local_kernel<<<(unsigned)((K + 255) / 256), 256, 0, stream>>>(hist, nb, (int)K, excl, btot);
        const int64_t place_threads = std::max<int64_t>(N, K + 1);
        ema_place_kernel<<<(unsigned)((place_threads + 255) / 256), 256, 0, stream>>>(reinterpret_cast<const long long*>(idx), lrank, hist,
                                                                                      excl, btot, (int)N, (int)K, keys_out, vals_out, seg);
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
    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(x_rows) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(sums_counts) & 15) == 0);
    const unsigned kblocks = (unsigned)((K + 7) / 8);
    if (vec) ema_rowsum_kernel<4><<<kblocks, 256, 0, stream>>>(x_rows, vals_out, seg, (int)K, (int)D, sums, counts);
    else ema_rowsum_kernel<1><<<kblocks, 256, 0, stream>>>(x_rows, vals_out, seg, (int)K, (int)D, sums, counts);
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
    const bool vec4 = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(sums_counts) | reinterpret_cast<uintptr_t>(ema_w) |
                                        reinterpret_cast<uintptr_t>(weight)) & 15) == 0;
    if (vec4)
        ema_apply_rows_kernel<true><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(sums, counts, cluster_partial, nblocks, (int)K,
                                                                                   (int)D, (int)cl.Dp, rows, decay_f, omd_f, eps_f,
                                                                                   keps_f, ema_cluster_size, ema_w, weight, cb, cnorm);
    else
        ema_apply_rows_kernel<false><<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(sums, counts, cluster_partial, nblocks, (int)K,
                                                                                    (int)D, (int)cl.Dp, rows, decay_f, omd_f, eps_f,
                                                                                    keps_f, ema_cluster_size, ema_w, weight, cb, cnorm);
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
