// Peer-memory collectives of the two exchange steps of the path, over NVLink 5 / NVSwitch:
//   * SUM all-reduce of fp32 ranges (EMA sums|counts, head gradients d_W|d_b|loss) — batch-sharded mode,
//   * MIN all-reduce of the packed (distance, index) int64 winners — codebook-sharded mode.
// Every rank owns one "peer buffer" of identical size that all ranks have mapped (peer pointers) and, when
// the NVSwitch multicast object exists, a multicast alias of it.  A collective is ONE kernel per rank:
//
//   barrier (all ranks' producers are done)  ->  rank r reduces slice r  ->  writes it to every rank
//   -> barrier (every slice has landed everywhere).
//
// Slice reduction: `multimem.ld_reduce` (the switch adds the replicas, the reducing GPU receives one copy)
// followed by `multimem.st` (the switch replicates the store) when a multicast alias is given; otherwise
// plain peer loads summed in rank order 0..g-1 and peer stores.  Either way each element is reduced exactly
// once, by its owner, and the same bits are delivered to every rank, so replicated state stays bit-identical.
//
// Barriers are one-way release stores of a growing epoch into arrival words in the first
// PERO_PEER_HEADER_BYTES of each peer's buffer ([block][source rank] u32) polled locally with acquire loads;
// nothing is ever reset, so the kernels are CUDA-graph replayable.
// A rank that never arrives makes the waiters trap after 30 s instead of hanging the GPU.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/pero_b200.h"
#include "knobs.h"

namespace {

constexpr int kMaxThreads = 512;
constexpr int kMaxWorld = PERO_PEER_MAX_WORLD;
constexpr int kMaxBlocks = PERO_PEER_MAX_BLOCKS;
// Barrier timeout.  Ranks may reach an exchange minutes apart (a checkpoint or an evaluation on one rank, a data-loader
// stall, lazy module loading), so the default matches NCCL's watchdog: 10 minutes.  The caller may store another value
// (milliseconds, u32) in the TIMEOUT word of its own buffer header.  On a timeout the kernel does NOT trap (a trap
// is a sticky error that kills the CUDA context of every waiting rank): it stores a non-zero code in the local ERROR
// word and returns without touching the payload again; the host reads that word (peer.PeerBuffer.check()).
constexpr unsigned long long kDefaultTimeoutMs = 600000ull;
constexpr int kTimeoutOffsetWords = PERO_PEER_TIMEOUT_OFFSET / 4;
constexpr int kErrorOffsetWords = PERO_PEER_ERROR_OFFSET / 4;
static_assert(kMaxWorld * kMaxBlocks * 4 <= 8192 && 8192 + kMaxBlocks * 4 <= PERO_PEER_TIMEOUT_OFFSET &&
              PERO_PEER_ERROR_OFFSET + 4 <= PERO_PEER_HEADER_BYTES, "flag words must fit the buffer header");

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Header of a peer buffer: arrival words [block][source rank] (written by the peers, one writer each) and one
// epoch word per block (local).  Every collective launch advances a block's epoch by 2 (two barriers); arrival
// words only ever grow, so nothing needs resetting and a CUDA graph can replay the launch.
constexpr int kEpochOffsetWords = 2048;     // byte 8192

__device__ __forceinline__ uint32_t* arrival_word(void* base, int blk, int src) {
    return reinterpret_cast<uint32_t*>(base) + blk * kMaxWorld + src;
}
__device__ __forceinline__ uint32_t* epoch_word(void* base, int blk) {
    return reinterpret_cast<uint32_t*>(base) + kEpochOffsetWords + blk;
}

__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All threads of block `blk` on every rank meet here.  Thread t < world publishes `target` in rank t's arrival
// word [blk][rank] with a one-way store and polls its own word [blk][t] (local memory, relaxed loads) until rank t
// has published the same epoch.
//   kPublish: the block's earlier peer / multicast stores must be performed before the flag (release store: the
//             closing barrier).  The opening barrier publishes nothing of its own — what the peers are about to
//             read was written by earlier kernels of this stream and is already in this GPU's L2 — so a relaxed
//             store suffices there.
//   kConsume: the block goes on to read peers' data (opening barrier): one acquire load after the poll.
// Returns false (for every thread of the block) when a peer did not arrive in time.
template <bool kPublish, bool kConsume>
__device__ __forceinline__ bool rank_barrier(void* const* bufs, int rank, int world, int blk, uint32_t target) {
    __shared__ int timed_out;
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    const int t = threadIdx.x;
    if (t < world) {
        uint32_t* theirs = arrival_word(bufs[t], blk, rank);
        if (kPublish) st_release_sys(theirs, target); else st_relaxed_sys(theirs, target);
        const uint32_t* mine = arrival_word(bufs[rank], blk, t);
        const uint32_t* hdr = reinterpret_cast<const uint32_t*>(bufs[rank]);
        const unsigned long long ms = hdr[kTimeoutOffsetWords] ? hdr[kTimeoutOffsetWords] : kDefaultTimeoutMs;
        const unsigned long long t0 = now_ns();
        unsigned spins = 0;
        while ((int32_t)(ld_relaxed_sys(mine) - target) < 0) {
            if ((++spins & 1023u) == 0 && now_ns() - t0 > ms * 1000000ull) {
                reinterpret_cast<uint32_t*>(bufs[rank])[kErrorOffsetWords] = 0x80000000u | ((uint32_t)blk << 8) | (uint32_t)t;
                atomicExch(&timed_out, 1);
                break;
            }
        }
        if (kConsume) (void)ld_acquire_sys(mine);
    }
    __syncthreads();
    return timed_out == 0;
}

// Epoch of this launch for the block (read before the first barrier, advanced after the second).
__device__ __forceinline__ uint32_t begin_collective(void* const* bufs, int rank, int blk) {
    return *epoch_word(bufs[rank], blk);
}
__device__ __forceinline__ void end_collective(void* const* bufs, int rank, int blk, uint32_t epoch) {
    if (threadIdx.x == 0) *epoch_word(bufs[rank], blk) = epoch + 2;
}

__device__ __forceinline__ float4 mc_ld_reduce_add(const char* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(char* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ long long mc_ld_reduce_min(const char* p) {
    long long v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.min.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(char* p, long long v) {
    asm volatile("multimem.st.relaxed.sys.global.b64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float4 peer_ld(const float4* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st(float4* p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ longlong2 peer_ld(const longlong2* p) {
    longlong2 v;
    asm volatile("ld.relaxed.sys.global.v2.s64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void peer_st(longlong2* p, longlong2 v) {
    asm volatile("st.relaxed.sys.global.v2.s64 [%0], {%1, %2};" :: "l"(p), "l"(v.x), "l"(v.y) : "memory");
}

struct SumF32 {
    using Vec = float4;
    static __device__ __forceinline__ Vec combine(Vec a, Vec b) {
        return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
    }
};
struct MinI64 {
    using Vec = longlong2;
    static __device__ __forceinline__ Vec combine(Vec a, Vec b) {
        return make_longlong2(a.x < b.x ? a.x : b.x, a.y < b.y ? a.y : b.y);
    }
};

struct Range { int64_t lo, hi; };
// Contiguous slice of `total` 16-byte vectors owned by `rank` (remainder spread over the first ranks).
__device__ __forceinline__ Range slice_of(int64_t total, int world, int rank) {
    const int64_t base = total / world, rem = total % world;
    Range r;
    r.lo = rank * base + (rank < rem ? rank : rem);
    r.hi = r.lo + base + (rank < rem ? 1 : 0);
    return r;
}

// Peer-pointer variant.  kWorld > 0 keeps all kWorld loads of a vector in flight at once (registers);
// kWorld == 0 is the generic loop.  With kEmulate the rank is blockIdx.y: g "ranks" of ONE cooperative launch
// on one GPU (the single-GPU test of the protocol, never the product path).
template <class Op, int kWorld, bool kEmulate>
__global__ void __launch_bounds__(kMaxThreads) peer_allreduce_kernel(void* const* __restrict__ bufs, int rank_arg, int world_arg,
                                                                  int64_t off_bytes, int64_t nvec) {
    using Vec = typename Op::Vec;
    const int world = kWorld > 0 ? kWorld : world_arg;
    const int rank = kEmulate ? (int)blockIdx.y : rank_arg;
    const uint32_t epoch = begin_collective(bufs, rank, blockIdx.x);
    if (!rank_barrier<false, true>(bufs, rank, world, blockIdx.x, epoch + 1)) return;
    const Range r = slice_of(nvec, world, rank);
    constexpr int U = kWorld == 2 ? 4 : (kWorld == 8 ? 1 : 2);
    const int kThreads = blockDim.x;
    const int64_t step = (int64_t)gridDim.x * kThreads * U;
    for (int64_t base = r.lo + (int64_t)blockIdx.x * kThreads * U; base < r.hi; base += step) {
        Vec acc[U];
        if (kWorld > 0) {
            Vec v[U][kWorld > 0 ? kWorld : 1];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + u * kThreads + threadIdx.x;
                if (i < r.hi) {
#pragma unroll
                    for (int p = 0; p < kWorld; ++p)
                        v[u][p] = peer_ld(reinterpret_cast<const Vec*>(static_cast<const char*>(bufs[p]) + off_bytes) + i);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                acc[u] = v[u][0];
#pragma unroll
                for (int p = 1; p < kWorld; ++p) acc[u] = Op::combine(acc[u], v[u][p]);
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t i = base + u * kThreads + threadIdx.x;
                if (i < r.hi) {
                    acc[u] = peer_ld(reinterpret_cast<const Vec*>(static_cast<const char*>(bufs[0]) + off_bytes) + i);
                    for (int p = 1; p < world; ++p)
                        acc[u] = Op::combine(acc[u], peer_ld(reinterpret_cast<const Vec*>(static_cast<const char*>(bufs[p]) + off_bytes) + i));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kThreads + threadIdx.x;
            if (i < r.hi) {
                for (int p = 0; p < world; ++p)
                    peer_st(reinterpret_cast<Vec*>(static_cast<char*>(bufs[p]) + off_bytes) + i, acc[u]);
            }
        }
    }
    if (!rank_barrier<true, false>(bufs, rank, world, blockIdx.x, epoch + 2)) return;
    end_collective(bufs, rank, blockIdx.x, epoch);
}

// Multicast variant: the switch reduces on load and replicates on store.
__global__ void __launch_bounds__(kMaxThreads) mc_allreduce_sum_f32_kernel(void* const* __restrict__ bufs, char* __restrict__ mc, int rank,
                                                                        int world, int64_t off_bytes, int64_t nvec) {
    const uint32_t epoch = begin_collective(bufs, rank, blockIdx.x);
    if (!rank_barrier<false, true>(bufs, rank, world, blockIdx.x, epoch + 1)) return;
    const Range r = slice_of(nvec, world, rank);
    constexpr int U = 8;
    const int kThreads = blockDim.x;
    char* base_ptr = mc + off_bytes;
    const int64_t step = (int64_t)gridDim.x * kThreads * U;
    for (int64_t base = r.lo + (int64_t)blockIdx.x * kThreads * U; base < r.hi; base += step) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kThreads + threadIdx.x;
            if (i < r.hi) v[u] = mc_ld_reduce_add(base_ptr + i * 16);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kThreads + threadIdx.x;
            if (i < r.hi) mc_st(base_ptr + i * 16, v[u]);
        }
    }
    if (!rank_barrier<true, false>(bufs, rank, world, blockIdx.x, epoch + 2)) return;
    end_collective(bufs, rank, blockIdx.x, epoch);
}

__global__ void __launch_bounds__(kMaxThreads) mc_allreduce_min_i64_kernel(void* const* __restrict__ bufs, char* __restrict__ mc, int rank,
                                                                        int world, int64_t off_bytes, int64_t n) {
    const uint32_t epoch = begin_collective(bufs, rank, blockIdx.x);
    if (!rank_barrier<false, true>(bufs, rank, world, blockIdx.x, epoch + 1)) return;
    const Range r = slice_of(n, world, rank);
    constexpr int U = 8;
    const int kThreads = blockDim.x;
    char* base_ptr = mc + off_bytes;
    const int64_t step = (int64_t)gridDim.x * kThreads * U;
    for (int64_t base = r.lo + (int64_t)blockIdx.x * kThreads * U; base < r.hi; base += step) {
        long long v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kThreads + threadIdx.x;
            if (i < r.hi) v[u] = mc_ld_reduce_min(base_ptr + i * 8);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = base + u * kThreads + threadIdx.x;
            if (i < r.hi) mc_st(base_ptr + i * 8, v[u]);
        }
    }
    if (!rank_barrier<true, false>(bufs, rank, world, blockIdx.x, epoch + 2)) return;
    end_collective(bufs, rank, blockIdx.x, epoch);
}

int check_args(void* const* peer_bufs, int rank, int world, int64_t off_bytes, int64_t n, int64_t per_vec, int n_blocks) {
    if (!peer_bufs) return PERO_ERR_NULL;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || n < 0) return PERO_ERR_BAD_SHAPE;
    if (n_blocks < 1 || n_blocks > kMaxBlocks) return PERO_ERR_BAD_SHAPE;
    if (off_bytes < PERO_PEER_HEADER_BYTES || (off_bytes & 15) || (n % per_vec)) return PERO_ERR_BAD_ALIGN;
    return PERO_OK;
}

// CTA size of the exchange kernels.  Small CTAs (few registers, no shared memory) slot in beside the resident
// GEMM CTAs instead of waiting for a whole SM; PERO_PEER_THREADS is a tuning knob.
int peer_threads() {
    int t = PERO_KNOB("PERO_PEER_THREADS", 256);      // dev build only
    if (t < 32) t = 32;
    if (t > kMaxThreads) t = kMaxThreads;
    return t / 32 * 32;
}

template <class Op, bool kEmulate>
cudaError_t launch_peer(void* const* bufs, int rank, int world, int64_t off, int64_t nvec, int n_blocks, cudaStream_t s) {
    dim3 grid(n_blocks, kEmulate ? world : 1);
    void* args[] = {(void*)&bufs, (void*)&rank, (void*)&world, (void*)&off, (void*)&nvec};
    const void* fn;
    switch (world) {
        case 2: fn = (const void*)peer_allreduce_kernel<Op, 2, kEmulate>; break;
        case 4: fn = (const void*)peer_allreduce_kernel<Op, 4, kEmulate>; break;
        case 8: fn = (const void*)peer_allreduce_kernel<Op, 8, kEmulate>; break;
        default: fn = (const void*)peer_allreduce_kernel<Op, 0, kEmulate>; break;
    }
    // The emulated ranks spin on one another inside one grid: a cooperative launch guarantees co-residency.
    return kEmulate ? cudaLaunchCooperativeKernel(fn, grid, dim3(peer_threads()), args, 0, s)
                    : cudaLaunchKernel(fn, grid, dim3(peer_threads()), args, 0, s);
}

}  // namespace

extern "C" {

int pero_peer_allreduce_sum_f32(void* const* peer_bufs, void* multicast_base, int rank, int world, int64_t offset_bytes,
                                int64_t n_elems, int n_blocks, pero_stream_t stream) {
    int rc = check_args(peer_bufs, rank, world, offset_bytes, n_elems, 4, n_blocks);
    if (rc != PERO_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (world == 1 || n_elems == 0) return PERO_OK;
    const int64_t nvec = n_elems / 4;
    if (multicast_base) {
        mc_allreduce_sum_f32_kernel<<<n_blocks, peer_threads(), 0, s>>>(peer_bufs, static_cast<char*>(multicast_base), rank, world,
                                                                 offset_bytes, nvec);
        return (int)cudaGetLastError();
    }
    return (int)launch_peer<SumF32, false>(peer_bufs, rank, world, offset_bytes, nvec, n_blocks, s);
}

int pero_peer_allreduce_min_i64(void* const* peer_bufs, void* multicast_base, int rank, int world, int64_t offset_bytes,
                                int64_t n_elems, int n_blocks, pero_stream_t stream) {
    int rc = check_args(peer_bufs, rank, world, offset_bytes, n_elems, 2, n_blocks);
    if (rc != PERO_OK) return rc;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (world == 1 || n_elems == 0) return PERO_OK;
    if (multicast_base) {
        mc_allreduce_min_i64_kernel<<<n_blocks, peer_threads(), 0, s>>>(peer_bufs, static_cast<char*>(multicast_base), rank, world,
                                                                 offset_bytes, n_elems);
        return (int)cudaGetLastError();
    }
    return (int)launch_peer<MinI64, false>(peer_bufs, rank, world, offset_bytes, n_elems / 2, n_blocks, s);
}

int pero_peer_allreduce_emulate(void* const* bufs_on_one_device, int world, int op, int64_t offset_bytes, int64_t n_elems,
                                int n_blocks, pero_stream_t stream) {
    int rc = check_args(bufs_on_one_device, 0, world, offset_bytes, n_elems, op == 0 ? 4 : 2, n_blocks);
    if (rc != PERO_OK) return rc;
    if (op != 0 && op != 1) return PERO_ERR_UNSUPPORTED;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (n_elems == 0) return PERO_OK;
    if (op == 0) return (int)launch_peer<SumF32, true>(bufs_on_one_device, 0, world, offset_bytes, n_elems / 4, n_blocks, s);
    return (int)launch_peer<MinI64, true>(bufs_on_one_device, 0, world, offset_bytes, n_elems / 2, n_blocks, s);
}

}  // extern "C"
