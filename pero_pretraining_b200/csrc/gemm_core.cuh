// Warp-specialised tcgen05 GEMM core for sm_100a, shared by every dense contraction on the
// quantize-and-predict path (nearest-codeword distances, head logits, dlogits recompute, dW, dh).
//
//   C[i, j] = sum_k A[i, k] * B[j, k]        A: [rows_a, Kd] bf16, B: [rows_b, Kd] bf16, both K-contiguous
//
// so both operands are K-major UMMA operands fed by TMA with SWIZZLE_128B and never transposed.
// One CTA (kCtaGroup == 1) or one CTA pair (kCtaGroup == 2, tcgen05 cta_group::2, UMMA M = 256) owns
// 128 (256) rows of A at a time and sweeps 256-wide column tiles of B.  fp32 accumulators live in TMEM,
// double-buffered across all 512 columns, so the epilogue of tile t overlaps the MMAs of tile t+1.
//
//   warp 0      TMA producer (one elected lane)
//   warp 1      MMA issuer   (one elected lane, leader CTA only)
//   warp 2      TMEM allocator
//   warps 4..7  epilogue: thread <-> accumulator row, tcgen05.ld 32 columns at a time
//
// With kAResident the A row block is loaded ONCE per row block into its own k-block slots and only B
// streams through the stage ring (8192/256 = 32 B/cycle/SM of L2 traffic in pair mode instead of 96);
// otherwise A and B k-blocks share the ring (long contractions: dW, dh).
//
// The epilogue is a policy class (see epilogues in vq_assign.cu / masked_ce.cu):
//   struct Epi { struct Params; struct State;
//     static __device__ void begin_rb(State&, const Params&, const TileCtx&);
//     static __device__ void tile    (State&, const Params&, const TileCtx&, uint32_t taddr);
//     static __device__ void end_rb  (State&, const Params&, const TileCtx&); };
// Row-reducing epilogues (arg-min, log-sum-exp) keep their running state in registers across the
// column sweep, so the rows x columns matrix never reaches HBM.
#pragma once
#include "ptx.cuh"

namespace pero {

constexpr int kBlockM = 128;        // accumulator rows per CTA (TMEM lanes)
constexpr int kBlockN = 256;        // accumulator columns per tile (UMMA N)
constexpr int kBlockK = 64;         // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kMaxStages = 8;
constexpr int kMaxAKb = 12;         // resident A: up to 12 k-blocks (Kd <= 768)
constexpr int kGemmThreads = 256;
constexpr int kABlockBytes = kBlockM * kBlockK * 2;   // 16 KiB

struct GemmShape {
    int rows_a;        // valid rows of A (output rows)
    int rows_b;        // valid rows of B (output columns)
    int num_kb;        // contraction length / 64 (operands are zero-padded to a multiple of 64)
    int num_rb;        // row blocks of 128 * kCtaGroup rows
    int num_ct;        // column tiles of 256
    int num_ks;        // contraction splits (partial results, 1 = none)
    int kb_per_split;  // k-blocks per split
    int split_mode;    // 0: balanced contiguous unit ranges; 1: worker = rb * fixed_s + s
    int fixed_s;
    int num_stages;    // ring depth chosen by the host from the shared-memory budget
};

struct TileCtx {
    int rb, ct, ks;     // row block, column tile, contraction split
    int row;            // global output row owned by this thread
    int col0;           // first global output column of this tile
    int worker;         // CTA (or pair) index
    int cta_rank;       // rank inside the pair
    int num_workers;
};

__host__ __device__ inline size_t gemm_smem_bytes(int cta_group, bool a_resident, int num_kb, int stages) {
    const size_t b_bytes = (size_t)(kBlockN / cta_group) * kBlockK * 2;
    const size_t stage = b_bytes + (a_resident ? 0 : kABlockBytes);
    const size_t a_res = a_resident ? (size_t)num_kb * kABlockBytes : 0;
    return 1024 /*align slack*/ + a_res + stage * stages + 1024 /*barriers*/;
}

__device__ __forceinline__ void unit_range(const GemmShape& sh, int worker, int num_workers, int& u0, int& u1) {
    const int per_rb = sh.num_ct * sh.num_ks;
    if (sh.split_mode == 0) {
        const long long total = (long long)sh.num_rb * per_rb;
        u0 = (int)(total * worker / num_workers);
        u1 = (int)(total * (worker + 1) / num_workers);
    } else {
        const int rb = worker / sh.fixed_s, s = worker % sh.fixed_s;
        if (rb >= sh.num_rb) { u0 = u1 = 0; return; }
        u0 = rb * per_rb + (int)((long long)per_rb * s / sh.fixed_s);
        u1 = rb * per_rb + (int)((long long)per_rb * (s + 1) / sh.fixed_s);
    }
}

template <int kCtaGroup, bool kAResident, class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmShape sh, const typename Epi::Params ep) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
    const bool leader = (cta_rank == 0);
    const int worker = blockIdx.x / kCtaGroup;
    const int num_workers = gridDim.x / kCtaGroup;

    constexpr uint32_t kBRows = kBlockN / kCtaGroup;              // B rows this CTA loads per tile
    constexpr uint32_t kBBlockBytes = kBRows * kBlockK * 2;
    constexpr uint32_t kStageBytes = kBBlockBytes + (kAResident ? 0 : kABlockBytes);
    const int stages = sh.num_stages;

    const uint32_t a_res = smem_base;                                        // resident A k-blocks
    const uint32_t ring = smem_base + (kAResident ? sh.num_kb * kABlockBytes : 0);
    const uint32_t bars = ring + stages * kStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
    auto afull_bar = [&](int k) { return bars + 8u * (2 * kMaxStages + k); };
    auto aempty_bar = [&](int k) { return bars + 8u * (2 * kMaxStages + kMaxAKb + k); };
    auto tfull_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 2 * kMaxAKb + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 2 * kMaxAKb + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 2 * kMaxAKb + 4);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), kCtaGroup); mbar_init(empty_bar(s), 1); }
        for (int k = 0; k < kMaxAKb; ++k) { mbar_init(afull_bar(k), kCtaGroup); mbar_init(aempty_bar(k), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 128 * kCtaGroup); }
        fence_mbar_init();
    }
    if constexpr (kCtaGroup == 2) cluster_sync_all();   // peer barriers must exist before remote arrives
    if (warp == 2) tmem_alloc<kCtaGroup>(tmem_slot, 512);
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    int u0, u1;
    unit_range(sh, worker, num_workers, u0, u1);
    const int per_rb = sh.num_ct * sh.num_ks;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            int prev_rb = -1, rbi = -1;
            for (int u = u0; u < u1; ++u) {
                const int rb = u / per_rb, rem = u - rb * per_rb;
                const int ct = rem / sh.num_ks, ks = rem - ct * sh.num_ks;
                const bool new_rb = (rb != prev_rb);
                if (new_rb) { ++rbi; prev_rb = rb; }
                const int row0 = (rb * kCtaGroup + (int)cta_rank) * kBlockM;
                const int brow0 = ct * kBlockN + (int)cta_rank * (int)kBRows;
                const int kb0 = ks * sh.kb_per_split;
                const int kb1 = min(sh.num_kb, kb0 + sh.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (kAResident && new_rb) {
                        mbar_wait(aempty_bar(kb), (rbi & 1) ^ 1);
                        if constexpr (kCtaGroup == 1) {
                            mbar_arrive_expect_tx(afull_bar(kb), kABlockBytes);
                            tma_load_2d(a_res + kb * kABlockBytes, &tmap_a, afull_bar(kb), kb * kBlockK, row0);
                        } else {
                            if (leader) mbar_arrive_expect_tx(afull_bar(kb), 2 * kABlockBytes);
                            else mbar_arrive_remote(afull_bar(kb), 0);
                            tma_load_2d_pair(a_res + kb * kABlockBytes, &tmap_a, afull_bar(kb), kb * kBlockK, row0);
                        }
                    }
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sbase = ring + stage * kStageBytes;
                    if constexpr (kCtaGroup == 1) {
                        mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
                        if constexpr (!kAResident) tma_load_2d(sbase + kBBlockBytes, &tmap_a, full_bar(stage), kb * kBlockK, row0);
                        tma_load_2d(sbase, &tmap_b, full_bar(stage), kb * kBlockK, brow0);
                    } else {
                        if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * kStageBytes);
                        else mbar_arrive_remote(full_bar(stage), 0);
                        if constexpr (!kAResident) tma_load_2d_pair(sbase + kBBlockBytes, &tmap_a, full_bar(stage), kb * kBlockK, row0);
                        tma_load_2d_pair(sbase, &tmap_b, full_bar(stage), kb * kBlockK, brow0);
                    }
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
            // Drain: every tcgen05.commit aimed at this CTA's barriers must have landed before the
            // CTA may exit (in pair mode they are multicast from the leader).
            for (int s = 0; s < stages; ++s) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (kAResident && rbi >= 0)
                for (int kb = 0; kb < sh.num_kb; ++kb) mbar_wait(aempty_bar(kb), rbi & 1);
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA)
        if (leader) {
            constexpr uint32_t idesc = make_idesc_bf16(kBlockM * kCtaGroup, kBlockN);
            int stage = 0; uint32_t phase = 0;
            int prev_rb = -1, rbi = -1, it = 0;
            for (int u = u0; u < u1; ++u, ++it) {
                const int rb = u / per_rb, rem = u - rb * per_rb;
                const int ct = rem / sh.num_ks, ks = rem - ct * sh.num_ks;
                (void)ct;
                const bool new_rb = (rb != prev_rb);
                if (new_rb) { ++rbi; prev_rb = rb; }
                const bool last_of_rb = (u + 1 == u1) || ((u + 1) / per_rb != rb);
                const int kb0 = ks * sh.kb_per_split;
                const int kb1 = min(sh.num_kb, kb0 + sh.kb_per_split);
                const int buf = it & 1;
                mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * kBlockN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (kAResident && new_rb) mbar_wait(afull_bar(kb), rbi & 1);
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sbase = ring + stage * kStageBytes;
                        const uint32_t a_addr = kAResident ? (a_res + kb * kABlockBytes) : (sbase + kBBlockBytes);
                        const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                        const uint64_t bdesc = make_kmajor_sw128_desc(sbase);
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            // +32 B per UMMA_K step inside the 128-byte swizzle row: +2 in the >>4 address field
                            umma_bf16<kCtaGroup>(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit<kCtaGroup>(empty_bar(stage));
                        if (kAResident && last_of_rb) umma_commit<kCtaGroup>(aempty_bar(kb));
                        if (kb + 1 == kb1) umma_commit<kCtaGroup>(tfull_bar(buf));
                    }
                    __syncwarp();
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (both CTAs of a pair)
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        typename Epi::State st;
        TileCtx cx;
        cx.worker = worker; cx.cta_rank = (int)cta_rank; cx.num_workers = num_workers;
        int prev_rb = -1, it = 0;
        for (int u = u0; u < u1; ++u, ++it) {
            const int rb = u / per_rb, rem = u - rb * per_rb;
            cx.rb = rb; cx.ct = rem / sh.num_ks; cx.ks = rem - cx.ct * sh.num_ks;
            cx.row = (rb * kCtaGroup + (int)cta_rank) * kBlockM + q * 32 + lane;
            cx.col0 = cx.ct * kBlockN;
            if (rb != prev_rb) { prev_rb = rb; Epi::begin_rb(st, ep, cx); }
            const bool last_of_rb = (u + 1 == u1) || ((u + 1) / per_rb != rb);
            const int buf = it & 1;
            mbar_wait(tfull_bar(buf), (it >> 1) & 1);
            tc_fence_after();
            Epi::tile(st, ep, cx, tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBlockN);
            tc_fence_before();
            if constexpr (kCtaGroup == 1) mbar_arrive(tempty_bar(buf));
            else mbar_arrive_remote(tempty_bar(buf), 0);
            if (last_of_rb) Epi::end_rb(st, ep, cx);
        }
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 2) tmem_dealloc<kCtaGroup>(tmem_base, 512);
}

}  // namespace pero
