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
//   warps 0..7  epilogue: thread <-> accumulator row, tcgen05.ld 32 columns at a time.  Two warps per SM
//               sub-partition: warps 0..3 reduce columns 0..127 of a tile, warps 4..7 columns 128..255,
//               so each scheduler always has a second warp to issue from while one waits on TMEM.
//               A per-column vector (|c|^2 or the bias) is staged through shared memory one tile ahead.
//   warp 8      TMA producer (one elected lane)
//   warp 9      MMA issuer   (one elected lane, leader CTA only)
//   warp 10     barrier init + TMEM allocator
//   The single-lane roles sit at the HIGHEST warp ids on purpose: the warp scheduler prefers the highest
//   eligible warp id, and a producer / MMA issuer that loses issue slots to ALU-heavy epilogue warps
//   stalls the tensor pipe (measured: 2400 instead of 2048 cycles per 256-deep tile).
//
// With kARes >= 1 the A row block is loaded ONCE per row block into its own k-block slots and only B streams
// through the stage ring; kARes == 2 keeps TWO such sets (short contractions, Kd <= 384), so the next row block's A is
// loaded while the current one is still being multiplied and a row-block change does not drain the MMA queue.
// kARes == 0: A and B k-blocks share the ring (long contractions: dW, dh).
//
// The epilogue is a policy class (see epilogues.cuh / masked_ce.cu):
//   struct Epi { static constexpr bool kColVec; struct Params; struct State;
//     static __device__ void begin_rb(State&, const Params&, const TileCtx&);
//     static __device__ void tile    (State&, const Params&, const TileCtx&, uint32_t taddr);   // 128 columns
//     static __device__ void end_rb  (State&, const Params&, const TileCtx&); };
// Row-reducing epilogues (arg-min, log-sum-exp) keep their running state in registers across the
// column sweep, so the rows x columns matrix never reaches HBM.
#pragma once
#include "ptx.cuh"

// 1: one elected lane of the MMA warp walks the whole issue loop; 0: the warp walks it and elects a lane per k-block
#ifndef PERO_MMA_SINGLE_LANE
#define PERO_MMA_SINGLE_LANE 0
#endif

namespace pero {

constexpr int kBlockM = 128;        // accumulator rows per CTA (TMEM lanes)
constexpr int kBlockN = 256;        // accumulator columns per tile (UMMA N)
constexpr int kBlockK = 64;         // bf16 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kMaxStages = 8;
constexpr int kMaxAKb = 12;         // resident A: up to 12 k-block slots (Kd <= 768, or two sets of Kd <= 384)
constexpr int kGemmThreads = 384;
constexpr int kHalfN = kBlockN / 2; // columns per epilogue warp group
constexpr int kEpiThreads = 256;
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
#ifdef PERO_DEV_BUILD
    int fake_b;        // timing study: B is loaded only during the first pass over the ring
#endif
    int pdl;           // programmatic dependent launch: bit 0 = release the next kernel of the stream at once (it does
                       // not read this kernel's output), bit 1 = before exiting, wait for the previous kernel (this
                       // kernel was released early by it, does not read its output, and must not be seen to finish
                       // first), bit 2 = wait for the previous kernel after the set-up (barriers, TMEM, descriptor
                       // prefetch), before the first global access: the set-up overlaps the producer kernel's tail,
                       // bit 3 (with bit 2 and a resident A) = the B operand and the column vector are NOT written by
                       // the previous kernel: the first ring of B loads is issued before the wait, only the A loads
                       // and the epilogue's global accesses come after it,
                       // bit 4 = release the next kernel of the stream LATE: when this CTA has requested its last
                       // operand load (about one unit before it exits), so that a dependent GEMM -- which cannot be
                       // co-resident anyway -- is dispatched while the last tiles drain instead of after a launch gap
    unsigned long long* timeline;   // debug: per-unit clock64 stamps of worker 0 ([unit][8]); NULL in production
};

struct TileCtx {
    int rb, ct, ks;     // row block, column tile, contraction split
    int row;            // global output row owned by this thread
    int col0;           // first global output column of this thread's half tile (128 columns)
    int half;           // 0: columns 0..127 of the tile, 1: columns 128..255
    const float* cv;    // shared-memory copy of Params::colvec[col0 .. col0 + 128) (when Epi::kColVec)
    uint8_t* scratch;   // warp-private shared memory, Epi::kScratchPerWarp bytes (16-byte aligned)
    int worker;         // CTA (or pair) index
    int cta_rank;       // rank inside the pair
    int num_workers;
};

__host__ __device__ inline size_t gemm_smem_bytes(int cta_group, int a_sets, int num_kb, int stages,
                                                  int scratch_per_warp) {
    const size_t b_bytes = (size_t)(kBlockN / cta_group) * kBlockK * 2;
    const size_t stage = b_bytes + (a_sets ? 0 : kABlockBytes);
    const size_t a_res = (size_t)a_sets * num_kb * kABlockBytes;
    return 1024 /*align slack*/ + a_res + stage * stages + 1024 /*barriers*/ + 2 * kBlockN * 4 /*column vector x2*/ +
           (size_t)scratch_per_warp * (kEpiThreads / 32);
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

// Walks the units [u0, u1) of one worker in (row block, column tile, contraction split) order without
// a division per unit.
struct UnitIter {
    int u, u1, rb, ct, ks, num_ct, num_ks;
    __device__ __forceinline__ UnitIter(const GemmShape& sh, int u0, int u1_)
        : u(u0), u1(u1_), num_ct(sh.num_ct), num_ks(sh.num_ks) {
        const int per_rb = num_ct * num_ks;
        rb = u0 / per_rb;
        const int rem = u0 - rb * per_rb;
        ct = rem / num_ks;
        ks = rem - ct * num_ks;
    }
    __device__ __forceinline__ bool valid() const { return u < u1; }
    __device__ __forceinline__ bool last_of_rb() const { return (u + 1 == u1) || (ks + 1 == num_ks && ct + 1 == num_ct); }
    __device__ __forceinline__ int next_ct() const { return (ks + 1 < num_ks) ? ct : ((ct + 1 == num_ct) ? 0 : ct + 1); }
    __device__ __forceinline__ void next() {
        ++u;
        if (++ks == num_ks) { ks = 0; if (++ct == num_ct) { ct = 0; ++rb; } }
    }
};

__device__ __forceinline__ void stamp(const GemmShape& sh, bool on, int unit, int slot) {
    if (on) sh.timeline[unit * 8 + slot] = clock64();
}
// debug: per-CTA wall-clock (ns) stamps after the per-unit area: [4096 + cta * 4 + slot]
__device__ __forceinline__ void stamp_cta(const GemmShape& sh, bool on, int block, int slot) {
    if (on && sh.timeline) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        sh.timeline[4096 + block * 4 + slot] = t;
    }
}

// Epi::kMaxRegs registers per thread (104 where the epilogue fits: 384 threads -> 39 K of the SM's 64 K registers; 128
// otherwise -> 48 K): the remaining 25 K (16 K) let two or three CTAs (one CTA) of a bandwidth-bound kernel or of the
// peer-exchange kernel run on the same SM, so those kernels overlap a resident GEMM instead of waiting for it (or, worse,
// keeping the next GEMM's CTA off the SM).
// kMajor (bit 0: A, bit 1: B) marks operands that are read "transposed" (MN-major): with both bits set A is stored
// [k][rows_a] and B [k][rows_b] (the contraction index
// runs over the rows of the stored matrices), i.e. C = A^T B without a transposed copy of either in HBM.  TMA
// fetches boxes of {64 contiguous MN-elements, 64 k-rows}; the UMMA descriptors are MN-major.
// The kernel body: `block` / `nblocks` are this CTA's index and the CTA count of ITS GEMM (a dual launch runs two
// GEMMs in one grid, see gemm_dual_kernel).
template <int kCtaGroup, int kARes, class Epi, int kMajor>
__device__ __forceinline__ void gemm_tn_body(const CUtensorMap& tmap_a, const CUtensorMap& tmap_b, const GemmShape& sh,
                                             const typename Epi::Params& ep, const int block, const int nblocks) {
    constexpr bool kAResident = kARes != 0;
    constexpr int kASets = kARes == 2 ? 2 : 1;
    constexpr bool kAMn = (kMajor & 1) != 0, kBMn = (kMajor & 2) != 0;
    static_assert(!(kMajor && kAResident), "MN-major operands are streamed through the ring");
    if (sh.pdl & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    stamp_cta(sh, threadIdx.x == 0, block, 0);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
    const bool leader = (cta_rank == 0);
    const int worker = block / kCtaGroup;
    const int num_workers = nblocks / kCtaGroup;

    constexpr uint32_t kBRows = kBlockN / kCtaGroup;              // B rows this CTA loads per tile
    constexpr uint32_t kBBlockBytes = kBRows * kBlockK * 2;
    constexpr uint32_t kStageBytes = kBBlockBytes + (kAResident ? 0 : kABlockBytes);
    const int stages = sh.num_stages;

    const uint32_t a_res = smem_base;                                        // resident A k-blocks
    const uint32_t ring = smem_base + (kAResident ? kASets * sh.num_kb * kABlockBytes : 0);
    const uint32_t bars = ring + stages * kStageBytes;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (kMaxStages + s); };
    auto afull_bar = [&](int k) { return bars + 8u * (2 * kMaxStages + k); };
    auto aempty_bar = [&](int k) { return bars + 8u * (2 * kMaxStages + kMaxAKb + k); };
    auto tfull_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 2 * kMaxAKb + b); };
    auto tempty_bar = [&](int b) { return bars + 8u * (2 * kMaxStages + 2 * kMaxAKb + 2 + b); };
    const uint32_t tmem_slot = bars + 8u * (2 * kMaxStages + 2 * kMaxAKb + 4);
    float* const cv_smem = reinterpret_cast<float*>(smem_raw + (bars + 1024u - smem_u32(smem_raw)));   // [2][256]
    uint8_t* const scratch_smem = reinterpret_cast<uint8_t*>(cv_smem) + 2 * kBlockN * 4;

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 10 && lane == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(full_bar(s), kCtaGroup); mbar_init(empty_bar(s), 1); }
        for (int k = 0; k < kMaxAKb; ++k) { mbar_init(afull_bar(k), kCtaGroup); mbar_init(aempty_bar(k), 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kEpiThreads * kCtaGroup); }
        fence_mbar_init();
    }
    if constexpr (kCtaGroup == 2) cluster_sync_all();   // peer barriers must exist before remote arrives
    if (warp == 10) tmem_alloc<kCtaGroup>(tmem_slot, 512);
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    // bit 3: only the threads that touch memory written by the previous kernel wait, and as late as they can (below)
    const bool late_wait = kAResident && (sh.pdl & 12) == 12;
    if ((sh.pdl & 4) && !late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");

    stamp_cta(sh, threadIdx.x == 0, block, 1);
    int u0, u1;
    unit_range(sh, worker, num_workers, u0, u1);
    const bool tl = sh.timeline != nullptr && worker == 0 && cta_rank == 0;

    if (warp == 8) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one_sync()) {
            int stage = 0; uint32_t phase = 0;
            int prev_rb = -1, rbi = -1;
            // late_wait: B never comes from the kernel this launch programmatically depends on, so the first ring of B
            // k-blocks is requested at once; the A loads (and everything else) follow the wait.
            int pre_issued = 0;
            bool waited = !late_wait;
#ifdef PERO_DEV_BUILD
            bool fake_started = false;
#endif
            if constexpr (kAResident) {
                if (late_wait) {
                    for (UnitIter ui(sh, u0, u1); ui.valid() && pre_issued < stages; ui.next()) {
                        const int brow0 = ui.ct * kBlockN + (int)cta_rank * (int)kBRows;
                        for (int kb = 0; kb < sh.num_kb && pre_issued < stages; ++kb, ++pre_issued) {
                            const uint32_t sbase = ring + pre_issued * kStageBytes;        // first pass over the ring: all free
                            if constexpr (kCtaGroup == 1) {
                                mbar_arrive_expect_tx(full_bar(pre_issued), kStageBytes);
                                tma_load_2d(sbase, &tmap_b, full_bar(pre_issued), kb * kBlockK, brow0);
                            } else {
                                if (leader) mbar_arrive_expect_tx(full_bar(pre_issued), 2 * kStageBytes);
                                else mbar_arrive_remote(full_bar(pre_issued), 0);
                                tma_load_2d_pair(sbase, &tmap_b, full_bar(pre_issued), kb * kBlockK, brow0);
                            }
                        }
                    }
                }
            }
            for (UnitIter ui(sh, u0, u1); ui.valid(); ui.next()) {
                const bool new_rb = (ui.rb != prev_rb);
                if (new_rb) { ++rbi; prev_rb = ui.rb; }
                const int row0 = (ui.rb * kCtaGroup + (int)cta_rank) * kBlockM;
                const int brow0 = ui.ct * kBlockN + (int)cta_rank * (int)kBRows;
                const int kb0 = ui.ks * sh.kb_per_split;
                const int kb1 = min(sh.num_kb, kb0 + sh.kb_per_split);
                const int aset = (kASets == 2) ? (rbi & 1) : 0;
                const int ause = (kASets == 2) ? (rbi >> 1) : rbi;        // how often this A set has been filled before
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (kb == kb0) stamp(sh, tl, ui.u - u0, 6);
                    if (kAResident && new_rb) {
                        if (!waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); waited = true; }
                        const int slot = aset * sh.num_kb + kb;
                        mbar_wait(aempty_bar(slot), (ause & 1) ^ 1);
                        if constexpr (kCtaGroup == 1) {
                            mbar_arrive_expect_tx(afull_bar(slot), kABlockBytes);
                            tma_load_2d(a_res + slot * kABlockBytes, &tmap_a, afull_bar(slot), kb * kBlockK, row0);
                        } else {
                            if (leader) mbar_arrive_expect_tx(afull_bar(slot), 2 * kABlockBytes);
                            else mbar_arrive_remote(afull_bar(slot), 0);
                            tma_load_2d_pair(a_res + slot * kABlockBytes, &tmap_a, afull_bar(slot), kb * kBlockK, row0);
                        }
                    }
                    if (pre_issued > 0) {               // this k-block's B was requested ahead of the wait
                        --pre_issued;
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                        if (kb + 1 == kb1) stamp(sh, tl, ui.u - u0, 7);
                        continue;
                    }
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sbase = ring + stage * kStageBytes;
#ifdef PERO_DEV_BUILD
                    if (sh.fake_b && (phase || fake_started)) {          // timing study only: stage contents are stale
                        fake_started = true;
                        if constexpr (kCtaGroup == 1) mbar_arrive(full_bar(stage));
                        else { if (leader) mbar_arrive(full_bar(stage)); else mbar_arrive_remote(full_bar(stage), 0); }
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
#endif
                    {
                        // K-major operand: one box {64 k-elements, rows}.  MN-major operand: boxes of {64 MN-elements,
                        // 64 k-rows} = 8 KiB each (2 for A, kBRows / 64 for B), the contraction index running over rows.
                        constexpr uint32_t kBox = 64u * kBlockK * 2u;
                        if constexpr (kCtaGroup == 1) {
                            mbar_arrive_expect_tx(full_bar(stage), kStageBytes);
                        } else {
                            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * kStageBytes);
                            else mbar_arrive_remote(full_bar(stage), 0);
                        }
                        auto load = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1) {
                            if constexpr (kCtaGroup == 1) tma_load_2d(dst, m, full_bar(stage), c0, c1);
                            else tma_load_2d_pair(dst, m, full_bar(stage), c0, c1);
                        };
                        if constexpr (!kAResident) {
                            if constexpr (kAMn) {
#pragma unroll
                                for (int j = 0; j < kBlockM / 64; ++j) load(sbase + kBBlockBytes + j * kBox, &tmap_a, row0 + 64 * j, kb * kBlockK);
                            } else {
                                load(sbase + kBBlockBytes, &tmap_a, kb * kBlockK, row0);
                            }
                        }
                        if constexpr (kBMn) {
#pragma unroll
                            for (int j = 0; j < (int)kBRows / 64; ++j) load(sbase + j * kBox, &tmap_b, brow0 + 64 * j, kb * kBlockK);
                        } else {
                            load(sbase, &tmap_b, kb * kBlockK, brow0);
                        }
                    }
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                    if (kb + 1 == kb1) stamp(sh, tl, ui.u - u0, 7);
                }
            }
            if (sh.pdl & 16) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
            // Drain: every tcgen05.commit aimed at this CTA's barriers must have landed before the
            // CTA may exit (in pair mode they are multicast from the leader).
            for (int s = 0; s < stages; ++s) {
                mbar_wait(empty_bar(stage), phase ^ 1);
                if (++stage == stages) { stage = 0; phase ^= 1; }
            }
            if (kAResident && rbi >= 0) {
                for (int set = 0; set < kASets; ++set) {
                    const int uses = (kASets == 2) ? ((rbi - set) >= 0 ? (rbi - set) / 2 + 1 : 0) : rbi + 1;
                    if (uses == 0) continue;
                    for (int kb = 0; kb < sh.num_kb; ++kb) mbar_wait(aempty_bar(set * sh.num_kb + kb), (uses - 1) & 1);
                }
            }
        }
    } else if (warp == 9) {
        // ------------------------------------------------------------ MMA issuer (leader CTA)
        // The whole warp walks the loop (uniform control flow) and one elected lane issues.  The full-
        // barrier probe of the NEXT stage is issued before the MMAs of the current one, so its latency
        // hides behind the issue instead of adding ~100 cycles per k-block.
        // (Measured, round 2: handing alternate units to a second issuing warp that has already passed the waits of
        // the next unit does not shorten a unit -- the 16 MMAs of a 256 x 256 x 256 unit take ~2700 cycles back to
        // back either way -- so the single issuer stays.)
        constexpr bool single = PERO_MMA_SINGLE_LANE != 0;
        if (leader && u0 < u1 && (single ? elect_one_sync() : 1u)) {
            constexpr uint32_t idesc = make_idesc_bf16(kBlockM * kCtaGroup, kBlockN, kAMn, kBMn);
            int stage = 0; uint32_t phase = 0;
            int prev_rb = -1, rbi = -1, it = 0;
            uint32_t ready = 0;
            for (UnitIter ui(sh, u0, u1); ui.valid(); ui.next(), ++it) {
                const bool new_rb = (ui.rb != prev_rb);
                if (new_rb) { ++rbi; prev_rb = ui.rb; }
                const bool last_of_rb = ui.last_of_rb();
                const int kb0 = ui.ks * sh.kb_per_split;
                const int kb1 = min(sh.num_kb, kb0 + sh.kb_per_split);
                const int buf = it & 1;
                const int aset = (kASets == 2) ? (rbi & 1) : 0;
                const int ause = (kASets == 2) ? (rbi >> 1) : rbi;
                stamp(sh, tl && (single || lane == 0), it, 0);
                mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1);
                stamp(sh, tl && (single || lane == 0), it, 1);
                const uint32_t tmem_d = tmem_base + buf * kBlockN;
                for (int kb = kb0; kb < kb1; ++kb) {
                    if (kAResident && new_rb) mbar_wait(afull_bar(aset * sh.num_kb + kb), ause & 1);
                    if (!ready) mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (kb == kb0) stamp(sh, tl && (single || lane == 0), it, 2);
                    int nstage = stage + 1; uint32_t nphase = phase;
                    if (nstage == stages) { nstage = 0; nphase ^= 1; }
                    // probe only if another k-block follows: a probe of a phase that never completes would stall
                    ready = (kb + 1 < kb1 || ui.u + 1 < u1) ? mbar_test_wait(full_bar(nstage), nphase) : 0u;
                    const uint32_t sbase = ring + stage * kStageBytes;
                    const uint32_t a_addr = kAResident ? (a_res + (aset * sh.num_kb + kb) * kABlockBytes) : (sbase + kBBlockBytes);
                    constexpr uint32_t kMnBox = 64u * kBlockK * 2u;
                    const uint64_t adesc = kAMn ? make_mnmajor_sw128_desc(a_addr, kMnBox) : make_kmajor_sw128_desc(a_addr);
                    const uint64_t bdesc = kBMn ? make_mnmajor_sw128_desc(sbase, kMnBox) : make_kmajor_sw128_desc(sbase);
                    // address-field step per UMMA_K (16 contraction elements): K-major +32 B inside the 128-byte
                    // swizzle row (+2 in the >>4 field); MN-major +16 k-rows = two 1024-byte atoms (+128)
                    constexpr uint64_t kAStep = kAMn ? (2048u >> 4) : 2u, kBStep = kBMn ? (2048u >> 4) : 2u;
                    if (single || elect_one_sync()) {
#pragma unroll
                        for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                            umma_bf16<kCtaGroup>(tmem_d, adesc + kAStep * k, bdesc + kBStep * k, idesc,
                                                 (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit<kCtaGroup>(empty_bar(stage));
                        if (kAResident && last_of_rb) umma_commit<kCtaGroup>(aempty_bar(aset * sh.num_kb + kb));
                        if (kb + 1 == kb1) { umma_commit<kCtaGroup>(tfull_bar(buf)); stamp(sh, tl, it, 3); }
                    }
                    if (!single) __syncwarp();
                    stage = nstage; phase = nphase;
                }
            }
        }
    } else if (warp < 8) {
        // ------------------------------------------------------------ epilogue (both CTAs of a pair)
        if (late_wait) asm volatile("griddepcontrol.wait;" ::: "memory");     // before the first access to the previous kernel's output
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = warp >> 2;                   // which 128 columns of every tile
        const int ht = q * 32 + lane;                 // thread index inside its half group
        typename Epi::State st;
        TileCtx cx;
        cx.worker = worker; cx.cta_rank = (int)cta_rank; cx.num_workers = num_workers; cx.half = half; cx.cv = nullptr;
        cx.scratch = scratch_smem + warp * Epi::kScratchPerWarp;
        UnitIter ui(sh, u0, u1);
        float pre = 0.f;                              // column-vector element of the NEXT tile, loaded a tile ahead
        if constexpr (Epi::kColVec) {
            if (ui.valid()) pre = __ldg(ep.colvec + ui.ct * kBlockN + half * kHalfN + ht);
        }
        int prev_rb = -1, it = 0;
        for (; ui.valid(); ui.next(), ++it) {
            cx.rb = ui.rb; cx.ct = ui.ct; cx.ks = ui.ks;
            cx.row = (ui.rb * kCtaGroup + (int)cta_rank) * kBlockM + q * 32 + lane;
            cx.col0 = ui.ct * kBlockN + half * kHalfN;
            const int buf = it & 1;
            if constexpr (Epi::kColVec) {
                float* dst = cv_smem + buf * kBlockN + half * kHalfN;
                dst[ht] = pre;
                asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");     // this half group only
                cx.cv = dst;
                if (ui.u + 1 < u1) pre = __ldg(ep.colvec + ui.next_ct() * kBlockN + half * kHalfN + ht);
            }
            if (ui.rb != prev_rb) { prev_rb = ui.rb; Epi::begin_rb(st, ep, cx); }
            const bool last_of_rb = ui.last_of_rb();
            mbar_wait(tfull_bar(buf), (it >> 1) & 1);
            tc_fence_after();
            stamp(sh, tl && warp == 0 && lane == 0, it, 4);
            Epi::tile(st, ep, cx, tmem_base + ((uint32_t)(q * 32) << 16) + buf * kBlockN + half * kHalfN);
            tc_fence_before();
            if constexpr (kCtaGroup == 1) mbar_arrive(tempty_bar(buf));
            else mbar_arrive_remote(tempty_bar(buf), 0);
            stamp(sh, tl && warp == 0 && lane == 0, it, 5);
            if (last_of_rb) Epi::end_rb(st, ep, cx);
        }
        stamp_cta(sh, threadIdx.x == 0, block, 2);
    }

    // ---------------------------------------------------------------- teardown
    tc_fence_before();
    if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 10) tmem_dealloc<kCtaGroup>(tmem_base, 512);
    if (sh.pdl & 2) asm volatile("griddepcontrol.wait;" ::: "memory");
    stamp_cta(sh, threadIdx.x == 0, block, 3);
}

template <int kCtaGroup, int kARes, class Epi, int kMajor = 0>
__global__ void __maxnreg__(Epi::kMaxRegs)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmShape sh, const __grid_constant__ typename Epi::Params ep) {
    gemm_tn_body<kCtaGroup, kARes, Epi, kMajor>(tmap_a, tmap_b, sh, ep, (int)blockIdx.x, (int)gridDim.x);
}

// Two independent streamed GEMMs (same epilogue class, operand majorness kMajor1 / kMajor2) in ONE grid: the first
// `ctas1` CTAs run GEMM 1, the others GEMM 2.  One launch instead of two: the pair can be a programmatic dependent of
// the kernel in front as a whole (both halves wait for it after their set-up), which two back-to-back launches cannot
// express -- the second would wait for the first to complete.  Used for d_W = P^T A beside d_h = P W.
template <class Epi, int kMajor1, int kMajor2>
__global__ void __maxnreg__(Epi::kMaxRegs)
gemm_dual_kernel(const __grid_constant__ CUtensorMap tmap_a1, const __grid_constant__ CUtensorMap tmap_b1, const GemmShape sh1,
                 const __grid_constant__ typename Epi::Params ep1, const __grid_constant__ CUtensorMap tmap_a2,
                 const __grid_constant__ CUtensorMap tmap_b2, const GemmShape sh2,
                 const __grid_constant__ typename Epi::Params ep2, const int ctas1) {
    if ((int)blockIdx.x < ctas1)
        gemm_tn_body<2, 0, Epi, kMajor1>(tmap_a1, tmap_b1, sh1, ep1, (int)blockIdx.x, ctas1);
    else
        gemm_tn_body<2, 0, Epi, kMajor2>(tmap_a2, tmap_b2, sh2, ep2, (int)blockIdx.x - ctas1, (int)gridDim.x - ctas1);
}

}  // namespace pero
