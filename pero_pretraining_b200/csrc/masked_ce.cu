// Masked-label cross-entropy over codebook logits:
//   LinearHead                masked_pretraining/model.py:104-105
//   MaskedCrossEntropyLoss    masked_pretraining/model.py:78-82 (+ :84-93 by a second call)
// Only the M masked frames go through the head: gather -> logits GEMM (tcgen05) -> online log-sum-exp in
// the GEMM epilogue.  The backward recomputes the logits tile by tile, turns them into
// dlogits = (softmax - onehot) * scale in the epilogue (bf16, both orientations) and runs two more
// tcgen05 GEMMs for d_W = dlogits^T @ h and d_h = dlogits @ W (split over the label axis, fixed-order
// reduction), then scatters d_h back to the frame positions.  No atomics anywhere: deterministic.
#include <cub/device/device_select.cuh>
#include <algorithm>
#include <cuda_bf16.h>
#include <thrust/iterator/counting_iterator.h>
#include "../../include/pero_b200.h"
#include "epilogues.cuh"
#include "gemm_host.cuh"
#include "layout.h"

namespace pero {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// One MUFU.EX2 (rel. error ~2^-22, denormals flushed): exp2f() costs three more instructions per element
// for its denormal-input scaling, and the epilogues below are instruction-issue bound.
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Prepared head: bf16 W [V, Dhp] | bias fp32 [Vt] (-inf beyond V so padded label columns vanish from the
// log-sum-exp).  Dhp, Vp: rounded up to 64; Vt: rounded up to 256.  ONE bf16 copy serves all three GEMMs that touch the
// head: the logits GEMMs read it K-major (contraction over Dh), d_h = dlogits W reads it MN-major (contraction over V).
struct HeadLayout { int64_t Dhp, Vp, Vt; size_t w_off, bias_off, total; };
inline HeadLayout head_layout(int64_t V, int64_t Dh) {
    HeadLayout l;
    l.Dhp = round_up(Dh, 64); l.Vp = round_up(V, 64); l.Vt = round_up(V, 256);
    l.w_off = 0;
    l.bias_off = align256((size_t)V * l.Dhp * 2);
    l.total = l.bias_off + align256((size_t)l.Vt * 4);
    return l;
}

constexpr int kMaxLseSplits = 64;
constexpr int kDbRows = 8;          // masked frames per partial row of the d_b column sums

struct CeWsLayout {
    int64_t Dhp, Vp, Pp, Mp64, Mpad, S, KS;
    size_t a_off, lab_off, inv_off, p_off, pm_off, ps_off, zlab_off, rowloss_off, ticket_off, zlin_off, pcnt_off, dbpart_off,
        planes_off, cm_off, total;
};
inline CeWsLayout ce_ws_layout(int64_t N, int64_t M, int64_t V, int64_t Dh) {
    CeWsLayout l;
    // The logits GEMMs (forward LSE, backward dlogits) run on CTA pairs: 256-row blocks, 74 workers.
    l.Dhp = round_up(Dh, 64); l.Vp = round_up(V, 64); l.Mp64 = round_up(M, 64); l.Mpad = round_up(M, 256);
    // Row pitch of the dlogits matrix P: a pitch that is a multiple of 4 KiB would put the same 128-byte column
    // stripe of every row into the same few L2 sets (column sums and 8-rows-per-instruction stores walk exactly
    // that pattern); 64 extra elements stagger the rows.
    l.Pp = ((l.Vp * 2) % 4096 == 0) ? l.Vp + 64 : l.Vp;
    const int64_t num_rb_pair = l.Mpad / 256, num_rb = (M + 127) / 128, num_ct = (V + 255) / 256, num_ct_dh = (Dh + 255) / 256;
    int64_t S = 74 / num_rb_pair;
    if (S < 1) S = 1;
    if (S > num_ct) S = num_ct;
    if (S > kMaxLseSplits) S = kMaxLseSplits;
    l.S = S;
    int64_t KS = 148 / (num_rb * num_ct_dh);
    if (KS < 1) KS = 1;
    if (KS > l.Vp / 64) KS = l.Vp / 64;
    if (KS > 16) KS = 16;
    l.KS = KS;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    l.a_off = take((size_t)M * l.Dhp * 2);
    l.lab_off = take((size_t)l.Mpad * 4);
    l.inv_off = take((size_t)N * 4);
    l.p_off = take((size_t)M * l.Pp * 2);
    l.pm_off = take((size_t)2 * S * l.Mpad * 4);
    l.ps_off = take((size_t)2 * S * l.Mpad * 4);
    l.zlab_off = take((size_t)l.Mpad * 4);
    l.rowloss_off = take((size_t)l.Mpad * 4);
    l.ticket_off = take(256);
    l.zlin_off = take((size_t)l.Mpad * 4);              // evaluation: label logit known before the sweep
    l.pcnt_off = take((size_t)2 * S * l.Mpad * 4);      // evaluation: per-split counts of logits above the label's
    l.dbpart_off = take((size_t)((M + kDbRows - 1) / kDbRows) * l.Vp * 4);   // d_b partial column sums, one row per 8 masked frames
    l.planes_off = take((size_t)KS * M * Dh * 4);
    l.cm_off = take((size_t)(l.Vp / 32) * l.Mpad * 4);     // PERO_CE_KEEP_LOGITS: chunk maxima of the kept exponentials
    l.total = off;
    return l;
}

// ------------------------------------------------------------------------------------------------ prep
__global__ void __launch_bounds__(256)
head_prepare_kernel(const float* __restrict__ W, const float* __restrict__ bias, int V, int Dh, int Dhp, int Vt,
                    __nv_bfloat16* __restrict__ wb, float* __restrict__ bias_out) {
    // bf16 copy of W, rows zero-padded to Dhp columns: 4 elements per quad (16-byte loads when Dh % 4 == 0), 4 independent
    // quads per thread and iteration.  The grid is small on purpose (2 CTAs per SM): the kernel is issued at the start of a
    // step, beside the frame preparation of the quantizer, and must not take every thread slot of the machine.
    const long long quads = (long long)V * (Dhp / 4);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool vec = (Dh & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
    for (long long q0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; q0 < quads; q0 += 4 * stride) {
        float x[4][4];
        long long vv[4]; int dd[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long q = q0 + u * stride;
            const long long v = q / (Dhp / 4);
            const int d = (int)(q - v * (Dhp / 4)) * 4;
            vv[u] = v; dd[u] = d;
#pragma unroll
            for (int j = 0; j < 4; ++j) x[u][j] = 0.f;
            if (q < quads) {
                if (vec) {
                    if (d < Dh) { const float4 t = __ldg(reinterpret_cast<const float4*>(W + v * Dh + d)); x[u][0] = t.x; x[u][1] = t.y; x[u][2] = t.z; x[u][3] = t.w; }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (d + j < Dh) x[u][j] = __ldg(W + v * Dh + d + j);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (q0 + u * stride < quads) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(x[u][0], x[u][1]), hi = __floats2bfloat162_rn(x[u][2], x[u][3]);
                uint2 o;
                o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(wb + vv[u] * Dhp + dd[u]) = o;
            }
        }
    }
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < Vt; v += stride)
        bias_out[v] = v < V ? (bias ? __ldg(bias + v) : 0.f) : -CUDART_INF_F;
}

// Gather the masked frames into the GEMM operand A [M, Dhp] bf16 (rows = masked frames: the K-major operand of the
// logits GEMMs and, read through MN-major descriptors, the operand of d_W = dlogits^T A) and the inverse
// frame -> masked-row map inv [N] at the selected frames (see masked_row_of).  One warp per masked row.
// Also clears the ticket word that ce_finalize_kernel's last block uses.  Nothing here reads the labels, so a
// caller whose labels are produced late (by the quantizer of the same step) can gather ahead of them.
template <typename T>
__global__ void __launch_bounds__(256)
ce_gather_kernel(const T* __restrict__ h, const int* __restrict__ rows, int M, int Dh, int Dhp,
                 __nv_bfloat16* __restrict__ a, int* __restrict__ inv, unsigned int* __restrict__ ticket) {
    // the logits GEMM behind this kernel sets itself up meanwhile and waits for this grid before its first load
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) *ticket = 0u;
    if (m >= M) return;
    const int r = __ldg(rows + m);
    const T* src = h + (size_t)r * Dh;
    __nv_bfloat16* dst = a + (size_t)m * Dhp;
    if constexpr (sizeof(T) == 4) {
        if ((Dh & 3) == 0) {                            // 16-byte loads, 8-byte stores
            for (int d = 4 * lane; d < Dhp; d += 128) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (d < Dh) x = __ldg(reinterpret_cast<const float4*>(src + d));
                const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
                uint2 o;
                o.x = *reinterpret_cast<const uint32_t*>(&lo); o.y = *reinterpret_cast<const uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(dst + d) = o;
            }
            if (lane == 0) inv[r] = m;
            return;
        }
    }
    for (int d = 2 * lane; d < Dhp; d += 64) {          // Dhp is a multiple of 64: every lane writes whole pairs
        const float x0 = d < Dh ? (float)src[d] : 0.f, x1 = d + 1 < Dh ? (float)src[d + 1] : 0.f;
        *reinterpret_cast<__nv_bfloat162*>(dst + d) = __floats2bfloat162_rn(x0, x1);
    }
    if (lane == 0) inv[r] = m;
}

// Label of masked row `row` (-1 beyond M; -2 when the label is outside [0, V): the reference's F.cross_entropy raises a
// device assert there, here the row's loss becomes NaN and no one-hot is subtracted).
// packed: `labels` holds the packed (distance, index) winners of pero_vq_assign (the label is the low 32 bits), which
// lets the head of the same step start right behind the distance GEMM, without waiting for pero_vq_unpack.
__device__ __forceinline__ int masked_label(const long long* __restrict__ labels, const int* __restrict__ rows, int row, int M, int V,
                                            int packed) {
    if (row >= M) return -1;
    long long l = __ldg(labels + __ldg(rows + row));
    if (packed) l = (long long)(unsigned long long)(l & 0xffffffffll);
    return (l < 0 || l >= V) ? -2 : (int)l;
}

// The frame -> masked-row map `inv` is written only at the selected frames; every other entry keeps whatever the
// workspace held.  A reader validates a candidate m = inv[n] against `rows`: it is n's row iff 0 <= m < M and
// rows[m] == n (rows are distinct, so a stale or garbage entry can never pass for the wrong frame).
__device__ __forceinline__ int masked_row_of(const int* __restrict__ inv, const int* __restrict__ rows, long long n, int M) {
    const int m = __ldg(inv + n);
    return ((unsigned)m < (unsigned)M && (long long)__ldg(rows + m) == n) ? m : -1;
}

// ------------------------------------------------------------------------------------------------ epilogues
// Online log-sum-exp over the label axis; one partial (max, sum) per row, column split and column half.
// kRank (evaluation, masked_pretraining/tester.py:70-93): the label's logit is known before the sweep (zl_in) and
// the epilogue also counts the labels whose logit is strictly larger — the label's 0-based rank, from which the
// top-k errors follow without the [M, V] logits ever existing.
// kStore (training forward, PERO_CE_KEEP_LOGITS): the sweep also leaves, in the workspace's P area [M, Pp], the
// exponentials of the logits relative to the maximum of their own 32-column chunk, e = exp(z - cmax) in (0, 1], as bf16,
// and the chunk maxima cm [Vp / 32][Mpad] in fp32.  They are the numerators of the softmax up to one fp32 factor per
// (row, chunk): the backward turns P into the dlogits in place with a multiplication, P = e * exp(cm - lse) * scale
// (ce_dlogits_inplace_kernel) -- no second logits GEMM and no second exponential per element, and the bf16 rounding of P
// is a RELATIVE 2^-9 whatever the magnitude of the logits (bf16 logits would lose absolute precision as they grow).
// The log-sum-exp itself is accumulated from the fp32 exponentials, as without kStore.
// Staging as in DlogitsEpi: one SWIZZLE_128B tile of 32 rows x 64 columns per warp, handed to a TMA store.
template <bool kRank, bool kStore = false>
struct LseEpiT {
    static constexpr int kMaxRegs = kRank ? 128 : 104;
    static constexpr bool kColVec = true;
    // 8 warps x 192 B = 128 rows x (max, sum, count): half 1 -> half 0 hand-over; kStore: the 8 staging tiles come first
    // (1024-byte aligned, as the swizzle pattern requires), the hand-over area behind them
    static constexpr int kTileBytes = kStore ? 4096 : 0;
    static constexpr int kScratchPerWarp = 192 + kTileBytes;
    struct Params {
        CUtensorMap tmap_p;   // kStore: P [M, Vp] bf16, box {64 columns, 32 rows}
        const float* colvec;  // bias [Vt], -inf beyond V
        const int* rows;      // [M] frame of every masked row
        const long long* labels;   // [N] label of every frame
        float* pm; float* ps; // [S, Mpad] partial max / sum(exp(z - max)) of every worker's column range
        float* zlab;          // [Mpad] logit at the label
        const float* zl_in;   // kRank: [Mpad] label logit computed ahead of the sweep
        int* pcnt;            // kRank: [S, Mpad] partial counts of logits above zl_in
        int M, Mpad, S, V, packed;
        int Vp;               // kStore: label columns of P to write (multiple of 64)
        float* cm;            // kStore: [Vp / 32][Mpad] chunk maxima
    };
    struct State { float m, s, zl, zin; int label, cnt; bool has; };
    static __device__ __forceinline__ void begin_rb(State& st, const Params& ep, const TileCtx& cx) {
        st.m = -CUDART_INF_F; st.s = 0.f; st.zl = 0.f; st.has = false;
        st.label = masked_label(ep.labels, ep.rows, cx.row, ep.M, ep.V, ep.packed);
        st.cnt = 0;
        st.zin = (kRank && cx.row < ep.M) ? __ldg(ep.zl_in + cx.row) : CUDART_INF_F;
    }
    static __device__ __forceinline__ void tile(State& st, const Params& ep, const TileCtx& cx, uint32_t taddr) {
        const float4* cv = reinterpret_cast<const float4*>(cx.cv);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        // cx.scratch = base + warp * kScratchPerWarp; this warp's staging tile is base + warp * 4096
        const uint32_t stage = smem_u32(cx.scratch) - (uint32_t)(warp * 192);
        for_each_chunk(taddr, [&](int c, const uint32_t (&r)[32]) {
            const int col = cx.col0 + c * 32;
            float z[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = cv[c * 8 + i];
                z[4 * i + 0] = __uint_as_float(r[4 * i + 0]) + b.x;
                z[4 * i + 1] = __uint_as_float(r[4 * i + 1]) + b.y;
                z[4 * i + 2] = __uint_as_float(r[4 * i + 2]) + b.z;
                z[4 * i + 3] = __uint_as_float(r[4 * i + 3]) + b.w;
            }
            float m8[4];
#pragma unroll
            for (int g = 0; g < 4; ++g)
                m8[g] = fmaxf(fmaxf(fmaxf(z[8 * g], z[8 * g + 1]), fmaxf(z[8 * g + 2], z[8 * g + 3])),
                              fmaxf(fmaxf(z[8 * g + 4], z[8 * g + 5]), fmaxf(z[8 * g + 6], z[8 * g + 7])));
            const float cmax = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
            const unsigned rel = (unsigned)(st.label - col);
            if (rel < 32u) {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (rel == (unsigned)j) st.zl = z[j];
                st.has = true;
            }
            if constexpr (kRank) {
                int above = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) above += (z[j] > st.zin) ? 1 : 0;
                // the label's own column is never counted, however its two evaluations round
                if (rel < 32u && st.zl > st.zin) above -= 1;
                st.cnt += above;
            }
            if constexpr (!kStore) {
                if (cmax > -CUDART_INF_F) {
                    const float mn = fmaxf(st.m, cmax);
                    const float mn2 = mn * kLog2e;
                    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        a0 += ex2_fast(fmaf(z[j + 0], kLog2e, -mn2));
                        a1 += ex2_fast(fmaf(z[j + 1], kLog2e, -mn2));
                        a2 += ex2_fast(fmaf(z[j + 2], kLog2e, -mn2));
                        a3 += ex2_fast(fmaf(z[j + 3], kLog2e, -mn2));
                    }
                    st.s = st.s * ex2_fast((st.m - mn) * kLog2e) + ((a0 + a1) + (a2 + a3));
                    st.m = mn;
                }
            } else {
                // exponentials relative to the chunk's own maximum (an all-padding chunk: exp2(-inf - 0) = 0 everywhere)
                const bool live = cmax > -CUDART_INF_F;
                const float c2 = live ? cmax * kLog2e : 0.f;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    z[j + 0] = ex2_fast(fmaf(z[j + 0], kLog2e, -c2)); a0 += z[j + 0];
                    z[j + 1] = ex2_fast(fmaf(z[j + 1], kLog2e, -c2)); a1 += z[j + 1];
                    z[j + 2] = ex2_fast(fmaf(z[j + 2], kLog2e, -c2)); a2 += z[j + 2];
                    z[j + 3] = ex2_fast(fmaf(z[j + 3], kLog2e, -c2)); a3 += z[j + 3];
                }
                if (live) {
                    const float mn = fmaxf(st.m, cmax);
                    st.s = st.s * ex2_fast((st.m - mn) * kLog2e) + ((a0 + a1) + (a2 + a3)) * ex2_fast((cmax - mn) * kLog2e);
                    st.m = mn;
                }
                if (col < ep.Vp) {                                  // warp-uniform: Vp is a multiple of 64
                    ep.cm[(size_t)(col >> 5) * ep.Mpad + cx.row] = cmax;      // 32 consecutive rows per warp store
                    if ((c & 1) == 0) {                             // the previous store must have read the tile
                        if (lane == 0) tma_store_wait_read();
                        __syncwarp();
                    }
                    const uint32_t rowaddr = stage + (uint32_t)lane * 128u;
                    const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        __nv_bfloat162 o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(z[8 * i + 2 * j], z[8 * i + 2 * j + 1]);
                        const uint32_t piece = (uint32_t)((c & 1) * 4 + i) ^ sw;
                        const uint4 v = *reinterpret_cast<uint4*>(&o[0]);
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                                     ::"r"(rowaddr + piece * 16u), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
                    }
                    if (c & 1) {
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) {
                            tma_store_2d(&ep.tmap_p, stage, col - 32, cx.row - lane);
                            tma_store_commit();
                        }
                    }
                }
            }
        });
    }
    // The two column halves of a tile belong to two threads of the CTA that own the same row: half 1 hands its partial
    // (max, sum[, count]) to half 0 through shared memory, half 0 combines (half 0 first: a fixed order) and writes ONE
    // slot per worker, which halves what the readers of the partials have to fetch.
    static __device__ __forceinline__ void end_rb(State& st, const Params& ep, const TileCtx& cx) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        float* xch = reinterpret_cast<float*>(cx.scratch - warp * kScratchPerWarp + 8 * kTileBytes);     // [128][3]
        const int r = (warp & 3) * 32 + lane;                                                // row inside the CTA's 128
        const int slot = cx.worker % ep.S;
        // the staging tile must stay valid until the last store has read it (the CTA may exit right after this)
        if (kStore && lane == 0) tma_store_wait_read();
        if (st.has) ep.zlab[cx.row] = st.zl;
        else if (st.label == -2 && slot == 0 && cx.half == 0) ep.zlab[cx.row] = CUDART_NAN_F;      // label outside [0, V)
        if (cx.half == 1) {
            xch[3 * r + 0] = st.m; xch[3 * r + 1] = st.s; xch[3 * r + 2] = __int_as_float(st.cnt);
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");
        if (cx.half == 0) {
            const float m1 = xch[3 * r + 0], s1 = xch[3 * r + 1];
            const float mn = fmaxf(st.m, m1);
            float sum = 0.f;
            if (mn > -CUDART_INF_F) sum = st.s * ex2_fast((st.m - mn) * kLog2e) + s1 * ex2_fast((m1 - mn) * kLog2e);
            ep.pm[(size_t)slot * ep.Mpad + cx.row] = mn;
            ep.ps[(size_t)slot * ep.Mpad + cx.row] = sum;
            if constexpr (kRank) ep.pcnt[(size_t)slot * ep.Mpad + cx.row] = st.cnt + __float_as_int(xch[3 * r + 2]);
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");                                         // the hand-over area is free again
    }
};
using LseEpi = LseEpiT<false>;
using LseStoreEpi = LseEpiT<false, true>;
using EvalEpi = LseEpiT<true>;

// dlogits tile = (exp(z - lse) - [col == label]) * scale, written bf16 as P [M, Vp] only.  Both gradient GEMMs read
// this one copy: d_h = P W consumes it K-major, d_W = P^T A through MN-major descriptors.
// Every warp stages 32 rows x 64 columns (two 32-column chunks) in its own 4 KiB shared-memory tile, laid out as a
// SWIZZLE_128B TMA box (row r at r * 128 B, its 16-byte piece j at position j ^ (r & 7): conflict-free 16-byte
// stores), and one lane hands the tile to a TMA store (cp.async.bulk.tensor, clipped against M and Vp by the
// descriptor): no per-row predicates, no address arithmetic, full 128-byte row segments on the way to L2
// (16 us instead of 24 us for the kernel at the bench shape, round 2).
struct DlogitsEpi {
    static constexpr int kMaxRegs = 128;
    static constexpr bool kColVec = true;
    static constexpr int kScratchPerWarp = 4096;
    struct Params {
        CUtensorMap tmap_p;   // P (first column = the range's first label column), box {64 columns, 32 rows}
        const float* colvec;  // bias [Vt], -inf beyond V (already offset to the range's first column)
        const int* rows; const long long* labels;
        const float* lse; const float* grad_scale;
        float inv_count;
        const float* pm; const float* ps;   // when not NULL: the forward's LSE partials [slots, Mpad]; the log-sum-exp
        int slots, Mpad;                    // is rebuilt from them instead of being read from `lse`
        __nv_bfloat16* p;      // already offset to the first label column of the range
        int M, Vp;             // Vp: label columns of the range to write (multiple of 64)
        int p_pitch;           // row pitch of P (the full padded label count)
        int col_base;          // global index of the range's first label column
        int V, packed;
    };
    struct State { float lse2, scale; int label; };
    static __device__ __forceinline__ void begin_rb(State& st, const Params& ep, const TileCtx& cx) {
        const bool ok = cx.row < ep.M;
        st.label = masked_label(ep.labels, ep.rows, cx.row, ep.M, ep.V, ep.packed);
        if (ep.pm) {
            // The forward's per-worker partials (max, sum) of this row are combined here, in slot order and with the same
            // arithmetic for every CTA that owns the row, instead of waiting for ce_finalize_kernel.  Only the half-0
            // thread of a row does it (all loads of a batch of 16 slots in flight at once) and hands the result to its
            // half-1 twin through shared memory; the rebuild hides behind the first accumulator tile of the row block.
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            float* xch = reinterpret_cast<float*>(cx.scratch - warp * kScratchPerWarp);      // warp 0's staging tile: free here
            const int r = (warp & 3) * 32 + lane;
            if (cx.half == 0) {
                float l2 = CUDART_INF_F;
                if (ok) {
                    float mx = -CUDART_INF_F, sum = 0.f;
                    for (int s0 = 0; s0 < ep.slots; s0 += 16) {
                        float pmv[16], psv[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const bool in = s0 + j < ep.slots;
                            pmv[j] = in ? __ldg(ep.pm + (size_t)(s0 + j) * ep.Mpad + cx.row) : -CUDART_INF_F;
                            psv[j] = in ? __ldg(ep.ps + (size_t)(s0 + j) * ep.Mpad + cx.row) : 0.f;
                        }
                        float bm = mx;
#pragma unroll
                        for (int j = 0; j < 16; ++j) bm = fmaxf(bm, pmv[j]);
                        if (bm > -CUDART_INF_F) {
                            const float bm2 = bm * kLog2e;
                            sum *= ex2_fast(fmaf(mx, kLog2e, -bm2));                            // 2^-inf = 0 on the first batch
#pragma unroll
                            for (int j = 0; j < 16; ++j) sum = fmaf(psv[j], ex2_fast(fmaf(pmv[j], kLog2e, -bm2)), sum);
                            mx = bm;
                        }
                    }
                    l2 = fmaf(mx, kLog2e, log2f(sum));
                }
                xch[r] = l2;
                st.lse2 = l2;
            }
            asm volatile("bar.sync 3, 256;" ::: "memory");
            if (cx.half == 1) st.lse2 = xch[r];
            asm volatile("bar.sync 3, 256;" ::: "memory");        // before the staging tile is used for stores again
        } else {
            st.lse2 = ok ? __ldg(ep.lse + cx.row) * kLog2e : CUDART_INF_F;
        }
        st.scale = ok ? ep.inv_count * (ep.grad_scale ? __ldg(ep.grad_scale) : 1.0f) : 0.f;
    }
    static __device__ __forceinline__ void tile(State& st, const Params& ep, const TileCtx& cx, uint32_t taddr) {
        const float4* cv = reinterpret_cast<const float4*>(cx.cv);
        const int lane = threadIdx.x & 31;
        const int row_base = cx.row - lane;
        const uint32_t stage = smem_u32(cx.scratch);
        for_each_chunk(taddr, [&](int c, const uint32_t (&r)[32]) {
            const int col = cx.col0 + c * 32;
            if (col >= ep.Vp) return;                           // warp-uniform: Vp is a multiple of 64
            float g[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 b = cv[c * 8 + i];
                g[4 * i + 0] = ex2_fast(fmaf(__uint_as_float(r[4 * i + 0]) + b.x, kLog2e, -st.lse2)) * st.scale;
                g[4 * i + 1] = ex2_fast(fmaf(__uint_as_float(r[4 * i + 1]) + b.y, kLog2e, -st.lse2)) * st.scale;
                g[4 * i + 2] = ex2_fast(fmaf(__uint_as_float(r[4 * i + 2]) + b.z, kLog2e, -st.lse2)) * st.scale;
                g[4 * i + 3] = ex2_fast(fmaf(__uint_as_float(r[4 * i + 3]) + b.w, kLog2e, -st.lse2)) * st.scale;
            }
            const unsigned rel = (unsigned)(st.label - ep.col_base - col);    // (p - 1) * scale at the label column
            if (rel < 32u) {
#pragma unroll
                for (int j = 0; j < 32; ++j) if (rel == (unsigned)j) g[j] -= st.scale;
            }
            __nv_bfloat162 o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __floats2bfloat162_rn(g[2 * j], g[2 * j + 1]);
            if ((c & 1) == 0) {                             // the previous store must have read the tile
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
            }
            const uint32_t rowaddr = stage + (uint32_t)lane * 128u;
            const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t piece = (uint32_t)((c & 1) * 4 + i) ^ sw;
                const uint4 v = *reinterpret_cast<uint4*>(&o[4 * i]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                             ::"r"(rowaddr + piece * 16u), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
            }
            if (c & 1) {
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&ep.tmap_p, stage, col - 32, row_base);
                    tma_store_commit();
                }
            }
        });
    }
    static __device__ __forceinline__ void end_rb(State&, const Params&, const TileCtx&) {
        // the staging tile must stay valid until the last store has read it (the CTA may exit right after this)
        if ((threadIdx.x & 31) == 0) tma_store_wait_read();
    }
};

// ------------------------------------------------------------------------------------------------ small kernels
// lse[m] = log-sum-exp combined over the column-split partials; rowloss[m] = lse[m] - z[label].
// The block that finishes last (ticket) sums rowloss in a fixed order into loss_sum: no second launch, and the
// result does not depend on which block that is.
__global__ void __launch_bounds__(256)
ce_finalize_kernel(const float* __restrict__ pm, const float* __restrict__ ps, const float* __restrict__ zlab, int M, int Mpad,
                   int slots, float* __restrict__ lse, float* __restrict__ rowloss, unsigned int* __restrict__ ticket,
                   float* __restrict__ loss_sum) {
    // 32 rows per block; thread (r = tid % 32, g = tid / 32) combines slots g, g + 8, ... of row r: every load
    // instruction of a warp reads 32 consecutive rows of one slot (one 128-byte line), then the 8 groups are merged
    // through shared memory in a fixed order.
    // a dlogits GEMM launched right behind (backward on the forward's workspace) recomputes the log-sum-exp from the
    // same partials and never reads this kernel's output: it may start at once
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ float gm[8][33], gs[8][33];
    const int r = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int m = blockIdx.x * 32 + r;
    const int lane = r;
    float mx = -CUDART_INF_F, sum = 0.f;
    if (m < M) {
        for (int s = g; s < slots; s += 8) {
            const float pmv = __ldg(pm + (size_t)s * Mpad + m), psv = __ldg(ps + (size_t)s * Mpad + m);
            const float nm = fmaxf(mx, pmv);
            if (nm > -CUDART_INF_F) sum = sum * exp2f((mx - nm) * kLog2e) + psv * exp2f((pmv - nm) * kLog2e);
            mx = nm;
        }
    }
    gm[g][r] = mx; gs[g][r] = sum;
    __syncthreads();
    if (g == 0 && m < M) {
        float tm = gm[0][r];
#pragma unroll
        for (int k = 1; k < 8; ++k) tm = fmaxf(tm, gm[k][r]);
        float ts = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) ts += gs[k][r] * exp2f((gm[k][r] - tm) * kLog2e);      // exp2(-inf) = 0 on empty groups
        const float l = tm + log2f(ts) * kLn2;
        lse[m] = l;
        rowloss[m] = l - zlab[m];
    }
    __shared__ bool last;
    __shared__ float sh[8];
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    float part = 0.f;
    for (int i = threadIdx.x; i < M; i += blockDim.x) part += __ldcg(rowloss + i);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sh[w];
        loss_sum[0] = t;
    }
}

// Evaluation: zl[m] = <A[m, :], W[label_m, :]> + bias[label_m] from the same bf16 operands the GEMM reads
// (fp32 accumulation); one warp per masked row.
__global__ void __launch_bounds__(256)
ce_label_logit_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ wb, const float* __restrict__ bias,
                      const int* __restrict__ rows, const long long* __restrict__ labels, int M, int V, int Dhp, int packed,
                      float* __restrict__ zl) {
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const int label = masked_label(labels, rows, m, M, V, packed);
    if (label < 0) { if (lane == 0) zl[m] = CUDART_NAN_F; return; }      // outside [0, V): never index W with it
    const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(a + (size_t)m * Dhp);
    const __nv_bfloat162* w = reinterpret_cast<const __nv_bfloat162*>(wb + (size_t)label * Dhp);
    float acc = 0.f;
    for (int i = lane; i < Dhp / 2; i += 32) {
        const float2 xv = __bfloat1622float2(x[i]), wv = __bfloat1622float2(w[i]);
        acc = fmaf(xv.x, wv.x, acc);
        acc = fmaf(xv.y, wv.y, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) zl[m] = acc + __ldg(bias + label);
}

struct TopKs { int k[8]; int n; };
// rank[m] = number of labels with a larger logit than the frame's own label (sum of the per-split counts);
// errors[i] += [rank[m] >= k_i]  (integer atomics: exact, order-free).
__global__ void __launch_bounds__(256)
ce_rank_finalize_kernel(const int* __restrict__ pcnt, int M, int Mpad, int slots, TopKs ks, int* __restrict__ rank,
                        unsigned long long* __restrict__ errors) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    int r = 0;
    if (m < M) {
        for (int s = 0; s < slots; ++s) r += __ldg(pcnt + (size_t)s * Mpad + m);
        if (rank) rank[m] = r;
    }
    for (int i = 0; i < ks.n; ++i) {
        const unsigned ballot = __ballot_sync(0xffffffffu, m < M && r >= ks.k[i]);
        if ((threadIdx.x & 31) == 0 && ballot) atomicAdd(errors + i, (unsigned long long)__popc(ballot));
    }
}

__global__ void ce_sum_kernel(const float* __restrict__ v, int M, float* __restrict__ out);

// d_b[v] = sum_m P[m, v] in two passes that both stream contiguous memory:
//   ce_db_partial_kernel   block r adds the kDbRows consecutive rows r*8 .. r*8+7 of P (whole rows: 16-byte loads,
//                          every warp instruction reads 512 contiguous bytes) into part[r, :]
//   ce_db_reduce_block     d_b[v] = sum_r part[r, v] in a fixed order, 32 labels per block
// Fixed orders: bit-identical from run to run.
__global__ void __launch_bounds__(256)
ce_db_partial_kernel(const __nv_bfloat16* __restrict__ p, int M, int p_pitch, int v_begin, int v_len8, int vp,
                     float* __restrict__ part, int wait_for_previous, uint4* __restrict__ zero_fill, long long zero_vec) {
    // side-by-side schedule: this kernel also clears d_h (zero_vec 16-byte words) while the gradient GEMMs run, so
    // that the scatter behind them only has to write the masked frames
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // a programmatic successor waits for this grid itself
    if (zero_fill) {
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < zero_vec; i += stride)
            zero_fill[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const int m0 = blockIdx.x * kDbRows;
    float* dst = part + (size_t)blockIdx.x * vp + v_begin;
    for (int c = threadIdx.x; c < v_len8; c += blockDim.x) {        // 8 columns per thread
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
        for (int r = 0; r < kDbRows; ++r) {
            if (m0 + r < M) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4*>(p + (size_t)(m0 + r) * p_pitch + v_begin) + c);
                const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __bfloat1622float2(h2[j]);
                    acc[2 * j] += f.x; acc[2 * j + 1] += f.y;
                }
            }
        }
        float4* o = reinterpret_cast<float4*>(dst + 8 * c);
        o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    // side-by-side schedule: this kernel was released early by the d_h GEMM in front of it and must not be seen to
    // finish before that one (see pero_masked_ce_bwd_range)
    if (wait_for_previous) asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Backward on what the forward left in P (PERO_CE_KEEP_LOGITS: e = exp(z - cmax) per 32-column chunk, cm = the maxima):
//   P[m, v] <- e[m, v] * exp(cm[chunk(v), m] - lse_m) * scale - [v == label_m] * scale      in place,
// and the d_b partial column sums of the same 8 rows in the same pass (fp32 values, before the bf16 rounding).
// One block = 8 masked rows (one row group of the d_b partials) x 2048 label columns (256 threads x 8 columns: 8
// independent 16-byte loads in flight per thread, every warp instruction touches 512 contiguous bytes; the 8 columns of a
// thread lie in one chunk, whose maxima for the 8 rows are 32 contiguous bytes).  One exponential per (row, thread), three
// instructions per element, one round trip to memory per block; three blocks per SM are resident and the hardware
// back-fills the rest.  The log-sum-exp of a row is rebuilt from the forward's per-worker partials by one warp (slots
// in lane order, fixed shuffle tree: bit-identical from run to run and for every column range) while the loads are in
// flight.
__global__ void __launch_bounds__(256, 3)
ce_dlogits_inplace_kernel(__nv_bfloat16* __restrict__ p, int M, int p_pitch, int v_begin, int v_len8, int vp,
                          const float* __restrict__ pm, const float* __restrict__ ps, int slots, int Mpad,
                          const float* __restrict__ cm, const int* __restrict__ rows, const long long* __restrict__ labels,
                          int V, int packed, const float* __restrict__ grad_scale, float inv_count, float* __restrict__ part,
                          int wait_for_previous, uint4* __restrict__ zero_fill, long long zero_vec) {
    // launched programmatically behind the forward's GEMM (or its finalize): everything read here is theirs
    if (wait_for_previous) asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");    // the gradient GEMMs set themselves up meanwhile
    __shared__ float s_lse2[kDbRows];
    __shared__ int s_label[kDbRows];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float scale = inv_count * (grad_scale ? __ldg(grad_scale) : 1.0f);
    const int groups = (M + kDbRows - 1) / kDbRows;
    // consecutive blocks are consecutive row groups of one column chunk
    const int chunk = blockIdx.x / groups, grp = blockIdx.x - chunk * groups;
    const int m0 = grp * kDbRows;
    const int c = chunk * 256 + (int)threadIdx.x;
    const bool c_ok = c < v_len8;
    uint4 raw[kDbRows];
#pragma unroll
    for (int r = 0; r < kDbRows; ++r)
        raw[r] = (m0 + r < M && c_ok) ? *(reinterpret_cast<const uint4*>(p + (size_t)(m0 + r) * p_pitch + v_begin) + c)
                                      : make_uint4(0u, 0u, 0u, 0u);
    // chunk maxima of the 8 rows (m0 is a multiple of 8 and Mpad of 256: rows beyond M exist in the array)
    float cmx[kDbRows];
    {
        const float4* cp = reinterpret_cast<const float4*>(cm + (size_t)((v_begin + 8 * (c_ok ? c : 0)) >> 5) * Mpad + m0);
        const float4 u = __ldg(cp), v = __ldg(cp + 1);
        cmx[0] = u.x; cmx[1] = u.y; cmx[2] = u.z; cmx[3] = u.w; cmx[4] = v.x; cmx[5] = v.y; cmx[6] = v.z; cmx[7] = v.w;
    }
    {
        const int m = m0 + warp;                   // 8 warps <-> 8 rows
        float l2 = CUDART_INF_F;
        if (m < M) {
            float mx = -CUDART_INF_F, sum = 0.f;
            for (int s0 = 0; s0 < slots; s0 += 32) {
                const bool in = s0 + lane < slots;
                const float pmv = in ? __ldg(pm + (size_t)(s0 + lane) * Mpad + m) : -CUDART_INF_F;
                const float psv = in ? __ldg(ps + (size_t)(s0 + lane) * Mpad + m) : 0.f;
                float bm = pmv;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
                bm = fmaxf(bm, mx);
                float t = (bm > -CUDART_INF_F) ? psv * ex2_fast((pmv - bm) * kLog2e) : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                sum = ((bm > -CUDART_INF_F) ? sum * ex2_fast((mx - bm) * kLog2e) : 0.f) + t;
                mx = bm;
            }
            l2 = fmaf(mx, kLog2e, log2f(sum));
        }
        if (lane == 0) {
            s_lse2[warp] = l2;
            s_label[warp] = masked_label(labels, rows, m, M, V, packed);
        }
    }
    __syncthreads();
    // every block also clears its share of d_h (zero_vec 16-byte words: the scatter behind the gradient GEMMs then writes
    // the masked frames only); plain stores, nothing waits for them here
    if (zero_fill) {
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < zero_vec; i += stride)
            zero_fill[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (!c_ok) return;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int col = v_begin + 8 * c;
    uint4* prow = reinterpret_cast<uint4*>(p + (size_t)m0 * p_pitch + v_begin) + c;
    const int rows_here = min(kDbRows, M - m0);
#pragma unroll
    for (int r = 0; r < kDbRows; ++r) {
        if (r < rows_here) {
            // exp(cm - lse) * scale: the softmax factor of this (row, chunk); exp2(-inf) = 0 on padding chunks
            const float f = ex2_fast(fmaf(cmx[r], kLog2e, -s_lse2[r])) * scale;
            const uint32_t w[4] = {raw[r].x, raw[r].y, raw[r].z, raw[r].w};
            float g[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) {           // bf16 pair -> fp32: low half shifted up, high half masked
                g[2 * j] = __uint_as_float(w[j] << 16) * f;
                g[2 * j + 1] = __uint_as_float(w[j] & 0xffff0000u) * f;
            }
            const unsigned rel = (unsigned)(s_label[r] - col);
            if (rel < 8u) {                         // the row's label lies in this thread's 8 columns (one thread in 256 per row)
                switch (rel) {
                    case 0: g[0] -= scale; break;
                    case 1: g[1] -= scale; break;
                    case 2: g[2] -= scale; break;
                    case 3: g[3] -= scale; break;
                    case 4: g[4] -= scale; break;
                    case 5: g[5] -= scale; break;
                    case 6: g[6] -= scale; break;
                    default: g[7] -= scale; break;
                }
            }
            uint4 o;
            __nv_bfloat162 t;
            t = __floats2bfloat162_rn(g[0], g[1]); o.x = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(g[2], g[3]); o.y = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(g[4], g[5]); o.z = *reinterpret_cast<uint32_t*>(&t);
            t = __floats2bfloat162_rn(g[6], g[7]); o.w = *reinterpret_cast<uint32_t*>(&t);
            *prow = o;
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += g[j];
        }
        prow = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(prow) + p_pitch);
    }
    float4* o4 = reinterpret_cast<float4*>(part + (size_t)grp * vp + v_begin + 8 * c);
    o4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
}

// One block finishes 32 labels: thread (c = tid % 32, g = tid / 32) adds the partial rows g, g + 8, ... of label
// v0 + c (all loads independent: one round trip), then the 8 group sums are added in ascending g.
__device__ __forceinline__ void ce_db_reduce_block(const float* __restrict__ part, int nparts, int vp, int v0, int v_end,
                                                   float* __restrict__ db) {
    __shared__ float grp[8][33];
    const int c = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int v = v0 + c;
    float t = 0.f;
    if (v < v_end) {
#pragma unroll 8
        for (int r = g; r < nparts; r += 8) t += __ldg(part + (size_t)r * vp + v);
    }
    grp[g][c] = t;
    __syncthreads();
    if (g == 0 && v < v_end) {
        float sum = grp[0][c];
#pragma unroll
        for (int k = 1; k < 8; ++k) sum += grp[k][c];
        db[v] = sum;
    }
}

__global__ void __launch_bounds__(256)
ce_db_reduce_kernel(const float* __restrict__ part, int nparts, int vp, int v_begin, int v_end, float* __restrict__ db) {
    ce_db_reduce_block(part, nparts, vp, v_begin + (int)blockIdx.x * 32, v_end, db);
}

// The first db_blocks blocks (if any) finish d_b from its partial sums (ce_db_reduce_block), the others scatter.
template <typename T>
__global__ void __launch_bounds__(256)
ce_dh_scatter_kernel(const float* __restrict__ planes, const int* __restrict__ inv, const int* __restrict__ rows, long long N,
                     int M, int Dh, int KS, T* __restrict__ dh, int db_blocks, const float* __restrict__ dbpart, int nparts,
                     int vp, int V, float* __restrict__ db, int masked_only) {
    if ((int)blockIdx.x < db_blocks) {
        ce_db_reduce_block(dbpart, nparts, vp, (int)blockIdx.x * 32, V, db);
        return;
    }
    const int scatter_blocks = (int)gridDim.x - db_blocks;
    const int sblock = (int)blockIdx.x - db_blocks;
    // one thread per 4 consecutive channels (Dh % 4 == 0): 16-byte plane reads, 16/8-byte stores
    const int g4 = Dh >> 2;
    const long long total = (masked_only ? (long long)M : N) * g4;
    const long long stride = (long long)scatter_blocks * blockDim.x;
    for (long long j = (long long)sblock * blockDim.x + threadIdx.x; j < total; j += stride) {
        // masked_only: d_h was cleared beforehand and only the rows of the masked frames are written
        long long n = j / g4;
        const int g = (int)(j - n * g4);
        int m;
        if (masked_only) { m = (int)n; n = __ldg(rows + m); } else { m = masked_row_of(inv, rows, n, M); }
        const long long i = n * g4 + g;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m >= 0) {
            // all plane loads of a group of 8 are in flight before the (fixed-order) adds
            for (int k0 = 0; k0 < KS; k0 += 8) {
                float4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    v[k] = (k0 + k < KS) ? __ldg(reinterpret_cast<const float4*>(planes + ((size_t)(k0 + k) * M + m) * Dh) + g)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (k0 + k < KS) { s.x += v[k].x; s.y += v[k].y; s.z += v[k].z; s.w += v[k].w; }
                }
            }
        }
        if constexpr (sizeof(T) == 4) {
            reinterpret_cast<float4*>(dh)[i] = s;
        } else {
            __nv_bfloat162 lo = __floats2bfloat162_rn(s.x, s.y), hi = __floats2bfloat162_rn(s.z, s.w);
            uint2 o;
            o.x = *reinterpret_cast<uint32_t*>(&lo); o.y = *reinterpret_cast<uint32_t*>(&hi);
            reinterpret_cast<uint2*>(dh)[i] = o;
        }
    }
}


// ------------------------------------------------------------------------------------------------ logits-in CE
// MaskedCrossEntropyLoss.forward(output, labels, mask) for callers that already hold the logits
// (masked_pretraining/model.py:78-82): one warp per selected row, online log-sum-exp, 16-byte loads.
template <typename T>
__global__ void __launch_bounds__(256)
ce_logits_fwd_kernel(const T* __restrict__ logits, const int* __restrict__ rows, const long long* __restrict__ labels, int M, int V,
                     float* __restrict__ lse, float* __restrict__ rowloss) {
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const int r = __ldg(rows + m);
    const T* z = logits + (size_t)r * V;
    float mx = -CUDART_INF_F, s = 0.f;
    for (int v = lane; v < V; v += 32) {
        const float x = (float)z[v];
        if (x > mx) { s = s * exp2f((mx - x) * kLog2e); mx = x; }
        s += exp2f((x - mx) * kLog2e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, s, o);
        const float nm = fmaxf(mx, om);
        s = (nm == -CUDART_INF_F) ? 0.f : s * exp2f((mx - nm) * kLog2e) + os * exp2f((om - nm) * kLog2e);
        mx = nm;
    }
    if (lane == 0) {
        const float l = mx + log2f(s) * kLn2;
        lse[m] = l;
        rowloss[m] = l - (float)z[__ldg(labels + r)];
    }
}

__global__ void __launch_bounds__(1024)
ce_sum_kernel(const float* __restrict__ v, int M, float* __restrict__ out) {
    __shared__ float sh[32];
    float part = 0.f;
    for (int m = threadIdx.x; m < M; m += blockDim.x) part += v[m];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = sh[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) out[0] = t;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
ce_logits_bwd_kernel(const T* __restrict__ logits, const int* __restrict__ rows, const long long* __restrict__ labels,
                     const float* __restrict__ lse, const float* __restrict__ grad_scale, float inv_count, int M, int V,
                     T* __restrict__ d_logits) {
    const int m = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (m >= M) return;
    const int r = __ldg(rows + m);
    const int label = (int)__ldg(labels + r);
    const float l2 = __ldg(lse + m) * kLog2e;
    const float scale = inv_count * (grad_scale ? __ldg(grad_scale) : 1.0f);
    const T* z = logits + (size_t)r * V;
    T* d = d_logits + (size_t)r * V;
    for (int v = lane; v < V; v += 32) {
        float p = exp2f(fmaf((float)z[v], kLog2e, -l2));
        if (v == label) p -= 1.0f;
        d[v] = (T)(p * scale);
    }
}

struct MaskPred {
    const void* mask; int dtype; int want; const long long* labels;
    __device__ __forceinline__ bool operator()(int i) const {
        long long v;
        if (dtype == 0) v = static_cast<const long long*>(mask)[i];
        else if (dtype == 1) v = static_cast<const int*>(mask)[i];
        else v = static_cast<const unsigned char*>(mask)[i];
        return v == want && (labels == nullptr || labels[i] >= 0);
    }
};

template <typename T>
int launch_ce_gather(const void* h, const int32_t* rows, int M, int Dh, const CeWsLayout& l, char* ws, cudaStream_t stream) {
    ce_gather_kernel<T><<<(unsigned)((M + 7) / 8), 256, 0, stream>>>(
        static_cast<const T*>(h), rows, M, Dh, (int)l.Dhp, reinterpret_cast<__nv_bfloat16*>(ws + l.a_off),
        reinterpret_cast<int*>(ws + l.inv_off), reinterpret_cast<unsigned int*>(ws + l.ticket_off));
    return (int)cudaGetLastError();
}
inline int ce_gather(const void* h, int h_is_bf16, const int32_t* rows, int M, int Dh, const CeWsLayout& l, char* ws,
                     cudaStream_t stream) {
    return h_is_bf16 ? launch_ce_gather<__nv_bfloat16>(h, rows, M, Dh, l, ws, stream)
                     : launch_ce_gather<float>(h, rows, M, Dh, l, ws, stream);
}

// The logits GEMMs keep the masked rows resident in shared memory while the label tiles stream by when the hidden
// dimension allows it (Dh <= 512: 8 k-blocks = 128 KB); wider heads stream both operands through the ring.
template <class Epi>
int launch_logits_gemm(const void* a, int M, int Dhp, const void* w, int vlen, int split_mode, int fixed_s,
                       const typename Epi::Params& ep, cudaStream_t st, size_t budget, int pdl) {
    if (Dhp / kBlockK <= 8)
        return launch_gemm_tn<2, 1, Epi>(a, M, Dhp, w, vlen, Dhp, Dhp, 1, split_mode, fixed_s, 0, ep, st, nullptr, budget, 0, pdl);
    return launch_gemm_tn<2, 0, Epi>(a, M, Dhp, w, vlen, Dhp, Dhp, 1, split_mode, fixed_s, 0, ep, st, nullptr, budget, 0, pdl);
}


}  // namespace pero

using namespace pero;

extern "C" {

size_t pero_head_bytes(int64_t V, int64_t Dh) {
    if (V <= 0 || Dh <= 0) return 0;
    return head_layout(V, Dh).total;
}

int pero_head_prepare(const float* W, const float* bias, int64_t V, int64_t Dh, void* head, size_t head_bytes,
                      pero_stream_t stream) {
    if (!W || !head) return PERO_ERR_NULL;
    if (V <= 0 || Dh <= 0 || V > (1ll << 24) || Dh > 65536) return PERO_ERR_BAD_SHAPE;
    const HeadLayout l = head_layout(V, Dh);
    if (head_bytes < l.total) return PERO_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(head) & 255) return PERO_ERR_BAD_ALIGN;
    char* base = static_cast<char*>(head);
    long long blocks = ((long long)V * (l.Dhp / 4) + 1023) / 1024;
    if (blocks > 148 * 2) blocks = 148 * 2;
    if (blocks < 1) blocks = 1;
    head_prepare_kernel<<<(unsigned)blocks, 256, 0, stream>>>(W, bias, (int)V, (int)Dh, (int)l.Dhp, (int)l.Vt,
                                                            reinterpret_cast<__nv_bfloat16*>(base + l.w_off),
                                                            reinterpret_cast<float*>(base + l.bias_off));
    return (int)cudaGetLastError();
}

size_t pero_masked_ce_workspace_bytes(int64_t N, int64_t M, int64_t V, int64_t Dh) {
    if (N <= 0 || M <= 0 || V <= 0 || Dh <= 0) return 0;
    return ce_ws_layout(N, M, V, Dh).total;
}

static int ce_check(const void* h, int64_t N, int64_t Dh, const int32_t* rows, int64_t M, const int64_t* labels,
                    const void* head, int64_t V, void* workspace, size_t workspace_bytes, bool h_optional = false) {
    if ((!h && !h_optional) || !rows || !labels || !head || !workspace) return PERO_ERR_NULL;
    if (N <= 0 || M <= 0 || M > N || V <= 0 || Dh <= 0 || N > (1ll << 31) - 256 || V > (1ll << 24) || Dh > 16384)
        return PERO_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(head) & 255)) return PERO_ERR_BAD_ALIGN;
    if (workspace_bytes < ce_ws_layout(N, M, V, Dh).total) return PERO_ERR_WORKSPACE;
    return PERO_OK;
}

int pero_masked_ce_gather(const void* h, int h_is_bf16, int64_t N, int64_t Dh, const int32_t* rows, int64_t M, int64_t V,
                          void* workspace, size_t workspace_bytes, pero_stream_t stream) {
    if (M == 0) return PERO_ERR_BAD_SHAPE;
    if (!h || !rows || !workspace) return PERO_ERR_NULL;
    if (N <= 0 || M < 0 || M > N || V <= 0 || Dh <= 0 || N > (1ll << 31) - 256 || Dh > 16384) return PERO_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return PERO_ERR_BAD_ALIGN;
    const CeWsLayout l = ce_ws_layout(N, M, V, Dh);
    if (workspace_bytes < l.total) return PERO_ERR_WORKSPACE;
    return ce_gather(h, h_is_bf16, rows, (int)M, (int)Dh, l, static_cast<char*>(workspace), reinterpret_cast<cudaStream_t>(stream));
}

int pero_masked_ce_fwd(const void* h, int flags, int64_t N, int64_t Dh, const int32_t* rows, int64_t M,
                       const int64_t* labels, const void* head, int64_t V, float* loss_sum, float* lse,
                       void* workspace, size_t workspace_bytes, pero_stream_t stream) {
    if (M == 0) return PERO_ERR_BAD_SHAPE;   // the reference returns NaN on an empty mask; the host wrapper handles it
    const int h_is_bf16 = flags & PERO_CE_H_BF16, packed = (flags & PERO_CE_LABELS_PACKED) ? 1 : 0;
    int rc = ce_check(h, N, Dh, rows, M, labels, head, V, workspace, workspace_bytes, /*h_optional=*/true);
    if (rc) return rc;
    if ((loss_sum == nullptr) != (lse == nullptr)) return PERO_ERR_NULL;      // both (finalize now) or neither (pero_masked_ce_loss later)
    const CeWsLayout l = ce_ws_layout(N, M, V, Dh);
    const HeadLayout hl = head_layout(V, Dh);
    char* ws = static_cast<char*>(workspace);
    const char* hb = static_cast<const char*>(head);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // The gather leaves the bf16 operand A and the inverse row map in the workspace: a backward call that is handed
    // the same workspace (h = NULL) starts directly with its GEMM.  h == NULL here: pero_masked_ce_gather already ran
    // on this workspace for the same (h, rows).
    if (h) {
        rc = ce_gather(h, h_is_bf16, rows, (int)M, (int)Dh, l, ws, st);
        if (rc) return rc;
    }
    auto fill = [&](auto& ep) {
        ep.colvec = reinterpret_cast<const float*>(hb + hl.bias_off);
        ep.rows = rows; ep.labels = reinterpret_cast<const long long*>(labels);
        ep.pm = reinterpret_cast<float*>(ws + l.pm_off);
        ep.ps = reinterpret_cast<float*>(ws + l.ps_off);
        ep.zlab = reinterpret_cast<float*>(ws + l.zlab_off);
        ep.zl_in = nullptr; ep.pcnt = nullptr;
        ep.M = (int)M; ep.Mpad = (int)l.Mpad; ep.S = (int)l.S; ep.V = (int)V; ep.packed = packed; ep.Vp = (int)l.Vp;
        ep.cm = reinterpret_cast<float*>(ws + l.cm_off);
    };
    // behind its own gather: the set-up overlaps the gather, the first load waits for it (programmatic launch); the
    // kernel behind is released when the last operand load of a CTA has been requested (bit 4)
    float* pm = reinterpret_cast<float*>(ws + l.pm_off);
    float* ps = reinterpret_cast<float*>(ws + l.ps_off);
    float* zlab = reinterpret_cast<float*>(ws + l.zlab_off);
    if (flags & PERO_CE_KEEP_LOGITS) {
        // training forward: the bf16 logits stay in the workspace (P area); the backward converts them in place
        LseStoreEpi::Params ep;
        fill(ep);
        rc = make_tmap_bf16(&ep.tmap_p, ws + l.p_off, (uint64_t)M, (uint64_t)l.Vp, (uint64_t)l.Pp, 32);
        if (rc) return rc;
        rc = launch_logits_gemm<LseStoreEpi>(ws + l.a_off, (int)M, (int)l.Dhp, hb + hl.w_off, (int)V, /*split_mode=*/1, (int)l.S, ep,
                                             st, kSmemBudget, (h ? 4 : 0) | 16);
    } else {
        LseEpi::Params ep;
        fill(ep);
        rc = launch_logits_gemm<LseEpi>(ws + l.a_off, (int)M, (int)l.Dhp, hb + hl.w_off, (int)V, /*split_mode=*/1, (int)l.S, ep, st,
                                        kSmemBudgetShared, (h ? 4 : 0) | 16);
    }
    if (rc || !lse) return rc;
    // lse / rowloss / loss_sum from the partials.  A backward GEMM launched right behind on the same workspace rebuilds
    // the log-sum-exp from the partials itself and never reads this kernel's output.
    float* rowloss = reinterpret_cast<float*>(ws + l.rowloss_off);
    ce_finalize_kernel<<<(unsigned)((M + 31) / 32), 256, 0, st>>>(pm, ps, zlab, (int)M, (int)l.Mpad, (int)l.S, lse,
                                                              rowloss, reinterpret_cast<unsigned int*>(ws + l.ticket_off),
                                                              loss_sum);
    return (int)cudaGetLastError();
}

int pero_masked_ce_loss(int64_t N, int64_t Dh, int64_t M, int64_t V, float* loss_sum, float* lse, void* workspace,
                        size_t workspace_bytes, pero_stream_t stream) {
    if (!loss_sum || !lse || !workspace) return PERO_ERR_NULL;
    if (N <= 0 || M <= 0 || M > N || V <= 0 || Dh <= 0) return PERO_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return PERO_ERR_BAD_ALIGN;
    const CeWsLayout l = ce_ws_layout(N, M, V, Dh);
    if (workspace_bytes < l.total) return PERO_ERR_WORKSPACE;
    char* ws = static_cast<char*>(workspace);
    ce_finalize_kernel<<<(unsigned)((M + 31) / 32), 256, 0, stream>>>(
        reinterpret_cast<const float*>(ws + l.pm_off), reinterpret_cast<const float*>(ws + l.ps_off),
        reinterpret_cast<const float*>(ws + l.zlab_off), (int)M, (int)l.Mpad, (int)l.S, lse,
        reinterpret_cast<float*>(ws + l.rowloss_off), reinterpret_cast<unsigned int*>(ws + l.ticket_off), loss_sum);
    return (int)cudaGetLastError();
}

int pero_masked_ce_eval(const void* h, int flags, int64_t N, int64_t Dh, const int32_t* rows, int64_t M,
                        const int64_t* labels, const void* head, int64_t V, const int32_t* topk_host, int num_topk,
                        float* loss_sum, float* lse, int32_t* rank, int64_t* errors, void* workspace, size_t workspace_bytes,
                        pero_stream_t stream) {
    if (M == 0) return PERO_ERR_BAD_SHAPE;
    const int h_is_bf16 = flags & PERO_CE_H_BF16;
    int rc = ce_check(h, N, Dh, rows, M, labels, head, V, workspace, workspace_bytes);
    if (rc) return rc;
    if (!loss_sum || !lse || !errors || !topk_host) return PERO_ERR_NULL;
    if (num_topk < 1 || num_topk > 8) return PERO_ERR_BAD_SHAPE;
    TopKs ks;
    ks.n = num_topk;
    for (int i = 0; i < 8; ++i) ks.k[i] = i < num_topk ? topk_host[i] : 0;
    for (int i = 0; i < num_topk; ++i) if (ks.k[i] < 1) return PERO_ERR_BAD_SHAPE;
    const CeWsLayout l = ce_ws_layout(N, M, V, Dh);
    const HeadLayout hl = head_layout(V, Dh);
    char* ws = static_cast<char*>(workspace);
    const char* hb = static_cast<const char*>(head);
    cudaError_t e = cudaMemsetAsync(errors, 0, (size_t)num_topk * 8, stream);
    if (e != cudaSuccess) return (int)e;
    rc = ce_gather(h, h_is_bf16, rows, (int)M, (int)Dh, l, ws, stream);
    if (rc) return rc;
    EvalEpi::Params ep;
    ep.colvec = reinterpret_cast<const float*>(hb + hl.bias_off);
    ep.rows = rows; ep.labels = reinterpret_cast<const long long*>(labels); ep.V = (int)V;
    ep.pm = reinterpret_cast<float*>(ws + l.pm_off);
    ep.ps = reinterpret_cast<float*>(ws + l.ps_off);
    ep.zlab = reinterpret_cast<float*>(ws + l.zlab_off);
    float* zl_in = reinterpret_cast<float*>(ws + l.zlin_off);
    ep.zl_in = zl_in;
    ep.pcnt = reinterpret_cast<int*>(ws + l.pcnt_off);
    ep.M = (int)M; ep.Mpad = (int)l.Mpad; ep.S = (int)l.S; ep.packed = (flags & PERO_CE_LABELS_PACKED) ? 1 : 0;
    ce_label_logit_kernel<<<(unsigned)((M + 7) / 8), 256, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(ws + l.a_off), reinterpret_cast<const __nv_bfloat16*>(hb + hl.w_off), ep.colvec,
        rows, ep.labels, (int)M, (int)V, (int)l.Dhp, ep.packed, zl_in);
    rc = launch_logits_gemm<EvalEpi>(ws + l.a_off, (int)M, (int)l.Dhp, hb + hl.w_off, (int)V, /*split_mode=*/1, (int)l.S, ep, stream,
                                     kSmemBudgetShared, /*pdl=*/0);
    if (rc) return rc;
    float* rowloss = reinterpret_cast<float*>(ws + l.rowloss_off);
    ce_finalize_kernel<<<(unsigned)((M + 31) / 32), 256, 0, stream>>>(ep.pm, ep.ps, ep.zlab, (int)M, (int)l.Mpad, (int)l.S, lse,
                                                                  rowloss, reinterpret_cast<unsigned int*>(ws + l.ticket_off),
                                                                  loss_sum);
    ce_rank_finalize_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(ep.pcnt, (int)M, (int)l.Mpad, (int)l.S, ks, rank,
                                                                         reinterpret_cast<unsigned long long*>(errors));
    return (int)cudaGetLastError();
}

int pero_masked_ce_bwd_range(const void* h, int flags, int64_t N, int64_t Dh, const int32_t* rows, int64_t M,
                             const int64_t* labels, const void* head, int64_t V, const float* lse,
                             const float* grad_scale, float inv_count, int64_t v_begin, int64_t v_end, void* d_h, float* d_W,
                             float* d_b, void* workspace, size_t workspace_bytes, pero_stream_t stream) {
    if (M == 0) return PERO_ERR_BAD_SHAPE;
    const int h_is_bf16 = flags & PERO_CE_H_BF16, packed = (flags & PERO_CE_LABELS_PACKED) ? 1 : 0;
    int rc = ce_check(h, N, Dh, rows, M, labels, head, V, workspace, workspace_bytes, /*h_optional=*/true);
    if (rc) return rc;
    // h == NULL: the workspace is the one pero_masked_ce_fwd ran on for the same (h, rows, labels) and still holds
    // the gathered operands (A, A^T, labels, inverse row map): no second gather.
    // Phases (a data-parallel caller exchanges d_W | d_b of one label range while the next is being computed):
    //   d_W, d_b given  -> dlogits P / P^T of label columns [v_begin, v_end) into the workspace, rows [v_begin, v_end)
    //                      of d_W and d_b; then, if d_h is given too (requires the full range), d_h
    //   d_W == d_b == NULL, d_h given -> d_h from the P that earlier calls left in the SAME workspace for ALL columns
    //   d_W given, d_b == d_h == NULL -> only dlogits + d_W of the range (the exchange of d_W starts the moment its GEMM
    //                      is done); d_W == NULL, d_b and d_h given (full range) -> d_h AND d_b from the P in the workspace:
    //                      the column sums run beside the d_h GEMM and are exchanged later, as a small message of their own
    //   d_b only (d_W == d_h == NULL) -> the column sums of the P in the workspace (full range), e.g. on a side stream
    //                      beside the d_W / d_h GEMMs, so that d_b | loss can be exchanged early as a small message
    const bool dh_only = (!d_W && d_h);               // second phase: no dlogits, no d_W
    const bool dw_only = (d_W && !d_b && !d_h);
    const bool db_only = (!d_W && !d_h && d_b);
    const bool late_db = (dh_only && d_b);
    if (db_only) {
        if (Dh % 4 != 0 || v_begin != 0 || v_end != V) return PERO_ERR_BAD_SHAPE;
        const CeWsLayout l = ce_ws_layout(N, M, V, Dh);
        char* ws = static_cast<char*>(workspace);
        const __nv_bfloat16* P = reinterpret_cast<const __nv_bfloat16*>(ws + l.p_off);
        float* dbpart = reinterpret_cast<float*>(ws + l.dbpart_off);
        const int db_nparts = (int)((M + kDbRows - 1) / kDbRows);
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        // (PERO_CE_KEEP_LOGITS: the in-place dlogits pass of the earlier calls has left the partial sums already)
        if (!(h == nullptr && (flags & PERO_CE_KEEP_LOGITS)))
            ce_db_partial_kernel<<<(unsigned)db_nparts, 256, 0, st>>>(P, (int)M, (int)l.Pp, 0, (int)(l.Vp / 8), (int)l.Vp, dbpart, 0,
                                                                     (uint4*)nullptr, 0ll);
        ce_db_reduce_kernel<<<(unsigned)((V + 31) / 32), 256, 0, st>>>(dbpart, db_nparts, (int)l.Vp, 0, (int)V, d_b);
        return (int)cudaGetLastError();
    }
    if ((!lse && h) || (!dh_only && !d_W) || (!dh_only && !dw_only && !d_b)) return PERO_ERR_NULL;      // h == NULL: the log-sum-exp comes from the forward's partials
    if (Dh % 4 != 0) return PERO_ERR_BAD_SHAPE;
    if (v_begin < 0 || v_end > V || v_begin >= v_end || (v_begin % 256) != 0 || (v_end != V && (v_end % 256) != 0))
        return PERO_ERR_BAD_SHAPE;
    const bool full_range = (v_begin == 0 && v_end == V);
    if (d_h && !full_range && (!dh_only || late_db)) return PERO_ERR_UNSUPPORTED;
    const CeWsLayout l = ce_ws_layout(N, M, V, Dh);
    const HeadLayout hl = head_layout(V, Dh);
    char* ws = static_cast<char*>(workspace);
    const char* hb = static_cast<const char*>(head);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int* inv = reinterpret_cast<int*>(ws + l.inv_off);
    __nv_bfloat16* P = reinterpret_cast<__nv_bfloat16*>(ws + l.p_off);
    float* dbpart = reinterpret_cast<float*>(ws + l.dbpart_off);
    const int db_nparts = (int)((M + kDbRows - 1) / kDbRows);
    const int store_pairs = PERO_KNOB("PERO_CE_STORE_PAIRS", 1);     // dev build, 0: gradient GEMMs on single CTAs
    // When one call produces both gradients, the d_W and d_h GEMMs each get half of the SM pairs and run side by
    // side, with the d_b column sums filling in beside them: none of the three reads another's output, so each
    // kernel releases its successor at once (programmatic dependent launch) and every successor, before it exits,
    // waits for its predecessor, so that stream order still implies "all three are done".  Each GEMM CTA then walks
    // two tiles, and the store of one overlaps the loads of the next.
    const int pdl_on = PERO_KNOB("PERO_CE_PDL", 1);                  // dev build, 0: the gradient GEMMs run one after the other
    const bool side_by_side = pdl_on && store_pairs && !dh_only && d_h != nullptr && full_range;
    const bool db_beside_dh = late_db && pdl_on && store_pairs;      // second phase: column sums beside the d_h GEMM
    // Side by side, the two GEMMs share the SM pairs: the split of the pairs and the number of label-axis slices of d_h are
    // chosen together so that both halves finish at the same time (k-block steps of the slowest worker), with as few d_h
    // planes as that allows (the planes are written by the GEMM and read again by the scatter).
    int ks_use = (int)l.KS, dw_share = device_sm_count() / 4, dh_share = device_sm_count() / 4;
    if (side_by_side) {
        const int T = device_sm_count() / 2;
        const long long u_w = ((V + 255) / 256) * ((Dh + 255) / 256), kb_w = l.Mp64 / 64;
        const long long t_h = ((M + 255) / 256) * ((Dh + 255) / 256), kb_h = l.Vp / 64;
        long long best = -1;
        for (int ks = 1; ks <= (int)l.KS; ++ks) {
            const long long kbps = (kb_h + ks - 1) / ks, planes_n = (kb_h + kbps - 1) / kbps, units_h = t_h * planes_n;
            for (int wh = 1; wh < T; ++wh) {
                const int ww = T - wh;
                const long long tw = ((u_w + ww - 1) / ww) * kb_w, th = ((units_h + wh - 1) / wh) * kbps;
                const long long cost = std::max(tw, th) * 64 + planes_n;        // time first, then fewer planes
                if (best < 0 || cost < best) { best = cost; ks_use = ks; dw_share = ww; dh_share = wh; }
            }
        }
    }
    // PERO_CE_KEEP_LOGITS (with h == NULL): the forward left the bf16 logits in P; the first phase converts them in place
    // and produces the d_b partial sums in the same pass, so no later phase has to compute them
    const bool logits_in_ws = (h == nullptr) && (flags & PERO_CE_KEEP_LOGITS) != 0;
    bool use_dual = false, dh_cleared = false;
    GemmLaunch dual_w, dual_h;
    StoreTmaEpi::Params dual_ep_w;
    // d_h = P W runs as split-K planes [ks][M, Dh] over slices of the label axis
    float* planes = reinterpret_cast<float*>(ws + l.planes_off);
    // the number of planes actually produced is recomputed exactly as launch_gemm_tn does
    const int dh_num_kb = (int)(l.Vp / 64);
    const int dh_kb_per = (dh_num_kb + ks_use - 1) / ks_use;
    const int dh_planes = (dh_num_kb + dh_kb_per - 1) / dh_kb_per;
    StoreTmaEpi::Params sht;
    const bool dh_tma = d_h != nullptr && PERO_KNOB("PERO_STORE_TMA", 1) != 0 && store_pairs &&
                        make_tmap_f32_store(&sht.tmap_out, planes, (uint64_t)dh_planes, (uint64_t)M, (uint64_t)Dh, (uint64_t)Dh,
                                            (uint64_t)M * Dh) == PERO_OK;
    if (!dh_only) {
        if (h && v_begin == 0) {
            rc = ce_gather(h, h_is_bf16, rows, (int)M, (int)Dh, l, ws, st);
            if (rc) return rc;
        }
        const int64_t vlen = v_end - v_begin;
        const int64_t vp_range = (v_end == V ? l.Vp - v_begin : vlen);      // the last range also writes P's zero padding columns
        // Directly behind the forward on the same workspace (h == NULL): the log-sum-exp comes from the forward's partials
        // (every label range of such a backward uses them, so that walking the label axis range by range gives the same
        // bits as one call), and this GEMM need not wait for ce_finalize_kernel: it is released by it at once and only
        // waits for it before exiting.  Behind its own gather: the set-up overlaps the gather, then waits for it.
        const bool from_partials = (h == nullptr);
        // the set-up overlaps the tail of the kernel in front (this call's gather, or the forward's GEMM / finalize,
        // which release their successor early); the first global access waits for it
        const int dl_pdl = (!pdl_on || v_begin != 0) ? 0 : 4;
        auto fill = [&](auto& ep) {
            ep.colvec = reinterpret_cast<const float*>(hb + hl.bias_off) + v_begin;
            ep.rows = rows; ep.labels = reinterpret_cast<const long long*>(labels); ep.V = (int)V; ep.packed = packed;
            ep.lse = lse; ep.grad_scale = grad_scale; ep.inv_count = inv_count;
            ep.pm = from_partials ? reinterpret_cast<const float*>(ws + l.pm_off) : nullptr;
            ep.ps = from_partials ? reinterpret_cast<const float*>(ws + l.ps_off) : nullptr;
            ep.slots = (int)l.S; ep.Mpad = (int)l.Mpad;
            ep.p = P + v_begin; ep.M = (int)M;
            ep.Vp = (int)vp_range;
            ep.p_pitch = (int)l.Pp; ep.col_base = (int)v_begin;
        };
        const void* wslice = hb + hl.w_off + (size_t)v_begin * l.Dhp * 2;
        if (logits_in_ws) {
            // (side by side: the pass also clears d_h, so that the scatter only writes the masked frames)
            const long long zero_bytes = (long long)N * Dh * (h_is_bf16 ? 2 : 4);
            dh_cleared = side_by_side && (zero_bytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(d_h) & 15) == 0);
            // the forward left the softmax numerators in P: one in-place streaming pass (+ the d_b partial sums of the
            // range) instead of the logits GEMM; launched programmatically behind the kernel in front
            cudaLaunchConfig_t cfg = {};
            const long long tasks = (long long)db_nparts * ((vp_range / 8 + 255) / 256);
            if (tasks > (1ll << 30)) return PERO_ERR_BAD_SHAPE;
            const int grid = (int)tasks;
            cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(256); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = pdl_on ? 1 : 0;
            cudaError_t e = cudaLaunchKernelEx(&cfg, ce_dlogits_inplace_kernel, P, (int)M, (int)l.Pp, (int)v_begin, (int)(vp_range / 8),
                                               (int)l.Vp, reinterpret_cast<const float*>(ws + l.pm_off),
                                               reinterpret_cast<const float*>(ws + l.ps_off), (int)l.S, (int)l.Mpad,
                                               reinterpret_cast<const float*>(ws + l.cm_off), (const int*)rows, reinterpret_cast<const long long*>(labels), (int)V, packed, grad_scale,
                                               inv_count, dbpart, pdl_on ? 1 : 0, dh_cleared ? static_cast<uint4*>(d_h) : (uint4*)nullptr,
                                               dh_cleared ? zero_bytes / 16 : 0ll);
            if (e != cudaSuccess) return (int)e;
        } else {
            DlogitsEpi::Params ep;
            fill(ep);
            // store descriptor of P[:, v_begin : v_begin + vp_range]: boxes of {64 columns, 32 rows}, clipped at M rows
            rc = make_tmap_bf16(&ep.tmap_p, P + v_begin, (uint64_t)M, (uint64_t)vp_range, (uint64_t)l.Pp, 32);
            if (rc) return rc;
            rc = launch_logits_gemm<DlogitsEpi>(ws + l.a_off, (int)M, (int)l.Dhp, wslice, (int)vlen, 0, 1, ep, st, kSmemBudget, dl_pdl);
        }
        if (rc) return rc;

        // d_W [v_begin:v_end, Dh] = P[:, v_begin:v_end]^T @ A: both operands are read as stored (rows = masked
        // frames = the contraction index) through MN-major descriptors; no transposed copy of either exists.
        const bool store_tma = PERO_KNOB("PERO_STORE_TMA", 1) != 0;
        StoreTmaEpi::Params swt;
        const bool dw_tma = store_tma && store_pairs &&
                            make_tmap_f32_store(&swt.tmap_out, d_W + (size_t)v_begin * Dh, 1, (uint64_t)vlen, (uint64_t)Dh, (uint64_t)Dh,
                                                (uint64_t)vlen * Dh) == PERO_OK;
        StoreEpi::Params sw;
        sw.out = d_W + (size_t)v_begin * Dh; sw.ld = Dh; sw.split_stride = 0; sw.rows = (int)vlen; sw.cols = (int)Dh;
        const int dw_workers = side_by_side ? dw_share : 0;
        // behind the in-place dlogits pass the gradient GEMMs are programmatic dependents: set-up during its tail, first
        // load after it (bit 2)
        const int dw_pdl = ((side_by_side || pdl_on) ? 1 : 0) | ((logits_in_ws && pdl_on) ? 4 : 0);
        use_dual = side_by_side && dw_tma && dh_tma && PERO_KNOB("PERO_CE_DUAL", 1) != 0;
        if (use_dual) {
            // d_W and d_h in ONE grid (gemm_dual_kernel), launched below once the d_h half is prepared
            dual_ep_w = swt;
            rc = prepare_gemm_tn<2, 0, StoreTmaEpi, 3>(dual_w, P + v_begin, (int)vlen, (int)l.Pp, ws + l.a_off, (int)Dh, (int)l.Dhp,
                                                       (int)l.Mp64, 1, 0, 1, dw_workers, nullptr, kSmemBudgetShared, (int)M, dw_pdl);
        } else if (dw_tma)
            rc = launch_gemm_tn<2, 0, StoreTmaEpi, 3>(P + v_begin, (int)vlen, (int)l.Pp, ws + l.a_off, (int)Dh, (int)l.Dhp, (int)l.Mp64,
                                                         1, 0, 1, dw_workers, swt, st, nullptr, kSmemBudgetShared, (int)M, dw_pdl);
        else if (store_pairs)
            rc = launch_gemm_tn<2, 0, StoreEpi, 3>(P + v_begin, (int)vlen, (int)l.Pp, ws + l.a_off, (int)Dh, (int)l.Dhp, (int)l.Mp64,
                                                      1, 0, 1, dw_workers, sw, st, nullptr, kSmemBudgetShared, (int)M, dw_pdl);
        else
            rc = launch_gemm_tn<1, 0, StoreEpi, 3>(P + v_begin, (int)vlen, (int)l.Pp, ws + l.a_off, (int)Dh, (int)l.Dhp, (int)l.Mp64,
                                                      1, 0, 1, 0, sw, st, nullptr, kSmemBudgetShared, (int)M);
        if (rc) return rc;
        // d_b: partial column sums of P now, unless the side-by-side schedule below runs them beside the GEMMs; the
        // final sums on their own when this call stops after d_W | d_b (they are exchanged next), otherwise by the
        // leading blocks of the scatter launch
        if (!side_by_side && !dw_only && !logits_in_ws) {
            // beside the d_W GEMM (released by it at once, waits for it before exiting) when PDL is on
            const int vlen8 = (int)(vp_range / 8);      // P's padding columns are zeros
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)db_nparts); cfg.blockDim = dim3(256); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = pdl_on ? 1 : 0;
            cudaError_t e = cudaLaunchKernelEx(&cfg, ce_db_partial_kernel, (const __nv_bfloat16*)P, (int)M, (int)l.Pp, (int)v_begin,
                                               vlen8, (int)l.Vp, dbpart, pdl_on ? 1 : 0, (uint4*)nullptr, 0ll);
            if (e != cudaSuccess) return (int)e;
        }
        if (!d_h && !dw_only)
            ce_db_reduce_kernel<<<(unsigned)((vlen + 31) / 32), 256, 0, st>>>(dbpart, db_nparts, (int)l.Vp, (int)v_begin,
                                                                              (int)v_end, d_b);
    }

    bool scatter_masked_only = false;
    if (d_h) {
        // planes[ks] [M, Dh] = P [M, Vp] @ W [V, Dh] over the ks-th slice of the label axis: P is the K-major operand, the
        // head's one bf16 copy is read MN-major (boxes of {64 hidden channels, 64 labels}; labels beyond V read as zero)
        StoreEpi::Params sh;
        sh.out = planes; sh.ld = Dh; sh.split_stride = (long long)M * Dh; sh.rows = (int)M; sh.cols = (int)Dh;
        const int dh_workers = side_by_side ? dh_share : 0;
        const int dh_pdl = side_by_side ? 3 : (db_beside_dh ? 1 : 0);
        if (use_dual) {
            rc = prepare_gemm_tn<2, 0, StoreTmaEpi, 2>(dual_h, P, (int)M, (int)l.Pp, hb + hl.w_off, (int)Dh, (int)l.Dhp, (int)l.Vp,
                                                       ks_use, 0, 1, dh_workers, nullptr, kSmemBudgetShared, (int)V, dual_w.sh.pdl);
            if (rc) return rc;
            rc = launch_gemm_dual<StoreTmaEpi, 3, 2>(dual_w, dual_ep_w, dual_h, sht, st);
        } else if (dh_tma)
            rc = launch_gemm_tn<2, 0, StoreTmaEpi, 2>(P, (int)M, (int)l.Pp, hb + hl.w_off, (int)Dh, (int)l.Dhp, (int)l.Vp, ks_use, 0, 1,
                                                      dh_workers, sht, st, nullptr, kSmemBudgetShared, (int)V, dh_pdl);
        else if (store_pairs)
            rc = launch_gemm_tn<2, 0, StoreEpi, 2>(P, (int)M, (int)l.Pp, hb + hl.w_off, (int)Dh, (int)l.Dhp, (int)l.Vp, ks_use, 0, 1,
                                                   dh_workers, sh, st, nullptr, kSmemBudgetShared, (int)V, dh_pdl);
        else
            rc = launch_gemm_tn<1, 0, StoreEpi, 2>(P, (int)M, (int)l.Pp, hb + hl.w_off, (int)Dh, (int)l.Dhp, (int)l.Vp, ks_use, 0, 1, 0,
                                                   sh, st, nullptr, kSmemBudgetShared, (int)V);
        if (rc) return rc;
        if (late_db && !db_beside_dh && !logits_in_ws) {        // no programmatic launches: plain column sums behind the GEMM
            ce_db_partial_kernel<<<(unsigned)db_nparts, 256, 0, st>>>((const __nv_bfloat16*)P, (int)M, (int)l.Pp, 0, (int)(l.Vp / 8),
                                                                     (int)l.Vp, dbpart, 0, (uint4*)nullptr, 0ll);
        }
        if (dh_cleared) {
            scatter_masked_only = true;
        } else if (side_by_side || db_beside_dh) {
            // third member of the side-by-side group: released by the d_h GEMM as soon as that one has started
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)db_nparts); cfg.blockDim = dim3(256); cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            const long long zero_vec = (long long)N * Dh * (h_is_bf16 ? 2 : 4) / 16;       // Dh % 4 == 0 and, for bf16,
            const bool can_zero = (((long long)N * Dh * (h_is_bf16 ? 2 : 4)) % 16 == 0) &&    // whole 16-byte words
                                  ((reinterpret_cast<uintptr_t>(d_h) & 15) == 0);
            scatter_masked_only = can_zero;
            // (logits_in_ws: the partial sums exist already and this launch only clears d_h)
            cudaError_t e = cudaLaunchKernelEx(&cfg, ce_db_partial_kernel, (const __nv_bfloat16*)P, (int)M, (int)l.Pp, 0,
                                               logits_in_ws ? 0 : (int)(l.Vp / 8), (int)l.Vp, dbpart, 1,
                                               can_zero ? static_cast<uint4*>(d_h) : (uint4*)nullptr, can_zero ? zero_vec : 0ll);
            if (e != cudaSuccess) return (int)e;
        }
        const long long total = (scatter_masked_only ? M : N) * (Dh / 4);
        long long blocks = (total + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        const int db_blocks = (dh_only && !late_db) ? 0 : (int)((V + 31) / 32);
        blocks += db_blocks;
        const int ks_eff = dh_planes;
        // (a plain launch on purpose: released programmatically, the ~800 CTAs of this grid park on every SM until
        // the GEMMs in front are done and starve the kernels of the other chains -- measured, round 2)
        if (h_is_bf16)
            ce_dh_scatter_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, st>>>(planes, inv, rows, N, (int)M, (int)Dh, ks_eff,
                                                                                 static_cast<__nv_bfloat16*>(d_h), db_blocks, dbpart,
                                                                                 db_nparts, (int)l.Vp, (int)V, d_b,
                                                                                 scatter_masked_only ? 1 : 0);
        else
            ce_dh_scatter_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(planes, inv, rows, N, (int)M, (int)Dh, ks_eff,
                                                                         static_cast<float*>(d_h), db_blocks, dbpart, db_nparts,
                                                                         (int)l.Vp, (int)V, d_b, scatter_masked_only ? 1 : 0);
    }
    return (int)cudaGetLastError();
}

int pero_masked_ce_bwd(const void* h, int flags, int64_t N, int64_t Dh, const int32_t* rows, int64_t M,
                       const int64_t* labels, const void* head, int64_t V, const float* lse,
                       const float* grad_scale, float inv_count, void* d_h, float* d_W, float* d_b,
                       void* workspace, size_t workspace_bytes, pero_stream_t stream) {
    return pero_masked_ce_bwd_range(h, flags, N, Dh, rows, M, labels, head, V, lse, grad_scale, inv_count, 0, V, d_h, d_W,
                                    d_b, workspace, workspace_bytes, stream);
}

int pero_ce_logits_fwd(const void* logits, int is_bf16, int64_t N, int64_t V, const int32_t* rows, int64_t M,
                       const int64_t* labels, float* loss_sum, float* lse, void* workspace, size_t workspace_bytes,
                       pero_stream_t stream) {
    if (!logits || !rows || !labels || !loss_sum || !lse || !workspace) return PERO_ERR_NULL;
    if (N <= 0 || M <= 0 || M > N || V <= 0 || V > (1ll << 30)) return PERO_ERR_BAD_SHAPE;
    if (workspace_bytes < align256((size_t)M * 4)) return PERO_ERR_WORKSPACE;
    float* rowloss = static_cast<float*>(workspace);
    const unsigned grid = (unsigned)((M + 7) / 8);
    if (is_bf16)
        ce_logits_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), rows,
                                                                      reinterpret_cast<const long long*>(labels), (int)M, (int)V, lse, rowloss);
    else
        ce_logits_fwd_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(logits), rows,
                                                              reinterpret_cast<const long long*>(labels), (int)M, (int)V, lse, rowloss);
    ce_sum_kernel<<<1, 1024, 0, stream>>>(rowloss, (int)M, loss_sum);
    return (int)cudaGetLastError();
}

int pero_ce_logits_bwd(const void* logits, int is_bf16, int64_t N, int64_t V, const int32_t* rows, int64_t M,
                       const int64_t* labels, const float* lse, const float* grad_scale, float inv_count, int zero_init,
                       void* d_logits, pero_stream_t stream) {
    if (!logits || !rows || !labels || !lse || !d_logits) return PERO_ERR_NULL;
    if (N <= 0 || M <= 0 || M > N || V <= 0 || V > (1ll << 30)) return PERO_ERR_BAD_SHAPE;
    if (zero_init) {
        cudaError_t e = cudaMemsetAsync(d_logits, 0, (size_t)N * V * (is_bf16 ? 2 : 4), stream);
        if (e != cudaSuccess) return (int)e;
    }
    const unsigned grid = (unsigned)((M + 7) / 8);
    if (is_bf16)
        ce_logits_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), rows,
            reinterpret_cast<const long long*>(labels), lse, grad_scale, inv_count, (int)M, (int)V, static_cast<__nv_bfloat16*>(d_logits));
    else
        ce_logits_bwd_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(logits), rows,
            reinterpret_cast<const long long*>(labels), lse, grad_scale, inv_count, (int)M, (int)V, static_cast<float*>(d_logits));
    return (int)cudaGetLastError();
}

size_t pero_mask_compact_workspace_bytes(int64_t N) {
    if (N <= 0) return 0;
    size_t bytes = 0;
    MaskPred pred{nullptr, 0, 1, nullptr};
    cudaError_t e = cub::DeviceSelect::If(nullptr, bytes, thrust::counting_iterator<int>(0), (int32_t*)nullptr,
                                          (int32_t*)nullptr, (int)N, pred);
    if (e != cudaSuccess || bytes == 0) {      // no device to query (CPU-only host): conservative bound
        (void)cudaGetLastError();
        bytes = (size_t)N + (1u << 20);
    }
    return align256(bytes);
}

int pero_mask_compact(const void* mask, int mask_dtype, int want_value, const int64_t* labels_or_null, int64_t N,
                      int32_t* rows, int32_t* count, void* workspace, size_t workspace_bytes,
                      pero_stream_t stream) {
    if (!mask || !rows || !count || !workspace) return PERO_ERR_NULL;
    if (N <= 0 || N > (1ll << 31) - 256 || mask_dtype < 0 || mask_dtype > 2) return PERO_ERR_BAD_SHAPE;
    size_t bytes = workspace_bytes;
    MaskPred pred{mask, mask_dtype, want_value, reinterpret_cast<const long long*>(labels_or_null)};
    if (workspace_bytes < pero_mask_compact_workspace_bytes(N)) return PERO_ERR_WORKSPACE;
    cudaError_t e = cub::DeviceSelect::If(workspace, bytes, thrust::counting_iterator<int>(0), rows, count, (int)N, pred, stream);
    return (int)e;
}

}  // extern "C"
