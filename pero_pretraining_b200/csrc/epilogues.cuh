// Epilogue policies for gemm_core.cuh.  Each epilogue thread owns ONE accumulator row and walks the
// 256 columns of a tile 32 at a time (tcgen05.ld 32x32b.x32).
#pragma once
#include "gemm_core.cuh"
#include <math_constants.h>

namespace pero {

// ------------------------------------------------------------------------------------------------
// Nearest codeword: d[row, col] = |c_col|^2 - 2 * <x_row, c_col>  (|x_row|^2 is constant along the
// row and cannot change the arg-min; reference: models/autoencoders.py:212-217).  The running
// (min, argmin) stays in registers across the column sweep; strict '<' in ascending column order keeps
// torch.argmin's first-index-on-ties rule.  Results of different workers on the same row are merged
// with one signed 64-bit atomicMin on (order_key(d) << 32 | index): min distance first, then lowest index.
struct ArgminEpi {
    struct Params {
        const float* cnorm;            // [num_ct * 256] |c|^2 in fp32, +inf beyond the last codeword
        long long* packed;             // [rows] pre-set to kPackedEmpty
        int rows;
        int index_offset;              // global index of this shard's codeword 0
    };
    struct State { float best; int besti; };

    static __device__ __forceinline__ void begin_rb(State& st, const Params&, const TileCtx&) {
        st.best = CUDART_INF_F; st.besti = 0;
    }
    static __device__ __forceinline__ void tile(State& st, const Params& ep, const TileCtx& cx, uint32_t taddr) {
        float tb = CUDART_INF_F; int tj = 0;
#pragma unroll 1
        for (int c = 0; c < kBlockN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            const float4* cn = reinterpret_cast<const float4*>(ep.cnorm + cx.col0 + c * 32);
            float4 n[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) n[i] = __ldg(cn + i);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float nn[4] = {n[i].x, n[i].y, n[i].z, n[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float d = fmaf(__uint_as_float(r[i * 4 + e]), -2.0f, nn[e]);
                    if (d < tb) { tb = d; tj = c * 32 + i * 4 + e; }
                }
            }
        }
        if (tb < st.best) { st.best = tb; st.besti = cx.col0 + tj; }
    }
    static __device__ __forceinline__ void end_rb(State& st, const Params& ep, const TileCtx& cx) {
        if (cx.row < ep.rows) {
            atomicMin(ep.packed + cx.row, pack_dist_index(st.best, st.besti + ep.index_offset));
        }
    }
};

// ------------------------------------------------------------------------------------------------
// Plain fp32 store C[row, col] (one plane per contraction split).  128 contiguous bytes per thread per
// step, so every 32-byte sector written is full.
struct StoreEpi {
    struct Params {
        float* out;
        long long ld;             // elements between output rows
        long long split_stride;   // elements between split planes
        int rows, cols;
    };
    struct State {};
    static __device__ __forceinline__ void begin_rb(State&, const Params&, const TileCtx&) {}
    static __device__ __forceinline__ void tile(State&, const Params& ep, const TileCtx& cx, uint32_t taddr) {
        float* dst = ep.out + (long long)cx.ks * ep.split_stride + (long long)cx.row * ep.ld + cx.col0;
        const bool row_ok = cx.row < ep.rows;
        const bool vec_ok = ((ep.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.out) & 15) == 0) &&
                            ((ep.split_stride & 3) == 0);
#pragma unroll 1
        for (int c = 0; c < kBlockN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
            if (!row_ok) continue;
            const int col = cx.col0 + c * 32;
            if (vec_ok && col + 32 <= ep.cols) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    reinterpret_cast<float4*>(dst + c * 32)[i] =
                        make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]),
                                    __uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3]));
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col + j < ep.cols) dst[c * 32 + j] = __uint_as_float(r[j]);
            }
        }
    }
    static __device__ __forceinline__ void end_rb(State&, const Params&, const TileCtx&) {}
};

}  // namespace pero
