// 1x1 projections around the quantizer: VQVAE.quantize, models/autoencoders.py:142-147, wraps the
// quantizer in two 1x1 Conv2d layers (:114-115).  A 1x1 convolution over [n_lines, C, H, W] is the GEMM
//     y[n, :] = W x[n, :] + b          over the N = n_lines * H * W frames,
// so the encoder projection is folded into the distance GEMM's operand preparation: ONE pass splits the channels-first
// fp32 features into bf16 (hi, lo) rows, a tcgen05 GEMM multiplies them with the split weights, and its epilogue adds
// the bias and writes exactly the two copies the quantizer consumes (fp32 rows for gather / EMA, bf16 rows [N, Dp] as
// the distance GEMM's K-major operand) and resets the packed winners.  The projected NCHW tensor never exists.
// The decoder projection commutes with the gather: W_d e[idx] + b_d = (E W_d^T + b_d)[idx], so it is the same GEMM
// over the K codewords followed by a row gather into the channels-first output (a label-production loop with a fixed
// codebook projects the codebook once).
//
// Precision: the reference runs these convolutions in fp32.  bf16 operands alone would lose 8 bits, so every fp32
// operand v is split into hi = bf16(v), lo = bf16(v - hi) and the contraction is laid out as
//     A3 = [a_hi | a_lo | a_hi],  B3 = [b_hi | b_hi | b_lo]      (3 x Cp columns)
// i.e. a_hi b_hi + a_lo b_hi + a_hi b_lo with fp32 accumulation in TMEM: the dropped a_lo b_lo term and the rounding of
// lo are ~2^-16 relative, 60 times finer than the TF32 arithmetic cuDNN uses by default for these layers.
#include <cuda_bf16.h>
#include "../../include/pero_b200.h"
#include "epilogues.cuh"
#include "gemm_host.cuh"
#include "layout.h"

namespace pero {

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(v);
    lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Channels-first features [n_lines, C, HW] -> split rows [N, 3 * Cp] = [hi | lo | hi], 64(c) x 32(hw) tiles through
// shared memory (reads coalesced along hw, writes along c).
__global__ void __launch_bounds__(256)
proj_split_cf_kernel(const float* __restrict__ x, int C, int Cp, int HW, __nv_bfloat16* __restrict__ a3) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ float tile[64][33];
    const int nl = blockIdx.z, hw0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* xl = x + (size_t)nl * C * HW;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + ty + i * 8, hw = hw0 + tx;
        tile[ty + i * 8][tx] = (c < C && hw < HW) ? __ldg(xl + (size_t)c * HW + hw) : 0.f;
    }
    __syncthreads();
    const int c = c0 + 2 * tx;
    if (c >= Cp) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int hw = hw0 + ty + i * 8;
        if (hw >= HW) continue;
        const size_t n = (size_t)nl * HW + hw;
        __nv_bfloat162 hi, lo;
        split_bf16(tile[2 * tx][ty + i * 8], hi.x, lo.x);
        split_bf16(tile[2 * tx + 1][ty + i * 8], hi.y, lo.y);
        __nv_bfloat16* row = a3 + n * (size_t)(3 * Cp) + c;
        *reinterpret_cast<__nv_bfloat162*>(row) = hi;
        *reinterpret_cast<__nv_bfloat162*>(row + Cp) = lo;
        *reinterpret_cast<__nv_bfloat162*>(row + 2 * Cp) = hi;
    }
}

// Row-major fp32 [R, C] -> split rows [R, 3 * Cp]; lo_slot = 1: [hi | lo | hi] (the A side), 2: [hi | hi | lo] (B side).
// Block 0 also writes the zero-padded bias vector the epilogue reads ([Dpad], Dpad a multiple of 256).
__global__ void __launch_bounds__(256)
proj_split_rows_kernel(const float* __restrict__ x, int C, int Cp, long long R, int lo_slot, __nv_bfloat16* __restrict__ out,
                       const float* __restrict__ bias, int D, int Dpad, float* __restrict__ bias_pad) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (bias_pad && blockIdx.x == 0)
        for (int j = threadIdx.x; j < Dpad; j += blockDim.x) bias_pad[j] = (bias && j < D) ? __ldg(bias + j) : 0.f;
    const long long pairs = R * (Cp / 2);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < pairs; p += stride) {
        const long long r = p / (Cp / 2);
        const int c = (int)(p - r * (Cp / 2)) * 2;
        const float a = c < C ? __ldg(x + r * C + c) : 0.f;
        const float b = c + 1 < C ? __ldg(x + r * C + c + 1) : 0.f;
        __nv_bfloat162 hi, lo;
        split_bf16(a, hi.x, lo.x);
        split_bf16(b, hi.y, lo.y);
        __nv_bfloat16* row = out + r * (long long)(3 * Cp) + c;
        *reinterpret_cast<__nv_bfloat162*>(row) = hi;
        *reinterpret_cast<__nv_bfloat162*>(row + Cp) = (lo_slot == 1) ? lo : hi;
        *reinterpret_cast<__nv_bfloat162*>(row + 2 * Cp) = (lo_slot == 1) ? hi : lo;
    }
}

// y = acc + bias: fp32 rows [rows, cols] (row pitch ld; 32 x 32 chunks transposed through a warp-private tile so that a
// store instruction writes four full 128-byte lines, as StoreEpi does) and, optionally, the bf16 copy [rows, cols_b]
// (cols_b = cols rounded up to 64; the padding columns receive acc = 0 + bias_pad = 0) that the distance GEMM reads as
// its K-major operand, and the reset of the packed (distance, index) winners of these rows.
struct ProjEpi {
    static constexpr int kMaxRegs = 128;
    static constexpr bool kColVec = true;
    static constexpr int kScratchPerWarp = 32 * 144;
    struct Params {
        const float* colvec;      // bias, zero padded to num_ct * 256
        float* out;               // [rows, cols] fp32 or NULL
        long long ld;
        __nv_bfloat16* outb;      // [rows, cols_b] bf16 or NULL
        long long pitch_b;
        long long* packed;        // [rows] or NULL: reset to "empty"
        int rows, cols, cols_b;
    };
    struct State {};
    static __device__ __forceinline__ void begin_rb(State&, const Params&, const TileCtx&) {}
    static __device__ __forceinline__ void tile(State&, const Params& ep, const TileCtx& cx, uint32_t taddr) {
        const int lane = threadIdx.x & 31;
        const int row_base = cx.row - lane;
        if (ep.packed && cx.ct == 0 && cx.half == 0 && cx.row < ep.rows) ep.packed[cx.row] = kPackedEmpty;
        float* base = ep.out ? ep.out + (long long)row_base * ep.ld + cx.col0 : nullptr;
        const bool vec_ok = ((ep.ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(ep.out) & 15) == 0);
        for_each_chunk(taddr, [&](int c, const uint32_t (&r)[32]) {
            const int col = cx.col0 + c * 32;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + cx.cv[c * 32 + j];
            if (ep.outb && cx.row < ep.rows && col + 32 <= ep.cols_b) {
                uint4* dst = reinterpret_cast<uint4*>(ep.outb + (long long)cx.row * ep.pitch_b + col);
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 o;
                    __nv_bfloat162 t;
                    t = __floats2bfloat162_rn(v[8 * g + 0], v[8 * g + 1]); o.x = *reinterpret_cast<uint32_t*>(&t);
                    t = __floats2bfloat162_rn(v[8 * g + 2], v[8 * g + 3]); o.y = *reinterpret_cast<uint32_t*>(&t);
                    t = __floats2bfloat162_rn(v[8 * g + 4], v[8 * g + 5]); o.z = *reinterpret_cast<uint32_t*>(&t);
                    t = __floats2bfloat162_rn(v[8 * g + 6], v[8 * g + 7]); o.w = *reinterpret_cast<uint32_t*>(&t);
                    dst[g] = o;
                }
            }
            if (!base) return;
            if (vec_ok && col + 32 <= ep.cols) {
                float4* srow = reinterpret_cast<float4*>(cx.scratch + lane * 144);
#pragma unroll
                for (int i = 0; i < 8; ++i) srow[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int rr = 4 * k + (lane >> 3), piece = lane & 7;
                    const float4 q = *reinterpret_cast<const float4*>(cx.scratch + rr * 144 + piece * 16);
                    if (row_base + rr < ep.rows)
                        *reinterpret_cast<float4*>(base + (long long)rr * ep.ld + c * 32 + piece * 4) = q;
                }
                __syncwarp();
            } else if (cx.row < ep.rows) {
                float* dst = base + (long long)lane * ep.ld + c * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col + j < ep.cols) dst[j] = v[j];
            }
        });
    }
    static __device__ __forceinline__ void end_rb(State&, const Params&, const TileCtx&) {}
};

// out[nl, c, hw] = table[idx[nl * HW + hw], c]: table rows are read coalesced along c, transposed through shared memory
// in 64(c) x 32(hw) tiles and written coalesced along hw (the channels-first tensor the decoder expects).
__global__ void __launch_bounds__(256)
gather_rows_cf_kernel(const float* __restrict__ table, const long long* __restrict__ idx, int C, int HW, long long K,
                      float* __restrict__ out) {
    __shared__ float tile[64][33];
    const int nl = blockIdx.z, hw0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = c0 + 2 * tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int hwl = ty + i * 8, hw = hw0 + hwl;
        float q0 = 0.f, q1 = 0.f;
        if (hw < HW) {
            long long k = __ldg(idx + (size_t)nl * HW + hw);
            k = k < 0 ? 0 : (k >= K ? K - 1 : k);          // an index outside the table cannot be dereferenced
            const float* row = table + (size_t)k * C;
            if (c < C) q0 = __ldg(row + c);
            if (c + 1 < C) q1 = __ldg(row + c + 1);
        }
        tile[2 * tx][hwl] = q0;
        tile[2 * tx + 1][hwl] = q1;
    }
    __syncthreads();
    float* ol = out + (size_t)nl * C * HW;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int cc = c0 + ty + i * 8, hw = hw0 + tx;
        if (cc < C && hw < HW) ol[(size_t)cc * HW + hw] = tile[ty + i * 8][tx];
    }
}

struct ProjWsLayout { int64_t Cp, Dpad; size_t a_off, b_off, bias_off, total; };
inline ProjWsLayout proj_ws_layout(int64_t N, int64_t C, int64_t D) {
    ProjWsLayout l;
    l.Cp = round_up(C, 64); l.Dpad = round_up(D, 256);
    l.a_off = 0;
    l.b_off = align256((size_t)N * 3 * l.Cp * 2);
    l.bias_off = l.b_off + align256((size_t)D * 3 * l.Cp * 2);
    l.total = l.bias_off + align256((size_t)l.Dpad * 4);
    return l;
}

}  // namespace pero

using namespace pero;

extern "C" {

size_t pero_proj_workspace_bytes(int64_t N, int64_t C, int64_t D) {
    if (N <= 0 || C <= 0 || D <= 0) return 0;
    return proj_ws_layout(N, C, D).total;
}

int pero_proj_forward(const float* x, int64_t n_lines, int64_t frames_per_line, int channels_first, int64_t C,
                      const float* weight, const float* bias, int64_t D, float* out_rows, void* out_bf16,
                      int64_t* packed_reset, void* workspace, size_t workspace_bytes, pero_stream_t stream) {
    if (n_lines < 0 || frames_per_line < 0) return PERO_ERR_BAD_SHAPE;
    const int64_t N = n_lines * frames_per_line;
    if (N == 0) return PERO_OK;
    if (!x || !weight || !workspace || (!out_rows && !out_bf16)) return PERO_ERR_NULL;
    if (C <= 0 || D <= 0 || N > (1ll << 31) - 256 || C > 16384 || D > (1ll << 24)) return PERO_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return PERO_ERR_BAD_ALIGN;
    if (out_bf16 && (reinterpret_cast<uintptr_t>(out_bf16) & 15)) return PERO_ERR_BAD_ALIGN;
    const ProjWsLayout l = proj_ws_layout(N, C, D);
    if (workspace_bytes < l.total) return PERO_ERR_WORKSPACE;
    char* ws = static_cast<char*>(workspace);
    __nv_bfloat16* a3 = reinterpret_cast<__nv_bfloat16*>(ws + l.a_off);
    __nv_bfloat16* b3 = reinterpret_cast<__nv_bfloat16*>(ws + l.b_off);
    float* bias_pad = reinterpret_cast<float*>(ws + l.bias_off);
    const int Cp = (int)l.Cp;
    cudaStream_t st = (cudaStream_t)stream;

    {   // weights [D, C] -> [hi | hi | lo] rows (+ the padded bias)
        long long blocks = (D * (Cp / 2) + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        proj_split_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(weight, (int)C, Cp, D, 2, b3, bias, (int)D, (int)l.Dpad, bias_pad);
    }
    if (channels_first) {
        if (n_lines > 65535) return PERO_ERR_BAD_SHAPE;
        dim3 grid((unsigned)((frames_per_line + 31) / 32), (unsigned)(Cp / 64), (unsigned)n_lines);
        proj_split_cf_kernel<<<grid, 256, 0, st>>>(x, (int)C, Cp, (int)frames_per_line, a3);
    } else {
        long long blocks = (N * (Cp / 2) + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        proj_split_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, (int)C, Cp, N, 1, a3, nullptr, 0, 0, nullptr);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;

    ProjEpi::Params ep;
    ep.colvec = bias_pad;
    ep.out = out_rows; ep.ld = D;
    ep.outb = static_cast<__nv_bfloat16*>(out_bf16); ep.pitch_b = round_up(D, 64);
    ep.packed = reinterpret_cast<long long*>(packed_reset);
    ep.rows = (int)N; ep.cols = (int)D; ep.cols_b = (int)round_up(D, 64);
    // streamed pair GEMM over the 3 * Cp contraction; the set-up overlaps the split pass in front (bit 2)
    return launch_gemm_tn<2, 0, ProjEpi>(a3, (int)N, 3 * Cp, b3, (int)D, 3 * Cp, 3 * Cp, 1, 0, 1, 0, ep, st, nullptr, kSmemBudget,
                                         0, /*pdl=*/4);
}

int pero_gather_rows_cf(const float* table, const int64_t* idx, int64_t n_lines, int64_t frames_per_line, int64_t K,
                        int64_t C, float* out, pero_stream_t stream) {
    if (n_lines < 0 || frames_per_line < 0) return PERO_ERR_BAD_SHAPE;
    if (n_lines * frames_per_line == 0) return PERO_OK;
    if (!table || !idx || !out) return PERO_ERR_NULL;
    if (K <= 0 || C <= 0 || n_lines > 65535 || C > 65535 * 64) return PERO_ERR_BAD_SHAPE;
    dim3 grid((unsigned)((frames_per_line + 31) / 32), (unsigned)((C + 63) / 64), (unsigned)n_lines);
    gather_rows_cf_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table, reinterpret_cast<const long long*>(idx), (int)C,
                                                                 (int)frames_per_line, K, out);
    return (int)cudaGetLastError();
}

}  // extern "C"
