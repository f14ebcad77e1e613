// Nearest-codeword assignment: operand preparation, the tcgen05 distance
// GEMM with the fused bias + arg-min epilogue, and the (distance, index) unpack.
// Reference behaviour: models/autoencoders.py:205-217, scripts/produce_kmeans_labels.py:72-76.
#include <cuda_bf16.h>
#include <stdlib.h>
#include "../../include/pero_b200.h"
#include "epilogues.cuh"
#include "gemm_host.cuh"
#include "layout.h"

namespace pero {

// One warp per codeword: bf16 copy (zero-padded to Dp columns) and |c|^2 from the fp32 weights.
// |c|^2 can reach 1e12 after the reference's cold-start EMA step, so it is never taken from bf16.
__global__ void codebook_prepare_kernel(const float* __restrict__ w, int K, int D, int Dp, int Kp,
                                        __nv_bfloat16* __restrict__ cb, float* __restrict__ cnorm) {
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (k >= Kp) return;
    if (k >= K) { if (lane == 0) cnorm[k] = CUDART_INF_F; return; }
    const float* row = w + (size_t)k * D;
    __nv_bfloat16* dst = cb + (size_t)k * Dp;
    float s = 0.f;
    for (int d = lane; d < Dp; d += 32) {
        const float v = d < D ? row[d] : 0.f;
        s = fmaf(v, v, s);
        dst[d] = __float2bfloat16_rn(v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) cnorm[k] = s;
}

// Frames arrive channels-first [n_lines, D, HW] (the NCHW tensor of VectorQuantizer.forward with H*W
// collapsed).  One pass transposes 64(d) x 32(hw) tiles through shared memory into the row-major
// bf16 GEMM operand [N, Dp] and, optionally, the fp32 rows the gather/EMA stages read coalesced.
__global__ void __launch_bounds__(256)
frames_prepare_cf_kernel(const float* __restrict__ x, int D, int Dp, int HW, long long N,
                         __nv_bfloat16* __restrict__ xb, float* __restrict__ xr, long long* __restrict__ packed) {
    // the distance GEMM behind this kernel sets itself up meanwhile and waits for this grid before its first load
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ float tile[64][33];
    const int nl = blockIdx.z, hw0 = blockIdx.x * 32, d0 = blockIdx.y * 64;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const float* xl = x + (size_t)nl * D * HW;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int d = d0 + ty + i * 8, hw = hw0 + tx;
        tile[ty + i * 8][tx] = (d < D && hw < HW) ? __ldg(xl + (size_t)d * HW + hw) : 0.f;
    }
    __syncthreads();
    const int d = d0 + 2 * tx;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int hw = hw0 + ty + i * 8;
        if (hw >= HW) continue;
        const size_t n = (size_t)nl * HW + hw;
        const float a = tile[2 * tx][ty + i * 8], b = tile[2 * tx + 1][ty + i * 8];
        if (d < Dp) *reinterpret_cast<__nv_bfloat162*>(xb + n * Dp + d) = __floats2bfloat162_rn(a, b);
        if (xr) {
            if (d < D) xr[n * D + d] = a;
            if (d + 1 < D) xr[n * D + d + 1] = b;
        }
    }
    if (packed) {
        const long long lin = ((long long)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 256 + threadIdx.x;
        if (lin < N) packed[lin] = kPackedEmpty;
    }
}

// Frames already stored as rows [N, D] (the kmeans labeller flattens before cdist).
__global__ void __launch_bounds__(256)
frames_prepare_rows_kernel(const float* __restrict__ x, int D, int Dp, long long N,
                           __nv_bfloat16* __restrict__ xb, float* __restrict__ xr, long long* __restrict__ packed) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long pairs = N * (Dp / 2);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long p = t0; p < pairs; p += stride) {
        const long long n = p / (Dp / 2);
        const int d = (int)(p - n * (Dp / 2)) * 2;
        const float a = d < D ? __ldg(x + n * D + d) : 0.f;
        const float b = d + 1 < D ? __ldg(x + n * D + d + 1) : 0.f;
        *reinterpret_cast<__nv_bfloat162*>(xb + n * Dp + d) = __floats2bfloat162_rn(a, b);
        if (xr && xr != x) {
            if (d < D) xr[n * D + d] = a;
            if (d + 1 < D) xr[n * D + d + 1] = b;
        }
    }
    if (packed) for (long long n = t0; n < N; n += stride) packed[n] = kPackedEmpty;
}

__global__ void packed_init_kernel(long long* packed, long long N) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) packed[i] = kPackedEmpty;
}

__global__ void unpack_kernel(const long long* __restrict__ packed, long long N,
                              long long* __restrict__ idx, float* __restrict__ dmin) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const long long p = packed[i];
    if (idx) idx[i] = (long long)((unsigned long long)p & 0xffffffffull);
    if (dmin) dmin[i] = float_from_order_key((int32_t)(p >> 32));
}

// bit0: CTA pairs (cta_group::2); bit1: resident A row block.  PERO_ASSIGN_VARIANT overrides.
int assign_variant_default(int num_kb) {
    const int env = PERO_KNOB("PERO_ASSIGN_VARIANT", -1);       // dev build only
    int v = env >= 0 ? env : 3;
    if (num_kb > 8) v &= ~2;   // resident A beyond D = 512 leaves no room for the B ring
    return v;
}

int run_assign_gemm(const __nv_bfloat16* xb, long long N, int Dp, const CodebookLayout& cl, const void* codebook,
                    long long K, int index_offset, long long* packed, cudaStream_t stream, int pdl = 0) {
    ArgminEpi::Params ep;
    ep.colvec = reinterpret_cast<const float*>(static_cast<const char*>(codebook) + cl.cnorm_off);
    ep.packed = packed; ep.rows = (int)N; ep.index_offset = index_offset;
    const void* cb = static_cast<const char*>(codebook) + cl.cb_off;
    const int v = assign_variant_default(Dp / kBlockK);
    // Short contractions (D <= 256): two resident A sets, so that the next row block's frames are loaded while the
    // current block is still being multiplied (a worker changes row block every 32 column tiles at K = 8192).
    const bool dbl = (v & 2) && (Dp / kBlockK <= 4) && PERO_KNOB("PERO_A_DOUBLE", 1) != 0;
    const size_t budget = dbl ? kSmemBudget : ((Dp / kBlockK <= 4) ? kSmemBudgetShared : kSmemBudget);
    switch (v & 3) {
        case 0: return launch_gemm_tn<1, 0, ArgminEpi>(xb, (int)N, Dp, cb, (int)K, Dp, Dp, 1, 0, 1, 0, ep, stream, nullptr,
                                                            budget, 0, pdl);
        case 1: return launch_gemm_tn<2, 0, ArgminEpi>(xb, (int)N, Dp, cb, (int)K, Dp, Dp, 1, 0, 1, 0, ep, stream, nullptr,
                                                            budget, 0, pdl);
        case 2: return launch_gemm_tn<1, 1, ArgminEpi>(xb, (int)N, Dp, cb, (int)K, Dp, Dp, 1, 0, 1, 0, ep, stream, nullptr,
                                                            budget, 0, pdl);
        default:
            if (dbl) return launch_gemm_tn<2, 2, ArgminEpi>(xb, (int)N, Dp, cb, (int)K, Dp, Dp, 1, 0, 1, 0, ep, stream, nullptr,
                                                             budget, 0, pdl);
            return launch_gemm_tn<2, 1, ArgminEpi>(xb, (int)N, Dp, cb, (int)K, Dp, Dp, 1, 0, 1, 0, ep, stream, nullptr,
                                                    budget, 0, pdl);
    }
}

}  // namespace pero

using namespace pero;

extern "C" {

size_t pero_vq_codebook_bytes(int64_t K, int64_t D) {
    if (K <= 0 || D <= 0) return 0;
    return codebook_layout(K, D).total;
}

int pero_vq_codebook_prepare(const float* weight, int64_t K, int64_t D, void* codebook, size_t codebook_bytes,
                             pero_stream_t stream) {
    if (!weight || !codebook) return PERO_ERR_NULL;
    if (K <= 0 || D <= 0 || K > (1ll << 30) || D > 65536) return PERO_ERR_BAD_SHAPE;
    const CodebookLayout cl = codebook_layout(K, D);
    if (codebook_bytes < cl.total) return PERO_ERR_WORKSPACE;
    if (reinterpret_cast<uintptr_t>(codebook) & 255) return PERO_ERR_BAD_ALIGN;
    char* base = static_cast<char*>(codebook);
    const int warps = 8;
    codebook_prepare_kernel<<<(unsigned)((cl.Kp + warps - 1) / warps), warps * 32, 0, stream>>>(
        weight, (int)K, (int)D, (int)cl.Dp, (int)cl.Kp, reinterpret_cast<__nv_bfloat16*>(base + cl.cb_off),
        reinterpret_cast<float*>(base + cl.cnorm_off));
    return (int)cudaGetLastError();
}

size_t pero_vq_assign_workspace_bytes(int64_t N, int64_t K, int64_t D) {
    if (N <= 0 || K <= 0 || D <= 0) return 0;
    return assign_ws_layout(N, D).total;
}

int pero_vq_packed_init(int64_t* packed, int64_t N, pero_stream_t stream) {
    if (!packed) return PERO_ERR_NULL;
    if (N <= 0) return N == 0 ? PERO_OK : PERO_ERR_BAD_SHAPE;
    packed_init_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(reinterpret_cast<long long*>(packed), N);
    return (int)cudaGetLastError();
}

int pero_vq_unpack(const int64_t* packed, int64_t N, int64_t* idx, float* dmin, pero_stream_t stream) {
    if (!packed) return PERO_ERR_NULL;
    if (N <= 0) return N == 0 ? PERO_OK : PERO_ERR_BAD_SHAPE;
    unpack_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<const long long*>(packed), N, reinterpret_cast<long long*>(idx), dmin);
    return (int)cudaGetLastError();
}

int pero_vq_assign(const float* x, int64_t n_lines, int64_t frames_per_line, int channels_first, int64_t K,
                   int64_t D, const void* codebook, int64_t index_offset, int64_t* idx, float* dmin,
                   int64_t* packed_io, float* x_rows, void* workspace, size_t workspace_bytes,
                   pero_stream_t stream) {
    if (n_lines < 0 || frames_per_line < 0) return PERO_ERR_BAD_SHAPE;
    const int64_t N = n_lines * frames_per_line;
    if (N == 0) return PERO_OK;
    const bool init_packed = (channels_first & PERO_ASSIGN_INIT_PACKED) != 0;
    channels_first &= 1;
    if (!x || !codebook || !workspace) return PERO_ERR_NULL;
    if (K <= 0 || D <= 0 || N > (1ll << 31) - 256 || K + index_offset > (1ll << 31) - 1 || index_offset < 0 || D > 65536)
        return PERO_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(codebook) & 255))
        return PERO_ERR_BAD_ALIGN;
    const AssignWsLayout wl = assign_ws_layout(N, D);
    if (workspace_bytes < wl.total) return PERO_ERR_WORKSPACE;
    const CodebookLayout cl = codebook_layout(K, D);
    char* ws = static_cast<char*>(workspace);
    __nv_bfloat16* xb = reinterpret_cast<__nv_bfloat16*>(ws + wl.xb_off);
    long long* packed_ws = reinterpret_cast<long long*>(ws + wl.packed_off);
    long long* packed = packed_io ? reinterpret_cast<long long*>(packed_io) : packed_ws;
    long long* packed_reset = packed_io ? (init_packed ? packed : nullptr) : packed_ws;      // reset by the preparation pass
    const int Dp = (int)cl.Dp;

    if (channels_first) {
        if (n_lines > 65535) return PERO_ERR_BAD_SHAPE;
        dim3 grid((unsigned)((frames_per_line + 31) / 32), (unsigned)(Dp / 64), (unsigned)n_lines);
        frames_prepare_cf_kernel<<<grid, 256, 0, stream>>>(x, (int)D, Dp, (int)frames_per_line, N, xb, x_rows,
                                                           packed_reset);
    } else {
        const long long pairs = N * (Dp / 2);
        long long blocks = (pairs + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        frames_prepare_rows_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, (int)D, Dp, N, xb, x_rows,
                                                                         packed_reset);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;

    int rc = run_assign_gemm(xb, N, Dp, cl, codebook, K, (int)index_offset, packed, (cudaStream_t)stream,
                             /*pdl=*/PERO_KNOB("PERO_ASSIGN_EARLY_B", 1) ? 12 : 4);
    if (rc) return rc;
    if (!packed_io && (idx || dmin)) {
        unpack_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(packed, N, reinterpret_cast<long long*>(idx), dmin);
        e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return PERO_OK;
}

int pero_vq_assign_bf16(const void* x_bf16, int64_t N, int64_t K, int64_t D, const void* codebook, int64_t index_offset,
                        int64_t* packed_io, pero_stream_t stream) {
    if (N == 0) return PERO_OK;
    if (!x_bf16 || !codebook || !packed_io) return PERO_ERR_NULL;
    if (N < 0 || K <= 0 || D <= 0 || N > (1ll << 31) - 256 || K + index_offset > (1ll << 31) - 1 || index_offset < 0)
        return PERO_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(x_bf16) & 15) || (reinterpret_cast<uintptr_t>(codebook) & 255)) return PERO_ERR_BAD_ALIGN;
    const CodebookLayout cl = codebook_layout(K, D);
    // Programmatic launch: the set-up (barriers, TMEM, descriptor prefetch) overlaps the tail of whatever runs in front
    // and the first global access waits for it; the kernel behind is released at once (it cannot be co-resident anyway).
    // A label-production loop that assigns batch after batch therefore pays the ~4 us prologue once, not per batch.
    return run_assign_gemm(static_cast<const __nv_bfloat16*>(x_bf16), N, (int)cl.Dp, cl, codebook, K, (int)index_offset,
                           reinterpret_cast<long long*>(packed_io), (cudaStream_t)stream,
                           PERO_KNOB("PERO_ASSIGN_BF16_PDL", 1) ? (1 | 4) : 0);
}

#ifdef PERO_DEV_BUILD
int pero_debug_set_timeline(void* device_buffer, int slots) {
    g_debug_timeline = static_cast<unsigned long long*>(device_buffer);
    g_debug_timeline_slots = device_buffer ? (slots > 0 ? slots : 0) : 0;
    return PERO_OK;
}
#endif

int pero_gemm_tn_bf16(const void* a_bf16, int64_t rows_a, const void* b_bf16, int64_t rows_b, int64_t kd,
                       int variant, int num_splits, float* out, pero_stream_t stream) {
    if (!a_bf16 || !b_bf16 || !out) return PERO_ERR_NULL;
    StoreEpi::Params ep;
    ep.out = out; ep.ld = rows_b; ep.split_stride = rows_a * rows_b; ep.rows = (int)rows_a; ep.cols = (int)rows_b;
    const int ra = (int)rows_a, rb = (int)rows_b, k = (int)kd;
#ifdef PERO_DEV_BUILD
    if (variant & 16) {          // timeline: `out` receives clock64 stamps [unit][8] of worker 0 (u64)
        NullEpi::Params np; np.out = nullptr;
        unsigned long long* tl = reinterpret_cast<unsigned long long*>(out);
        const bool pair = variant & 1, res = variant & 2;
        if (pair && res) return launch_gemm_tn<2, 1, NullEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, np, stream, tl);
        if (pair) return launch_gemm_tn<2, 0, NullEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, np, stream, tl);
        if (res) return launch_gemm_tn<1, 1, NullEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, np, stream, tl);
        return launch_gemm_tn<1, 0, NullEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, np, stream, tl);
    }
    if (variant & 12) {          // measurement only: bit2 = no TMEM reads at all, bit3 = TMEM reads without math
        NullEpi::Params np; np.out = out;
        LoadEpi::Params lp; lp.out = out;
        const bool pair = variant & 1, res = variant & 2, load = variant & 8;
        if (load) {
            if (pair && res) return launch_gemm_tn<2, 1, LoadEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, lp, stream);
            if (pair) return launch_gemm_tn<2, 0, LoadEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, lp, stream);
            if (res) return launch_gemm_tn<1, 1, LoadEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, lp, stream);
            return launch_gemm_tn<1, 0, LoadEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, lp, stream);
        }
        if (pair && res) return launch_gemm_tn<2, 1, NullEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, np, stream);
        if (pair) return launch_gemm_tn<2, 0, NullEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, np, stream);
        if (res) return launch_gemm_tn<1, 1, NullEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, np, stream);
        return launch_gemm_tn<1, 0, NullEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, np, stream);
    }
#else
    if (variant & (4 | 8 | 16)) return PERO_ERR_UNSUPPORTED;      // measurement variants exist in the dev build only
#endif
    if (variant & 32) {          // MN-major operands: a is [kd, rows_a], b is [kd, rows_b]; out = a^T b
        const int kp = (k + 63) / 64 * 64;
        if (variant & 1)
            return launch_gemm_tn<2, 0, StoreEpi, 3>(a_bf16, ra, ra, b_bf16, rb, rb, kp, num_splits, 0, 1, 0, ep, stream,
                                                            nullptr, kSmemBudget, k);
        return launch_gemm_tn<1, 0, StoreEpi, 3>(a_bf16, ra, ra, b_bf16, rb, rb, kp, num_splits, 0, 1, 0, ep, stream, nullptr,
                                                        kSmemBudget, k);
    }
    switch (variant & 3) {
        case 0: return launch_gemm_tn<1, 0, StoreEpi>(a_bf16, ra, k, b_bf16, rb, k, k, num_splits, 0, 1, 0, ep, stream);
        case 1: return launch_gemm_tn<2, 0, StoreEpi>(a_bf16, ra, k, b_bf16, rb, k, k, num_splits, 0, 1, 0, ep, stream);
        case 2: return launch_gemm_tn<1, 1, StoreEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, ep, stream);
        default: return launch_gemm_tn<2, 1, StoreEpi>(a_bf16, ra, k, b_bf16, rb, k, k, 1, 0, 1, 0, ep, stream);
    }
}

}  // extern "C"
