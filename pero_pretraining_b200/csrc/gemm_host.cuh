// Host side of gemm_core.cuh: TMA tensor-map construction (driver entry point fetched through the
// runtime, so the library does not link libcuda) and the launch helper that sizes shared memory,
// the stage ring and the grid.
#pragma once
#include "gemm_core.cuh"
#include "errors.h"
#include "knobs.h"
#include <atomic>
#include <stdlib.h>
#include <string.h>

namespace pero {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// Tiled tensor map (rank 2 or 3), SWIZZLE_128B, out-of-bounds elements read as zero / not written.
// The encode is a pure function of its arguments and costs ~3 us of host time per call; a training loop asks for the
// same few descriptors every step (the caching allocator hands back the same addresses), so the last results are
// memoised per thread.  Nothing here refers to device state: a stale entry cannot exist.
inline int make_tmap(CUtensorMap* out, CUtensorMapDataType dtype, int elem_bytes, const void* base, int rank,
                     const uint64_t* dims, const uint64_t* stride_elems /* rank - 1 entries */, const uint32_t* box) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return PERO_ERR_DRIVER;
    if (reinterpret_cast<uintptr_t>(base) & 15u) return PERO_ERR_BAD_ALIGN;
    for (int i = 0; i + 1 < rank; ++i)
        if ((stride_elems[i] * elem_bytes) & 15u) return PERO_ERR_BAD_ALIGN;
    struct Key { const void* base; uint64_t d[3], s[2]; uint32_t b[3]; int rank, dtype; };
    struct Entry { Key k; CUtensorMap map; bool valid; };
    constexpr int kSlots = 32;
    static thread_local Entry cache[kSlots] = {};
    Key key = {};
    key.base = base; key.rank = rank; key.dtype = (int)dtype;
    for (int i = 0; i < rank; ++i) { key.d[i] = dims[i]; key.b[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) key.s[i] = stride_elems[i];
    uint64_t hsh = reinterpret_cast<uintptr_t>(base) >> 8;
    hsh = (hsh ^ (key.d[1] * 0x9E3779B97F4A7C15ull) ^ (key.d[0] << 17) ^ (key.s[0] << 29) ^ key.b[1] ^ ((uint64_t)dtype << 50) ^
           (key.d[2] << 7)) * 0xD6E8FEB86659FD93ull;
    Entry& e = cache[(hsh >> 40) % kSlots];
    if (e.valid && memcmp(&e.k, &key, sizeof(Key)) == 0) {
        *out = e.map;
        return PERO_OK;
    }
    // The encode is a DRIVER call and needs the primary context current on the calling thread.  A thread whose
    // first CUDA action is this call (e.g. PyTorch's autograd worker entering a backward that starts with a
    // GEMM) has none yet: cudaSetDevice binds it (legal during stream capture, once per thread).
    static thread_local bool context_bound = false;
    if (!context_bound) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaSetDevice(dev) != cudaSuccess) return PERO_ERR_DRIVER;
        context_bound = true;
    }
    cuuint64_t gd[3] = {1, 1, 1};
    cuuint64_t gs[2] = {0, 0};
    cuuint32_t bx[3] = {1, 1, 1}, estr[3] = {1, 1, 1};
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = stride_elems[i] * elem_bytes;
    CUresult r = enc(out, dtype, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return PERO_ERR_DRIVER;
    e.k = key; e.map = *out; e.valid = true;
    return PERO_OK;
}

// Row-major bf16 matrix [rows, cols] with row pitch `pitch_elems` (multiple of 8 elements), tiled in
// boxes of {64 columns, box_rows rows}.
inline int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                          uint32_t box_rows) {
    if ((pitch_elems & 7u) || box_rows == 0 || box_rows > 256) return PERO_ERR_BAD_ALIGN;
    const uint64_t dims[2] = {cols, rows}, strides[1] = {pitch_elems};
    const uint32_t box[2] = {(uint32_t)kBlockK, box_rows};
    return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, 2, dims, strides, box);
}

// fp32 output [planes][rows][cols] (row pitch ld, plane pitch plane_stride, both multiples of 4 elements), stored in
// boxes of {32 columns = 128 B, 32 rows, 1 plane}: the staging tile of StoreTmaEpi.
inline int make_tmap_f32_store(CUtensorMap* out, const void* base, uint64_t planes, uint64_t rows, uint64_t cols, uint64_t ld,
                               uint64_t plane_stride) {
    const uint64_t dims[3] = {cols, rows, planes}, strides[2] = {ld, plane_stride};
    const uint32_t box[3] = {32, 32, 1};
    return make_tmap(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, 3, dims, strides, box);
}

constexpr int kMaxDevices = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < kMaxDevices) ? dev : 0;
}

// Memo of an immutable device property (per device; racing first calls store the same value).
inline int device_sm_count() {
    static std::atomic<int> sms[kMaxDevices];
    const int dev = current_device();
    int v = sms[dev].load(std::memory_order_relaxed);
    if (v <= 0) {
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        if (v <= 0) v = 148;
        sms[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device, idempotent setting: it is applied once per
// (kernel, device) and remembered in a per-kernel table of atomics (thread-safe; no other global state).
template <class Kernel>
inline cudaError_t ensure_max_dynamic_smem(Kernel kern, std::atomic<bool>* done_per_device, size_t bytes) {
    const int dev = current_device();
    if (done_per_device[dev].load(std::memory_order_acquire)) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) done_per_device[dev].store(true, std::memory_order_release);
    return e;
}

#ifdef PERO_DEV_BUILD
// Dev build only (pero_debug_set_timeline): while slots remain, every GEMM launch records worker 0's clock64 stamps
// in the next 64 KiB slot of the caller's buffer; once the slots are used up, launches run without stamps.
inline unsigned long long* g_debug_timeline = nullptr;
inline int g_debug_timeline_slots = 0;
#endif

constexpr size_t kSmemBudget = 227 * 1024;
// Budget for GEMMs that run beside other chains of the step (masked CE): leaves ~27 KB of shared memory and, with
// the 128-register cap of the kernel, 16 K registers per SM for a co-resident bandwidth-bound or exchange CTA.
constexpr size_t kSmemBudgetShared = 200 * 1024;
constexpr size_t kSmemFloor = 120 * 1024;   // > half an SM: never two TMEM-hungry CTAs on one SM

// Picks the deepest ring that fits; returns 0 when even 2 stages do not fit.
inline int pick_stages(int cta_group, int a_sets, int num_kb, int scratch_per_warp, size_t budget = kSmemBudget) {
    for (int s = kMaxStages; s >= 2; --s)
        if (gemm_smem_bytes(cta_group, a_sets, num_kb, s, scratch_per_warp) <= budget) return s;
    return 0;
}

// A: [rows_a, kd] bf16, pitch_a elements; B: [rows_b, kd] bf16, pitch_b elements; kd % 64 == 0.
// `workers` = number of CTAs (kCtaGroup == 1) or CTA pairs (== 2); 0 = fill the machine.
// kMajor bit 0 / bit 1: A / B is MN-major, i.e. stored [k_rows, rows_a] / [k_rows, rows_b] (row pitch pitch_a / pitch_b);
// both bits: C = A^T B.  kd is the
// contraction length rounded up to 64 and k_rows (<= kd, 0 = kd) the rows that exist (TMA zero-fills the rest).
// kARes: 0 = A streams through the ring with B, 1 = resident A row block, 2 = two resident A sets (num_kb <= 6).
struct GemmLaunch {
    CUtensorMap ta, tb;
    GemmShape sh;
    int workers;
    size_t smem;
};

template <int kCtaGroup, int kARes, class Epi, int kMajor = 0>
int prepare_gemm_tn(GemmLaunch& g, const void* a, int rows_a, int pitch_a, const void* b, int rows_b, int pitch_b, int kd,
                    int num_ks, int split_mode, int fixed_s, int workers, unsigned long long* timeline = nullptr,
                    size_t smem_budget = kSmemBudget, int k_rows = 0, int pdl = 0) {
    if (rows_a <= 0 || rows_b <= 0 || kd <= 0 || (kd % kBlockK) != 0) return PERO_ERR_BAD_SHAPE;
    GemmShape& sh = g.sh;
    sh.timeline = timeline;
#ifdef PERO_DEV_BUILD
    if (!timeline && g_debug_timeline && g_debug_timeline_slots > 0) {     // debug: one 64 KiB slot per GEMM launch
        sh.timeline = g_debug_timeline;
        g_debug_timeline += 8192;
        --g_debug_timeline_slots;
    }
#endif
    sh.rows_a = rows_a; sh.rows_b = rows_b;
    sh.num_kb = kd / kBlockK;
    sh.num_rb = (rows_a + kBlockM * kCtaGroup - 1) / (kBlockM * kCtaGroup);
    sh.num_ct = (rows_b + kBlockN - 1) / kBlockN;
    sh.num_ks = num_ks < 1 ? 1 : num_ks;
    sh.kb_per_split = (sh.num_kb + sh.num_ks - 1) / sh.num_ks;
    sh.num_ks = (sh.num_kb + sh.kb_per_split - 1) / sh.kb_per_split;   // no empty splits
    sh.split_mode = split_mode; sh.fixed_s = fixed_s < 1 ? 1 : fixed_s;
    if (PERO_KNOB("PERO_PDL", 1) == 0) pdl = 0;       // dev build: no programmatic dependent launches at all
    sh.pdl = pdl;
#ifdef PERO_DEV_BUILD
    sh.fake_b = PERO_KNOB("PERO_GEMM_FAKE_B", 0);
#endif
    constexpr int kASets = kARes == 2 ? 2 : (kARes ? 1 : 0);
    if (kARes && (sh.num_ks != 1 || sh.num_kb * kASets > kMaxAKb)) return PERO_ERR_BAD_SHAPE;
    sh.num_stages = pick_stages(kCtaGroup, kASets, sh.num_kb, Epi::kScratchPerWarp, smem_budget);
    {
        const int cap = PERO_KNOB("PERO_GEMM_STAGES", 0);        // dev build: ring-depth experiments
        if (cap >= 2 && cap < sh.num_stages) sh.num_stages = cap;
    }
    if (sh.num_stages < 2) return PERO_ERR_BAD_SHAPE;

    int rc;
    const uint64_t kr = (uint64_t)(k_rows > 0 ? k_rows : kd);
    if constexpr (kMajor & 1) rc = make_tmap_bf16(&g.ta, a, kr, (uint64_t)rows_a, (uint64_t)pitch_a, 64);
    else rc = make_tmap_bf16(&g.ta, a, (uint64_t)rows_a, kr, (uint64_t)pitch_a, kBlockM);
    if (rc) return rc;
    if constexpr (kMajor & 2) rc = make_tmap_bf16(&g.tb, b, kr, (uint64_t)rows_b, (uint64_t)pitch_b, 64);
    else rc = make_tmap_bf16(&g.tb, b, (uint64_t)rows_b, kr, (uint64_t)pitch_b, kBlockN / kCtaGroup);
    if (rc) return rc;

    const long long units = (long long)sh.num_rb * sh.num_ct * sh.num_ks;
    int max_workers = device_sm_count() / kCtaGroup;
    {
        const int cap = PERO_KNOB("PERO_GEMM_MAX_CTAS", -1);      // dev build: co-residency experiments
        if (cap > 0 && cap / kCtaGroup < max_workers) max_workers = cap / kCtaGroup;
    }
    if (split_mode == 1) workers = sh.num_rb * sh.fixed_s;
    else if (workers <= 0 || workers > max_workers) workers = max_workers;
    if (split_mode == 0 && workers > units) workers = (int)units;
    g.workers = workers;
    g.smem = gemm_smem_bytes(kCtaGroup, kASets, sh.num_kb, sh.num_stages, Epi::kScratchPerWarp);
    if (g.smem < kSmemFloor) g.smem = kSmemFloor;
    return PERO_OK;
}

template <int kCtaGroup, int kARes, class Epi, int kMajor = 0>
int launch_gemm_tn(const void* a, int rows_a, int pitch_a, const void* b, int rows_b, int pitch_b, int kd,
                   int num_ks, int split_mode, int fixed_s, int workers, const typename Epi::Params& ep,
                   cudaStream_t stream, unsigned long long* timeline = nullptr, size_t smem_budget = kSmemBudget,
                   int k_rows = 0, int pdl = 0) {
    GemmLaunch g;
    int rc = prepare_gemm_tn<kCtaGroup, kARes, Epi, kMajor>(g, a, rows_a, pitch_a, b, rows_b, pitch_b, kd, num_ks, split_mode, fixed_s,
                                                            workers, timeline, smem_budget, k_rows, pdl);
    if (rc) return rc;
    auto kern = gemm_tn_kernel<kCtaGroup, kARes, Epi, kMajor>;
    {
        static std::atomic<bool> attr_done[kMaxDevices];
        cudaError_t e = ensure_max_dynamic_smem(kern, attr_done, kSmemBudget);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(g.workers * kCtaGroup));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtaGroup; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (g.sh.pdl & 6) {      // may start while the previous kernel of the stream is still running (see GemmShape::pdl)
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, g.ta, g.tb, g.sh, ep);
    return e == cudaSuccess ? PERO_OK : (int)e;
}

// Two prepared streamed pair-GEMMs (prepare_gemm_tn<2, 0, Epi, kMajor1 / kMajor2>) in one grid (gemm_dual_kernel): the
// first g1.workers pairs run GEMM 1, the next g2.workers pairs GEMM 2.  The launch is programmatic when either shape
// asks for it (pdl bits 1 / 2).
template <class Epi, int kMajor1, int kMajor2>
int launch_gemm_dual(const GemmLaunch& g1, const typename Epi::Params& ep1, const GemmLaunch& g2, const typename Epi::Params& ep2,
                     cudaStream_t stream) {
    auto kern = gemm_dual_kernel<Epi, kMajor1, kMajor2>;
    {
        static std::atomic<bool> attr_done[kMaxDevices];
        cudaError_t e = ensure_max_dynamic_smem(kern, attr_done, kSmemBudget);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((g1.workers + g2.workers) * 2));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = g1.smem > g2.smem ? g1.smem : g2.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if ((g1.sh.pdl | g2.sh.pdl) & 6) {
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, g1.ta, g1.tb, g1.sh, ep1, g2.ta, g2.tb, g2.sh, ep2, g1.workers * 2);
    return e == cudaSuccess ? PERO_OK : (int)e;
}

}  // namespace pero
