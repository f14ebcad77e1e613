// Host side of gemm_core.cuh: TMA tensor-map construction (driver entry point fetched through the
// runtime, so the library does not link libcuda) and the launch helper that sizes shared memory,
// the stage ring and the grid.
#pragma once
#include "gemm_core.cuh"
#include "errors.h"
#include "knobs.h"
#include <atomic>
#include <stdlib.h>

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

// Row-major bf16 matrix [rows, cols] with row pitch `pitch_elems` (multiple of 8 elements), tiled in
// boxes of {64 columns, box_rows rows}, SWIZZLE_128B, out-of-bounds elements read as zero.
inline int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                          uint32_t box_rows) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return PERO_ERR_DRIVER;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || (pitch_elems & 7u) || box_rows == 0 || box_rows > 256)
        return PERO_ERR_BAD_ALIGN;
    // The encode is a pure function of its arguments and costs ~3 us of host time per call; a training loop asks
    // for the same few descriptors every step (the caching allocator hands back the same addresses), so the last
    // results are memoised per thread.  Nothing here refers to device state: a stale entry cannot exist.
    struct Key { const void* base; uint64_t rows, cols, pitch; uint32_t box_rows; };
    struct Entry { Key k; CUtensorMap map; bool valid; };
    constexpr int kSlots = 32;
    static thread_local Entry cache[kSlots] = {};
    const Key key{base, rows, cols, pitch_elems, box_rows};
    uint64_t hsh = reinterpret_cast<uintptr_t>(base) >> 8;
    hsh = (hsh ^ (rows * 0x9E3779B97F4A7C15ull) ^ (cols << 17) ^ (pitch_elems << 29) ^ box_rows) * 0xD6E8FEB86659FD93ull;
    Entry& e = cache[(hsh >> 40) % kSlots];
    if (e.valid && e.k.base == key.base && e.k.rows == key.rows && e.k.cols == key.cols && e.k.pitch == key.pitch &&
        e.k.box_rows == key.box_rows) {
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
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return PERO_ERR_DRIVER;
    e.k = key; e.map = *out; e.valid = true;
    return PERO_OK;
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
inline int pick_stages(int cta_group, bool a_resident, int num_kb, int scratch_per_warp, size_t budget = kSmemBudget) {
    for (int s = kMaxStages; s >= 2; --s)
        if (gemm_smem_bytes(cta_group, a_resident, num_kb, s, scratch_per_warp) <= budget) return s;
    return 0;
}

// A: [rows_a, kd] bf16, pitch_a elements; B: [rows_b, kd] bf16, pitch_b elements; kd % 64 == 0.
// `workers` = number of CTAs (kCtaGroup == 1) or CTA pairs (== 2); 0 = fill the machine.
// kMnMajor: A is stored [k_rows, rows_a] and B [k_rows, rows_b] (row pitches pitch_a / pitch_b), C = A^T B; kd is the
// contraction length rounded up to 64 and k_rows (<= kd, 0 = kd) the rows that exist (TMA zero-fills the rest).
template <int kCtaGroup, bool kAResident, class Epi, bool kMnMajor = false>
int launch_gemm_tn(const void* a, int rows_a, int pitch_a, const void* b, int rows_b, int pitch_b, int kd,
                   int num_ks, int split_mode, int fixed_s, int workers, const typename Epi::Params& ep,
                   cudaStream_t stream, unsigned long long* timeline = nullptr, size_t smem_budget = kSmemBudget,
                   int k_rows = 0, int pdl = 0) {
    if (rows_a <= 0 || rows_b <= 0 || kd <= 0 || (kd % kBlockK) != 0) return PERO_ERR_BAD_SHAPE;
    GemmShape sh;
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
    if (kAResident && (sh.num_ks != 1 || sh.num_kb > kMaxAKb)) return PERO_ERR_BAD_SHAPE;
    sh.num_stages = pick_stages(kCtaGroup, kAResident, sh.num_kb, Epi::kScratchPerWarp, smem_budget);
    if (sh.num_stages < 2) return PERO_ERR_BAD_SHAPE;

    CUtensorMap ta, tb;
    int rc;
    if constexpr (kMnMajor) {
        const uint64_t kr = (uint64_t)(k_rows > 0 ? k_rows : kd);
        rc = make_tmap_bf16(&ta, a, kr, (uint64_t)rows_a, (uint64_t)pitch_a, 64);
        if (rc) return rc;
        rc = make_tmap_bf16(&tb, b, kr, (uint64_t)rows_b, (uint64_t)pitch_b, 64);
    } else {
        rc = make_tmap_bf16(&ta, a, (uint64_t)rows_a, (uint64_t)kd, (uint64_t)pitch_a, kBlockM);
        if (rc) return rc;
        rc = make_tmap_bf16(&tb, b, (uint64_t)rows_b, (uint64_t)kd, (uint64_t)pitch_b, kBlockN / kCtaGroup);
    }
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

    size_t smem = gemm_smem_bytes(kCtaGroup, kAResident, sh.num_kb, sh.num_stages, Epi::kScratchPerWarp);
    if (smem < kSmemFloor) smem = kSmemFloor;
    auto kern = gemm_tn_kernel<kCtaGroup, kAResident, Epi, kMnMajor>;
    {
        static std::atomic<bool> attr_done[kMaxDevices];
        cudaError_t e = ensure_max_dynamic_smem(kern, attr_done, kSmemBudget);
        if (e != cudaSuccess) return (int)e;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(workers * kCtaGroup));
    cfg.blockDim = dim3(kGemmThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kCtaGroup; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (pdl & 6) {      // may start while the previous kernel of the stream is still running (see GemmShape::pdl)
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.numAttrs = 2;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ta, tb, sh, ep);
    return e == cudaSuccess ? PERO_OK : (int)e;
}

}  // namespace pero
