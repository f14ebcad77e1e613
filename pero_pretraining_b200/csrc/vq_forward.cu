// VectorQuantizer.forward as ONE call of the C ABI (models/autoencoders.py:204-241): assign -> quantize with the
// straight-through value -> (training, decay > 0) EMA accumulate + apply.  Host code only: it carves one
// caller-owned workspace and enqueues the same kernels as the stage-level entry points, so the Python module
// crosses the ABI once per forward instead of four times.
#include <cuda_runtime.h>

#include "../../include/pero_b200.h"
#include "layout.h"

using namespace pero;

namespace {
struct FwdLayout { size_t assign_off, assign_bytes, rows_off, sums_off, ema_off, ema_bytes, total; };
FwdLayout fwd_layout(int64_t N, int64_t K, int64_t D, bool ema) {
    FwdLayout l;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    l.assign_bytes = pero_vq_assign_workspace_bytes(N, K, D);
    l.assign_off = take(l.assign_bytes);
    l.rows_off = take((size_t)N * D * 4);
    l.ema_bytes = ema ? pero_vq_ema_workspace_bytes(N, K, D) : 0;
    l.sums_off = take(ema ? ((size_t)K * D + K) * 4 : 0);
    l.ema_off = take(l.ema_bytes);
    l.total = off;
    return l;
}
}  // namespace

extern "C" {

size_t pero_vq_forward_workspace_bytes(int64_t N, int64_t K, int64_t D, int update_ema) {
    if (N <= 0 || K <= 0 || D <= 0) return 0;
    return fwd_layout(N, K, D, update_ema != 0).total;
}

int pero_vq_forward(const float* x, int64_t n_lines, int64_t frames_per_line, int channels_first, int64_t K, int64_t D,
                    void* codebook, size_t codebook_bytes, float* weight, float* ema_w, float* ema_cluster_size, double decay,
                    double epsilon, int update_ema, float* quantized, int64_t* idx, void* workspace, size_t workspace_bytes,
                    pero_stream_t stream) {
    if (!x || !codebook || !weight || !quantized || !idx || !workspace) return PERO_ERR_NULL;
    if (update_ema && (!ema_w || !ema_cluster_size)) return PERO_ERR_NULL;
    const int64_t N = n_lines * frames_per_line;
    if (n_lines <= 0 || frames_per_line <= 0 || K <= 0 || D <= 0) return PERO_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return PERO_ERR_BAD_ALIGN;
    const FwdLayout l = fwd_layout(N, K, D, update_ema != 0);
    if (workspace_bytes < l.total) return PERO_ERR_WORKSPACE;
    if (codebook_bytes < pero_vq_codebook_bytes(K, D)) return PERO_ERR_WORKSPACE;
    char* ws = static_cast<char*>(workspace);
    float* x_rows = reinterpret_cast<float*>(ws + l.rows_off);
    int rc = pero_vq_assign(x, n_lines, frames_per_line, channels_first, K, D, codebook, 0, idx, nullptr, nullptr, x_rows,
                            ws + l.assign_off, l.assign_bytes, stream);
    if (rc) return rc;
    // the quantized value is gathered from the codebook as it was for this step's assignment (:222), before the
    // EMA update below replaces it (:235-237)
    rc = pero_vq_gather_st(x_rows, idx, weight, n_lines, frames_per_line, channels_first, K, D, quantized, stream);
    if (rc || !update_ema) return rc;
    float* sums = reinterpret_cast<float*>(ws + l.sums_off);
    rc = pero_vq_ema_accumulate(x_rows, idx, N, K, D, sums, ws + l.ema_off, l.ema_bytes, stream);
    if (rc) return rc;
    return pero_vq_ema_apply(sums, K, D, decay, epsilon, ema_w, ema_cluster_size, weight, codebook, codebook_bytes,
                             ws + l.ema_off, l.ema_bytes, stream);
}

}  // extern "C"
