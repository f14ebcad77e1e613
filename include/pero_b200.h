/* pero_b200.h — C ABI of the B200-native quantize-and-predict path of DCGM/pero-pretraining.
 *
 * The reference has no FFI layer: its boundary for this path is a set of torch.nn.Module methods
 * (SURVEY.md §8b).  Every entry point below replaces the body of one of those methods (or a
 * group of torch ops inside it) and cites it as  file:line  relative to the reference root.
 * The Python host side (pero_pretraining_b200/) keeps the reference's module signatures and calls
 * these functions through ctypes; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless stated; sizes are int64_t; no torch types.
 *   - The caller owns every buffer including workspaces; the library allocates nothing persistent,
 *     keeps no mutable state between calls (only per-device memos of immutable facts: SM count, "shared-memory
 *     attribute already set", and a per-thread memo of encoded TMA descriptors) and is re-entrant.  All work is enqueued on `stream`; no call
 *     synchronises, so sequences of calls are CUDA-graph capturable.
 *   - Return 0 on success; negative = PERO_ERR_* below; positive = a cudaError_t.  Never throws, never
 *     prints.  pero_strerror() names any code.
 *   - "frames" are the rows the path works on: N = n_lines * frames_per_line feature vectors of
 *     dimension D (one per 8-px column of a 40-px text line); "codebook" is K codewords x D.
 */
#ifndef PERO_B200_H_
#define PERO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* pero_stream_t;

#define PERO_OK 0
#define PERO_ERR_BAD_SHAPE (-1)
#define PERO_ERR_BAD_ALIGN (-2)
#define PERO_ERR_WORKSPACE (-3)
#define PERO_ERR_ARCH (-4)
#define PERO_ERR_NULL (-5)
#define PERO_ERR_DRIVER (-6)
#define PERO_ERR_UNSUPPORTED (-7)

int pero_version(void);
const char* pero_strerror(int code);
/* 0 when the current device is sm_100 (B200); PERO_ERR_ARCH otherwise. */
int pero_check_device(void);

/* ------------------------------------------------------------------ nearest-codeword assignment
 * Replaces  models/autoencoders.py:212-217  (distances + argmin of VectorQuantizer.forward) and
 *           scripts/produce_kmeans_labels.py:72-76 (cdist + argmin of the FQ / PQ-AE labeller).
 *
 * pero_vq_codebook_prepare: fp32 codebook [K, D] -> opaque device blob holding the bf16 operand
 *   (rows padded to a multiple of 64 columns) and |c|^2 in fp32 computed from the fp32 weights.
 *   Re-run whenever the codebook changes (pero_vq_ema_apply refreshes it itself).
 */
size_t pero_vq_codebook_bytes(int64_t K, int64_t D);
int pero_vq_codebook_prepare(const float* weight, int64_t K, int64_t D, void* codebook, size_t codebook_bytes,
                             pero_stream_t stream);

/* pero_vq_assign: for every frame, the index of the nearest codeword (lowest index on exact ties).
 *   x               fp32 frames; channels_first = 1: [n_lines, D, frames_per_line] (the NCHW tensor the
 *                   quantizer receives, H*W collapsed; autoencoders.py:205-209 permutes it),
 *                   channels_first = 0: [n_lines * frames_per_line, D] rows (kmeans labeller).
 *                   Bit 1 (PERO_ASSIGN_INIT_PACKED = 2, or'ed in): `packed_io` is reset to "empty" by the frame
 *                   preparation pass of this call (the first / only shard of a merge): no pero_vq_packed_init launch.
 *   index_offset    added to every index (codebook shard k0 when the codebook is sharded).
 *   idx   [N] int64 or NULL, dmin [N] fp32 or NULL (|c|^2 - 2<x,c> of the winner, i.e. the squared
 *                   distance minus |x|^2), written only when `packed_io` is NULL.
 *   packed_io [N] int64 or NULL: when given, results are min-merged into it as the SIGNED word
 *                   (order_key(dmin) << 32 | index) and idx/dmin are left untouched — the caller
 *                   all-reduces it with a plain int64 MIN over codebook shards and calls
 *                   pero_vq_unpack.  Must be pre-set to INT64_MAX (pero_vq_packed_init).
 *   x_rows [N, D] fp32 or NULL: row-major copy of the frames for the gather / EMA stages.
 */
#define PERO_ASSIGN_INIT_PACKED 2
size_t pero_vq_assign_workspace_bytes(int64_t N, int64_t K, int64_t D);
int pero_vq_assign(const float* x, int64_t n_lines, int64_t frames_per_line, int channels_first, int64_t K,
                   int64_t D, const void* codebook, int64_t index_offset, int64_t* idx, float* dmin,
                   int64_t* packed_io, float* x_rows, void* workspace, size_t workspace_bytes,
                   pero_stream_t stream);
/* The distance GEMM + arg-min alone, for callers that already hold the frames as bf16 rows
 * [N, Dp] (Dp = D rounded up to 64, zero padded): min-merges into packed_io (see above). */
int pero_vq_assign_bf16(const void* x_bf16, int64_t N, int64_t K, int64_t D, const void* codebook, int64_t index_offset,
                        int64_t* packed_io, pero_stream_t stream);
int pero_vq_packed_init(int64_t* packed, int64_t N, pero_stream_t stream);
int pero_vq_unpack(const int64_t* packed, int64_t N, int64_t* idx, float* dmin, pero_stream_t stream);

/* ------------------------------------------------------------------ quantize + straight-through
 * Replaces  models/autoencoders.py:218-222, 239-241  (one-hot, encodings @ weight, straight-through,
 * permute back).  out = x + (weight[idx] - x), evaluated in fp32 exactly as the reference does,
 * written channels-first [n_lines, D, frames_per_line] (or as rows when channels_first = 0).
 */
int pero_vq_gather_st(const float* x_rows, const int64_t* idx, const float* weight, int64_t n_lines,
                      int64_t frames_per_line, int channels_first, int64_t K, int64_t D, float* out,
                      pero_stream_t stream);
/* The same pass, also producing the quantisation loss of  models/autoencoders.py:193-202  (calculate_loss):
 *     m = mean((out - x)^2),  loss_out[0] = scale_a * m + scale_b * m   (each product rounded to fp32, see pero_mse_fwd)
 * from per-block partial sums combined in a fixed order: `out` and the frames are not read a second time. */
size_t pero_vq_gather_st_mse_workspace_bytes(int64_t n_lines, int64_t frames_per_line, int channels_first, int64_t D);
int pero_vq_gather_st_mse(const float* x_rows, const int64_t* idx, const float* weight, int64_t n_lines,
                          int64_t frames_per_line, int channels_first, int64_t K, int64_t D, float* out, float scale_a,
                          float scale_b, float* loss_out, void* workspace, size_t workspace_bytes, pero_stream_t stream);

/* ------------------------------------------------------------------ EMA codebook update
 * Replaces  models/autoencoders.py:225-237.
 * pero_vq_ema_accumulate: deterministic sort-based segmented sum:
 *     sums[k, :] = sum of x_rows[n, :] over frames with idx[n] == k (ascending n),  counts[k] = #frames.
 *   `sums_counts` is ONE contiguous fp32 buffer [K*D + K] (sums then counts) so that data-parallel
 *   ranks all-reduce it with a single SUM before pero_vq_ema_apply.
 * pero_vq_ema_apply:
 *     cs <- cs*decay + (1-decay)*counts;  n = sum(cs);  cs <- (cs + eps) / (n + K*eps) * n
 *     ema_w <- ema_w*decay + (1-decay)*sums;  weight <- ema_w / cs[:, None]
 *   and, when `codebook` is not NULL, refreshes the prepared blob for the next assign.
 *   Its workspace needs 256 bytes (any pero_vq_ema_workspace_bytes() buffer is large enough).
 */
size_t pero_vq_ema_workspace_bytes(int64_t N, int64_t K, int64_t D);
int pero_vq_ema_accumulate(const float* x_rows, const int64_t* idx, int64_t N, int64_t K, int64_t D,
                           float* sums_counts, void* workspace, size_t workspace_bytes, pero_stream_t stream);
int pero_vq_ema_apply(const float* sums_counts, int64_t K, int64_t D, double decay, double epsilon, float* ema_w,
                      float* ema_cluster_size, float* weight, void* codebook, size_t codebook_bytes,
                      void* workspace, size_t workspace_bytes, pero_stream_t stream);
/* counts[k] = #frames with idx == k as int64 (models/autoencoders.py:165, torch.bincount). */
int pero_vq_counts(const int64_t* idx, int64_t N, int64_t K, int64_t* counts, pero_stream_t stream);

/* ------------------------------------------------------------------ mini-batch k-means centre update (codebook FIT)
 * The step of scikit-learn's MiniBatchKMeans (sklearn/cluster/_k_means_minibatch.pyx, update_center_dense — the
 * fitter behind  scripts/fit_kmeans.py:20-32 ; scikit-learn is an unpinned dependency of the reference, 1.9.0 here)
 * after the batch has been assigned (pero_vq_assign) and summed per centre (pero_vq_ema_accumulate):
 *     for centres with n > 0 members:  c <- (c * w + sum) * (1 / (w + n)),  w <- w + n
 * centers [K, D] and weight_sums [K] are updated in place; `codebook` (or NULL) is the prepared blob to refresh. */
int pero_kmeans_update(const float* sums_counts, int64_t K, int64_t D, float* centers, float* weight_sums, void* codebook,
                       size_t codebook_bytes, pero_stream_t stream);

/* ------------------------------------------------------------------ the whole quantizer forward in one call
 * Replaces  models/autoencoders.py:204-241  (VectorQuantizer.forward): pero_vq_assign -> pero_vq_gather_st ->
 * (update_ema != 0: decay > 0 and training) pero_vq_ema_accumulate -> pero_vq_ema_apply, enqueued back to back on
 * `stream` out of ONE workspace.  quantized [n_lines, D, frames_per_line] (or rows), idx [N] int64.  The
 * single-process path; data-parallel callers use the stage-level calls with the exchange between accumulate and
 * apply.  `codebook` is the prepared blob of `weight` (refreshed by the EMA update). */
size_t pero_vq_forward_workspace_bytes(int64_t N, int64_t K, int64_t D, int update_ema);
int pero_vq_forward(const float* x, int64_t n_lines, int64_t frames_per_line, int channels_first, int64_t K, int64_t D,
                    void* codebook, size_t codebook_bytes, float* weight, float* ema_w, float* ema_cluster_size,
                    double decay, double epsilon, int update_ema, float* quantized, int64_t* idx, void* workspace,
                    size_t workspace_bytes, pero_stream_t stream);

/* ------------------------------------------------------------------ commitment / latent loss
 * Replaces  models/autoencoders.py:193-202  (VectorQuantizer.calculate_loss = mse_loss terms).
 * pero_mse_fwd: m = mean((a - b)^2); out[0] = scale_a * m + scale_b * m (each product rounded to fp32,
 *               like q_latent_loss + commitment_cost * e_latent_loss); deterministic two-stage reduction.
 * pero_mse_bwd: g_b = coef * grad_out[0] * (b - a), and g_a = -g_b when g_a != NULL
 *               (coef = 2 * weight / numel).
 */
size_t pero_mse_workspace_bytes(int64_t numel);
int pero_mse_fwd(const float* a, const float* b, int64_t numel, float scale_a, float scale_b, float* out,
                 void* workspace, size_t workspace_bytes, pero_stream_t stream);
int pero_mse_bwd(const float* a, const float* b, int64_t numel, float coef, const float* grad_out, float* g_a,
                 float* g_b, pero_stream_t stream);
/* Backward of VectorQuantizer.forward + calculate_loss(quantized, inputs) in one pass
 * (models/autoencoders.py:239 straight-through, :198 commitment term):
 *   g_inputs = g_quantized + coef * grad_loss[0] * (inputs - quantized),  coef = 2 * commitment_cost / numel.
 * All tensors share one layout (the NCHW layout of the module's input/output). */
int pero_vq_st_commit_bwd(const float* g_quantized, const float* quantized, const float* inputs, int64_t numel, float coef,
                          const float* grad_loss, float* g_inputs, pero_stream_t stream);

/* ------------------------------------------------------------------ masked-label cross-entropy
 * Replaces  masked_pretraining/model.py:104-105 (LinearHead) + :78-82 (MaskedCrossEntropyLoss):
 * gather the masked frames, logits = h @ W^T + b over them only, mean cross-entropy with an online
 * log-sum-exp; the [M, V] logits never reach HBM in the forward.
 *
 * pero_head_prepare: fp32 head [V, Dh] (+ bias [V]) -> opaque blob with ONE bf16 copy of W and the bias (the logits
 *   GEMMs read it K-major, d_h = dlogits W reads the same copy MN-major).
 * pero_masked_ce_fwd:
 *   flags    PERO_CE_H_BF16 (1): h is bf16 (else fp32); PERO_CE_LABELS_PACKED (2): `labels` holds the packed
 *            (distance, index) winners of pero_vq_assign instead of plain int64 labels (label = low 32 bits), so that
 *            the head of a step whose labels are that step's codeword indices can start right behind the distance
 *            GEMM, without waiting for pero_vq_unpack;
 *            PERO_CE_KEEP_LOGITS (4): training forward -- the sweep also leaves in the workspace, as bf16, the
 *            exponentials of the masked frames' logits relative to the maximum of their 32-label chunk, plus those
 *            maxima in fp32 (the softmax numerators up to one fp32 factor per row and chunk; the loss is still taken
 *            from the fp32 accumulators).  A backward on the same workspace (h = NULL) given the same flag turns them
 *            into the dlogits in place, in one streaming pass that also yields the d_b partial sums, instead of
 *            recomputing the logits GEMM; dlogits are rounded to bf16 once, exactly as on the recompute route.
 *   h        hidden states [N, Dh]
 *   rows     [M] int32 frame indices of the masked frames, ascending (mask == 1 order)
 *   labels   [N] int64 (only labels[rows[m]] are read, by the GEMM epilogues, not by the gather; a label outside
 *            [0, V) -- where the reference's F.cross_entropy raises a device assert -- makes the loss NaN)
 *   loss_sum [1] fp32: sum over masked frames of (lse - logit[label]);  lse [M] fp32 saved for backward.  Both NULL:
 *            the log-sum-exp partials stay in the workspace and pero_masked_ce_loss() produces loss_sum / lse later
 *            (a backward on the same workspace does not need them), which takes the finalize launch off the
 *            forward -> backward critical path.
 *   h = NULL: pero_masked_ce_gather already ran on this workspace for the same (h, rows).  The gather reads neither the
 *            labels nor the head, so a caller whose labels are produced late in the step (the quantizer's indices)
 *            can issue it ahead of them, off the critical path.
 *   Any hidden size works; up to Dh = 512 the masked rows stay resident in shared memory during the label sweep.
 * pero_masked_ce_bwd: gradients of  loss = grad_scale[0] * inv_count * loss_sum:
 *   d_h [N, Dh] (same dtype as h, zero on unmasked frames), d_W [V, Dh] fp32, d_b [V] fp32.
 *   Two-phase use: a first call with d_h = NULL produces d_W, d_b (ready to be all-reduced); a second call with
 *   d_W = d_b = NULL and the SAME workspace produces d_h from the dlogits the first call left there.
 *   h = NULL: the workspace is the one pero_masked_ce_fwd ran on for the same (h, rows, labels) and has not been
 *   written since; the operands gathered there are reused (PERO_CE_H_BF16 must still describe d_h's dtype) and the
 *   log-sum-exp is rebuilt from the forward's partials (`lse` may then be NULL).
 *   `rows` must be ascending (the order of mask == 1): the frame -> masked-row map is a binary search over it.
 */
size_t pero_head_bytes(int64_t V, int64_t Dh);
int pero_head_prepare(const float* W, const float* bias, int64_t V, int64_t Dh, void* head, size_t head_bytes,
                      pero_stream_t stream);
size_t pero_masked_ce_workspace_bytes(int64_t N, int64_t M, int64_t V, int64_t Dh);
int pero_masked_ce_gather(const void* h, int h_is_bf16, int64_t N, int64_t Dh, const int32_t* rows, int64_t M, int64_t V,
                          void* workspace, size_t workspace_bytes, pero_stream_t stream);
#define PERO_CE_H_BF16 1
#define PERO_CE_LABELS_PACKED 2
#define PERO_CE_KEEP_LOGITS 4
int pero_masked_ce_fwd(const void* h, int flags, int64_t N, int64_t Dh, const int32_t* rows, int64_t M,
                       const int64_t* labels, const void* head, int64_t V, float* loss_sum, float* lse,
                       void* workspace, size_t workspace_bytes, pero_stream_t stream);
int pero_masked_ce_loss(int64_t N, int64_t Dh, int64_t M, int64_t V, float* loss_sum, float* lse, void* workspace,
                        size_t workspace_bytes, pero_stream_t stream);
int pero_masked_ce_bwd(const void* h, int flags, int64_t N, int64_t Dh, const int32_t* rows, int64_t M,
                       const int64_t* labels, const void* head, int64_t V, const float* lse,
                       const float* grad_scale, float inv_count, void* d_h, float* d_W, float* d_b,
                       void* workspace, size_t workspace_bytes, pero_stream_t stream);
/* Evaluation of the head on the masked frames without materialising logits
 * (replaces  masked_pretraining/tester.py:70-93  Tester._update_errors + the loss of :57-64): one sweep of the logits
 * GEMM yields the loss terms AND, per masked frame, the number of labels whose logit is strictly larger than the
 * frame's own label's (its 0-based rank); errors[i] = #frames with rank >= topk_host[i] ("label not among the
 * top-k predictions"; exact logit ties are broken in favour of the label).
 *   topk_host  HOST array of num_topk (1..8) values k >= 1, e.g. {1, 3, 10}
 *   rank  [M] int32 or NULL;  errors [num_topk] int64 (device).  Workspace: pero_masked_ce_workspace_bytes(). */
int pero_masked_ce_eval(const void* h, int flags, int64_t N, int64_t Dh, const int32_t* rows, int64_t M,
                        const int64_t* labels, const void* head, int64_t V, const int32_t* topk_host, int num_topk,
                        float* loss_sum, float* lse, int32_t* rank, int64_t* errors, void* workspace,
                        size_t workspace_bytes, pero_stream_t stream);
/* The same backward restricted to the label columns [v_begin, v_end) (v_begin and v_end multiples of 256, or
 * v_end = V): rows [v_begin, v_end) of d_W and d_b (pointers to the FULL arrays) and the matching columns of the
 * dlogits kept in the workspace.  A data-parallel caller walks the label axis range by range and exchanges each
 * range of d_W while the next one is being computed; once all ranges are done, a call with d_W = d_b = NULL and
 * d_h given produces d_h.  d_h together with d_W/d_b is accepted only for the full range. */
int pero_masked_ce_bwd_range(const void* h, int flags, int64_t N, int64_t Dh, const int32_t* rows, int64_t M,
                             const int64_t* labels, const void* head, int64_t V, const float* lse,
                             const float* grad_scale, float inv_count, int64_t v_begin, int64_t v_end, void* d_h,
                             float* d_W, float* d_b, void* workspace, size_t workspace_bytes, pero_stream_t stream);
/* Logits-in variant for callers that already hold logits [N, V] (fp32 or bf16):
 * MaskedCrossEntropyLoss.forward(output, labels, mask), masked_pretraining/model.py:78-82.
 * workspace: >= 4*M bytes rounded up to 256.  d_logits has the dtype of logits; zero_init != 0 clears it
 * first (rows not listed get zero gradient). */
int pero_ce_logits_fwd(const void* logits, int is_bf16, int64_t N, int64_t V, const int32_t* rows, int64_t M,
                       const int64_t* labels, float* loss_sum, float* lse, void* workspace, size_t workspace_bytes,
                       pero_stream_t stream);
int pero_ce_logits_bwd(const void* logits, int is_bf16, int64_t N, int64_t V, const int32_t* rows, int64_t M,
                       const int64_t* labels, const float* lse, const float* grad_scale, float inv_count, int zero_init,
                       void* d_logits, pero_stream_t stream);
/* Ordered compaction of a {0,1} mask without a host sync: rows[0..count) = indices with mask != 0
 * (optionally also requiring labels[i] >= 0), count[0] = their number.  mask_dtype: 0 int64, 1 int32,
 * 2 uint8/bool. */
size_t pero_mask_compact_workspace_bytes(int64_t N);
int pero_mask_compact(const void* mask, int mask_dtype, int want_value, const int64_t* labels_or_null, int64_t N,
                      int32_t* rows, int32_t* count, void* workspace, size_t workspace_bytes,
                      pero_stream_t stream);

/* ------------------------------------------------------------------ the callers either side of the head
 * pero_mask_pixels: TransformerEncoder.mask (models/transformers.py:53-68) on the device, in place:
 *     x[n, c, h, t*pw : (t+1)*pw] = tile[c, h, :]   for every masked frame r = n * frames_per_line + t in `rows`
 *   x [n_lines, C, H, W] fp32, tile [C, H, patch_width] fp32 (the reference's fixed noise tile: np.random.seed(42);
 *   np.random.rand(1, C, patch_h, patch_w)), rows = the masked-frame list of pero_masked_ce_fwd.
 * pero_head_argmax_prepare: the head as a prepared "codebook" (pero_vq_codebook_bytes(V, Dh) bytes) such that
 *   pero_vq_assign / pero_vq_assign_bf16 on the hidden states returns  argmax_v (h.W_v + b_v)  for every frame, lowest
 *   index on ties (masked_pretraining/visualizer.py:32: torch.argmax(output['output'], dim=-1)) -- the [N, V] logits
 *   of the evaluation / visualisation forward are never materialised. */
int pero_mask_pixels(float* x, int64_t n_lines, int64_t C, int64_t H, int64_t W, const int32_t* rows, int64_t M,
                     int64_t frames_per_line, int64_t patch_width, const float* tile, pero_stream_t stream);
int pero_head_argmax_prepare(const float* W, const float* bias, int64_t V, int64_t Dh, void* codebook, size_t codebook_bytes,
                             pero_stream_t stream);

/* ------------------------------------------------------------------ peer-memory collectives (NVLink 5 / NVSwitch)
 * The reference is single-process; these are the two exchange steps of the sharded path:
 *   batch-sharded    SUM of the EMA buffer [K*D + K] between pero_vq_ema_accumulate and pero_vq_ema_apply
 *                    (models/autoencoders.py:225-237 evaluated on the concatenated batch) and of the head
 *                    gradients d_W | d_b | loss_sum (masked_pretraining/model.py:78-82 on the concatenated batch);
 *   codebook-sharded MIN of the packed (distance, index) winners (models/autoencoders.py:217 argmin over all K).
 * Every rank owns one peer buffer of the same size which every rank has mapped: peer_bufs is a DEVICE array of
 * `world` base pointers (entry r = rank r's buffer, entry `rank` = the local one).  The first
 * PERO_PEER_HEADER_BYTES of every buffer are flag words owned by the library and must be zero before the
 * first call; payload ranges live at offset_bytes >= PERO_PEER_HEADER_BYTES (16-byte aligned, the same offset
 * on every rank) and are reduced IN PLACE: on return (in stream order) every rank's range holds the same bits.
 * multicast_base: the NVSwitch multicast alias of the buffers (switch-side reduction, multimem.ld_reduce /
 * multimem.st) or NULL (peer loads summed in rank order + peer stores).  One kernel per call, n_blocks CTAs
 * (1..PERO_PEER_MAX_BLOCKS; 16-32 saturate NVLink and leave the other SMs to the GEMMs running beside it);
 * all ranks must issue the same sequence of calls with the same n_blocks.  n_elems % 4 == 0 (f32) / % 2 == 0 (i64).
 * Header words the CALLER may touch (all others belong to the library): the u32 at PERO_PEER_TIMEOUT_OFFSET is the
 * barrier timeout in milliseconds (0 = the default, 600 000 ms = NCCL's watchdog); the u32 at PERO_PEER_ERROR_OFFSET is
 * set non-zero by a kernel whose peer did not arrive in time — that kernel returns WITHOUT trapping (the CUDA context
 * stays usable) and leaves the payload unreduced, so a caller that cannot rule out stragglers reads the word back
 * (peer.PeerBuffer.check()).  Ranks that do rank-local long work between steps should barrier on the host first.
 * pero_peer_allreduce_emulate: the same protocol with all `world` buffers on ONE device and the ranks played
 * by blockIdx.y of one cooperative launch (op 0 = f32 sum, 1 = i64 min) — single-GPU test of the protocol only.
 */
#define PERO_PEER_HEADER_BYTES 16384
#define PERO_PEER_TIMEOUT_OFFSET 12288
#define PERO_PEER_ERROR_OFFSET 12292
#define PERO_PEER_MAX_WORLD 16
#define PERO_PEER_MAX_BLOCKS 64
int pero_peer_allreduce_sum_f32(void* const* peer_bufs, void* multicast_base, int rank, int world, int64_t offset_bytes,
                                int64_t n_elems, int n_blocks, pero_stream_t stream);
int pero_peer_allreduce_min_i64(void* const* peer_bufs, void* multicast_base, int rank, int world, int64_t offset_bytes,
                                int64_t n_elems, int n_blocks, pero_stream_t stream);
int pero_peer_allreduce_emulate(void* const* bufs_on_one_device, int world, int op, int64_t offset_bytes, int64_t n_elems,
                                int n_blocks, pero_stream_t stream);

/* ------------------------------------------------------------------ 1x1 projections around the quantizer
 * Replaces  models/autoencoders.py:114-115, 143-147  (VQVAE.encoder_projection_layer / decoder_projection_layer, the two
 * 1x1 Conv2d layers of VQVAE.quantize).  A 1x1 convolution is y[n, :] = W x[n, :] + b over the N = n_lines *
 * frames_per_line frames.
 * pero_proj_forward: x fp32, channels_first = 1: [n_lines, C, frames_per_line] (the NCHW tensor with H*W collapsed),
 *   0: rows [N, C]; weight [D, C] (the conv weight [D, C, 1, 1]), bias [D] or NULL.  fp32-grade arithmetic on the tensor
 *   cores: every operand is split into bf16 hi + lo parts, three partial products, fp32 accumulation (error ~2^-16
 *   relative).  Outputs, each optional (at least one): out_rows fp32 [N, D]; out_bf16 bf16 [N, Dp] (Dp = D rounded up
 *   to 64, zero padded): exactly the operand pero_vq_assign_bf16 reads, so the projected NCHW tensor never exists;
 *   packed_reset [N] int64 or NULL is set to "empty" (saves the pero_vq_packed_init launch in front of the assign).
 * pero_gather_rows_cf: out[l, c, t] = table[idx[l * frames_per_line + t], c]  (table [K, C] fp32, out channels-first):
 *   with table = pero_proj_forward(codebook rows, decoder weight, bias) this is decoder_projection_layer(quantized),
 *   because the 1x1 projection commutes with the row gather.  Indices are clamped to [0, K). */
size_t pero_proj_workspace_bytes(int64_t N, int64_t C, int64_t D);
int pero_proj_forward(const float* x, int64_t n_lines, int64_t frames_per_line, int channels_first, int64_t C,
                      const float* weight, const float* bias, int64_t D, float* out_rows, void* out_bf16,
                      int64_t* packed_reset, void* workspace, size_t workspace_bytes, pero_stream_t stream);
int pero_gather_rows_cf(const float* table, const int64_t* idx, int64_t n_lines, int64_t frames_per_line, int64_t K,
                        int64_t C, float* out, pero_stream_t stream);

/* ------------------------------------------------------------------ the GEMM core on its own (parity-test entry)
 * C[rows_a, rows_b] = A @ B^T through the same tcgen05 core every contraction of the path uses (bf16 operands with row
 * pitch `kd`, fp32 out [num_splits, rows_a, rows_b], one plane per contraction split).  variant: bit0 = CTA pairs
 * (cta_group::2), bit1 = resident A, bit5 = MN-major operands (A stored [kd, rows_a], B [kd, rows_b], C = A^T B).
 * tests/test_gpu_assign.py checks it against an fp64 matmul; nothing on the product path calls it. */
int pero_gemm_tn_bf16(const void* a_bf16, int64_t rows_a, const void* b_bf16, int64_t rows_b, int64_t kd,
                      int variant, int num_splits, float* out, pero_stream_t stream);

#ifdef PERO_DEV_BUILD
/* Dev build only (make DEV=1; not part of the production export table).  The next `slots` GEMM launches each write
 * clock64 stamps of their worker 0 into the next 64 KiB (8192 u64) slot of device_buffer, later launches none; NULL
 * switches it off.  Slot layout ([unit][8] u64: mma_top, mma_tempty_ok, mma_first_full, mma_issued, epi_tfull_ok,
 * epi_done, prod_first, prod_last).  The dev build also honours the PERO_* tuning knobs listed in csrc/knobs.h. */
int pero_debug_set_timeline(void* device_buffer, int slots);
#endif

#ifdef __cplusplus
}
#endif
#endif /* PERO_B200_H_ */
