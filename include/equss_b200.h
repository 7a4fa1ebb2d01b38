/*
 * equss_b200.h -- C-ABI of the B200-native EQUSS product-quantization hot path.
 *
 * The reference (pitlover/Expand-and-Quantize-for-Unsupervised-Semantic-Segmentation) is pure
 * Python/PyTorch and has no FFI layer; its boundary for this path is a set of nn.Module methods.
 * Each entry point below replaces the PyTorch op sequence cited next to it (paths relative to the
 * reference root).  The Python host side (package `equss_b200`) binds these with ctypes and keeps
 * the reference's module names / signatures; INTEGRATION.md shows the stub a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all tensors are dense fp32 unless stated; indices are int32 [M][N] (subspace-major);
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on that stream;
 *   - return value: 0 = ok, negative = error (see EQUSS_ERR_*); the message is available from
 *     equss_last_error_string().  No entry point throws, allocates device memory, retains pointers
 *     or falls back to the CPU: an unsupported shape is an explicit error.
 *
 * Activation layout (`equss_zdesc`): element (pixel n, channel c) of the expanded feature map with
 * n = b*HW + s lives at  z[b*stride_b + s*stride_s + c*stride_c]:
 *   flat  (n, D) row-major  (model/quantizer.py:396 "z_flat = z")  : stride_b=HW*D, stride_s=D, stride_c=1
 *   NCHW  (B, D, h, w)      (model/quantizer.py:112 permute path)  : stride_b=D*HW, stride_s=1, stride_c=HW
 */
#ifndef EQUSS_B200_H_
#define EQUSS_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EQUSS_OK                 0
#define EQUSS_ERR_INVALID_ARG   -1   /* null pointer, non-positive size, bad enum          */
#define EQUSS_ERR_UNSUPPORTED   -2   /* shape outside what the sm_100a kernels implement   */
#define EQUSS_ERR_CUDA          -3   /* a CUDA runtime / driver call failed                */
#define EQUSS_ERR_NO_DEVICE     -4   /* no sm_100 device visible: there is no CPU fallback */

/* z-side normalisation applied per (pixel, subspace) row before the distance
 * (model/quantizer.py:419-455; the codebook side is normalised by the caller, it is M*K*d elements). */
#define EQUSS_NORM_NONE    0   /* "none"                                                        */
#define EQUSS_NORM_L2      1   /* "l2":     z / max(||z||_2, 1e-12)         (F.normalize)        */
#define EQUSS_NORM_ZNORM   2   /* "z_norm": (z - mean) / (std_unbiased + 1e-5) over the d dims  */
#define EQUSS_NORM_AFFINE  3   /* "z_trainable": (z - mean[c]) / denom[c], vectors of length D   */

#define EQUSS_LAYOUT_FLAT  0
#define EQUSS_LAYOUT_NCHW  1

/* Which assign kernel to run (equss_pq_assign `algo`). AUTO picks the tcgen05 kernel whenever the
 * shape is supported and the exact SIMT kernel otherwise; both are CUDA, neither is a fallback to CPU. */
#define EQUSS_ASSIGN_AUTO     0
#define EQUSS_ASSIGN_SIMT     1   /* exact fp32 CUDA-core scan (validator + odd shapes)          */
#define EQUSS_ASSIGN_TCGEN05  2   /* tcgen05/TMEM GEMM + fused argmin + exact fp32 re-score: the fp16-split kernel
                                     for l2 rows with d in {16,32,64}, else the split-tf32 kernel            */
#define EQUSS_ASSIGN_TCGEN05_TF32 3 /* force the split-tf32 tcgen05 kernel (all norm modes, d in {8,16,32,64})  */

typedef struct equss_zdesc {
  int64_t n_pixels;   /* N = B*HW                                  */
  int64_t hw;         /* pixels per image (N for a flat (n,D) view) */
  int64_t stride_b;   /* element strides, see header comment        */
  int64_t stride_s;
  int64_t stride_c;
  int32_t dim;        /* D = M*d                                    */
  int32_t layout;     /* EQUSS_LAYOUT_* (a hint; strides are authoritative) */
} equss_zdesc;

/* ---------------------------------------------------------------------------------------------
 * library / device
 * ------------------------------------------------------------------------------------------- */
const char* equss_last_error_string(void);            /* thread-local, never NULL                */
int  equss_version(void);                              /* 10000*major + 100*minor + patch         */
int  equss_device_check(int device);                   /* 0 if `device` is sm_100; else NO_DEVICE  */
/* number of kernels this library launched since load (bench.py's `gpu_launches`) */
int64_t equss_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1  distance + argmin            replaces model/quantizer.py:457-467 (twins quantizer_v2.py:262-270,
 *                                  dino_pqgo.py:646-654, dino_new_vq.py:391-397) executed M times.
 *   idx[m][n] = argmin_k ( (sum z_norm^2 + cnorm2[m][k]) - 2 * <z_norm[n, m*d:(m+1)*d], codebook_norm[m][k]> )
 *   first minimal index wins (torch.argmin).  codebook_norm: [M][K][d].  cnorm2: [M][K] = sum(c^2).
 *   norm_a / norm_b: per-channel vectors [D] for EQUSS_NORM_AFFINE (mean, denominator), else NULL.
 *   margin_out (optional, may be NULL): [M][N] fp32 relative gap between best and second-best fp32
 *   distance, used by the near-tie audit.
 *   workspace: device scratch of equss_pq_assign_workspace_bytes() bytes (may be NULL if that is 0).
 * ------------------------------------------------------------------------------------------- */
int64_t equss_pq_assign_workspace_bytes(int64_t n_pixels, int M, int K, int d, int algo);
int equss_pq_assign(const float* z, const equss_zdesc* zd,
                    const float* codebook_norm, const float* cnorm2, int M, int K, int d,
                    int norm_mode, const float* norm_a, const float* norm_b,
                    int32_t* idx_out, float* margin_out,
                    void* workspace, int64_t workspace_bytes, int algo, void* stream);

/* cnorm2[m][k] = sum_j codebook_norm[m][k][j]^2   (model/quantizer.py:459) */
int equss_pq_cnorm2(const float* codebook_norm, int M, int K, int d, float* cnorm2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3  gather + losses + straight-through value     replaces model/quantizer.py:474,514,534-536
 *                                                  (param variant :153,175-184).
 *   q = gather_src[m][idx[m][n]]            gather_src: [M][K][d] (codebook_norm, or raw weight for
 *                                           dino_new_vq.py:403 / dino_pqgo.py:665)
 *   out(n, m*d+j)  = z_norm + (q - z_norm)  evaluated in fp32 exactly like the reference's STE line
 *   sqerr[m]      += sum_{n,j} (z_norm - q)^2     (double accumulators, caller zeroes them;
 *                                                  mse = sqerr[m] / (N*d))
 *   out uses the same strides as z.  znorm_out (optional) receives z_norm in the same layout.
 * ------------------------------------------------------------------------------------------- */
int equss_pq_gather_loss(const float* z, const equss_zdesc* zd,
                         const float* gather_src, const int32_t* idx, int M, int K, int d,
                         int norm_mode, const float* norm_a, const float* norm_b,
                         float* out, float* znorm_out, double* sqerr, void* stream);

/* K1 + K3 in one pass over the activations (the tcgen05 assign kernel gathers each tile from shared memory a few
 * tiles after assigning it): idx as equss_pq_assign, out / sqerr as equss_pq_gather_loss.  Available for l2 rows
 * with d in {16, 32} and K <= 256 (equss_pq_assign_gather_supported); other shapes: call K1 then K3.
 * Unlike equss_pq_gather_loss this entry point OVERWRITES sqerr (it is zeroed by the call's first kernel).
 * workspace: equss_pq_assign_workspace_bytes(..., EQUSS_ASSIGN_TCGEN05) bytes. */
int equss_pq_assign_gather_supported(const equss_zdesc* zd, int M, int K, int d, int norm_mode);
int equss_pq_assign_gather(const float* z, const equss_zdesc* zd,
                           const float* codebook_norm, const float* cnorm2, const float* gather_src,
                           int M, int K, int d, int norm_mode,
                           int32_t* idx_out, float* out, double* sqerr,
                           void* workspace, int64_t workspace_bytes, void* stream);

/* backward of K3 w.r.t. z for the straight-through output and the commitment/codebook MSE terms
 * (SURVEY 8b "Autograd"):  g_znorm = grad_out + coef[m] * (z_norm - q),   grad_z = J_norm(z)^T g_znorm
 *   coef[m] = 2*beta*grad_loss_m/(N*d) is supplied by the caller ([M] fp32 on device).
 *   Supported norm modes: NONE, L2, ZNORM, AFFINE.
 *   grad_codebook (optional, [M][K][d], caller-zeroed): += cb_coef[m] * (q - z_norm) scattered by idx
 *   (gradient of the codebook loss w.r.t. the gathered rows, model/quantizer.py:175). */
int equss_pq_gather_loss_bwd(const float* z, const equss_zdesc* zd,
                             const float* gather_src, const int32_t* idx, int M, int K, int d,
                             int norm_mode, const float* norm_a, const float* norm_b,
                             const float* grad_out, const float* coef,
                             float* grad_z,
                             const float* cb_coef, float* grad_codebook, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4  per-code counts and sums     replaces model/quantizer.py:485-488 (one_hot + one_hot^T @ z_flat)
 *   counts[m][k] += #{n : idx[m][n]==k}          (fp32, like the reference's float one-hot sum)
 *   sums[m][k][:] += sum_{n: idx==k} src(n)      src = raw z rows (use_norm=0, quantizer.py:488,
 *                                                dino_new_vq.py:411) or z_norm rows (use_norm=1,
 *                                                quantizer_v2.py:266,281)
 *   `packed` is ONE buffer [M][K][d+1]: column 0..d-1 = sums, column d = count, so that the
 *   data-parallel exchange (K5, quantizer.py:490-491) is a single all-reduce.  Caller zeroes it.
 * ------------------------------------------------------------------------------------------- */
int equss_pq_accumulate(const float* z, const equss_zdesc* zd, const int32_t* idx,
                        int M, int K, int d, int use_norm,
                        int norm_mode, const float* norm_a, const float* norm_b,
                        float* packed, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K6  EMA codebook update          replaces EmbeddingEMA.update, model/quantizer.py:233-254
 *   vq_count  <- decay*vq_count  + (1-decay)*count
 *   weight_avg<- decay*weight_avg+ (1-decay)*sum
 *   n = sum_k vq_count;  weight = weight_avg / ((vq_count+eps)/(n+K*eps)*n)
 *   packed: [M][K][d+1] as produced by K4 (after the all-reduce). All state tensors are [M][K](x[d]).
 *   exact_count (optional, [M][K]): += count  (the non-persistent `vq_count` of EMAVectorQuantizer,
 *   quantizer.py:493).  unused_out (optional, [M] int32): number of codes with count==0 (:509).
 * ------------------------------------------------------------------------------------------- */
int equss_ema_update(const float* packed, int M, int K, int d, double decay, double eps,
                     float* vq_count, float* weight_avg, float* weight,
                     float* exact_count, int32_t* unused_out, void* stream);

/* K6+K7 fused tail of the EMA training step   replaces model/quantizer.py:493-532 (the EMA update, both
 *   get_histogram_count calls, codebook-usage, codebook-sum, commitment / loss scalars, and the wrapper's mean over
 *   subspaces, :607-608) -- about fifteen small launches on the eager path -- by ONE launch.
 *   State update exactly as equss_ema_update (exact_count is required here).  sqerr: [M] fp64 squared errors from K3
 *   (or NULL: the caller computes the loss terms itself, e.g. through autograd).
 *   stats_out[10] = { total-p10, total-p50, total-p90, current-p10, current-p50, current-p90, codebook-usage,
 *                     codebook-sum, commitment-loss, loss = beta * commitment }, each the mean over the M subspaces.
 *   scratch: equss_pq_train_tail_scratch_floats(M) floats, zero before the first call (the kernel leaves it reusable).
 *   K <= 1024. */
int equss_pq_train_tail_scratch_floats(int M);
int equss_pq_train_tail(const float* packed, int M, int K, int d, double decay, double eps,
                        float* vq_count, float* weight_avg, float* weight, float* exact_count,
                        const double* sqerr, int64_t n_pixels, double beta,
                        float* scratch, float* stats_out, void* stream);

/* K5 + K6 + K7 in one kernel over peer memory: as equss_pq_train_tail, but the data-parallel SUM of the packed statistics
 * (model/quantizer.py:490-491, 2*M all_reduce_tensor calls in the reference) is taken inside the kernel from the ranks'
 * buffers over NVLink: peer_packed is a DEVICE array of `world` device pointers to every rank's packed [M][K][d+1]
 * buffer (symmetric memory, own rank included), summed in rank order so all replicas obtain identical bits.  The
 * reduced statistics are written to packed_out (local).  The caller orders the ranks: every rank's K4 must be complete
 * and visible (a device-side barrier over the symmetric-memory signal pads) before the launch, and a buffer may be
 * re-zeroed only after all peers have read it: the host mirror alternates two buffers and passes the OTHER one as
 * zero_next (optional, [M][K][d+1]), which this launch clears for the next step -- by the barrier above every peer is
 * past its reads of it. */
int equss_pq_train_tail_peers(const void* const* peer_packed, int world, float* packed_out, int M, int K, int d,
                              double decay, double eps, float* vq_count, float* weight_avg, float* weight,
                              float* exact_count, const double* sqerr, int64_t n_pixels, double beta,
                              float* scratch, float* stats_out, float* zero_next, void* stream);

/* Codebook-side normalisation (model/quantizer.py:421 "l2", :426 "z_norm", "none") and cnorm2 of the result in one
 * launch: codebook_norm [M][K][d], cnorm2 [M][K].  (The "z_trainable" flavours normalise across codes / with learned
 * statistics and stay with the caller.) */
int equss_pq_prepare_codebook(const float* codebook, int M, int K, int d, int norm_mode,
                              float* codebook_norm, float* cnorm2, void* stream);

/* K7  codebook-usage percentiles      replaces get_histogram_count, model/quantizer.py:15-30 (a Python loop with
 *   ~6K tensor->bool host syncs per subspace in the reference).  count: M rows of K entries, element (m, k) at
 *   count[m*row_stride + k*k_stride] (so the count column of the packed K4 buffer can be read in place).
 *   out: [M][3] = first rank whose descending cumulative usage reaches 10 / 50 / 90 %, divided by K; NaN where the
 *   reference returns None. */
int equss_usage_percentiles(const float* count, int64_t row_stride, int64_t k_stride, int M, int K,
                            float* out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  soft assignment              replaces model/quantizer.py:468 + :609 (softmax(-distance), concatenated
 *                                  over subspaces on the last dim; dino_pqgo.py:655 divides by jsd_ts)
 *   prob[n][m*K + k] = softmax_k( -distance[n][m][k] / temperature )     prob: [N][M*K] fp32
 * ------------------------------------------------------------------------------------------- */
int equss_pq_distance_prob(const float* z, const equss_zdesc* zd,
                           const float* codebook_norm, const float* cnorm2, int M, int K, int d,
                           int norm_mode, const float* norm_a, const float* norm_b,
                           float temperature, float* prob, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2b fused consumers of the soft assignment   replaces model/dino_new_vq.py:447-450, i.e.
 *     JSDLoss (model/loss.py:508-525) and EntropyLoss (model/loss.py:490-505) applied to
 *     torch.chunk(softmax(-distance / T), 2, dim=0), WITHOUT materialising the N x (K*M) probabilities.
 *   With p = rows [0, N/2) and q = rows [N/2, N) of the soft assignment of subspace m:
 *     kl_sum[m]      += sum_{n,k} (p+e) log((p+e)/mix) + (q+e) log((q+e)/mix),  mix = (p+q+e)/2, e = 1e-6
 *                       (jsd_m = 0.5 * kl_sum[m] / (N/2), KLDivLoss "batchmean")
 *     prob_sum[m][k] += sum_n p[n][k]          (entropy_m = sum_k a log(a + 1e-8), a = prob_sum / (N/2))
 *   Both outputs are fp64, caller-zeroed.  N must be even; K <= 256, d in {8,16,32,64}
 *   (equss_pq_soft_stats_supported); other shapes: materialise with equss_pq_distance_prob.
 * ------------------------------------------------------------------------------------------- */
int equss_pq_soft_stats_supported(int K, int d);
int equss_pq_soft_stats(const float* z, const equss_zdesc* zd,
                        const float* codebook_norm, const float* cnorm2, int M, int K, int d,
                        int norm_mode, const float* norm_a, const float* norm_b,
                        float temperature, double* kl_sum, double* prob_sum, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K13 per-channel moments of the activations   replaces model/quantizer.py:433-434 (torch.mean(z, 0) and
 *     torch.mean(z*z, 0), evaluated per subspace) for the running statistics of the "z_trainable" mode.
 *   sums[0][c] += sum_n z[n][c],  sums[1][c] += sum_n z[n][c]^2     sums: fp64 [2][D], caller-zeroed.
 *   One pass over z for all D = M*d channels; the caller divides by N and all-reduces the [2][D] pair once
 *   (the reference issues two all_reduce_tensor("mean") calls per subspace, :437-438).
 * ------------------------------------------------------------------------------------------- */
int equss_channel_moments(const float* z, const equss_zdesc* zd, double* sums, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K8  cluster / linear probe at label resolution   replaces model/evaluator.py:53-54,67-70,95-106
 *   step 1 (token resolution):  logits[b][s][j] = <feat[b,:,s], w[j,:]> + bias[j]
 *           feat: NCHW (B, D, h, w);  wmat_t: K-major weights [D][C_pad] (column j = probe channel j: columns
 *           0..C'-1 = L2-normalised cluster centres, further columns = linear-probe weights, padding
 *           columns zero); bias: [C_total] or NULL;
 *           logits: [B*h*w][C_pad] with C_pad = equss_probe_cpad(C_total).
 *   step 2 (label resolution):  bilinear (align_corners=False) interpolation of the token logits,
 *           argmax over each head's channel range, int64 predictions, fused confusion histogram.
 *           Because bilinear interpolation is linear and the per-pixel L2 norm is a positive scalar,
 *           argmax_j <normalize(interp(x)), c_j> == argmax_j interp(<x, c_j>)  (SURVEY 7.4).
 *   head h covers channels [head_off[h], head_off[h]+head_cnt[h]) of the logits.
 *   preds_out[h]: (B,H,W) int64 or NULL.  confusion[h]: [rows_h][C] int64 accumulated in place
 *   (rows = prediction, cols = label, model/metric.py:53-57) or NULL; rows_h = conf_rows[h].
 *   Masking as UnSegMetrics.update: 0<=label<C and 0<=pred<C (metric.py:49).
 * ------------------------------------------------------------------------------------------- */
int equss_probe_cpad(int c_total);
int equss_probe_logits(const float* feat, int B, int D, int h, int w,
                       const float* wmat_t, const float* bias, int c_total,
                       float* logits, void* stream);
/* Tensor-core variant of step 1 (tcgen05 split-tf32 GEMM, fp32-level accuracy): the weights are first packed into
 * an operand image (equss_probe_image_bytes() bytes, rebuilt whenever the weights change).  Supported when
 * D % 32 == 0, C_pad <= 64 and h*w % 32 == 0 (equss_probe_logits_tc_supported); equss_probe_logits covers the rest. */
int64_t equss_probe_image_bytes(int D, int c_total);
int equss_probe_build_image(const float* wmat_t, int D, int c_total, void* image, void* stream);
int equss_probe_logits_tc_supported(int D, int h, int w, int c_total);
int equss_probe_logits_tc(const float* feat, int B, int D, int h, int w,
                          const void* image, const float* bias, int c_total,
                          float* logits, void* stream);
/* Host-only helper (no GPU work): the persistent schedule equss_probe_argmax_confusion walks.  n_items work items
 * (image x row block x column slice) on `ctas` persistent CTAs (ctas <= n_items): out3 = { n_full, n_sched, parts } --
 * items [0, n_full) are processed whole, each of the remaining n_items - n_full items as `parts` row ranges of
 * rows_per_block / parts rows, n_sched = n_full + (n_items - n_full) * parts slots in total; CTA c handles slots
 * c, c + ctas, ...  Exposed so that the host logic is testable without a device (tests/test_probe_argmax_core.py). */
int equss_probe_argmax_schedule(int n_items, int ctas, int rows_per_block, int32_t* out3);
int equss_probe_argmax_confusion(const float* logits, int B, int h, int w, int c_total,
                                 const int64_t* label, int H, int W, int num_classes,
                                 int n_heads, const int32_t* head_off_host, const int32_t* head_cnt_host,
                                 int64_t* const* preds_out_host, int64_t* const* confusion_host,
                                 const int32_t* conf_rows_host, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K8b probe losses at label resolution   replaces model/evaluator.py:65-80 (masked cross-entropy of the
 *     linear probe) and :95-106 (ClusterLookup, alpha=None: minus the mean cosine between the upsampled feature and
 *     its nearest cluster centre), forward and backward, from the token-resolution logits of K8 step 1.
 *   equss_token_gram: gram[b][y][x][5] = <x,x>, <x,right>, <x,down>, <x,down-right>, <x,down-left> over the D channels
 *     of feat (B, D, h, w); the norm of the bilinearly interpolated feature vector follows from these 2x2 terms.
 *   equss_probe_losses:  sums[0] += sum over valid pixels (0 <= label < C) of  logsumexp(v_lin) - v_lin[label]
 *                        sums[1] += sum over ALL label pixels of  max_j v_clu[j] / max(|up(x)|, 1e-12)
 *                        n_valid += number of valid pixels       (linear_loss = sums[0]/n_valid, cluster_loss = -sums[1]/(B*H*W))
 *     grad_logits (optional, [B*h*w][C_pad], caller-zeroed): the transpose of the interpolation applied to the per-pixel
 *     gradients of the two SUMS above w.r.t. the interpolated logits (softmax - onehot on the linear channels,
 *     1/norm at the winning cluster channel); the caller scales by 1/n_valid, -1/(B*H*W) and the upstream gradients and
 *     contracts with the features (a [C_pad x N] x [N x D] library GEMM) to obtain the probe-parameter gradients.
 *   Heads as in equss_probe_argmax_confusion: cluster channels [off_cluster, +cnt_cluster), linear [off_linear, +cnt_linear).
 * ------------------------------------------------------------------------------------------- */
int equss_token_gram(const float* feat, int B, int D, int h, int w, float* gram, void* stream);
int equss_probe_losses_supported(int h, int w, int H, int W, int c_total, int cnt_cluster, int cnt_linear,
                                 int off_cluster, int off_linear);
int equss_probe_losses(const float* logits, const float* gram, int B, int h, int w, int c_total,
                       const int64_t* label, int H, int W, int num_classes,
                       int off_cluster, int cnt_cluster, int off_linear, int cnt_linear,
                       double* sums, uint64_t* n_valid, float* grad_logits, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K9  confusion histogram          replaces UnSegMetrics.update, model/metric.py:44-58
 *   confusion[pred][label] += 1 for every position with 0<=label<C and 0<=pred<C.
 *   confusion: [rows][C] int64 with rows = C + extra_classes.
 * ------------------------------------------------------------------------------------------- */
int equss_confusion_update(const int64_t* preds, const int64_t* label, int64_t n,
                           int num_classes, int rows, int64_t* confusion, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K11 global-feature kNN           replaces data/precompute_knns.py:313-315 (einsum + topk)
 *   For each query row, the k database rows with the largest inner product, sorted by decreasing
 *   similarity (ties: lower index first).  queries: [nq][F], db: [n][F] fp32 (caller normalises,
 *   precompute_knns.py:169).  idx_out: [nq][k] int64 (the `nns` array), sim_out: [nq][k] fp32 or NULL.
 *   1 <= k <= 32.  The result is an exact fp32 top-k on every path; for F % 64 == 0 the candidates are screened by an
 *   fp16 tensor-core GEMM inside a proven error margin and only the survivors are re-scored (no similarity matrix).
 * ------------------------------------------------------------------------------------------- */
int64_t equss_knn_workspace_bytes(int64_t nq, int64_t n, int F, int k);
int equss_knn_topk(const float* queries, int64_t nq, const float* db, int64_t n, int F, int k,
                   int64_t* idx_out, float* sim_out,
                   void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K12 expansion head  replaces the 1x1 convolutions of model/dino_pqgo.py:104-112,127-128 and
 *                                   model/blocks/module.py:20-44:  code = cluster1(x) + cluster2(x)
 *   One tcgen05 split-tf32 GEMM (fp32-level accuracy), called twice by the host mirror:
 *     out[r][o] = act( sum_{k<C1} A1[r][k] W[o][k] + sum_{k<C2} A2[r][k] W[o][C1+k] + bias[o] ),   r = b*hw + s
 *   a1: NCHW [B][C1][hw] (a1_nchw = 1, needs hw % 4 == 0; read in place, no permute) or flat [B*hw][C1];
 *   a2: flat [B*hw][C2] or NULL (C2 = 0);  w: [n_out][C1+C2] row-major (Conv2d weight (o, c, 1, 1), branches
 *   concatenated along c);  out: flat [B*hw][out_ld] -- the (pixel, channel) layout the PQ entry points take as
 *   EQUSS layout FLAT.  relu != 0 applies max(., 0) to the output (the hidden layer of cluster2).
 *   C1 % 16 == 0 and C2 % 16 == 0 (equss_head_gemm_supported).
 * ------------------------------------------------------------------------------------------- */
int equss_head_gemm_supported(int C1, int C2, int hw, int a1_nchw);
int equss_head_gemm(const float* a1, int a1_nchw, int C1, const float* a2, int C2, int B, int hw,
                    const float* w, const float* bias, int n_out, int relu, float* out, int64_t out_ld,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * K14 STEGO feature correlation   replaces the no-grad half of STEGOLoss.helper, model/loss.py:679-687:
 *     fd[n][p][q] = < f1[n,:,p] / max(|f1[n,:,p]|, 1e-10) , f2[n,:,q] / max(|f2[n,:,q]|, 1e-10) >
 *     pointwise != 0:  fd[n][p][:] -= mean_q fd[n][p][:]
 *   f1, f2: [n][C][P] sampled backbone features (P = feature_samples^2 <= 128 positions); fd: [n][P][P].
 *   sums (fp64 [2], caller-zeroed): += sum of fd before / after the row centring; the reference's final
 *   "fd - fd.mean() + old_mean" is the scalar (sums[0] - sums[1]) / (n*P*P) added to every element by the caller.
 * ------------------------------------------------------------------------------------------- */
int equss_stego_feature_corr(const float* f1, const float* f2, int n, int C, int P, int pointwise,
                             float* fd, double* sums, void* stream);

#ifdef __cplusplus
}
#endif
#endif  /* EQUSS_B200_H_ */
