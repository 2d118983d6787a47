/*
 * clc_b200.h -- C ABI of libclc_b200.so: the CLC conditional-latent hot path
 * (match + CLM fusion + ChARM entropy stage + bpp) as hand-written sm_100a kernels.
 *
 * The reference (ydchen0806/CLC) is pure Python; it has no FFI.  Each entry point
 * below replaces the PyTorch operator sequence of one reference function (cited
 * per function as file:line into the reference tree) and is what a ctypes/cffi
 * binding on the reference side would bind.
 *
 * Conventions:
 *   - plain pointers + sizes, no torch types; all tensors are fp32, NCHW, device memory
 *     on the calling thread's current CUDA device; index tensors are int32.
 *   - every call returns 0 on success or a negative clc_status; nothing throws.
 *   - no allocation, no host synchronisation, no global mutable state: the caller passes
 *     outputs and workspace; kernels are enqueued on `stream` (a cudaStream_t).
 *   - "bs" arguments are batch strides in ELEMENTS, so channel-slice views of a
 *     [B, 320, h, w] tensor (CLC_run.py:535 `y.chunk(5, 1)`) are accepted without copies.
 *   - accumulator outputs (`double* log2_sum`, parameter gradients, g_ref) are ADDED to;
 *     the caller zero-initialises them.
 */
#ifndef CLC_B200_H_
#define CLC_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define CLC_API __attribute__((visibility("default")))
#else
#define CLC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef enum clc_status {
  CLC_OK = 0,
  CLC_ERR_INVALID_ARGUMENT = -1, /* NULL where required, non-positive size, bad mode      */
  CLC_ERR_UNSUPPORTED = -2,      /* shape outside what the kernels implement               */
  CLC_ERR_WORKSPACE = -3,        /* workspace too small (see *_workspace_bytes)            */
  CLC_ERR_CUDA = -4,             /* a CUDA runtime / driver call failed (launch error)     */
  CLC_ERR_ARCH = -5              /* device is not sm_100 (tcgen05 / TMA path unavailable)  */
} clc_status;

CLC_API int clc_version(void);                 /* ABI version, currently 2 */
CLC_API const char* clc_strerror(int status);  /* static string */
/* Text of the last CUDA error seen by the calling thread (empty if none). */
CLC_API const char* clc_last_cuda_error(void);
/* Number of CUDA kernels this library has enqueued in the calling process (statistics). */
CLC_API uint64_t clc_kernel_launch_count(void);

/* Per-kernel tracing (the reference only has wall-clock timers, train_CLC.py:126-217).
 * Between clc_trace_start(stream) and clc_trace_stop() the library records one CUDA event on
 * `stream` after every kernel it enqueues (all calls must target that stream; not capturable
 * into a CUDA graph).  clc_trace_get(i) returns kernel i's label and the device time in ms
 * since the previous event, i.e. that kernel's duration when launches are back to back. */
CLC_API int clc_trace_start(void* stream);
/* Records an unlabeled event: the next kernel's time is measured from here (used after enqueueing
 * foreign work, e.g. a spin kernel that lets the host run ahead of the device). */
CLC_API int clc_trace_mark(void);
CLC_API int clc_trace_stop(void);
CLC_API int clc_trace_count(void);
CLC_API int clc_trace_get(int i, const char** name, float* ms);

/* ------------------------------------------------------------------------------------
 * Entropy stage
 * ---------------------------------------------------------------------------------- */

/* GaussianConditional.forward + ste_round, fused.
 * Replaces: compressai GaussianConditional.forward as called at CLC_run.py:569
 *           (restated in-tree at CLC_run.py:718-736), ste_round CLC_run.py:35-36,:571,
 *           and (optionally) the bpp reduction of train_CLC.py:48-51 for this tensor.
 * Tensors are [B, CS] with CS contiguous elements per batch item and the given batch strides.
 *   noise == NULL : eval,  outputs = round(y-mean)+mean          ("dequantize")
 *   noise != NULL : train, outputs = y + noise                   ("noise"; noise ~ U(-.5,.5))
 *   lik      = max( .5erfc(-(.5-v)/(s*sqrt2)) - .5erfc(-(-.5-v)/(s*sqrt2)), lik_bound ),
 *              v = |outputs - mean|, s = max(scale, scale_bound)
 *   y_hat    = round(y-mean)+mean        (optional, may be NULL)
 *   outputs  = first return value of GaussianConditional.forward (optional, may be NULL)
 *   log2_sum += sum(log2(lik))           (optional, may be NULL)
 * mean may be NULL (zero mean). */
CLC_API int clc_gc_fwd(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
               const float* mean, int64_t mean_bs, const float* noise, int64_t noise_bs,
               float* lik, int64_t lik_bs, float* y_hat, int64_t y_hat_bs,
               float* outputs, int64_t outputs_bs, double* log2_sum,
               int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream);

/* Backward of clc_gc_fwd (autograd semantics of the reference incl. both LowerBound gates:
 * gradient passes iff x >= bound or grad < 0; ste_round: d/dy = 1, d/dmean = 0).
 *   lik     : the likelihood tensor written by clc_gc_fwd (required)
 *   g_lik   : dL/dlik, or NULL -> dL/dlik = bpp_coef / lik  (bpp loss formed analytically,
 *             bpp_coef = dL/dbpp * 1/(-ln2 * num_pixels), train_CLC.py:48-51)
 *   g_y_hat : dL/dy_hat, may be NULL
 *   g_mean  : may be NULL when mean == NULL. */
CLC_API int clc_gc_bwd(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
               const float* mean, int64_t mean_bs, const float* noise, int64_t noise_bs,
               const float* lik, int64_t lik_bs, const float* g_lik, int64_t g_lik_bs,
               float bpp_coef, const float* g_y_hat, int64_t g_y_hat_bs,
               float* g_y, int64_t g_y_bs, float* g_scale, int64_t g_scale_bs,
               float* g_mean, int64_t g_mean_bs,
               int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream);

/* Training-mode quantisation noise generated INSIDE the kernels (compressai EntropyModel.quantize "noise",
 * `inputs + U(-1/2, 1/2)`, as reached from CLC_run.py:526 / :569) instead of read from a tensor: saves the
 * uniform_() launch and 8 B per element.  Philox4x32-10, counter = (element index / 4, state[1] + rng_offset),
 * key = state[0].  `rng_state` is a DEVICE uint64[2] {seed, base offset}: a captured CUDA graph draws fresh
 * noise on every replay once its owner advances the base (clc_rng_advance / clc_bpp_finalize); `rng_offset`
 * separates the calls of one step.  The backward regenerates the forward's sample from the same
 * (rng_state contents, rng_offset).  Same distribution as torch's uniform_(-.5, .5), not the same bits: pass an
 * explicit `noise` tensor to the plain entry points for bit-reproducible parity runs. */
CLC_API int clc_gc_fwd_rng(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                   const float* mean, int64_t mean_bs, const uint64_t* rng_state, uint64_t rng_offset,
                   float* lik, int64_t lik_bs, float* y_hat, int64_t y_hat_bs,
                   float* outputs, int64_t outputs_bs, double* log2_sum,
                   int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream);
CLC_API int clc_gc_bwd_rng(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                   const float* mean, int64_t mean_bs, const uint64_t* rng_state, uint64_t rng_offset,
                   const float* lik, int64_t lik_bs, const float* g_lik, int64_t g_lik_bs,
                   float bpp_coef, const float* g_y_hat, int64_t g_y_hat_bs,
                   float* g_y, int64_t g_y_bs, float* g_scale, int64_t g_scale_bs,
                   float* g_mean, int64_t g_mean_bs,
                   int64_t B, int64_t CS, float scale_bound, float lik_bound, void* stream);
/* state[1] += n (one-thread kernel; capturable). */
CLC_API int clc_rng_advance(uint64_t* rng_state, uint64_t n, void* stream);

/* Latent residual prediction add:  y_hat += 0.5 * tanh(lrp)   (CLC_run.py:582-583). */
CLC_API int clc_lrp_add_fwd(float* y_hat, int64_t y_hat_bs, const float* lrp, int64_t lrp_bs,
                    int64_t B, int64_t CS, void* stream);
/* g_lrp = g * 0.5 * (1 - tanh(lrp)^2);  (g_y_hat = g, no kernel needed). */
CLC_API int clc_lrp_add_bwd(const float* g, int64_t g_bs, const float* lrp, int64_t lrp_bs,
                    float* g_lrp, int64_t g_lrp_bs, int64_t B, int64_t CS, void* stream);

/* Quantised symbols + scale-table indexes for the entropy coder.
 * Replaces: GaussianConditional.quantize(y, "symbols", mean) CLC_run.py:690 and
 *           GaussianConditional.build_indexes(scale) CLC_run.py:689 (63 passes upstream).
 *   symbols = int32(round(y - mean));
 *   indexes = (T-1) - #{ t in table[0..T-2] : max(scale, scale_bound) <= t }
 * `scale_table` is a device pointer to T ascending floats (T <= 256). */
CLC_API int clc_gc_symbols_indexes(const float* y, int64_t y_bs, const float* scale, int64_t scale_bs,
                           const float* mean, int64_t mean_bs, const float* scale_table, int T,
                           int32_t* symbols, int64_t symbols_bs, int32_t* indexes, int64_t indexes_bs,
                           int64_t B, int64_t CS, float scale_bound, void* stream);

/* EntropyBottleneck.forward (factorised prior, filters (3,3,3,3)) + z STE round, fused.
 * Replaces: compressai EntropyBottleneck.forward called at CLC_run.py:526, `_get_medians`
 *           :528 and the z straight-through round :529-530.
 * z is [B, C, S] contiguous.  matrix[i]/bias[i]/factor[i] are device pointers to the module's
 * _matrix{i} [C,f(i+1),f(i)], _bias{i} [C,f(i+1),1], _factor{i} [C,f(i+1),1] with f=(1,3,3,3,3,1);
 * quantiles is [C,1,3] (median = quantiles[c][1]).
 *   noise == NULL : eval, outputs = round(z-med)+med;  else outputs = z + noise
 *   lik = max(|sigmoid(s*u) - sigmoid(s*l)|, lik_bound), l,u = logits(outputs -/+ .5),
 *         s = -sign(l+u)
 *   z_hat = round(z-med)+med (optional);  outputs optional;  log2_sum += sum(log2 lik). */
CLC_API int clc_eb_fwd(const float* z, const float* noise, const float* const matrix[5],
               const float* const bias[5], const float* const factor[4], const float* quantiles,
               float* lik, float* z_hat, float* outputs, double* log2_sum,
               int64_t B, int64_t C, int64_t S, float lik_bound, void* stream);

/* Backward of clc_eb_fwd.  g_lik may be NULL -> bpp_coef / lik.  g_z_hat may be NULL.
 * Parameter gradients are ACCUMULATED into g_matrix/g_bias/g_factor (same shapes as the
 * parameters; any of the three arrays may be NULL to skip parameter gradients). */
CLC_API int clc_eb_bwd(const float* z, const float* noise, const float* const matrix[5],
               const float* const bias[5], const float* const factor[4], const float* quantiles,
               const float* lik, const float* g_lik, float bpp_coef, const float* g_z_hat,
               float* g_z, float* const g_matrix[5], float* const g_bias[5], float* const g_factor[4],
               int64_t B, int64_t C, int64_t S, float lik_bound, void* stream);

/* clc_eb_fwd / clc_eb_bwd with in-kernel noise (see clc_gc_fwd_rng); element index = offset inside z. */
CLC_API int clc_eb_fwd_rng(const float* z, const uint64_t* rng_state, uint64_t rng_offset,
                   const float* const matrix[5], const float* const bias[5], const float* const factor[4],
                   const float* quantiles, float* lik, float* z_hat, float* outputs, double* log2_sum,
                   int64_t B, int64_t C, int64_t S, float lik_bound, void* stream);
CLC_API int clc_eb_bwd_rng(const float* z, const uint64_t* rng_state, uint64_t rng_offset,
                   const float* const matrix[5], const float* const bias[5], const float* const factor[4],
                   const float* quantiles, const float* lik, const float* g_lik, float bpp_coef,
                   const float* g_z_hat, float* g_z, float* const g_matrix[5], float* const g_bias[5],
                   float* const g_factor[4], int64_t B, int64_t C, int64_t S, float lik_bound, void* stream);

/* bpp of RateDistortionLoss from the accumulated log2 sums (train_CLC.py:48-51, eval.py:27-31):
 *   *bpp = -(log2_sums[0] + ... + log2_sums[n-1]) / num_pixels      (device doubles, one-thread kernel)
 * and, when rng_state != NULL, rng_state[1] += rng_advance (next step's noise). */
CLC_API int clc_bpp_finalize(const double* log2_sums, int32_t n, double num_pixels, double* bpp,
                             uint64_t* rng_state, uint64_t rng_advance, void* stream);
/* Zero-fill of an accumulator (cudaMemsetAsync on `stream`: a memset node, not a kernel). */
CLC_API int clc_zero(void* p, size_t bytes, void* stream);

/* Rate term of RateDistortionLoss for an arbitrary likelihood tensor.
 * Replaces: `torch.log(likelihoods).sum()` train_CLC.py:48-51, eval.py:27-31.
 *   fwd: *log2_sum += sum_i log2(lik[i]);
 *   bwd: g_lik[i] = coef * (coef_dev ? *coef_dev : 1) / lik[i]   (coef_dev: optional DEVICE
 *        float64 scalar, so an upstream autograd scalar never needs a host read). */
CLC_API int clc_log2_sum_fwd(const float* lik, int64_t n, double* log2_sum, void* stream);
CLC_API int clc_log2_sum_bwd(const float* lik, float coef, const double* coef_dev, float* g_lik,
                             int64_t n, void* stream);

/* ------------------------------------------------------------------------------------
 * Reference matching (Patch_Matching.py)
 * ---------------------------------------------------------------------------------- */

/* Addressing of query patches: element (patch, c, dy, dx) of problem n lives at
 *   q + (n / q_repeat) * q_sn + (patch / npx) * q_spy + (patch % npx) * q_spx
 *     + c * q_sc + dy * q_sy + dx
 * which covers both an extracted [P, C, ph, pw] patch tensor (npx = P, q_spx = C*ph*pw,
 * q_sc = ph*pw, q_sy = pw) and patches addressed in place inside an NCHW image
 * (q_spy = ph*W, q_spx = pw, q_sc = H*W, q_sy = W) -- the reshape/permute copy of
 * Patch_Matching.py:172 is never made. */
typedef struct clc_patch_view {
  const float* q;
  int64_t q_sn, q_spy, q_spx, q_sc, q_sy;
  int32_t npx;      /* patches per row of the patch grid                                   */
  int32_t q_repeat; /* consecutive problems sharing one query image (= n_refs when the refs
                       of an image are stacked along the problem axis), >= 1              */
} clc_patch_view;

/* Pearson correlation map, materialised (exact fp32 FMA path; "fp32 mode").
 * Replaces: L2_or_pearson_corr Patch_Matching.py:854-910 (+ `cross_corr * mask` :182-183).
 *   r    : [NP, C, fh, fw] reference-side features, one per problem
 *   mask : NULL or [P, fh-ph+1, fw-pw+1] (shared by all problems)
 *   corr : out [NP, P, fh-ph+1, fw-pw+1] */
CLC_API int clc_pearson_corr(const clc_patch_view* qv, const float* r, const float* mask, float* corr,
                     int64_t NP, int32_t P, int32_t C, int32_t ph, int32_t pw,
                     int32_t fh, int32_t fw, void* workspace, size_t workspace_bytes, void* stream);
CLC_API size_t clc_pearson_corr_workspace_bytes(int64_t NP, int32_t P, int32_t C, int32_t ph, int32_t pw,
                                        int32_t fh, int32_t fw);

/* Row-wise top-k (values descending, ties -> lowest index), warp-shuffle selection.
 * Replaces: torch.topk(cross_corr, k, dim=2) Patch_Matching.py:224.  k <= 32.
 *   x [R, L] -> val [R, k], idx [R, k] (int32 position in the row). */
CLC_API int clc_topk_rows(const float* x, int64_t R, int64_t L, int32_t k, float* val, int32_t* idx,
                  void* stream);

/* create_gaussian_masks Patch_Matching.py:779-807, evaluated in fp64 on the device and
 * rounded to fp32 like the reference's numpy path.  mask: out [P, img_h-ph+1, img_w-pw+1]. */
CLC_API int clc_gaussian_mask(float* mask, int32_t img_h, int32_t img_w, int32_t ph, int32_t pw, void* stream);

/* Fused match: correlation GEMM on tcgen05 tensor cores (bf16 operands staged by TMA,
 * fp32 accumulate in TMEM) with the Pearson normalisation, Gaussian mask and a per-patch
 * candidate top-KC selection fused into the epilogue (the P x L map is never written),
 * followed by exact fp32 FMA re-scoring of the candidates and the final warp-shuffle top-k.
 * Replaces: L2_or_pearson_corr + mask + torch.topk (Patch_Matching.py:181-183,:224).
 *   q_img : [NQ, C, H, W] query latents, problem n uses image n / q_repeat
 *   r     : [NP, C, H, W] reference latents (same spatial size as the query)
 *   gaussian_mask : 0 = no mask, 1 = create_gaussian_masks(H, W, ph, pw)
 *   val, idx : out [NP, P, k], P = (H/ph)*(W/pw); idx = oy*(W-pw+1)+ox
 *   n_uncertified : optional device int32, SET to the number of patches whose candidate set could not
 *                   be certified to contain the exact top-k (margin below the bf16 screening error
 *                   bound; such a patch may differ from the fp32 path at a near-tie), may be NULL */
CLC_API int clc_match_topk_tc(const float* q_img, const float* r, int64_t NP, int32_t q_repeat,
                      int32_t C, int32_t H, int32_t W, int32_t ph, int32_t pw, int32_t k,
                      int32_t gaussian_mask, float* val, int32_t* idx, int32_t* n_uncertified,
                      float temperature, float* aligned, float* weights,
                      void* workspace, size_t workspace_bytes, void* stream);
/*   aligned : optional out [NP, C, H, W] (16-byte aligned).  When non-NULL the gather + blend of
 *             SI_Wraper (Patch_Matching.py:225-238, is_stack = False, same-scale features: the
 *             gathered feature map is `r` itself) is fused into the re-scoring kernel:
 *             aligned = sum_j softmax_j(val * temperature) * window_j(r), reassembled in place of
 *             the query patches -- identical to clc_gather_blend_fwd(r, idx, val, ...).
 *   weights : optional out [NP, P, k], the softmax weights (needed by clc_match_bwd). */
CLC_API size_t clc_match_topk_tc_workspace_bytes(int64_t NP, int32_t q_repeat, int32_t C, int32_t H,
                                         int32_t W, int32_t ph, int32_t pw, int32_t k);
/* Channels-last fp32 copy of the reference latents [NP, H*W, C] that clc_match_topk_tc leaves in
 * its workspace (valid until the workspace is reused); pass it to clc_match_bwd as `r_cl` to skip
 * that call's own transpose.  Returns NULL for an unsupported shape. */
CLC_API const float* clc_match_topk_tc_ref_cl(void* workspace, int64_t NP, int32_t q_repeat, int32_t C,
                                              int32_t H, int32_t W, int32_t ph, int32_t pw, int32_t k);

/* clc_match_topk_tc with the SimpleCLM elementwise fusion (models/CLM.py:170-182; clc_clm_fuse_fwd) folded into its
 * last kernel: problem n = image * R + reference (q_repeat = R); the R references of one (image, patch) run as a
 * thread-block cluster and exchange their blended tiles through distributed shared memory.
 *   aligned : out [NP, C, H, W] (required);  att : plane (r, b) of [H*W] logits at att + r*att_sr + b*att_sb
 *   fused   : out [NP/R, C, H, W] = sum_r aligned_r * softmax_r(att) * sigmoid(att_r) + q_img
 * Identical results to clc_match_topk_tc followed by clc_clm_fuse_fwd.  R <= 8. */
CLC_API int clc_match_clm_fwd(const float* q_img, const float* r, int64_t NP, int32_t R, int32_t C, int32_t H,
                              int32_t W, int32_t ph, int32_t pw, int32_t k, int32_t gaussian_mask, float* val,
                              int32_t* idx, int32_t* n_uncertified, float temperature, float* aligned,
                              float* weights, const float* att, int64_t att_sr, int64_t att_sb, float* fused,
                              void* workspace, size_t workspace_bytes, void* stream);

/* softmax(value*T) weights + gather of the k matched patches + weighted sum + tile
 * reassembly (or channel stacking).
 * Replaces: SI_Wraper Patch_Matching.py:226-238.
 *   feat : [NP, C, fh, fw] features gathered from; fh = npy*gh, fw = npx*gw where (gh, gw) is
 *          the patch size at this feature scale.  idx indexes a correlation map of width
 *          corr_w (oy = idx / corr_w, ox = idx % corr_w).  For the multi-scale reuse
 *          (:198-208) the caller passes indices of the sub-sampled map, its width, and the
 *          scaled patch size.
 *   out  : [NP, C, fh, fw] (is_stack = 0) or [NP, k*C, fh, fw] (is_stack = 1)
 *   weights : out [NP, P, k] softmax weights (needed by the backward), may be NULL if is_stack */
CLC_API int clc_gather_blend_fwd(const float* feat, const int32_t* idx, const float* val, float temperature,
                         float* out, float* weights, int64_t NP, int32_t C, int32_t fh, int32_t fw,
                         int32_t gh, int32_t gw, int32_t corr_w, int32_t k, int32_t is_stack,
                         void* stream);
/* Backward: g_feat (ACCUMULATED, zero-init by caller) and g_val [NP, P, k]. */
CLC_API int clc_gather_blend_bwd(const float* feat, const int32_t* idx, const float* weights,
                         float temperature, const float* g_out, float* g_feat, float* g_val,
                         int64_t NP, int32_t C, int32_t fh, int32_t fw, int32_t gh, int32_t gw,
                         int32_t corr_w, int32_t k, int32_t is_stack, void* stream);

/* Backward of the (masked) Pearson correlation at the k selected positions only, with the
 * reference's autograd semantics: the conv2d weights are detached query patches
 * (Patch_Matching.py:869) so d(xy)/dq is dropped, while the query still receives gradient
 * through x_sum / sum_x_square / x_mean (:880-887) and the reference side through all terms.
 *   g_val : [NP, P, k] dL/d(masked corr value);  g_r ACCUMULATED [NP, C, fh, fw];
 *   g_q   : ACCUMULATED, addressed like the query through `gqv` (same strides as qv), may be NULL */
CLC_API int clc_pearson_topk_bwd(const clc_patch_view* qv, const float* r, const float* mask,
                         const int32_t* idx, const float* g_val, float* g_r, float* g_q,
                         int64_t NP, int32_t P, int32_t C, int32_t ph, int32_t pw,
                         int32_t fh, int32_t fw, int32_t k, void* stream);

/* Fused backward of clc_gather_blend_fwd (is_stack = 0) + the masked Pearson top-k values for the
 * single-scale wiring in which the gathered feature map IS the matched reference (feat == r and
 * the gather patch equals the match patch): equivalent to clc_gather_blend_bwd followed by
 * clc_pearson_topk_bwd, with one pass of reductions and ONE scatter-add per window element.
 *   g_out : [NP, C, fh, fw] dL/d(blended reference);  weights : [NP, P, k] from the forward
 *   g_r ACCUMULATED [NP, C, fh, fw];  g_q ACCUMULATED through `qv` (may be NULL);
 *   g_val : optional out [NP, P, k] = dL/d(masked corr value) (may be NULL)
 *   r_cl  : optional channels-last fp32 copy of r, [NP, fh*fw, C] (e.g. clc_match_topk_tc_ref_cl);
 *           NULL -> the call transposes r itself (one more kernel)
 *   flags : CLC_MATCH_BWD_OVERWRITE_G_R -> g_r is WRITTEN (no zero-init needed, no read-modify-write)
 *   workspace : clc_match_bwd_workspace_bytes(...) bytes enable the channels-last fast path
 *               (coalesced float4 loads, vector atomics); NULL selects the workspace-free kernel. */
#define CLC_MATCH_BWD_OVERWRITE_G_R 1
#define CLC_MATCH_BWD_WS_ZEROED 2 /* the caller already ran clc_match_bwd_zero_workspace on `workspace` */
CLC_API int clc_match_bwd(const clc_patch_view* qv, const float* r, const float* r_cl, const float* mask,
                          const int32_t* idx, const float* weights, float temperature, const float* g_out,
                          float* g_r, float* g_q, float* g_val, int64_t NP, int32_t P, int32_t C, int32_t ph,
                          int32_t pw, int32_t fh, int32_t fw, int32_t k, int32_t flags, void* workspace,
                          size_t workspace_bytes, void* stream);
CLC_API size_t clc_match_bwd_workspace_bytes(int64_t NP, int32_t C, int32_t fh, int32_t fw);
/* Zeroes the gradient scratch inside `workspace` (one memset).  Lets the caller take the memset off
 * the critical path (e.g. on a parallel stream / CUDA-graph branch during the forward pass) and then
 * pass CLC_MATCH_BWD_WS_ZEROED. */
CLC_API int clc_match_bwd_zero_workspace(void* workspace, size_t workspace_bytes, int64_t NP, int32_t C,
                                         int32_t fh, int32_t fw, void* stream);

/* Backward of [match -> SimpleCLM elementwise fusion] in one call: clc_clm_fuse_bwd folded into clc_match_bwd
 * (models/CLM.py:170-182 + Patch_Matching.py:218-240, :854-910, reference autograd semantics as above).
 * The gradient of the aligned references, g_aligned_r = g_fused * softmax_r(att) * sigmoid(att_r), is formed on
 * the fly and never written; the R CTAs of one (image, patch) publish their G_r = sum_c g_fused_c * aligned_r,c to
 * the workspace and the transposing kernel that ends the call forms g_att from them (fixed order r = 0..R-1).
 *   qv      : query patches addressed in place, q_repeat == R (problem n = image * R + reference)
 *   r_cl    : channels-last fp32 copy of the references [NP, fh*fw, C] (clc_match_topk_tc_ref_cl), required
 *   g_fused : [NP/R, C, fh, fw] dL/d(fused feature);  att / g_att : plane (r, b) at + r*att_sr + b*att_sb
 *   aligned : [NP, C, fh, fw] blended references of the forward pass
 *   g_r, g_q, g_val, flags, workspace : as clc_match_bwd (workspace required)
 * Shapes outside the fused kernel (patches other than 4x4, k > 4, C/4*ph not in {128,...,384}, R > 8) return
 * CLC_ERR_UNSUPPORTED: call clc_clm_fuse_bwd + clc_match_bwd instead. */
CLC_API int clc_match_clm_bwd(const clc_patch_view* qv, const float* r_cl, const float* mask, const int32_t* idx,
                              const float* weights, float temperature, const float* g_fused, const float* att,
                              int64_t att_sr, int64_t att_sb, const float* aligned, float* g_r, float* g_q,
                              float* g_val, float* g_att, int64_t NP, int32_t R, int32_t P, int32_t C, int32_t ph,
                              int32_t pw, int32_t fh, int32_t fw, int32_t k, int32_t flags, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------
 * CLM conditional fusion (elementwise part; the 1x1 / 3x3 convolutions stay nn.Conv2d)
 * ---------------------------------------------------------------------------------- */

/* Replaces: SimpleCLM.forward CLM.py:170-182 (sigmoid gate, softmax over refs, weighted sum, +y).
 *   ref_t : R x B planes of [C, S]; plane (r, b) starts at ref_t + r*ref_sr + b*ref_sb
 *   att   : R x B planes of [S];    plane (r, b) starts at att   + r*att_sr + b*att_sb
 *   (strides in elements: [R,B,C,S] stacking has ref_sr = B*C*S, ref_sb = C*S; the [B,R,C,S]
 *    layout produced by the match stage has ref_sr = C*S, ref_sb = R*C*S -- no transpose copy)
 *   y [B, C, S] -> out [B, C, S]
 *   out = sum_r softmax_r(att) * ref_t[r] * sigmoid(att[r]) + y */
CLC_API int clc_clm_fuse_fwd(const float* ref_t, int64_t ref_sr, int64_t ref_sb, const float* att,
                             int64_t att_sr, int64_t att_sb, const float* y, float* out,
                             int32_t R, int64_t B, int32_t C, int64_t S, void* stream);
/* g_ref_t / g_att are written with the same strides as ref_t / att; g_y = g_out (no kernel). */
CLC_API int clc_clm_fuse_bwd(const float* ref_t, int64_t ref_sr, int64_t ref_sb, const float* att,
                             int64_t att_sr, int64_t att_sb, const float* g_out, float* g_ref_t,
                             float* g_att, int32_t R, int64_t B, int32_t C, int64_t S, void* stream);

/* ---- CLM variant (a): similarity-softmax alignment (models/CLM.py:5-128), forward ----
 * Batch index nb = r*B + b over the R references of B images ([R, B, ...] stacking); NB = R*B. */

/* Replaces: the HW x HW `torch.bmm(y_t^T, y_ref_t) / T` + softmax (CLM.py:104-107) as consumed by
 * DeformableAlignment.forward's accumulation loop (:16-20), which only ever uses the COLUMN SUMS of the map:
 *   colsum[nb, p] = sum_q softmax_p(y_t[nb % B_y][:, q] . ref_t[nb][:, :] / T)[p]
 *   y_t [B_y, C, HW], ref_t [NB, C, HW] -> colsum [NB, HW].  Two passes over register tiles; the map is never
 *   written.  workspace: clc_clm_sim_colsum_workspace_bytes(NB, HW) (row max / row sum). */
CLC_API size_t clc_clm_sim_colsum_workspace_bytes(int64_t NB, int64_t HW);
CLC_API int clc_clm_sim_colsum(const float* y_t, const float* ref_t, int64_t NB, int64_t B_y, int32_t C, int64_t HW,
                               float temperature, float* colsum, void* workspace, size_t workspace_bytes,
                               void* stream);
/* Replaces: weighted_x (CLM.py:16-20) + torch.cat([x, weighted_x], 1) (:22):
 *   out[nb, 0:C] = x[nb],  out[nb, C:2C] = x[nb] * colsum[nb]      x [NB, C, HW] -> out [NB, 2C, HW] */
CLC_API int clc_clm_weighted_concat(const float* x, const float* colsum, float* out, int64_t NB, int32_t C,
                                    int64_t HW, void* stream);
/* Replaces: DeformableAlignment.deform_conv (CLM.py:35-60), the Python loops over B*H*W*9 taps.
 *   x [NB, C, H, W]; offset [NB, 18, H, W] (channel 2k = dh, 2k+1 = dw of tap k, the view of :29);
 *   modulation [NB, 9, H, W] -- the convolution output when modulation_is_logit (sigmoid fused, :25), else
 *   already sigmoided;  out [NB, C, H, W]. */
CLC_API int clc_clm_deform_fwd(const float* x, const float* offset, const float* modulation,
                               int32_t modulation_is_logit, float* out, int64_t NB, int32_t C, int32_t H, int32_t W,
                               void* stream);
/* Replaces: CLM.py:117-126  out = sum_r softmax_r(att)[r] * aligned[r] + y.
 *   aligned [R, B, C, S], att [R, B, S], y [B, C, S] -> out [B, C, S].  R <= 8. */
CLC_API int clc_clm_attention_sum_fwd(const float* aligned, const float* att, const float* y, float* out, int32_t R,
                                      int64_t B, int32_t C, int64_t S, void* stream);

/* ---- Swin window attention core (WMSA inside the SWAtten parameter networks, CLC_run.py:107-193,
 * :399-409, and inside g_a / g_s / h_a / h_s), forward ----
 * Replaces everything between WMSA's two Linear layers: roll, window partition, q/k/v slicing, q k^T * scale +
 * relative-position bias, the shifted-window mask, softmax, . v, window reverse, roll back.
 *   qkv [B, H, W, 3C] = embedding_layer(x) on the UN-rolled, UN-partitioned tokens (channel (t*heads + h)*head_dim + d,
 *   t = q, k, v);  rel_table [heads, 2*window-1, 2*window-1] (the module's relative_position_params);
 *   out [B, H, W, C] (channel h*head_dim + d), ready for the output Linear.  window == 8, head_dim in {8, 16, 32}. */
CLC_API int clc_window_attention_fwd(const float* qkv, const float* rel_table, float* out, int64_t B, int32_t H,
                                     int32_t W, int32_t C, int32_t head_dim, int32_t window, int32_t shifted,
                                     float scale, void* stream);

/* ---- reference retrieval, search half (dataloader_ref_cluster.py:64, :162) ----
 * Replaces: sklearn NearestNeighbors(algorithm='ball_tree').kneighbors, called per sample on the host.
 *   queries [Q, D], dict [N, D] (fp32, D % 4 == 0, D <= 6400) -> neg_d2 [Q, N] = -(squared Euclidean distance);
 *   follow with clc_topk_rows(neg_d2, Q, N, k, val, idx): idx = the k nearest entries (nearest first, ties ->
 *   lowest index), distance = sqrt(-val). */
CLC_API int clc_knn_neg_sqdist(const float* queries, const float* dict, int32_t Q, int32_t N, int32_t D,
                               float* neg_d2, void* stream);

/* ---- range coder: compress() / decompress() (CLC_run.py:629-716, :738-814) ----
 * HOST functions (no stream argument, plain host pointers): one rANS stream is a sequential
 * recurrence, so the state machine runs on the host; its inputs (symbols, scale-table indexes) come
 * from clc_gc_symbols_indexes on the device.  Wire format of compressai.ans (rans64: 64-bit state,
 * 32-bit words, 16-bit probabilities, 4-bit bypass groups). */

/* Replaces: compressai._CXX.pmf_to_quantized_cdf used by EntropyModel._pmf_to_cdf [upstream], reached
 * from CLC.update() CLC_run.py:486-491.  pmf [n] -> cdf_out [n + 1], cdf_out[n] = 2^precision, every
 * symbol keeps a non-zero frequency. */
CLC_API int clc_pmf_to_quantized_cdf(const float* pmf, int32_t n, int32_t precision, int32_t* cdf_out);

/* Replaces: BufferedRansEncoder.encode_with_indexes + flush (CLC_run.py:654, :712-713) and
 * RansEncoder.encode_with_indexes (EntropyBottleneck.compress, CLC_run.py:643).
 *   symbols, indexes [n]; cdfs [n_cdfs, cdf_stride]; cdf_sizes, offsets [n_cdfs]  (all host int32)
 *   out : 4-byte aligned host buffer, out_capacity >= clc_rans_encode_capacity(n) bytes;
 *   the stream is out[0 .. *out_bytes). */
CLC_API int clc_rans_encode(const int32_t* symbols, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                            int32_t n_cdfs, int32_t cdf_stride, const int32_t* cdf_sizes,
                            const int32_t* offsets, uint8_t* out, size_t out_capacity, size_t* out_bytes);
CLC_API size_t clc_rans_encode_capacity(int64_t n);

/* Replaces: RansDecoder.set_stream / decode_stream (CLC_run.py:758-760, :793) and decode_with_indexes
 * (EntropyBottleneck.decompress, CLC_run.py:749).  state[2] = {rANS state, next word}; {0, 0} starts a
 * stream, later calls continue it (one call per slice).  out [n] host int32 symbols. */
CLC_API int clc_rans_decode(const uint8_t* stream, size_t stream_bytes, uint64_t* state, const int32_t* indexes,
                            int64_t n, const int32_t* cdfs, int32_t n_cdfs, int32_t cdf_stride,
                            const int32_t* cdf_sizes, const int32_t* offsets, int32_t* out);

/* ------------------------------------------------------------------------------------
 * Data-parallel exchange over NVLink peer memory (the reference trains with threaded
 * nn.DataParallel, train_CLC.py:74-79,:472-473, whose gradient is the mean over the global batch)
 * ---------------------------------------------------------------------------------- */

/* Peer-visible device memory: a zero-filled cudaMalloc region exported / opened through CUDA IPC (HOST
 * functions; the 64-byte handles travel over the caller's own transport, e.g. torch.distributed). */
CLC_API int clc_peer_alloc(size_t bytes, void** ptr);
CLC_API int clc_peer_free(void* ptr);
CLC_API int clc_peer_export(void* ptr, uint8_t handle[64]);
CLC_API int clc_peer_open(const uint8_t handle[64], void** ptr);
CLC_API int clc_peer_close(void* ptr);
/* Bytes of the peer region clc_peer_allreduce needs on every rank (0 = unsupported arguments). */
CLC_API size_t clc_peer_allreduce_bytes(int32_t n_stat, int32_t n_grads, int32_t world);

/* One-shot all-reduce of the path's per-step exchange in ONE kernel per rank (no ring): every rank publishes
 * {stat[n_stat] doubles, grads[n_grads] floats} in its peer region, signals all ranks through NVLink, waits
 * for all of them and sums the world's contributions in fixed rank order (identical bits on every rank):
 *   stat  <- sum over ranks            (the bpp statistic: sum log2 likelihoods)
 *   grads <- grad_scale * sum over ranks  (1/world: the mean gradient of the EntropyBottleneck parameters)
 *   regions : HOST array of `world` device pointers, regions[r] = rank r's region (own or clc_peer_open'ed)
 *   state   : device uint64[3], zero-initialised once; the kernel keeps its step counter there, so the launch can
 *             be captured into a CUDA graph.  Every rank must enqueue the same sequence of calls.  world <= 8. */
CLC_API int clc_peer_allreduce(void* const* regions, int32_t rank, int32_t world, double* stat, int32_t n_stat,
                               float* grads, int32_t n_grads, float grad_scale, uint64_t* state, void* stream);

#ifdef CLC_DEBUG_ABI
/* ------------------------------------------------------------------------------------
 * Bring-up / test hooks of the tcgen05 match kernel (no reference counterpart).  They exist ONLY in
 * libclc_b200_dbg.so, the -DCLC_DEBUG_ABI build of the same sources that tests/test_match_tc_gpu.py and
 * scripts/ load; the production library exports none of them and has no mutable global state.
 * ---------------------------------------------------------------------------------- */

/* Timing aid: bit i of `mask` enables the i-th kernel of the multi-kernel entry points
 * (clc_match_topk_tc: prepass, gemm, rescore; clc_match_bwd: main, transpose).  Default 0xff = all; bits 8-15 are kernel-specific experiment switches (0 in production).
 * Results are only meaningful with every stage on. */
CLC_API void clc_debug_set_stage_mask(int mask);
/* wall-clock (ns, %globaltimer) phase stamps of the first 64 CTAs of the last re-scoring kernel: [64][16] */
CLC_API int clc_debug_rescore_stamps(long long* host_out);
CLC_API int clc_debug_bwd_stamps(long long* host_out);      /* same for the match backward kernel */

/* clc_match_topk_tc that additionally dumps the raw bf16-GEMM accumulators
 * xy[NP, P, H*W] (linear window origins oy*W+ox, wrapped ones included). */
CLC_API int clc_debug_match_tc_xy(const float* q_img, const float* r, int64_t NP, int32_t q_repeat, int32_t C,
                                  int32_t H, int32_t W, int32_t ph, int32_t pw, int32_t k,
                                  int32_t gaussian_mask, float* val, int32_t* idx, float* xy, void* workspace,
                                  size_t workspace_bytes, void* stream);
/* clc_match_topk_tc that additionally records per-CTA clock64 stamps of the GEMM kernel's
 * pipeline stages into timing[148][16] (int64). */
CLC_API int clc_debug_match_tc_timing(const float* q_img, const float* r, int64_t NP, int32_t q_repeat,
                                      int32_t C, int32_t H, int32_t W, int32_t ph, int32_t pw, int32_t k,
                                      int32_t gaussian_mask, float* val, int32_t* idx, long long* timing,
                                      void* workspace, size_t workspace_bytes, void* stream);
#endif /* CLC_DEBUG_ABI */

#ifdef __cplusplus
}
#endif
#endif /* CLC_B200_H_ */
