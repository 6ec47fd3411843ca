/* aline_b200 -- C ABI of the B200-native ALINE rollout + sPCE hot path.
 *
 * The reference (huangdaolang/ALINE) is pure Python/PyTorch and has no FFI of
 * its own; the boundary this library sits behind is the Python call surface
 * listed in SURVEY.md section 8(b).  Each entry point below names the reference
 * function (file:line under the reference tree) whose arithmetic it replaces.
 * The Python mirror in aline_b200/ binds these with ctypes.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - all tensors are dense, row-major, fp32 unless stated; indices are int64;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and
 *     no entry point synchronises the device;
 *   - the library never allocates or frees caller memory: scratch space is
 *     passed in, its size obtained from the matching *_scratch_bytes() query;
 *   - return value 0 = success; otherwise aline_last_error() describes the
 *     failure (thread-local string).
 */
#ifndef ALINE_B200_H
#define ALINE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALINE_ABI_VERSION 4

int aline_abi_version(void);
const char* aline_last_error(void);
/* Number of kernels this library has launched in the calling process (for bench accounting). */
uint64_t aline_kernel_launches(void);

/* ---------------------------------------------------------------- sPCE ---- */

enum aline_task {
    ALINE_TASK_LOCATION = 0,     /* tasks/location_finding.py:110-130,149-164 */
    ALINE_TASK_CES = 1,          /* tasks/ces.py:169-210 + distributions/censored_sigmoid_normal.py:47-86 */
    ALINE_TASK_PSYCHOMETRIC = 2  /* tasks/psychometric.py:107-134,178-195 */
};

/* Simulator likelihood description (host struct, passed by pointer).
 *   location:     dim_x = D, K sources, dim_theta = K*D; c0 = noise_scale, c1 = base_signal, c2 = max_signal
 *   ces:          dim_x = 6, dim_theta = 5 [rho, a1, a2, a3, log u]; c0 = noise_scale, c1 = epsilon
 *   psychometric: dim_x = 1, dim_theta = 4 [alpha, beta, gamma, lambda]
 */
typedef struct aline_lik {
    int32_t task;
    int32_t dim_x;
    int32_t K;
    int32_t dim_theta;
    float c0, c1, c2, c3;
} aline_lik;

/* Prior of the contrastive draws for device-side generation (Philox4x32-10, key = seed, counter =
 * (global row, column, call)).  ALINE_PRIOR_BOX: theta_i ~ U(lo_i, hi_i) (tasks/location_finding.py:85-98 uniform prior,
 * tasks/psychometric.py:70-89); ALINE_PRIOR_CES: rho = 0.01 + 0.99 U, alpha ~ Dirichlet(1,1,1), log u ~ N(lo[4], hi[4])
 * (tasks/ces.py:52-81). */
enum { ALINE_PRIOR_BOX = 0, ALINE_PRIOR_CES = 1 };
typedef struct aline_prior {
    int32_t kind;
    int32_t dim_theta;
    float lo[16];
    float hi[16];
} aline_prior;

/* thetas [n_rows, B, dim_theta] <- prior draws of the global rows row_offset .. row_offset + n_rows - 1 (replaces
 * Task.sample_theta((n_rows, B)) on the evaluation path; statistical, not value, parity with torch's generator). */
int aline_prior_sample(const aline_prior* prior, uint64_t seed, int64_t row_offset, int64_t n_rows, int32_t B,
                       float* thetas, void* stream);

/* Task.sample_batch on the device (replaces tasks/location_finding.py:167-192, tasks/ces.py:213-234,
 * tasks/psychometric.py:197-222 on the evaluation path, so that eval_boed's M-loop never touches the host generator):
 * for the rollouts g = batch_offset .. batch_offset + B - 1 draw theta_0 [B, dim_theta] from `prior`, n_points designs
 * x [B, n_points, dim_x] ~ U(x_lo, x_hi) per coordinate (the normalised designs the model sees) and their outcomes
 * y [B, n_points, 1] simulated at design_scale * x.  Philox streams keyed by (seed, g, point), disjoint from the
 * contrastive draws of aline_prior_sample; statistical parity with torch's generator. */
int aline_sample_batch(const aline_lik* lik, const aline_prior* prior, uint64_t seed, int64_t batch_offset, int32_t B,
                       int32_t n_points, float x_lo, float x_hi, float design_scale, float* theta, float* x, float* y,
                       void* stream);

/* aline_spce_history with the contrastive rows 1 .. n_rows-1 generated INSIDE the fused pass from the same streams as
 * aline_prior_sample (global row = row_offset + local row), so the draws never touch HBM: thetas holds only row 0
 * (theta_0, [1,B,dim_theta]).  Location K=1, D=2 with a box prior; seq [n_rows,B] is scratch.  *redo_flag (device) is set
 * non-zero when a shifted sum under/overflowed: the caller then materialises the draws with aline_prior_sample and calls
 * aline_spce_history (the robust path) -- same values, since both evaluate the same function of (seed, row, column). */
/* Rows of the `seq` scratch array aline_spce_history_device_prior will touch for this problem: n_rows when the history
 * needs several passes (the accumulated log-likelihood is carried between them), 1 when the whole history is evaluated
 * in one pass (T <= 36: only theta_0's row) -- then neither the draws nor their running sums ever exist in HBM. */
int64_t aline_spce_device_prior_seq_rows(const aline_lik* lik, int64_t n_rows, int32_t T);
int aline_spce_history_device_prior(const aline_lik* lik, const aline_prior* prior, uint64_t seed, int64_t row_offset,
                                    const float* y, const float* xi, const float* theta0, float* seq, int64_t n_rows,
                                    int32_t B, int32_t T, float* out_m, float* out_s, float* out_lp0, int32_t* redo_flag,
                                    void* scratch, size_t scratch_bytes, void* stream);

/* Scratch bytes needed by aline_spce_step / aline_spce_history / aline_log_likelihood for B trajectories
 * and T history points (T = 1 for the step and element-wise entry points).  Pure host arithmetic. */
size_t aline_spce_scratch_bytes(int32_t B, int32_t T);

/* History points aline_spce_history covers per pass over thetas for this likelihood and B (a history of
 * T points takes ceil(T / pass_len) passes; seq may be NULL only when that is 1).  Pure host arithmetic. */
int32_t aline_spce_pass_len(const aline_lik* lik, int32_t B);

/* One EIGStepLoss.step + the two logsumexp reductions of EIGStepLoss.forward
 * (loss/eig.py:174-209) for one history point, over this caller's shard of rows:
 *     seq[l,b] += log p(y[b] | xi[b], thetas[l,b,:])        for all n_rows rows
 *     out_m[b], out_s[b] = running max and sum exp(seq - max) over rows >= skip_rows
 * A single-GPU caller passes the reference's [L+1,B,..] tensors with skip_rows = 1
 * (row 0 is theta_0, excluded from the contrastive sum; its value is seq[0,b]).
 *   y [B], xi [B,dim_x] (unnormalised design), thetas [n_rows,B,dim_theta], seq [n_rows,B] in/out,
 *   out_m, out_s, out_lp0 [B] (out_lp0 = seq[0,b] after the update; required when skip_rows = 1).
 *   bad_flag: device int32, OR-ed with 1 when a NaN/inf log-likelihood is produced
 *             (the reference raises ArithmeticError, censored_sigmoid_normal.py:83-84).
 */
int aline_spce_step(const aline_lik* lik, const float* y, const float* xi, const float* thetas,
                    float* seq, int64_t n_rows, int32_t B, int32_t skip_rows,
                    float* out_m, float* out_s, float* out_lp0, int32_t* bad_flag,
                    void* scratch, size_t scratch_bytes, void* stream);

/* Fused whole-history evaluation (compute_EIG_from_history, utils/eval.py:43-80, and
 * EIGBounds.compute_seq_logprobs, loss/eig.py:22-48): T history points in one call,
 * thetas streamed once per chunk of history points, step-wise partials for every t.
 *   y [B,T], xi [B,T,dim_x], thetas [n_rows,B,dim_theta],
 *   seq [n_rows,B] in/out accumulator (zero it for a fresh evaluation); may be NULL (= zeros, nothing
 *       written back) when T <= aline_spce_pass_len(lik, B), i.e. when the whole history fits one pass,
 *   out_m, out_s [B,T]: partials after history point t over rows >= skip_rows,
 *   out_lp0 [B,T]: accumulated log-likelihood of row 0 after history point t (written only if skip_rows > 0).
 */
int aline_spce_history(const aline_lik* lik, const float* y, const float* xi, const float* thetas,
                       float* seq, int64_t n_rows, int32_t B, int32_t T, int32_t skip_rows,
                       float* out_m, float* out_s, float* out_lp0, int32_t* bad_flag,
                       void* scratch, size_t scratch_bytes, void* stream);

/* aline_spce_history with flags.  ALINE_SPCE_SEQ_SCRATCH: seq holds zeros on entry and its final content is not
 * needed by the caller; with theta_0 as row 0 (skip_rows = 1) this enables the shifted-accumulation fast pass
 * (location likelihoods): the contrastive sums are accumulated relative to theta_0's own log-likelihood, one
 * exp2 per evaluation, and recomputed by the robust kernels only if a sum under- or overflowed. */
#define ALINE_SPCE_SEQ_SCRATCH 1
/* ALINE_SPCE_LAST_ONLY: the caller only reads the entries of the LAST history point (out_*[b][T-1]) -- the
 * reference's compute_EIG_from_history(stepwise=False), utils/eval.py:72-74, which keeps the final step's losses.
 * The fast pass then takes one exponential per contrastive row instead of one per (row, history point); the other
 * entries of out_m / out_s are unspecified (finite).  A hint: paths without a last-only variant ignore it. */
#define ALINE_SPCE_LAST_ONLY 2
int aline_spce_history_ex(const aline_lik* lik, const float* y, const float* xi, const float* thetas,
                          float* seq, int64_t n_rows, int32_t B, int32_t T, int32_t skip_rows,
                          float* out_m, float* out_s, float* out_lp0, int32_t* bad_flag,
                          void* scratch, size_t scratch_bytes, int32_t flags, void* stream);

/* Combine R shards' partials into the EIGStepLoss.forward outputs
 *     pce_loss = logsumexp_{l=0..L} seq - seq[0],   nmc_loss = logsumexp_{l=1..L} seq - seq[0]
 * (loss/eig.py:200-202).  m, s [R,n]; lp0 [n]; outputs [n].
 */
int aline_lse_combine(const float* m, const float* s, const float* lp0, int32_t R, int64_t n,
                      float* pce_loss, float* nmc_loss, void* stream);

/* Element-wise simulator log-likelihood, thetas [n_rows,B,dim_theta] -> out [n_rows,B]
 * (Task.log_likelihood with y [1,B,1], xi [1,B,dim_x]). */
int aline_log_likelihood(const aline_lik* lik, const float* y, const float* xi, const float* thetas,
                         float* out, int64_t n_rows, int32_t B, int32_t* bad_flag,
                         void* scratch, size_t scratch_bytes, void* stream);

/* ------------------------------------------------------- model forward ---- */

/* ALINE model description (host struct).  `params` is the DEVICE pointer of the packed fp32 parameter blob:
 * every nn.Linear weight transposed to [in][out], in the order
 *   x_embedder {W1 [dx][EH], b1, W2 [EH][d], b2}, y_embedder {same with dy}, theta_tokens [n_theta_tok][d],
 *   per encoder layer {Wq, Wk, Wv [d][d], bq, bk, bv, Wo [d][d], bo, norm1 w, b, linear1 [d][ff], b,
 *                      linear2 [ff][d], b, norm2 w, b},
 *   acquisition head {W1 [d + time_token][HH], b1, w2 [HH], b2 (padded to 4)},
 *   per GMM component {W1 [d][HH], b1, W2 [3][HH], b2 (padded to 4)}
 * (segments whose length is not a multiple of 4 floats are zero-padded to one; aline_model_param_count gives the
 * total).  State-dict source: model/embedder.py:47-65, model/encoder.py:76-79, model/head.py:27-33,214-224. */
typedef struct aline_model {
    int32_t d;            /* dim_embedding: 32 or 64 */
    int32_t ff;           /* encoder dim_feedforward */
    int32_t n_head;       /* d / n_head must be 8 */
    int32_t n_layer;
    int32_t dim_x, dim_y;
    int32_t n_theta_tok;  /* learnable theta tokens (theta / mix embedding), 0 in data mode */
    int32_t n_comp;       /* GMM components */
    int32_t emb_hidden;   /* hidden width of the x / y embedders */
    int32_t head_hidden;  /* hidden width of the acquisition and GMM heads */
    int32_t time_token;   /* acquisition head takes [z ; t] */
    float std_min;
    const float* params;
    uint64_t n_params;
} aline_model;

uint64_t aline_model_param_count(const aline_model* m);

/* Process-wide kernel-selection switches (A/B measurements, tests).  "query_tc4": which fast tensor-core candidate
 * stream runs when both are possible -- -1 automatic (two threads per row above 32 keys), 0 never, 1 always.
 * "query_fold": the one-thread-per-row stream with the query projection folded into the key operand and the output
 * projection into the value operand (13 instead of 19 MMA phases per tile; csrc/query_tc3.cu; the context kernels
 * emit the folded operands) -- -1 / 2 up to 32 keys (default), 1 up to 48 keys, 0 never.
 * "ces_fast_pow": power arithmetic of the CES likelihood -- 1 (default) exp2 / log2 form with hi + lo products,
 * 0 eight powf per evaluation (2.7x slower; csrc/lik.cuh). */
int aline_set_option(const char* name, int32_t value);

/* Embedder on the candidate queries (model/embedder.py:143-147): query_x [B,nq,dx] -> eq [B,d,nq] (k-major). */
int aline_embed_queries(const aline_model* m, const float* query_x, int32_t B, int32_t nq, float* eq, void* stream);

/* The same embedding with a second, optional output layout: eq_rm [B,nq,d] (row-major: one candidate = d consecutive
 * floats), the input of the two-threads-per-row tensor-core query stream (csrc/query_tc4.cu).  eq or eq_rm may be
 * NULL (not both). */
int aline_embed_queries_ex(const aline_model* m, const float* query_x, int32_t B, int32_t nq, float* eq, float* eq_rm,
                           void* stream);

/* Context + target tokens of every rollout through embedder and all encoder layers (model/embedder.py:128-214,
 * model/encoder.py:128-141 restricted to the rows that are keys).  cx [B,ctx_cap,dx], cy [B,ctx_cap] hold n_c valid
 * context points; target_x [B,n_td,dx] the data-target inputs (mix / data embedding), followed by the model's
 * theta tokens.  tgt_slot [n_td + n_theta_tok] int32: position of target i among the targets the queries attend
 * to, or -1 (batch.target_mask, model/encoder.py:108-124).  Writes per layer the key / value rows
 * kv [n_layer,B,kv_slots,2,d] (slots 0..n_c-1 context, then the selected targets) and, if not NULL,
 * the target encodings z_tgt [B, n_td + n_theta_tok, d].
 * tckv (optional, d = 32, n_keys_tc = n_c + number of selected targets <= 160): the same keys / values as the bf16
 * operand blocks of the fast tensor-core query stream, aline_tc_kv_bytes(m, B, n_keys_tc) bytes, fully rewritten
 * by every call.  With nkp = n_keys_tc rounded up to 16, block (layer l, rollout b) sits at byte (l*B + b)*208*nkp:
 *   K part, 80*nkp bytes: 5 chunks of nkp rows x 16 B; chunk h < 4, row j = bf16 (K[j] - K[0])[8h .. 8h+8)
 *           (scores are taken relative to key 0; rows j >= n_keys_tc are 0), chunk 4 row j = [mask_j, 0 x 7] with
 *           mask_j = 0 for a real key, -200 for a padding row;
 *   V part, 128*nkp bytes: per head h, nkp/8 chunks of 16 rows x 16 B (row = 8 consecutive keys): rows 0..7 =
 *           V[key][8h + row], row 8 = 1 for real keys (yields the softmax denominator), rows 9..15 = 0. */
int aline_ctx_stack(const aline_model* m, const float* cx, const float* cy, int32_t B, int32_t n_c, int32_t ctx_cap,
                    const float* target_x, int32_t n_td, const int32_t* tgt_slot, float* kv, int32_t kv_slots,
                    float* z_tgt, void* tckv, int32_t n_keys_tc, void* stream);

/* aline_ctx_stack with one more optional output: z_ctx [B,n_c,d], the final-layer encodings of the context tokens,
 * the input of ValueHead (model/head.py:368-370). */
int aline_ctx_stack_ex(const aline_model* m, const float* cx, const float* cy, int32_t B, int32_t n_c, int32_t ctx_cap,
                    const float* target_x, int32_t n_td, const int32_t* tgt_slot, float* kv, int32_t kv_slots,
                    float* z_tgt, float* z_ctx, void* tckv, int32_t n_keys_tc, void* stream);

/* ValueHead.forward (model/head.py:84-111): value[b] = mean_t( w2 . relu(W1 z_ctx[b,t] + b1) + b2 ); W1 [ff,d], b1 [ff],
 * w2 [ff], b2 [1] are the module's own tensors (head.value_head.predictor.{0,2}.{weight,bias}). */
int aline_value_head(const float* z_ctx, int32_t B, int32_t n_c, int32_t d, int32_t ff, const float* w1, const float* b1,
                     const float* w2, const float* b2, float* value, void* stream);
uint64_t aline_tc_kv_bytes(const aline_model* m, int32_t B, int32_t n_keys);

/* Every live candidate through all encoder layers + the acquisition MLP (model/head.py:27-31, pre-softmax):
 * logits [B,nq] (-inf for retired candidates; alive [B,nq] uint8 or NULL = all live), optionally the query
 * encodings zq [B,nq,d].  n_keys = n_c + number of selected targets. */
int aline_query_stream(const aline_model* m, const float* eq, const uint8_t* alive, int32_t B, int32_t nq,
                       const float* kv, int32_t n_keys, int32_t kv_slots, float t_value, float* logits, float* zq,
                       void* stream);

/* Tensor-core (tcgen05, bf16 operands / fp32 accumulate) variant of aline_query_stream for d = 32.
 * tc_weights: device blob of aline_tc_weight_bytes(m) bytes, two sections, every matrix bf16 in the core-matrix
 * tiled layout (8-column chunks; chunk c of an R-row matrix at byte c*R*16, 16 bytes per row):
 *   (1) general kernel: per layer Wq [d][d], Wo [d][d], linear1 [ff][d], linear2 [d][ff]; acquisition W1[:, :d];
 *   (2) fast kernel: the same matrices with their bias folded in as 16 extra input columns
 *       [b_hi, b_lo, wt, wt, 0 x 12] (bf16 hi / lo split of the bias; wt = time-token column of the acquisition
 *       W1, else 0) that multiply the operand columns [1, 1, t_hi, t_lo, 0 ...]: per layer c*Wq [d][d+16]
 *       (c = log2(e)/sqrt(8)), Wo [d][d+16], linear1 [ff][d+16], linear2 [d][16+ff] (bias columns FIRST);
 *       acquisition W1 [HH][d+16].
 * n_keys <= aline_tc_max_keys(m).  tckv: operand blocks written by aline_ctx_stack for these n_keys; when given
 * and n_keys <= aline_tc_fast_max_keys(m) the fast kernel runs (all contractions incl. Q K^T and P V on the
 * tensor cores, softmax relative to key 0); should a softmax row overflow (a score > 127 log2-units above key
 * 0's) the general kernel recomputes the launch.  Otherwise the general kernel runs (FFMA attention from kv). */
uint64_t aline_tc_weight_bytes(const aline_model* m);
int32_t aline_tc_max_keys(const aline_model* m);
int32_t aline_tc_fast_max_keys(const aline_model* m);
int aline_query_stream_tc(const aline_model* m, const void* tc_weights, const float* eq, const uint8_t* alive, int32_t B,
                          int32_t nq, const float* kv, int32_t n_keys, int32_t kv_slots, float t_value, float* logits,
                          float* zq, const void* tckv, void* stream);

/* aline_query_stream_tc with the row-major embeddings eq_rm [B,nq,d] (aline_embed_queries_ex) as an additional,
 * optional input: when given together with tckv and the shape has one (d = 32, ff = head_hidden = 128, n_keys <= 48),
 * the fast kernel is the two-threads-per-row variant (6-8 warps per scheduler instead of 4).  eq may then be NULL
 * unless the general kernel can be reached (n_keys beyond the fast kernel's limit). */
int aline_query_stream_tc_ex(const aline_model* m, const void* tc_weights, const float* eq, const float* eq_rm,
                             const uint8_t* alive, int32_t B, int32_t nq, const float* kv, int32_t n_keys,
                             int32_t kv_slots, float t_value, float* logits, float* zq, const void* tckv, void* stream);

/* Softmax over the live candidates, first-argmax, log-prob (model/head.py:355-358) and, if cx != NULL, the
 * in-place Task.update_batch (tasks/base_task.py:133-154): append (qx, qy)[idx] at context position n_c, retire
 * the candidate.  idx_out[b*idx_stride] = index within the compacted live set (the reference's design_out.idx),
 * idx_orig_out[b] = index into the original candidate array, zt [B,nq] = probabilities (alive must be NULL). */
int aline_select(const float* logits, uint8_t* alive, int32_t B, int32_t nq, const float* qx, const float* qy,
                 int32_t dx, int32_t dy, float* cx, float* cy, int32_t n_c, int32_t ctx_cap, int64_t* idx_out,
                 int32_t idx_stride, float* logp_out, int32_t logp_stride, int64_t* idx_orig_out, float* zt,
                 void* stream);

/* aline_select in TRAIN mode (model/head.py:350-354): idx ~ Categorical(zt) instead of the argmax, drawn by inverse CDF
 * over the live candidates from one Philox4x32-10 uniform keyed by (seed, rollout b, step) -- fused with the softmax,
 * the log-prob (Categorical.log_prob = log(clamp(p, eps, 1 - eps))) and the optional in-place append.  Statistical,
 * not value, parity with torch.multinomial. */
int aline_select_sample(const float* logits, uint8_t* alive, int32_t B, int32_t nq, const float* qx, const float* qy,
                        int32_t dx, int32_t dy, float* cx, float* cy, int32_t n_c, int32_t ctx_cap, int64_t* idx_out,
                        int32_t idx_stride, float* logp_out, int32_t logp_stride, int64_t* idx_orig_out, float* zt,
                        uint64_t seed, int32_t step, void* stream);

/* GMMTargetHead.forward (model/head.py:152-186, 252-266): z [n_tok,d] -> means, stds, weights [n_tok,n_comp]. */
int aline_gmm_head(const aline_model* m, const float* z, int64_t n_tok, float* means, float* stds, float* weights,
                   void* stream);

/* calculate_gmm_variance (utils/misc.py:244-279), the uncertainty-sampling baseline's acquisition score:
 * var = sum_c w_c (sigma_c^2 + (mu_c - sum_c w_c mu_c)^2).  means/stds [n,C]; weights [n / tok_per_w, C] (tok_per_w = 1:
 * per-token weights; = n_query: one weight row per rollout, the reference's 2-D weights case) -> out [n]. */
int aline_gmm_variance(const float* means, const float* stds, const float* weights, int64_t n, int32_t C,
                       int64_t tok_per_w, float* out, void* stream);

/* GMMTargetHead.forward fused with calculate_gmm_variance: z [n_tok,d] -> variance [n_tok]; posterior_out_query
 * (model/head.py:366) is never materialised (notebooks/eval_al.ipynb cell 1, acquisition "uncertainty_sampling"). */
int aline_gmm_head_variance(const aline_model* m, const float* z, int64_t n_tok, float* variance, void* stream);

/* compute_ll (utils/eval.py:200-207): value [n], means/stds/weights [n,C] -> out [n]. */
int aline_gmm_log_likelihood(const float* value, const float* means, const float* stds, const float* weights,
                             int64_t n, int32_t C, float* out, void* stream);

/* Task.update_batch out of place (tasks/base_task.py:103-154) for one (query, context) pair of tensors:
 * query [B,N,D], ctx [B,M,D], idx [B] -> new_query [B,N-1,D], new_ctx [B,M+1,D]. */
int aline_move_selected(const float* query, const float* ctx, const int64_t* idx, int32_t B, int32_t N, int32_t M,
                        int32_t D, float* new_query, float* new_ctx, void* stream);

/* get_traces' T-step loop (utils/eval.py:21-30), resident: T x (ctx_stack, query_stream, select+append) enqueued
 * back to back on `stream`, no host synchronisation.  cx / cy must have room for n_c0 + T points.
 * t_values_host: per-step time-token value (host array of T floats) or NULL.  idx_hist, logp_hist [B,T].
 * tc_weights: NULL = fp32 FFMA query stream; otherwise the bf16 blob of aline_query_stream_tc, with tckv a buffer of
 * aline_tc_kv_bytes(m, B, min(n_c0 + T - 1 + n_sel, aline_tc_fast_max_keys(m))) bytes for the fast kernel's operand
 * blocks (NULL: general tcgen05 kernel only). */
int aline_rollout(const aline_model* m, const float* qx, const float* qy, uint8_t* alive, const float* eq, float* cx,
                  float* cy, int32_t B, int32_t nq, int32_t n_c0, int32_t ctx_cap, const float* target_x, int32_t n_td,
                  const int32_t* tgt_slot, int32_t n_sel, float* kv, int32_t kv_slots, float* logits, int32_t T,
                  const float* t_values_host, int64_t* idx_hist, float* logp_hist, const void* tc_weights, void* tckv,
                  void* stream);

/* aline_rollout with the row-major candidate embeddings eq_rm [B,nq,d] (optional, see aline_query_stream_tc_ex). */
int aline_rollout_ex(const aline_model* m, const float* qx, const float* qy, uint8_t* alive, const float* eq,
                     const float* eq_rm, float* cx, float* cy, int32_t B, int32_t nq, int32_t n_c0, int32_t ctx_cap,
                     const float* target_x, int32_t n_td, const int32_t* tgt_slot, int32_t n_sel, float* kv,
                     int32_t kv_slots, float* logits, int32_t T, const float* t_values_host, int64_t* idx_hist,
                     float* logp_hist, const void* tc_weights, void* tckv, void* stream);

/* ------------------------------------------------------- GP prior draws ---- */

/* Kernel families of GPTask (tasks/gaussian_process.py:59-60): 0 rbf, 1 matern12, 2 matern32, 3 matern52. */

/* Global scratch bytes aline_gp_sample needs (0 when the packed triangle fits in shared memory, N <= ~335). */
size_t aline_gp_scratch_bytes(int32_t B, int32_t N);

/* GPTask.generate_gp_data (tasks/gaussian_process.py:366-417), batched: for every b
 *     K = scale[b] * k_{type[b]}(x[b], x[b]; lengthscales[b]) + jitter I;  L = chol(K);  y[b] = L z[b] + noise eps[b]
 * x [B,N,dx], lengthscales [B,dx], scale [B], kernel_type [B] int32, z / eps / y [B,N].  Optional outputs (NULL to
 * skip): L_out, K_out [B,N,N]; info [B] int32 = 0 or 1 if the matrix was not positive definite (y = NaN). */
int aline_gp_sample(const float* x, int32_t B, int32_t N, int32_t dim_x, const float* lengthscales, const float* scale,
                    const int32_t* kernel_type, const float* z, const float* eps, float jitter, float noise_scale,
                    float* y, float* L_out, float* K_out, int32_t* info, void* scratch, size_t scratch_bytes,
                    void* stream);

/* GPTask.compute_kernel_matrix (tasks/gaussian_process.py:194-317) for one pair: x1 [N,dx], x2 [M,dx],
 * lengthscales [dx], scale [1] (device) -> K [N,M]. */
int aline_gp_kernel_matrix(const float* x1, const float* x2, int32_t N, int32_t M, int32_t dim_x,
                           const float* lengthscales, const float* scale, int32_t kernel_type, float* K, void* stream);

/* Hardware self-test of the tcgen05 / TMEM / TMA building blocks: D [128,N] = A [128,K] * B [N,K]^T with bf16-rounded
 * operands and fp32 accumulation (one thread block).  B_packed (optional): B already packed as bf16 in the
 * core-matrix tiled layout (csrc/tc.cuh), fetched with one TMA bulk copy instead of element-wise staging. */
int aline_tc_selftest(const float* A, const float* B, int32_t N, int32_t K, float* D, const void* B_packed,
                      void* stream);
/* Same product with the A operand staged in tensor memory (tcgen05.st, packed bf16 pairs) instead of shared memory. */
int aline_tc_selftest_tmem_a(const float* A, const float* B, int32_t N, int32_t K, float* D, void* stream);

/* CensoredSigmoidNormal(loc, scale, lower_lim, upper_lim).log_prob(value), element-wise over n entries
 * (distributions/censored_sigmoid_normal.py:47-86).  bad_flag as above. */
int aline_censored_sigmoid_normal_log_prob(const float* loc, const float* scale, const float* value,
                                           float lower_lim, float upper_lim, int64_t n, float* out,
                                           int32_t* bad_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALINE_B200_H */
