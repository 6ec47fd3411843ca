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

#define ALINE_ABI_VERSION 1

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

/* CensoredSigmoidNormal(loc, scale, lower_lim, upper_lim).log_prob(value), element-wise over n entries
 * (distributions/censored_sigmoid_normal.py:47-86).  bad_flag as above. */
int aline_censored_sigmoid_normal_log_prob(const float* loc, const float* scale, const float* value,
                                           float lower_lim, float upper_lim, int64_t n, float* out,
                                           int32_t* bad_flag, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ALINE_B200_H */
