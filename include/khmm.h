/*
 * khmm.h -- C ABI of libkhmm.so, the B200 (sm_100a) engine behind kerehmm_b200.
 *
 * The reference (keryil/kereHMM) is pure Python and has NO plugin / FFI interface
 * (SURVEY.md 8(b)); its boundary is the method surface of AbstractHMM / DiscreteHMM /
 * GaussianHMM.  Each entry point below names the reference method whose inner loops it
 * replaces (paths relative to the reference checkout).  The Python classes in
 * kerehmm_b200/ bind these through ctypes (kerehmm_b200/_lib.py); INTEGRATION.md shows
 * the stub a kereHMM maintainer would add to call them from the reference classes.
 *
 * Conventions
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer on `device`
 *     unless its name ends in `_host`; the caller owns all buffers (no allocation inside);
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on that stream;
 *   - return value: 0 = KHMM_OK, otherwise an error code; khmm_last_error() returns a
 *     thread-local message;
 *   - batch layout: observations obs[B][T] (row per sequence), optional lengths[B]
 *     (int32, NULL = every sequence has T steps, 1 <= lengths[b] <= T);
 *     alpha/beta/gamma/V are [B][T][N], scale is [B][T], per-sequence results are [B];
 *   - transition matrix A[N][N] is row-major "from i to j" exactly as
 *     AbstractHMM.transitionMatrix; AT is its transpose; the discrete emission table is
 *     SYMBOL-MAJOR Bsm[M][N] (Bsm[k][j] = emissionDistributions[j].probabilities[k]) so
 *     that the N emission values of one observation are contiguous;
 *   - 1 <= N <= 1024.
 */
#ifndef KHMM_H
#define KHMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KHMM_OK 0
#define KHMM_ERR_INVALID 1     /* bad argument */
#define KHMM_ERR_CUDA 2        /* CUDA runtime error (message has the cudaError string) */
#define KHMM_ERR_UNSUPPORTED 3 /* shape outside what the kernels cover */

/* dtype codes (observations, paths) */
#define KHMM_U8 0
#define KHMM_U16 1
#define KHMM_I32 2
#define KHMM_I64 3
#define KHMM_F32 4
#define KHMM_F64 5

/* gamma flavours for khmm_backward_* */
#define KHMM_GAMMA_REFERENCE 0 /* alpha*beta*c_t, AbstractHMM.gamma (abstract_hmm.py:291-297) */
#define KHMM_GAMMA_POSTERIOR 1 /* alpha*beta: the true state posterior */

typedef void *khmm_stream_t;

const char *khmm_last_error(void);
int khmm_version(void);
/* Fills SM count / compute capability / memory of `device`; fails if there is no CUDA device. */
int khmm_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem);

/* out[c][r] = in[r][c] for a rows x cols matrix (used for AT and for state-major <-> symbol-major). */
int khmm_transpose_f32(int rows, int cols, const float *in, float *out, khmm_stream_t stream);
int khmm_transpose_f64(int rows, int cols, const double *in, double *out, khmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Viterbi -- AbstractHMM.viterbi_path (kerehmm/abstract_hmm.py:203-239), discrete emissions.
 * float64 max-plus DP in the reference's association order, first-index argmax ties, psi
 * backpointers kept in HBM (1 byte per entry when N <= 256, else 2), then a backtrace.
 * logpi[N], logA[N][N], logBsm[M][N] are the HOST-computed np.log tables (device copies):
 * log is never evaluated on the device so results are bit-identical to the reference.
 * backptr: workspace of khmm_viterbi_workspace_bytes(B,T,N) bytes.
 * path: [B][T] of path_dtype (U8/U16/I32/I64); logp: [B] float64 = max_j delta[len-1][j].
 */
size_t khmm_viterbi_workspace_bytes(int64_t B, int64_t T, int N);
int khmm_viterbi_discrete_f64(int64_t B, int64_t T, int N, int M,
                              const double *logpi, const double *logA, const double *logBsm,
                              const void *obs, int obs_dtype, const int32_t *lengths,
                              void *backptr, size_t backptr_bytes,
                              void *path, int path_dtype, double *logp,
                              khmm_stream_t stream);

/* EXTENSION: the same DP for 1-D Gaussian emissions.  The reference's viterbi_path fails on
 * float observations (abstract_hmm.py:234-237), so there is no reference output to match; the definition
 * is log b_j(x) = (-0.5 z) z + logc_j with z = (x - mean_j) / sd_j, one IEEE rounding per operation in that order, and
 * logc_j = -log(sd_j) - log(sqrt(2 pi)) computed on the HOST.  mean_sd_logc is [3][N] float64: mean, sd, logc.
 * obs: [B][T] f32 or f64.  Bit-exact against oracle/hmm_oracle.py::viterbi_gaussian_1d ("parity unpinned" beyond it). */
int khmm_viterbi_gauss_f64(int64_t B, int64_t T, int N,
                           const double *logpi, const double *logA, const double *mean_sd_logc,
                           const void *obs, int obs_dtype, const int32_t *lengths,
                           void *backptr, size_t backptr_bytes,
                           void *path, int path_dtype, double *logp,
                           khmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Scaled forward -- AbstractHMM.forward (abstract_hmm.py:250-285) with the emission
 * evaluation (abstract_hmm.py:353-367 -> distribution.py:35-36 / :151-155) fused in.
 * alpha (may be NULL), scale (may be NULL; c_t is the SUM, not its reciprocal), loglik
 * (may be NULL; [B] float64 = sum_t log c_t, the true log P(O)).
 * Gaussian: D = 1, pdf = exp(-((x-mean)/sd)^2/2)/(sd*sqrt(2*pi)); `sd` is the reference's
 * `variance` attribute, which it passes to scipy as the scale.
 */
int khmm_forward_discrete_f32(int64_t B, int64_t T, int N, int M,
                              const float *pi, const float *A, const float *Bsm,
                              const void *obs, int obs_dtype, const int32_t *lengths,
                              float *alpha, float *scale, double *loglik, khmm_stream_t stream);
int khmm_forward_discrete_f64(int64_t B, int64_t T, int N, int M,
                              const double *pi, const double *A, const double *Bsm,
                              const void *obs, int obs_dtype, const int32_t *lengths,
                              double *alpha, double *scale, double *loglik, khmm_stream_t stream);
int khmm_forward_gauss_f32(int64_t B, int64_t T, int N,
                           const float *pi, const float *A, const float *mean, const float *sd,
                           const void *obs, int obs_dtype, const int32_t *lengths,
                           float *alpha, float *scale, double *loglik, khmm_stream_t stream);
int khmm_forward_gauss_f64(int64_t B, int64_t T, int N,
                           const double *pi, const double *A, const double *mean, const double *sd,
                           const void *obs, int obs_dtype, const int32_t *lengths,
                           double *alpha, double *scale, double *loglik, khmm_stream_t stream);
/* Emissions given as a dense matrix E[B][T][N] (multivariate Gaussians, user emissions). */
int khmm_forward_dense_f32(int64_t B, int64_t T, int N, const float *pi, const float *A, const float *E,
                           const int32_t *lengths, float *alpha, float *scale, double *loglik,
                           khmm_stream_t stream);
int khmm_forward_dense_f64(int64_t B, int64_t T, int N, const double *pi, const double *A, const double *E,
                           const int32_t *lengths, double *alpha, double *scale, double *loglik,
                           khmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Scaled backward -- AbstractHMM.backward (abstract_hmm.py:327-351), ghmm convention:
 * beta[len-1] = 1, beta[t][i] = sum_j A[i][j] b_j(o_{t+1}) beta[t+1][j] / c_{t+1}.
 * Takes AT (transpose of A) so the per-state reads coalesce.  scale: forward's c [B][T].
 * beta (may be NULL).  If alpha != NULL and gamma != NULL the kernel also writes
 * gamma = alpha*beta*c (KHMM_GAMMA_REFERENCE, abstract_hmm.py:291-297) or alpha*beta
 * (KHMM_GAMMA_POSTERIOR) in the same pass.
 */
int khmm_backward_discrete_f32(int64_t B, int64_t T, int N, int M, const float *AT, const float *Bsm,
                               const void *obs, int obs_dtype, const int32_t *lengths, const float *scale,
                               const float *alpha, float *beta, float *gamma, int gamma_kind,
                               khmm_stream_t stream);
int khmm_backward_discrete_f64(int64_t B, int64_t T, int N, int M, const double *AT, const double *Bsm,
                               const void *obs, int obs_dtype, const int32_t *lengths, const double *scale,
                               const double *alpha, double *beta, double *gamma, int gamma_kind,
                               khmm_stream_t stream);
int khmm_backward_gauss_f32(int64_t B, int64_t T, int N, const float *AT, const float *mean, const float *sd,
                            const void *obs, int obs_dtype, const int32_t *lengths, const float *scale,
                            const float *alpha, float *beta, float *gamma, int gamma_kind,
                            khmm_stream_t stream);
int khmm_backward_gauss_f64(int64_t B, int64_t T, int N, const double *AT, const double *mean, const double *sd,
                            const void *obs, int obs_dtype, const int32_t *lengths, const double *scale,
                            const double *alpha, double *beta, double *gamma, int gamma_kind,
                            khmm_stream_t stream);
int khmm_backward_dense_f32(int64_t B, int64_t T, int N, const float *AT, const float *E,
                            const int32_t *lengths, const float *scale,
                            const float *alpha, float *beta, float *gamma, int gamma_kind,
                            khmm_stream_t stream);
int khmm_backward_dense_f64(int64_t B, int64_t T, int N, const double *AT, const double *E,
                            const int32_t *lengths, const double *scale,
                            const double *alpha, double *beta, double *gamma, int gamma_kind,
                            khmm_stream_t stream);

/* AbstractHMM.gamma (abstract_hmm.py:291-297) on caller-supplied arrays: out = alpha*beta*scale[...,None]
 * (KHMM_GAMMA_REFERENCE) or alpha*beta (KHMM_GAMMA_POSTERIOR); rows = B*T. */
int khmm_gamma_f32(int64_t rows, int N, const float *alpha, const float *beta, const float *scale, float *out,
                   int gamma_kind, khmm_stream_t stream);
int khmm_gamma_f64(int64_t rows, int N, const double *alpha, const double *beta, const double *scale, double *out,
                   int gamma_kind, khmm_stream_t stream);
/* AbstractHMM.zi (abstract_hmm.py:299-325) materialised for ONE sequence (API compatibility; the
 * training path never materialises it): zi[t][i][j] = alpha[t][i]*A[i][j]*E[t+1][j]*beta[t+1][j],
 * t < T-1, with E the dense emission matrix [T][N] and the reference's left-to-right product order. */
int khmm_zi_f64(int64_t T, int N, const double *A, const double *E, const double *alpha, const double *beta,
                double *zi, khmm_stream_t stream);
/* Dense emission matrix E[b][t][j] = b_j(o_{b,t}) (AbstractHMM.emission_probability(None, o),
 * abstract_hmm.py:362-366) for discrete tables and 1-D Gaussians. */
int khmm_emissions_discrete_f64(int64_t B, int64_t T, int N, int M, const double *Bsm, const void *obs,
                                int obs_dtype, double *E, khmm_stream_t stream);
int khmm_emissions_gauss_f64(int64_t B, int64_t T, int N, const double *mean, const double *sd, const void *obs,
                             int obs_dtype, double *E, khmm_stream_t stream);
/* Forward / backward with FUSED full-covariance Gaussian emissions, D <= KHMM_GAUSS_FULL_DMAX features
 * (GaussianDistribution with a covariance matrix, distribution.py:153; AbstractHMM.forward / backward,
 * abstract_hmm.py:250-285, 327-351): no [B][T][N] emission matrix is materialised.  means[N][D], Linv[N][D][D],
 * lognorm[N] as for khmm_emissions_gauss_full_f64 below (float64 whatever the recursion's type: the density is
 * evaluated in float64 with that function's operation order, then cast); obs[B][T][D] (f32/f64). */
#define KHMM_GAUSS_FULL_DMAX 8
int khmm_forward_gaussfull_f32(int64_t B, int64_t T, int N, int D, const float *pi, const float *A, const double *means,
                               const double *Linv, const double *lognorm, const void *obs, int obs_dtype,
                               const int32_t *lengths, float *alpha, float *scale, double *loglik, khmm_stream_t stream);
int khmm_forward_gaussfull_f64(int64_t B, int64_t T, int N, int D, const double *pi, const double *A, const double *means,
                               const double *Linv, const double *lognorm, const void *obs, int obs_dtype,
                               const int32_t *lengths, double *alpha, double *scale, double *loglik, khmm_stream_t stream);
int khmm_backward_gaussfull_f32(int64_t B, int64_t T, int N, int D, const float *AT, const double *means, const double *Linv,
                                const double *lognorm, const void *obs, int obs_dtype, const int32_t *lengths,
                                const float *scale, const float *alpha, float *beta, float *gamma, int gamma_kind,
                                khmm_stream_t stream);
int khmm_backward_gaussfull_f64(int64_t B, int64_t T, int N, int D, const double *AT, const double *means, const double *Linv,
                                const double *lognorm, const void *obs, int obs_dtype, const int32_t *lengths,
                                const double *scale, const double *alpha, double *beta, double *gamma, int gamma_kind,
                                khmm_stream_t stream);
/* Full-covariance Gaussians (distribution.py:153): means[N][D], Linv[N][D][D] = inverse of the lower
 * Cholesky factor of each covariance (row-major, lower triangular), lognorm[N] =
 * -0.5*(D*log(2*pi) + log det cov); obs[B][T][D] (f32/f64).  E = exp(lognorm - 0.5*|Linv (x-mean)|^2). */
int khmm_emissions_gauss_full_f64(int64_t B, int64_t T, int N, int D, const double *means, const double *Linv,
                                  const double *lognorm, const void *obs, int obs_dtype, double *E,
                                  khmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused forward-backward log-likelihood (no alpha/beta in HBM) for small state spaces
 * (N <= KHMM_SMALL_N_MAX), emission evaluation fused, everything but the observations in
 * registers: one thread per (sequence, direction) in packed FP32, or -- N = 5..8 -- one warp per
 * 16 sequences with the matrix-vector products as mma.sync tiles on split operands (same
 * fp32-class accuracy).  loglik[B] = sum_t log c_t (forward); loglik_bwd[B] (may be
 * NULL: forward only) = the same likelihood recovered from the backward recursion, which runs
 * concurrently with its own scaling s_t: log(sum_i pi_i b_i(o_0) beta_0(i)) = sum_t log s_t +
 * log(pi . w_0)  -- equal to loglik up to rounding.
 */
#define KHMM_SMALL_N_MAX 16
int khmm_fwdbwd_loglik_gauss_f32(int64_t B, int64_t T, int N,
                                 const float *pi, const float *A, const float *mean, const float *sd,
                                 const void *obs, int obs_dtype, const int32_t *lengths,
                                 double *loglik, double *loglik_bwd, khmm_stream_t stream);
int khmm_fwdbwd_loglik_discrete_f32(int64_t B, int64_t T, int N, int M,
                                    const float *pi, const float *A, const float *Bsm,
                                    const void *obs, int obs_dtype, const int32_t *lengths,
                                    double *loglik, double *loglik_bwd, khmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Tensor-core forward-backward log-likelihood for large state spaces: N a multiple of 128, 128 <= N <= 1024
 * (the Python layer pads any 65 <= N <= 1024 with unreachable states), any batch size; the *_ragged_* entry
 * points below take per-sequence lengths.  ALL T-1 steps of the forward (abstract_hmm.py:266-270) and of the
 * self-scaled backward (:339-349) recursion run in ONE cooperative launch of a persistent tcgen05 kernel
 * (kind::tf32 or kind::f16 with bf16 operands, fp32 accumulation in TMEM, TMA-staged operands): the work is a
 * set of (128-sequence row tile, direction) chains of column-tile jobs that wait for their inputs on per-chain
 * counters in global memory instead of on a kernel boundary; the emission evaluation, the deferred scaling and the
 * row sums are fused in the epilogue.  Where a cooperative launch is unavailable the same
 * arithmetic runs with one launch per step.  A and AT: the transition matrix and its transpose.  workspace:
 * khmm_fwdbwd_loglik_tc_workspace_bytes(B, N) bytes, 256-byte aligned.  loglik_bwd may be NULL (forward only).
 * tf32 multiplies: log-likelihoods agree with the float64 reference to ~1e-6 relative (stated bound 2e-5 in the
 * tests); bf16: 7e-6 at T = 1024.
 */
/* operand_precision */
#define KHMM_TC_TF32 0 /* tf32 operands (10-bit mantissa), fp32 accumulation */
#define KHMM_TC_BF16 1 /* bf16 operands (8-bit mantissa), fp32 accumulation: twice the MMA rate; needs N % 256 == 0 and
                          >= one (128 sequences x 256 states x direction) job per SM, else KHMM_ERR_UNSUPPORTED */
size_t khmm_fwdbwd_loglik_tc_workspace_bytes(int64_t B, int N);
int khmm_fwdbwd_loglik_tc_gauss_f32(int64_t B, int64_t T, int N, const float *pi, const float *A, const float *AT,
                                    const float *mean, const float *sd, const void *obs, int obs_dtype,
                                    void *workspace, size_t workspace_bytes, double *loglik, double *loglik_bwd,
                                    int operand_precision, khmm_stream_t stream);
int khmm_fwdbwd_loglik_tc_discrete_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A,
                                       const float *AT, const float *Bsm, const void *obs, int obs_dtype,
                                       void *workspace, size_t workspace_bytes, double *loglik, double *loglik_bwd,
                                       int operand_precision, khmm_stream_t stream);
/* The same for ragged batches: lengths[b] in [1, T] (device int32[B]; NULL = equal lengths), rows padded to T with
 * anything.  Both recursions run through the padding with the emission factor set to 1 (a continuation that emits
 * "anything" leaves the likelihood unchanged), so a ragged batch costs what the padded batch costs. */
int khmm_fwdbwd_loglik_tc_gauss_ragged_f32(int64_t B, int64_t T, int N, const float *pi, const float *A, const float *AT,
                                           const float *mean, const float *sd, const void *obs, int obs_dtype,
                                           const int32_t *lengths, void *workspace, size_t workspace_bytes, double *loglik,
                                           double *loglik_bwd, int operand_precision, khmm_stream_t stream);
int khmm_fwdbwd_loglik_tc_discrete_ragged_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A,
                                              const float *AT, const float *Bsm, const void *obs, int obs_dtype,
                                              const int32_t *lengths, void *workspace, size_t workspace_bytes,
                                              double *loglik, double *loglik_bwd, int operand_precision,
                                              khmm_stream_t stream);


/* ---------------------------------------------------------------------------------------
 * Baum-Welch E-step -- the statistics DiscreteHMM.do_pass forms (discrete_hmm.py:117-159)
 * from forward, backward, AbstractHMM.zi (abstract_hmm.py:299-325) and gamma (:291-297),
 * with the reference's weighting (gamma carries c_t, zi carries c_{t+1}, gamma[0] is
 * overwritten by smooth1d(gamma[0])), summed over the sequences
 * of the batch (Q8).  Call order for one batch chunk:
 *   khmm_forward_discrete_*   -> alpha, scale (stash)
 *   khmm_estep_discrete_*     -> adds into counts[], writes V[b][t][j] = b_j(o_t) beta_t(j)
 *   khmm_xi_accumulate_*      -> adds sum_{b,t<len-1} alpha[b][t][i] * V[b][t+1][j] into counts.xi
 * counts is ONE float64 buffer of khmm_counts_len(N,M) doubles (zeroed by the caller before
 * the first chunk; this is the buffer that is all-reduced across GPUs):
 *   [ pi: N | xi_raw: N*N | gam: N | emis_num (symbol-major [M][N]): M*N | emis_den: N |
 *     loglik_sum: 1 | nseq: 1 | flags: 1 ]
 * xi_raw is sum alpha (x) V WITHOUT the elementwise A factor (the M-step applies it).
 * flags != 0 means some sequence hit the reference's ZeroDivisionError case
 * (every gamma[0] entry < 1e-10, util.py:57-58).
 */
size_t khmm_counts_len(int N, int M);
int khmm_estep_discrete_f32(int64_t B, int64_t T, int N, int M, const float *AT, const float *Bsm,
                            const void *obs, int obs_dtype, const int32_t *lengths,
                            const float *alpha, const float *scale, float *V, double *counts,
                            khmm_stream_t stream);
int khmm_estep_discrete_f64(int64_t B, int64_t T, int N, int M, const double *AT, const double *Bsm,
                            const void *obs, int obs_dtype, const int32_t *lengths,
                            const double *alpha, const double *scale, double *V, double *counts,
                            khmm_stream_t stream);
int khmm_xi_accumulate_f32(int64_t B, int64_t T, int N, const int32_t *lengths,
                           const float *alpha, const float *V, double *xi_raw, khmm_stream_t stream);
int khmm_xi_accumulate_f64(int64_t B, int64_t T, int N, const int32_t *lengths,
                           const double *alpha, const double *V, double *xi_raw, khmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused tensor-core E-step for N = 128 states (config 4; the Python layer pads 8 <= N < 128 with unreachable
 * states; per-sequence lengths: the *_ragged_* entry point below).  One call does, for a batch chunk, what the
 * three calls above do -- forward (abstract_hmm.py:250-285), backward (:327-351), gamma (:291-297), zi summed
 * over t (:299-325, discrete_hmm.py:155) and the count sums of DiscreteHMM.do_pass (discrete_hmm.py:117-159)
 * with the same reference weighting -- in two persistent tcgen05 kernels (fp32 accumulation in TMEM):
 *   forward   thread = sequence; the row W_t = c_t alpha_t stays in tensor memory as the next step's A operand
 *             and is stashed once in HBM (TMA tensor stores);
 *   backward  thread = state, four quarter-tiles of 32 sequences per CTA as independent pipelines: the stash
 *             comes back by TMA, zi accumulates in TMEM over the whole tile, emission values are read and the
 *             emission counts reduced (red.global.add) straight from registers, one coalesced 128-byte row per
 *             warp instruction.
 * Adds into `counts` (same layout as above).  workspace: khmm_estep_tc_workspace_bytes(B,T,N,M) bytes,
 * 1024-byte aligned (holds the fp32 stash).  alpha_out / scale_out (may be NULL): caller-visible copies of the
 * stash -- tile-major [ceil(B/128)][T][128][128] floats, W[b/128][t][b%128][:] = c_t * alpha_t, one contiguous
 * 64 KB block per (tile of 128 sequences, step) -- and of the scale factors, tile-major [ceil(B/128)][T][128],
 * c[b/128][t][b%128] (used by the tests).
 * This entry point multiplies in tf32: statistics agree with the float64 reference to ~1e-3 relative per
 * (sequence, step) term, unbiased, so re-estimated parameters converge as 1/sqrt(rows) down to the systematic
 * 1e-5 ... 1e-4 left by the rounding of A (tests state the bounds); khmm_estep_tc_discrete_ex_f32 below takes the
 * operand precision (the Python layer's default is KHMM_TC_BF16X3).
 */
size_t khmm_estep_tc_workspace_bytes(int64_t B, int64_t T, int N, int M);
int khmm_estep_tc_discrete_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *AT,
                               const float *Bsm, const void *obs, int obs_dtype, void *workspace,
                               size_t workspace_bytes, double *counts, float *alpha_out, float *scale_out,
                               khmm_stream_t stream);

/* The same with per-sequence lengths (1 <= lengths[b] <= T, device int32[B]; NULL = every sequence has T steps): rows
 * are padded to T, the padding symbols may be anything.  The statistics are those of the unpadded sequences
 * (DiscreteHMM.do_pass on each sequence, summed); the recursion simply runs through the padding and three per-row
 * switches remove it from the counts, so a ragged batch costs what the padded batch costs. */
int khmm_estep_tc_discrete_ragged_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *AT,
                                      const float *Bsm, const void *obs, int obs_dtype, const int32_t *lengths,
                                      void *workspace, size_t workspace_bytes, double *counts, float *alpha_out,
                                      float *scale_out, khmm_stream_t stream);

/* The same with a choice of operand precision for the recursion's contraction (alpha_{t-1} A and A (b . beta_{t+1})):
 *   KHMM_TC_TF32    one tf32 MMA (10-bit mantissa).  The rounding of A itself is a fixed perturbation of the model and
 *                   leaves a systematic 1e-5 (median) ... 1e-4 (worst entry) relative error in the counts.
 *   KHMM_TC_BF16X3  split operands: x = hi + lo with hi = bf16(x), lo = bf16(x - hi), three bf16 MMAs
 *                   (hi hi + hi lo + lo hi) into the same fp32 accumulator: 2^-17 operand precision in the same 64 KB
 *                   of shared memory, 1.5 x the tensor time of the contraction.  Counts and re-estimated parameters
 *                   then agree with the float64 reference to float32 accuracy (tests state the bound).
 * The accumulation of zi (the second contraction) stays tf32 in both modes: its per-term rounding errors are unbiased
 * and average out over the B*T rows it sums. */
#define KHMM_TC_BF16X3 2
int khmm_estep_tc_discrete_ex_f32(int64_t B, int64_t T, int N, int M, const float *pi, const float *A, const float *AT,
                                  const float *Bsm, const void *obs, int obs_dtype, const int32_t *lengths,
                                  void *workspace, size_t workspace_bytes, double *counts, float *alpha_out,
                                  float *scale_out, int operand_precision, khmm_stream_t stream);

/* Measurement hook (bench.py's roofline): khmm_estep_tc_timing(1) arms CUDA events on the launching stream around the
 * forward and the backward kernel of every following khmm_estep_tc_* call (per device, a ring of the last 64 calls;
 * nothing is created or recorded while disarmed).  khmm_estep_tc_kernel_ms(k, ...) returns the durations of the call
 * k calls back (0 = the most recent) on the current device and waits for that call to finish;
 * khmm_estep_tc_last_kernel_ms is k = 0. */
int khmm_estep_tc_timing(int enable);
int khmm_estep_tc_kernel_ms(int calls_back, float *forward_ms_host, float *backward_ms_host);
int khmm_estep_tc_last_kernel_ms(float *forward_ms_host, float *backward_ms_host);

/* Development switches: name = "KHMM_RC_MODE" | "KHMM_RC_BN" | "KHMM_VIT_SCREEN" | "KHMM_BW_BACKWARD" |
 * "KHMM_SMALL_MMA" (DESIGN.md 4a), value -1
 * restores the library's default.  The environment variable of the same name sets the initial value and is read once per
 * process; no production call reads the environment. */
int khmm_debug_switch(const char *name, int value);
/* Development hook: which backward kernel the fused E-step uses -- 3 = thread per state, four quarter-tile pipelines per
 * CTA (the default), 1 = the round-1 thread-per-sequence kernel, 0 = back to the default / $KHMM_BW_BACKWARD.  Both compute
 * the same statistics in both operand precisions (tests/test_gpu_train_loop.py runs them against each other). */
int khmm_debug_bw_backward_version(int version);
/* Development hook: arm (device buffer of 2*8*32 int64) or disarm (NULL) a clock64 timeline of the warp roles
 * of the backward and forward kernels (block 0, steps 8..15; tools/estep_forward_timeline.py prints it). */
int khmm_debug_bw_trace(int64_t *dev_buf);
/* Same for the whole-recursion kernel (device buffer of 4*2*16 int64: block 0, steps 8..11, its two lanes). */
int khmm_debug_rc_trace(int64_t *dev_buf);

/* ---------------------------------------------------------------------------------------
 * Batched ancestral sampling -- AbstractHMM.simulate / transition / emit (abstract_hmm.py:444-469,
 * distribution.py:38-39, :157-161).  Per step: transition (from pi for the first step, then from row
 * A[current]), then emit from the new state.  cum_pi[N], cumA[N][N] (row-wise), cumB[N][M] (state-major,
 * row-wise) are the float64 CDFs numpy.random.choice builds (cumsum(p) / cumsum(p)[-1]); the kernel does
 * choice's right-sided searchsorted.  Gaussian emissions: mean + sd * z, z by Box-Muller in float64.
 * uniforms == NULL: Philox4x32-10 keyed by `seed`, counter = (sequence, step) -- a sequence's stream is
 * independent of the batch shape.  uniforms != NULL: [B][T][3] float64 in [0,1) used instead
 * (state draw, emission draw, second Box-Muller draw): with identical uniforms states and symbols are
 * identical to the NumPy procedure (tests).  The reference's MT19937 stream is not reproduced.
 * states/symbols: [B][T] int32; values: [B][T] float64. */
int khmm_simulate_discrete(int64_t B, int64_t T, int N, int M, const double *cum_pi, const double *cumA,
                           const double *cumB, uint64_t seed, const double *uniforms, int32_t *states,
                           int32_t *symbols, khmm_stream_t stream);
int khmm_simulate_gauss(int64_t B, int64_t T, int N, const double *cum_pi, const double *cumA, const double *mean,
                        const double *sd, uint64_t seed, const double *uniforms, int32_t *states, double *values,
                        khmm_stream_t stream);
/* Known-answer hook for the generator: n blocks of (4 counter words, 2 key words) -> 4 output words each. */
int khmm_philox4x32_10(int n, const uint32_t *ctr_key, uint32_t *out, khmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Baum-Welch for 1-D Gaussian emissions -- an EXTENSION: GaussianHMM.do_pass (gaussian_hmm.py:29-124) is
 * stale code that raises at HEAD, so this is textbook EM on the reference's scaled recursions:
 * gamma = alpha*beta, xi_t = alpha_t(i) A_ij b_j(x_{t+1}) beta_{t+1}(j) / c_{t+1}; pi = mean of smooth1d(gamma_0),
 * A rows = smooth1d(rownorm(xi/gam)) (util.py:56-60; the 2-D rule of util.py:46-55 raises for rows with several
 * tiny entries); mean_i = sum gamma x / sum gamma,
 * sd_i = sqrt(max(sum gamma x^2 / sum gamma - mean_i^2, sd_min^2)) (sd = the `variance` attribute).
 * Call order per chunk: khmm_forward_gauss_* -> khmm_backward_gauss_* (beta) -> khmm_estep_gauss_* (adds
 * into counts, writes V[b][t][j] = b_j(x_t) beta_t(j) / c_t) -> khmm_xi_accumulate_* (alpha, V, counts.xi)
 * -> all-reduce -> khmm_mstep_gauss_f64.  counts: khmm_counts_len_gauss(N) doubles
 *   [ pi: N | xi_raw: N*N | gam: N | S0 = sum gamma: N | S1 = sum gamma x: N | S2 = sum gamma x^2: N |
 *     loglik_sum, nseq, flags ].
 * mstep status bits: 1 ZeroDivisionError case (every entry of a row < 1e-10), 4 sanity_check failure,
 * 8 a state with no posterior mass (mean undefined). */
size_t khmm_counts_len_gauss(int N);
int khmm_estep_gauss_f32(int64_t B, int64_t T, int N, const float *mean, const float *sd, const void *obs,
                         int obs_dtype, const int32_t *lengths, const float *alpha, const float *beta,
                         const float *scale, float *V, double *counts, khmm_stream_t stream);
int khmm_estep_gauss_f64(int64_t B, int64_t T, int N, const double *mean, const double *sd, const void *obs,
                         int obs_dtype, const int32_t *lengths, const double *alpha, const double *beta,
                         const double *scale, double *V, double *counts, khmm_stream_t stream);
int khmm_mstep_gauss_f64(int N, const double *counts, const double *A, double sd_min, double *pi_new, double *A_new,
                         double *mean_new, double *sd_new, int32_t *status, khmm_stream_t stream);

/* M-step on the device -- discrete_hmm.py:132-161 + util.smooth_probabilities (util.py:39-63),
 * float64, operation for operation.  Reads counts (after the all-reduce), A (old, [N][N]);
 * writes pi_new[N], A_new[N][N], B_new state-major [N][M].  status[0] (device int32) gets a
 * bit mask: 1 = ZeroDivisionError case, 2 = the 2-D smoothing ValueError case (2 <= k < n
 * small entries in a row, util.py:51-53), 4 = a sanity_check (isclose(sum,1)) failure. */
int khmm_mstep_discrete_f64(int N, int M, const double *counts, const double *A,
                            double *pi_new, double *A_new, double *B_new, int32_t *status,
                            khmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Device-resident training loop.  In the reference every Baum-Welch iteration ends on the host:
 * DiscreteHMM.do_pass installs the re-estimated pi / A / B as new NumPy arrays (discrete_hmm.py:168-173) and
 * AbstractHMM.train (abstract_hmm.py:400-441) calls it once per iteration.  khmm_train_commit_discrete_* keeps that
 * hand-over on the GPU: it installs the outputs of khmm_mstep_discrete_f64 (pi_src[N], A_src[N][N], B_src[N][M]
 * state-major, float64) as the master parameters pi64 / A64 / B64 and writes the operand copies the recursion
 * kernels read -- pi_op[Np], A_op[Np][Np], AT_op[Np][Np] and the SYMBOL-MAJOR table Bsm_op[M][Np] -- padded to
 * Np >= N states with unreachable states (start probability 0, no transitions, zero emission row).
 * status_it (may be NULL): the M-step's status word of this iteration; sticky (may be NULL): int32[2] =
 * {status bits of the first failing iteration, that iteration + 1}.  Once either is non-zero nothing is installed
 * any more: the masters keep the parameters of the last good iteration (the reference raises inside do_pass at
 * that point; the host raises the same exception when it reads `sticky` after the loop).
 * Passing the masters themselves as sources (pi_src == pi64, ...) only (re)writes the operand copies.
 */
int khmm_train_commit_discrete_f32(int N, int M, int Np, int iteration, const int32_t *status_it, int32_t *sticky,
                                   const double *pi_src, const double *A_src, const double *B_src, double *pi64,
                                   double *A64, double *B64, float *pi_op, float *A_op, float *AT_op, float *Bsm_op,
                                   khmm_stream_t stream);
int khmm_train_commit_discrete_f64(int N, int M, int Np, int iteration, const int32_t *status_it, int32_t *sticky,
                                   const double *pi_src, const double *A_src, const double *B_src, double *pi64,
                                   double *A64, double *B64, double *pi_op, double *A_op, double *AT_op, double *Bsm_op,
                                   khmm_stream_t stream);

/* min_max[0..1] (device int64) = smallest / largest symbol among the live positions (t < lengths[b]) of obs[B][T].
 * DiscreteDistribution.get_probability indexes a NumPy table with the symbol (distribution.py:35-36): a symbol >= M
 * (or < -M) raises IndexError and a negative one wraps; the kernels clamp, so the host checks the range once per
 * batch with this reduction and raises / wraps before any kernel sees the data. */
int khmm_symbol_range(int64_t B, int64_t T, const void *obs, int obs_dtype, const int32_t *lengths, int64_t *min_max,
                      khmm_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Multi-GPU: one process per GPU, sequences sharded across ranks, and ONE collective per Baum-Welch iteration --
 * the in-place sum of the packed float64 count buffer (layout above) over all ranks, after which every rank runs
 * the same M-step on identical bits (no parameter broadcast).  The reference is single-process; the sums are the
 * per-sequence statistics of DiscreteHMM.do_pass (discrete_hmm.py:117-159) added over all sequences (SURVEY.md
 * 8(e)).  NCCL is bound at run time (libnccl.so.2 already loaded in the process, else found by the loader, else
 * $KHMM_NCCL_LIB); without it these return KHMM_ERR_UNSUPPORTED.
 *   khmm_comm_unique_id     rank 0 creates the 128-byte rendezvous id; the caller ships it to the other ranks
 *                           (any side channel: torch.distributed store, MPI, a file);
 *   khmm_comm_init_rank     collective over all `world` ranks; binds `device`; *comm_out is an opaque handle;
 *   khmm_allreduce_counts   counts[n] += the other ranks' counts, asynchronous on `stream` (ncclAllReduce, sum, f64);
 *   khmm_comm_info          rank / world as given at init and the communicator's own rank count (ncclCommCount);
 *   khmm_comm_destroy       releases the communicator (NULL is a no-op).
 */
#define KHMM_COMM_ID_BYTES 128
int khmm_comm_nccl_version(int *version_host);
int khmm_comm_unique_id(void *id128_host);
int khmm_comm_init_rank(const void *id128_host, int rank, int world, int device, void **comm_out);
int khmm_comm_info(void *comm, int *rank_host, int *world_host, int *nccl_world_host);
int khmm_allreduce_counts(void *comm, double *counts, size_t n, khmm_stream_t stream);
int khmm_comm_destroy(void *comm);

#ifdef __cplusplus
}
#endif
#endif /* KHMM_H */
