/* b200ref.h -- C-ABI of the B200-native draw-comparison hot path.
 *
 * Drop-in boundary for mcmc-ref's diagnostics / stats / compare path.  The
 * reference is pure Python and has no FFI; each entry point below names the
 * reference interface it replaces (paths relative to the reference checkout).
 * The reference-side binding a maintainer would add is the ctypes stub shown in
 * INTEGRATION.md (mirrored by mcmc-db_b200/mcmc_ref_b200/_native.py).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - "_dev" pointers are CUDA device pointers owned by the caller.  The library
 *     never frees caller memory and allocates device memory only inside the
 *     caller-supplied workspace `ws` (size from b200ref_workspace_bytes).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *   - calls are synchronous from the caller's view: they return after the work
 *     on `stream` has completed (unless documented otherwise), so outputs are
 *     valid on return.  ctypes releases the GIL during the call.
 *   - return 0 on success, negative error code otherwise; the message is in
 *     b200ref_last_error() (thread-local).
 *   - canonical layout: x[P][C][N] fp64, draw index fastest (an Arrow/Parquet
 *     parameter column with rows sorted chain-major is exactly one x[p]).
 *   - non-finite input is rejected with B200REF_ERR_NONFINITE (documented
 *     divergence: the reference's behaviour on NaN is order-dependent).
 */
#ifndef B200REF_H
#define B200REF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200REF_VERSION 100

#define B200REF_OK 0
#define B200REF_ERR_ARG (-1)        /* bad shape / null pointer / nq too large   */
#define B200REF_ERR_WORKSPACE (-2)  /* workspace too small                        */
#define B200REF_ERR_CUDA (-3)       /* a CUDA runtime call or kernel failed       */
#define B200REF_ERR_NONFINITE (-4)  /* NaN or +-inf in the input draws            */

#define B200REF_MAX_QUANTILES 8

/* Output block of the fused pass: struct-of-arrays, each [P] unless noted.
 * Any pointer may be NULL (that output is skipped).  All device pointers. */
typedef struct b200ref_outputs {
  double* rhat;        /* max(rhat_bulk, rhat_folded)  -- diagnostics.py:13-40   */
  double* rhat_bulk;   /* _rhat(split(rank_normalize(x)))                         */
  double* rhat_tail;   /* _rhat(split(rank_normalize(fold(x))))                   */
  double* ess_bulk;    /* diagnostics.py:43-56                                    */
  double* ess_tail;    /* diagnostics.py:59-73                                    */
  int64_t* lag_bulk;   /* first lag with rho < 0 (n if none) -- diagnostics.py:172-176 */
  int64_t* lag_tail;
  double* mean;        /* compare.py:58-64 / backends_numpy.py:41                 */
  double* std;         /* population std (ddof = 0)                               */
  double* quantiles;   /* [P][nq] numpy-linear -- backends_numpy.py:44            */
  double* order_stats; /* [P][nq][2] the two selected order statistics s[lo], s[lo+1] */
  double* median;      /* pooled statistics.median -- diagnostics.py:97           */
} b200ref_outputs;

/* Version of this ABI (B200REF_VERSION). */
int b200ref_version(void);

/* Message of the last error on this thread ("" if none). */
const char* b200ref_last_error(void);

/* Number of sm_100a kernels this library has launched in this process
 * (monotone counter; bench.py reports the delta over its timed region). */
uint64_t b200ref_kernel_launches(void);

/* Per-kernel device timing for bench.py's roofline block.  While enabled, every
 * kernel launch is bracketed by CUDA events on its launching stream.
 * b200ref_profile_read synchronises the device, accumulates elapsed milliseconds and
 * launch counts per kernel id into ms_out / launches_out (n entries each), optionally
 * resets the accumulators, and returns the number of kernel ids. */
int b200ref_profile(int enable);
int b200ref_profile_read(double* ms_out, uint64_t* launches_out, int n, int reset);
const char* b200ref_profile_name(int id);

/* Workspace needed to process P parameters of C chains x N draws in one call
 * (the library loops over parameter batches internally if given less, down to
 * the size of one batch; see b200ref_workspace_min_bytes). */
size_t b200ref_workspace_bytes(int64_t P, int64_t C, int64_t N);
size_t b200ref_workspace_min_bytes(int64_t C, int64_t N);

/* Pooled average ranks and normal scores of x and of fold(x).
 * Replaces _rank_normalize (diagnostics.py:101-133) and _fold_chains (:93-98).
 * Outputs [P][C][N] in the original positions; any may be NULL.
 * Ranks are exact (half-integers); z follows AS241 in the stdlib's operation order. */
int b200ref_ranks_f64(const double* x_dev, int64_t P, int64_t C, int64_t N,
                      double* avg_rank_dev, double* z_dev,
                      double* avg_rank_folded_dev, double* z_folded_dev,
                      void* ws_dev, size_t ws_bytes, void* stream);

/* Rank-normalised split R-hat, bulk ESS, folded ("tail") ESS, plus pooled
 * mean/std/quantiles of x, for P parameters.
 * Replaces split_rhat / ess_bulk / ess_tail (diagnostics.py:13-73) and the
 * per-column body of NumpyBackend.stats (backends_numpy.py:40-47) /
 * compute_basic_stats (compare.py:58-64).
 * q_host: nq quantile levels in [0,1] (host pointer, nq <= B200REF_MAX_QUANTILES). */
int b200ref_diagnostics_f64(const double* x_dev, int64_t P, int64_t C, int64_t N,
                            const double* q_host, int nq, const b200ref_outputs* out,
                            void* ws_dev, size_t ws_bytes, void* stream);

/* Same, with the truncation rule of the ESS autocorrelation sum selectable.
 *   B200REF_ESS_FIRST_NEGATIVE  the reference's rule (diagnostics.py:172-176): stop at the first rho_t < 0
 *   B200REF_ESS_GEYER           EXTENSION, not in the reference: Geyer's initial monotone sequence on
 *                               the same rho_t (rho_0 := 1, pairs rho_2k + rho_2k+1, stop before the
 *                               first negative pair, pairs capped by their predecessor,
 *                               ESS = m n / max(-1 + 2 sum of pairs, 1 / log10(m n))); defined by the oracle,
 *                               parity unpinned. */
#define B200REF_ESS_FIRST_NEGATIVE 0
#define B200REF_ESS_GEYER 1
int b200ref_diagnostics_ex_f64(const double* x_dev, int64_t P, int64_t C, int64_t N,
                               const double* q_host, int nq, int ess_mode, const b200ref_outputs* out,
                               void* ws_dev, size_t ws_bytes, void* stream);

/* Chain statistics of CALLER-SUPPLIED scores: split R-hat (_split_chains + _rhat, diagnostics.py:76-85, :136-151) and
 * ESS (_ess / _autocorr, :154-193) of zb_dev [P][C][N] (-> rhat_bulk, ess_bulk, lag_bulk) and zf_dev [P][C][N]
 * (-> rhat_tail, ess_tail, lag_tail; null: zb again); out->rhat = max of the two.  No ranking happens here: this is
 * the second half of the fused pass, exposed for inputs whose rank-normalisation was done separately -- the host
 * layer uses it for RAGGED chains, which the reference rank-normalises over all draws and then truncates to the
 * shortest chain (diagnostics.py:140,158).  Workspace: b200ref_workspace_bytes(P, C, N). */
int b200ref_chain_stats_f64(const double* zb_dev, const double* zf_dev, int64_t P, int64_t C, int64_t N, int ess_mode,
                            const b200ref_outputs* out, void* ws_dev, size_t ws_bytes, void* stream);

/* The same diagnostics for draws in the reference's NUMPY layout: `Draws.to_numpy()` (draws.py:20-29) returns
 * [rows][P] with rows ordered (chain, draw) and the PARAMETER index fastest.  x_rows_dev is that matrix
 * ([C*N][P] fp64, device); the library transposes parameter blocks into its workspace (tiled transposing load)
 * and runs the canonical path on them, so results are bit-identical to b200ref_diagnostics_ex_f64 on the
 * transposed tensor.  Workspace: b200ref_rows_workspace_bytes (the canonical workspace of a parameter block
 * plus the block's transposed draws; the library sizes the block to what it is given). */
#define B200REF_LAYOUT_PCN 0  /* x[P][C][N], draw fastest (canonical: one Arrow / Parquet column per parameter) */
#define B200REF_LAYOUT_ROWS 1 /* x[C*N][P], parameter fastest (Draws.to_numpy)                                 */
size_t b200ref_rows_workspace_bytes(int64_t P, int64_t C, int64_t N);
int b200ref_diagnostics_layout_f64(const double* x_dev, int layout, int64_t P, int64_t C, int64_t N,
                                   const double* q_host, int nq, int ess_mode, const b200ref_outputs* out,
                                   void* ws_dev, size_t ws_bytes, void* stream);
/* out[p][r] = in[r][p0 + p] for PB parameter columns of an [M][P] row-major matrix (the transposing load itself). */
int b200ref_transpose_rows_f64(const double* in_dev, int64_t M, int64_t P, int64_t p0, int64_t PB, double* out_dev,
                               void* stream);

/* mean / std(ddof=0) / numpy-linear quantiles over the pooled M draws of each of
 * P parameters; no diagnostics.  Replaces NumpyBackend.stats (backends_numpy.py:17-49).
 * Only out->mean, std, quantiles, order_stats, median are written. */
int b200ref_stats_f64(const double* x_dev, int64_t P, int64_t M,
                      const double* q_host, int nq, const b200ref_outputs* out,
                      void* ws_dev, size_t ws_bytes, void* stream);

/* rel = |act - ref| / max(|ref|, 1e-12);  passed = rel <= tol (NaN fails).
 * Replaces the inner loop of compare_stats (compare.py:38-43) over n values.
 * all_passed_dev (optional, 1 byte) receives the AND of passed[]. */
int b200ref_compare_f64(const double* ref_dev, const double* act_dev, int64_t n, double tol,
                        double* rel_err_dev, uint8_t* passed_dev, uint8_t* all_passed_dev,
                        void* stream);

/* Quality-gate verdicts of convert._checks (convert.py:164-172) over P parameters:
 * verdicts_host[0] = all(ess_bulk > 400), verdicts_host[1] = all(rhat < 1.01). */
int b200ref_checks_f64(const double* rhat_dev, const double* ess_bulk_dev, int64_t P,
                       uint8_t* verdicts_host, void* stream);

/* HOST-buffer form of the full diagnostics + compare pass (the end-to-end call).
 * cand_host [P][C][N] candidate draws and ref_host [P][Mref] reference draws are
 * host buffers (pinned for full speed); the library stages them through the
 * workspace in parameter batches, overlapping H2D copies with kernels, and
 * writes host outputs:
 *   out_host   [P][B200REF_SUMMARY_K] doubles, see B200REF_S_* below
 * Replaces reference.compare + diagnostics_for_model's per-parameter loop
 * (reference.py:76-122).  Returns the overall compare verdict in *passed_host. */
#define B200REF_SUMMARY_K 20
#define B200REF_S_RHAT 0
#define B200REF_S_ESS_BULK 1
#define B200REF_S_ESS_TAIL 2
#define B200REF_S_MEAN 3
#define B200REF_S_STD 4
#define B200REF_S_Q5 5
#define B200REF_S_Q50 6
#define B200REF_S_Q95 7
#define B200REF_S_REF_MEAN 8
#define B200REF_S_REF_STD 9
#define B200REF_S_REF_Q5 10
#define B200REF_S_REF_Q50 11
#define B200REF_S_REF_Q95 12
#define B200REF_S_REL_MEAN 13
#define B200REF_S_REL_STD 14
#define B200REF_S_FLAGS 15 /* bit0 mean passed, bit1 std passed (stored as double) */
/* EXTENSION columns (north-star MCSE / z-score check; the reference has neither -- compare.py is
 * relative-error only -- so parity is pinned against the oracle's definition, and these never
 * change the compare verdict):
 *   mcse_mean   = std / sqrt(ess_bulk)                       (candidate)
 *   ref_se_mean = ref_std / sqrt(Mref)                       (reference draws are thinned: taken as independent)
 *   z_mean      = (mean - ref_mean) / sqrt(mcse_mean^2 + ref_se_mean^2)
 *   z_pass      = 1.0 if |z_mean| <= z_crit else 0.0        (NaN fails) */
#define B200REF_S_MCSE_MEAN 16
#define B200REF_S_REF_SE_MEAN 17
#define B200REF_S_Z_MEAN 18
#define B200REF_S_Z_PASS 19

/* Device-resident form: inputs already in HBM, summary written to device.  With a workspace of
 * b200ref_full_pass_workspace_bytes() the statistics of the reference draws run on an auxiliary
 * stream underneath the candidate pass; with less they run after it (same results). */
size_t b200ref_full_pass_workspace_bytes(int64_t P, int64_t C, int64_t N, int64_t Mref);
int b200ref_full_pass_f64(const double* cand_dev, int64_t P, int64_t C, int64_t N,
                          const double* ref_dev, int64_t Mref, double tol, double z_crit,
                          double* summary_dev /* [P][B200REF_SUMMARY_K] */, uint8_t* passed_host,
                          void* ws_dev, size_t ws_bytes, void* stream);

int b200ref_full_pass_host_f64(const double* cand_host, int64_t P, int64_t C, int64_t N,
                               const double* ref_host, int64_t Mref, double tol, double z_crit,
                               double* summary_host /* [P][B200REF_SUMMARY_K] */, uint8_t* passed_host,
                               void* ws_dev, size_t ws_bytes);

/* Ingestion of the reference's long draws table (rows = (chain, draw), one double column per
 * parameter; store.py:79-95).  Replaces the per-parameter re-bucketing of
 * convert._chains_from_table (convert.py:150-161):
 *   row_order:   *flags_host bit0 = some adjacent row pair is out of (chain, draw) order,
 *                bit1 = an adjacent (chain, draw) pair repeats.  chain / draw are device int64 [rows].
 *   gather_rows: out[p][r] = in[p][perm[r]] for P parameter columns of M rows (device pointers;
 *                perm = the row permutation that sorts by (chain, draw)); not in place. */
int b200ref_row_order_i64(const int64_t* chain_dev, const int64_t* draw_dev, int64_t rows, int* flags_host,
                          void* stream);
int b200ref_gather_rows_f64(const double* in_dev, const int64_t* perm_dev, int64_t P, int64_t M, double* out_dev,
                            void* stream);

/* Batched log-density of the logistic-regression GLM (Stan: y ~ bernoulli_logit(X * beta)):
 *   lp[s] = sum_n ( y[n] * eta - log(1 + exp(eta)) ),  eta = sum_d X[n][d] * beta[s][d]
 * X [N][D], y [N] (0/1), beta [S][D] draws, lp [S]; all fp64 device pointers, row-major.
 * EXTENSION: the reference packages the Stan text and data of such models (e.g.
 * stan_models/wells_data-wells_dist.stan:9-11, provenance.py:539-606) but has no log-density
 * code; parity is pinned against the fp64 oracle only (1e-6 relative).  The contraction runs
 * on the tcgen05 tensor cores with split-bf16 operands and a fused likelihood epilogue. */
size_t b200ref_glm_workspace_bytes(int64_t N, int64_t D, int64_t S);

/* The same fused contraction + likelihood for the GLM families of the packaged models
 * (provenance.py:539-640, 704-734; Stan `~` semantics: terms without parameters are dropped):
 *   B200REF_GLM_BERNOULLI_LOGIT  sum_n y eta - log(1 + exp(eta))
 *   B200REF_GLM_BINOMIAL_LOGIT   sum_n y eta - trials[n] log(1 + exp(eta))          (trials_dev [N])
 *   B200REF_GLM_POISSON_LOG      sum_n y eta - exp(eta)
 *   B200REF_GLM_NORMAL           -N log sigma[s] - sum_n (y - eta)^2 / (2 sigma[s]^2) (sigma_dev [S])
 * trials_dev / sigma_dev may be null for the families that do not use them. */
#define B200REF_GLM_BERNOULLI_LOGIT 0
#define B200REF_GLM_BINOMIAL_LOGIT 1
#define B200REF_GLM_POISSON_LOG 2
#define B200REF_GLM_NORMAL 3
int b200ref_glm_lp_f64(int family, const double* X_dev, const double* y_dev, const double* trials_dev,
                       const double* beta_dev, const double* sigma_dev, int64_t N, int64_t D, int64_t S,
                       double* lp_dev, void* ws_dev, size_t ws_bytes, void* stream);
/* Same, with flags.  B200REF_GLM_REUSE_DESIGN: the caller states that X / y / trials (pointers, contents, N, D) are
 * those of the previous call made with this workspace; the library then reuses its staged copies of the design
 * (the bf16 hi/lo split of X is 0.2 ms and 340 MB of traffic at config 3) if the workspace header agrees, and
 * stages them afresh otherwise. */
#define B200REF_GLM_REUSE_DESIGN 1
int b200ref_glm_lp_ex_f64(int family, const double* X_dev, const double* y_dev, const double* trials_dev,
                          const double* beta_dev, const double* sigma_dev, int64_t N, int64_t D, int64_t S, int flags,
                          double* lp_dev, void* ws_dev, size_t ws_bytes, void* stream);
int b200ref_glm_logistic_lp_f64(const double* X_dev, const double* y_dev, const double* beta_dev,
                                int64_t N, int64_t D, int64_t S, double* lp_dev,
                                void* ws_dev, size_t ws_bytes, void* stream);

/* Synthetic workload of SURVEY.md 8(d) (bench/test utility, not on the hot path):
 * x[p][c][i] = mu_p + sigma_p * AR1(phi_p), counter-based on (seed, p0 + p, c, i)
 * so values do not depend on tiling or GPU count.  kind 0 = candidate [P][C][N],
 * kind 1 = reference posterior [P][C][N] iid (every 89th param mean-shifted). */
int b200ref_synth_fill_f64(double* x_dev, int64_t p0, int64_t P, int64_t C, int64_t N,
                           uint64_t seed, int kind, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200REF_H */
