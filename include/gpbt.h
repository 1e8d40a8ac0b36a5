/*
 * gpbt.h -- C ABI of libgpbt_b200.so: the B200 (sm_100a) implementation of GPBayesTools-HIC's
 * batched "PCA-GP emulator -> Gaussian log-likelihood" hot path.
 *
 * The reference is pure Python and has no FFI; the operator interfaces this ABI sits behind are
 * the Python call signatures (paths are under /root/reference):
 *
 *   Emulator.predict(X, return_cov, extra_std)     src/emulator.py:465-605
 *   Chain._predict(X, extra_std)                   src/mcmc.py:153-166
 *   mvn_loglike(y, cov)                            src/mcmc.py:23-65
 *   Chain.log_likelihood(X, ..., finite)           src/mcmc.py:188-222
 *   Chain.log_posterior(X, ...)                    src/mcmc.py:261-299
 *
 * Conventions
 *   - all floating point data is IEEE binary64, row-major, contiguous unless a leading dimension
 *     is given;  N = walkers (rows of X), p = parameters, n = design points, q = emulated PCs,
 *     m = observables of one emulator, M = observables of the whole chain, Q = PCs of the chain.
 *   - "_dev" arguments are device pointers on the CUDA device that was current when the handle
 *     was created; "_host" arguments are host pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device entry
 *     points only enqueue work; they never synchronise.
 *   - every function returns 0 on success, a positive cudaError_t value on a CUDA failure, or a
 *     negative GPBT_E* code.  gpbt_last_error() returns a thread-local message.
 *   - there is no CPU implementation behind any entry point.
 */
#ifndef GPBT_H
#define GPBT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpbt_emulator* gpbt_emulator_t; /* device-resident trained state of one emulator */
typedef struct gpbt_chain* gpbt_chain_t;       /* emulators + experimental data + workspaces    */

enum { GPBT_KERNEL_RBF = 0, GPBT_KERNEL_MATERN32 = 1, GPBT_KERNEL_PCGP = 2 /* gpbt_emulator_create_pcgp only */ };
enum { GPBT_FLAG_NO_PCA = 1, GPBT_FLAG_EXP_DIAG = 2 };
enum { GPBT_PATH_AUTO = 0, GPBT_PATH_DENSE = 1, GPBT_PATH_LOWRANK = 2, GPBT_PATH_DIAG = 3 };
enum {
  GPBT_EINVAL = -1,   /* bad argument                                              */
  GPBT_ESHAPE = -2,   /* shape outside what the kernels support (message says why) */
  GPBT_ENOTAPPLICABLE = -3 /* requested path cannot represent this chain           */
};

const char* gpbt_last_error(void);
int gpbt_version(void);

/* ---- emulator state --------------------------------------------------------------------- *
 * Replaces the sklearn objects a trained reference Emulator carries (src/emulator.py:309-363):
 *   Xtr   [n,p]    gp.X_train_ (shared by all GPs, copy_X_train=False, :312)
 *   ell   [q,p]    gp.kernel_.k1.k2.length_scale      c [q]  gp.kernel_.k1.k1.constant_value
 *   sn    [q]      gp.kernel_.k2.noise_level          alpha [q,n]  gp.alpha_
 *   Linv  [q,n,n]  inverse of gp.L_ (lower triangular; the caller inverts L_ once in FP64)
 *   A     [q,m]    _trans_matrix[:npc]  (PCA mode; ignored with GPBT_FLAG_NO_PCA)
 *   mu    [m]      scaler.mean_         scale [m]  scaler.scale_ (used with GPBT_FLAG_NO_PCA)
 *   Ctrunc[m,m]    _cov_trunc           (PCA mode; NULL with GPBT_FLAG_NO_PCA)
 * All pointers are HOST pointers; the data is copied to the current device.                 */
int gpbt_emulator_create(gpbt_emulator_t* out, int p, int n, int q, int m, int kernel_kind,
                         int flags, const double* Xtr_host, const double* ell_host,
                         const double* c_host, const double* sn_host, const double* alpha_host,
                         const double* Linv_host, const double* A_host, const double* mu_host,
                         const double* scale_host, const double* Ctrunc_host);
int gpbt_emulator_destroy(gpbt_emulator_t emu);

/* surmise PCGP / PCSK emulator state, the object behind EmulatorBAND.predict
 * (src/emulator_BAND.py:270-292 fit, :386-478 predict = self.emu.predict(x, theta).mean() / .covx()).
 * The arithmetic is inside surmise 0.2.1 (requirements.txt), which is NOT part of the reference tree
 * and not installed in the build image: this entry point follows the published form of
 * emulationmethods/PCGP.py `predict` and its parity against surmise itself is UNPINNED.
 *   theta [n,p]   fitinfo['theta']                  ell [q,p]  exp(hypcov[:-1]) of each PC's GP
 *   amp   [q]     (1 - nug) / (1 + exp(hypcov[-1])) off [q]    (1 - nug) exp(hypcov[-1]) / (1 + exp(hypcov[-1]))
 *   sig2  [q]     emulist[k]['sig2']                pw  [q,n]  emulist[k]['pw']
 *   VhT   [q,n,n] transpose of emulist[k]['Vh']     A   [q,m]  (pct * scale[:,None]).T
 *   offset [m]    fitinfo['offset']                 extra_cov [m,m] constant added to every covariance
 *                                                   (zeros for covx(); diag(extravar) for the variant)
 * Per walker:  r = amp * prod_d(1 + s_d) exp(-sum_d s_d) + off,  s_d = |x_d - theta_d| / ell_d;
 *   z_mean = r . pw,  z_var = sig2 |1 - |r Vh|^2|;  mean = z_mean A + offset,
 *   cov = A^T diag(z_var) A + extra_cov.  GPBT_FLAG_EXP_DIAG as for gpbt_emulator_create.
 * extra_std passed to gpbt_pc_predict is ignored, as EmulatorBAND.predict ignores it.
 * All pointers are HOST pointers.                                                             */
int gpbt_emulator_create_pcgp(gpbt_emulator_t* out, int p, int n, int q, int m, int flags,
                              const double* theta_host, const double* ell_host, const double* amp_host,
                              const double* off_host, const double* sig2_host, const double* pw_host,
                              const double* VhT_host, const double* A_host, const double* offset_host,
                              const double* extra_cov_host);

/* Optional "parameterTrafoPCA" pre-transform in front of kernel (a) (src/emulator.py:492-551,
 * src/emulator_BAND.py:393-452): walkers arrive with p_in model parameters; the columns listed in
 * `keep` are copied through, and each of n_groups (<= 3) parameter groups is replaced by ncomp[g]
 * principal components of the npts[g]-point curve it parametrises on linspace(grid_lo, grid_hi):
 *   kind 0  zeta/s(T)       4 columns  (src/emulator.py:100-106)
 *   kind 1  eta/s(mu_B)     3 columns  (src/emulator.py:109-115)
 *   kind 2  y_loss(y_init)  3 columns  (src/emulator.py:118-124)
 * idx is [n_groups][4]; Wt[g] is [ncomp[g]][npts[g]] and b[g] is [ncomp[g]], the StandardScaler and
 * PCA of the reference folded into  PC_j = sum_t f_t Wt[j][t] + b_j.  n_keep + sum(ncomp) must
 * equal the p the emulator was created with.  All pointers are HOST pointers.                 */
int gpbt_emulator_set_param_trafo(gpbt_emulator_t emu, int p_in, const int* keep, int n_keep,
                                  int n_groups, const int* kinds, const int* idx, const int* ncomp,
                                  const int* npts, const double* grid_lo, const double* grid_hi,
                                  const double* const* Wt, const double* const* b);
/* number of columns of X this emulator expects (p_in with a pre-transform, else p)              */
int gpbt_emulator_input_dim(gpbt_emulator_t emu);

/* ---- kernel (a): per-(walker, PC) cross-kernel + GP mean + predictive variance ----------- *
 * Replaces [gp.predict(X, return_cov=True) for gp in self.gps] + diagonal extraction + the
 * extra_std**2 term (src/emulator.py:553, 573-579; sklearn _gpr.py:446-466).
 *   z_mean[i*ldz + j], z_var[i*ldz + j]  for walker i, PC j;   extra_std_dev may be NULL.     */
int gpbt_pc_predict(gpbt_emulator_t emu, const double* X_dev, const double* extra_std_dev,
                    double* z_mean_dev, double* z_var_dev, int64_t ldz, int64_t N, void* stream);

/* ---- kernel (b): PCA back-transform to observable space ---------------------------------- *
 * Replaces _inverse_transform / scaler.inverse_transform, exp(), np.dot(gp_var,_var_trans) +
 * _cov_trunc and the exp_and_cov_diagonal rewrite (src/emulator.py:558-601).
 *   mean_dev[i*ld_mean + col_off + o]                         (always written)
 *   cov_dev [i*ld_cov*ld_cov + (col_off+o)*ld_cov + col_off + o']   (skipped when NULL)
 * With ld_cov > m the rows col_off..col_off+m-1 of each walker's matrix are written across ALL
 * ld_cov columns (zeros outside the diagonal block), which is how Chain._predict's block-diagonal
 * covariance (src/mcmc.py:153-166) is assembled without a separate memset.                   */
int gpbt_backtransform(gpbt_emulator_t emu, const double* z_mean_dev, const double* z_var_dev,
                       int64_t ldz, double* mean_dev, int64_t ld_mean, double* cov_dev,
                       int64_t ld_cov, int64_t col_off, int64_t N, void* stream);

/* Mean and only the DIAGONAL of the observable-space covariance (var_dev[i*ld + col_off + o]); the
 * form posterior-predictive / sensitivity sweeps need (examples/SensitivityAnalysis.ipynb cell 4,
 * BASELINE config 5), where N * m^2 covariances could not be stored, let alone returned.        */
int gpbt_backtransform_diag(gpbt_emulator_t emu, const double* z_mean_dev, const double* z_var_dev,
                            int64_t ldz, double* mean_dev, double* var_dev, int64_t ld,
                            int64_t col_off, int64_t N, void* stream);

/* ---- kernel (c): batched Cholesky log-likelihood ------------------------------------------ *
 * Replaces list(map(mvn_loglike, dY, cov)) (src/mcmc.py:23-65, 293) including dY = mean - y_exp
 * and cov + expdata_cov (src/mcmc.py:288-290):
 *   lp[i] = -1/2 y_i^T C_i^-1 y_i - sum(log(diag(chol(C_i)))),   y_i = mean_i - y_exp (y_exp may
 *   be NULL), C_i = cov_i + cov_add (cov_add may be NULL).
 * cov_dev [N,m,m] MAY BE OVERWRITTEN (the in-place kernels leave the Cholesky factor in its lower
 * triangle, as dpotrf does to its own copy; batches of 256+ matrices go through the fused kernels, which
 * keep their packed factor in a work buffer and leave cov_dev untouched).  Walkers whose matrix is not positive definite get lp = notpd_value and
 * increment *n_notpd_dev (may be NULL); the reference's own check is broken (both branches test
 * info < 0, src/mcmc.py:44-54) and would return garbage there.                               */
int gpbt_mvn_loglike(const double* mean_dev, const double* y_exp_dev, double* cov_dev,
                     const double* cov_add_dev, double* lp_dev, int* n_notpd_dev,
                     double notpd_value, int64_t N, int m, void* stream);

/* ---- chain: everything between X[N,p] and lp[N] ------------------------------------------ *
 * lo/hi [p]: Chain.min/max; y_exp [M]: Chain.expdata; cov_exp [M,M]: Chain.expdata_cov
 * (src/mcmc.py:104-142).  The optional low-rank factors describe the exact identity
 *     C_w = F + U^T diag(v_w) U,  F = blockdiag(Ctrunc_e) + cov_exp  (walker independent)
 * through  L_F^-1 U^T = Qb R :  R [Q,Q] upper triangular, c0 = Qb^T L_F^-1 (mu - y_exp) [Q],
 * s_perp = |(I - Qb Qb^T) L_F^-1 (mu - y_exp)|^2, logdetF_half = sum(log(diag(L_F))).
 * Pass R_host = NULL when the chain has a no-PCA / exp-diag emulator (dense path only).
 * All pointers are HOST pointers.                                                             */
int gpbt_chain_create(gpbt_chain_t* out, const gpbt_emulator_t* emus, int n_emu, int p,
                      const double* lo_host, const double* hi_host, const double* y_exp_host,
                      const double* cov_exp_host, const double* R_host, const double* c0_host,
                      double s_perp, double logdetF_half);
int gpbt_chain_destroy(gpbt_chain_t chain);

/* Chain._predict on device buffers: mean [N,M], cov [N,M,M] (cov may be NULL).               */
int gpbt_chain_predict(gpbt_chain_t chain, const double* X_dev, double extra_std_scale,
                       double* mean_dev, double* cov_dev, int64_t N, void* stream);

/* Chain.log_posterior / log_likelihood: bounds mask (strict, src/mcmc.py:275-276), emulators,
 * likelihood and the constant 2*log(1e-16) (src/mcmc.py:296-297).  Rows outside the box get
 * oob_value (-inf, or -1e300 for finite=True).  path: GPBT_PATH_AUTO picks LOWRANK when the
 * chain was created with R; DIAG (element-wise) when every emulator is no-PCA / exp-diag and
 * cov_exp is diagonal; else DENSE ((a) -> (b) -> (c) with the covariance in HBM).  Asking for a
 * path the chain cannot take returns GPBT_ENOTAPPLICABLE.
 * n_notpd_dev (may be NULL) counts non-positive-definite walkers.                            */
int gpbt_log_posterior(gpbt_chain_t chain, const double* X_dev, double oob_value, double* lp_dev,
                       int* n_notpd_dev, int64_t N, int path, void* stream);

/* gpbt_log_posterior with the all-gather fused into the last kernel: every lp[w] is additionally
 * stored to peers[r][peer_off + w], r < n_peers (<= 16), where peers[] are device pointers valid in
 * this process -- buffers of the other GPUs mapped over NVLink (CUDA IPC / symmetric memory), or
 * local ones.  peers_host is a HOST array of those pointers.  The caller orders the ranks (one
 * barrier after the call) before anyone reads the gathered vector.  Replaces the NCCL all-gather of
 * the per-walker log-posteriors, the only collective of the multi-GPU path.                       */
int gpbt_log_posterior_scatter(gpbt_chain_t chain, const double* X_dev, double oob_value, double* lp_dev,
                               double* const* peers_host, int n_peers, int64_t peer_off,
                               int* n_notpd_dev, int64_t N, int path, void* stream);

/* Same with HOST buffers: H2D of X, kernels, D2H of lp, one synchronisation.
 * This is the call behind Chain.log_posterior(X: np.ndarray) -> np.ndarray.                  */
int gpbt_log_posterior_host(gpbt_chain_t chain, const double* X_host, double oob_value,
                            double* lp_host, int* n_notpd_host, int64_t N, int path);

/* ---- device-resident ensemble sampler --------------------------------------------------- *
 * Replaces emcee.EnsembleSampler(nwalkers, ndim, chain.log_posterior, pool=chain) with its
 * default stretch move, as driven by LoggingEnsembleSampler.run_mcmc / Chain.run_mcmc
 * (src/mcmc.py:68-92, 372-412).  Walkers, log-posteriors and the history of every step stay on
 * the device; a step is a CUDA graph around the log-posterior path, so nothing crosses PCIe
 * until the chain is read back.  emcee (>= 3.1.4, requirements.txt) is not part of the reference
 * tree; the move follows the published algorithm (Goodman & Weare 2010; red/blue split).
 * All calls are synchronous and run on the chain's own stream.                               */
typedef struct gpbt_ensemble* gpbt_ensemble_t;

/* a: stretch scale (emcee default 2.0); randomize_split: shuffle the two walker sets every
 * step (emcee default) or keep even / odd walkers apart; seed: Philox key                    */
int gpbt_ensemble_create(gpbt_ensemble_t* out, gpbt_chain_t chain, int n_walkers, double a,
                         int randomize_split, uint64_t seed);
int gpbt_ensemble_destroy(gpbt_ensemble_t ens);

/* walker positions X_host [n_walkers, p]; lp_host [n_walkers] or NULL (= evaluate them)      */
int gpbt_ensemble_set_state(gpbt_ensemble_t ens, const double* X_host, const double* lp_host);
int gpbt_ensemble_get_state(gpbt_ensemble_t ens, double* X_host, double* lp_host);

/* Advance n_steps steps, appending every step to the device history.
 * Test hooks (all NULL in production = device Philox streams): u_host [n_steps, 2, n_half, 2]
 * uniforms (stretch draw, accept draw) and partner_host [n_steps, 2, n_half] indices into the
 * complementary set, with n_half = ceil(n_walkers / 2); perm_host [n_steps, n_walkers] the split
 * permutation of each step (first n_half entries = first set).
 * use_graph: 0 = enqueue kernel by kernel, 1 = replay a captured CUDA graph.                 */
int gpbt_ensemble_run(gpbt_ensemble_t ens, int64_t n_steps, const double* u_host,
                      const int32_t* partner_host, const int32_t* perm_host, int use_graph);

/* The same step in pieces, for callers that evaluate the proposals themselves -- e.g. one process
 * per GPU with replicated walkers (same seed on every rank -> identical proposals), each rank
 * evaluating a slice and the log-posteriors all-gathered (gpbt_log_posterior_scatter):
 *   gpbt_ensemble_prepare(n_steps)            room in the history for n_steps more steps
 *   gpbt_ensemble_begin_half(half)            (half 0: new split of the walkers) + stretch proposals
 *   gpbt_ensemble_copy_proposals(half, first, n, dst_dev)   proposal rows [first, first+n) -> dst_dev[n,p]
 *                                             (rows past the active set repeat its last row)
 *   gpbt_ensemble_end_half(half, lp_new_dev)  accept with lp_new_dev[i] = log-posterior of proposal i;
 *                                             half 1 also appends the step to the history
 * These only enqueue on `stream`; the caller orders them against its own work there.          */
int gpbt_ensemble_prepare(gpbt_ensemble_t ens, int64_t n_steps);
int gpbt_ensemble_begin_half(gpbt_ensemble_t ens, int half, void* stream);
int gpbt_ensemble_copy_proposals(gpbt_ensemble_t ens, int half, int64_t first, int64_t n, double* dst_dev,
                                 void* stream);
int gpbt_ensemble_end_half(gpbt_ensemble_t ens, int half, const double* lp_new_dev, void* stream);

/* steps in the history since the last reset                                                  */
int64_t gpbt_ensemble_steps(gpbt_ensemble_t ens);
/* make room for n_steps more steps of history in one allocation (optional; run grows it)     */
int gpbt_ensemble_reserve(gpbt_ensemble_t ens, int64_t n_steps);
/* history rows [first, first + n): chain_host [n, n_walkers, p], lp_host [n, n_walkers]
 * (either may be NULL); accepted_host [n_walkers] accepted proposals per walker since the
 * last reset; n_notpd_host: non-positive-definite covariances met since creation            */
int gpbt_ensemble_read(gpbt_ensemble_t ens, int64_t first, int64_t n, double* chain_host,
                       double* lp_host, int64_t* accepted_host, int64_t* n_notpd_host);
/* forget the history and the acceptance counts; the walkers stay where they are              */
int gpbt_ensemble_reset(gpbt_ensemble_t ens);

/* HOST helper (no GPU work): one sweep of Chain.tempexchange (src/mcmc.py:679-693) over n_picks
 * caller-drawn picks rt in [1, n) and log-uniform draws; swaps order[rt-1], order[rt] where
 * (lp[order[rt]] - lp[order[rt-1]]) * (1/temps[rt-1] - 1/temps[rt]) > log_u[i].  order is updated
 * in place.  The swaps depend on each other, so the loop is sequential; it is here because in
 * Python it dominates a PTLMC iteration at thousands of chains.                                 */
int gpbt_host_temp_exchange(const double* lp_host, const double* temps_host, int64_t n,
                            const int64_t* picks_host, const double* log_u_host, int64_t n_picks,
                            int64_t* order_host);

/* ---- device-resident PTLMC sampling loop ------------------------------------------------------- *
 * Replaces the iteration loop of Chain.samplerPTLMC (src/mcmc.py:623-671, the branch without
 * gradients, which is the one Chain.log_posterior drives) and Chain.tempexchange (:679-693): Gaussian
 * proposal theta + sqrt(2) rho_T (xi C^1/2) with rho_T = stride(tau) T^(1/3), tempered Metropolis
 * accept, five sweeps of random adjacent-temperature swaps, step-size tuning every tenth of the
 * first n_tune iterations, record of the T = 1 chains afterwards.  All n_chains chains stay on the
 * device; an iteration is proposal kernel -> log-posterior path -> one single-CTA kernel.  Random
 * numbers are counter-based (Philox, keyed by seed): reproducible, but not NumPy's global stream
 * (the host driver gpbt_b200/ptlmc.py follows that one draw for draw).  The start-up stage of the
 * reference's sampler (:560-621) stays with the caller.  Synchronous, on the chain's own stream.   */
typedef struct gpbt_ptlmc* gpbt_ptlmc_t;

/* temps_host [n_chains]: the ladder, hottest first, the last n_chains - n_hot entries at T = 1;
 * root_host [p, p]: C^1/2 of the proposal; goal: target acceptance (0.25 in the reference)        */
int gpbt_ptlmc_create(gpbt_ptlmc_t* out, gpbt_chain_t chain, int n_chains, int n_hot, const double* temps_host,
                      const double* root_host, double goal, uint64_t seed);
int gpbt_ptlmc_destroy(gpbt_ptlmc_t s);
/* starting points theta_host [n_chains, p] (their log-posteriors are evaluated here), tau: the
 * step-size parameter (reference: -1); restarts the iteration count                               */
int gpbt_ptlmc_set_state(gpbt_ptlmc_t s, const double* theta_host, double tau);
/* the next n_steps iterations of a run of n_tune + n_keep                                         */
int gpbt_ptlmc_run(gpbt_ptlmc_t s, int64_t n_tune, int64_t n_keep, int64_t n_steps);
/* saved_host [n_chains - n_hot, n_keep, p] (the reference's sampler_info['theta']), theta_host
 * [n_chains, p] / lp_host [n_chains] the current state, info_host[4] = {tau, stride, accepted
 * proposals of the T = 1 chains after tuning, non-positive-definite covariances seen}; any may be
 * NULL                                                                                            */
int gpbt_ptlmc_read(gpbt_ptlmc_t s, double* saved_host, double* theta_host, double* lp_host, double* info_host);

/* ---- single-process multi-GPU fan-out of the host boundary call -------------------------------- *
 * The reference's samplers are single-process and pass ONE host array to Chain.log_posterior /
 * log_likelihood (src/mcmc.py:188-222, 261-299; pocoMC hands over all active particles, :798-804).
 * A fan-out owns replicas of one chain on distinct devices (created by the caller with
 * gpbt_set_device + gpbt_emulator_create + gpbt_chain_create per device) and one host worker thread per
 * replica.  gpbt_fanout_log_posterior_host splits X_host [N,p] into contiguous row blocks, each worker
 * stages its block through pinned memory to its GPU, runs gpbt_log_posterior there and copies its lp
 * block into lp_host; the call returns when all blocks are back.  No collective is involved.
 * max_devices: 0 = automatic (one device per `fanout_min_rows` rows, default 256, at most all),
 * else the number of replicas to use.  *devices_used (may be NULL) reports how many took part.
 * The chains stay owned by the caller and must outlive the fan-out.                             */
typedef struct gpbt_fanout* gpbt_fanout_t;
int gpbt_device_count(void);
int gpbt_set_device(int device);
int gpbt_get_device(void);
int gpbt_fanout_create(gpbt_fanout_t* out, const gpbt_chain_t* chains, int n_chains);
int gpbt_fanout_destroy(gpbt_fanout_t fanout);
int gpbt_fanout_size(gpbt_fanout_t fanout);
int gpbt_fanout_log_posterior_host(gpbt_fanout_t fanout, const double* X_host, double oob_value,
                                   double* lp_host, int* n_notpd_host, int64_t N, int path,
                                   int max_devices, int* devices_used);

/* Tuning overrides (tests, tuning tools).  The library reads GPBT_PC_TILE, GPBT_CHOL,
 * GPBT_LOWRANK_GENERIC, GPBT_NO_ZEROCOPY, GPBT_ENSEMBLE_SPLIT_KERNELS, GPBT_FANOUT_MIN_ROWS,
 * GPBT_CHOL_BATCH, GPBT_CHOL_STREAMS and GPBT_CHOL_PIPE from the environment once, at load; this call
 * changes one value afterwards.
 * keys: "pc_tile" (8|16|32), "chol" (warp|batch|staged|cta|fused), "lowrank_generic", "no_zerocopy",
 * "ensemble_split_kernels", "fanout_min_rows"; the fused Cholesky's schedule: "chol_batch" (walkers per
 * sub-batch), "chol_streams", "chol_pipe" (kernel (a) pipelined over that many sub-batches), "chol_lag"
 * (sub-batches start one step apart), "chol_prio", "cf_debug"; value NULL or "" restores the default. */
int gpbt_set_option(const char* key, const char* value);

/* bytes of device workspace the chain currently holds (grows with the largest N seen)        */
int64_t gpbt_chain_workspace_bytes(gpbt_chain_t chain);

/* test hook: y[i] = the kernels' internal exp(x[i]) for x <= 0 (accuracy is pinned by a test)  */
int gpbt_debug_exp_neg(const double* x_dev, double* y_dev, int64_t n, void* stream);

/* tuning hook: with option "cf_debug" the fused Cholesky records clock64 stamps of its phases for the
 * first 32 walkers, [16 launches][32 walkers][8 tiles][8 stamps] int64; this copies them out        */
int gpbt_debug_timing_read(void* dst_host, int64_t bytes);
/* tuning / debugging hook: work buffers of the fused Cholesky of walker w of the last dense call
 * (what: 0 factor panels, 1 t, 2 Dinv, 3 raw diagonal block, 4 z_var, 5 mean, 6 log det, 7 |t|^2)   */
int gpbt_debug_fused_read(gpbt_chain_t chain, int what, int64_t w, double* dst_host, int64_t count);

/* number of kernel launches issued by this library since load (bench.py's gpu_launches)      */
int64_t gpbt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* GPBT_H */
