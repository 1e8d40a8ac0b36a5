// Device-resident PTLMC sampling loop (included at the end of gpbt_api.cu): gpbt_ptlmc_* of include/gpbt.h.
// Kernels and algorithm: ptlmc.cuh.  The start-up stage of the reference's sampler (ranking of the candidate
// points, L-BFGS-B polish, src/mcmc.py:560-621) stays on the host -- it is a few hundred N = 1 calls made once;
// this object takes over at the loop of src/mcmc.py:623-671.
#include "ptlmc.cuh"

struct gpbt_ptlmc {
  gpbt_chain_t ch = nullptr;
  int n = 0, p = 0, n_hot = 0;
  double goal = 0.25;
  uint64_t seed = 0;
  PtlmcCtl* ctl = nullptr;
  double *temps = nullptr, *cbrt_t = nullptr, *gap = nullptr, *root = nullptr;
  double *theta[2] = {nullptr, nullptr}, *lp[2] = {nullptr, nullptr}, *prop = nullptr, *lp_prop = nullptr;
  double* saved = nullptr;
  int* sw_slot = nullptr;
  double *sw_gap = nullptr, *sw_logu = nullptr;
  int64_t n_keep = 0, n_tune = 0, k = 0;      // k: iterations done in the current run
  int cur = 0;                                 // which of theta[] / lp[] holds the state
  int* notpd = nullptr;
  bool has_state = false;
};

extern "C" int gpbt_ptlmc_destroy(gpbt_ptlmc_t s) {
  if (!s) return 0;
  cudaSetDevice(s->ch ? s->ch->device : 0);
  if (s->ch && s->ch->stream) cudaStreamSynchronize(s->ch->stream);
  void* ptrs[] = {s->ctl, s->temps, s->cbrt_t, s->gap, s->root, s->theta[0], s->theta[1], s->lp[0], s->lp[1],
                  s->prop, s->lp_prop, s->saved, s->notpd, s->sw_slot, s->sw_gap, s->sw_logu};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  delete s;
  return 0;
}

extern "C" int gpbt_ptlmc_create(gpbt_ptlmc_t* out, gpbt_chain_t ch, int n_chains, int n_hot, const double* temps_host,
                                 const double* root_host, double goal, uint64_t seed) {
  if (!out || !ch || n_chains < 1 || n_hot < 0 || n_hot >= n_chains || !temps_host || !root_host)
    return fail(GPBT_EINVAL, "gpbt_ptlmc_create: bad argument");
  for (int i = 0; i < n_chains; i++)
    if (!(temps_host[i] > 0.0)) return fail(GPBT_EINVAL, "gpbt_ptlmc_create: temperatures must be positive");
  CU(cudaSetDevice(ch->device));
  if (ptlmc_step_smem_bytes(n_chains) > (size_t)max_optin_smem())
    return fail(GPBT_ESHAPE, "gpbt_ptlmc_create: %d chains do not fit in one CTA's shared memory", n_chains);
  gpbt_ptlmc* s = new gpbt_ptlmc();
  s->ch = ch; s->n = n_chains; s->p = ch->p; s->n_hot = n_hot; s->goal = goal; s->seed = seed;
  const int n = s->n, p = s->p;
  std::vector<double> cb(n), gap(n, 0.0);
  for (int i = 0; i < n; i++) {
    cb[i] = pow(temps_host[i], 1.0 / 3.0);      // (temps ** (1 / 3), src/mcmc.py:620)
    if (i > 0) gap[i] = 1.0 / temps_host[i - 1] - 1.0 / temps_host[i];
  }
  auto up = [&](double** dst, const double* src, size_t cnt) -> int {
    CU(cudaMalloc(dst, cnt * sizeof(double)));
    if (src) CU(cudaMemcpy(*dst, src, cnt * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
  };
  int rc = 0;
  rc = rc ? rc : up(&s->temps, temps_host, n);
  rc = rc ? rc : up(&s->cbrt_t, cb.data(), n);
  rc = rc ? rc : up(&s->gap, gap.data(), n);
  rc = rc ? rc : up(&s->root, root_host, (size_t)p * p);
  for (int b = 0; b < 2 && !rc; b++) {
    rc = up(&s->theta[b], nullptr, (size_t)n * p);
    rc = rc ? rc : up(&s->lp[b], nullptr, n);
  }
  rc = rc ? rc : up(&s->prop, nullptr, (size_t)n * p);
  rc = rc ? rc : up(&s->lp_prop, nullptr, n);
  rc = rc ? rc : up(&s->sw_gap, nullptr, n);
  rc = rc ? rc : up(&s->sw_logu, nullptr, n);
  auto raw = [&](void** dst, size_t bytes) -> int {
    CU(cudaMalloc(dst, bytes));
    return 0;
  };
  rc = rc ? rc : raw(reinterpret_cast<void**>(&s->ctl), sizeof(PtlmcCtl));
  rc = rc ? rc : raw(reinterpret_cast<void**>(&s->notpd), sizeof(int));
  rc = rc ? rc : raw(reinterpret_cast<void**>(&s->sw_slot), (size_t)n * sizeof(int));
  if (rc) {
    gpbt_ptlmc_destroy(s);
    return rc;
  }
  *out = s;
  return 0;
}

extern "C" int gpbt_ptlmc_set_state(gpbt_ptlmc_t s, const double* theta_host, double tau) {
  if (!s || !theta_host) return fail(GPBT_EINVAL, "gpbt_ptlmc_set_state: bad argument");
  CU(cudaSetDevice(s->ch->device));
  cudaStream_t st = s->ch->stream;
  s->cur = 0;
  s->k = 0;
  CU(cudaMemcpyAsync(s->theta[0], theta_host, (size_t)s->n * s->p * sizeof(double), cudaMemcpyHostToDevice, st));
  PtlmcCtl c;
  c.tau = tau; c.hits = 0.0; c.stride = ptlmc_stride(tau); c.accepted = 0;
  CU(cudaMemcpyAsync(s->ctl, &c, sizeof c, cudaMemcpyHostToDevice, st));
  CU(cudaMemsetAsync(s->notpd, 0, sizeof(int), st));
  if (int r = log_posterior_impl(s->ch, s->theta[0], -INFINITY, s->lp[0], s->notpd, s->n, GPBT_PATH_AUTO, st, nullptr, 0,
                                 0, /*zero_counter=*/false))
    return r;
  CU(cudaStreamSynchronize(st));
  s->has_state = true;
  return 0;
}

// Iterations k = done .. done + n_steps - 1 of a run of n_tune + n_keep iterations (the first call of a run
// sizes the record of the T = 1 chains).  Everything is enqueued on the chain's stream; the call returns after
// the last iteration has finished.
extern "C" int gpbt_ptlmc_run(gpbt_ptlmc_t s, int64_t n_tune, int64_t n_keep, int64_t n_steps) {
  if (!s || n_tune < 0 || n_keep < 0 || n_steps < 0) return fail(GPBT_EINVAL, "gpbt_ptlmc_run: bad argument");
  if (!s->has_state) return fail(GPBT_EINVAL, "gpbt_ptlmc_run: no state (gpbt_ptlmc_set_state)");
  CU(cudaSetDevice(s->ch->device));
  cudaStream_t st = s->ch->stream;
  const int n = s->n, p = s->p, n_cold = n - s->n_hot;
  if (s->k == 0) {
    if (s->saved) cudaFree(s->saved);
    s->saved = nullptr;
    s->n_tune = n_tune;
    s->n_keep = n_keep;
    if (n_keep > 0) CU(cudaMalloc(&s->saved, (size_t)n_cold * n_keep * p * sizeof(double)));
  } else if (n_tune != s->n_tune || n_keep != s->n_keep) {
    return fail(GPBT_EINVAL, "gpbt_ptlmc_run: the run in progress has n_tune = %lld, n_keep = %lld", (long long)s->n_tune,
                (long long)s->n_keep);
  }
  if (s->k + n_steps > n_tune + n_keep) return fail(GPBT_EINVAL, "gpbt_ptlmc_run: more steps than the run has");
  const size_t smem_step = ptlmc_step_smem_bytes(n);
  if (int r = ensure_dynamic_smem<ptlmc_step_kernel>(smem_step)) return r;
  const int warps = 4, p2 = (p + 1) / 2;
  for (int64_t it = 0; it < n_steps; it++, s->k++) {
    PtlmcParams prm;
    prm.ctl = s->ctl; prm.seed = s->seed; prm.n = n; prm.p = p; prm.n_hot = s->n_hot;
    prm.temps = s->temps; prm.cbrt_t = s->cbrt_t; prm.gap = s->gap; prm.root = s->root;
    prm.theta_in = s->theta[s->cur]; prm.lp_in = s->lp[s->cur]; prm.prop = s->prop; prm.lp_prop = s->lp_prop;
    prm.theta_out = s->theta[s->cur ^ 1]; prm.lp_out = s->lp[s->cur ^ 1]; prm.saved = s->saved;
    prm.sw_slot = s->sw_slot; prm.sw_gap = s->sw_gap; prm.sw_logu = s->sw_logu;
    prm.k = s->k; prm.n_tune = n_tune; prm.n_keep = n_keep; prm.goal = s->goal;
    CU(launch_pdl(ptlmc_propose_kernel, dim3((unsigned)((n + warps - 1) / warps)), warps * 32,
                  (size_t)warps * 2 * p2 * sizeof(double), st, prm));
    LAUNCH_CHECK();
    if (int r = log_posterior_impl(s->ch, s->prop, -INFINITY, s->lp_prop, s->notpd, n, GPBT_PATH_AUTO, st, nullptr, 0, 0,
                                   /*zero_counter=*/false))
      return r;
    CU(launch_pdl(ptlmc_step_kernel, dim3(1), kPtStepThreads, smem_step, st, prm));
    LAUNCH_CHECK();
    s->cur ^= 1;
  }
  CU(cudaStreamSynchronize(st));
  return 0;
}

// saved_host [n - n_hot, n_keep, p] (or NULL), theta_host [n, p] / lp_host [n] the current state (or NULL),
// info_host[4] = {tau, stride, accepted proposals of the T = 1 chains after tuning, non-PD covariances seen}
extern "C" int gpbt_ptlmc_read(gpbt_ptlmc_t s, double* saved_host, double* theta_host, double* lp_host, double* info_host) {
  if (!s) return fail(GPBT_EINVAL, "gpbt_ptlmc_read: bad argument");
  CU(cudaSetDevice(s->ch->device));
  cudaStream_t st = s->ch->stream;
  CU(cudaStreamSynchronize(st));
  const int n = s->n, p = s->p;
  if (saved_host && s->saved)
    CU(cudaMemcpy(saved_host, s->saved, (size_t)(n - s->n_hot) * s->n_keep * p * sizeof(double), cudaMemcpyDeviceToHost));
  if (theta_host) CU(cudaMemcpy(theta_host, s->theta[s->cur], (size_t)n * p * sizeof(double), cudaMemcpyDeviceToHost));
  if (lp_host) CU(cudaMemcpy(lp_host, s->lp[s->cur], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
  if (info_host) {
    PtlmcCtl c;
    int bad = 0;
    CU(cudaMemcpy(&c, s->ctl, sizeof c, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(&bad, s->notpd, sizeof bad, cudaMemcpyDeviceToHost));
    info_host[0] = c.tau; info_host[1] = c.stride; info_host[2] = (double)c.accepted; info_host[3] = (double)bad;
  }
  return 0;
}
