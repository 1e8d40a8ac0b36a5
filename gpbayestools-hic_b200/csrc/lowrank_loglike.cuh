// Exact low-rank evaluation of the walker log-likelihood for PCA-mode emulators.
//
// For PCA-mode emulators the reference's per-walker covariance (src/emulator.py:584-587,
// src/mcmc.py:153-166,288-290) is a walker-independent matrix plus a rank-Q term:
//     C_w = F + U^T diag(v_w) U,     F = blockdiag(_cov_trunc_e) + expdata_cov,
//     dy_w = U^T z_w + r0,           U = blockdiag(_trans_matrix_e[:npc_e]),  r0 = mu - y_exp.
// With L_F = chol(F) and the thin QR factorisation  L_F^-1 U^T = Qb R  (host, once per chain):
//     log det C_w        = log det F + log det (I_Q + R diag(v_w) R^T)
//     dy^T C_w^-1 dy     = s_perp + c^T (I_Q + R diag(v_w) R^T)^-1 c,    c = R z_w + c0
// (c0 = Qb^T L_F^-1 r0, s_perp = |(I - Qb Qb^T) L_F^-1 r0|^2).  Both identities are exact, involve
// no cancellation (S = I + R D R^T has eigenvalues >= 1 for v >= 0) and need a QxQ Cholesky per
// walker instead of the MxM one of mvn_loglike (src/mcmc.py:23-65).
//
// One warp per walker; S lives in shared memory.  The kernel also applies the bounds mask and the
// constant 2*log(1e-16) of Chain.log_posterior (src/mcmc.py:275-276, 296-297).
#pragma once
#include "common.cuh"

namespace gpbt {

struct LowrankParams {
  const double* __restrict__ X;       // [N, p]
  const double* __restrict__ lo;      // [p]
  const double* __restrict__ hi;      // [p]
  const double* __restrict__ z_mean;  // [N, Q]
  const double* __restrict__ z_var;   // [N, Q]
  const double* __restrict__ R;       // [Q, Q] upper triangular, row-major
  const double* __restrict__ c0;      // [Q]
  double* __restrict__ lp;            // [N]
  int* __restrict__ n_notpd;          // may be null
  double s_perp, logdetF_half, oob_value, sys_const;
  int64_t N;
  int p, Q;
};

constexpr int kLrWarps = 4;

__host__ __device__ inline int lowrank_stride(int Q) { return Q | 1; }
inline size_t lowrank_smem_bytes(int Q) {
  return sizeof(double) * ((size_t)Q * Q + (size_t)kLrWarps * ((size_t)Q * lowrank_stride(Q) + 3 * (size_t)Q));
}

__global__ void __launch_bounds__(kLrWarps * 32) lowrank_loglike_kernel(const LowrankParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Q = prm.Q, ld = lowrank_stride(Q);
  double* Rs = reinterpret_cast<double*>(smem_raw);                 // [Q][Q]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* S = Rs + Q * Q + (size_t)warp * (Q * ld + 3 * Q);         // [Q][ld]
  double* cv = S + Q * ld;                                          // [Q]
  double* vs = cv + Q;                                              // [Q]
  double* zs = vs + Q;                                              // [Q]

  for (int i = threadIdx.x; i < Q * Q; i += blockDim.x) Rs[i] = prm.R[i];
  __syncthreads();

  const int64_t w = (int64_t)blockIdx.x * kLrWarps + warp;
  if (w >= prm.N) return;

  // strict bounds mask (NaN compares false -> outside)
  bool ok = true;
  for (int d = lane; d < prm.p; d += 32) {
    const double x = prm.X[w * prm.p + d];
    ok = ok && (x > prm.lo[d]) && (x < prm.hi[d]);
  }
  if (!__all_sync(0xffffffffu, ok)) {
    if (lane == 0) prm.lp[w] = prm.oob_value;
    return;
  }

  for (int a = lane; a < Q; a += 32) {
    zs[a] = prm.z_mean[w * Q + a];
    vs[a] = prm.z_var[w * Q + a];
  }
  __syncwarp();
  for (int a = lane; a < Q; a += 32) {
    double s = prm.c0[a];
    for (int k = a; k < Q; k++) s = fma(Rs[a * Q + k], zs[k], s);
    cv[a] = s;
  }
  // S (lower triangle) = I + R diag(v) R^T ; R upper triangular -> k runs from max(a,b) = a
  for (int idx = lane; idx < Q * (Q + 1) / 2; idx += 32) {
    int a = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while ((a + 1) * (a + 2) / 2 <= idx) a++;
    while (a * (a + 1) / 2 > idx) a--;
    const int b = idx - a * (a + 1) / 2;
    double s = (a == b) ? 1.0 : 0.0;
    for (int k = a; k < Q; k++) s = fma(Rs[a * Q + k] * vs[k], Rs[b * Q + k], s);
    S[a * ld + b] = s;
  }
  __syncwarp();

  // in-place Cholesky (lower), right-looking, one column per step
  double logdet = 0.0;
  bool pd = true;
  for (int b = 0; b < Q; b++) {
    const double d = S[b * ld + b];
    if (!(d > 0.0)) { pd = false; break; }
    const double l = sqrt(d);
    logdet += log(l);
    const double inv = 1.0 / l;
    for (int a = b + 1 + lane; a < Q; a += 32) S[a * ld + b] *= inv;
    __syncwarp();
    if (lane == 0) S[b * ld + b] = l;  // after the barrier: every lane has read the old diagonal
    for (int a = b + 1 + lane; a < Q; a += 32) {
      const double lab = S[a * ld + b];
      for (int c = b + 1; c <= a; c++) S[a * ld + c] = fma(-lab, S[c * ld + b], S[a * ld + c]);
    }
    __syncwarp();
  }
  if (!pd) {
    if (lane == 0) {
      prm.lp[w] = prm.oob_value;
      if (prm.n_notpd) atomicAdd(prm.n_notpd, 1);
    }
    return;
  }
  // forward solve t = L^-1 c (column oriented), quad = |t|^2
  double quad = 0.0;
  for (int b = 0; b < Q; b++) {
    const double tb = cv[b] / S[b * ld + b];
    quad = fma(tb, tb, quad);
    __syncwarp();
    for (int a = b + 1 + lane; a < Q; a += 32) cv[a] = fma(-S[a * ld + b], tb, cv[a]);
    __syncwarp();
  }
  if (lane == 0)
    prm.lp[w] = -0.5 * (prm.s_perp + quad) - logdet - prm.logdetF_half + prm.sys_const;
}

}  // namespace gpbt
