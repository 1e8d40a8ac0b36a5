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

constexpr int kMaxPeers = 16;

struct LowrankParams {
  const double* __restrict__ X;       // [N, p]
  const double* __restrict__ lo;      // [p]
  const double* __restrict__ hi;      // [p]
  const double* __restrict__ z_mean;  // [N, ldz], this block's PCs start at column z_off
  const double* __restrict__ z_var;   // [N, ldz]
  int64_t ldz;
  int z_off;
  // Block-separable chains (experimental covariance block diagonal over the emulators): log L is a
  // sum over emulator blocks, evaluated by one launch per block; all but the first ADD their term to
  // lp (and to the peer buffers) instead of storing it.
  int accumulate;
  const double* __restrict__ R;       // [Q, Q] upper triangular, row-major
  const double* __restrict__ c0;      // [Q]
  double* __restrict__ lp;            // [N]
  // fused all-gather: when n_peers > 0 every result also goes to peers[r][peer_off + w] -- buffers of
  // the other GPUs mapped into this process (NVLink peer stores), so no collective follows the kernel
  double* peers[kMaxPeers];
  int n_peers;
  int64_t peer_off;
  int* __restrict__ n_notpd;          // may be null
  double s_perp, logdetF_half, oob_value, sys_const;
  int64_t N;
  int p, Q;
};

constexpr int kLrWarps = 4;

// `term` is this block's contribution; `failed` (out of bounds / not positive definite) forces the
// out-of-bounds value whatever the other blocks say
__device__ __forceinline__ void lowrank_store(const LowrankParams& prm, int64_t w, double term, bool failed = false) {
  double v = term;
  if (failed) {
    v = prm.oob_value;
  } else if (prm.accumulate) {
    const double prev = prm.lp[w];
    v = (prev == prm.oob_value) ? prev : prev + term;   // an earlier block already failed this walker
  }
  prm.lp[w] = v;
  for (int r = 0; r < prm.n_peers; r++) prm.peers[r][prm.peer_off + w] = v;
}

__host__ __device__ inline int lowrank_stride(int Q) { return Q | 1; }
inline size_t lowrank_smem_bytes(int Q) {
  return sizeof(double) * (2 * (size_t)Q * lowrank_stride(Q) +
                           (size_t)kLrWarps * ((size_t)Q * lowrank_stride(Q) + 4 * (size_t)Q));
}

// Lanes own rows (a = lane, lane + 32, ...); S is kept column-major in shared memory
// (S[c * ld + a]) so every per-column step is one conflict-free vector operation across the warp
// and the loops over columns are warp-uniform.
__global__ void __launch_bounds__(kLrWarps * 32) lowrank_loglike_kernel(const LowrankParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  const int Q = prm.Q, ld = lowrank_stride(Q);
  double* Rt = reinterpret_cast<double*>(smem_raw);  // Rt[k * ld + a] = R[a][k]   (zero for k < a)
  double* Rr = Rt + (size_t)Q * ld;                  // Rr[a * ld + k] = R[a][k]   (row-major, broadcast reads)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* S = Rr + (size_t)Q * ld + (size_t)warp * ((size_t)Q * ld + 4 * Q);  // [Q cols][ld]
  double* cv = S + (size_t)Q * ld;                   // [Q]
  double* vs = cv + Q;                               // [Q]
  double* zs = vs + Q;                               // [Q]
  double* dg = zs + Q;                               // [Q] pivots d_b (for the log-determinant)

  for (int i = threadIdx.x; i < Q * Q; i += blockDim.x) {
    const int a = i / Q, k = i - a * Q;
    const double r = (k >= a) ? prm.R[i] : 0.0;
    Rt[k * ld + a] = r;
    Rr[a * ld + k] = r;
  }
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
    if (lane == 0) lowrank_store(prm, w, 0.0, true);
    return;
  }

  for (int a = lane; a < Q; a += 32) {
    zs[a] = prm.z_mean[w * prm.ldz + prm.z_off + a];
    vs[a] = prm.z_var[w * prm.ldz + prm.z_off + a];
  }
  __syncwarp();
  // c = c0 + R z ;  S[:, b] (rows a >= b) = delta_ab + sum_k R[a][k] v_k R[b][k]
  for (int a = lane; a < Q; a += 32) {
    double s = prm.c0[a];
    for (int k = 0; k < Q; k++) s = fma(Rt[k * ld + a], zs[k], s);
    cv[a] = s;
  }
  for (int b = 0; b < Q; b++) {
    for (int a = b + lane; a < Q; a += 32) {
      double s = (a == b) ? 1.0 : 0.0;
      for (int k = b; k < Q; k++) s = fma(Rt[k * ld + a], vs[k] * Rr[b * ld + k], s);
      S[b * ld + a] = s;
    }
  }
  __syncwarp();

  // right-looking Cholesky, one column per step; the forward solve t = L^-1 c rides along
  bool pd = true;
  double quad = 0.0;
  for (int b = 0; b < Q; b++) {
    const double d = S[b * ld + b];
    if (!(d > 0.0)) { pd = false; break; }
    const double inv = rsqrt(d);
    const double tb = cv[b] * inv;                   // t_b = c_b / l_bb
    quad = fma(tb, tb, quad);
    __syncwarp();
    for (int a = b + 1 + lane; a < Q; a += 32) {
      const double lab = S[b * ld + a] * inv;        // L[a][b]
      S[b * ld + a] = lab;
      cv[a] = fma(-lab, tb, cv[a]);
    }
    if (lane == 0) dg[b] = d;
    __syncwarp();
    for (int c = b + 1; c < Q; c++) {                // trailing update, column by column
      const double lcb = S[b * ld + c];
      for (int a = c + lane; a < Q; a += 32) S[c * ld + a] = fma(-S[b * ld + a], lcb, S[c * ld + a]);
    }
    __syncwarp();
  }
  if (!pd) {
    if (lane == 0) {
      lowrank_store(prm, w, 0.0, true);
      if (prm.n_notpd) atomicAdd(prm.n_notpd, 1);
    }
    return;
  }
  double logdet2 = 0.0;                              // sum log d_b = 2 sum log l_bb
  for (int b = lane; b < Q; b += 32) logdet2 += log(dg[b]);
  logdet2 = warp_sum(logdet2);
  if (lane == 0)
    lowrank_store(prm, w, -0.5 * (prm.s_perp + quad) - 0.5 * logdet2 - prm.logdetF_half + prm.sys_const);
}

// ---- register-resident variant, Q <= 32 ----------------------------------------------------------
// Lane a owns row a of S in registers (QP = Q rounded up to a multiple of 4 is a compile-time
// constant, so every index is static); pivots and L[c][b] travel by warp shuffle.  About 1.1k warp
// instructions per walker at Q = 20 instead of the ~10k of the shared-memory kernel above, which
// remains the fallback for Q > 32.  Rows/columns Q..QP-1 are identity padding (v = z = c0 = 0),
// which leaves both the determinant and the quadratic form unchanged.
template <int QP>
__global__ void __launch_bounds__(kLrWarps * 32) lowrank_loglike_reg_kernel(const LowrankParams prm) {
  constexpr int LD = QP + 2;                 // even stride (16-byte row alignment), conflict-light
  __shared__ __align__(16) double Rr[QP * LD];            // R row-major, zero below the diagonal
  __shared__ __align__(16) double zv[kLrWarps][2 * QP];   // per warp: z[QP], v[QP]
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  const int Q = prm.Q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < QP * LD; i += blockDim.x) {
    const int a = i / LD, k = i - a * LD;
    double r = 0.0;
    if (a < Q && k < Q && k >= a) r = prm.R[a * Q + k];
    else if (a == k && a >= Q && a < QP) r = 1.0;
    Rr[i] = r;
  }
  __syncthreads();
  const int64_t w = (int64_t)blockIdx.x * kLrWarps + warp;
  if (w >= prm.N) return;

  bool ok = true;
  for (int d = lane; d < prm.p; d += 32) {
    const double x = prm.X[w * prm.p + d];
    ok = ok && (x > prm.lo[d]) && (x < prm.hi[d]);
  }
  if (!__all_sync(0xffffffffu, ok)) {
    if (lane == 0) lowrank_store(prm, w, 0.0, true);
    return;
  }
  double* zs = zv[warp];
  double* vs = zs + QP;
  for (int a = lane; a < QP; a += 32) {
    zs[a] = a < Q ? prm.z_mean[w * prm.ldz + prm.z_off + a] : 0.0;
    vs[a] = a < Q ? prm.z_var[w * prm.ldz + prm.z_off + a] : 0.0;
  }
  __syncwarp();

  const int a = lane < QP ? lane : QP - 1;   // idle lanes shadow the last row (results unused)
  double rd[QP];                             // R[a][k] * v_k
  double cv = (lane < Q) ? prm.c0[a] : 0.0;
#pragma unroll
  for (int k = 0; k < QP; k += 2) {
    const double2 r = *reinterpret_cast<const double2*>(&Rr[a * LD + k]);
    const double2 z = *reinterpret_cast<const double2*>(&zs[k]);
    const double2 v = *reinterpret_cast<const double2*>(&vs[k]);
    cv = fma(r.x, z.x, cv);
    cv = fma(r.y, z.y, cv);
    rd[k] = r.x * v.x;
    rd[k + 1] = r.y * v.y;
  }
  double S[QP];                              // row a of S = I + R diag(v) R^T (entries b <= a are used)
#pragma unroll
  for (int b = 0; b < QP; b++) {
    double s0 = (a == b) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
    for (int k = b & ~1; k < QP; k += 2) {   // R[b][k] = 0 for k < b, so starting at the even k below b is exact
      const double2 r = *reinterpret_cast<const double2*>(&Rr[b * LD + k]);
      s0 = fma(rd[k], r.x, s0);
      s1 = fma(rd[k + 1], r.y, s1);
    }
    S[b] = s0 + s1;
  }

  bool pd = true;
  double quad = 0.0, pivot = 1.0;
#pragma unroll
  for (int b = 0; b < QP; b++) {
    const double d = __shfl_sync(0xffffffffu, S[b], b);
    pd = pd && (d > 0.0);
    const double inv = rsqrt(d);
    if (lane == b) pivot = d;
    const double lab = S[b] * inv;                                   // L[a][b] (meaningful for a > b)
    const double tb = __shfl_sync(0xffffffffu, cv, b) * inv;          // t_b = c_b / l_bb
    quad = fma(tb, tb, quad);
    cv = fma(-lab, tb, cv);
#pragma unroll
    for (int c = b + 1; c < QP; c++) {
      const double lcb = __shfl_sync(0xffffffffu, lab, c);
      S[c] = fma(-lab, lcb, S[c]);
    }
  }
  double logdet2 = (lane < QP) ? log(pivot) : 0.0;                    // sum log d_b = 2 sum log l_bb
  logdet2 = warp_sum(logdet2);
  if (lane == 0) {
    if (!pd) {
      lowrank_store(prm, w, 0.0, true);
      if (prm.n_notpd) atomicAdd(prm.n_notpd, 1);
    } else {
      lowrank_store(prm, w, -0.5 * (prm.s_perp + quad) - 0.5 * logdet2 - prm.logdetF_half + prm.sys_const);
    }
  }
}

}  // namespace gpbt
