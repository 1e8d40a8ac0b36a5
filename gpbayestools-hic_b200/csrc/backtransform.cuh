// Kernel (b): PC space -> observable space.
//
// Replaces Emulator._inverse_transform / scaler.inverse_transform, the optional exp(), the
// covariance contraction np.dot(gp_var, _var_trans).reshape(N,m,m) + _cov_trunc and the
// exp_and_cov_diagonal rewrite (reference: src/emulator.py:366-375, 558-601):
//     PCA mode      mean = z A + mu                   cov = sum_k v_k A_k^T A_k + Ctrunc
//     no-PCA mode   mean = z * scale + mu             cov = diag(v)
//     exp_diag      mean = exp(mean)                  cov = diag(diag(cov) * mean^2)
// Output goes into Chain._predict's layout (src/mcmc.py:153-166): mean[N, ld_mean] at column
// col_off; cov[N, ld_cov, ld_cov] rows col_off..col_off+m-1, written across all ld_cov columns
// (zeros outside the diagonal block) so no separate memset is needed.
//
// The covariance is a [m x q] x [q x m] product per walker with 8 m^2 output bytes for 2 q m^2
// flops: at q = 20 it sits on the FP64 ridge (5 flop/B), so it runs on the tensor pipe (DMMA)
// and writes each 8x8 accumulator tile straight to HBM as 64-byte row segments.
#pragma once
#include "common.cuh"

namespace gpbt {

struct BacktransformParams {
  const double* __restrict__ z_mean;  // [N, ldz] (already offset to this emulator's PCs)
  const double* __restrict__ z_var;   // [N, ldz]
  const double* __restrict__ A;       // [q, m_ld] zero padded rows, m_ld >= m (PCA mode)
  const double* __restrict__ mu;      // [m]
  const double* __restrict__ scale;   // [m] (no-PCA mode)
  const double* __restrict__ Ctrunc;  // [m, m] (PCA mode)
  double* __restrict__ mean;          // [N, ld_mean]
  double* __restrict__ cov;           // [N, ld_cov, ld_cov] or null
  int64_t ldz, ld_mean, ld_cov, col_off, N;
  int q, m, m_ld, flags;
};

// ---- mean (and the diagonal-covariance modes) ------------------------------------------------
// one CTA per walker; threads stride over observables
__global__ void __launch_bounds__(128) backtransform_mean_kernel(const BacktransformParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* zs = reinterpret_cast<double*>(smem_raw);  // [q]
  double* vs = zs + prm.q;                           // [q]
  const int64_t w = blockIdx.x;
  const bool no_pca = prm.flags & 1, exp_diag = prm.flags & 2;
  for (int k = threadIdx.x; k < prm.q; k += blockDim.x) {
    zs[k] = prm.z_mean[w * prm.ldz + k];
    vs[k] = prm.z_var[w * prm.ldz + k];
  }
  __syncthreads();
  const bool diag_cov = no_pca || exp_diag;
  for (int o = threadIdx.x; o < prm.m; o += blockDim.x) {
    double mean, var = 0.0;
    if (no_pca) {
      mean = fma(zs[o], prm.scale[o], prm.mu[o]);
      var = vs[o];
    } else {
      double s = 0.0;
      for (int k = 0; k < prm.q; k++) s = fma(zs[k], prm.A[(size_t)k * prm.m_ld + o], s);
      mean = s + prm.mu[o];
      if (exp_diag) {
        double d = prm.Ctrunc[(size_t)o * prm.m + o];
        for (int k = 0; k < prm.q; k++) {
          const double a = prm.A[(size_t)k * prm.m_ld + o];
          d = fma(vs[k], a * a, d);
        }
        var = d;
      }
    }
    if (exp_diag) {
      mean = exp(mean);
      const double f = sqrt(var) * mean;  // (fstd * mean)**2, src/emulator.py:598-599
      var = f * f;
    }
    prm.mean[w * prm.ld_mean + prm.col_off + o] = mean;
    if (diag_cov && prm.cov != nullptr) {
      // whole row (col_off + o) of this walker's matrix: zeros + the diagonal entry
      double* row = prm.cov + ((size_t)w * prm.ld_cov + prm.col_off + o) * prm.ld_cov;
      for (int64_t cidx = 0; cidx < prm.ld_cov; cidx++) row[cidx] = (cidx == prm.col_off + o) ? var : 0.0;
    }
  }
}

// ---- dense covariance (PCA mode) on the FP64 tensor pipe --------------------------------------
// grid = (row tiles of 32, N).  CTA = 4 warps; warp w handles column tiles w, w+4, ... of 32
// columns.  D[i][j] = sum_k (v_k A[k][i]) * A[k][j]: A-operand fragment (row i = g, k = t) and
// B-operand fragment (k = t, col j = g) both come from the same [q_pad][m_ld] array in shared
// memory; m_ld = 4 mod 8 keeps those loads bank-conflict free.
constexpr int kBtThreads = 128;
constexpr int kBtRows = 32;

inline size_t backtransform_smem_bytes(int q_pad, int m_ld) {
  return sizeof(double) * ((size_t)q_pad * m_ld + (size_t)q_pad * kBtRows + q_pad);
}

__global__ void __launch_bounds__(kBtThreads) backtransform_cov_kernel(const BacktransformParams prm, int q_pad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = prm.m, m_ld = prm.m_ld;
  double* As = reinterpret_cast<double*>(smem_raw);  // [q_pad][m_ld]
  double* Av = As + (size_t)q_pad * m_ld;            // [q_pad][32]: v_k * A[k][i0 + r]
  double* vs = Av + (size_t)q_pad * kBtRows;         // [q_pad]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int64_t w = blockIdx.y;
  const int i0 = blockIdx.x * kBtRows;

  for (int k = tid; k < q_pad; k += kBtThreads) vs[k] = k < prm.q ? prm.z_var[w * prm.ldz + k] : 0.0;
  for (int idx = tid; idx < q_pad * m_ld; idx += kBtThreads) {
    const int k = idx / m_ld;
    As[idx] = k < prm.q ? prm.A[idx] : 0.0;
  }
  __syncthreads();
  for (int idx = tid; idx < q_pad * kBtRows; idx += kBtThreads) {
    const int k = idx / kBtRows, r = idx - k * kBtRows;
    Av[idx] = (i0 + r < m) ? vs[k] * As[(size_t)k * m_ld + i0 + r] : 0.0;
  }
  __syncthreads();

  const int64_t ldc = prm.ld_cov, off = prm.col_off;
  double* out = prm.cov + ((size_t)w * ldc + off + i0) * ldc;
  const int n_ct = (m + 31) / 32;
  for (int ct = warp; ct < n_ct; ct += kBtThreads / 32) {
    const int j0 = ct * 32;
    double acc[4][4][2];
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
      for (int nb = 0; nb < 4; nb++) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    for (int k0 = 0; k0 < q_pad; k0 += 4) {
      double a[4], b[4];
#pragma unroll
      for (int mb = 0; mb < 4; mb++) a[mb] = Av[(k0 + t) * kBtRows + 8 * mb + g];
#pragma unroll
      for (int nb = 0; nb < 4; nb++) {
        const int j = j0 + 8 * nb + g;
        b[nb] = j < m_ld ? As[(size_t)(k0 + t) * m_ld + j] : 0.0;
      }
#pragma unroll
      for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
    }
    // epilogue: + Ctrunc, store.  lane owns D[8mb + g][8nb + 2t + {0,1}]
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
      const int i = i0 + 8 * mb + g;
      if (i >= m) continue;
#pragma unroll
      for (int nb = 0; nb < 4; nb++) {
        const int j = j0 + 8 * nb + 2 * t;
#pragma unroll
        for (int h = 0; h < 2; h++)
          if (j + h < m)
            out[(size_t)(8 * mb + g) * ldc + off + j + h] = acc[mb][nb][h] + prm.Ctrunc[(size_t)i * m + j + h];
      }
    }
  }
  // zero the parts of these rows that lie outside the diagonal block (multi-emulator chains)
  if (ldc > m) {
    for (int r = warp; r < kBtRows && i0 + r < m; r += kBtThreads / 32)
      for (int64_t cidx = lane; cidx < ldc; cidx += 32)
        if (cidx < off || cidx >= off + m) out[(size_t)r * ldc + cidx] = 0.0;
  }
}

}  // namespace gpbt
