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
  double* __restrict__ var_diag;      // [N, ld_mean] or null: diag(cov) only (posterior-predictive sweeps)
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
      if (exp_diag || prm.var_diag != nullptr) {
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
    if (prm.var_diag != nullptr) prm.var_diag[w * prm.ld_mean + prm.col_off + o] = var;
    if (diag_cov && prm.cov != nullptr) {
      // whole row (col_off + o) of this walker's matrix: zeros + the diagonal entry
      double* row = prm.cov + ((size_t)w * prm.ld_cov + prm.col_off + o) * prm.ld_cov;
      for (int64_t cidx = 0; cidx < prm.ld_cov; cidx++) row[cidx] = (cidx == prm.col_off + o) ? var : 0.0;
    }
  }
}

// ---- dense covariance (PCA mode) on the FP64 tensor pipe --------------------------------------
// D[i][j] = Ctrunc[i][j] + sum_k (v_k A[k][i]) * A[k][j] per walker.  Persistent CTAs keep A in
// shared memory ([q_pad][m_ld], m_ld = 4 mod 8: the A-operand fragment (row i = g, k = t) and the
// B-operand fragment (k = t, col j = g) are both read from it without bank conflicts); their WARPS
// walk independently over work items (walker, 32-row tile rt, group of four n8 column blocks), so
// there is no block-level synchronisation after the prologue.  The accumulators start from the
// Ctrunc tile (its L2 latency overlaps other warps' tensor work); v_k multiplies the A fragment in
// registers.
// cov is symmetric, so only the column groups at or right of the diagonal 32x32 tile are computed
// and every off-diagonal 8x8 block is stored twice: directly (16-byte stores, 64-byte row segments
// per quad) and mirrored (lanes of equal t write 8 consecutive doubles of a row).  That halves the
// tensor work, which otherwise is as large as the HBM write stream (both ~30 k cycles per walker
// per SM at q = 20, m = 300).
constexpr int kBtThreads = 128;
constexpr int kBtWarps = kBtThreads / 32;
constexpr int kBtRows = 32;

inline size_t backtransform_smem_bytes(int q_pad, int m_ld) { return sizeof(double) * (size_t)q_pad * m_ld; }

constexpr int kBtNBI = 2;     // n8 column blocks per work item (a 32 x 16 tile)
constexpr int kBtGroup = 8;   // walkers that share one load of the Ctrunc tile

// number of (row tile, column group) items of one walker
__host__ __device__ inline int bt_items_per_walker(int m) {
  const int n_rt = (m + kBtRows - 1) / kBtRows, n_nb = (m + 7) / 8;
  int cnt = 0;
  for (int rt = 0; rt < n_rt; rt++) cnt += (n_nb - 4 * rt + kBtNBI - 1) / kBtNBI;
  return cnt;
}

// A work item is one 32 x 16 tile position for a group of kBtGroup consecutive walkers: the warp keeps
// the Ctrunc tile in registers and walks through the group, so Ctrunc is read from L2 once per eight
// output tiles and its latency is off the path of all but the first.
__global__ void __launch_bounds__(kBtThreads, 4) backtransform_cov_kernel(const BacktransformParams prm, int q_pad) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = prm.m, m_ld = prm.m_ld;
  double* As = reinterpret_cast<double*>(smem_raw);  // [q_pad][m_ld]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;

  for (int idx = tid; idx < q_pad * m_ld; idx += kBtThreads) {
    const int k = idx / m_ld;
    As[idx] = k < prm.q ? prm.A[idx] : 0.0;
  }
  __syncthreads();

  const int n_nb = (m + 7) / 8;
  const int per_walker = bt_items_per_walker(m);
  const int64_t n_groups = (prm.N + kBtGroup - 1) / kBtGroup;
  const int64_t n_items = n_groups * per_walker;
  const int64_t ldc = prm.ld_cov, off = prm.col_off;
  const bool vec_ok = (ldc % 2 == 0) && (off % 2 == 0) && ((reinterpret_cast<uintptr_t>(prm.cov) & 15) == 0);
  const bool ct_vec = (m % 2 == 0) && ((reinterpret_cast<uintptr_t>(prm.Ctrunc) & 15) == 0);
  const int64_t warp_stride = (int64_t)gridDim.x * kBtWarps;

  for (int64_t item = (int64_t)blockIdx.x * kBtWarps + warp; item < n_items; item += warp_stride) {
    const int64_t wg = item / per_walker;
    int rem = (int)(item - wg * per_walker);
    int rt = 0;
    for (;; rt++) {  // items of a walker group are ordered by row tile, then column group
      const int cnt = (n_nb - 4 * rt + kBtNBI - 1) / kBtNBI;
      if (rem < cnt) break;
      rem -= cnt;
    }
    const int i0 = rt * kBtRows, nb_first = 4 * rt, nb0 = nb_first + kBtNBI * rem;
    const int nb_hi = min(nb0 + kBtNBI, n_nb);

    // the Ctrunc tile; lane owns D[8mb + g][8nb + 2t + {0,1}]
    double ct[4][kBtNBI][2];
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
      const int i = i0 + 8 * mb + g;
#pragma unroll
      for (int nb = 0; nb < kBtNBI; nb++) {
        const int j = 8 * (nb0 + nb) + 2 * t;
        const bool in = (i < m) && (nb0 + nb < nb_hi);
        if (in && ct_vec && j + 1 < m) {
          const double2 c2 = ldg2(prm.Ctrunc + (size_t)i * m + j);
          ct[mb][nb][0] = c2.x;
          ct[mb][nb][1] = c2.y;
        } else {
          ct[mb][nb][0] = (in && j < m) ? __ldg(prm.Ctrunc + (size_t)i * m + j) : 0.0;
          ct[mb][nb][1] = (in && j + 1 < m) ? __ldg(prm.Ctrunc + (size_t)i * m + j + 1) : 0.0;
        }
      }
    }

    const int64_t w_end = min(prm.N, (wg + 1) * kBtGroup);
#pragma unroll 1
    for (int64_t w = wg * kBtGroup; w < w_end; w++) {
      double* wbase = prm.cov + (size_t)w * ldc * ldc + (size_t)off * ldc + off;  // block (0,0) of this emulator
      const double* zv = prm.z_var + w * prm.ldz;
      double acc[4][kBtNBI][2];
#pragma unroll
      for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < kBtNBI; nb++) {
          acc[mb][nb][0] = ct[mb][nb][0];
          acc[mb][nb][1] = ct[mb][nb][1];
        }
      for (int k0 = 0; k0 < q_pad; k0 += 4) {
        const double vk = (k0 + t < prm.q) ? zv[k0 + t] : 0.0;
        double a[4], b[kBtNBI];
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
          const int i = i0 + 8 * mb + g;
          a[mb] = (i < m_ld) ? vk * As[(size_t)(k0 + t) * m_ld + i] : 0.0;
        }
#pragma unroll
        for (int nb = 0; nb < kBtNBI; nb++) {
          const int j = 8 * (nb0 + nb) + g;
          b[nb] = j < m_ld ? As[(size_t)(k0 + t) * m_ld + j] : 0.0;
        }
#pragma unroll
        for (int nb = 0; nb < kBtNBI; nb++)
          if (nb0 + nb < nb_hi) {
#pragma unroll
            for (int mb = 0; mb < 4; mb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
          }
      }
#pragma unroll
      for (int mb = 0; mb < 4; mb++) {
        const int i = i0 + 8 * mb + g;
        if (i >= m) continue;
#pragma unroll
        for (int nb = 0; nb < kBtNBI; nb++) {
          const int j = 8 * (nb0 + nb) + 2 * t;
          if (nb0 + nb >= nb_hi || j >= m) continue;
          double* row = wbase + (size_t)i * ldc;
          if (vec_ok && j + 1 < m) {
            *reinterpret_cast<double2*>(row + j) = make_double2(acc[mb][nb][0], acc[mb][nb][1]);
          } else {
            row[j] = acc[mb][nb][0];
            if (j + 1 < m) row[j + 1] = acc[mb][nb][1];
          }
          // mirror of the blocks right of the diagonal 32x32 tile: cov[j][i] = cov[i][j]
          if (nb0 + nb >= nb_first + 4) {
            wbase[(size_t)j * ldc + i] = acc[mb][nb][0];
            if (j + 1 < m) wbase[(size_t)(j + 1) * ldc + i] = acc[mb][nb][1];
          }
        }
      }
      // zero the parts of these rows that lie outside the diagonal block (multi-emulator chains);
      // done by the item that owns the diagonal tile of the row tile
      if (ldc > m && rem == 0) {
        double* rows0 = prm.cov + ((size_t)w * ldc + off + i0) * ldc;
        for (int r = 0; r < kBtRows && i0 + r < m; r++)
          for (int64_t cidx = lane; cidx < ldc; cidx += 32)
            if (cidx < off || cidx >= off + m) rows0[(size_t)r * ldc + cidx] = 0.0;
      }
    }
  }
}

}  // namespace gpbt
