// Kernel (c): batched per-walker blocked Cholesky with the triangular solve and log-determinant
// fused, one log-likelihood per walker.
//
// Replaces list(map(mvn_loglike, dY, cov)) (reference: src/mcmc.py:23-65, 288-293):
//     y = mean - y_exp ;  C = cov + cov_add ;  U = dpotrf(C) ;  alpha = dpotrs(U, y)
//     lp = -1/2 y.alpha - sum(log(diag(U)))
// evaluated as  t = L^-1 y,  lp = -1/2 |t|^2 - sum(log(diag(L)))  with L = U^T.
//
// One CTA (4 warps) per walker, left-looking blocked factorisation with panel width 16.  The
// factor is written in place over the lower triangle of cov (HBM/L2 resident: a 300x300 FP64
// matrix is 720 KB, more than one SM's shared memory); only the current (m-J) x 16 panel lives in
// shared memory.  The panel update  P -= L[J:, :J] L[J:J+16, :J]^T  is the m^3/3 term and runs on
// the FP64 tensor pipe (DMMA.8x8x4) with fragments loaded straight from global memory -- rows of
// L are private to one warp, and the 16 rows of the B operand are re-read through L1.  The 16x16
// diagonal block is factorised by one warp; the panel's triangular solve is one thread per row.
// Several CTAs are resident per SM so one walker's sequential diagonal step overlaps another's
// tensor work.
#pragma once
#include "common.cuh"

namespace gpbt {

struct CholParams {
  const double* __restrict__ mean;     // [N, m]
  const double* __restrict__ y_exp;    // [m] or null
  double* cov;                         // [N, m, m], overwritten with L (lower)
  const double* __restrict__ cov_add;  // [m, m] or null
  double* __restrict__ lp;             // [N]
  int* __restrict__ n_notpd;           // or null
  const unsigned char* __restrict__ skip;  // [N] or null: nonzero -> leave lp[w] untouched
  double notpd_value, add_const;
  int64_t N;
  int m;
};

constexpr int kChThreads = 128;
constexpr int kChWarps = kChThreads / 32;
constexpr int kChNB = 16;        // panel width
constexpr int kChLd = kChNB + 1; // panel row stride in shared memory
constexpr int kChMBW = 4;        // m8 row blocks per warp per update pass

__host__ __device__ inline int chol_rows_pad(int m) {
  const int r = (int)round_up(m, 8);
  return r < kChNB ? kChNB : r;
}
inline size_t chol_smem_bytes(int m) {
  const int rows = chol_rows_pad(m);
  return sizeof(double) * ((size_t)rows * kChLd + (size_t)m + kChNB * kChLd + 64);
}

// four consecutive doubles of a row of L; 16-byte loads when the row stride keeps them aligned
__device__ __forceinline__ void load4(const double* p, bool aligned, double (&v)[4]) {
  if (aligned) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3];
  }
}

// One pass of the panel update: this warp's row blocks  blk = warp + 4*(MBW*pass + i), i < MBW.
template <int MBW>
__device__ __forceinline__ void panel_update(const double* Lw, int m, int J, int nrows, int nb, bool aligned,
                                             double* P, int warp, int lane, int pass) {
  const int g = lane >> 2, t = lane & 3;
  double acc[MBW][2][2];
#pragma unroll
  for (int i = 0; i < MBW; i++)
#pragma unroll
    for (int h = 0; h < 2; h++) acc[i][h][0] = acc[i][h][1] = 0.0;
  const int nblk = (nrows + 7) >> 3;
  int blk[MBW], rowi[MBW];
#pragma unroll
  for (int i = 0; i < MBW; i++) {
    blk[i] = warp + kChWarps * (MBW * pass + i);
    rowi[i] = min(J + 8 * blk[i] + g, m - 1);  // clamped: padded rows are computed but never stored
  }
  const int brow0 = min(J + g, m - 1), brow1 = min(J + 8 + g, m - 1);
  for (int k0 = 0; k0 < J; k0 += 16) {
    double b[2][4], a[MBW][4];
    load4(Lw + (size_t)brow0 * m + k0 + 4 * t, aligned, b[0]);
    load4(Lw + (size_t)brow1 * m + k0 + 4 * t, aligned, b[1]);
#pragma unroll
    for (int i = 0; i < MBW; i++)
      if (blk[i] < nblk) load4(Lw + (size_t)rowi[i] * m + k0 + 4 * t, aligned, a[i]);
#pragma unroll
    for (int s = 0; s < 4; s++)
#pragma unroll
      for (int i = 0; i < MBW; i++)
        if (blk[i] < nblk) {
          dmma884(acc[i][0][0], acc[i][0][1], a[i][s], b[0][s]);
          dmma884(acc[i][1][0], acc[i][1][1], a[i][s], b[1][s]);
        }
  }
#pragma unroll
  for (int i = 0; i < MBW; i++) {
    const int r = 8 * blk[i] + g;
    if (blk[i] < nblk && r < nrows) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int c = 8 * h + 2 * t;
        if (c < nb) P[r * kChLd + c] -= acc[i][h][0];
        if (c + 1 < nb) P[r * kChLd + c + 1] -= acc[i][h][1];
      }
    }
  }
}

__global__ void __launch_bounds__(kChThreads) chol_loglike_kernel(const CholParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = prm.m;
  const int rows_pad = chol_rows_pad(m);
  double* P = reinterpret_cast<double*>(smem_raw);  // [rows_pad][17] current panel (rows J..m-1)
  double* tv = P + (size_t)rows_pad * kChLd;        // [m] forward-solve vector t
  double* D = tv + m;                               // [16][17] factorised diagonal block
  double* red = D + kChNB * kChLd;                  // scratch
  __shared__ int s_bad;

  const int64_t w = blockIdx.x;
  if (prm.skip != nullptr && prm.skip[w]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* Lw = prm.cov + (size_t)w * m * m;
  const bool aligned = (m % 2 == 0) && ((reinterpret_cast<uintptr_t>(Lw) & 15) == 0);
  if (tid == 0) s_bad = 0;

  double logdet = 0.0;  // meaningful in warp 0
  for (int J = 0; J < m; J += kChNB) {
    const int nrows = m - J;                 // panel rows J .. m-1
    const int nb = min(kChNB, nrows);        // real columns in this panel
    // 1. load the panel (+ cov_add); padded columns become identity columns
    for (int idx = tid; idx < rows_pad * kChNB; idx += kChThreads) {
      const int r = idx / kChNB, c = idx - r * kChNB;
      double v = 0.0;
      if (r < nrows && c < nb) {
        v = Lw[(size_t)(J + r) * m + J + c];
        if (prm.cov_add) v += prm.cov_add[(size_t)(J + r) * m + J + c];
      } else if (r == c) {
        v = 1.0;  // identity padding of a narrow last panel
      }
      P[r * kChLd + c] = v;
    }
    __syncthreads();
    // 2. P -= L[J:, :J] L[J:J+16, :J]^T
    if (J > 0) {
      const int per_warp = ((nrows + 7) / 8 + kChWarps - 1) / kChWarps;
      for (int pass = 0; pass * kChMBW < per_warp; pass++)
        panel_update<kChMBW>(Lw, m, J, nrows, nb, aligned, P, warp, lane, pass);
      __syncthreads();
    }
    // 3. factorise the 16x16 diagonal block (warp 0; lane = row) and forward-solve 16 entries of t
    if (warp == 0) {
      const int r = lane;
      for (int c = 0; c < kChNB; c++) {
        const double d = P[c * kChLd + c];
        if (!(d > 0.0)) { if (lane == 0) s_bad = 1; break; }
        const double l = sqrt(d), inv = 1.0 / l;
        logdet += log(l);
        __syncwarp();
        if (r == c) P[c * kChLd + c] = l;
        if (r > c && r < kChNB) P[r * kChLd + c] *= inv;
        __syncwarp();
        if (r > c && r < kChNB) {
          const double lrc = P[r * kChLd + c];
          for (int c2 = c + 1; c2 <= r; c2++) P[r * kChLd + c2] = fma(-lrc, P[c2 * kChLd + c], P[r * kChLd + c2]);
        }
        __syncwarp();
      }
      for (int idx = lane; idx < kChNB * kChNB; idx += 32) {
        const int rr = idx / kChNB, cc = idx - rr * kChNB;
        D[rr * kChLd + cc] = P[rr * kChLd + cc];
      }
    }
    __syncthreads();
    if (s_bad) break;
    // 4. rows below the diagonal block: P[r, :] <- P[r, :] D^-T  (one thread per row)
    for (int r = kChNB + tid; r < nrows; r += kChThreads) {
      double x[kChNB];
#pragma unroll
      for (int c = 0; c < kChNB; c++) x[c] = P[r * kChLd + c];
#pragma unroll
      for (int c = 0; c < kChNB; c++) {
        double s = x[c];
#pragma unroll
        for (int k = 0; k < c; k++) s = fma(-x[k], D[c * kChLd + k], s);
        x[c] = s / D[c * kChLd + c];
      }
#pragma unroll
      for (int c = 0; c < kChNB; c++) P[r * kChLd + c] = x[c];
    }
    // 5. t[J:J+nb]: rhs = y - L[J:J+nb, :J] t[:J]  (warp per row, lanes over k), then D^-1
    for (int c = warp; c < nb; c += kChWarps) {
      double s = 0.0;
      const double* Lrow = Lw + (size_t)(J + c) * m;
      for (int k = lane; k < J; k += 32) s = fma(Lrow[k], tv[k], s);
      s = warp_sum(s);
      if (lane == 0) {
        double y = prm.mean[w * m + J + c];
        if (prm.y_exp) y -= prm.y_exp[J + c];
        red[c] = y - s;
      }
    }
    __syncthreads();
    if (warp == 0) {
      // sequential 16-step forward substitution, lane = row
      double rhs = lane < nb ? red[lane] : 0.0;
      for (int c = 0; c < nb; c++) {
        const double tc = __shfl_sync(0xffffffffu, rhs, c) / D[c * kChLd + c];
        if (lane == c) rhs = tc;
        else if (lane > c && lane < nb) rhs = fma(-D[lane * kChLd + c], tc, rhs);
      }
      if (lane < nb) tv[J + lane] = rhs;
    }
    // 6. write the panel back (lower part only matters; rows J.., real columns)
    for (int idx = tid; idx < nrows * kChNB; idx += kChThreads) {
      const int r = idx / kChNB, c = idx - r * kChNB;
      if (c < nb) Lw[(size_t)(J + r) * m + J + c] = P[r * kChLd + c];
    }
    __syncthreads();
  }
  if (warp == 0) {
    double out;
    if (s_bad) {
      out = prm.notpd_value;
      if (lane == 0 && prm.n_notpd) atomicAdd(prm.n_notpd, 1);
    } else {
      double q2 = 0.0;
      for (int k = lane; k < m; k += 32) q2 = fma(tv[k], tv[k], q2);
      q2 = warp_sum(q2);
      out = -0.5 * q2 - logdet + prm.add_const;
    }
    if (lane == 0) prm.lp[w] = out;
  }
}

}  // namespace gpbt
