// Kernel (c), stepped variant -- EXPERIMENTAL, opt-in with GPBT_CHOL=batch, not the default: measured
// 1.33 ms per 1024 walkers at m = 300 against 1.14 ms for chol_staged.cuh (profiles/
// r01_chol_stepped_launches.csv).  Kept because it passes the full parity suite and documents where
// the time goes when the latency chain is cut differently.
//
// The blocked Cholesky of ALL walkers advances panel by panel, two
// kernels per 32-column panel, instead of one CTA carrying one walker through its 19 dependent panel
// steps (chol_staged.cuh: latency bound, two walkers per SM, FP64 pipe 29 % busy).
//
//   update J   grid (row tiles of 64, walkers):  C[rows, J:J+32] -= L[rows, :J] L[J:J+32, :J]^T
//              -- the m^3/3 term -- as DMMA.8x8x4 over a 3-stage cp.async ring of 16-column operand
//              slices; thousands of independent CTAs, so the tensor pipe and HBM stay busy.
//   factor J   grid (walkers): 32x32 diagonal block factorised in registers by one warp (lane = row,
//              pivots / multipliers by shuffle), its inverse by forward substitution (lane = column);
//              rows below <- rows * Dinv^T (DMMA, 8 rows per warp pass); the forward solve
//              t[J:J+32] = Dinv (y - L[J:J+32, :J] t[:J]) and the log-determinant ride along; the
//              last panel's launch also emits lp[w] = -1/2 |t|^2 - 1/2 sum log(pivot) + const.
// The factor is written in place over the lower triangle of cov, as in the other variants.  Needs
// m even and 16-byte aligned matrices (cp.async rows); cov_add must already be folded in.
#pragma once
#include "chol_staged.cuh"

namespace gpbt {

constexpr int kSpNB = 32;       // panel width
constexpr int kSpRows = 64;     // rows per update CTA (4 warps x 16)
constexpr int kSpThreads = 128;
constexpr int kSpKC = 16;       // operand columns per stage (two 8-column sub-blocks)
constexpr int kSpStages = 3;
constexpr int kSpLd = 34;       // row stride of the diagonal block / its inverse in shared memory (16-byte rows)

struct SteppedWork {
  double* tvec;     // [N, m]  forward-solve vector t = L^-1 y
  double* logdet;   // [N]     sum of log(pivot)
  int* bad;         // [N]     nonzero: a non-positive pivot was met
};

constexpr size_t chol_step_update_smem_bytes() {
  return sizeof(double) * (size_t)kSpStages * 2 * (kSpRows + kSpNB) * 8;
}

__global__ void __launch_bounds__(kSpThreads) chol_step_update_kernel(const CholParams prm, int J) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);   // [stages][ A: [2][64][8] | B: [2][32][8] ]
  constexpr int kStage = 2 * (kSpRows + kSpNB) * 8;
  const int m = prm.m;
  const int64_t w = blockIdx.y;
  if (prm.skip != nullptr && prm.skip[w]) return;
  const int row0 = J + blockIdx.x * kSpRows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* Lw = prm.cov + (size_t)w * m * m;
  const int nst = J / kSpKC;

  auto issue = [&](int s) {
    if (s < nst) {
      double* A = ring + (size_t)(s % kSpStages) * kStage;
      double* B = A + 2 * kSpRows * 8;
      const int k0 = s * kSpKC;
      for (int idx = tid; idx < kSpRows * 8; idx += kSpThreads) {
        const int r = idx >> 3, q8 = idx & 7, sub = q8 >> 2, q4 = q8 & 3;
        cp_async16(A + ((size_t)sub * kSpRows + r) * 8 + 2 * q4,
                   Lw + (size_t)min(row0 + r, m - 1) * m + k0 + 8 * sub + 2 * q4);
      }
      for (int idx = tid; idx < kSpNB * 8; idx += kSpThreads) {
        const int r = idx >> 3, q8 = idx & 7, sub = q8 >> 2, q4 = q8 & 3;
        cp_async16(B + ((size_t)sub * kSpNB + r) * 8 + 2 * q4,
                   Lw + (size_t)min(J + r, m - 1) * m + k0 + 8 * sub + 2 * q4);
      }
    }
    cp_async_commit();
  };

  double acc[2][4][2];
#pragma unroll
  for (int mb = 0; mb < 2; mb++)
#pragma unroll
    for (int nb = 0; nb < 4; nb++) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;

  issue(0);
  issue(1);
#pragma unroll 1
  for (int s = 0; s < nst; s++) {
    cp_async_wait<1>();
    __syncthreads();
    issue(s + 2);
    const double* A = ring + (size_t)(s % kSpStages) * kStage;
    const double* B = A + 2 * kSpRows * 8;
#pragma unroll
    for (int sub = 0; sub < 2; sub++) {
      // logical k slot t of step {0,1} is the sub-block column 2t + {0,1} (same permutation for A and B)
      double2 b[4], a[2];
#pragma unroll
      for (int nb = 0; nb < 4; nb++)
        b[nb] = *reinterpret_cast<const double2*>(B + ((size_t)sub * kSpNB + 8 * nb + g) * 8 + 2 * t);
#pragma unroll
      for (int mb = 0; mb < 2; mb++)
        a[mb] = *reinterpret_cast<const double2*>(A + ((size_t)sub * kSpRows + 16 * warp + 8 * mb + g) * 8 + 2 * t);
#pragma unroll
      for (int mb = 0; mb < 2; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
          dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].x, b[nb].x);
          dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb].y, b[nb].y);
        }
    }
  }
  cp_async_wait<0>();
  // C[r][c] -= acc, lower triangle only
#pragma unroll
  for (int mb = 0; mb < 2; mb++) {
    const int r = row0 + 16 * warp + 8 * mb + g;
    if (r >= m) continue;
#pragma unroll
    for (int nb = 0; nb < 4; nb++) {
      const int c = J + 8 * nb + 2 * t;
      if (c > r || c >= m) continue;
      double* dst = Lw + (size_t)r * m + c;
      if (c + 1 <= r) {
        double2 v = *reinterpret_cast<double2*>(dst);
        v.x -= acc[mb][nb][0];
        v.y -= acc[mb][nb][1];
        *reinterpret_cast<double2*>(dst) = v;
      } else {
        dst[0] -= acc[mb][nb][0];
      }
    }
  }
}

inline size_t chol_step_factor_smem_bytes(int m) {
  return sizeof(double) * ((size_t)2 * kSpNB * kSpLd + kSpNB + (size_t)m);
}

__global__ void __launch_bounds__(kSpThreads) chol_step_factor_kernel(const CholParams prm, const SteppedWork wk, int J,
                                                                      int is_last) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* D = reinterpret_cast<double*>(smem_raw);   // [32][kSpLd] diagonal block -> its factor
  double* Dinv = D + kSpNB * kSpLd;                  // [32][kSpLd] inverse of the factor (lower)
  double* red = Dinv + kSpNB * kSpLd;                // [32] right-hand side of the t solve
  double* tv = red + kSpNB;                          // [m] (last panel only)
  __shared__ int s_bad;
  const int m = prm.m;
  const int64_t w = blockIdx.x;
  if (prm.skip != nullptr && prm.skip[w]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = min(kSpNB, m - J);
  double* Lw = prm.cov + (size_t)w * m * m;
  double* tw = wk.tvec + (size_t)w * m;
  if (tid == 0) s_bad = 0;

  // diagonal block (lower triangle; a narrow last panel is completed with identity rows)
  for (int idx = tid; idx < kSpNB * kSpNB; idx += kSpThreads) {
    const int r = idx >> 5, c = idx & 31;
    double v = (r == c) ? 1.0 : 0.0;
    if (r < nb && c <= r) v = Lw[(size_t)(J + r) * m + J + c];
    D[r * kSpLd + c] = v;
  }
  __syncthreads();

  if (warp == 0) {
    // ---- 32x32 Cholesky in registers: lane = row ---------------------------------------------------
    double S[kSpNB];
#pragma unroll
    for (int c = 0; c < kSpNB; c++) S[c] = D[lane * kSpLd + c];
    bool pd = true;
    double piv = 1.0;
#pragma unroll
    for (int b = 0; b < kSpNB; b++) {
      const double d = __shfl_sync(0xffffffffu, S[b], b);
      pd = pd && (d > 0.0);
      const double inv = rsqrt(d);
      if (lane == b) piv = d;
      const double lab = (lane == b) ? d * inv : S[b] * inv;
      S[b] = lab;
#pragma unroll
      for (int c = b + 1; c < kSpNB; c++) {
        const double lcb = __shfl_sync(0xffffffffu, lab, c);
        S[c] = fma(-lab, lcb, S[c]);
      }
    }
    if (!pd && lane == 0) s_bad = 1;
    const double lsum = warp_sum(lane < nb ? log(piv) : 0.0);
    if (lane == 0) wk.logdet[w] = (J == 0 ? 0.0 : wk.logdet[w]) + lsum;
#pragma unroll
    for (int c = 0; c < kSpNB; c++) D[lane * kSpLd + c] = (c <= lane) ? S[c] : 0.0;
    __syncwarp();
    // ---- inverse of the factor by forward substitution: lane = column of the identity ---------------
    //      x_i = (e_i - sum_{k<i} L_ik x_k) / L_ii ; the x_k of this lane stay in registers
    double x[kSpNB];
#pragma unroll
    for (int i = 0; i < kSpNB; i++) {
      double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
      for (int k = 0; k < i; k++) sacc = fma(-D[i * kSpLd + k], x[k], sacc);
      x[i] = (i < lane) ? 0.0 : sacc / D[i * kSpLd + i];
    }
#pragma unroll
    for (int i = 0; i < kSpNB; i++) Dinv[i * kSpLd + lane] = x[i];
  } else {
    // ---- right-hand side of the t solve: red[c] = y[J+c] - L[J+c, :J] . t[:J] -------------------------
    for (int c = warp - 1; c < nb; c += kSpThreads / 32 - 1) {
      double sdot = 0.0;
      const double* Lrow = Lw + (size_t)(J + c) * m;
      for (int k = lane; k < J; k += 32) sdot = fma(Lrow[k], tw[k], sdot);
      sdot = warp_sum(sdot);
      if (lane == 0) {
        double y = prm.mean[w * m + J + c];
        if (prm.y_exp) y -= prm.y_exp[J + c];
        red[c] = y - sdot;
      }
    }
  }
  __syncthreads();
  const int bad_now = s_bad;
  if (tid == 0 && (bad_now || J == 0)) wk.bad[w] = (J == 0 ? 0 : wk.bad[w]) | bad_now;

  // ---- factor of the diagonal block back to global; t[J:J+nb] = Dinv red ------------------------------
  for (int idx = tid; idx < nb * nb; idx += kSpThreads) {
    const int r = idx / nb, c = idx - r * nb;
    if (c <= r) Lw[(size_t)(J + r) * m + J + c] = D[r * kSpLd + c];
  }
  if (tid < nb) {
    double sx = 0.0;
    for (int k = 0; k <= tid; k++) sx = fma(Dinv[tid * kSpLd + k], red[k], sx);
    tw[J + tid] = sx;
    if (is_last) tv[J + tid] = sx;
  }
  // ---- rows below the block: rows <- rows * Dinv^T on the tensor pipe, 8 rows per warp pass --------------
  //      (logical k slot t of step {0,1} is column 8kk + 2t + {0,1}, for both operands)
  if (J + kSpNB < m) {
    const int g = lane >> 2, t = lane & 3;
    double2 bfr[4][4];   // [nb][kk]: Dinv[8nb + g][8kk + 2t .. +1]
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++)
#pragma unroll
      for (int kk = 0; kk < 4; kk++)
        bfr[nbk][kk] = *reinterpret_cast<const double2*>(&Dinv[(8 * nbk + g) * kSpLd + 8 * kk + 2 * t]);
    for (int r0 = J + kSpNB + 8 * warp; r0 < m; r0 += 8 * (kSpThreads / 32)) {
      const int r = min(r0 + g, m - 1);            // rows past the end are computed, never stored
      double* prow = Lw + (size_t)r * m + J;       // (J and m even: 16-byte aligned)
      double2 a[4];
#pragma unroll
      for (int kk = 0; kk < 4; kk++) a[kk] = *reinterpret_cast<const double2*>(prow + 8 * kk + 2 * t);
      double acc2[4][2];
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++) acc2[nbk][0] = acc2[nbk][1] = 0.0;
#pragma unroll
      for (int kk = 0; kk < 4; kk++)
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++) {
          dmma884(acc2[nbk][0], acc2[nbk][1], a[kk].x, bfr[nbk][kk].x);
          dmma884(acc2[nbk][0], acc2[nbk][1], a[kk].y, bfr[nbk][kk].y);
        }
      __syncwarp();   // every lane has read its part of these rows before anyone overwrites them
      if (r0 + g < m) {
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++)
          *reinterpret_cast<double2*>(prow + 8 * nbk + 2 * t) = make_double2(acc2[nbk][0], acc2[nbk][1]);
      }
    }
  }
  if (!is_last) return;
  // ---- last panel: lp[w] ------------------------------------------------------------------------------
  for (int k = tid; k < J; k += kSpThreads) tv[k] = tw[k];
  __syncthreads();
  if (warp == 0) {
    double q2 = 0.0;
    for (int k = lane; k < m; k += 32) q2 = fma(tv[k], tv[k], q2);
    q2 = warp_sum(q2);
    if (lane == 0) {
      const int bad = wk.bad[w] | bad_now;
      if (bad) {
        prm.lp[w] = prm.notpd_value;
        if (prm.n_notpd) atomicAdd(prm.n_notpd, 1);
      } else {
        prm.lp[w] = -0.5 * q2 - 0.5 * wk.logdet[w] + prm.add_const;
      }
    }
  }
}

// cov += cov_add over the lower triangle (the stepped kernels expect finished covariances)
__global__ void chol_step_add_kernel(const CholParams prm) {
  const int m = prm.m;
  const int64_t w = blockIdx.y;
  if (prm.skip != nullptr && prm.skip[w]) return;
  double* Lw = prm.cov + (size_t)w * m * m;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < m * m; idx += gridDim.x * blockDim.x) {
    const int r = idx / m, c = idx - r * m;
    if (c <= r) Lw[idx] += __ldg(prm.cov_add + idx);
  }
}

}  // namespace gpbt
