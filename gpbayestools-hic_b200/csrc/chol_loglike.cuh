// Kernel (c): batched per-walker blocked Cholesky with the triangular solve and log-determinant
// fused, one log-likelihood per walker.
//
// Replaces list(map(mvn_loglike, dY, cov)) (reference: src/mcmc.py:23-65, 288-293):
//     y = mean - y_exp ;  C = cov + cov_add ;  U = dpotrf(C) ;  alpha = dpotrs(U, y)
//     lp = -1/2 y.alpha - sum(log(diag(U)))
// evaluated as  t = L^-1 y,  lp = -1/2 |t|^2 - sum(log(diag(L)))  with L = U^T.
//
// One CTA (8 warps) per walker, left-looking blocked factorisation with panel width 16.  The
// factor is written in place over the lower triangle of cov (HBM/L2 resident: a 300x300 FP64
// matrix is 720 KB, more than one SM's shared memory); only the current (m-J) x 16 panel lives in
// shared memory.
//   update   P -= L[J:, :J] L[J:J+16, :J]^T  -- the m^3/3 term -- on the FP64 tensor pipe
//            (DMMA.8x8x4); fragments come straight from global/L2 into ping-pong registers one
//            8-column chunk ahead (rows of L are private to one warp; the 16 B-operand rows are
//            shared through L1)
//   diagonal 16x16 block: warp 0 alone updates it and factorises it in registers (lane = row, pivots
//            and multipliers by shuffle; 16 identity rows ride along and end as the inverse) while
//            the other 7 warps run the bulk of the update
//   solve    P[16:, :] <- P[16:, :] Dinv^T as a DMMA product; t[J:J+16] = Dinv (y - L t) rides along
// Two CTAs are resident per SM so one walker's sequential steps overlap the other's tensor work.
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

constexpr int kChThreads = 256;
constexpr int kChWarps = kChThreads / 32;
constexpr int kChNB = 16;   // panel width
constexpr int kChLd = 20;   // panel / diagonal-block row stride in shared memory (conflict-free DMMA fragment loads)
constexpr int kChMBW = 5;   // m8 row blocks per warp per update pass (one pass covers m <= 320)

__host__ __device__ inline int chol_rows_pad(int m) {
  const int r = (int)round_up(m, 8);
  return r < kChNB ? kChNB : r;
}
inline size_t chol_smem_bytes(int m) {
  return sizeof(double) * ((size_t)chol_rows_pad(m) * kChLd + (size_t)m + 2 * kChNB * kChLd + 2 * kChNB + 8);
}

// two consecutive doubles of a row of L (16-byte load when the row stride keeps it aligned)
__device__ __forceinline__ double2 load2(const double* p, bool aligned) {
  if (aligned) return *reinterpret_cast<const double2*>(p);
  return make_double2(p[0], p[1]);
}

// Panel update for the m8 row blocks  blk = first + stride * i,  i < MBW  (blocks >= nblk are idle).
// k runs in chunks of 8: logical k slot t of step s is the actual column k0 + 2t + s (the same
// permutation for the A and B operands); chunks alternate between two register buffers.
template <int MBW>
__device__ __forceinline__ void panel_update(const double* Lw, int m, int J, int nrows, int nb, bool aligned,
                                             double* P, int first, int stride, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const int nblk = (nrows + 7) >> 3;
  double acc[MBW][2][2];
  const double* arow[MBW];
  bool act[MBW];
#pragma unroll
  for (int i = 0; i < MBW; i++) {
    acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
    const int blk = first + stride * i;
    act[i] = blk < nblk;
    // clamped: rows of the padded tail are computed but never stored
    arow[i] = Lw + (size_t)min(J + 8 * blk + g, m - 1) * m + 2 * t;
  }
  const double* brow0 = Lw + (size_t)min(J + g, m - 1) * m + 2 * t;
  const double* brow1 = Lw + (size_t)min(J + 8 + g, m - 1) * m + 2 * t;
  double2 a0[MBW], a1[MBW], b0[2], b1[2];
  auto fetch = [&](double2 (&a)[MBW], double2 (&b)[2], int k0) {
    b[0] = load2(brow0 + k0, aligned);
    b[1] = load2(brow1 + k0, aligned);
#pragma unroll
    for (int i = 0; i < MBW; i++)
      if (act[i]) a[i] = load2(arow[i] + k0, aligned);
  };
  auto mma = [&](const double2 (&a)[MBW], const double2 (&b)[2]) {
#pragma unroll
    for (int i = 0; i < MBW; i++)
      if (act[i]) {
        dmma884(acc[i][0][0], acc[i][0][1], a[i].x, b[0].x);
        dmma884(acc[i][1][0], acc[i][1][1], a[i].x, b[1].x);
        dmma884(acc[i][0][0], acc[i][0][1], a[i].y, b[0].y);
        dmma884(acc[i][1][0], acc[i][1][1], a[i].y, b[1].y);
      }
  };
  fetch(a0, b0, 0);
#pragma unroll 1
  for (int k0 = 0; k0 < J; k0 += 16) {  // J is a multiple of 16: two chunks per trip
    fetch(a1, b1, k0 + 8);
    mma(a0, b0);
    if (k0 + 16 < J) fetch(a0, b0, k0 + 16);
    mma(a1, b1);
  }
#pragma unroll
  for (int i = 0; i < MBW; i++) {
    const int r = 8 * (first + stride * i) + g;
    if (act[i] && r < nrows) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const int c = 8 * h + 2 * t;
        if (c < nb) P[r * kChLd + c] -= acc[i][h][0];
        if (c + 1 < nb) P[r * kChLd + c + 1] -= acc[i][h][1];
      }
    }
  }
}

__global__ void __launch_bounds__(kChThreads, 2) chol_loglike_kernel(const CholParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = prm.m;
  const int rows_pad = chol_rows_pad(m);
  double* P = reinterpret_cast<double*>(smem_raw);  // [rows_pad][20] current panel (rows J..m-1)
  double* tv = P + (size_t)rows_pad * kChLd;        // [m] forward-solve vector t
  double* D = tv + m;                               // [16][20] Cholesky factor of the diagonal block
  double* Dinv = D + kChNB * kChLd;                 // [16][20] its inverse (lower triangular)
  double* red = Dinv + kChNB * kChLd;               // [16] right-hand side of the t solve
  int* s_bad = reinterpret_cast<int*>(red + kChNB);

  const int64_t w = blockIdx.x;
  if (prm.skip != nullptr && prm.skip[w]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* Lw = prm.cov + (size_t)w * m * m;
  const bool aligned = (m % 2 == 0) && ((reinterpret_cast<uintptr_t>(Lw) & 15) == 0);
  if (tid == 0) *s_bad = 0;

  double logpiv = 0.0;  // warp 0, lane b: sum of log(pivot) over the panels' columns b
  for (int J = 0; J < m; J += kChNB) {
    const int nrows = m - J;           // panel rows J .. m-1
    const int nb = min(kChNB, nrows);  // real columns in this panel
    const int nblk = (nrows + 7) >> 3;
    // 1. load the panel (+ cov_add); a narrow last panel is padded with identity columns
#pragma unroll 8
    for (int idx = tid; idx < rows_pad * kChNB; idx += kChThreads) {
      const int r = idx / kChNB, c = idx - r * kChNB;
      double v = 0.0;
      if (r < nrows && c < nb) {
        v = Lw[(size_t)(J + r) * m + J + c];
        if (prm.cov_add) v += __ldg(prm.cov_add + (size_t)(J + r) * m + J + c);
      } else if (r == c) {
        v = 1.0;
      }
      P[r * kChLd + c] = v;
    }
    __syncthreads();
    // pull the next panel (rows J+16.., columns J+16..J+31 of cov, still untouched input) towards L2
    // while this one is being processed: its load is otherwise a chain of exposed HBM latencies
    if (J + kChNB < m) {
      for (int r = tid; r < nrows - kChNB; r += kChThreads) {
        const double* nxt = Lw + (size_t)(J + kChNB + r) * m + J + kChNB;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + min(kChNB, nrows - kChNB) - 1));
      }
    }
    // 2+3. Warp 0 walks the critical path: update of the two diagonal row blocks, then the 16x16
    //      Cholesky.  Warps 1..7 meanwhile update the rows below and form the right-hand side of
    //      the t solve,  red[c] = y[J+c] - L[J+c, :J] . t[:J].
    if (warp == 0) {
      if (J > 0) {
        panel_update<2>(Lw, m, J, nrows, nb, aligned, P, 0, 1, lane);
        __syncwarp();
      }
      // lanes 0..15 own the rows of the diagonal block, lanes 16..31 the rows of an identity matrix:
      // the same right-looking elimination that leaves L in the first group leaves L^-T in the
      // second (row 16+j ends as column j of L^-1), so the inverse costs no extra code.
      const int r = lane & 15;
      double S[kChNB];
#pragma unroll
      for (int c = 0; c < kChNB; c++) S[c] = (lane < 16) ? P[r * kChLd + c] : (r == c ? 1.0 : 0.0);
      bool pd = true;
      double piv = 1.0;
#pragma unroll
      for (int b = 0; b < kChNB; b++) {
        const double d = __shfl_sync(0xffffffffu, S[b], b);
        pd = pd && (d > 0.0);
        const double inv = rsqrt(d);
        if (lane == b) piv = d;
        const double lab = (lane == b) ? d * inv : S[b] * inv;  // L[r][b]  |  (L^-T)[j][b];  diagonal = sqrt(d)
        S[b] = lab;
#pragma unroll
        for (int c = b + 1; c < kChNB; c++) {
          const double lcb = __shfl_sync(0xffffffffu, lab, c);
          S[c] = fma(-lab, lcb, S[c]);
        }
      }
      if (lane < 16) logpiv += log(piv);
      if (!pd && lane == 0) *s_bad = 1;
      if (lane < 16) {
#pragma unroll
        for (int c = 0; c < kChNB; c++) D[r * kChLd + c] = (c <= r) ? S[c] : 0.0;
      } else {
        // lane 16+j holds (L^-T)[j][c] = (L^-1)[c][j] for c >= j
#pragma unroll
        for (int c = 0; c < kChNB; c++) Dinv[c * kChLd + r] = (c >= r) ? S[c] : 0.0;
      }
    } else {
      if (J > 0) {
        const int bulk = nblk - 2;  // row blocks 2 .. nblk-1 over 7 warps, kChMBW per pass
        for (int base = 0; base < bulk; base += (kChWarps - 1) * kChMBW)
          panel_update<kChMBW>(Lw, m, J, nrows, nb, aligned, P, 2 + base + (warp - 1), kChWarps - 1, lane);
      }
      for (int c = warp - 1; c < nb; c += kChWarps - 1) {
        double sdot = 0.0;
        const double* Lrow = Lw + (size_t)(J + c) * m;
        for (int k = lane; k < J; k += 32) sdot = fma(Lrow[k], tv[k], sdot);
        sdot = warp_sum(sdot);
        if (lane == 0) {
          double y = prm.mean[w * m + J + c];
          if (prm.y_exp) y -= prm.y_exp[J + c];
          red[c] = y - sdot;
        }
      }
    }
    __syncthreads();
    if (*s_bad) break;
    // 4. rows below the diagonal block: P[r, :] <- P[r, :] Dinv^T  (DMMA; blocks 2.. dealt to warps
    //    1..7, the B fragments of Dinv^T held in registers).  Warp 0 meanwhile puts the factor of the
    //    diagonal block into P's first 16 rows and finishes t[J:J+16] = Dinv * red.
    if (warp > 0) {
      double bfr[2][4];
#pragma unroll
      for (int s = 0; s < 4; s++) {
        bfr[0][s] = Dinv[(size_t)g * kChLd + 4 * s + t];
        bfr[1][s] = Dinv[(size_t)(8 + g) * kChLd + 4 * s + t];
      }
      for (int blk = 2 + (warp - 1); blk < nblk; blk += kChWarps - 1) {
        double a[4], acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        double* prow = P + (size_t)(8 * blk + g) * kChLd;
#pragma unroll
        for (int s = 0; s < 4; s++) a[s] = prow[4 * s + t];
        __syncwarp();  // every lane has read its row fragment before anyone overwrites the rows
#pragma unroll
        for (int s = 0; s < 4; s++) {
          dmma884(acc[0][0], acc[0][1], a[s], bfr[0][s]);
          dmma884(acc[1][0], acc[1][1], a[s], bfr[1][s]);
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
          prow[8 * h + 2 * t] = acc[h][0];
          prow[8 * h + 2 * t + 1] = acc[h][1];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < kChNB * kChNB / 32; i++) {
        const int idx = lane + 32 * i, rr = idx >> 4, cc = idx & 15;
        P[rr * kChLd + cc] = D[rr * kChLd + cc];
      }
      if (lane < nb) {
        double sx = 0.0;
        for (int k = 0; k <= lane; k++) sx = fma(Dinv[lane * kChLd + k], red[k], sx);
        tv[J + lane] = sx;
      }
    }
    __syncthreads();
    // 5. write the panel back (rows J.., real columns)
#pragma unroll 4
    for (int idx = tid; idx < nrows * kChNB; idx += kChThreads) {
      const int r = idx / kChNB, c = idx - r * kChNB;
      if (c < nb) Lw[(size_t)(J + r) * m + J + c] = P[r * kChLd + c];
    }
    __syncthreads();
  }
  if (warp == 0) {
    double out;
    const int bad = *s_bad;
    double q2 = 0.0;
    for (int k = lane; k < m; k += 32) q2 = fma(tv[k], tv[k], q2);
    q2 = warp_sum(q2);
    const double ld2 = warp_sum(logpiv);  // sum log d = 2 sum log l
    if (bad) {
      out = prm.notpd_value;
      if (lane == 0 && prm.n_notpd) atomicAdd(prm.n_notpd, 1);
    } else {
      out = -0.5 * q2 - 0.5 * ld2 + prm.add_const;
    }
    if (lane == 0) prm.lp[w] = out;
  }
}

}  // namespace gpbt
