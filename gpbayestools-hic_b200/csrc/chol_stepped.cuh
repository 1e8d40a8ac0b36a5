// Kernel (c), stepped variant: the blocked Cholesky of ALL walkers advances panel by panel (32 columns),
// two launches per panel, instead of one CTA carrying one walker through its 19 dependent panel steps
// (chol_staged.cuh: latency bound, two walkers per SM, FP64 pipe 29 % busy).
//
//   diag J    grid (walkers), 4 warps: the 32x32 diagonal block  D = C[J:J+32, J:J+32] - L[J:J+32, :J] L[..]^T
//             by DMMA over a 6-stage cp.async ring; the right-hand side of the forward solve,
//             y[J:J+32] - L[J:J+32, :J] t[:J], rides on the same operand stream.  Warp 0 then factorises
//             the block as two 16x16 register factorisations (lanes 0-15: rows, pivots and multipliers
//             by shuffle; lanes 16-31: identity rows that come out as the inverse) around four
//             16x16x16 products on the tensor pipe.  Writes the factor in place, Dinv to a work buffer,
//             t[J:J+32] = Dinv rhs, accumulates log-determinant and |t|^2; the last panel emits lp[w].
//   below J   grid (64- or 32-row tiles below the block, walkers):
//             rows <- (C[rows, J:J+32] - L[rows, :J] L[J:J+32, :J]^T) Dinv^T.
//             The C tile is loaded straight into the accumulators before the operand stream starts
//             (one memory round trip per CTA, not two), the product runs with a negated A fragment, and
//             the triangular solve is a second DMMA pass over the tile parked in shared memory.
// The m^3/3 term is in `below`: thousands of independent CTAs.  The factor is written in place over the
// lower triangle of cov, as in the other variants.  Needs m even and 16-byte aligned matrices
// (cp.async rows); cov_add must already be folded in.
#pragma once
#include "chol_staged.cuh"

namespace gpbt {

constexpr int kSpNB = 32;       // panel width
constexpr int kSpThreads = 128;
constexpr int kSpKC = 16;       // operand columns per stage (two 8-column sub-blocks)
constexpr int kSpStages = 3;
constexpr int kSpLd = 34;       // row stride of 32-column tiles in shared memory (16-byte rows)

struct SteppedWork {
  double* tvec;     // [N, m]       forward-solve vector t = L^-1 y
  double* dinv;     // [N, 32, 32]  inverse of the current diagonal block's factor (lower)
  double* logdet;   // [N]          sum of log(pivot)
  double* tsq;      // [N]          |t|^2 so far
  int* bad;         // [N]          nonzero: a non-positive pivot was met
};

// One warp: OUT[16][16] (+)= sign * A[16][16] * B, all tiles in shared memory with row stride kSpLd.
//   BT = true :  B(k, n) = Bm[n][k]   (OUT = A Bm^T)        BT = false:  B(k, n) = Bm[k][n]   (OUT = A Bm)
// `init` (may be null) is the starting value of OUT.  Lane (g, t) owns OUT[8mb + g][8nb + 2t .. +1].
template <bool BT>
__device__ __forceinline__ void tile16_mma(const double* A, const double* Bm, const double* init, double* OUT,
                                           double sign, int g, int t) {
  double acc[2][2][2];
#pragma unroll
  for (int mb = 0; mb < 2; mb++)
#pragma unroll
    for (int nb2 = 0; nb2 < 2; nb2++) {
      acc[mb][nb2][0] = init ? init[(8 * mb + g) * kSpLd + 8 * nb2 + 2 * t] : 0.0;
      acc[mb][nb2][1] = init ? init[(8 * mb + g) * kSpLd + 8 * nb2 + 2 * t + 1] : 0.0;
    }
#pragma unroll
  for (int s = 0; s < 4; s++) {
    double a[2], b[2];
#pragma unroll
    for (int mb = 0; mb < 2; mb++) a[mb] = sign * A[(8 * mb + g) * kSpLd + 4 * s + t];
#pragma unroll
    for (int nb2 = 0; nb2 < 2; nb2++)
      b[nb2] = BT ? Bm[(8 * nb2 + g) * kSpLd + 4 * s + t] : Bm[(4 * s + t) * kSpLd + 8 * nb2 + g];
#pragma unroll
    for (int mb = 0; mb < 2; mb++)
#pragma unroll
      for (int nb2 = 0; nb2 < 2; nb2++) dmma884(acc[mb][nb2][0], acc[mb][nb2][1], a[mb], b[nb2]);
  }
  __syncwarp();   // OUT may alias an operand: everybody has read before anybody writes
#pragma unroll
  for (int mb = 0; mb < 2; mb++)
#pragma unroll
    for (int nb2 = 0; nb2 < 2; nb2++)
      *reinterpret_cast<double2*>(&OUT[(8 * mb + g) * kSpLd + 8 * nb2 + 2 * t]) =
          make_double2(acc[mb][nb2][0], acc[mb][nb2][1]);
  __syncwarp();
}

// ---- diag ------------------------------------------------------------------------------------------
constexpr int kSpDiagStages = 6;   // the diagonal update is a short, latency-bound stream: prefetch deep
inline size_t chol_step_diag_smem_bytes(int m) {
  // the ring (6 x 4 KB; the block and its inverse live over it once the stream is done) + red + t[:J]
  static_assert(2 * kSpNB * kSpLd <= kSpDiagStages * 2 * kSpNB * 8, "D and Dinv must fit over the ring");
  return sizeof(double) * ((size_t)kSpDiagStages * 2 * kSpNB * 8 + kSpNB + (size_t)m);
}

__global__ void __launch_bounds__(kSpThreads) chol_step_diag_kernel(const CholParams prm, const SteppedWork wk, int J,
                                                                    int is_last) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);       // [stages][2][32][8]
  double* D = ring;                                           // [32][kSpLd] diagonal block -> its factor (over the ring)
  double* Dinv = D + kSpNB * kSpLd;                           // [32][kSpLd] inverse of the factor (lower)
  double* red = ring + (size_t)kSpDiagStages * 2 * kSpNB * 8; // [32] right-hand side of the t solve
  double* tvs = red + kSpNB;                                  // [J] t[:J] of this walker
  __shared__ int s_bad;
  constexpr int kStage = 2 * kSpNB * 8;
  const int m = prm.m;
  const int64_t w = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  if (prm.skip != nullptr && prm.skip[w]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int nb = min(kSpNB, m - J);
  double* Lw = prm.cov + (size_t)w * m * m;
  double* tw = wk.tvec + (size_t)w * m;
  if (tid == 0) s_bad = 0;

  // ---- D = C[J:J+32, J:J+32] - L[J:J+32, :J] L[J:J+32, :J]^T : warp = 8 rows, all 32 columns -----------
  const int nst = J / kSpKC;
  auto issue = [&](int s) {
    if (s < nst) {
      double* B = ring + (size_t)(s % kSpDiagStages) * kStage;
      const int k0 = s * kSpKC;
      for (int idx = tid; idx < kSpNB * 8; idx += kSpThreads) {
        const int r = idx >> 3, q8 = idx & 7, sub = q8 >> 2, q4 = q8 & 3;
        cp_async16(B + ((size_t)sub * kSpNB + r) * 8 + 2 * q4,
                   Lw + (size_t)min(J + r, m - 1) * m + k0 + 8 * sub + 2 * q4);
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s0 = 0; s0 < kSpDiagStages - 1; s0++) issue(s0);
  // the forward solve's right-hand side  red = y[J:J+32] - L[J:J+32, :J] t[:J]  rides on the same operand
  // stream: thread (row rr, part) takes 4 of the 16 columns of every stage
  const int rr = tid >> 2, part = tid & 3;
  for (int k = tid; k < J; k += kSpThreads) tvs[k] = tw[k];
  double yrow = 0.0, racc = 0.0;
  if (part == 0 && rr < nb) {
    yrow = prm.mean[w * m + J + rr];
    if (prm.y_exp) yrow -= prm.y_exp[J + rr];
  }
  double acc[4][2];
  {
    const int r = 8 * warp + g;
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++) {
      const int c = 8 * nbk + 2 * t;
      // (a narrow last panel is completed with identity rows / columns; the upper triangle is never used)
      double2 v = make_double2(r == c ? 1.0 : 0.0, r == c + 1 ? 1.0 : 0.0);
      if (r < nb && c < nb) v = *reinterpret_cast<const double2*>(Lw + (size_t)(J + r) * m + J + c);
      acc[nbk][0] = v.x;
      acc[nbk][1] = v.y;
    }
  }
#pragma unroll 1
  for (int s = 0; s < nst; s++) {
    cp_async_wait<kSpDiagStages - 2>();
    __syncthreads();
    issue(s + kSpDiagStages - 1);
    const double* B = ring + (size_t)(s % kSpDiagStages) * kStage;
#pragma unroll
    for (int sub = 0; sub < 2; sub++) {
      const double2 a = *reinterpret_cast<const double2*>(B + ((size_t)sub * kSpNB + 8 * warp + g) * 8 + 2 * t);
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++) {
        const double2 b = *reinterpret_cast<const double2*>(B + ((size_t)sub * kSpNB + 8 * nbk + g) * 8 + 2 * t);
        dmma884(acc[nbk][0], acc[nbk][1], -a.x, b.x);
        dmma884(acc[nbk][0], acc[nbk][1], -a.y, b.y);
      }
      const double2 lv = *reinterpret_cast<const double2*>(B + ((size_t)sub * kSpNB + rr) * 8 + 2 * part);
      const double2 tv2 = *reinterpret_cast<const double2*>(tvs + s * kSpKC + 8 * sub + 2 * part);
      racc = fma(lv.x, tv2.x, racc);
      racc = fma(lv.y, tv2.y, racc);
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // the ring is free: D / Dinv take its place
  racc += __shfl_xor_sync(0xffffffffu, racc, 1);
  racc += __shfl_xor_sync(0xffffffffu, racc, 2);
  if (part == 0 && rr < nb) red[rr] = yrow - racc;
  {
    const int r = 8 * warp + g;
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++) {
      const int c = 8 * nbk + 2 * t;
      const bool pad = r >= nb || c >= nb;   // rows / columns of a narrow last panel stay identity
      *reinterpret_cast<double2*>(&D[r * kSpLd + c]) =
          pad ? make_double2(r == c ? 1.0 : 0.0, r == c + 1 ? 1.0 : 0.0) : make_double2(acc[nbk][0], acc[nbk][1]);
    }
  }
  __syncthreads();

  if (warp == 0) {
    // ---- 32x32 block as two 16x16 register factorisations (lanes 0-15: rows of the block, lanes
    //      16-31: identity rows that come out as the inverse) around three small products:
    //        L11, I11 = chol(D11);  L21 = D21 I11^T;  L22, I22 = chol(D22 - L21 L21^T);
    //        Dinv = [[I11, 0], [-I22 L21 I11, I22]]
    //      (a monolithic 32-pivot version is ~12 k instructions of straight-line code per warp and was
    //      measured 4x slower: it does not fit the instruction cache)
    bool pd = true;
    const int r = lane & 15;
    double lsum = 0.0;
#pragma unroll 1
    for (int blk = 0; blk < 2; blk++) {
      const int o = 16 * blk;
      if (blk == 1) {
        // L21 = D21 I11^T, then D22 -= L21 L21^T  (16x16x16 products on the tensor pipe)
        tile16_mma<true>(D + 16 * kSpLd, Dinv, nullptr, D + 16 * kSpLd, 1.0, g, t);
        tile16_mma<true>(D + 16 * kSpLd, D + 16 * kSpLd, D + 16 * kSpLd + 16, D + 16 * kSpLd + 16, -1.0, g, t);
      }
      double S[16];
#pragma unroll
      for (int c = 0; c < 16; c++) S[c] = (lane < 16) ? D[(o + r) * kSpLd + o + c] : (r == c ? 1.0 : 0.0);
      double piv = 1.0;
#pragma unroll
      for (int b = 0; b < 16; b++) {
        const double d = __shfl_sync(0xffffffffu, S[b], b);
        pd = pd && (d > 0.0);
        const double inv = rsqrt(d);
        if (lane == b) piv = d;
        const double lab = (lane == b) ? d * inv : S[b] * inv;
        S[b] = lab;
#pragma unroll
        for (int c = b + 1; c < 16; c++) {
          const double lcb = __shfl_sync(0xffffffffu, lab, c);
          S[c] = fma(-lab, lcb, S[c]);
        }
      }
      if (lane < 16 && o + lane < nb) lsum += log(piv);
      __syncwarp();
      if (lane < 16) {
#pragma unroll
        for (int c = 0; c < 16; c++) D[(o + r) * kSpLd + o + c] = (c <= r) ? S[c] : 0.0;
      } else {
#pragma unroll
        for (int c = 0; c < 16; c++) Dinv[(o + c) * kSpLd + o + r] = (c >= r) ? S[c] : 0.0;
      }
      __syncwarp();
    }
    if (!pd && lane == 0) s_bad = 1;
    lsum = warp_sum(lsum);
    if (lane == 0) wk.logdet[w] = (J == 0 ? 0.0 : wk.logdet[w]) + lsum;
    // Dinv21 = -I22 (L21 I11), through the (still unused) upper-right block of Dinv as scratch; the
    // upper-right blocks of the factor and of its inverse are zero
    tile16_mma<false>(D + 16 * kSpLd, Dinv, nullptr, Dinv + 16, 1.0, g, t);                      // M = L21 I11
    tile16_mma<false>(Dinv + 16 * kSpLd + 16, Dinv + 16, nullptr, Dinv + 16 * kSpLd, -1.0, g, t);  // -I22 M
    for (int idx = lane; idx < 256; idx += 32) {
      Dinv[(idx >> 4) * kSpLd + 16 + (idx & 15)] = 0.0;
      D[(idx >> 4) * kSpLd + 16 + (idx & 15)] = 0.0;
    }
  }
  __syncthreads();
  const int bad_now = s_bad;
  if (tid == 0 && (bad_now || J == 0)) wk.bad[w] = (J == 0 ? 0 : wk.bad[w]) | bad_now;

  // ---- factor of the diagonal block back to global, Dinv to the work buffer, t[J:J+nb] = Dinv red --------
  for (int idx = tid; idx < nb * nb; idx += kSpThreads) {
    const int r = idx / nb, c = idx - r * nb;
    if (c <= r) Lw[(size_t)(J + r) * m + J + c] = D[r * kSpLd + c];
  }
  double* dw = wk.dinv + (size_t)w * kSpNB * kSpNB;
  for (int idx = tid; idx < kSpNB * kSpNB; idx += kSpThreads) dw[idx] = Dinv[(idx >> 5) * kSpLd + (idx & 31)];
  if (warp == 0) {
    double sx = 0.0;
    if (lane < nb) {
      for (int k = 0; k <= lane; k++) sx = fma(Dinv[lane * kSpLd + k], red[k], sx);
      tw[J + lane] = sx;
    }
    const double q2 = warp_sum(sx * sx);
    if (lane == 0) {
      const double tot = (J == 0 ? 0.0 : wk.tsq[w]) + q2;
      wk.tsq[w] = tot;
      if (is_last) {
        const int bad = (J == 0 ? 0 : wk.bad[w]) | bad_now;
        if (bad) {
          prm.lp[w] = prm.notpd_value;
          if (prm.n_notpd) atomicAdd(prm.n_notpd, 1);
        } else {
          prm.lp[w] = -0.5 * tot - 0.5 * wk.logdet[w] + prm.add_const;
        }
      }
    }
  }
}

// ---- below -----------------------------------------------------------------------------------------
// MB = m8 row blocks per warp: a CTA covers 32 MB rows.  64-row tiles amortise the B-operand loads best;
// 32-row tiles are used for the panels where they cut the padding of the row range (e.g. 76 rows left:
// 96 instead of 128 computed).
template <int MB>
constexpr size_t chol_step_below_smem_bytes() {
  // the operand ring (the finished tile is parked over it for the triangular solve) + Dinv
  return sizeof(double) * ((size_t)kSpStages * 2 * (32 * MB + kSpNB) * 8 + (size_t)kSpNB * kSpLd);
}

template <int MB>
__global__ void __launch_bounds__(kSpThreads) chol_step_below_kernel(const CholParams prm, const SteppedWork wk, int J) {
  constexpr int kRows = 32 * MB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);   // [stages][ A: [2][64][8] | B: [2][32][8] ]
  constexpr int kStage = 2 * (kRows + kSpNB) * 8;
  double* Dv = ring + (size_t)kSpStages * kStage;        // [32][kSpLd] Dinv of this walker
  double* T = ring;                                      // [kRows][kSpLd] the tile, after the stream is done
  static_assert(kRows * kSpLd <= kSpStages * kStage, "tile must fit over the ring");
  const int m = prm.m;
  const int64_t w = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  if (prm.skip != nullptr && prm.skip[w]) return;
  const int row0 = J + kSpNB + blockIdx.x * kRows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* Lw = prm.cov + (size_t)w * m * m;
  const int nst = J / kSpKC;

  // Dinv first (its own cp.async group, the oldest), then the operand stream
  {
    const double* dw = wk.dinv + (size_t)w * kSpNB * kSpNB;
    for (int idx = tid; idx < kSpNB * 16; idx += kSpThreads) {
      const int r = idx >> 4, c2 = idx & 15;
      cp_async16(Dv + r * kSpLd + 2 * c2, dw + r * kSpNB + 2 * c2);
    }
    cp_async_commit();
  }
  auto issue = [&](int s) {
    if (s < nst) {
      double* A = ring + (size_t)(s % kSpStages) * kStage;
      double* B = A + 2 * kRows * 8;
      const int k0 = s * kSpKC;
      for (int idx = tid; idx < kRows * 8; idx += kSpThreads) {
        const int r = idx >> 3, q8 = idx & 7, sub = q8 >> 2, q4 = q8 & 3;
        cp_async16(A + ((size_t)sub * kRows + r) * 8 + 2 * q4,
                   Lw + (size_t)min(row0 + r, m - 1) * m + k0 + 8 * sub + 2 * q4);
      }
      for (int idx = tid; idx < kSpNB * 8; idx += kSpThreads) {
        const int r = idx >> 3, q8 = idx & 7, sub = q8 >> 2, q4 = q8 & 3;
        cp_async16(B + ((size_t)sub * kSpNB + r) * 8 + 2 * q4, Lw + (size_t)(J + r) * m + k0 + 8 * sub + 2 * q4);
      }
    }
    cp_async_commit();
  };
  issue(0);
  issue(1);

  // the C tile goes straight into the accumulators (these loads overlap the first stages)
  double acc[MB][4][2];
#pragma unroll
  for (int mb = 0; mb < MB; mb++) {
    const int r = min(row0 + 8 * MB * warp + 8 * mb + g, m - 1);
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++) {
      const double2 v = *reinterpret_cast<const double2*>(Lw + (size_t)r * m + J + 8 * nbk + 2 * t);
      acc[mb][nbk][0] = v.x;
      acc[mb][nbk][1] = v.y;
    }
  }
#pragma unroll 1
  for (int s = 0; s < nst; s++) {
    cp_async_wait<1>();
    __syncthreads();
    issue(s + 2);
    const double* A = ring + (size_t)(s % kSpStages) * kStage;
    const double* B = A + 2 * kRows * 8;
#pragma unroll
    for (int sub = 0; sub < 2; sub++) {
      // logical k slot t of step {0,1} is the sub-block column 2t + {0,1} (same permutation for A and B)
      double2 b[4], a[MB];
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++)
        b[nbk] = *reinterpret_cast<const double2*>(B + ((size_t)sub * kSpNB + 8 * nbk + g) * 8 + 2 * t);
#pragma unroll
      for (int mb = 0; mb < MB; mb++)
        a[mb] = *reinterpret_cast<const double2*>(A + ((size_t)sub * kRows + 8 * MB * warp + 8 * mb + g) * 8 + 2 * t);
#pragma unroll
      for (int mb = 0; mb < MB; mb++)
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++) {
          dmma884(acc[mb][nbk][0], acc[mb][nbk][1], -a[mb].x, b[nbk].x);
          dmma884(acc[mb][nbk][0], acc[mb][nbk][1], -a[mb].y, b[nbk].y);
        }
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // every warp is done with the ring; Dinv has landed
  // park the updated tile, then rows <- rows * Dinv^T (each warp reads back only its own 16 rows)
#pragma unroll
  for (int mb = 0; mb < MB; mb++)
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++)
      *reinterpret_cast<double2*>(&T[(8 * MB * warp + 8 * mb + g) * kSpLd + 8 * nbk + 2 * t]) =
          make_double2(acc[mb][nbk][0], acc[mb][nbk][1]);
  __syncwarp();
#pragma unroll
  for (int mb = 0; mb < MB; mb++) {
    double2 a[4];
#pragma unroll
    for (int kk = 0; kk < 4; kk++)
      a[kk] = *reinterpret_cast<const double2*>(&T[(8 * MB * warp + 8 * mb + g) * kSpLd + 8 * kk + 2 * t]);
    double acc2[4][2];
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++) acc2[nbk][0] = acc2[nbk][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < 4; kk++)
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++) {
        const double2 b = *reinterpret_cast<const double2*>(&Dv[(8 * nbk + g) * kSpLd + 8 * kk + 2 * t]);
        dmma884(acc2[nbk][0], acc2[nbk][1], a[kk].x, b.x);
        dmma884(acc2[nbk][0], acc2[nbk][1], a[kk].y, b.y);
      }
    const int r = row0 + 8 * MB * warp + 8 * mb + g;
    if (r < m) {
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++)
        *reinterpret_cast<double2*>(Lw + (size_t)r * m + J + 8 * nbk + 2 * t) = make_double2(acc2[nbk][0], acc2[nbk][1]);
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
