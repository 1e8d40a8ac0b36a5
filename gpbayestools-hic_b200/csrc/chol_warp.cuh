// Kernel (c), warp-per-walker variant: the same left-looking blocked Cholesky + fused forward solve
// and log-determinant as chol_loglike.cuh (reference: src/mcmc.py:23-65, 288-293), but one WARP owns
// one walker and the panel is 8 columns wide.
//
// Why: a walker's factorisation is a chain of ~m/8 dependent panel steps (load, update, 8x8
// diagonal block, solve, store).  With a whole CTA per walker every step ends in a block barrier
// and an SM holds two walkers; the tensor pipe idles while those two wait on L2/HBM or on the
// sequential diagonal block.  Here nothing synchronises beyond __syncwarp, a CTA is a single
// warp with a 19.5 KB panel, nine walkers are resident per SM and their steps interleave freely.
//
//   panel   P[(m-J) x 8] in shared memory, XOR-swizzled so that the DMMA fragment accesses of the
//           solve step and the 16-byte row accesses of load/store are bank-conflict free
//   update  P -= L[J:, :J] L[J:J+8, :J]^T on the tensor pipe (DMMA.8x8x4), eight m8 row blocks
//           per pass, A/B fragments from global/L2 through two ping-pong register buffers
//   8x8     Cholesky in registers (lane = row, shuffles); eight identity rows ride along in lanes
//           8..15 and end as the inverse
//   solve   P[8:, :] <- P[8:, :] Dinv^T as DMMA;  t[J:J+8] = Dinv (y - L t)
#pragma once
#include "chol_loglike.cuh"

namespace gpbt {

constexpr int kCwNB = 8;
constexpr int kCwPass = 8;    // m8 row blocks per update pass
constexpr int kCwDld = 12;    // row stride of the 8x8 inverse (conflict-free B fragments)

__host__ __device__ inline int cw_rows_pad(int m) { return (int)round_up(m, 8); }
inline size_t chol_warp_smem_bytes(int m) {
  return sizeof(double) * ((size_t)cw_rows_pad(m) * kCwNB + (size_t)cw_rows_pad(m) + 8 * kCwDld + 8 * 8 + 8);
}

// swizzled position of panel element (r, c), c < 8
__device__ __forceinline__ int cw_at(int r, int c) { return r * kCwNB + (c ^ (((r >> 1) & 1) << 2)); }

__global__ void __launch_bounds__(32, 9) chol_warp_kernel(const CholParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = prm.m;
  const int rows_pad = cw_rows_pad(m);
  double* P = reinterpret_cast<double*>(smem_raw);  // [rows_pad][8] swizzled
  double* tv = P + (size_t)rows_pad * kCwNB;        // [rows_pad] forward-solve vector
  double* Dinv = tv + rows_pad;                     // [8][12]
  double* Dm = Dinv + 8 * kCwDld;                   // [8][8] factor of the diagonal block
  double* red = Dm + 64;                            // [8]

  const int64_t w = blockIdx.x;
  if (prm.skip != nullptr && prm.skip[w]) return;
  const int lane = threadIdx.x;
  const int g = lane >> 2, t = lane & 3;
  double* Lw = prm.cov + (size_t)w * m * m;
  const bool aligned = (m % 2 == 0) && ((reinterpret_cast<uintptr_t>(Lw) & 15) == 0) &&
                       (prm.cov_add == nullptr || (reinterpret_cast<uintptr_t>(prm.cov_add) & 15) == 0);
  double logpiv = 0.0;
  bool pd_all = true;

  for (int J = 0; J < m; J += kCwNB) {
    const int nrows = m - J, nb = min(kCwNB, nrows), nblk = (nrows + 7) >> 3;
    // ---- 1. panel load: lane (g, t) takes row 8i + g, columns 2t, 2t+1 ---------------------------
#pragma unroll 4
    for (int i = 0; i < nblk; i++) {
      const int r = 8 * i + g, c = 2 * t;
      double2 v = make_double2(0.0, 0.0);
      if (r < nrows) {
        const double* src = Lw + (size_t)(J + r) * m + J + c;
        if (c + 1 < nb) {
          v = load2(src, aligned);
          if (prm.cov_add) {
            const double2 a = load2(prm.cov_add + (size_t)(J + r) * m + J + c, aligned);
            v.x += a.x; v.y += a.y;
          }
        } else if (c < nb) {
          v.x = src[0] + (prm.cov_add ? prm.cov_add[(size_t)(J + r) * m + J + c] : 0.0);
        }
      }
      if (c >= nb && r == c) v.x = 1.0;          // identity padding of a narrow last panel
      if (c + 1 >= nb && r == c + 1) v.y = 1.0;
      *reinterpret_cast<double2*>(&P[cw_at(r, c)]) = v;
    }
    if (J + kCwNB < m) {  // pull the next panel towards L2 while this one is processed
      for (int r = lane; r < nrows - kCwNB; r += 32)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(Lw + (size_t)(J + kCwNB + r) * m + J + kCwNB));
    }
    __syncwarp();

    // ---- 2. update with the columns already factorised -------------------------------------------
    if (J > 0) {
      const double* brow = Lw + (size_t)(J + g) * m + 2 * t;
      for (int b0 = 0; b0 < nblk; b0 += kCwPass) {
        double acc[kCwPass][2];
        const double* arow[kCwPass];
        bool act[kCwPass];
#pragma unroll
        for (int i = 0; i < kCwPass; i++) {
          acc[i][0] = acc[i][1] = 0.0;
          act[i] = b0 + i < nblk;
          arow[i] = Lw + (size_t)min(J + 8 * (b0 + i) + g, m - 1) * m + 2 * t;
        }
        double2 a0[kCwPass], a1[kCwPass], bb0, bb1;
        auto fetch = [&](double2 (&a)[kCwPass], double2& b, int k0) {
          b = load2(brow + k0, aligned);
#pragma unroll
          for (int i = 0; i < kCwPass; i++)
            if (act[i]) a[i] = load2(arow[i] + k0, aligned);
        };
        auto mma = [&](const double2 (&a)[kCwPass], const double2& b) {
#pragma unroll
          for (int i = 0; i < kCwPass; i++)
            if (act[i]) dmma884(acc[i][0], acc[i][1], a[i].x, b.x);
#pragma unroll
          for (int i = 0; i < kCwPass; i++)
            if (act[i]) dmma884(acc[i][0], acc[i][1], a[i].y, b.y);
        };
        // k in chunks of 8 (J is a multiple of 8); logical k slot t of step s is column k0 + 2t + s
        fetch(a0, bb0, 0);
#pragma unroll 1
        for (int k0 = 0; k0 < J; k0 += 16) {
          if (k0 + 8 < J) fetch(a1, bb1, k0 + 8);
          mma(a0, bb0);
          if (k0 + 8 < J) {
            if (k0 + 16 < J) fetch(a0, bb0, k0 + 16);
            mma(a1, bb1);
          }
        }
#pragma unroll
        for (int i = 0; i < kCwPass; i++) {
          const int r = 8 * (b0 + i) + g;
          if (act[i] && r < nrows) {
            double2* dst = reinterpret_cast<double2*>(&P[cw_at(r, 2 * t)]);
            double2 v = *dst;
            if (2 * t < nb) v.x -= acc[i][0];
            if (2 * t + 1 < nb) v.y -= acc[i][1];
            *dst = v;
          }
        }
      }
      // right-hand side of the t solve: red[c] = y[J+c] - L[J+c, :J] . t[:J]
      double dot[kCwNB];
#pragma unroll
      for (int c = 0; c < kCwNB; c++) dot[c] = 0.0;
      for (int k = lane; k < J; k += 32) {
        const double tk = tv[k];
#pragma unroll
        for (int c = 0; c < kCwNB; c++)
          if (c < nb) dot[c] = fma(Lw[(size_t)(J + c) * m + k], tk, dot[c]);
      }
#pragma unroll
      for (int c = 0; c < kCwNB; c++) {
        const double s = warp_sum(dot[c]);
        if (lane == c) red[c] = s;
      }
    } else if (lane < kCwNB) {
      red[lane] = 0.0;
    }
    __syncwarp();
    if (lane < nb) {
      double y = prm.mean[w * m + J + lane];
      if (prm.y_exp) y -= prm.y_exp[J + lane];
      red[lane] = y - red[lane];
    } else if (lane < kCwNB) {
      red[lane] = 0.0;
    }

    // ---- 3. 8x8 diagonal block in registers; lanes 8..15 carry identity rows -> inverse ------------
    {
      const int r = lane & 7;
      double S[kCwNB];
#pragma unroll
      for (int c = 0; c < kCwNB; c++) S[c] = ((lane & 8) == 0) ? P[cw_at(r, c)] : (r == c ? 1.0 : 0.0);
      double piv = 1.0;
#pragma unroll
      for (int b = 0; b < kCwNB; b++) {
        const double d = __shfl_sync(0xffffffffu, S[b], b);
        pd_all = pd_all && (d > 0.0);
        const double inv = rsqrt(d);
        if (lane == b) piv = d;
        const double lab = (lane == b) ? d * inv : S[b] * inv;
        S[b] = lab;
#pragma unroll
        for (int c = b + 1; c < kCwNB; c++) {
          const double lcb = __shfl_sync(0xffffffffu, lab, c);
          S[c] = fma(-lab, lcb, S[c]);
        }
      }
      if (lane < kCwNB) {
        logpiv += log(piv);
#pragma unroll
        for (int c = 0; c < kCwNB; c++) Dm[r * 8 + c] = (c <= r) ? S[c] : 0.0;
      } else if (lane < 2 * kCwNB) {
        // lane 8+j holds (L^-T)[j][c] = (L^-1)[c][j], c >= j
#pragma unroll
        for (int c = 0; c < kCwNB; c++) Dinv[c * kCwDld + r] = (c >= r) ? S[c] : 0.0;
      }
    }
    __syncwarp();

    // ---- 4. solve: rows 8.. <- rows * Dinv^T (DMMA); t[J:J+8] = Dinv * red; factor into rows 0..7 ---
    {
      double bfr[2];
#pragma unroll
      for (int s = 0; s < 2; s++) bfr[s] = Dinv[g * kCwDld + 4 * s + t];  // B[k = 4s+t][n = g] = Dinv[g][4s+t]
      for (int blk = 1; blk < nblk; blk++) {
        const int r = 8 * blk + g;
        double c0 = 0.0, c1 = 0.0;
        dmma884(c0, c1, P[cw_at(r, t)], bfr[0]);
        dmma884(c0, c1, P[cw_at(r, 4 + t)], bfr[1]);
        __syncwarp();  // all fragment reads of this block precede its overwrite
        *reinterpret_cast<double2*>(&P[cw_at(r, 2 * t)]) = make_double2(c0, c1);
      }
      if (lane < nb) {
        double sx = 0.0;
        for (int k = 0; k <= lane; k++) sx = fma(Dinv[lane * kCwDld + k], red[k], sx);
        tv[J + lane] = sx;
      }
      for (int idx = lane; idx < 64; idx += 32) P[cw_at(idx >> 3, idx & 7)] = Dm[idx];
    }
    __syncwarp();

    // ---- 5. panel store ---------------------------------------------------------------------------
#pragma unroll 4
    for (int i = 0; i < nblk; i++) {
      const int r = 8 * i + g, c = 2 * t;
      if (r < nrows && c < nb) {
        const double2 v = *reinterpret_cast<const double2*>(&P[cw_at(r, c)]);
        double* dst = Lw + (size_t)(J + r) * m + J + c;
        if (aligned && c + 1 < nb) {
          *reinterpret_cast<double2*>(dst) = v;
        } else {
          dst[0] = v.x;
          if (c + 1 < nb) dst[1] = v.y;
        }
      }
    }
    __syncwarp();
  }

  double q2 = 0.0;
  for (int k = lane; k < m; k += 32) q2 = fma(tv[k], tv[k], q2);
  q2 = warp_sum(q2);
  const double ld2 = warp_sum(lane < kCwNB ? logpiv : 0.0);  // sum log d = 2 sum log l
  if (lane == 0) {
    if (!pd_all) {
      prm.lp[w] = prm.notpd_value;
      if (prm.n_notpd) atomicAdd(prm.n_notpd, 1);
    } else {
      prm.lp[w] = -0.5 * q2 - 0.5 * ld2 + prm.add_const;
    }
  }
}

}  // namespace gpbt
