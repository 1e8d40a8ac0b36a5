// Kernel (c), staged variant of the CTA-per-walker blocked Cholesky (chol_loglike.cuh) for large m.
//
// Measured on the register-fed kernel (clock64 at its barriers, m = 300): 58 % of a walker's time is
// the panel update, and that loop is latency bound -- every 16 columns cost two exposed L2 round
// trips (~1.5-2 k cycles) because registers only allow a prefetch distance of one 8-column chunk.
// Here the operand stream  L[J:, k0:k0+8]  (rows of the panel x 8 columns, 64 bytes per row) goes
// through a 3-stage shared-memory ring filled by cp.async (LDGSTS.128, all 256 threads, two stages
// ahead); each stage is consumed by every warp: B fragments from its first 16 rows, A fragments from
// the warp's own row blocks.  One __syncthreads per stage.
// Everything else (panel in shared memory, diagonal 16x16 block factorised in registers by warp 0
// with identity rows riding along for the inverse, DMMA triangular solve, fused forward solve and
// log-determinant, in-place factor in global memory) is as in chol_loglike.cuh; the panel uses a
// 16-double row stride with an XOR swizzle so that two CTAs still fit on an SM.
#pragma once
#include "chol_loglike.cuh"

namespace gpbt {

constexpr int kCsStages = 3;
constexpr int kCsCols = 8;    // columns per stage
constexpr int kCsMBW = 6;     // m8 row blocks per bulk warp (7 warps x 6 = 42 blocks: m <= 352)
constexpr int kCsMaxM = 8 * (2 + (kChWarps - 1) * kCsMBW);

inline size_t chol_staged_smem_bytes(int m) {
  const size_t rows = (size_t)chol_rows_pad(m);
  return sizeof(double) * (rows * kChNB + kCsStages * rows * kCsCols + (size_t)m + 2 * kChNB * kChLd + kChNB + 8);
}

// swizzled position of panel element (r, c), c < 16: conflict-free DMMA fragment reads (row g,
// column 4s + t) with a 16-double row stride; 16-byte column pairs stay together
__device__ __forceinline__ int cs_at(int r, int c) { return r * kChNB + (c ^ ((r & 3) << 2)); }

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(kChThreads, 2) chol_staged_kernel(const CholParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = prm.m;
  const int rows_pad = chol_rows_pad(m);
  double* P = reinterpret_cast<double*>(smem_raw);           // [rows_pad][16] swizzled panel (rows J..m-1)
  double* ring = P + (size_t)rows_pad * kChNB;               // [stages][rows_pad][8]
  double* tv = ring + (size_t)kCsStages * rows_pad * kCsCols;  // [m]
  double* D = tv + m;                                        // [16][20]
  double* Dinv = D + kChNB * kChLd;                          // [16][20]
  double* red = Dinv + kChNB * kChLd;                        // [16]
  int* s_bad = reinterpret_cast<int*>(red + kChNB);

  const int64_t w = blockIdx.x;
  if (prm.skip != nullptr && prm.skip[w]) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  double* Lw = prm.cov + (size_t)w * m * m;   // m is even and the base 16-byte aligned (checked by the launcher)
  if (tid == 0) *s_bad = 0;

  double logpiv = 0.0;
  for (int J = 0; J < m; J += kChNB) {
    const int nrows = m - J, nb = min(kChNB, nrows), nblk = (nrows + 7) >> 3;
    // ---- 1. panel load; a narrow last panel is padded with identity columns.  Without cov_add (the
    //         chain's dense path: kernel (b) already produced cov + cov_exp) the rows go straight into
    //         the swizzled panel by cp.async and overlap the first stages of the update -------------------
    const bool async_panel = prm.cov_add == nullptr;
    if (async_panel) {
      for (int idx = tid; idx < rows_pad * (kChNB / 2); idx += kChThreads) {
        const int r = idx >> 3, c = 2 * (idx & 7);
        double* dst = &P[cs_at(r, c)];
        if (r < nrows && c < nb) {   // nb is even (m and J are), so a pair is wholly inside or outside
          cp_async16(dst, Lw + (size_t)(J + r) * m + J + c);
        } else {
          dst[0] = (r == c) ? 1.0 : 0.0;
          dst[1] = (r == c + 1) ? 1.0 : 0.0;
        }
      }
      cp_async_commit();
    } else {
#pragma unroll 8
      for (int idx = tid; idx < rows_pad * kChNB; idx += kChThreads) {
        const int r = idx / kChNB, c = idx - r * kChNB;
        double v = 0.0;
        if (r < nrows && c < nb) {
          v = Lw[(size_t)(J + r) * m + J + c] + __ldg(prm.cov_add + (size_t)(J + r) * m + J + c);
        } else if (r == c) {
          v = 1.0;
        }
        P[cs_at(r, c)] = v;
      }
    }
    if (J + kChNB < m) {  // pull the next panel towards L2
      for (int r = tid; r < nrows - kChNB; r += kChThreads) {
        const double* nxt = Lw + (size_t)(J + kChNB + r) * m + J + kChNB;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + min(kChNB, nrows - kChNB) - 1));
      }
    }
    // ---- 2. update: P -= L[J:, :J] L[J:J+16, :J]^T, operands staged through the ring -----------------
    if (J > 0) {
      const int nst = J / kCsCols;
      // stage s <- columns [8s, 8s+8) of rows J .. m-1 : 64 bytes per row, 4 x 16-byte copies
      auto issue = [&](int s) {
        if (s < nst) {
          double* dst = ring + (size_t)(s % kCsStages) * rows_pad * kCsCols;
          const double* src = Lw + (size_t)J * m + (size_t)s * kCsCols;
          for (int idx = tid; idx < nrows * 4; idx += kChThreads) {
            const int r = idx >> 2, q4 = idx & 3;
            cp_async16(dst + r * kCsCols + 2 * q4, src + (size_t)r * m + 2 * q4);
          }
        }
        cp_async_commit();
      };
      // warp 0: the two diagonal row blocks; warps 1..7: blocks 2 + (warp-1) + 7 i
      const int first = warp == 0 ? 0 : 2 + (warp - 1), stride = warp == 0 ? 1 : kChWarps - 1;
      const int nmine = warp == 0 ? 2 : kCsMBW;
      double acc[kCsMBW][2][2];
      bool act[kCsMBW];
#pragma unroll
      for (int i = 0; i < kCsMBW; i++) {
        acc[i][0][0] = acc[i][0][1] = acc[i][1][0] = acc[i][1][1] = 0.0;
        act[i] = (i < nmine) && (first + stride * i < nblk);
      }
      issue(0);
      issue(1);
#pragma unroll 1
      for (int s = 0; s < nst; s++) {
        cp_async_wait<1>();   // this thread's copies of stage s have landed ...
        __syncthreads();      // ... everybody's have, and everybody is done with stage s-1
        issue(s + 2);         // refills the buffer stage s-1 lived in
        const double* st = ring + (size_t)(s % kCsStages) * rows_pad * kCsCols;
        // logical k slot t of step {0,1} is the stage column 2t + {0,1} (same permutation for A and B)
        const double2 b0 = *reinterpret_cast<const double2*>(st + (size_t)g * kCsCols + 2 * t);
        const double2 b1 = *reinterpret_cast<const double2*>(st + (size_t)(8 + g) * kCsCols + 2 * t);
#pragma unroll
        for (int i = 0; i < kCsMBW; i++)
          if (act[i]) {
            const int r = 8 * (first + stride * i) + g;   // rows beyond nrows hold stale data: never stored
            const double2 a = *reinterpret_cast<const double2*>(st + (size_t)r * kCsCols + 2 * t);
            dmma884(acc[i][0][0], acc[i][0][1], a.x, b0.x);
            dmma884(acc[i][1][0], acc[i][1][1], a.x, b1.x);
            dmma884(acc[i][0][0], acc[i][0][1], a.y, b0.y);
            dmma884(acc[i][1][0], acc[i][1][1], a.y, b1.y);
          }
      }
      cp_async_wait<0>();
#pragma unroll
      for (int i = 0; i < kCsMBW; i++) {
        const int r = 8 * (first + stride * i) + g;
        if (act[i] && r < nrows) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const int c = 8 * h + 2 * t;
            double2* dst = reinterpret_cast<double2*>(&P[cs_at(r, c)]);
            double2 v = *dst;
            if (c < nb) v.x -= acc[i][h][0];
            if (c + 1 < nb) v.y -= acc[i][h][1];
            *dst = v;
          }
        }
      }
    } else {
      cp_async_wait<0>();
      __syncthreads();   // panel complete
    }
    // ---- 3. warp 0: 16x16 diagonal block in registers (lanes 16..31: identity rows -> inverse);
    //         warps 1..7: right-hand side of the t solve, red[c] = y[J+c] - L[J+c, :J] . t[:J] -----------
    if (warp == 0) {
      __syncwarp();
      const int r = lane & 15;
      double S[kChNB];
#pragma unroll
      for (int c = 0; c < kChNB; c++) S[c] = (lane < 16) ? P[cs_at(r, c)] : (r == c ? 1.0 : 0.0);
      bool pd = true;
      double piv = 1.0;
#pragma unroll
      for (int b = 0; b < kChNB; b++) {
        const double d = __shfl_sync(0xffffffffu, S[b], b);
        pd = pd && (d > 0.0);
        const double inv = rsqrt(d);
        if (lane == b) piv = d;
        const double lab = (lane == b) ? d * inv : S[b] * inv;
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
#pragma unroll
        for (int c = 0; c < kChNB; c++) Dinv[c * kChLd + r] = (c >= r) ? S[c] : 0.0;
      }
    } else {
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
    // ---- 4. rows below the diagonal block <- rows * Dinv^T (DMMA, warps 1..7); warp 0: factor into
    //         P's first 16 rows and t[J:J+16] = Dinv * red -------------------------------------------------
    if (warp > 0) {
      double bfr[2][4];
#pragma unroll
      for (int s = 0; s < 4; s++) {
        bfr[0][s] = Dinv[(size_t)g * kChLd + 4 * s + t];
        bfr[1][s] = Dinv[(size_t)(8 + g) * kChLd + 4 * s + t];
      }
      for (int blk = 2 + (warp - 1); blk < nblk; blk += kChWarps - 1) {
        const int r = 8 * blk + g;
        double a[4], acc2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int s = 0; s < 4; s++) a[s] = P[cs_at(r, 4 * s + t)];
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 4; s++) {
          dmma884(acc2[0][0], acc2[0][1], a[s], bfr[0][s]);
          dmma884(acc2[1][0], acc2[1][1], a[s], bfr[1][s]);
        }
#pragma unroll
        for (int h = 0; h < 2; h++)
          *reinterpret_cast<double2*>(&P[cs_at(r, 8 * h + 2 * t)]) = make_double2(acc2[h][0], acc2[h][1]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < kChNB * kChNB / 32; i++) {
        const int idx = lane + 32 * i, rr = idx >> 4, cc = idx & 15;
        P[cs_at(rr, cc)] = D[rr * kChLd + cc];
      }
      if (lane < nb) {
        double sx = 0.0;
        for (int k = 0; k <= lane; k++) sx = fma(Dinv[lane * kChLd + k], red[k], sx);
        tv[J + lane] = sx;
      }
    }
    __syncthreads();
    // ---- 5. write the panel back ------------------------------------------------------------------------
#pragma unroll 4
    for (int idx = tid; idx < nrows * kChNB; idx += kChThreads) {
      const int r = idx / kChNB, c = idx - r * kChNB;
      if (c < nb) Lw[(size_t)(J + r) * m + J + c] = P[cs_at(r, c)];
    }
    __syncthreads();
  }
  if (warp == 0) {
    double out;
    const int bad = *s_bad;
    double q2 = 0.0;
    for (int k = lane; k < m; k += 32) q2 = fma(tv[k], tv[k], q2);
    q2 = warp_sum(q2);
    const double ld2 = warp_sum(logpiv);
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
