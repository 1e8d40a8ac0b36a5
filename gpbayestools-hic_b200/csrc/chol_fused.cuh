// Kernels (b) + (c) fused for the chain's dense path: the per-walker covariance
//     C_w = F + U^T diag(v_w) U,     F = blockdiag(Ctrunc_e) + expdata_cov,  U = blockdiag(A_e)
// (src/emulator.py:584-587 feeding src/mcmc.py:288-293) is never written to memory.  The blocked
// (left-looking, 32-column panels) Cholesky of ALL walkers advances panel by panel, one launch per panel,
// and every tile of C a launch needs is generated where it is consumed: accumulators start from the tile
// of F (walker independent, L2 resident), take the rank-Q term as Q/4 DMMA k-steps and then the
// -L L^T updates from the operand stream.  Only the factor is stored, and only below its diagonal
// blocks, in a PANEL-MAJOR layout
//     panel K (columns 32K .. 32K+31):  [sub-block s = 0..3][row r - (32K + 32)][8 columns],  32K + 32 <= r < Mg
// so the operands of a tile are contiguous 4 KB / 2 KB pieces that go into the shared-memory ring by TMA
// bulk copies (one elected warp, mbarrier completion) in exactly the layout the DMMA fragment loads want.
//
// Launch J (J = -32, 0, 32, ...), grid (1 + tiles below, walkers):
//   tile 0 (32 rows right below panel J = the NEXT diagonal block, look-ahead): its rows of panel J,
//     L[rows, J:J+32] = (C - L L^T) Dinv_J^T;  in the same operand stream the diagonal block
//     D = C[rows, rows] - L[rows, :J+32] L[rows, :J+32]^T and the forward-solve dot product
//     y[rows] - L[rows, :J+32] t[:J+32];  then the 32x32 factorisation (two 16x16 register
//     factorisations around tensor-pipe block products, as in chol_stepped.cuh), Dinv_{J+32},
//     t[rows], log-determinant; the last one emits lp[w].  J = -32 is the prologue (first block only).
//   tiles 1.. (64 rows each): L[rows, J:J+32] = (C - L L^T) Dinv_J^T.
// So there is ONE launch per panel (the stepped kernels of chol_stepped.cuh need two, and kernel (b)
// before them), no N m^2 covariance in HBM, and the factor traffic is half of the row-major layout's
// (only the lower blocks exist).  Launches are chained by programmatic dependent launch.
#pragma once
#include "chol_stepped.cuh"

namespace gpbt {

constexpr int kCfNB = 32;        // panel width
constexpr int kCfThreads = 128;
constexpr int kCfCtasPerSm = 4;  // by registers (128) and shared memory (54 KB)
constexpr int kCfRows = 64;      // rows of a regular tile (16 per warp)
constexpr int kCfStages = 3;     // ring depth, 16 columns per stage
constexpr int kCfLd = 40;        // row stride of the 32-column tiles (T, Dinv) in shared memory: all their fragment
                                 // accesses are 16-byte ones at [row g][8 k + 2 t]; a quarter warp (g = 0, 1; t = 0..3)
                                 // covers all eight 16-byte bank groups when the stride is 8 mod 16 doubles (34, the
                                 // stride of the 8-byte fragment loads elsewhere, made every one of them 2-way)
constexpr int kCfLdD = 34;       // row stride of the diagonal block in the factor kernel (lane = row, LDS.128 pairs)

// Everything one launch of the chain hands to the next (Dinv, raw blocks, t, log-determinants, flags) and
// what kernel (a) / the mean kernel produced in the same call is read with ld.global.cg (L2) or by the
// async copies, never through L1: with sub-batches on several streams an SM is never idle between two
// launches of a chain, its L1 is not invalidated, and a plain load can return a line cached before another
// SM rewrote it (observed: a few walkers per thousand off by O(1) in log L with two streams).
struct CholFusedParams {
  const double* __restrict__ Fp;     // packed panels of F (layout of the factor), rows >= M and columns >= M zero
  const double* __restrict__ Fd;     // [nP][32][32] diagonal blocks of F, identity padded
  const double* __restrict__ UT;     // [Mg][Qp]  U^T, zero padded
  // dense source (stand-alone mvn_loglike, chains with a no-PCA / exp-diag emulator): the covariances exist
  // in memory, cov_src [N][M][M] row major (+ cov_add [M][M] or null); then Fp / Fd / UT / z_var are unused
  // (Qp = 0) and the accumulators start from the tile of cov_src instead of the tile of F
  const double* __restrict__ cov_src;
  const double* __restrict__ cov_add;
  int dense_vec;                     // dense source: rows are 16-byte aligned and M is even
  const double* __restrict__ z_var;  // [N][ldz]  v_w
  const double* __restrict__ mean;   // [N][M]
  const double* __restrict__ y_exp;  // [M]
  const unsigned char* __restrict__ skip;   // [N] or null
  double* __restrict__ L;            // [N][Lstride] packed factor panels
  double* __restrict__ dinv;         // [N][32*32] inverse of the current diagonal block's factor
  double* __restrict__ draw;         // [N][32*32] the NEXT diagonal block, updated but not yet factorised
  double* __restrict__ tvec;         // [N][Mg]    t = L^-1 y
  double* __restrict__ logdet;       // [N]
  double* __restrict__ tsq;          // [N]
  int* __restrict__ bad;             // [N]
  double* __restrict__ lp;           // [N]
  int* __restrict__ n_notpd;
  double notpd_value, add_const;
  int64_t N, Lstride, ldz;
  int M, Mg, Q, Qp;
  int early_after;                   // CTAs (in launch order) from this one on wait for the predecessor FIRST (see the kernel)
  int flags;                         // debugging: 1 = every CTA waits for the predecessor first thing, 32 = none does
  long long* dbg;                    // null, or [16 launches][32 walkers][8 tiles][8] clock64 stamps (tuning)
};

// (C + cov_add)[r][c], [r][c + 1] of walker w from the dense source (c even); zero outside the matrix
__device__ __forceinline__ double2 cf_dense_pair(const CholFusedParams& prm, int64_t w, int r, int c) {
  const int M = prm.M;
  double2 v = make_double2(0.0, 0.0);
  if (r < M && c < M) {
    const double* row = prm.cov_src + ((size_t)w * M + r) * M;
    if (prm.dense_vec) {            // M even, 16-byte aligned bases: c + 1 < M as well
      v = __ldcg(reinterpret_cast<const double2*>(row + c));
      if (prm.cov_add != nullptr) {
        const double2 a = ldg2(prm.cov_add + (size_t)r * M + c);
        v.x += a.x;
        v.y += a.y;
      }
    } else {
      v.x = __ldcg(row + c);
      if (c + 1 < M) v.y = __ldcg(row + c + 1);
      if (prm.cov_add != nullptr) {
        v.x += __ldg(prm.cov_add + (size_t)r * M + c);
        if (c + 1 < M) v.y += __ldg(prm.cov_add + (size_t)r * M + c + 1);
      }
    }
  }
  return v;
}

__device__ __forceinline__ void cf_stamp(const CholFusedParams& prm, int J, int64_t w, int slot) {
  if (prm.dbg != nullptr && threadIdx.x == 0 && w < 32 && blockIdx.x < 8)
    prm.dbg[((((J / kCfNB + 1) & 15) * 32 + w) * 8 + blockIdx.x) * 8 + slot] = clock64();
}

// rows stored for panel K (those below its diagonal block), and the panel's offset in doubles
__host__ __device__ inline int cf_panel_rows(int Mg, int K) { return Mg - kCfNB * K - kCfNB; }
__host__ __device__ inline int64_t cf_panel_off(int Mg, int K) {
  // sum_{k<K} 32 * (Mg - 32k - 32)
  return (int64_t)kCfNB * ((int64_t)K * (Mg - kCfNB) - (int64_t)kCfNB * K * (K - 1) / 2);
}
__host__ __device__ inline int64_t cf_factor_doubles(int M) {
  const int Mg = (M + 15) / 16 * 16, nP = (Mg + kCfNB - 1) / kCfNB;
  return cf_panel_off(Mg, nP - 1) + (cf_panel_rows(Mg, nP - 1) > 0 ? (int64_t)kCfNB * cf_panel_rows(Mg, nP - 1) : 0);
}

constexpr int kCfStageDoubles = 2 * (kCfRows + kCfNB) * 8;   // A: [2][64][8], B: [2][32][8]
constexpr int kCfUbDoubles = kCfNB * 24;   // the panel's own rows of U^T (Qp <= 24), staged for the rank-Q term
inline size_t chol_fused_smem_bytes(int Mg) {
  // ring + Dinv of the current panel + t[:J+32] + U^T[J:J+32]
  return sizeof(double) * ((size_t)kCfStages * kCfStageDoubles + (size_t)kCfNB * kCfLd + (size_t)Mg + kCfUbDoubles);
}
static_assert(kCfLdD == kSpLd, "tile16_mma works on kSpLd-strided tiles");
static_assert(kCfRows * kCfLd <= kCfStages * kCfStageDoubles, "a regular tile parks its 64 rows over the ring");

// DENSE: the covariances come from memory (prm.cov_src) instead of being generated from F and U
template <bool DENSE>
__global__ void __launch_bounds__(kCfThreads, kCfCtasPerSm) chol_fused_panel_kernel(const CholFusedParams prm, int J, int64_t w_first) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* ring = reinterpret_cast<double*>(smem_raw);                    // [stages][A | B]
  double* Dv = ring + (size_t)kCfStages * kCfStageDoubles;               // [32][kCfLd] Dinv of panel J
  double* tvs = Dv + kCfNB * kCfLd;                                      // [J + 32] t so far
  double* Ubuf = tvs + prm.Mg;                                           // [32][Qp] U^T rows J .. J + 31
  __shared__ uint64_t full_bar[kCfStages], empty_bar[kCfStages];

  const int M = prm.M, Mg = prm.Mg, Qp = prm.Qp;
  const int64_t w = w_first + blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int K = J / kCfNB;                      // panel index (J = -32: prologue, K = -1)
  const bool tile0 = blockIdx.x == 0;
  const bool has_panel = J >= 0;
  // The predecessor (the factor kernel of this panel's diagonal block) writes Dinv_J and t[J:J+32]; the wait
  // for it sits in front of their first use, BEHIND the operand stream -- for the CTAs that can be resident
  // while it still runs.  A CTA that starts later got its slot from one of those, which had to pass its own
  // wait first: for it the wait returns at once, so it waits first thing and fetches Dinv_J together with
  // everything else instead of in a dependent round trip before the epilogue.
  const bool early = has_panel && ((prm.flags & 1) || (!(prm.flags & 32) &&
                     (int64_t)blockIdx.y * gridDim.x + blockIdx.x >= prm.early_after));
  auto fetch_dinv = [&]() {
    const double* dw = prm.dinv + (size_t)w * kCfNB * kCfNB;
    for (int idx = tid; idx < kCfNB * 16; idx += kCfThreads) {
      const int r = idx >> 4, c2 = idx & 15;
      cp_async16(Dv + r * kCfLd + 2 * c2, dw + r * kCfNB + 2 * c2);
    }
    cp_async_commit();
  };
  const int nst = has_panel ? J / 16 : 0;       // 16-column stages of the operand stream
  double* Lw = prm.L + (size_t)w * prm.Lstride;

  if (tid == 0) {
    for (int s = 0; s < kCfStages; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kCfThreads / 32);   // one arrival per warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  if (!tile0) {
    // =================================== regular tile: 64 rows =======================================
    const int r0 = J + 2 * kCfNB + kCfRows * (blockIdx.x - 1);      // first row
    const int Rv = min(kCfRows, Mg - r0);                            // valid rows (multiple of 16)
    const int rel0 = r0 - (J + kCfNB);                               // row within panel J's storage
    const int RkJ = cf_panel_rows(Mg, K);
    const bool active = 16 * warp < Rv;
    double acc[2][4][2];
    {
      // accumulators start from the tile of F (constant data: before the wait on the predecessor)
      const double* Fpan = prm.Fp + cf_panel_off(Mg, K);
#pragma unroll
      for (int mb = 0; mb < 2; mb++)
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++) {
          double2 v = make_double2(0.0, 0.0);
          if (active) {
            if (DENSE) v = cf_dense_pair(prm, w, r0 + 16 * warp + 8 * mb + g, J + 8 * nbk + 2 * t);
            else v = ldg2(Fpan + ((size_t)nbk * RkJ + rel0 + 16 * warp + 8 * mb + g) * 8 + 2 * t);
          }
          acc[mb][nbk][0] = v.x;
          acc[mb][nbk][1] = v.y;
        }
    }
    cf_stamp(prm, J, w, 0);
    pdl_launch_dependents();
    // No wait yet.  The predecessor is the factor kernel of this panel's diagonal block; it triggers this
    // launch only after ITS wait, i.e. once the previous panel launch has completed -- so the factor
    // panels < J, z_var and skip are final.  Only Dinv_J (and t[J:J+32] in tile 0) are still being written:
    // the wait sits in front of their first use, and the operand stream overlaps the factor kernel
    // (`early`: see the top).
    if (early) pdl_wait_prior_grids();
    if (prm.skip != nullptr && __ldcg(prm.skip + w)) return;
    __syncthreads();   // barriers initialised
    if (early) fetch_dinv();

    // operand stage s = columns 16 s .. 16 s + 15 = sub-blocks (2 s, 2 s + 1) mod 4 of panel s / 2
    auto issue = [&](int s) {
      if (s < nst && warp == 0) {
        uint64_t* bar = &full_bar[s % kCfStages];
        double* A = ring + (size_t)(s % kCfStages) * kCfStageDoubles;
        double* B = A + 2 * kCfRows * 8;
        const int Kp = s >> 1, sub0 = (s & 1) * 2;
        const int Rkp = cf_panel_rows(Mg, Kp);
        const double* pan = Lw + cf_panel_off(Mg, Kp);
        if (lane == 0) mbar_expect_tx(bar, (uint32_t)(2 * (Rv + kCfNB) * 64));
        __syncwarp();
        if (lane < 2) {
          tma_bulk_g2s(A + (size_t)lane * kCfRows * 8, pan + ((size_t)(sub0 + lane) * Rkp + (r0 - kCfNB * Kp - kCfNB)) * 8,
                       (uint32_t)(Rv * 64), bar);
        } else if (lane < 4) {
          tma_bulk_g2s(B + (size_t)(lane - 2) * kCfNB * 8,
                       pan + ((size_t)(sub0 + lane - 2) * Rkp + (J - kCfNB * Kp - kCfNB)) * 8, (uint32_t)(kCfNB * 64), bar);
        }
      }
    };
#pragma unroll
    for (int s0 = 0; s0 < kCfStages - 1; s0++) issue(s0);
    // rank-Q term of the covariance tile: acc += (v U^T[rows]) U^T[cols]^T.  The rows of U^T this tile needs
    // (64 on the A side, 32 on the B side) come into shared memory with one round of cp.async -- the third
    // ring slot and the Dinv buffer are free until the operand stream / the epilogue need them -- instead of
    // Q/4 rounds of dependent L2 loads per warp (which were most of a CTA's life in the first panels).
    const bool stage_u = !DENSE && 64 * Qp <= kCfStageDoubles && 32 * Qp <= kCfUbDoubles;
    if (stage_u) {
      double* UA = ring + 2 * kCfStageDoubles;     // [64][Qp]
      double* UB = Ubuf;                           // [32][Qp]
      const double* srcA = prm.UT + (size_t)r0 * Qp;
      const double* srcB = prm.UT + (size_t)J * Qp;
      for (int idx = tid; idx < Rv * Qp / 2; idx += kCfThreads) cp_async16(UA + 2 * idx, srcA + 2 * idx);
      for (int idx = tid; idx < kCfNB * Qp / 2; idx += kCfThreads) cp_async16(UB + 2 * idx, srcB + 2 * idx);
      cp_async_commit();
      double vk[6];
      const double* zv = prm.z_var + w * prm.ldz;
#pragma unroll
      for (int s6 = 0; s6 < 6; s6++) vk[s6] = (4 * s6 + t < prm.Q) ? __ldcg(zv + 4 * s6 + t) : 0.0;
      cp_async_wait<0>();
      __syncthreads();
      if (active) {
#pragma unroll
        for (int s6 = 0; s6 < 6; s6++)
          if (4 * s6 < Qp) {
            const int k = 4 * s6 + t;
            double a[2], b[4];
#pragma unroll
            for (int mb = 0; mb < 2; mb++) a[mb] = vk[s6] * UA[(16 * warp + 8 * mb + g) * Qp + k];
#pragma unroll
            for (int nbk = 0; nbk < 4; nbk++) b[nbk] = UB[(8 * nbk + g) * Qp + k];
#pragma unroll
            for (int mb = 0; mb < 2; mb++)
#pragma unroll
              for (int nbk = 0; nbk < 4; nbk++) dmma884(acc[mb][nbk][0], acc[mb][nbk][1], a[mb], b[nbk]);
          }
      }
    } else if (!DENSE && active) {
      const double* zv = prm.z_var + w * prm.ldz;
      for (int k0 = 0; k0 < Qp; k0 += 4) {
        const double vk = (k0 + t < prm.Q) ? __ldcg(zv + k0 + t) : 0.0;
        double a[2], b[4];
#pragma unroll
        for (int mb = 0; mb < 2; mb++) a[mb] = vk * __ldg(prm.UT + (size_t)(r0 + 16 * warp + 8 * mb + g) * Qp + k0 + t);
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++) b[nbk] = __ldg(prm.UT + (size_t)(J + 8 * nbk + g) * Qp + k0 + t);
#pragma unroll
        for (int mb = 0; mb < 2; mb++)
#pragma unroll
          for (int nbk = 0; nbk < 4; nbk++) dmma884(acc[mb][nbk][0], acc[mb][nbk][1], a[mb], b[nbk]);
      }
    }
    // No block-wide barrier in the stream: a warp that is done with a stage arrives on its "empty" barrier
    // and goes on; only warp 0 waits -- until all four have finished stage s -- before it refills the slot
    // of stage s - 1 with stage s + 2.  (Refilling as soon as everybody had finished stage s - 1, i.e. while
    // some warp was still inside stage s, corrupted about 3 walkers per million whenever kernels of another
    // stream shared the SM -- 8..41 of 4.9 M evaluations, 0 of 9.8 M with this wait, 0 of 20 M with a block
    // barrier per stage; tools/r02/fused_verify.py.  The protocol reads correct either way; the cause was not
    // found, the measured-safe order is kept.)
    if (prm.flags & 64) __syncthreads();   // (experiment: the early refill must not overtake the rank-Q term)
#pragma unroll 1
    for (int s = 0; s < nst; s++) {
      mbar_wait(&full_bar[s % kCfStages], (s / kCfStages) & 1);
      const double* A = ring + (size_t)(s % kCfStages) * kCfStageDoubles;
      const double* B = A + 2 * kCfRows * 8;
      if (active) {
#pragma unroll
        for (int sub = 0; sub < 2; sub++) {
          double2 b[4], a[2];
#pragma unroll
          for (int nbk = 0; nbk < 4; nbk++)
            b[nbk] = *reinterpret_cast<const double2*>(B + ((size_t)sub * kCfNB + 8 * nbk + g) * 8 + 2 * t);
#pragma unroll
          for (int mb = 0; mb < 2; mb++)
            a[mb] = *reinterpret_cast<const double2*>(A + ((size_t)sub * kCfRows + 16 * warp + 8 * mb + g) * 8 + 2 * t);
          // (the two k halves apart: eight independent accumulators between dependent DMMAs)
#pragma unroll
          for (int mb = 0; mb < 2; mb++)
#pragma unroll
            for (int nbk = 0; nbk < 4; nbk++) dmma884(acc[mb][nbk][0], acc[mb][nbk][1], -a[mb].x, b[nbk].x);
#pragma unroll
          for (int mb = 0; mb < 2; mb++)
#pragma unroll
            for (int nbk = 0; nbk < 4; nbk++) dmma884(acc[mb][nbk][0], acc[mb][nbk][1], -a[mb].y, b[nbk].y);
        }
      }
      if (prm.flags & 128) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s % kCfStages]);
      if (warp == 0 && s + kCfStages - 1 < nst) {
        if (!(prm.flags & 64)) mbar_wait(&empty_bar[s % kCfStages], (s / kCfStages) & 1);   // all four warps are done with stage s
        else if (s >= 1) mbar_wait(&empty_bar[(s - 1) % kCfStages], ((s - 1) / kCfStages) & 1);   // (experiment: with stage s - 1)
        issue(s + kCfStages - 1);
      }
    }
    cf_stamp(prm, J, w, 1);
    if (!early) {
      pdl_wait_prior_grids();          // Dinv_J is final
      fetch_dinv();
    }
    __syncthreads();     // every warp has left the operand stream: the ring is free
    // park the tile over the ring, rows <- rows Dinv^T (Dinv lower triangular: k blocks kk <= nbk only)
    double* T = ring;
    if (active) {
#pragma unroll
      for (int mb = 0; mb < 2; mb++)
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++)
          *reinterpret_cast<double2*>(&T[(16 * warp + 8 * mb + g) * kCfLd + 8 * nbk + 2 * t]) =
              make_double2(acc[mb][nbk][0], acc[mb][nbk][1]);
    }
    cp_async_wait<0>();
    __syncthreads();     // Dinv has landed
    cf_stamp(prm, J, w, 2);
    if (active) {
      double* Lpan = Lw + cf_panel_off(Mg, K);
#pragma unroll
      for (int mb = 0; mb < 2; mb++) {
        double2 a[4];
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
          a[kk] = *reinterpret_cast<const double2*>(&T[(16 * warp + 8 * mb + g) * kCfLd + 8 * kk + 2 * t]);
        double acc2[4][2];
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++) acc2[nbk][0] = acc2[nbk][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          double2 b[4];
#pragma unroll
          for (int nbk = kk; nbk < 4; nbk++) b[nbk] = *reinterpret_cast<const double2*>(&Dv[(8 * nbk + g) * kCfLd + 8 * kk + 2 * t]);
#pragma unroll
          for (int nbk = kk; nbk < 4; nbk++) dmma884(acc2[nbk][0], acc2[nbk][1], a[kk].x, b[nbk].x);
#pragma unroll
          for (int nbk = kk; nbk < 4; nbk++) dmma884(acc2[nbk][0], acc2[nbk][1], a[kk].y, b[nbk].y);
        }
        const int rel = rel0 + 16 * warp + 8 * mb + g;
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++)
          *reinterpret_cast<double2*>(Lpan + ((size_t)nbk * RkJ + rel) * 8 + 2 * t) = make_double2(acc2[nbk][0], acc2[nbk][1]);
      }
    }
    cf_stamp(prm, J, w, 3);
    return;
  }

  // ======================= tile 0: rows Jd .. Jd + 31, the next diagonal block =========================
  const int Jd = J + kCfNB;                       // first row / column of the diagonal block handled here
  // rows 8 wr .. 8 wr + 7 belong to this warp.  The diagonal block costs warp wr (wr + 1) of the 4 column
  // blocks, and warp i of every CTA sits on SM sub-partition i: the roles rotate with the walker, so that the
  // heavy rows do not all land on the same FP64 pipe.
  const int wr = (warp + (int)w) & 3;
  const int nb = min(kCfNB, M - Jd);              // its true size (identity padded to 32)
  const int Rv0 = min(kCfNB, Mg - Jd);            // rows that exist in the storage (16 or 32)
  const bool is_last = Jd + kCfNB >= M;
  const int RkJ = has_panel ? cf_panel_rows(Mg, K) : 0;
  cf_stamp(prm, J, w, 0);
  pdl_launch_dependents();
  // the prologue launch follows kernels outside the chain (z_var, mean, skip, a dense source): it waits
  // before it reads anything
  if (!has_panel || early) pdl_wait_prior_grids();
  double accp[4][2], accd[4][2];                  // warp: rows 8 warp .. 8 warp + 7, 32 columns each
  {
    const double* Fpan = prm.Fp + (has_panel ? cf_panel_off(Mg, K) : 0);
    const double* Fdb = prm.Fd + (size_t)(K + 1) * kCfNB * kCfNB;
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++) {
      double2 v = make_double2(0.0, 0.0), d;
      if (DENSE) {
        if (has_panel) v = cf_dense_pair(prm, w, Jd + 8 * wr + g, J + 8 * nbk + 2 * t);
        d = cf_dense_pair(prm, w, Jd + 8 * wr + g, Jd + 8 * nbk + 2 * t);   // (padding: identity, set below)
      } else {
        if (has_panel && 8 * wr < Rv0) v = ldg2(Fpan + ((size_t)nbk * RkJ + 8 * wr + g) * 8 + 2 * t);
        d = ldg2(Fdb + (size_t)(8 * wr + g) * kCfNB + 8 * nbk + 2 * t);
      }
      accp[nbk][0] = v.x;
      accp[nbk][1] = v.y;
      accd[nbk][0] = d.x;
      accd[nbk][1] = d.y;
    }
  }
  if (prm.skip != nullptr && __ldcg(prm.skip + w)) return;
  __syncthreads();
  if (early) fetch_dinv();

  auto issue0 = [&](int s) {
    if (s < nst && warp == 0) {
      uint64_t* bar = &full_bar[s % kCfStages];
      double* A = ring + (size_t)(s % kCfStages) * kCfStageDoubles;   // [2][32][8] rows Jd.. of the earlier panel
      double* B = A + 2 * kCfRows * 8;                                 // [2][32][8] rows J..
      const int Kp = s >> 1, sub0 = (s & 1) * 2;
      const int Rkp = cf_panel_rows(Mg, Kp);
      const double* pan = Lw + cf_panel_off(Mg, Kp);
      if (lane == 0) mbar_expect_tx(bar, (uint32_t)(2 * (Rv0 + kCfNB) * 64));
      __syncwarp();
      if (lane < 2) {
        tma_bulk_g2s(A + (size_t)lane * kCfNB * 8, pan + ((size_t)(sub0 + lane) * Rkp + (Jd - kCfNB * Kp - kCfNB)) * 8,
                     (uint32_t)(Rv0 * 64), bar);
      } else if (lane < 4) {
        tma_bulk_g2s(B + (size_t)(lane - 2) * kCfNB * 8,
                     pan + ((size_t)(sub0 + lane - 2) * Rkp + (J - kCfNB * Kp - kCfNB)) * 8, (uint32_t)(kCfNB * 64), bar);
      }
    }
  };
#pragma unroll
  for (int s0 = 0; s0 < kCfStages - 1; s0++) issue0(s0);
  if (has_panel) {
    // t[:J] is final (see the note on the wait in the regular tile); t[J:J+32] follows after the wait
    const double* tw = prm.tvec + (size_t)w * Mg;
    for (int k = tid; k < J; k += kCfThreads) tvs[k] = __ldcg(tw + k);
  }
  // the forward solve's right-hand side  y[rows] - L[rows, :Jd] t[:Jd]: thread (row rr, part) takes a
  // quarter of the columns of every stage
  const int rr = tid >> 2, part = tid & 3;
  double yrow = 0.0, racc = 0.0;
  if (part == 0 && rr < nb) {
    yrow = __ldcg(prm.mean + w * M + Jd + rr);
    if (prm.y_exp) yrow -= prm.y_exp[Jd + rr];
  }
  const bool stage_u0 = !DENSE && 64 * Qp <= kCfStageDoubles;
  if (stage_u0) {
    // rank-Q term for both tiles, U^T rows through shared memory (see the regular tile): [32][Qp] rows Jd..
    // (A side and the diagonal block's B side), [32][Qp] rows J.. (the panel's B side)
    double* UA = ring + 2 * kCfStageDoubles;
    double* UB = UA + kCfNB * Qp;
    const double* srcA = prm.UT + (size_t)Jd * Qp;
    for (int idx = tid; idx < Rv0 * Qp / 2; idx += kCfThreads) cp_async16(UA + 2 * idx, srcA + 2 * idx);
    for (int idx = Rv0 * Qp + tid; idx < kCfNB * Qp; idx += kCfThreads) UA[idx] = 0.0;
    if (has_panel) {
      const double* srcB = prm.UT + (size_t)J * Qp;
      for (int idx = tid; idx < kCfNB * Qp / 2; idx += kCfThreads) cp_async16(UB + 2 * idx, srcB + 2 * idx);
    }
    cp_async_commit();
    double vk[6];
    const double* zv = prm.z_var + w * prm.ldz;
#pragma unroll
    for (int s6 = 0; s6 < 6; s6++) vk[s6] = (4 * s6 + t < prm.Q) ? __ldcg(zv + 4 * s6 + t) : 0.0;
    cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int s6 = 0; s6 < 6; s6++)
      if (4 * s6 < Qp) {
        const int k = 4 * s6 + t;
        const double a = vk[s6] * UA[(8 * wr + g) * Qp + k];
#pragma unroll
        for (int nbk = 0; nbk < 4; nbk++) {
          if (has_panel) dmma884(accp[nbk][0], accp[nbk][1], a, UB[(8 * nbk + g) * Qp + k]);
          if (nbk <= wr) dmma884(accd[nbk][0], accd[nbk][1], a, UA[(8 * nbk + g) * Qp + k]);
        }
      }
  } else if (!DENSE) {
    // rank-Q term for both tiles
    const double* zv = prm.z_var + w * prm.ldz;
    const int row = min(Jd + 8 * wr + g, Mg - 1);
    for (int k0 = 0; k0 < Qp; k0 += 4) {
      const double vk = (k0 + t < prm.Q) ? __ldcg(zv + k0 + t) : 0.0;
      const double a = vk * __ldg(prm.UT + (size_t)row * Qp + k0 + t);
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++) {
        if (has_panel) {
          const double bp = __ldg(prm.UT + (size_t)(J + 8 * nbk + g) * Qp + k0 + t);
          dmma884(accp[nbk][0], accp[nbk][1], a, bp);
        }
        if (nbk <= wr) {
          const double bd = __ldg(prm.UT + (size_t)min(Jd + 8 * nbk + g, Mg - 1) * Qp + k0 + t);
          dmma884(accd[nbk][0], accd[nbk][1], a, bd);
        }
      }
    }
  }
  __syncthreads();   // tvs visible
#pragma unroll 1
  for (int s = 0; s < nst; s++) {
    mbar_wait(&full_bar[s % kCfStages], (s / kCfStages) & 1);
    const double* A = ring + (size_t)(s % kCfStages) * kCfStageDoubles;
    const double* B = A + 2 * kCfRows * 8;
#pragma unroll
    for (int sub = 0; sub < 2; sub++) {
      const double2 a = *reinterpret_cast<const double2*>(A + ((size_t)sub * kCfNB + 8 * wr + g) * 8 + 2 * t);
      // (of the diagonal block only the lower triangle is needed: column blocks nbk <= wr)
      double2 bp[4], bd[4];
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++) {
        bp[nbk] = *reinterpret_cast<const double2*>(B + ((size_t)sub * kCfNB + 8 * nbk + g) * 8 + 2 * t);
        if (nbk <= wr) bd[nbk] = *reinterpret_cast<const double2*>(A + ((size_t)sub * kCfNB + 8 * nbk + g) * 8 + 2 * t);
      }
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++) {
        dmma884(accp[nbk][0], accp[nbk][1], -a.x, bp[nbk].x);
        if (nbk <= wr) dmma884(accd[nbk][0], accd[nbk][1], -a.x, bd[nbk].x);
      }
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++) {
        dmma884(accp[nbk][0], accp[nbk][1], -a.y, bp[nbk].y);
        if (nbk <= wr) dmma884(accd[nbk][0], accd[nbk][1], -a.y, bd[nbk].y);
      }
      const double2 lv = *reinterpret_cast<const double2*>(A + ((size_t)sub * kCfNB + rr) * 8 + 2 * part);
      const double2 tv2 = *reinterpret_cast<const double2*>(tvs + 16 * s + 8 * sub + 2 * part);
      racc = fma(lv.x, tv2.x, racc);
      racc = fma(lv.y, tv2.y, racc);
    }
    if (prm.flags & 128) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s % kCfStages]);
    if (warp == 0 && s + kCfStages - 1 < nst) {
      if (!(prm.flags & 64)) mbar_wait(&empty_bar[s % kCfStages], (s / kCfStages) & 1);   // (see the regular tile)
      else if (s >= 1) mbar_wait(&empty_bar[(s - 1) % kCfStages], ((s - 1) / kCfStages) & 1);
      issue0(s + kCfStages - 1);
    }
  }
  cf_stamp(prm, J, w, 1);
  if (has_panel) {
    if (!early) {
      pdl_wait_prior_grids();          // Dinv_J and t[J:J+32] are final
      fetch_dinv();
    }
    if (tid < kCfNB) tvs[J + tid] = __ldcg(prm.tvec + (size_t)w * Mg + J + tid);
  }
  cp_async_wait<0>();
  __syncthreads();   // Dinv_J and t have landed; the ring is free
  cf_stamp(prm, J, w, 2);
  double* T = ring;                        // [32][kCfLd]  updated rows of panel J, then L[rows, J:J+32]
  if (has_panel) {
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++)
      *reinterpret_cast<double2*>(&T[(8 * wr + g) * kCfLd + 8 * nbk + 2 * t]) = make_double2(accp[nbk][0], accp[nbk][1]);
    __syncwarp();
    double2 a[4];
#pragma unroll
    for (int kk = 0; kk < 4; kk++) a[kk] = *reinterpret_cast<const double2*>(&T[(8 * wr + g) * kCfLd + 8 * kk + 2 * t]);
    double acc2[4][2];
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++) acc2[nbk][0] = acc2[nbk][1] = 0.0;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      double2 b[4];
#pragma unroll
      for (int nbk = kk; nbk < 4; nbk++) b[nbk] = *reinterpret_cast<const double2*>(&Dv[(8 * nbk + g) * kCfLd + 8 * kk + 2 * t]);
#pragma unroll
      for (int nbk = kk; nbk < 4; nbk++) dmma884(acc2[nbk][0], acc2[nbk][1], a[kk].x, b[nbk].x);
#pragma unroll
      for (int nbk = kk; nbk < 4; nbk++) dmma884(acc2[nbk][0], acc2[nbk][1], a[kk].y, b[nbk].y);
    }
    __syncwarp();    // every lane of this warp has read its rows of T
    const bool row_ok = 8 * wr < Rv0;   // (rows past the storage are neither written nor used)
    double* Lpan = Lw + cf_panel_off(Mg, K);
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++) {
      const double2 v = row_ok ? make_double2(acc2[nbk][0], acc2[nbk][1]) : make_double2(0.0, 0.0);
      *reinterpret_cast<double2*>(&T[(8 * wr + g) * kCfLd + 8 * nbk + 2 * t]) = v;
      if (row_ok) *reinterpret_cast<double2*>(Lpan + ((size_t)nbk * RkJ + 8 * wr + g) * 8 + 2 * t) = v;
    }
    __syncthreads();   // T = L[rows, J:J+32] complete
    // D -= L[rows, J:J+32] L[rows, J:J+32]^T;  rhs -= L[rows, J:J+32] t[J:J+32]
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      const double2 a2 = *reinterpret_cast<const double2*>(&T[(8 * wr + g) * kCfLd + 8 * kk + 2 * t]);
      double2 b[4];
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++)
        if (nbk <= wr) b[nbk] = *reinterpret_cast<const double2*>(&T[(8 * nbk + g) * kCfLd + 8 * kk + 2 * t]);
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++)
        if (nbk <= wr) dmma884(accd[nbk][0], accd[nbk][1], -a2.x, b[nbk].x);
#pragma unroll
      for (int nbk = 0; nbk < 4; nbk++)
        if (nbk <= wr) dmma884(accd[nbk][0], accd[nbk][1], -a2.y, b[nbk].y);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {   // (column pairs 2 part + 8 i: conflict-free 16-byte loads)
      const double2 lv = *reinterpret_cast<const double2*>(&T[rr * kCfLd + 2 * part + 8 * i]);
      racc = fma(lv.x, tvs[J + 2 * part + 8 * i], racc);
      racc = fma(lv.y, tvs[J + 2 * part + 8 * i + 1], racc);
    }
  }
  racc += __shfl_xor_sync(0xffffffffu, racc, 1);
  racc += __shfl_xor_sync(0xffffffffu, racc, 2);
  // hand the raw diagonal block (identity padded) and the right-hand side to the factor kernel
  if (part == 0 && rr < nb) prm.tvec[(size_t)w * Mg + Jd + rr] = yrow - racc;
  {
    // (not into dinv: the regular tiles of this walker may still have to read Dinv_J from there)
    double* dw = prm.draw + (size_t)w * kCfNB * kCfNB;
    const int r = 8 * wr + g;
#pragma unroll
    for (int nbk = 0; nbk < 4; nbk++) {
      const int c = 8 * nbk + 2 * t;
      const bool pad = r >= nb || c >= nb;   // nb is even for an even M; an odd M pads column c + 1 below
      double2 v = pad ? make_double2(r == c ? 1.0 : 0.0, r == c + 1 ? 1.0 : 0.0) : make_double2(accd[nbk][0], accd[nbk][1]);
      if (!pad && c + 1 >= nb) v.y = 0.0;
      *reinterpret_cast<double2*>(dw + r * kCfNB + c) = v;
    }
  }
  cf_stamp(prm, J, w, 3);
}

// Factorisation of the 32x32 diagonal blocks: one WARP per walker, everything in shared memory with
// ROLLED loops -- a few dozen instructions that stay in the instruction cache.  Why a kernel of its own:
// inside the panel kernel this serial piece (about 1500 dependent FP64 instructions) shares its SM
// sub-partition's FP64 pipe with three warps issuing 16-cycle DMMAs and gets one issue slot per ~70
// cycles: 50-75 us per block measured, with the 48 KB / 4-warp CTA slot blocked meanwhile (40 % of all
// slot time).  Here the blocks of all walkers are factorised in a window of their own (a few us per launch,
// FP64-throughput bound), or overlap the panel kernel of ANOTHER sub-batch on a second stream.
//   raw block D (written by tile 0) -> factor and inverse (register factorisation, see below);  Dinv to the
//   work buffer, t[Jd:Jd+32] = Dinv rhs, log-determinant, |t|^2; the last block emits lp[w].
constexpr int kCfFactorWarps = 4;
constexpr size_t kCfFactorSmem = sizeof(double) * kCfFactorWarps * (kCfNB * kCfLdD + kCfNB);

__global__ void __launch_bounds__(kCfFactorWarps * 32) chol_fused_factor_kernel(const CholFusedParams prm, int Jd,
                                                                                 int64_t w_first, int64_t n_walkers) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ONE 32x34 tile per walker (24 warps = walkers per SM): its four 16x16 blocks are reused as the
  // factorisation proceeds --  D11 -> L11 -> (dead) -> I22;   D12 (never an input) <- I11;
  // D21 -> L21 -> M = L21 I11 -> Dinv21 = -I22 M;   D22 -> D22 - L21 L21^T -> L22.
  double* D = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (kCfNB * kCfLdD + kCfNB);
  double* red = D + kCfNB * kCfLdD;                                                                // [32] right-hand side
  double* B11 = D, *B12 = D + 16, *B21 = D + 16 * kCfLdD, *B22 = D + 16 * kCfLdD + 16;
  const int64_t wi = (int64_t)blockIdx.x * kCfFactorWarps + warp;
  auto stamp = [&](int slot) {
    if (prm.dbg != nullptr && lane == 0 && w_first + wi < 32)
      prm.dbg[((((Jd / kCfNB) & 15) * 32 + (w_first + wi)) * 8 + 7) * 8 + slot] = clock64();
  };
  stamp(0);
  pdl_wait_prior_grids();      // (first the wait, then the trigger: whoever follows may read everything older
  pdl_launch_dependents();     //  than this grid before its own wait)
  if (wi >= n_walkers) return;
  stamp(1);
  const int64_t w = w_first + wi;
  if (prm.skip != nullptr && __ldcg(prm.skip + w)) return;
  const int M = prm.M, Mg = prm.Mg;
  const int nb = min(kCfNB, M - Jd);
  const bool first = Jd == 0, is_last = Jd + kCfNB >= M;
  const double* dr = prm.draw + (size_t)w * kCfNB * kCfNB;
  double* dw = prm.dinv + (size_t)w * kCfNB * kCfNB;
  {
    // all 8 KB of the block in flight at once (16 coalesced 16-byte loads per lane), then to shared memory
    double2 v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __ldcg(reinterpret_cast<const double2*>(dr + 2 * (lane + 32 * i)));
    const double rv = lane < nb ? __ldcg(prm.tvec + (size_t)w * Mg + Jd + lane) : 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int idx = 2 * (lane + 32 * i);
      *reinterpret_cast<double2*>(&D[(idx >> 5) * kCfLdD + (idx & 31)]) = v[i];
    }
    red[lane] = rv;
  }
  __syncwarp();
  stamp(2);

  // 32x32 block as two 16x16 REGISTER factorisations (lanes 0-15: rows of the block, pivots and
  // multipliers by shuffle; lanes 16-31: identity rows that come out as the inverse) around block products
  // on the tensor pipe:  L11, I11 = chol(D11);  L21 = D21 I11^T;  L22, I22 = chol(D22 - L21 L21^T);
  // Dinv = [[I11, 0], [-I22 L21 I11, I22]].  The critical path is ~170 cycles per pivot (shuffle, rsqrt,
  // multiply, shuffle, fma); a rolled shared-memory formulation measured 1000+ cycles per column.
  bool pd = true;
  const int g = lane >> 2, t = lane & 3;
  const int r = lane & 15;
  double lsum = 0.0;
#pragma unroll 1
  for (int blk = 0; blk < 2; blk++) {
    const int o = 16 * blk;
    if (blk == 1) {
      tile16_mma<true>(B21, B12, nullptr, B21, 1.0, g, t);      // L21 = D21 I11^T
      tile16_mma<true>(B21, B21, B22, B22, -1.0, g, t);         // D22 -= L21 L21^T
    }
    double S[16];
#pragma unroll
    for (int c = 0; c < 16; c++) S[c] = (lane < 16) ? D[(o + r) * kCfLdD + o + c] : (r == c ? 1.0 : 0.0);
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
    // the inverse of this block's factor (identity rows, lanes 16-31): I11 -> B12, I22 -> B11 (L11 is
    // dead by then); the factor itself is only needed for L22's ... nothing: L11 and L22 are not stored
    double* Iout = blk == 0 ? B12 : B11;
    if (lane >= 16) {
#pragma unroll
      for (int c = 0; c < 16; c++) Iout[c * kCfLdD + r] = (c >= r) ? S[c] : 0.0;
    }
    __syncwarp();
  }
  stamp(3);
  tile16_mma<false>(B21, B12, nullptr, B21, 1.0, g, t);     // M = L21 I11       (in place)
  tile16_mma<false>(B11, B21, nullptr, B21, -1.0, g, t);    // Dinv21 = -I22 M   (in place)
  // Dinv = [[I11 (B12), 0], [Dinv21 (B21), I22 (B11)]]
  stamp(4);
  lsum = warp_sum(lsum);                // sum of log(pivot) = log det of the block
  const int bad_now = pd ? 0 : 1;
  if (!is_last) {
    // lane -> (row parity, column pair): rows 2i + (lane >> 4), columns 2 (lane & 15), +1
    const int col = 2 * (lane & 15), cl = col & 15, hi = lane >> 4;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int row = 2 * i + hi;
      double2 v = make_double2(0.0, 0.0);
      if (i < 8) {
        if (col < 16) v = *reinterpret_cast<const double2*>(&B12[row * kCfLdD + cl]);
      } else {
        v = *reinterpret_cast<const double2*>(&(col < 16 ? B21 : B11)[(row - 16) * kCfLdD + cl]);
      }
      *reinterpret_cast<double2*>(dw + row * kCfNB + col) = v;
    }
  }
  double sx = 0.0;
  {
    // t = Dinv rhs: first 16 columns from I11 (rows < 16) or Dinv21, the other 16 from I22 (rows >= 16);
    // the zeros above the diagonals are stored, so both loops are uniform
    const double* ra = lane < 16 ? B12 + lane * kCfLdD : B21 + (lane - 16) * kCfLdD;
    const double* rb = B11 + (lane & 15) * kCfLdD;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k += 2) {
      const double2 a = *reinterpret_cast<const double2*>(ra + k), b = *reinterpret_cast<const double2*>(rb + k);
      s0 = fma(a.x, red[k], s0);
      s1 = fma(a.y, red[k + 1], s1);
      s2 = fma(b.x, red[16 + k], s2);
      s3 = fma(b.y, red[17 + k], s3);
    }
    sx = (s0 + s1) + (lane >= 16 ? s2 + s3 : 0.0);
  }
  if (lane < nb) {
    prm.tvec[(size_t)w * Mg + Jd + lane] = sx;
  }
  const double q2 = warp_sum(sx * sx);
  if (lane == 0) {
    const double ld = (first ? 0.0 : __ldcg(prm.logdet + w)) + lsum;
    const double tot = (first ? 0.0 : __ldcg(prm.tsq + w)) + q2;
    const int bad = (first ? 0 : __ldcg(prm.bad + w)) | bad_now;
    if (!is_last) {
      prm.logdet[w] = ld;
      prm.tsq[w] = tot;
      prm.bad[w] = bad;
    } else if (bad) {
      prm.lp[w] = prm.notpd_value;
      if (prm.n_notpd) atomicAdd(prm.n_notpd, 1);
    } else {
      prm.lp[w] = -0.5 * tot - 0.5 * ld + prm.add_const;
    }
  }
  stamp(5);
}

}  // namespace gpbt
