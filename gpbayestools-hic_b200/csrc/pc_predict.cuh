// Kernel (a): fused cross-kernel + GP mean + predictive variance for every (walker, PC).
//
// Replaces, for all q GPs of one emulator (reference: src/emulator.py:553,573-579; sklearn
// _gpr.py:446-466, kernels.py:1558-1571 / 1713-1729):
//     K      = c_j * kappa((x - X_train) / ell_j)                     [N, n]
//     z_mean = K @ alpha_j
//     V      = solve_triangular(L_j, K.T);  z_var = c_j + sn_j - sum(V*V, 0)   (+ extra_std^2)
// The triangular solve is evaluated as the product with the explicit lower-triangular inverse
// W_j = L_j^-1 (inverted once on the host in FP64): V^T = K W_j^T, so the O(N n^2) term is a
// triangular GEMM that runs on the FP64 tensor pipe (DMMA.8x8x4).
//
// KIND 2 is the surmise PCGP / PCSK predictor behind EmulatorBAND.predict (src/emulator_BAND.py:386-478;
// the arithmetic lives in surmise 0.2.1 emulationmethods/PCGP.py, which is not part of the reference
// tree -- restated from its published form, PARITY UNPINNED):
//     r      = (1 - nug_j) * (a_j * prod_d (1 + s_d) * exp(-sum_d s_d) + b_j),  s_d = |x_d - theta_d| / exp(gamma_jd)
//     z_mean = r @ pw_j,      z_var = sig2_j * | 1 - || r Vh_j ||^2 |
// with c = (1 - nug) a, sn = (1 - nug) b and W_j = Vh_j^T, a DENSE n x n matrix: phase 2 then runs over
// all k for every row block instead of k <= row.
//
// One CTA = one tile of TW walkers x one PC (TW = 16 with two CTAs resident per SM for large
// batches, so one CTA's phase 1 overlaps the other's phase 2; 32 when two do not fit; 8 for small N).
//   phase 1  the scaled design (rows X_train[k]/ell_j with alpha_j[k] appended) streams through
//            shared memory in 128-row stages by TMA bulk copy (cp.async.bulk + mbarrier, double
//            buffered); K^T[k][w] for k < n_pad goes to shared memory (FP64 pipe: distance + a
//            branch-free exp), z_mean on the fly.  Layout Kt[k*TW + (w ^ swz(k))] makes both the
//            phase-1 stores and the phase-2 B-fragment loads bank-conflict free.
//   phase 2  for each 32-row block of W_j: acc[32 x TW] += W_j[rows, k] * Kt[k, :] over k <= row,
//            A fragments straight from global/L2 (each element of W_j is needed by exactly one
//            warp of the CTA, so staging it in shared memory buys nothing), then square and
//            column-sum the accumulators.  Row blocks r and nrb-1-r are paired on one warp so
//            the triangular work is balanced.
#pragma once
#include "common.cuh"
#include "fast_exp.cuh"

namespace gpbt {

struct PcPredictParams {
  const double* __restrict__ X;      // [N, p] walkers
  const double* __restrict__ extra;  // [N] extra_std or nullptr
  const double* __restrict__ Xs;     // [q, n_pad, p_pad + 2]  row k = (X_train[k] / ell_j, 0-pad, alpha_j[k], 0)
  const double* __restrict__ ell;    // [q, p_pad]          (pad = 1)
  const double* __restrict__ c;      // [q]
  const double* __restrict__ sn;     // [q]
  const double* __restrict__ W;      // [q, n_pad, n_pad] lower-triangular inverse of L_j, zero padded
  const double* __restrict__ sig2;   // [q]  KIND 2 only
  double* __restrict__ z_mean;       // [N, ldz]
  double* __restrict__ z_var;        // [N, ldz]
  int64_t ldz;
  int64_t N;
  int p, p_pad, n, n_pad, q;
};

// XOR applied to the walker index of row k of Kt (kept inside the row for TW < 16)
template <int TW>
__host__ __device__ constexpr int kt_swizzle(int k) { return (((k >> 2) & 3) << 2) & (TW - 1); }

constexpr int kPcThreads = 256;
constexpr int kPcWarps = kPcThreads / kWarp;
constexpr int kRowBlock = 32;   // rows of W per warp task
constexpr int kKChunk = 16;     // k extent of one A-fragment fetch (two 16-byte loads per m8 block)
constexpr int kXsRows = 128;    // rows of the scaled design staged per TMA bulk copy
constexpr int kXsStages = 2;

template <int TW>
constexpr size_t pc_predict_smem_bytes(int n_pad, int p_pad) {
  return sizeof(double) * ((size_t)n_pad * TW + (size_t)kXsStages * kXsRows * (p_pad + 2) + (size_t)p_pad * TW +
                           2 * (size_t)kPcWarps * TW) + 64 * sizeof(double2) + 64;
}

// A fragments of one k chunk: lane (g, t) holds W[i0 + 8mb + g][16kc + 4t .. 16kc + 4t + 3]
template <int MB0>
__device__ __forceinline__ void load_a(double (&a)[4][4], const double* __restrict__ Wrow, int n_pad, int kc) {
  const double* base = Wrow + kc * kKChunk;
#pragma unroll
  for (int mb = MB0; mb < 4; mb++) {
    const double2 lo = ldg2(base + (size_t)(8 * mb) * n_pad);
    const double2 hi = ldg2(base + (size_t)(8 * mb) * n_pad + 2);
    a[mb][0] = lo.x; a[mb][1] = lo.y; a[mb][2] = hi.x; a[mb][3] = hi.y;
  }
}

// acc[mb][nt] += W-chunk * Kt-chunk for m8 blocks mb >= MB0.  Logical k slot t of step s is the
// actual column k0 + 4t + s: the same permutation is applied to the A and B operands.
template <int TW, int MB0>
__device__ __forceinline__ void mma_chunk(double (&acc)[4][TW / 8][2], const double (&a)[4][4],
                                          const double* Kt, int k0, int g, int t) {
  constexpr int NT = TW / 8;
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const int k = k0 + 4 * t + s;
    const double* row = Kt + (size_t)k * TW;
    const int swz = kt_swizzle<TW>(k);
    double b[NT];
#pragma unroll
    for (int nt = 0; nt < NT; nt++) b[nt] = row[(8 * nt + g) ^ swz];
#pragma unroll
    for (int mb = MB0; mb < 4; mb++)
#pragma unroll
      for (int nt = 0; nt < NT; nt++) dmma884(acc[mb][nt][0], acc[mb][nt][1], a[mb][s], b[nt]);
  }
}

// P2 > 0: p_pad / 2 known at compile time -- the walker's scaled coordinates stay in registers for
// the whole of phase 1 and the distance loop is fully unrolled (only warp-uniform, broadcast shared
// memory reads remain).  P2 == 0: generic p, coordinates re-read from shared memory.
template <int TW, int KIND, int P2>
__global__ void __launch_bounds__(kPcThreads, TW == 16 ? 2 : 1) pc_predict_kernel(const PcPredictParams prm) {
  constexpr int NT = TW / 8;  // n8 tiles across the walker dimension
  static_assert(TW == 8 || TW == 16 || TW == 32, "TW must be 8, 16 or 32");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int p = prm.p, p_pad = P2 > 0 ? 2 * P2 : prm.p_pad, n = prm.n, n_pad = prm.n_pad;
  const int xrow = p_pad + 2;                                       // doubles per staged design row
  double* Kt = reinterpret_cast<double*>(smem_raw);                 // [n_pad][TW] swizzled
  double* xst = Kt + (size_t)n_pad * TW;                            // [stages][kXsRows][xrow]
  double* xs = xst + (size_t)kXsStages * kXsRows * xrow;            // [p_pad/2][TW][2]
  double* red_mean = xs + (size_t)p_pad * TW;                       // [warps][TW]
  double* red_ssq = red_mean + kPcWarps * TW;                       // [warps][TW]
  double2* etab = reinterpret_cast<double2*>(red_ssq + kPcWarps * TW);    // [64] 2^(j/64) (hi, lo)
  uint64_t* bars = reinterpret_cast<uint64_t*>(etab + 64);                // [stages]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.y;
  const int64_t w0 = (int64_t)blockIdx.x * TW;
  const double* __restrict__ Xs_j = prm.Xs + (size_t)j * n_pad * xrow;
  const int n_xchunk = (n_pad + kXsRows - 1) / kXsRows;
  auto chunk_bytes = [&](int c) { return (uint32_t)(min(kXsRows, n_pad - c * kXsRows) * xrow * sizeof(double)); };

  // ---- TMA: first two chunks of the scaled design X_train / ell_j into shared memory ---------
  if (tid == 0) {
    for (int s = 0; s < kXsStages; s++) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int c = 0; c < kXsStages && c < n_xchunk; c++) {
      mbar_expect_tx(&bars[c], chunk_bytes(c));
      tma_bulk_g2s(xst + (size_t)c * kXsRows * xrow, Xs_j + (size_t)c * kXsRows * xrow, chunk_bytes(c), &bars[c]);
    }
  }
  // (the design is constant: its first TMA copies are in flight before this kernel waits for whoever
  // produced X -- a sampler's proposal kernel, the parameter pre-transform)
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  // ---- the scaled walker tile: xs[d/2][w] = (x_d, x_d+1) / ell_j (true division, as sklearn's
  //      X / length_scale) ------------------------------------------------------------------
  for (int idx = tid; idx < p_pad * TW; idx += kPcThreads) {
    const int d = idx / TW, w = idx - d * TW;
    double v = 0.0;
    if (d < p && w0 + w < prm.N) v = prm.X[(w0 + w) * p + d] / prm.ell[(size_t)j * p_pad + d];
    xs[((size_t)(d >> 1) * TW + w) * 2 + (d & 1)] = v;
  }
  if (tid < 64) etab[tid] = kExpTable[tid];
  __syncthreads();

  // ---- phase 1: Kt and the mean ------------------------------------------------------------
  const double cj = prm.c[j];
  const double offj = KIND == 2 ? prm.sn[j] : 0.0;
  {
    constexpr int LW = TW < 32 ? TW : 32;  // walkers covered by one warp pass
    constexpr int KPW = 32 / LW;           // k rows handled per warp pass (TW < 32)
    constexpr int UNR = 4;   // rows per warp pass (8 measured no faster)
    const int w = lane % LW;               // walker within the tile
    const int ksub = lane / LW;            // which of the KPW rows this lane takes
    const double2* xs2 = reinterpret_cast<const double2*>(xs);
    double2 xreg[P2 > 0 ? P2 : 1];
    if (P2 > 0) {
#pragma unroll
      for (int d2 = 0; d2 < P2; d2++) xreg[d2] = xs2[d2 * TW + w];
    }
    double msum = 0.0;
    for (int c = 0; c < n_xchunk; c++) {
      const int stage = c % kXsStages;
      const double* xc = xst + (size_t)stage * kXsRows * xrow;
      const int rows_c = min(kXsRows, n_pad - c * kXsRows);
      mbar_wait(&bars[stage], (c / kXsStages) & 1);
      // rows of this chunk are dealt to warps in groups of UNR*KPW for ILP
      for (int rb = warp * UNR * KPW; rb < rows_c; rb += kPcWarps * UNR * KPW) {
        double acc[UNR], acc1[UNR];
        const double2* xr[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          acc[u] = 0.0;
          acc1[u] = KIND == 2 ? 1.0 : 0.0;   // KIND 2: acc = sum of s_d, acc1 = prod of (1 + s_d)
          xr[u] = reinterpret_cast<const double2*>(xc + (size_t)min(rb + u * KPW + ksub, rows_c - 1) * xrow);
        }
        if (P2 > 0) {
#pragma unroll
          for (int d2 = 0; d2 < P2; d2++) {
#pragma unroll
            for (int u = 0; u < UNR; u++) {
              const double2 tr = xr[u][d2];
              const double e0 = xreg[d2].x - tr.x, e1 = xreg[d2].y - tr.y;
              if (KIND == 2) {
                const double s0 = fabs(e0), s1 = fabs(e1);
                acc[u] += s0 + s1;
                acc1[u] = fma(acc1[u], s0, acc1[u]);
                acc1[u] = fma(acc1[u], s1, acc1[u]);
              } else {
                acc[u] = fma(e0, e0, acc[u]);
                acc1[u] = fma(e1, e1, acc1[u]);
              }
            }
          }
        } else {
          const double2* xq = xs2 + w;
#pragma unroll 3
          for (int d2 = 0; d2 < (p_pad >> 1); d2++) {
            const double2 x = xq[d2 * TW];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
              const double2 tr = xr[u][d2];
              const double e0 = x.x - tr.x, e1 = x.y - tr.y;
              if (KIND == 2) {
                const double s0 = fabs(e0), s1 = fabs(e1);
                acc[u] += s0 + s1;
                acc1[u] = fma(acc1[u], s0, acc1[u]);
                acc1[u] = fma(acc1[u], s1, acc1[u]);
              } else {
                acc[u] = fma(e0, e0, acc[u]);
                acc1[u] = fma(e1, e1, acc1[u]);
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          const double d2sum = acc[u] + acc1[u];
          const int kl = rb + u * KPW + ksub;
          const int k = c * kXsRows + kl;
          double kv;
          if (KIND == 2) {
            kv = fma(cj, acc1[u] * exp_neg(-acc[u], etab), offj);
          } else if (KIND == 0) {
            kv = cj * exp_neg(-0.5 * d2sum, etab);
          } else {
            const double r = sqrt(d2sum) * 1.7320508075688772;  // sqrt(3) rounded as np.sqrt(3)
            kv = cj * ((1.0 + r) * exp_neg(-r, etab));
          }
          if (k >= n) kv = 0.0;
          if (kl < rows_c) {
            Kt[(size_t)k * TW + (w ^ kt_swizzle<TW>(k))] = kv;
            msum = fma(kv, xr[u][p_pad >> 1].x, msum);   // alpha_j[k] rides in the staged row
          }
        }
      }
      __syncthreads();  // every warp is done with this stage
      if (tid == 0 && c + kXsStages < n_xchunk) {
        const int cn = c + kXsStages;
        mbar_expect_tx(&bars[stage], chunk_bytes(cn));
        tma_bulk_g2s(xst + (size_t)stage * kXsRows * xrow, Xs_j + (size_t)cn * kXsRows * xrow, chunk_bytes(cn),
                     &bars[stage]);
      }
    }
    // combine the KPW partial sums that belong to the same walker, then park per-warp partials
    for (int o = LW; o < 32; o <<= 1) msum += __shfl_xor_sync(0xffffffffu, msum, o);
    if (lane < LW) red_mean[warp * TW + w] = msum;
  }
  // (the trailing __syncthreads of the last chunk also publishes Kt)

  // ---- phase 2: ssq[w] = sum_i (sum_{k<=i} W[i][k] K[w][k])^2 on the FP64 tensor pipe ---------
  const double* __restrict__ W_j = prm.W + (size_t)j * n_pad * n_pad;
  const int nrb = n_pad / kRowBlock;
  const int g = lane >> 2, t = lane & 3;  // fragment row group / k slot
  double ssq[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) ssq[nt][0] = ssq[nt][1] = 0.0;

  // task list of this warp: pairs of row blocks (r, nrb-1-r) for r = warp, warp + kPcWarps, ...
  for (int pr = warp; 2 * pr < nrb; pr += kPcWarps) {
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      const int rb = half == 0 ? pr : nrb - 1 - pr;
      if (half == 1 && rb == pr) break;
      const int i0 = rb * kRowBlock;
      const double* __restrict__ Wrow = W_j + (size_t)(i0 + g) * n_pad + 4 * t;
      double acc[4][NT][2];
#pragma unroll
      for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) acc[mb][nt][0] = acc[mb][nt][1] = 0.0;
      // k chunks 0 .. 2rb+1 touch this row block; they are consumed in pairs from two register
      // buffers (no copies, loads run one full chunk ahead of their use)
      double a0[4][4], a1[4][4];
      load_a<0>(a0, Wrow, n_pad, 0);
      if (KIND == 2) {
        // dense W: every k chunk contributes to every row block
        const int nch = n_pad / kKChunk;
#pragma unroll 1
        for (int cp = 0; 2 * cp < nch; cp++) {
          load_a<0>(a1, Wrow, n_pad, 2 * cp + 1);
          mma_chunk<TW, 0>(acc, a0, Kt, (2 * cp) * kKChunk, g, t);
          load_a<0>(a0, Wrow, n_pad, min(2 * cp + 2, nch - 1));   // (last pass: a harmless reload)
          mma_chunk<TW, 0>(acc, a1, Kt, (2 * cp + 1) * kKChunk, g, t);
        }
      } else {
#pragma unroll 1
      for (int cp = 0; cp < rb; cp++) {
        load_a<0>(a1, Wrow, n_pad, 2 * cp + 1);
        mma_chunk<TW, 0>(acc, a0, Kt, (2 * cp) * kKChunk, g, t);
        load_a<0>(a0, Wrow, n_pad, 2 * cp + 2);
        mma_chunk<TW, 0>(acc, a1, Kt, (2 * cp + 1) * kKChunk, g, t);
      }
      // last pair: chunk 2rb holds the diagonal; in chunk 2rb+1 the m8 blocks 0 and 1 (rows
      // i0 .. i0+15) lie strictly above the diagonal and are skipped
      load_a<2>(a1, Wrow, n_pad, 2 * rb + 1);
      mma_chunk<TW, 0>(acc, a0, Kt, (2 * rb) * kKChunk, g, t);
      mma_chunk<TW, 2>(acc, a1, Kt, (2 * rb + 1) * kKChunk, g, t);
      }
#pragma unroll
      for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) {
          ssq[nt][0] = fma(acc[mb][nt][0], acc[mb][nt][0], ssq[nt][0]);
          ssq[nt][1] = fma(acc[mb][nt][1], acc[mb][nt][1], ssq[nt][1]);
        }
    }
  }
  // column sums over the 8 row groups (lanes with equal t own the same walkers)
#pragma unroll
  for (int nt = 0; nt < NT; nt++)
#pragma unroll
    for (int h = 0; h < 2; h++) {
      double v = ssq[nt][h];
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (g == 0) red_ssq[warp * TW + 8 * nt + 2 * t + h] = v;
    }
  __syncthreads();

  if (tid < TW && w0 + tid < prm.N) {
    double mean = 0.0, s2 = 0.0;
#pragma unroll
    for (int wp = 0; wp < kPcWarps; wp++) {
      mean += red_mean[wp * TW + tid];
      s2 += red_ssq[wp * TW + tid];
    }
    double var = (cj + prm.sn[j]) - s2;  // diag(kernel_(X)) = c + sn; no clamping (return_cov branch)
    if (KIND == 2) var = prm.sig2[j] * fabs(1.0 - s2);   // extra_std is accepted and ignored by EmulatorBAND
    if (KIND != 2 && prm.extra != nullptr) {
      const double e = prm.extra[w0 + tid];
      var += e * e;
    }
    prm.z_mean[(w0 + tid) * prm.ldz + j] = mean;
    prm.z_var[(w0 + tid) * prm.ldz + j] = var;
  }
}

}  // namespace gpbt
