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
// One CTA = one tile of TW walkers x one PC.
//   phase 1  K^T[k][w] for k < n_pad into shared memory (FP64 pipe: distance + exp), z_mean on
//            the fly.  Layout Kt[k*TW + (w ^ swz(k))] makes both the phase-1 stores and the
//            phase-2 B-fragment loads bank-conflict free.
//   phase 2  for each 32-row block of W_j: acc[32 x TW] += W_j[rows, k] * Kt[k, :] over k <= row,
//            A fragments straight from global/L2 (each element of W_j is needed by exactly one
//            warp of the CTA, so staging it in shared memory buys nothing), then square and
//            column-sum the accumulators.  Row blocks r and nrb-1-r are paired on one warp so
//            the triangular work is balanced.
#pragma once
#include "common.cuh"

namespace gpbt {

struct PcPredictParams {
  const double* __restrict__ X;      // [N, p] walkers
  const double* __restrict__ extra;  // [N] extra_std or nullptr
  const double* __restrict__ Xs;     // [q, n_pad, p_pad]  X_train / ell_j, zero padded
  const double* __restrict__ ell;    // [q, p_pad]          (pad = 1)
  const double* __restrict__ c;      // [q]
  const double* __restrict__ sn;     // [q]
  const double* __restrict__ alpha;  // [q, n_pad] zero padded
  const double* __restrict__ W;      // [q, n_pad, n_pad] lower-triangular inverse of L_j, zero padded
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
constexpr int kRowBlock = 32;  // rows of W per warp task
constexpr int kKChunk = 16;    // k extent of one A-fragment fetch (two 16-byte loads per m8 block)

template <int TW>
constexpr size_t pc_predict_smem_bytes(int n_pad, int p_pad) {
  return sizeof(double) * ((size_t)n_pad * TW + (size_t)p_pad * TW + 2 * (size_t)kPcWarps * TW);
}

template <int TW, int KIND>
__global__ void __launch_bounds__(kPcThreads, 1) pc_predict_kernel(const PcPredictParams prm) {
  constexpr int NT = TW / 8;  // n8 tiles across the walker dimension
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* Kt = reinterpret_cast<double*>(smem_raw);      // [n_pad][TW] swizzled
  double* xs = Kt + (size_t)prm.n_pad * TW;              // [p_pad][TW]
  double* red_mean = xs + (size_t)prm.p_pad * TW;        // [warps][TW]
  double* red_ssq = red_mean + kPcWarps * TW;            // [warps][TW]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.y;
  const int64_t w0 = (int64_t)blockIdx.x * TW;
  const int p = prm.p, p_pad = prm.p_pad, n = prm.n, n_pad = prm.n_pad;

  // ---- stage the scaled walker tile: xs[d][w] = X[w0+w][d] / ell_j[d] (true division, as
  //      sklearn's X / length_scale) ---------------------------------------------------------
  for (int idx = tid; idx < p_pad * TW; idx += kPcThreads) {
    const int d = idx / TW, w = idx - d * TW;
    double v = 0.0;
    if (d < p && w0 + w < prm.N) v = prm.X[(w0 + w) * p + d] / prm.ell[(size_t)j * p_pad + d];
    xs[idx] = v;
  }
  __syncthreads();

  // ---- phase 1: Kt and the mean ------------------------------------------------------------
  const double cj = prm.c[j];
  const double* __restrict__ Xs_j = prm.Xs + (size_t)j * n_pad * p_pad;
  const double* __restrict__ alpha_j = prm.alpha + (size_t)j * n_pad;
  {
    constexpr int WPR = TW / 32 > 0 ? TW / 32 : 1;  // warps needed to cover one k row (TW=32 -> 1)
    static_assert(TW == 8 || TW == 16 || TW == 32 || TW == 64, "TW must be 8, 16, 32 or 64");
    constexpr int LW = TW < 32 ? TW : 32;           // walkers covered by one warp pass
    constexpr int KPW = 32 / LW;                    // k rows handled per warp pass (TW<32)
    constexpr int UNR = 4;
    const int wl = lane % LW;                       // walker within tile (first slab)
    const int ksub = lane / LW;                     // which of the KPW rows this lane takes
    double msum[WPR];
#pragma unroll
    for (int s = 0; s < WPR; s++) msum[s] = 0.0;
    // rows are dealt to warps in groups of UNR*KPW for ILP
    for (int kb = warp * UNR * KPW; kb < n_pad; kb += kPcWarps * UNR * KPW) {
#pragma unroll
      for (int s = 0; s < WPR; s++) {
        const int w = wl + 32 * s;
        double acc[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) acc[u] = 0.0;
        for (int d = 0; d < p_pad; d += 2) {
          const double x0 = xs[d * TW + w], x1 = xs[(d + 1) * TW + w];
#pragma unroll
          for (int u = 0; u < UNR; u++) {
            const int k = min(kb + u * KPW + ksub, n_pad - 1);
            const double2 t = ldg2(Xs_j + (size_t)k * p_pad + d);
            const double e0 = x0 - t.x, e1 = x1 - t.y;
            acc[u] = fma(e0, e0, acc[u]);
            acc[u] = fma(e1, e1, acc[u]);
          }
        }
#pragma unroll
        for (int u = 0; u < UNR; u++) {
          const int k = kb + u * KPW + ksub;
          double kv;
          if (KIND == 0) {
            kv = cj * exp(-0.5 * acc[u]);
          } else {
            const double r = sqrt(acc[u]) * 1.7320508075688772;  // sqrt(3) rounded as np.sqrt(3)
            kv = cj * ((1.0 + r) * exp(-r));
          }
          if (k >= n) kv = 0.0;
          if (k < n_pad) {
            Kt[(size_t)k * TW + (w ^ kt_swizzle<TW>(k))] = kv;
            msum[s] = fma(kv, alpha_j[k], msum[s]);
          }
        }
      }
    }
    // combine the KPW partial sums that belong to the same walker, then park per-warp partials
#pragma unroll
    for (int s = 0; s < WPR; s++) {
      double v = msum[s];
      for (int o = LW; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane < LW) red_mean[warp * TW + wl + 32 * s] = v;
    }
  }
  __syncthreads();

  // ---- phase 2: ssq[w] = sum_i (sum_{k<=i} W[i][k] K[w][k])^2 on the FP64 tensor pipe ---------
  const double* __restrict__ W_j = prm.W + (size_t)j * n_pad * n_pad;
  const int nrb = n_pad / kRowBlock;
  const int g = lane >> 2, t = lane & 3;  // fragment row group / k slot
  double ssq[NT][2];
#pragma unroll
  for (int nt = 0; nt < NT; nt++) ssq[nt][0] = ssq[nt][1] = 0.0;

  // task list of this warp: pairs (r, nrb-1-r) for r = warp, warp + kPcWarps, ...
  for (int pr = warp; 2 * pr < nrb; pr += kPcWarps) {
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
      const int rb = half == 0 ? pr : nrb - 1 - pr;
      if (half == 1 && rb == pr) break;
      const int i0 = rb * kRowBlock;
      double acc[4][NT][2];
#pragma unroll
      for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nt = 0; nt < NT; nt++) acc[mb][nt][0] = acc[mb][nt][1] = 0.0;

      const int nchunk = (i0 + kRowBlock) / kKChunk;  // k chunks 0 .. nchunk-1 touch rows <= i0+31
      // A fragments for one chunk: lane holds W[i0 + 8mb + g][k0 + 4t .. k0 + 4t + 3]
      double a_cur[4][4], a_nxt[4][4];
      auto load_a = [&](double (&a)[4][4], int kc) {
        const double* base = W_j + (size_t)(i0 + g) * n_pad + kc * kKChunk + 4 * t;
#pragma unroll
        for (int mb = 0; mb < 4; mb++) {
          const double2 lo = ldg2(base + (size_t)(8 * mb) * n_pad);
          const double2 hi = ldg2(base + (size_t)(8 * mb) * n_pad + 2);
          a[mb][0] = lo.x; a[mb][1] = lo.y; a[mb][2] = hi.x; a[mb][3] = hi.y;
        }
      };
      load_a(a_cur, 0);
#pragma unroll 1
      for (int kc = 0; kc < nchunk; kc++) {
        if (kc + 1 < nchunk) load_a(a_nxt, kc + 1);
        const int k0 = kc * kKChunk;
        // m8 blocks whose rows all lie above this chunk's columns hold only zeros of the
        // (strictly upper) triangle: skip them.  rows of block mb: i0+8mb .. i0+8mb+7
        const int mb_first = (k0 > i0) ? ((k0 - i0) >> 3) : 0;
#pragma unroll
        for (int s = 0; s < 4; s++) {
          // logical k slot t of step s  <->  actual k = k0 + 4t + s  (same permutation for A and B)
          const int k = k0 + 4 * t + s;
          double b[NT];
#pragma unroll
          for (int nt = 0; nt < NT; nt++) b[nt] = Kt[(size_t)k * TW + ((8 * nt + g) ^ kt_swizzle<TW>(k))];
#pragma unroll
          for (int mb = 0; mb < 4; mb++) {
            if (mb >= mb_first) {
#pragma unroll
              for (int nt = 0; nt < NT; nt++) dmma884(acc[mb][nt][0], acc[mb][nt][1], a_cur[mb][s], b[nt]);
            }
          }
        }
        if (kc + 1 < nchunk) {
#pragma unroll
          for (int mb = 0; mb < 4; mb++)
#pragma unroll
            for (int s = 0; s < 4; s++) a_cur[mb][s] = a_nxt[mb][s];
        }
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
    if (prm.extra != nullptr) {
      const double e = prm.extra[w0 + tid];
      var += e * e;
    }
    prm.z_mean[(w0 + tid) * prm.ldz + j] = mean;
    prm.z_var[(w0 + tid) * prm.ldz + j] = var;
  }
}

}  // namespace gpbt
