// Device-resident parallel-tempering Metropolis iteration: the sampling loop of the reference's PTLMC
// (src/mcmc.py:623-671, no-gradient branch -- the one Chain.log_posterior exercises) and its temperature
// exchange (tempexchange, src/mcmc.py:679-693), with the chains kept in HBM.
//
//   proposal   theta' = theta + sqrt(2) rho_T (xi C^1/2),  rho_T = stride(tau) T^(1/3),  xi ~ N(0, I)
//   accept     log u < (lp' - lp) / T
//   exchange   5 sweeps; a sweep is n random slots rt in [1, n), taken IN ORDER:
//                 swap chains rt-1 and rt  if  (lp[rt] - lp[rt-1]) (1/T[rt-1] - 1/T[rt]) > log u
//   tuning     every 10th of the first n_tune iterations: tau += (hits / 10 - goal) / sqrt(1 + k / 10)
//   record     after tuning: the T = 1 chains
//
// One iteration = ptlmc_propose_kernel -> the log-posterior path -> ptlmc_step_kernel (ONE CTA: accept in parallel,
// the sweeps on one warp, permutation / tuning / record in parallel again).  Random numbers: counter-based
// Philox keyed by (seed, iteration, index, purpose), so a run is reproducible and has a NumPy restatement
// (oracle/ptlmc_oracle.py) that the tests follow draw for draw; it does not follow NumPy's global stream --
// the host driver (gpbt_b200/ptlmc.py) does that.
//
// The sweeps are sequential by definition (a swap changes what the next slot sees).  Slots that are at least two
// apart commute, so a warp takes 32 slots at a time: the lanes whose slot touches no chain of an earlier slot of
// the window are applied at once, the others (with n = 8192 chains four windows in five have none) one by one
// in order -- slot for slot the sequential result.  Drawing the slots (Philox, log) and finding those lanes is
// done by all warps beforehand; the walk through the windows is a few shared-memory reads per window.
#pragma once
#include "common.cuh"
#include "ensemble.cuh"

namespace gpbt {

constexpr uint32_t kPtTagNormal = 16, kPtTagAccept = 17, kPtTagSweep = 20;   // Philox purposes
constexpr int kPtSweeps = 5;
constexpr int kPtStepThreads = 1024;

struct PtlmcCtl {
  double tau, hits, stride;
  long long accepted;        // accepted proposals of the T = 1 chains after tuning (diagnostics)
};

struct PtlmcParams {
  PtlmcCtl* ctl;
  uint64_t seed;
  int n, p, n_hot;                       // chains, parameters, chains above T = 1 (they come first)
  const double* __restrict__ temps;      // [n]
  const double* __restrict__ cbrt_t;     // [n]  T^(1/3)
  const double* __restrict__ gap;        // [n]  1/T[i-1] - 1/T[i]  (gap[0] unused)
  const double* __restrict__ root;       // [p, p]  C^1/2
  const double* theta_in;                // [n, p]
  const double* lp_in;                   // [n]
  double* prop;                          // [n, p]
  const double* lp_prop;                 // [n]
  double* theta_out;                     // [n, p]
  double* lp_out;                        // [n]
  double* saved;                         // [n - n_hot, n_keep, p]
  int* sw_slot;                          // [n]  scratch of one exchange sweep: slot,
  double* sw_gap;                        // [n]    the ladder gap at the slot,
  double* sw_logu;                       // [n]    log of the uniform
  long long k, n_tune, n_keep;
  double goal;
};

__host__ __device__ inline double ptlmc_stride(double tau) {
  const double e = exp(2.0 * tau);
  return 2.0 * (1.0 + (e - 1.0) / (e + 1.0));
}

// standard normals 2 e2, 2 e2 + 1 of chain i in iteration k (Box-Muller on two 53-bit uniforms)
__device__ __forceinline__ double2 ptlmc_normal_pair(uint64_t seed, long long k, int i, int e2, int p2) {
  uint32_t r[4];
  philox4x32(seed, (uint64_t)k, (uint32_t)(i * p2 + e2), kPtTagNormal, r);
  const double u1 = 1.0 - u01(r[0], r[1]);          // (0, 1]
  const double u2 = u01(r[2], r[3]);
  const double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincos(6.283185307179586 * u2, &s, &c);
  return make_double2(rad * c, rad * s);
}

// one warp per chain
__global__ void __launch_bounds__(128) ptlmc_propose_kernel(const PtlmcParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int p = prm.p, p2 = (p + 1) / 2;
  double* xi = reinterpret_cast<double*>(smem_raw) + (size_t)warp * 2 * p2;
  const int i = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= prm.n) return;
  for (int e2 = lane; e2 < p2; e2 += 32) {
    const double2 z = ptlmc_normal_pair(prm.seed, prm.k, i, e2, p2);
    xi[2 * e2] = z.x;
    xi[2 * e2 + 1] = z.y;
  }
  __syncwarp();
  // (what the previous iteration's kernels wrote is read from L2, never through L1: see chol_fused.cuh)
  const double scale = 1.4142135623730951 * (__ldcg(&prm.ctl->stride) * prm.cbrt_t[i]);
  for (int d = lane; d < p; d += 32) {
    double s = 0.0;
    for (int e = 0; e < p; e++) s = fma(xi[e], prm.root[(size_t)e * p + d], s);
    prm.prop[(size_t)i * p + d] = __ldcg(prm.theta_in + (size_t)i * p + d) + scale * s;
  }
}

// accept + exchange + permutation + tuning + record: ONE CTA.  Shared memory: lp [n], order [n], one mask per
// window of 32 slots, taken [n].
inline size_t ptlmc_step_smem_bytes(int n) {
  return (size_t)n * (sizeof(double) + sizeof(int) + 1) + (size_t)((n + 31) / 32) * sizeof(unsigned int) + 64;
}

__global__ void __launch_bounds__(kPtStepThreads) ptlmc_step_kernel(const PtlmcParams prm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  const int n = prm.n, p = prm.p;
  double* lp_s = reinterpret_cast<double*>(smem_raw);
  int* order = reinterpret_cast<int*>(lp_s + n);
  unsigned int* win_marked = reinterpret_cast<unsigned int*>(order + n);
  unsigned char* taken = reinterpret_cast<unsigned char*>(win_marked + (n + 31) / 32);
  __shared__ int s_hits, s_cold_hits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_hits = s_cold_hits = 0;
  __syncthreads();

  // ---- Metropolis accept, tempered: log u < (lp' - lp) / T ------------------------------------
  int mine = 0, mine_cold = 0;
  for (int i = tid; i < n; i += blockDim.x) {
    uint32_t r[4];
    philox4x32(prm.seed, (uint64_t)prm.k, (uint32_t)i, kPtTagAccept, r);
    const double logu = log(u01(r[0], r[1]));
    const double lpo = __ldcg(prm.lp_in + i), lpn = __ldcg(prm.lp_prop + i);
    const bool take = logu < (lpn - lpo) / prm.temps[i];     // false for NaN and for lp' = -inf
    lp_s[i] = take ? lpn : lpo;
    taken[i] = take;
    order[i] = i;
    mine += take;
    mine_cold += take && i >= prm.n_hot;
  }
  if (mine) atomicAdd(&s_hits, mine);
  if (mine_cold) atomicAdd(&s_cold_hits, mine_cold);
  __syncthreads();

  // ---- temperature exchange: kPtSweeps sweeps of n random slots -----------------------------------------
  // Per sweep, all warps first draw the slots (Philox, log, the ladder gap of the slot) into a scratch array and
  // mark, per window of 32 consecutive slots, the lanes whose slot touches a chain that an EARLIER slot of the
  // window touches.  Warp 0 then walks through the windows in order: the unmarked lanes commute with everything
  // before them in the window and with each other, so they are applied at once; the marked ones (usually none,
  // sometimes one or two) follow one by one in lane order.  That is the sequential sweep, slot for slot.
  if (n > 1) {
    const int n_win = (n + 31) / 32;
    for (int sweep = 0; sweep < kPtSweeps; sweep++) {
      for (int win = warp; win < n_win; win += (int)(blockDim.x >> 5)) {
        const int jdx = 32 * win + lane;
        const bool valid = jdx < n;
        uint32_t r[4];
        philox4x32(prm.seed, (uint64_t)prm.k, (uint32_t)jdx, kPtTagSweep + sweep, r);
        int rt = 1 + (int)(u01(r[0], r[1]) * (double)(n - 1));
        rt = min(rt, n - 1);
        bool conflict = false;
        for (int d = 1; d < 32; d++) {
          const int other = __shfl_up_sync(0xffffffffu, rt, d);
          conflict = conflict || (lane >= d && abs(rt - other) <= 1);
        }
        const unsigned int marked = __ballot_sync(0xffffffffu, conflict && valid);
        if (valid) {
          prm.sw_slot[jdx] = rt;
          prm.sw_logu[jdx] = log(u01(r[2], r[3]));
          prm.sw_gap[jdx] = prm.gap[rt];
        }
        if (lane == 0) win_marked[win] = marked;
      }
      __syncthreads();
      if (warp == 0) {
        auto apply = [&](int rt, double gap, double logu) {
          const int a = order[rt - 1], b = order[rt];
          if ((lp_s[b] - lp_s[a]) * gap > logu) {
            order[rt - 1] = b;
            order[rt] = a;
          }
        };
        // (the draws of window w + 1 are in flight while window w is applied)
        int rt = 1;
        double gap = 0.0, logu = 0.0;
        if (lane < n) { rt = prm.sw_slot[lane]; gap = prm.sw_gap[lane]; logu = prm.sw_logu[lane]; }
        for (int win = 0; win < n_win; win++) {
          const int jn = 32 * (win + 1) + lane;
          int rt_n = 1;
          double gap_n = 0.0, logu_n = 0.0;
          if (jn < n) { rt_n = prm.sw_slot[jn]; gap_n = prm.sw_gap[jn]; logu_n = prm.sw_logu[jn]; }
          const bool valid = 32 * win + lane < n;
          unsigned int marked = win_marked[win];
          if (valid && !((marked >> lane) & 1u)) apply(rt, gap, logu);
          __syncwarp();
          while (marked) {
            const int l = __ffs(marked) - 1;
            marked &= marked - 1;
            if (lane == l) apply(rt, gap, logu);
            __syncwarp();
          }
          rt = rt_n; gap = gap_n; logu = logu_n;
        }
      }
      __syncthreads();
    }
  }

  // ---- the new state, position by position: chain order[pos], moved or not ---------------------------
  for (int idx = tid; idx < n * p; idx += blockDim.x) {
    const int pos = idx / p, d = idx - pos * p;
    const int o = order[pos];
    const double v = taken[o] ? __ldcg(prm.prop + (size_t)o * p + d) : __ldcg(prm.theta_in + (size_t)o * p + d);
    prm.theta_out[idx] = v;
    if (prm.k >= prm.n_tune && pos >= prm.n_hot)
      prm.saved[((size_t)(pos - prm.n_hot) * prm.n_keep + (prm.k - prm.n_tune)) * p + d] = v;
  }
  for (int pos = tid; pos < n; pos += blockDim.x) prm.lp_out[pos] = lp_s[order[pos]];

  // ---- step-size tuning (src/mcmc.py:663-667) ------------------------------------------------------
  if (tid == 0) {
    PtlmcCtl* c = prm.ctl;
    double hits = __ldcg(&c->hits) + (double)s_hits / (double)n;
    if (prm.k < prm.n_tune && prm.k % 10 == 0) {
      const double tau = __ldcg(&c->tau) + 1.0 / sqrt(1.0 + (double)prm.k / 10.0) * (hits / 10.0 - prm.goal);
      c->tau = tau;
      c->stride = ptlmc_stride(tau);
      hits = 0.0;
    }
    c->hits = hits;
    if (prm.k >= prm.n_tune) c->accepted = __ldcg(&c->accepted) + s_cold_hits;
  }
}

}  // namespace gpbt
