// Device-resident affine-invariant ensemble sampler: the walkers, their log-posteriors and the
// whole chain history stay in HBM; one sampler step is a fixed sequence of small kernels around the
// log-posterior path, captured once in a CUDA graph and replayed.
//
// The algorithm is the one the reference drives through emcee.EnsembleSampler with its default
// move (src/mcmc.py:68-92, 372-412; emcee >= 3.1.4 is not vendored in the reference tree, so this
// follows the published algorithm: Goodman & Weare 2010, "stretch move" with a red/blue split):
//   every step the walkers are split at random into two sets; for each set in turn every walker s
//   draws z with density g(z) ~ 1/sqrt(z) on [1/a, a] (z = ((a-1)u+1)^2/a), a partner c from the
//   other set, proposes y = c - (c - s) z and accepts with probability min(1, z^(p-1) pi(y)/pi(s)).
//
// Everything that changes between graph replays (step counter, history pointers, optional
// host-supplied random streams used by the parity tests) lives in one EnsembleCtl block in device
// memory, so the captured launch arguments never go stale.
#pragma once
#include <cstdint>

#include "common.cuh"

namespace gpbt {

struct EnsembleCtl {
  long long step;         // steps taken since the last reset (index into the history)
  long long run_first;    // value of `step` when the current run started (index 0 of ru / rp / perm)
  const double* ru;       // [run_steps, 2, n_half, 2] uniforms (z draw, accept draw) or null -> Philox
  const int* rp;          // [run_steps, 2, n_half] partner indices into the complementary set, or null
  const int* perm_in;     // [run_steps, n_walkers] split permutations, or null -> Philox keys
  double* hist_x;         // [hist_cap, n_walkers, p]
  double* hist_lp;        // [hist_cap, n_walkers]
};

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: no per-walker generator state ------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

__device__ __forceinline__ void philox4x32(uint64_t seed, uint64_t ctr_lo, uint32_t idx, uint32_t tag,
                                           uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr_lo, (uint32_t)(ctr_lo >> 32), idx, tag};
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = c[i];
}

// 53-bit uniform in [0, 1), like numpy's random_sample
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

constexpr uint32_t kTagSplit = 2;   // tags 0 / 1: uniforms of the two half steps; 3 / 4: partner draws

// Split key of walker j: 40 random bits above the walker index, so keys are distinct and "sort by
// (random, index)" is a plain unsigned comparison.
__device__ __forceinline__ unsigned long long split_key(uint64_t seed, long long step, int j) {
  uint32_t r[4];
  philox4x32(seed, (uint64_t)step, (uint32_t)j, kTagSplit, r);
  return ((((unsigned long long)r[0] << 32) | r[1]) & ~0xFFFFFFull) | (unsigned long long)j;
}

// Large ensembles: ranking n keys against each other is O(n^2) (5 ms per step at 262144 walkers), so
// above kEnsembleRankMaxWalkers the split permutation is a keyed bijection of [0, n) instead: a 4-round
// Feistel network on the next power of four, round function = Philox, cycle-walked back into range
// (at most 4 iterations expected).  perm[i] = feistel(i): O(n), no communication between threads.
constexpr int kEnsembleRankMaxWalkers = 2048;
constexpr uint32_t kTagFeistel = 8;   // tags 8..11: the four round functions

__device__ __forceinline__ int split_feistel(uint64_t seed, long long step, int i, int nw) {
  int hbits = 1;
  while ((1ll << (2 * hbits)) < (long long)nw) hbits++;
  const uint32_t mask = (1u << hbits) - 1u;
  uint32_t x = (uint32_t)i;
  do {
    uint32_t l = x >> hbits, r = x & mask;
#pragma unroll 1
    for (uint32_t rd = 0; rd < 4; rd++) {
      uint32_t f[4];
      philox4x32(seed, (uint64_t)step, r, kTagFeistel + rd, f);
      const uint32_t nl = r;
      r = l ^ (f[0] & mask);
      l = nl;
    }
    x = (l << hbits) | r;
  } while (x >= (uint32_t)nw);
  return (int)x;
}

// ---- the three pieces of a step, as device functions shared by the two launch shapes ---------------
//
// Random split of the walkers into two sets: perm[0:n0] is set 0, perm[n0:] set 1.  perm is the
// ranking of one Philox key per walker.  Every CTA builds the whole key table
// in shared memory and ranks its own walkers [j_first, j_first + blockDim.x * j_iters) against it.
__device__ __forceinline__ void split_walkers(const EnsembleCtl* __restrict__ ctl, uint64_t seed, int nw,
                                              int randomize, int* __restrict__ perm, unsigned long long* keys,
                                              int j_first, int j_stride) {
  const long long step = ctl->step;
  if (ctl->perm_in) {
    for (int j = j_first; j < nw; j += j_stride) perm[j] = ctl->perm_in[(step - ctl->run_first) * nw + j];
    return;
  }
  if (!randomize) {   // fixed split: even walkers first, then odd ones (emcee's inds = arange % 2)
    for (int j = j_first; j < nw; j += j_stride) perm[(j & 1) ? (nw + 1) / 2 + (j >> 1) : (j >> 1)] = j;
    return;
  }
  for (int k = threadIdx.x; k < nw; k += blockDim.x) keys[k] = split_key(seed, step, k);
  __syncthreads();
  for (int j = j_first; j < nw; j += j_stride) {
    const unsigned long long kj = keys[j];
    int rank = 0;
#pragma unroll 8
    for (int k = 0; k < nw; k++) rank += keys[k] < kj;
    perm[rank] = j;
  }
}

struct EnsembleBuffers {
  const EnsembleCtl* ctl;
  uint64_t seed;
  double a;
  int nw, p, n_half, randomize;
  int* perm;            // [nw]
  double* x;            // [nw, p] walkers
  double* lp;           // [nw]
  double* q;            // [n_half, p] proposals of the active set
  double* factor;       // [n_half] (p - 1) ln z
  double* u_acc;        // [n_half] accept draws
  const double* lp_new; // [n_half] log-posterior of the proposals
  int* notpd_call;      // non-PD covariances met by the last log-posterior call (reset here)
  long long* notpd_total;
  long long* accepted;  // [nw]
};

// Stretch proposal of the i-th walker of the active set:  q[i, :] = c - (c - s) z,  factor = (p-1) ln z
__device__ __forceinline__ void propose_one(const EnsembleBuffers& b, int half, int i) {
  const int nw = b.nw, p = b.p, n0 = (nw + 1) / 2;
  const int ns = half == 0 ? n0 : nw - n0, nc = nw - ns;
  const int off_s = half == 0 ? 0 : n0, off_c = half == 0 ? n0 : 0;
  const long long step = b.ctl->step;
  double uz, ua;
  int r;
  if (b.ctl->ru) {
    const long long row = ((step - b.ctl->run_first) * 2 + half) * b.n_half + i;
    uz = b.ctl->ru[2 * row];
    ua = b.ctl->ru[2 * row + 1];
    r = b.ctl->rp[row];
  } else {
    uint32_t w[4], v[4];
    philox4x32(b.seed, (uint64_t)step, (uint32_t)i, (uint32_t)half, w);
    philox4x32(b.seed, (uint64_t)step, (uint32_t)i, 3u + (uint32_t)half, v);
    uz = u01(w[0], w[1]);
    ua = u01(w[2], w[3]);
    // unbiased to 2^-64 * nc: high word of a 64 x 64 -> 128-bit product
    const unsigned long long r64 = ((unsigned long long)v[0] << 32) | v[1];
    r = (int)__umul64hi(r64, (unsigned long long)nc);
  }
  const double t = __dadd_rn(__dmul_rn(b.a - 1.0, uz), 1.0);
  const double z = __ddiv_rn(__dmul_rn(t, t), b.a);
  const double* s = b.x + (size_t)b.perm[off_s + i] * p;
  const double* c = b.x + (size_t)b.perm[off_c + r] * p;
  // rounded operation by operation (no FMA contraction) so that a host restatement reproduces the
  // proposal bit for bit
  for (int d = 0; d < p; d++) b.q[(size_t)i * p + d] = __dsub_rn(c[d], __dmul_rn(__dsub_rn(c[d], s[d]), z));
  b.factor[i] = (double)(p - 1) * log(z);
  b.u_acc[i] = ua;
}

// Metropolis accept of the i-th walker of the active set
__device__ __forceinline__ void accept_one(const EnsembleBuffers& b, int half, int i) {
  const int n0 = (b.nw + 1) / 2, p = b.p;
  const int j = b.perm[(half == 0 ? 0 : n0) + i];
  const double lnew = b.lp_new[i];
  const double diff = b.factor[i] + lnew - b.lp[j];
  if (diff > log(b.u_acc[i])) {   // false for NaN (-inf against -inf): stay, as emcee does
    for (int d = 0; d < p; d++) b.x[(size_t)j * p + d] = b.q[(size_t)i * p + d];
    b.lp[j] = lnew;
    b.accepted[j] += 1;
  }
}

// the log-posterior call in front counted its non-PD covariances into notpd_call: fold and reset
// (this replaces a memset node per call)
__device__ __forceinline__ void fold_notpd(const EnsembleBuffers& b) {
  *b.notpd_total += *b.notpd_call;
  *b.notpd_call = 0;
}

__device__ __forceinline__ int active_size(int nw, int half) {
  const int n0 = (nw + 1) / 2;
  return half == 0 ? n0 : nw - n0;
}

// ---- launch shape 1: any ensemble size, one kernel per piece ----------------------------------------
// The split of a large ensemble is two kernels: the Philox keys go to global memory once, then
// CTAs of 128 walkers rank their keys against all the others, staged through shared memory in
// tiles (n_walkers^2 comparisons spread over n_walkers / 128 CTAs).
constexpr int kRankThreads = 128, kRankTile = 2048;

__global__ void ensemble_keys_kernel(const EnsembleCtl* __restrict__ ctl, uint64_t seed, int nw, int randomize,
                                     unsigned long long* __restrict__ keys, int* __restrict__ perm) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nw) return;
  const long long step = ctl->step;
  if (ctl->perm_in) {
    perm[j] = ctl->perm_in[(step - ctl->run_first) * nw + j];
  } else if (!randomize) {
    perm[(j & 1) ? (nw + 1) / 2 + (j >> 1) : (j >> 1)] = j;
  } else if (nw > kEnsembleRankMaxWalkers) {
    perm[j] = split_feistel(seed, step, j, nw);
  } else {
    keys[j] = split_key(seed, step, j);
  }
}

__global__ void __launch_bounds__(kRankThreads) ensemble_rank_kernel(const EnsembleCtl* __restrict__ ctl, int nw,
                                                                      int randomize,
                                                                      const unsigned long long* __restrict__ keys,
                                                                      int* __restrict__ perm) {
  __shared__ unsigned long long tile[kRankTile];
  if (ctl->perm_in || !randomize || nw > kEnsembleRankMaxWalkers) return;   // perm was written by ensemble_keys_kernel
  const int j = blockIdx.x * kRankThreads + threadIdx.x;
  const unsigned long long kj = j < nw ? keys[j] : 0ull;
  int rank = 0;
  for (int k0 = 0; k0 < nw; k0 += kRankTile) {
    const int kn = min(kRankTile, nw - k0);
    __syncthreads();
    for (int k = threadIdx.x; k < kn; k += kRankThreads) tile[k] = keys[k0 + k];
    __syncthreads();
#pragma unroll 16
    for (int k = 0; k < kn; k++) rank += tile[k] < kj;
  }
  if (j < nw) perm[rank] = j;
}

__global__ void ensemble_propose_kernel(const EnsembleBuffers b, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < active_size(b.nw, half)) propose_one(b, half, i);
}

__global__ void ensemble_accept_kernel(const EnsembleBuffers b, int half) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) fold_notpd(b);
  if (i < active_size(b.nw, half)) accept_one(b, half, i);
}

// Append the ensemble to the history.  The last CTA to finish advances the step counter (every CTA
// has read it by then), so a step stays a fixed kernel sequence with no host involvement.
__global__ void __launch_bounds__(256) ensemble_record_kernel(EnsembleCtl* __restrict__ ctl, int nw, int p,
                                                              const double* __restrict__ x,
                                                              const double* __restrict__ lp,
                                                              unsigned int* __restrict__ done) {
  const long long step = ctl->step;
  double* hx = ctl->hist_x + (size_t)step * nw * p;
  double* hl = ctl->hist_lp + (size_t)step * nw;
  const int stride = gridDim.x * blockDim.x, t = blockIdx.x * blockDim.x + threadIdx.x;
  for (int k = t; k < nw * p; k += stride) hx[k] = x[k];
  for (int k = t; k < nw; k += stride) hl[k] = lp[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1u) == gridDim.x - 1) {
      *done = 0;
      ctl->step = step + 1;
    }
  }
}

// rows [first, first + n) of the proposal buffer into dst (row index clamped to the active set: callers
// that shard the proposals over ranks pad their last block by repeating the last row)
__global__ void ensemble_gather_rows_kernel(const double* __restrict__ q, int64_t ns, int p, int64_t first, int64_t n,
                                            double* __restrict__ dst) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * p) return;
  const int64_t r = idx / p;
  const int d = (int)(idx - r * p);
  const int64_t src = min(first + r, ns - 1);
  dst[idx] = q[src * p + d];
}

// ---- launch shape 2: small ensembles (a step is launch-latency bound), one CTA, three kernels -------
//   begin = split + propose(0);   mid = accept(0) + propose(1);   end = accept(1) + record
constexpr int kEnsembleFusedMaxWalkers = 2048;

__global__ void __launch_bounds__(1024) ensemble_begin_kernel(const EnsembleBuffers b) {
  extern __shared__ unsigned long long keys[];
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  split_walkers(b.ctl, b.seed, b.nw, b.randomize, b.perm, keys, threadIdx.x, blockDim.x);
  __syncthreads();   // perm (global) is read by other threads of this CTA below
  for (int i = threadIdx.x; i < active_size(b.nw, 0); i += blockDim.x) propose_one(b, 0, i);
}

__global__ void __launch_bounds__(1024) ensemble_mid_kernel(const EnsembleBuffers b) {
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  if (threadIdx.x == 0) fold_notpd(b);
  for (int i = threadIdx.x; i < active_size(b.nw, 0); i += blockDim.x) accept_one(b, 0, i);
  __syncthreads();   // the second set draws partners from the updated first set
  for (int i = threadIdx.x; i < active_size(b.nw, 1); i += blockDim.x) propose_one(b, 1, i);
}

__global__ void __launch_bounds__(1024) ensemble_end_kernel(const EnsembleBuffers b, EnsembleCtl* ctl) {
  pdl_launch_dependents();
  pdl_wait_prior_grids();
  if (threadIdx.x == 0) fold_notpd(b);
  for (int i = threadIdx.x; i < active_size(b.nw, 1); i += blockDim.x) accept_one(b, 1, i);
  __syncthreads();
  const long long step = ctl->step;
  double* hx = ctl->hist_x + (size_t)step * b.nw * b.p;
  double* hl = ctl->hist_lp + (size_t)step * b.nw;
  for (int k = threadIdx.x; k < b.nw * b.p; k += blockDim.x) hx[k] = b.x[k];
  for (int k = threadIdx.x; k < b.nw; k += blockDim.x) hl[k] = b.lp[k];
  __syncthreads();   // everyone has read ctl->step
  if (threadIdx.x == 0) ctl->step = step + 1;
}

}  // namespace gpbt
