// C ABI of libgpbt_b200.so (declared in include/gpbt.h): handle management, state upload with the
// padding the kernels want, launch configuration and the host-buffer convenience entry point.
#include <functional>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "../../include/gpbt.h"
#include "backtransform.cuh"
#include "chol_loglike.cuh"
#include "chol_staged.cuh"
#include "chol_fused.cuh"
#include "chol_stepped.cuh"
#include "chol_warp.cuh"
#include "common.cuh"
#include "ensemble.cuh"
#include "lowrank_loglike.cuh"
#include "param_trafo.cuh"
#include "pc_predict.cuh"

using namespace gpbt;

namespace {

thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};
// bumped whenever a workspace is reallocated: captured CUDA graphs that hold the old pointers are stale
std::atomic<int64_t> g_ws_generation{0};

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess)                                                                      \
      return fail((int)e_, "%s failed: %s (%s:%d)", #x, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

#define LAUNCH_CHECK()                                                                  \
  do {                                                                                  \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                 \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess)                                                              \
      return fail((int)e_, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// 2*log(0 + 1e-16) - 0/scale: the extra_std prior term of the reference with extra_std == 0
// (src/mcmc.py:199, 281, 296-297)
const double kSysConst = 2.0 * std::log(0.0 + 1e-16);

template <typename T>
int upload(T** dst, const std::vector<T>& v) {
  CU(cudaMalloc(dst, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty()) CU(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

constexpr int kMaxDevices = 64;

int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev;
}

// Properties of a device the launch configuration depends on, queried once per device.
struct DeviceInfo {
  int sm_count = 0;        // multiprocessors
  int smem_optin = 0;      // largest dynamic shared memory a block may opt in to
  int smem_per_sm = 0;     // shared memory of one multiprocessor (all resident blocks together)
  int l2_bytes = 0;
  bool ready = false;
};

const DeviceInfo& device_info() {
  static DeviceInfo table[kMaxDevices];
  static DeviceInfo overflow;
  static std::mutex mtx;
  const int dev = current_device();
  DeviceInfo& d = (dev >= 0 && dev < kMaxDevices) ? table[dev] : overflow;
  if (!d.ready || &d == &overflow) {
    std::lock_guard<std::mutex> lock(mtx);
    cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&d.smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    cudaDeviceGetAttribute(&d.l2_bytes, cudaDevAttrL2CacheSize, dev);
    d.ready = true;
  }
  return d;
}

int max_optin_smem() { return device_info().smem_optin; }

// Tuning overrides.  The environment (GPBT_PC_TILE, GPBT_CHOL, GPBT_LOWRANK_GENERIC, GPBT_NO_ZEROCOPY,
// GPBT_ENSEMBLE_SPLIT_KERNELS, GPBT_FANOUT_MIN_ROWS, GPBT_CHOL_BATCH) is read ONCE, when the library is
// loaded; afterwards gpbt_set_option changes a value (tests, tuning tools).  Nothing on a call path
// touches getenv.
struct Options {
  std::atomic<int> pc_tile{0};              // 0 = automatic, else 8 | 16 | 32
  std::atomic<int> chol{0};                 // 0 = automatic, else 'w'arp | 'b'atch (stepped) | 's'taged | 'c'ta | 'f'used
  std::atomic<int> lowrank_generic{0};
  std::atomic<int> no_zerocopy{0};
  std::atomic<int> ensemble_split_kernels{0};
  std::atomic<int64_t> fanout_min_rows{0};  // 0 = built-in default
  std::atomic<int64_t> chol_batch{0};       // walkers per sub-batch of the fused Cholesky, 0 = automatic
  std::atomic<int64_t> chol_streams{0};     // streams the fused Cholesky spreads its sub-batches over, 0 = automatic
  std::atomic<int64_t> chol_pipe{0};        // dense path: sub-batches kernel (a) is pipelined over, 0 = automatic, 1 = off
  std::atomic<int> chol_lag{0};             // 1: sub-batch b + 1 starts behind the first factor kernel of sub-batch b
  std::atomic<int> chol_prio{0};            // 1: aux streams of the fused Cholesky at the highest stream priority
  std::atomic<int> cf_debug{0};             // record clock64 stamps of the fused Cholesky (tuning tool)
};
Options g_opt;
long long* g_cf_dbg = nullptr;            // managed buffer of the fused Cholesky's timing stamps (option cf_debug)
constexpr size_t kCfDbgBytes = 16 * 32 * 8 * 8 * sizeof(long long);
constexpr int kCfMaxStreams = 5;
// aux streams / events the fused Cholesky spreads its sub-batches over (see launch_chol_fused)
struct FusedStreams {
  cudaStream_t aux[kCfMaxStreams] = {};   // [0] unused: the caller's stream
  cudaEvent_t done[kCfMaxStreams] = {}, fork = nullptr;
  void destroy() {
    for (int i = 0; i < kCfMaxStreams; i++) {
      if (aux[i]) cudaStreamDestroy(aux[i]);
      if (done[i]) cudaEventDestroy(done[i]);
      aux[i] = nullptr; done[i] = nullptr;
    }
    if (fork) cudaEventDestroy(fork);
    fork = nullptr;
  }
};


int set_option_value(const char* key, const char* value) {
  const std::string k = key ? key : "";
  const char* v = (value && value[0]) ? value : nullptr;
  if (k == "pc_tile") {
    const int t = v ? atoi(v) : 0;
    if (t != 0 && t != 8 && t != 16 && t != 32) return GPBT_EINVAL;
    g_opt.pc_tile = t;
  } else if (k == "chol") {
    g_opt.chol = v ? (int)v[0] : 0;
  } else if (k == "lowrank_generic") {
    g_opt.lowrank_generic = v ? atoi(v) : 0;
  } else if (k == "no_zerocopy") {
    g_opt.no_zerocopy = v ? atoi(v) : 0;
  } else if (k == "ensemble_split_kernels") {
    g_opt.ensemble_split_kernels = v ? atoi(v) : 0;
  } else if (k == "fanout_min_rows") {
    g_opt.fanout_min_rows = v ? atoll(v) : 0;
  } else if (k == "chol_batch") {
    g_opt.chol_batch = v ? atoll(v) : 0;
  } else if (k == "chol_streams") {
    g_opt.chol_streams = v ? atoll(v) : 0;
  } else if (k == "chol_pipe") {
    g_opt.chol_pipe = v ? atoll(v) : 0;
  } else if (k == "chol_lag") {
    g_opt.chol_lag = v ? atoi(v) : 0;
  } else if (k == "chol_prio") {
    g_opt.chol_prio = v ? atoi(v) : 0;
  } else if (k == "cf_debug") {
    g_opt.cf_debug = v ? atoi(v) : 0;
  } else {
    return GPBT_EINVAL;
  }
  return 0;
}

struct OptionsFromEnvironment {
  OptionsFromEnvironment() {
    static const char* const names[][2] = {
        {"GPBT_PC_TILE", "pc_tile"}, {"GPBT_CHOL", "chol"}, {"GPBT_LOWRANK_GENERIC", "lowrank_generic"},
        {"GPBT_NO_ZEROCOPY", "no_zerocopy"}, {"GPBT_ENSEMBLE_SPLIT_KERNELS", "ensemble_split_kernels"},
        {"GPBT_FANOUT_MIN_ROWS", "fanout_min_rows"}, {"GPBT_CHOL_BATCH", "chol_batch"},
        {"GPBT_CHOL_STREAMS", "chol_streams"}, {"GPBT_CHOL_PIPE", "chol_pipe"}};
    for (const auto& n : names)
      if (const char* e = getenv(n[0])) set_option_value(n[1], e);
  }
} g_options_from_environment;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device setting: remember the largest size
// configured for each kernel on each device (the kernel is a template argument, so every kernel --
// every instantiation of a kernel template -- has its own table)
template <auto Kernel>
int ensure_dynamic_smem(size_t bytes) {
  static size_t configured[kMaxDevices] = {0};
  const int dev = current_device();
  if (dev >= 0 && dev < kMaxDevices && bytes <= configured[dev]) return 0;
  CU(cudaFuncSetAttribute(Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  // ask for the largest shared-memory carve-out explicitly: kernels launched from a captured graph do
  // not get the per-launch carve-out heuristic of stream launches, and the kernels that come through
  // here are sized for a given number of resident CTAs per SM
  if (bytes > 48 * 1024)
    CU(cudaFuncSetAttribute(Kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  if (dev >= 0 && dev < kMaxDevices) configured[dev] = bytes;
  return 0;
}

// launch with programmatic stream serialisation (see common.cuh); only for kernels that execute
// pdl_wait_prior_grids() before they touch a predecessor's output
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, unsigned threads, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace

struct gpbt_emulator {
  int p, p_pad, n, n_pad, q, q_pad, m, m_ld, kind, flags, device;
  double *Xs, *ell, *c, *sn, *W, *A, *mu, *scale, *Ctrunc;
  double* sig2;           // PCGP kind only
  // optional parameter-function PCA pre-transform: X [N, p_in] -> theta [N, p]
  bool has_trafo;
  int p_in;
  ParamTrafoParams trafo;
  int* keep_dev;
  double* trafo_buf[2 * kPtMaxGroups];
  double* theta;          // workspace [theta_cap, p]
  int64_t theta_cap;
};

struct gpbt_chain {
  std::vector<gpbt_emulator_t> emus;
  std::vector<int> q_off, m_off;
  int p, Q, M, device;
  bool has_lowrank, has_diag;
  // cov_exp does not couple the emulators (and every emulator is PCA-mode): the dense path lets
  // kernel (b) start from base_like[e] = Ctrunc_e + cov_exp[block e], so kernel (c) reads finished
  // covariances and needs no cov_add pass
  bool exp_blockdiag = false;
  std::vector<double*> base_like;
  // fused (b)+(c) dense path (chol_fused.cuh): F = blockdiag(Ctrunc) + cov_exp in the factor's packed
  // panel layout, its diagonal blocks, U^T; per-walker factor / work vectors grown on demand
  bool has_fused = false;
  int Mg = 0, Qp = 0;
  int64_t Lstride = 0, cap_fused = 0;
  double *cf_Fp = nullptr, *cf_Fd = nullptr, *cf_UT = nullptr;
  double *cf_L = nullptr, *cf_dinv = nullptr, *cf_draw = nullptr, *cf_tvec = nullptr, *cf_logdet = nullptr, *cf_tsq = nullptr, *cf_mean = nullptr;
  int* cf_bad = nullptr;
  FusedStreams cf_fs;
  bool lr_separable = false;          // R is block diagonal over the emulators
  std::vector<double*> R_blocks;      // per-emulator q_e x q_e copies of the diagonal blocks of R
  double s_perp, logdetF_half;
  double *lo, *hi, *y_exp, *cov_exp, *R, *c0;
  // workspaces (grown on demand)
  int64_t cap_rows = 0, cap_dense = 0, cap_x = 0, cap_diag = 0;
  double *dmean = nullptr, *dvar = nullptr;
  double *z_mean = nullptr, *z_var = nullptr, *extra = nullptr, *mean = nullptr, *cov = nullptr;
  double *x_dev = nullptr, *lp_dev = nullptr;
  unsigned char* skip = nullptr;
  int* notpd_dev = nullptr;
  cudaStream_t stream = nullptr;
  // The workspaces below are shared by every call on this chain.  Calls may come in on different streams (the
  // chain's own for the host API and the sampler, the caller's for the device API): when the stream changes,
  // the previous one is drained first, so two calls never use the buffers at the same time.
  cudaStream_t last_stream = nullptr;
  bool last_stream_valid = false;
  int64_t ws_bytes = 0;
  // small-batch path: pinned host buffers mapped into the device address space (zero copy)
  double *zc_x_host = nullptr, *zc_x_dev = nullptr;      // [kZeroCopyRows, p]
  double *zc_lp_host = nullptr, *zc_lp_dev = nullptr;    // [kZeroCopyRows] + one slot holding the int counter
};

// Batches up to this many walkers skip the explicit H2D / D2H copies: X is read and lp is written
// by the kernels directly in mapped pinned host memory (a few KB over PCIe), which removes three
// cudaMemcpyAsync calls -- about 20 us of a 52 us call at N = 64.
constexpr int64_t kZeroCopyRows = 512;

extern "C" const char* gpbt_last_error(void) { return g_err.c_str(); }
extern "C" int gpbt_version(void) { return 100; }
extern "C" int64_t gpbt_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// emulator state
// ------------------------------------------------------------------------------------------------
namespace {
int emulator_upload(gpbt_emulator* e, int p, int n, int q, int m, int kernel_kind, int flags, const double* Xtr,
                    const double* ell, const double* c, const double* sn, const double* alpha, const double* Linv,
                    const double* A, const double* mu, const double* scale, const double* Ctrunc);
}

extern "C" int gpbt_emulator_create(gpbt_emulator_t* out, int p, int n, int q, int m, int kernel_kind,
                                    int flags, const double* Xtr, const double* ell, const double* c,
                                    const double* sn, const double* alpha, const double* Linv,
                                    const double* A, const double* mu, const double* scale,
                                    const double* Ctrunc) {
  if (!out || p <= 0 || n <= 0 || q <= 0 || m <= 0 || !Xtr || !ell || !c || !sn || !alpha || !Linv || !mu)
    return fail(GPBT_EINVAL, "gpbt_emulator_create: null or non-positive argument");
  if (kernel_kind != GPBT_KERNEL_RBF && kernel_kind != GPBT_KERNEL_MATERN32)
    return fail(GPBT_EINVAL, "gpbt_emulator_create: unknown kernel kind %d (PCGP emulators: gpbt_emulator_create_pcgp)",
                kernel_kind);
  const bool no_pca = flags & GPBT_FLAG_NO_PCA;
  if (no_pca && (q != m || !scale)) return fail(GPBT_EINVAL, "no-PCA mode needs q == m and scale");
  if (!no_pca && (!A || !Ctrunc)) return fail(GPBT_EINVAL, "PCA mode needs A and Ctrunc");

  gpbt_emulator* e = new gpbt_emulator();
  memset(e, 0, sizeof *e);
  const int rc = emulator_upload(e, p, n, q, m, kernel_kind, flags, Xtr, ell, c, sn, alpha, Linv, A, mu, scale, Ctrunc);
  if (rc) {   // nothing half-built escapes: the error message set by the failing call survives
    gpbt_emulator_destroy(e);
    return rc;
  }
  *out = e;
  return 0;
}

namespace {
int emulator_upload(gpbt_emulator* e, int p, int n, int q, int m, int kernel_kind, int flags, const double* Xtr,
                    const double* ell, const double* c, const double* sn, const double* alpha, const double* Linv,
                    const double* A, const double* mu, const double* scale, const double* Ctrunc) {
  const bool no_pca = flags & GPBT_FLAG_NO_PCA;
  e->p = p; e->n = n; e->q = q; e->m = m; e->kind = kernel_kind; e->flags = flags;
  e->p_pad = (int)round_up(p, 2);
  e->n_pad = (int)round_up(n, 32);
  e->q_pad = (int)round_up(q, 4);
  e->m_ld = m;
  while (e->m_ld % 8 != 4) e->m_ld++;  // conflict-free fragment loads in backtransform_cov_kernel
  cudaGetDevice(&e->device);
  const int P = e->p_pad, NP = e->n_pad;

  // staged design rows: (X_train[i] / ell_j  [true division, as sklearn's X / length_scale],
  // zero pad to p_pad, alpha_j[i], 0)
  std::vector<double> h;
  const int XR = P + 2;
  h.assign((size_t)q * NP * XR, 0.0);
  for (int j = 0; j < q; j++)
    for (int i = 0; i < n; i++) {
      double* row = &h[((size_t)j * NP + i) * XR];
      for (int d = 0; d < p; d++) row[d] = Xtr[(size_t)i * p + d] / ell[(size_t)j * p + d];
      row[P] = alpha[(size_t)j * n + i];
    }
  if (int r = upload(&e->Xs, h)) return r;
  h.assign((size_t)q * P, 1.0);
  for (int j = 0; j < q; j++)
    for (int d = 0; d < p; d++) h[(size_t)j * P + d] = ell[(size_t)j * p + d];
  if (int r = upload(&e->ell, h)) return r;
  h.assign(c, c + q);
  if (int r = upload(&e->c, h)) return r;
  h.assign(sn, sn + q);
  if (int r = upload(&e->sn, h)) return r;
  h.assign((size_t)q * NP * NP, 0.0);
  const bool dense_w = kernel_kind == GPBT_KERNEL_PCGP;   // full rows, not just the lower triangle
  for (int j = 0; j < q; j++)
    for (int i = 0; i < n; i++)
      memcpy(&h[((size_t)j * NP + i) * NP], Linv + ((size_t)j * n + i) * n,
             (size_t)(dense_w ? n : i + 1) * sizeof(double));
  if (int r = upload(&e->W, h)) return r;
  h.assign(mu, mu + m);
  if (int r = upload(&e->mu, h)) return r;
  if (scale) {
    h.assign(scale, scale + m);
    if (int r = upload(&e->scale, h)) return r;
  }
  if (!no_pca) {
    h.assign((size_t)e->q_pad * e->m_ld, 0.0);
    for (int k = 0; k < q; k++) memcpy(&h[(size_t)k * e->m_ld], A + (size_t)k * m, m * sizeof(double));
    if (int r = upload(&e->A, h)) return r;
    h.assign(Ctrunc, Ctrunc + (size_t)m * m);
    if (int r = upload(&e->Ctrunc, h)) return r;
  }
  return 0;
}
}  // namespace

extern "C" int gpbt_emulator_create_pcgp(gpbt_emulator_t* out, int p, int n, int q, int m, int flags,
                                         const double* theta, const double* ell, const double* amp,
                                         const double* off, const double* sig2, const double* pw,
                                         const double* VhT, const double* A, const double* offset,
                                         const double* extra_cov) {
  if (!out || p <= 0 || n <= 0 || q <= 0 || m <= 0 || !theta || !ell || !amp || !off || !sig2 || !pw || !VhT || !A ||
      !offset || !extra_cov)
    return fail(GPBT_EINVAL, "gpbt_emulator_create_pcgp: null or non-positive argument");
  if (flags & GPBT_FLAG_NO_PCA) return fail(GPBT_EINVAL, "gpbt_emulator_create_pcgp: there is no no-PCA mode");
  gpbt_emulator* e = new gpbt_emulator();
  memset(e, 0, sizeof *e);
  int rc = emulator_upload(e, p, n, q, m, GPBT_KERNEL_PCGP, flags, theta, ell, amp, off, pw, VhT, A, offset, nullptr,
                           extra_cov);
  if (!rc) {
    std::vector<double> h(sig2, sig2 + q);
    rc = upload(&e->sig2, h);
  }
  if (rc) {
    gpbt_emulator_destroy(e);
    return rc;
  }
  *out = e;
  return 0;
}

extern "C" int gpbt_emulator_set_param_trafo(gpbt_emulator_t e, int p_in, const int* keep, int n_keep,
                                             int n_groups, const int* kinds, const int* idx, const int* ncomp,
                                             const int* npts, const double* grid_lo, const double* grid_hi,
                                             const double* const* Wt, const double* const* b) {
  if (!e || !keep || n_groups < 1 || n_groups > kPtMaxGroups || !kinds || !idx || !ncomp || !npts || !Wt || !b)
    return fail(GPBT_EINVAL, "gpbt_emulator_set_param_trafo: bad argument");
  int p_out = n_keep;
  for (int g = 0; g < n_groups; g++) p_out += ncomp[g];
  if (p_out != e->p) return fail(GPBT_EINVAL, "param trafo produces %d columns, emulator was trained on %d", p_out, e->p);
  e->has_trafo = false;   // a second call replaces the first transform
  if (e->keep_dev) { cudaFree(e->keep_dev); e->keep_dev = nullptr; }
  for (double*& buf : e->trafo_buf)
    if (buf) { cudaFree(buf); buf = nullptr; }
  ParamTrafoParams& T = e->trafo;
  memset(&T, 0, sizeof T);
  T.p_in = p_in; T.p_out = p_out; T.n_keep = n_keep; T.n_groups = n_groups;
  std::vector<int> hk(keep, keep + n_keep);
  if (int r = upload(&e->keep_dev, hk)) return r;
  T.keep = e->keep_dev;
  int off = n_keep;
  for (int g = 0; g < n_groups; g++) {
    ParamTrafoGroup& G = T.grp[g];
    if (kinds[g] < 0 || kinds[g] > 2) return fail(GPBT_EINVAL, "unknown parametrisation kind %d", kinds[g]);
    G.kind = kinds[g];
    G.nargs = kinds[g] == 0 ? 4 : 3;
    for (int i = 0; i < G.nargs; i++) {
      G.idx[i] = idx[g * kPtMaxArgs + i];
      if (G.idx[i] < 0 || G.idx[i] >= p_in) return fail(GPBT_EINVAL, "param trafo column %d outside X", G.idx[i]);
    }
    G.npts = npts[g]; G.ncomp = ncomp[g]; G.out_off = off; G.g0 = grid_lo[g]; G.g1 = grid_hi[g];
    off += ncomp[g];
    std::vector<double> hw(Wt[g], Wt[g] + (size_t)G.ncomp * G.npts), hb(b[g], b[g] + G.ncomp);
    if (int r = upload(&e->trafo_buf[2 * g], hw)) return r;
    if (int r = upload(&e->trafo_buf[2 * g + 1], hb)) return r;
    G.Wt = e->trafo_buf[2 * g];
    G.b = e->trafo_buf[2 * g + 1];
  }
  e->p_in = p_in;
  e->has_trafo = true;
  return 0;
}

extern "C" int gpbt_emulator_input_dim(gpbt_emulator_t e) { return e ? (e->has_trafo ? e->p_in : e->p) : 0; }

extern "C" int gpbt_emulator_destroy(gpbt_emulator_t e) {
  if (!e) return 0;
  if (e->keep_dev) cudaFree(e->keep_dev);
  if (e->theta) cudaFree(e->theta);
  for (double* p : e->trafo_buf)
    if (p) cudaFree(p);
  double* ptrs[] = {e->Xs, e->ell, e->c, e->sn, e->W, e->A, e->mu, e->scale, e->Ctrunc, e->sig2};
  for (double* p : ptrs)
    if (p) cudaFree(p);
  delete e;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// kernel (a)
// ------------------------------------------------------------------------------------------------
namespace {

template <int TW, int KIND, int P2>
int launch_pc_predict(const PcPredictParams& prm, cudaStream_t st) {
  const size_t smem = pc_predict_smem_bytes<TW>(prm.n_pad, prm.p_pad);
  if (int r = ensure_dynamic_smem<pc_predict_kernel<TW, KIND, P2>>(smem)) return r;
  dim3 grid((unsigned)((prm.N + TW - 1) / TW), (unsigned)prm.q);
  CU(launch_pdl(pc_predict_kernel<TW, KIND, P2>, grid, kPcThreads, smem, st, prm));
  LAUNCH_CHECK();
  return 0;
}

// compile-time parameter-count specialisations (p_pad = 2 * P2 <= 24); generic p otherwise
template <int TW, int KIND>
int launch_pc_predict_p(const PcPredictParams& prm, cudaStream_t st) {
  switch (prm.p_pad / 2) {
#define GPBT_P2(P) case P: return launch_pc_predict<TW, KIND, P>(prm, st);
    GPBT_P2(1) GPBT_P2(2) GPBT_P2(3) GPBT_P2(4) GPBT_P2(5) GPBT_P2(6)
    GPBT_P2(7) GPBT_P2(8) GPBT_P2(9) GPBT_P2(10) GPBT_P2(11) GPBT_P2(12)
#undef GPBT_P2
    default: return launch_pc_predict<TW, KIND, 0>(prm, st);
  }
}

// Walker-tile width.  16 walkers x 2 resident CTAs per SM is the default for large batches: one
// CTA's phase 1 (distance + exp, issue/latency bound) overlaps the other's phase 2 (DMMA bound),
// measured 0.86 ms vs 0.94 ms for one 32-wide CTA per SM at config 2.  32 is used when two 16-wide
// CTAs do not fit in shared memory; 8 for small batches (more CTAs, lower latency).
// GPBT_PC_TILE=8|16|32 overrides (tuning / tests).
// tile_rows: the batch size the tile width is chosen for (a caller that splits one batch into sub-batches
// passes the size of the whole, so that the split does not change a single bit of the result).
template <int KIND>
int dispatch_pc_predict(const PcPredictParams& prm, cudaStream_t st, int64_t tile_rows) {
  const DeviceInfo& di = device_info();
  const size_t limit = (size_t)di.smem_optin;
  const size_t per_sm = (size_t)di.smem_per_sm - 2048;
  int tw = 16;
  if (2 * (pc_predict_smem_bytes<16>(prm.n_pad, prm.p_pad) + 1024) > per_sm &&
      pc_predict_smem_bytes<32>(prm.n_pad, prm.p_pad) <= limit)
    tw = 32;
  if (tw == 16 && pc_predict_smem_bytes<16>(prm.n_pad, prm.p_pad) > limit) tw = 8;
  if (tw == 8 && pc_predict_smem_bytes<8>(prm.n_pad, prm.p_pad) > limit)
    return fail(GPBT_ESHAPE, "pc_predict: n = %d design points do not fit in shared memory", prm.n);
  const int64_t want = 2 * di.sm_count;
  while (tw > 8 && ((tile_rows + tw - 1) / tw) * prm.q < want) tw >>= 1;
  if (const int v = g_opt.pc_tile.load()) tw = v;
  if (tw == 32) return launch_pc_predict_p<32, KIND>(prm, st);
  if (tw == 16) return launch_pc_predict_p<16, KIND>(prm, st);
  return launch_pc_predict<8, KIND, 0>(prm, st);
}

int run_pc_predict(gpbt_emulator_t e, const double* X, const double* extra, double* zm, double* zv,
                   int64_t ldz, int64_t N, cudaStream_t st, int64_t tile_rows = 0) {
  if (N <= 0) return 0;
  if (tile_rows < N) tile_rows = N;
  if (e->has_trafo) {
    if (N > e->theta_cap) {
      // (stream-ordered work that still reads the old buffer has been enqueued before this free;
      // cudaFree synchronises the device)
      if (e->theta) cudaFree(e->theta);
      e->theta = nullptr;
      const int64_t cap = std::max<int64_t>(N, 2 * e->theta_cap);
      e->theta_cap = 0;
      CU(cudaMalloc(&e->theta, (size_t)cap * e->p * sizeof(double)));
      e->theta_cap = cap;
      g_ws_generation++;
    }
    ParamTrafoParams T = e->trafo;
    T.X = X; T.theta = e->theta; T.N = N;
    param_trafo_kernel<<<(unsigned)((N * 32 + 127) / 128), 128, 0, st>>>(T);
    LAUNCH_CHECK();
    X = e->theta;
  }
  PcPredictParams prm;
  prm.X = X; prm.extra = extra; prm.Xs = e->Xs; prm.ell = e->ell; prm.c = e->c; prm.sn = e->sn;
  prm.W = e->W; prm.sig2 = e->sig2; prm.z_mean = zm; prm.z_var = zv; prm.ldz = ldz; prm.N = N;
  prm.p = e->p; prm.p_pad = e->p_pad; prm.n = e->n; prm.n_pad = e->n_pad; prm.q = e->q;
  switch (e->kind) {
    case GPBT_KERNEL_RBF: return dispatch_pc_predict<0>(prm, st, tile_rows);
    case GPBT_KERNEL_MATERN32: return dispatch_pc_predict<1>(prm, st, tile_rows);
    default: return dispatch_pc_predict<2>(prm, st, tile_rows);
  }
}

int run_backtransform(gpbt_emulator_t e, const double* zm, const double* zv, int64_t ldz, double* mean,
                      int64_t ld_mean, double* cov, int64_t ld_cov, int64_t col_off, int64_t N,
                      cudaStream_t st, double* var_diag = nullptr, const double* base = nullptr) {
  if (N <= 0) return 0;
  BacktransformParams prm;
  prm.var_diag = var_diag;
  prm.z_mean = zm; prm.z_var = zv; prm.A = e->A; prm.mu = e->mu; prm.scale = e->scale;
  prm.Ctrunc = base ? base : e->Ctrunc; prm.mean = mean; prm.cov = cov; prm.ldz = ldz; prm.ld_mean = ld_mean;
  prm.ld_cov = ld_cov; prm.col_off = col_off; prm.N = N; prm.q = e->q; prm.m = e->m; prm.m_ld = e->m_ld;
  prm.flags = e->flags;
  backtransform_mean_kernel<<<(unsigned)N, 128, 2 * e->q * sizeof(double), st>>>(prm);
  LAUNCH_CHECK();
  const bool diag = e->flags & (GPBT_FLAG_NO_PCA | GPBT_FLAG_EXP_DIAG);
  if (cov != nullptr && !diag) {
    const size_t smem = backtransform_smem_bytes(e->q_pad, e->m_ld);
    if (smem > (size_t)max_optin_smem())
      return fail(GPBT_ESHAPE, "backtransform: q*m = %d*%d does not fit in shared memory", e->q, e->m);
    if (int r = ensure_dynamic_smem<backtransform_cov_kernel>(smem)) return r;
    const int64_t items = ((N + kBtGroup - 1) / kBtGroup) * bt_items_per_walker(e->m);   // one per warp
    const DeviceInfo& di = device_info();
    int per_sm = (int)std::min<size_t>(4, ((size_t)di.smem_per_sm - 4096) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    const unsigned grid = (unsigned)std::min<int64_t>((items + kBtWarps - 1) / kBtWarps, (int64_t)di.sm_count * per_sm);
    backtransform_cov_kernel<<<grid, kBtThreads, smem, st>>>(prm, e->q_pad);
    LAUNCH_CHECK();
  }
  return 0;
}

// Work vectors of the stepped Cholesky (Dinv, t, log-determinant, |t|^2, non-PD flag per walker): one
// grow-only buffer per (device, stream) -- calls on different streams may overlap on the GPU, calls on
// one stream cannot.
struct SteppedCache {
  double* buf = nullptr;
  size_t bytes = 0;
};
std::map<std::pair<int, cudaStream_t>, SteppedCache> g_stepped;
std::mutex g_stepped_mutex;

void release_stepped_buffer(int device, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_stepped_mutex);
  auto it = g_stepped.find(std::make_pair(device, st));
  if (it == g_stepped.end()) return;
  if (it->second.buf) cudaFree(it->second.buf);
  g_stepped.erase(it);
}

int run_chol_stepped(const CholParams& prm, cudaStream_t st) {
  const int m = prm.m;
  const int64_t N = prm.N;
  const size_t need = (size_t)N * ((size_t)m + kSpNB * kSpNB + 3) * sizeof(double);
  SteppedCache c;
  {
    std::lock_guard<std::mutex> lock(g_stepped_mutex);
    SteppedCache& slot = g_stepped[std::make_pair(current_device(), st)];
    if (need > slot.bytes) {
      // (work of earlier calls on this stream that still uses the old buffer has been enqueued before this
      // free; cudaFree synchronises the device)
      if (slot.buf) cudaFree(slot.buf);
      slot.buf = nullptr;
      slot.bytes = 0;
      CU(cudaMalloc(&slot.buf, need));
      slot.bytes = need;
      g_ws_generation++;
    }
    c = slot;
  }
  SteppedWork wk;
  wk.dinv = c.buf;                                  // first: 16-byte aligned rows for cp.async
  wk.tvec = wk.dinv + (size_t)N * kSpNB * kSpNB;
  wk.logdet = wk.tvec + (size_t)N * m;
  wk.tsq = wk.logdet + N;
  wk.bad = reinterpret_cast<int*>(wk.tsq + N);
  if (prm.cov_add != nullptr) {
    chol_step_add_kernel<<<dim3(8, (unsigned)N), 256, 0, st>>>(prm);
    LAUNCH_CHECK();
  }
  if (int r = ensure_dynamic_smem<chol_step_diag_kernel>(chol_step_diag_smem_bytes(m))) return r;
  if (int r = ensure_dynamic_smem<chol_step_below_kernel<2>>(chol_step_below_smem_bytes<2>())) return r;
  if (int r = ensure_dynamic_smem<chol_step_below_kernel<1>>(chol_step_below_smem_bytes<1>())) return r;
  for (int J = 0; J < m; J += kSpNB) {
    const bool last = J + kSpNB >= m;
    CU(launch_pdl(chol_step_diag_kernel, dim3((unsigned)N), kSpThreads, chol_step_diag_smem_bytes(m), st, prm, wk, J,
                  last ? 1 : 0));
    LAUNCH_CHECK();
    if (!last) {
      // 64-row tiles, or 32-row tiles where that trims the padded part of the row range
      const int rows = m - J - kSpNB;
      if ((rows + 31) / 32 * 32 < (rows + 63) / 64 * 64) {
        const dim3 grid((unsigned)((rows + 31) / 32), (unsigned)N);
        CU(launch_pdl(chol_step_below_kernel<1>, grid, kSpThreads, chol_step_below_smem_bytes<1>(), st, prm, wk, J));
      } else {
        const dim3 grid((unsigned)((rows + 63) / 64), (unsigned)N);
        CU(launch_pdl(chol_step_below_kernel<2>, grid, kSpThreads, chol_step_below_smem_bytes<2>(), st, prm, wk, J));
      }
      LAUNCH_CHECK();
    }
  }
  return 0;
}

int run_chol_fused_dense(const CholParams& cp, cudaStream_t st);   // (defined with the fused launcher below)

int run_chol(const double* mean, const double* y_exp, double* cov, const double* cov_add, double* lp,
             int* n_notpd, const unsigned char* skip, double notpd_value, double add_const, int64_t N, int m,
             cudaStream_t st) {
  if (N <= 0) return 0;
  CholParams prm;
  prm.mean = mean; prm.y_exp = y_exp; prm.cov = cov; prm.cov_add = cov_add; prm.lp = lp;
  prm.n_notpd = n_notpd; prm.skip = skip; prm.notpd_value = notpd_value; prm.add_const = add_const;
  prm.N = N; prm.m = m;
  // Small matrices: warp-per-walker kernel, nine walkers resident per SM (their matrices stay in
  // L2: 1332 x 8m^2 bytes <= ~64 MB).  Larger ones: panel-synchronous kernels over the whole batch,
  // or CTA-per-walker for small batches.  GPBT_CHOL=warp|batch|staged|cta overrides.
  const int which = g_opt.chol.load();
  const size_t wsmem = chol_warp_smem_bytes(m);
  bool use_warp = m <= 80;
  if (which == 'w') use_warp = true;
  if (which == 'c' || which == 's' || which == 'b' || which == 'f') use_warp = false;
  if (use_warp && wsmem <= (size_t)max_optin_smem()) {
    if (int r = ensure_dynamic_smem<chol_warp_kernel>(wsmem)) return r;
    chol_warp_kernel<<<(unsigned)N, 32, wsmem, st>>>(prm);
    LAUNCH_CHECK();
    return 0;
  }
  // stepped variant: all walkers advance panel by panel (two launches per 32 columns).  Faster than the
  // per-walker kernels once there are enough walkers to fill the machine per launch (measured at
  // m = 300: 0.92 vs 0.96 ms at N = 512, 9.4 vs 11.2 ms at N = 8192); below that its 19 dependent
  // launches cost more than one latency-bound kernel.  GPBT_CHOL=batch / staged / cta force a variant.
  // batches: the fused kernels (chol_fused.cuh) with the covariances as a dense source -- packed factor in
  // a work buffer (cov is left untouched), one launch per 32-column panel + a factor kernel; before them
  // the stepped kernels (two launches per panel, in place), still selectable
  if (which ? which == 'f' : N >= 256) return run_chol_fused_dense(prm, st);
  const bool aligned_rows = (m % 2 == 0) && ((reinterpret_cast<uintptr_t>(cov) & 15) == 0);
  const bool want_stepped = which ? which == 'b' : false;
  if (aligned_rows && want_stepped) return run_chol_stepped(prm, st);
  // staged kernel (operand stream through a cp.async ring): needs 16-byte aligned rows and its
  // fixed block assignment covers m <= 352
  const size_t ssmem = chol_staged_smem_bytes(m);
  const bool can_stage = (m % 2 == 0) && m <= kCsMaxM && ((reinterpret_cast<uintptr_t>(cov) & 15) == 0) &&
                         ssmem <= (size_t)max_optin_smem();
  if (can_stage && which != 'c') {
    if (int r = ensure_dynamic_smem<chol_staged_kernel>(ssmem)) return r;
    chol_staged_kernel<<<(unsigned)N, kChThreads, ssmem, st>>>(prm);
    LAUNCH_CHECK();
    return 0;
  }
  const size_t smem = chol_smem_bytes(m);
  if (smem > (size_t)max_optin_smem())
    return fail(GPBT_ESHAPE, "mvn_loglike: m = %d observables exceed the shared-memory panel", m);
  if (int r = ensure_dynamic_smem<chol_loglike_kernel>(smem)) return r;
  chol_loglike_kernel<<<(unsigned)N, kChThreads, smem, st>>>(prm);
  LAUNCH_CHECK();
  return 0;
}

// bounds mask for the dense path: skip[w] = outside, lp[w] = oob_value there
__global__ void bounds_mask_kernel(const double* __restrict__ X, const double* __restrict__ lo,
                                   const double* __restrict__ hi, int p, int64_t N, double oob,
                                   unsigned char* __restrict__ skip, double* __restrict__ lp) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= N) return;
  bool ok = true;
  for (int d = 0; d < p; d++) {
    const double x = X[w * p + d];
    ok = ok && (x > lo[d]) && (x < hi[d]);
  }
  skip[w] = ok ? 0 : 1;
  if (!ok) lp[w] = oob;
}

// Diagonal path: every emulator is no-PCA / exp-diag (diagonal model covariance, src/emulator.py:
// 588-601) and expdata_cov is diagonal (as _read_in_exp_data_pickle builds it, src/mcmc.py:318-322):
// the Cholesky of mvn_loglike (src/mcmc.py:23-65) degenerates to element-wise work.
//   lp = -1/2 sum_o dy_o^2 / c_o - 1/2 sum_o log c_o + const,  c_o = var_o + sigma_exp_o^2
// One warp per walker; applies the bounds mask like the other paths.
__global__ void diag_loglike_kernel(const double* __restrict__ X, const double* __restrict__ lo,
                                    const double* __restrict__ hi, int p, const double* __restrict__ mean,
                                    const double* __restrict__ var, const double* __restrict__ y_exp,
                                    const double* __restrict__ cov_exp, int M, int64_t N, double oob, double add_const,
                                    double* __restrict__ lp, int* __restrict__ n_notpd) {
  const int lane = threadIdx.x & 31;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= N) return;
  bool ok = true;
  for (int d = lane; d < p; d += 32) {
    const double x = X[w * p + d];
    ok = ok && (x > lo[d]) && (x < hi[d]);
  }
  if (!__all_sync(0xffffffffu, ok)) {
    if (lane == 0) lp[w] = oob;
    return;
  }
  double quad = 0.0, logdet = 0.0;
  bool pd = true;
  for (int o = lane; o < M; o += 32) {
    const double c = var[w * M + o] + cov_exp[(size_t)o * M + o];
    const double dy = mean[w * M + o] - y_exp[o];
    pd = pd && (c > 0.0);
    quad += dy * dy / c;
    logdet += log(c);
  }
  quad = warp_sum(quad);
  logdet = warp_sum(logdet);
  pd = __all_sync(0xffffffffu, pd);
  if (lane == 0) {
    if (!pd) {
      lp[w] = oob;
      if (n_notpd) atomicAdd(n_notpd, 1);
    } else {
      lp[w] = -0.5 * quad - 0.5 * logdet + add_const;
    }
  }
}

// copy a finished result vector into the peer-mapped buffers of the other GPUs (paths whose last
// kernel does not store to peers itself)
struct PeerList {
  double* p[kMaxPeers];
  int n;
};
__global__ void scatter_to_peers_kernel(const double* __restrict__ src, PeerList peers, int64_t off, int64_t N) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= N) return;
  const double v = src[w];
  for (int r = 0; r < peers.n; r++) peers.p[r][off + w] = v;
}

// extra_std_arr = extra_std * X[:, -1]   (src/mcmc.py:157)
__global__ void extra_std_kernel(const double* __restrict__ X, int p, int64_t N, double scale,
                                 double* __restrict__ out) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < N) out[w] = scale * X[w * p + p - 1];
}

}  // namespace

__global__ void debug_exp_neg_kernel(const double* __restrict__ x, double* __restrict__ y, int64_t n) {
  __shared__ double2 tab[64];
  if (threadIdx.x < 64) tab[threadIdx.x] = kExpTable[threadIdx.x];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = exp_neg(x[i], tab);
}

extern "C" int gpbt_debug_fused_read(gpbt_chain_t ch, int what, int64_t w, double* dst_host, int64_t count) {
  if (!ch || !dst_host || w < 0 || count < 0) return fail(GPBT_EINVAL, "gpbt_debug_fused_read: bad argument");
  const double* src = nullptr;
  switch (what) {
    case 0: src = ch->cf_L + (size_t)w * ch->Lstride; break;
    case 1: src = ch->cf_tvec + (size_t)w * ch->Mg; break;
    case 2: src = ch->cf_dinv + (size_t)w * 1024; break;
    case 3: src = ch->cf_draw + (size_t)w * 1024; break;
    case 4: src = ch->z_var + (size_t)w * ch->Q; break;
    case 5: src = ch->cf_mean + (size_t)w * ch->M; break;
    case 6: src = ch->cf_logdet + w; break;
    case 7: src = ch->cf_tsq + w; break;
    default: return fail(GPBT_EINVAL, "gpbt_debug_fused_read: unknown buffer %d", what);
  }
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(dst_host, src, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int gpbt_debug_timing_read(void* dst_host, int64_t bytes) {
  if (!dst_host || bytes < 0 || (size_t)bytes > kCfDbgBytes) return fail(GPBT_EINVAL, "gpbt_debug_timing_read: bad argument");
  if (!g_cf_dbg) return fail(GPBT_EINVAL, "gpbt_debug_timing_read: nothing recorded (option cf_debug)");
  CU(cudaDeviceSynchronize());
  memcpy(dst_host, g_cf_dbg, (size_t)bytes);
  return 0;
}

extern "C" int gpbt_debug_exp_neg(const double* x, double* y, int64_t n, void* stream) {
  if (!x || !y || n < 0) return fail(GPBT_EINVAL, "gpbt_debug_exp_neg: bad argument");
  if (n == 0) return 0;
  debug_exp_neg_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int gpbt_pc_predict(gpbt_emulator_t emu, const double* X, const double* extra, double* zm,
                               double* zv, int64_t ldz, int64_t N, void* stream) {
  if (!emu || !X || !zm || !zv || ldz < emu->q || N < 0) return fail(GPBT_EINVAL, "gpbt_pc_predict: bad argument");
  if (emu->device != current_device())
    return fail(GPBT_EINVAL, "emulator lives on device %d, current device is %d", emu->device, current_device());
  return run_pc_predict(emu, X, extra, zm, zv, ldz, N, (cudaStream_t)stream);
}

extern "C" int gpbt_backtransform(gpbt_emulator_t emu, const double* zm, const double* zv, int64_t ldz,
                                  double* mean, int64_t ld_mean, double* cov, int64_t ld_cov,
                                  int64_t col_off, int64_t N, void* stream) {
  if (!emu || !zm || !zv || !mean || N < 0 || ld_mean < col_off + emu->m || (cov && ld_cov < col_off + emu->m))
    return fail(GPBT_EINVAL, "gpbt_backtransform: bad argument");
  // the covariance kernel's grid.y carries the walker index: chunk it
  const int64_t step = 32768;
  for (int64_t s = 0; s < N; s += step) {
    const int64_t nn = std::min(step, N - s);
    int r = run_backtransform(emu, zm + s * ldz, zv + s * ldz, ldz, mean + s * ld_mean, ld_mean,
                              cov ? cov + (size_t)s * ld_cov * ld_cov : nullptr, ld_cov, col_off, nn,
                              (cudaStream_t)stream);
    if (r) return r;
  }
  return 0;
}

extern "C" int gpbt_backtransform_diag(gpbt_emulator_t emu, const double* zm, const double* zv, int64_t ldz,
                                       double* mean, double* var, int64_t ld, int64_t col_off, int64_t N,
                                       void* stream) {
  if (!emu || !zm || !zv || !mean || !var || N < 0 || ld < col_off + emu->m)
    return fail(GPBT_EINVAL, "gpbt_backtransform_diag: bad argument");
  return run_backtransform(emu, zm, zv, ldz, mean, ld, nullptr, 0, col_off, N, (cudaStream_t)stream, var);
}

extern "C" int gpbt_mvn_loglike(const double* mean, const double* y_exp, double* cov, const double* cov_add,
                                double* lp, int* n_notpd, double notpd_value, int64_t N, int m, void* stream) {
  if (!mean || !cov || !lp || N < 0 || m <= 0) return fail(GPBT_EINVAL, "gpbt_mvn_loglike: bad argument");
  return run_chol(mean, y_exp, cov, cov_add, lp, n_notpd, nullptr, notpd_value, 0.0, N, m, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// chain
// ------------------------------------------------------------------------------------------------
namespace {
int chain_build(gpbt_chain* ch, const gpbt_emulator_t* emus, int n_emu, int p, const double* lo, const double* hi,
                const double* y_exp, const double* cov_exp, const double* R, const double* c0, double s_perp,
                double logdetF_half);
}

extern "C" int gpbt_chain_create(gpbt_chain_t* out, const gpbt_emulator_t* emus, int n_emu, int p,
                                 const double* lo, const double* hi, const double* y_exp,
                                 const double* cov_exp, const double* R, const double* c0, double s_perp,
                                 double logdetF_half) {
  if (!out || !emus || n_emu <= 0 || !lo || !hi || !y_exp || !cov_exp)
    return fail(GPBT_EINVAL, "gpbt_chain_create: null argument");
  gpbt_chain* ch = new gpbt_chain();
  const int rc = chain_build(ch, emus, n_emu, p, lo, hi, y_exp, cov_exp, R, c0, s_perp, logdetF_half);
  if (rc) {
    gpbt_chain_destroy(ch);
    return rc;
  }
  *out = ch;
  return 0;
}

namespace {
int chain_build(gpbt_chain* ch, const gpbt_emulator_t* emus, int n_emu, int p, const double* lo, const double* hi,
                const double* y_exp, const double* cov_exp, const double* R, const double* c0, double s_perp,
                double logdetF_half) {
  ch->p = p; ch->Q = 0; ch->M = 0;
  cudaGetDevice(&ch->device);
  bool all_pca = true;
  for (int i = 0; i < n_emu; i++) {
    if (!emus[i] || gpbt_emulator_input_dim(emus[i]) != p)
      return fail(GPBT_EINVAL, "emulator %d: parameter count mismatch", i);
    ch->emus.push_back(emus[i]);
    ch->q_off.push_back(ch->Q);
    ch->m_off.push_back(ch->M);
    ch->Q += emus[i]->q;
    ch->M += emus[i]->m;
    if (emus[i]->flags & (GPBT_FLAG_NO_PCA | GPBT_FLAG_EXP_DIAG)) all_pca = false;
  }
  if (R && !all_pca)
    return fail(GPBT_ENOTAPPLICABLE, "low-rank factors given for a chain with a no-PCA / exp-diag emulator");
  if (R && !c0) return fail(GPBT_EINVAL, "R given without c0");
  ch->has_lowrank = R != nullptr;
  bool all_diag = true;
  for (int i = 0; i < n_emu; i++)
    if (!(emus[i]->flags & (GPBT_FLAG_NO_PCA | GPBT_FLAG_EXP_DIAG))) all_diag = false;
  bool exp_diag_only = true;
  for (int r = 0; r < ch->M && exp_diag_only; r++)
    for (int c2 = 0; c2 < ch->M; c2++)
      if (r != c2 && cov_exp[(size_t)r * ch->M + c2] != 0.0) { exp_diag_only = false; break; }
  ch->has_diag = all_diag && exp_diag_only;
  ch->s_perp = s_perp; ch->logdetF_half = logdetF_half;
  std::vector<double> h;
  h.assign(lo, lo + p); if (int r = upload(&ch->lo, h)) return r;
  h.assign(hi, hi + p); if (int r = upload(&ch->hi, h)) return r;
  h.assign(y_exp, y_exp + ch->M); if (int r = upload(&ch->y_exp, h)) return r;
  h.assign(cov_exp, cov_exp + (size_t)ch->M * ch->M); if (int r = upload(&ch->cov_exp, h)) return r;
  {
    bool bd = all_pca;
    for (int r2 = 0; r2 < ch->M && bd; r2++)
      for (int c2 = 0; c2 < ch->M; c2++) {
        int er = 0, ec = 0;
        while (er + 1 < n_emu && r2 >= ch->m_off[er + 1]) er++;
        while (ec + 1 < n_emu && c2 >= ch->m_off[ec + 1]) ec++;
        if (er != ec && cov_exp[(size_t)r2 * ch->M + c2] != 0.0) { bd = false; break; }
      }
    ch->exp_blockdiag = bd;
    if (bd) {
      for (int e = 0; e < n_emu; e++) {
        const int m0 = ch->m_off[e], me = emus[e]->m;
        std::vector<double> ct((size_t)me * me);
        CU(cudaMemcpy(ct.data(), emus[e]->Ctrunc, ct.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (int a = 0; a < me; a++)
          for (int b2 = 0; b2 < me; b2++) ct[(size_t)a * me + b2] += cov_exp[(size_t)(m0 + a) * ch->M + m0 + b2];
        double* d = nullptr;
        if (int r = upload(&d, ct)) return r;
        ch->base_like.push_back(d);
      }
    }
  }
  if (all_pca) {
    // constants of the fused dense path: F and U on the host, then packed
    const int M = ch->M, Q = ch->Q;
    ch->Mg = (int)round_up(M, 16);
    ch->Qp = (int)round_up(Q, 4);
    ch->Lstride = cf_factor_doubles(M);
    const int Mg = ch->Mg, Qp = ch->Qp, nP = (Mg + kCfNB - 1) / kCfNB;
    std::vector<double> F(cov_exp, cov_exp + (size_t)M * M), UT((size_t)Mg * Qp, 0.0);
    for (int e = 0; e < n_emu; e++) {
      const int m0 = ch->m_off[e], q0 = ch->q_off[e], me = emus[e]->m, qe = emus[e]->q, mld = emus[e]->m_ld;
      std::vector<double> ct((size_t)me * me), a((size_t)emus[e]->q_pad * mld);
      CU(cudaMemcpy(ct.data(), emus[e]->Ctrunc, ct.size() * sizeof(double), cudaMemcpyDeviceToHost));
      CU(cudaMemcpy(a.data(), emus[e]->A, a.size() * sizeof(double), cudaMemcpyDeviceToHost));
      for (int r = 0; r < me; r++)
        for (int c2 = 0; c2 < me; c2++) F[(size_t)(m0 + r) * M + m0 + c2] += ct[(size_t)r * me + c2];
      for (int k = 0; k < qe; k++)
        for (int o = 0; o < me; o++) UT[(size_t)(m0 + o) * Qp + q0 + k] = a[(size_t)k * mld + o];
    }
    std::vector<double> Fp((size_t)std::max<int64_t>(ch->Lstride, 1), 0.0), Fd((size_t)nP * kCfNB * kCfNB, 0.0);
    for (int K = 0; K < nP; K++) {
      for (int r = 0; r < kCfNB; r++)
        for (int c2 = 0; c2 < kCfNB; c2++) {
          const int gr = kCfNB * K + r, gc = kCfNB * K + c2;
          Fd[((size_t)K * kCfNB + r) * kCfNB + c2] = (gr < M && gc < M) ? F[(size_t)gr * M + gc] : (r == c2 ? 1.0 : 0.0);
        }
      const int rows = cf_panel_rows(Mg, K);
      const int64_t off = cf_panel_off(Mg, K);
      for (int sub = 0; sub < 4; sub++)
        for (int rel = 0; rel < rows; rel++)
          for (int c8 = 0; c8 < 8; c8++) {
            const int gr = kCfNB * K + kCfNB + rel, gc = kCfNB * K + 8 * sub + c8;
            if (gr < M && gc < M) Fp[off + ((size_t)sub * rows + rel) * 8 + c8] = F[(size_t)gr * M + gc];
          }
    }
    if (int r = upload(&ch->cf_Fp, Fp)) return r;
    if (int r = upload(&ch->cf_Fd, Fd)) return r;
    if (int r = upload(&ch->cf_UT, UT)) return r;
    ch->has_fused = true;
  }
  ch->R = nullptr; ch->c0 = nullptr;
  if (R) {
    h.assign(R, R + (size_t)ch->Q * ch->Q); if (int r = upload(&ch->R, h)) return r;
    h.assign(c0, c0 + ch->Q); if (int r = upload(&ch->c0, h)) return r;
    // block diagonal over the emulators?  (exact zeros: the host builds R per block in that case)
    bool sep = n_emu > 1;
    for (int a = 0; a < ch->Q && sep; a++)
      for (int b2 = 0; b2 < ch->Q; b2++) {
        int ea = 0, eb = 0;
        while (ea + 1 < n_emu && a >= ch->q_off[ea + 1]) ea++;
        while (eb + 1 < n_emu && b2 >= ch->q_off[eb + 1]) eb++;
        if (ea != eb && R[(size_t)a * ch->Q + b2] != 0.0) { sep = false; break; }
      }
    ch->lr_separable = sep;
    if (sep) {
      for (int e = 0; e < n_emu; e++) {
        const int q0 = ch->q_off[e], qe = emus[e]->q;
        h.assign((size_t)qe * qe, 0.0);
        for (int a = 0; a < qe; a++)
          for (int b2 = 0; b2 < qe; b2++) h[(size_t)a * qe + b2] = R[(size_t)(q0 + a) * ch->Q + q0 + b2];
        double* d = nullptr;
        if (int r = upload(&d, h)) return r;
        ch->R_blocks.push_back(d);
      }
    }
  }
  CU(cudaMalloc(&ch->notpd_dev, sizeof(int)));
  CU(cudaStreamCreateWithFlags(&ch->stream, cudaStreamNonBlocking));
  return 0;
}
}  // namespace

namespace {
void release_stepped_buffer(int device, cudaStream_t st);   // defined next to run_chol_stepped
void release_fused_dense(int device, cudaStream_t st);      // defined next to run_chol_fused_dense
}

extern "C" int gpbt_chain_destroy(gpbt_chain_t ch) {
  if (!ch) return 0;
  if (ch->stream) release_stepped_buffer(ch->device, ch->stream);
  if (ch->stream) release_fused_dense(ch->device, ch->stream);
  ch->cf_fs.destroy();
  for (double* d : ch->R_blocks) cudaFree(d);
  for (double* d : ch->base_like) cudaFree(d);
  if (ch->zc_x_host) cudaFreeHost(ch->zc_x_host);
  if (ch->zc_lp_host) cudaFreeHost(ch->zc_lp_host);
  void* ptrs[] = {ch->lo, ch->hi, ch->y_exp, ch->cov_exp, ch->R, ch->c0, ch->z_mean, ch->z_var, ch->extra,
                  ch->mean, ch->cov, ch->x_dev, ch->lp_dev, ch->skip, ch->notpd_dev, ch->dmean, ch->dvar,
                  ch->cf_Fp, ch->cf_Fd, ch->cf_UT, ch->cf_L, ch->cf_dinv, ch->cf_draw, ch->cf_tvec, ch->cf_logdet, ch->cf_tsq, ch->cf_bad, ch->cf_mean};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (ch->stream) cudaStreamDestroy(ch->stream);
  delete ch;
  return 0;
}

extern "C" int64_t gpbt_chain_workspace_bytes(gpbt_chain_t ch) { return ch ? ch->ws_bytes : 0; }

namespace {

// see gpbt_chain::last_stream
int chain_enter_stream(gpbt_chain* ch, cudaStream_t st) {
  if (ch->last_stream_valid && ch->last_stream != st) CU(cudaStreamSynchronize(ch->last_stream));
  ch->last_stream = st;
  ch->last_stream_valid = true;
  return 0;
}

// rows of PC-space workspace (z_mean, z_var, extra, skip): grows to the largest N seen
int ensure_rows(gpbt_chain* ch, int64_t N) {
  if (N <= ch->cap_rows) return 0;
  const int64_t cap = std::max<int64_t>(N, 2 * ch->cap_rows);
  if (ch->z_mean) { cudaFree(ch->z_mean); cudaFree(ch->z_var); cudaFree(ch->extra); cudaFree(ch->skip); }
  ch->z_mean = ch->z_var = ch->extra = nullptr; ch->skip = nullptr;
  ch->ws_bytes -= ch->cap_rows * (2 * ch->Q * 8 + 9);
  ch->cap_rows = 0;   // a failed allocation below leaves an empty, consistent workspace
  CU(cudaMalloc(&ch->z_mean, (size_t)cap * ch->Q * sizeof(double)));
  CU(cudaMalloc(&ch->z_var, (size_t)cap * ch->Q * sizeof(double)));
  CU(cudaMalloc(&ch->extra, (size_t)cap * sizeof(double)));
  CU(cudaMalloc(&ch->skip, (size_t)cap));
  ch->ws_bytes += cap * (2 * ch->Q * 8 + 9);
  ch->cap_rows = cap;
  g_ws_generation++;
  return 0;
}

// rows of observable-space workspace for the dense path (mean [rows, M], cov [rows, M, M])
int64_t dense_chunk_rows(const gpbt_chain* ch, int64_t N) {
  const size_t per_row = (size_t)ch->M * ch->M * 8 + (size_t)ch->M * 8;
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  size_t budget = std::min<size_t>((size_t)24 << 30, free_b / 2 + (size_t)ch->cap_dense * per_row);
  int64_t rows = (int64_t)std::max<size_t>(budget / per_row, 1);
  rows = std::min<int64_t>(rows, 32768);
  return std::min<int64_t>(rows, N);
}

int ensure_dense(gpbt_chain* ch, int64_t rows) {
  if (rows <= ch->cap_dense) return 0;
  if (ch->mean) { cudaFree(ch->mean); cudaFree(ch->cov); }
  ch->mean = ch->cov = nullptr;
  ch->ws_bytes -= ch->cap_dense * ((int64_t)ch->M * ch->M * 8 + (int64_t)ch->M * 8);
  ch->cap_dense = 0;
  CU(cudaMalloc(&ch->mean, (size_t)rows * ch->M * sizeof(double)));
  CU(cudaMalloc(&ch->cov, (size_t)rows * ch->M * ch->M * sizeof(double)));
  ch->ws_bytes += rows * ((int64_t)ch->M * ch->M * 8 + (int64_t)ch->M * 8);
  ch->cap_dense = rows;
  g_ws_generation++;
  return 0;
}

int ensure_io(gpbt_chain* ch, int64_t N) {
  if (N <= ch->cap_x) return 0;
  const int64_t cap = std::max<int64_t>(N, 2 * ch->cap_x);
  if (ch->x_dev) { cudaFree(ch->x_dev); cudaFree(ch->lp_dev); }
  ch->x_dev = ch->lp_dev = nullptr;
  ch->ws_bytes -= ch->cap_x * ((int64_t)ch->p * 8 + 8);
  ch->cap_x = 0;
  CU(cudaMalloc(&ch->x_dev, (size_t)cap * ch->p * sizeof(double)));
  CU(cudaMalloc(&ch->lp_dev, (size_t)cap * sizeof(double)));
  ch->ws_bytes += cap * ((int64_t)ch->p * 8 + 8);
  ch->cap_x = cap;
  return 0;
}

// fused dense path: per-walker factor panels and work vectors, mean [rows, M]
int64_t fused_bytes_per_row(const gpbt_chain* ch) {
  return (ch->Lstride + 2 * kCfNB * kCfNB + ch->Mg + 2 + ch->M) * (int64_t)sizeof(double) + (int64_t)sizeof(int);
}

int64_t fused_chunk_rows(const gpbt_chain* ch, int64_t N) {
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const size_t per_row = (size_t)fused_bytes_per_row(ch);
  const size_t budget = std::min<size_t>((size_t)24 << 30, free_b / 2 + (size_t)ch->cap_fused * per_row);
  int64_t rows = (int64_t)std::max<size_t>(budget / per_row, 1);
  rows = std::min<int64_t>(rows, 65536);
  return std::min<int64_t>(rows, N);
}

int ensure_fused(gpbt_chain* ch, int64_t rows) {
  if (rows <= ch->cap_fused) return 0;
  void* old[] = {ch->cf_L, ch->cf_dinv, ch->cf_draw, ch->cf_tvec, ch->cf_logdet, ch->cf_tsq, ch->cf_bad, ch->cf_mean};
  for (void* p : old)
    if (p) cudaFree(p);
  ch->cf_L = ch->cf_dinv = ch->cf_draw = ch->cf_tvec = ch->cf_logdet = ch->cf_tsq = ch->cf_mean = nullptr;
  ch->cf_bad = nullptr;
  ch->ws_bytes -= ch->cap_fused * fused_bytes_per_row(ch);
  ch->cap_fused = 0;
  CU(cudaMalloc(&ch->cf_L, (size_t)rows * std::max<int64_t>(ch->Lstride, 1) * sizeof(double)));
  CU(cudaMalloc(&ch->cf_dinv, (size_t)rows * kCfNB * kCfNB * sizeof(double)));
  CU(cudaMalloc(&ch->cf_draw, (size_t)rows * kCfNB * kCfNB * sizeof(double)));
  CU(cudaMalloc(&ch->cf_tvec, (size_t)rows * ch->Mg * sizeof(double)));
  CU(cudaMalloc(&ch->cf_logdet, (size_t)rows * sizeof(double)));
  CU(cudaMalloc(&ch->cf_tsq, (size_t)rows * sizeof(double)));
  CU(cudaMalloc(&ch->cf_bad, (size_t)rows * sizeof(int)));
  CU(cudaMalloc(&ch->cf_mean, (size_t)rows * ch->M * sizeof(double)));
  ch->ws_bytes += rows * fused_bytes_per_row(ch);
  ch->cap_fused = rows;
  g_ws_generation++;
  return 0;
}

// (b) + (c) fused: lp[w] from z_var (kernel (a)) and mean, for walkers [0, N) of the chunk.  One launch
// per 32-column panel; the walkers go through in sub-batches whose factors stay L2 resident
// (option "chol_batch", default: what fits in ~60 % of the L2).
// Launches of the fused Cholesky over walkers [0, N): sub-batches round-robin over up to kCfMaxStreams
// streams (the caller's + aux[1..]): while one sub-batch is in its factor kernel or in the tail of a panel
// launch, the DMMA work of another fills the machine.  Options "chol_streams" / "chol_batch" override the
// defaults.  prm carries everything but the debugging fields.
// produce(w0, nw): enqueue on the caller's stream whatever fills z_var / mean of walkers [w0, w0 + nw)
using FusedProduce = std::function<int(int64_t, int64_t)>;

int launch_chol_fused(CholFusedParams prm, int64_t N, FusedStreams* fs, cudaStream_t st,
                      const FusedProduce* produce = nullptr) {
  prm.dbg = nullptr;
  prm.flags = g_opt.cf_debug.load() >> 4;      // (cf_debug = 16 * flags + record-stamps bit)
  prm.early_after = device_info().sm_count * kCfCtasPerSm;   // the CTAs that can be resident at once
  if (g_opt.cf_debug.load() & 1) {
    if (!g_cf_dbg) CU(cudaMallocManaged(&g_cf_dbg, kCfDbgBytes));
    prm.dbg = g_cf_dbg;
  }
  const int Mg = prm.Mg;
  const size_t smem = chol_fused_smem_bytes(Mg);
  if (smem > (size_t)max_optin_smem()) return fail(GPBT_ESHAPE, "fused Cholesky: M = %d observables do not fit", prm.M);
  const bool dense = prm.cov_src != nullptr;
  if (int r = dense ? ensure_dynamic_smem<chol_fused_panel_kernel<true>>(smem)
                    : ensure_dynamic_smem<chol_fused_panel_kernel<false>>(smem))
    return r;
  if (int r = ensure_dynamic_smem<chol_fused_factor_kernel>(kCfFactorSmem)) return r;
  void (*panel)(CholFusedParams, int, int64_t) = dense ? chol_fused_panel_kernel<true> : chol_fused_panel_kernel<false>;
  // Pipelined form (option "chol_pipe" = number of sub-batches; off by default): the producer -- kernel (a) and
  // the mean -- runs per sub-batch on the caller's stream, the Cholesky launches of sub-batch b follow on an
  // aux stream as soon as ITS inputs are there, so kernel (a) of sub-batch b + 1 shares the machine with the
  // panel launches of sub-batch b instead of running in front of all of them.  Measured at config 2, 4096
  // walkers, best of 20 on three boxes: 3.44 ms against 3.38 ms for kernel (a) once + two sub-batches side by
  // side -- two kernel-(a) CTAs hold all 64 K registers of an SM, so the kernels never share an SM, they take
  // turns, and the last sub-batch ends alone.  (One box ran the side-by-side form at 3.65 ms in every
  // repetition -- see "chol_lag" -- and the pipelined form at 3.49.)
  int64_t pipe = produce ? g_opt.chol_pipe.load() : 1;
  if (pipe <= 0) pipe = 1;
  const bool pipelined = pipe > 1;
  if (produce && !pipelined)
    if (int r = (*produce)(0, N)) return r;
  int n_streams = (int)g_opt.chol_streams.load();
  if (n_streams <= 0) n_streams = pipelined ? kCfMaxStreams - 1 : (N >= 2048 ? 2 : 1);
  n_streams = std::min(n_streams, pipelined ? kCfMaxStreams - 1 : kCfMaxStreams);
  int64_t batch = g_opt.chol_batch.load();
  if (pipelined) batch = round_up((N + pipe - 1) / pipe, 64);   // (whole walker tiles of kernel (a))
  else if (batch <= 0) batch = std::max<int64_t>(512, (N + n_streams - 1) / n_streams);
  batch = std::min<int64_t>(batch, 32768);          // grid.y carries the walker
  const int64_t n_batches = (N + batch - 1) / batch;
  n_streams = (int)std::min<int64_t>(n_streams, n_batches);
  const int first_aux = pipelined ? 0 : 1;          // pipelined: every sub-batch on an aux stream
  if (n_streams > first_aux) {
    int prio_least = 0, prio_greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    for (int i = 1; i < n_streams + 1 - first_aux; i++)
      if (!fs->aux[i]) {
        CU(cudaStreamCreateWithPriority(&fs->aux[i], cudaStreamNonBlocking, g_opt.chol_prio.load() ? prio_greatest : prio_least));
        CU(cudaEventCreateWithFlags(&fs->done[i], cudaEventDisableTiming));
      }
    if (!fs->fork) CU(cudaEventCreateWithFlags(&fs->fork, cudaEventDisableTiming));
    if (!pipelined) {
      CU(cudaEventRecord(fs->fork, st));
      for (int i = 1; i < n_streams; i++) CU(cudaStreamWaitEvent(fs->aux[i], fs->fork, 0));
    }
  }
  // Sub-batches that start at the same instant can stay in lockstep: their factor kernels (a few warps per
  // SM, latency bound, ~20 us each) then coincide instead of hiding behind the other one's panel launch --
  // 0.2 ms per 4096 walkers on one of four boxes.  Option "chol_lag" (off by default) starts a sub-batch one
  // step behind its predecessor, after that one's first factor kernel; it costs 0.06 ms where there is no
  // lockstep to break (3.44 against 3.38 ms, best of 20).
  const bool lag = !pipelined && n_streams > 1 && g_opt.chol_lag.load();
  for (int64_t b = 0; b < n_batches; b++) {
    const int64_t w0 = b * batch, nw = std::min(batch, N - w0);
    cudaStream_t sb;
    if (pipelined) {
      sb = fs->aux[1 + b % n_streams];
      if (int r = (*produce)(w0, nw)) return r;
      CU(cudaEventRecord(fs->fork, st));          // (a wait takes the event as recorded now: one event serves all)
      CU(cudaStreamWaitEvent(sb, fs->fork, 0));
    } else {
      sb = (b % n_streams == 0) ? st : fs->aux[b % n_streams];
    }
    if (lag && b > 0 && b < n_streams) CU(cudaStreamWaitEvent(sb, fs->fork, 0));
    for (int J = -kCfNB; J + kCfNB < Mg; J += kCfNB) {
      const int tiles = (J >= 0 && Mg > J + 2 * kCfNB) ? (Mg - J - 2 * kCfNB + kCfRows - 1) / kCfRows : 0;
      if (J == 0 && lag && b + 1 < n_streams) CU(cudaEventRecord(fs->fork, sb));   // (behind factor(0))
      if (prm.flags & 4) {   // plain stream order, no programmatic launch
        panel<<<dim3((unsigned)(1 + tiles), (unsigned)nw), kCfThreads, smem, sb>>>(prm, J, w0);
        LAUNCH_CHECK();
        chol_fused_factor_kernel<<<dim3((unsigned)((nw + kCfFactorWarps - 1) / kCfFactorWarps)), kCfFactorWarps * 32,
                                   kCfFactorSmem, sb>>>(prm, J + kCfNB, w0, nw);
        LAUNCH_CHECK();
        continue;
      }
      CU(launch_pdl(panel, dim3((unsigned)(1 + tiles), (unsigned)nw), kCfThreads, smem, sb, prm, J, w0));
      LAUNCH_CHECK();
      CU(launch_pdl(chol_fused_factor_kernel, dim3((unsigned)((nw + kCfFactorWarps - 1) / kCfFactorWarps)),
                    kCfFactorWarps * 32, kCfFactorSmem, sb, prm, J + kCfNB, w0, nw));
      LAUNCH_CHECK();
    }
  }
  for (int i = 1; i < n_streams + 1 - first_aux; i++) {
    CU(cudaEventRecord(fs->done[i], fs->aux[i]));
    CU(cudaStreamWaitEvent(st, fs->done[i], 0));
  }
  return 0;
}

int run_chol_fused(gpbt_chain* ch, double* lp, int* n_notpd, double notpd_value, int64_t N, cudaStream_t st,
                   const FusedProduce* produce = nullptr) {
  CholFusedParams prm;
  prm.Fp = ch->cf_Fp; prm.Fd = ch->cf_Fd; prm.UT = ch->cf_UT; prm.cov_src = nullptr; prm.cov_add = nullptr; prm.dense_vec = 0;
  prm.z_var = ch->z_var; prm.mean = ch->cf_mean;
  prm.y_exp = ch->y_exp; prm.skip = ch->skip; prm.L = ch->cf_L; prm.dinv = ch->cf_dinv; prm.draw = ch->cf_draw; prm.tvec = ch->cf_tvec;
  prm.logdet = ch->cf_logdet; prm.tsq = ch->cf_tsq; prm.bad = ch->cf_bad; prm.lp = lp; prm.n_notpd = n_notpd;
  prm.notpd_value = notpd_value; prm.add_const = kSysConst; prm.N = N; prm.Lstride = ch->Lstride; prm.ldz = ch->Q;
  prm.M = ch->M; prm.Mg = ch->Mg; prm.Q = ch->Q; prm.Qp = ch->Qp;
  return launch_chol_fused(prm, N, &ch->cf_fs, st, produce);
}

// Stand-alone batched mvn_loglike on materialised covariances through the same kernels (dense source).
// Work buffers and streams: one grow-only set per (device, stream).
struct FusedDenseCache {
  double* buf = nullptr;
  size_t bytes = 0;
  FusedStreams fs;
};
std::map<std::pair<int, cudaStream_t>, FusedDenseCache> g_fused_dense;
std::mutex g_fused_dense_mutex;

int run_chol_fused_dense(const CholParams& cp, cudaStream_t st) {
  const int M = cp.m, Mg = (int)round_up(M, 16);
  const int64_t N = cp.N, Lstride = std::max<int64_t>(cf_factor_doubles(M), 1);
  const size_t per_row = (size_t)(Lstride + 2 * kCfNB * kCfNB + Mg + 2) * sizeof(double) + sizeof(int);
  FusedDenseCache* c;
  {
    std::lock_guard<std::mutex> lock(g_fused_dense_mutex);
    c = &g_fused_dense[std::make_pair(current_device(), st)];
    if ((size_t)N * per_row > c->bytes) {
      if (c->buf) cudaFree(c->buf);
      c->buf = nullptr;
      c->bytes = 0;
      CU(cudaMalloc(&c->buf, (size_t)N * per_row));
      c->bytes = (size_t)N * per_row;
      g_ws_generation++;
    }
  }
  CholFusedParams prm;
  prm.Fp = prm.Fd = prm.UT = prm.z_var = nullptr;
  prm.dense_vec = (M % 2 == 0) && ((reinterpret_cast<uintptr_t>(cp.cov) & 15) == 0) &&
                  (cp.cov_add == nullptr || (reinterpret_cast<uintptr_t>(cp.cov_add) & 15) == 0);
  prm.cov_src = cp.cov; prm.cov_add = cp.cov_add; prm.mean = cp.mean; prm.y_exp = cp.y_exp; prm.skip = cp.skip;
  prm.L = c->buf;
  prm.dinv = prm.L + (size_t)N * Lstride;
  prm.draw = prm.dinv + (size_t)N * kCfNB * kCfNB;
  prm.tvec = prm.draw + (size_t)N * kCfNB * kCfNB;
  prm.logdet = prm.tvec + (size_t)N * Mg;
  prm.tsq = prm.logdet + N;
  prm.bad = reinterpret_cast<int*>(prm.tsq + N);
  prm.lp = cp.lp; prm.n_notpd = cp.n_notpd; prm.notpd_value = cp.notpd_value; prm.add_const = cp.add_const;
  prm.N = N; prm.Lstride = Lstride; prm.ldz = 0; prm.M = M; prm.Mg = Mg; prm.Q = 0; prm.Qp = 0;
  return launch_chol_fused(prm, N, &c->fs, st);
}

void release_fused_dense(int device, cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_fused_dense_mutex);
  auto it = g_fused_dense.find(std::make_pair(device, st));
  if (it == g_fused_dense.end()) return;
  if (it->second.buf) cudaFree(it->second.buf);
  it->second.fs.destroy();
  g_fused_dense.erase(it);
}

// Chain._predict into (mean, cov) for rows [0, N) of X; cov may be null
// (row0 / tile_rows: X, mean and cov are those of rows [row0, row0 + N) of a batch of tile_rows rows whose
// PC-space work rows the chain holds from row 0 -- the pipelined dense path)
int chain_predict_rows(gpbt_chain* ch, const double* X, double extra_scale, double* mean, double* cov,
                       int64_t N, cudaStream_t st, bool with_exp = false, int64_t row0 = 0, int64_t tile_rows = 0) {
  if (int r = ensure_rows(ch, row0 + N)) return r;
  const double* extra = nullptr;
  if (extra_scale != 0.0) {
    extra_std_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(X, ch->p, N, extra_scale, ch->extra + row0);
    LAUNCH_CHECK();
    extra = ch->extra + row0;
  }
  for (size_t e = 0; e < ch->emus.size(); e++) {
    gpbt_emulator_t emu = ch->emus[e];
    double* zm = ch->z_mean + row0 * ch->Q + ch->q_off[e];
    double* zv = ch->z_var + row0 * ch->Q + ch->q_off[e];
    if (int r = run_pc_predict(emu, X, extra, zm, zv, ch->Q, N, st, tile_rows)) return r;
    if (int r = run_backtransform(emu, zm, zv, ch->Q, mean, ch->M,
                                  cov, ch->M, ch->m_off[e], N, st, nullptr, with_exp ? ch->base_like[e] : nullptr))
      return r;
  }
  return 0;
}

}  // namespace

extern "C" int gpbt_chain_predict(gpbt_chain_t ch, const double* X, double extra_std_scale, double* mean,
                                  double* cov, int64_t N, void* stream) {
  if (!ch || !X || !mean || N < 0) return fail(GPBT_EINVAL, "gpbt_chain_predict: bad argument");
  if (int r = chain_enter_stream(ch, (cudaStream_t)stream)) return r;
  const int64_t step = 32768;
  for (int64_t s = 0; s < N; s += step) {
    const int64_t nn = std::min(step, N - s);
    if (int r = chain_predict_rows(ch, X + s * ch->p, extra_std_scale, mean + s * ch->M,
                                   cov ? cov + (size_t)s * ch->M * ch->M : nullptr, nn, (cudaStream_t)stream))
      return r;
  }
  return 0;
}

namespace {
int log_posterior_impl(gpbt_chain_t ch, const double* X, double oob_value, double* lp, int* n_notpd, int64_t N,
                       int path, void* stream, double* const* peers, int n_peers, int64_t peer_off,
                       bool zero_counter = true);

int scatter_result(const double* lp, double* const* peers, int n_peers, int64_t peer_off, int64_t N, cudaStream_t st) {
  if (n_peers <= 0 || N <= 0) return 0;
  PeerList pl;
  pl.n = n_peers;
  for (int r = 0; r < n_peers; r++) pl.p[r] = peers[r];
  scatter_to_peers_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(lp, pl, peer_off, N);
  LAUNCH_CHECK();
  return 0;
}
}  // namespace

extern "C" int gpbt_log_posterior(gpbt_chain_t ch, const double* X, double oob_value, double* lp, int* n_notpd,
                                  int64_t N, int path, void* stream) {
  return log_posterior_impl(ch, X, oob_value, lp, n_notpd, N, path, stream, nullptr, 0, 0);
}

extern "C" int gpbt_log_posterior_scatter(gpbt_chain_t ch, const double* X, double oob_value, double* lp,
                                          double* const* peers_host, int n_peers, int64_t peer_off, int* n_notpd,
                                          int64_t N, int path, void* stream) {
  if (n_peers < 0 || n_peers > kMaxPeers || (n_peers > 0 && !peers_host) || peer_off < 0)
    return fail(GPBT_EINVAL, "gpbt_log_posterior_scatter: bad peer list");
  return log_posterior_impl(ch, X, oob_value, lp, n_notpd, N, path, stream, peers_host, n_peers, peer_off);
}

namespace {
int log_posterior_impl(gpbt_chain_t ch, const double* X, double oob_value, double* lp, int* n_notpd, int64_t N,
                       int path, void* stream, double* const* peers, int n_peers, int64_t peer_off, bool zero_counter) {
  if (!ch || !X || !lp || N < 0) return fail(GPBT_EINVAL, "gpbt_log_posterior: bad argument");
  if (ch->device != current_device())
    return fail(GPBT_EINVAL, "chain lives on device %d, current device is %d", ch->device, current_device());
  if (N == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (int r = chain_enter_stream(ch, st)) return r;
  if (path == GPBT_PATH_AUTO)
    path = ch->has_lowrank ? GPBT_PATH_LOWRANK : (ch->has_diag ? GPBT_PATH_DIAG : GPBT_PATH_DENSE);
  if (path == GPBT_PATH_DIAG && !ch->has_diag)
    return fail(GPBT_ENOTAPPLICABLE, "diagonal path requested but the covariance of this chain is not diagonal");
  if (path == GPBT_PATH_LOWRANK && !ch->has_lowrank)
    return fail(GPBT_ENOTAPPLICABLE, "low-rank path requested but the chain has no low-rank factors");
  if (n_notpd && zero_counter) CU(cudaMemsetAsync(n_notpd, 0, sizeof(int), st));

  if (path == GPBT_PATH_LOWRANK) {
    if (int r = ensure_rows(ch, N)) return r;
    for (size_t e = 0; e < ch->emus.size(); e++)
      if (int r = run_pc_predict(ch->emus[e], X, nullptr, ch->z_mean + ch->q_off[e], ch->z_var + ch->q_off[e],
                                 ch->Q, N, st))
        return r;
    // One launch over all Q PCs, or -- when R is block diagonal over the emulators (the experimental
    // covariance does not couple them) -- one launch per emulator block: log L is then a sum of
    // per-block terms, every block fits the register kernel (q_e <= 32) and the cubic cost is in q_e.
    const int n_blocks = ch->lr_separable ? (int)ch->emus.size() : 1;
    for (int blk = 0; blk < n_blocks; blk++) {
      const int q0 = ch->lr_separable ? ch->q_off[blk] : 0;
      const int qb = ch->lr_separable ? ch->emus[blk]->q : ch->Q;
      LowrankParams prm;
      prm.X = X; prm.lo = ch->lo; prm.hi = ch->hi; prm.z_mean = ch->z_mean; prm.z_var = ch->z_var;
      prm.ldz = ch->Q; prm.z_off = q0; prm.accumulate = blk > 0;
      prm.R = ch->lr_separable ? ch->R_blocks[blk] : ch->R; prm.c0 = ch->c0 + q0; prm.lp = lp; prm.n_notpd = n_notpd;
      // the walker-independent constants enter once, with the first block
      prm.s_perp = blk == 0 ? ch->s_perp : 0.0;
      prm.logdetF_half = blk == 0 ? ch->logdetF_half : 0.0;
      prm.oob_value = oob_value; prm.sys_const = blk == 0 ? kSysConst : 0.0;
      prm.N = N; prm.p = ch->p; prm.Q = qb;
      // peers receive the running value of every block; the last launch leaves the final one
      prm.n_peers = n_peers; prm.peer_off = peer_off;
      for (int r = 0; r < n_peers; r++) prm.peers[r] = peers[r];
      const unsigned grid = (unsigned)((N + kLrWarps - 1) / kLrWarps);
      if (qb <= 32 && !g_opt.lowrank_generic.load()) {
        switch ((qb + 3) / 4) {
#define GPBT_Q(QP) \
  case QP / 4: CU(launch_pdl(lowrank_loglike_reg_kernel<QP>, dim3(grid), kLrWarps * 32, 0, st, prm)); break;
          GPBT_Q(4) GPBT_Q(8) GPBT_Q(12) GPBT_Q(16) GPBT_Q(20) GPBT_Q(24) GPBT_Q(28) GPBT_Q(32)
#undef GPBT_Q
        }
        LAUNCH_CHECK();
        continue;
      }
      const size_t smem = lowrank_smem_bytes(qb);
      if (smem > (size_t)max_optin_smem()) return fail(GPBT_ESHAPE, "low-rank path: Q = %d too large", qb);
      if (int r = ensure_dynamic_smem<lowrank_loglike_kernel>(smem)) return r;
      CU(launch_pdl(lowrank_loglike_kernel, dim3(grid), kLrWarps * 32, smem, st, prm));
      LAUNCH_CHECK();
    }
    return 0;
  }

  if (path == GPBT_PATH_DIAG) {
    // (a) -> mean + diag(cov) -> element-wise likelihood; workspace: mean/var [N, M]
    const int64_t chunk = std::min<int64_t>(N, 1 << 18);
    if (int r = ensure_rows(ch, chunk)) return r;
    if (chunk > ch->cap_diag) {
      if (ch->dmean) { cudaFree(ch->dmean); cudaFree(ch->dvar); }
      ch->dmean = ch->dvar = nullptr;
      ch->ws_bytes -= ch->cap_diag * (int64_t)ch->M * 16;
      ch->cap_diag = 0;
      CU(cudaMalloc(&ch->dmean, (size_t)chunk * ch->M * sizeof(double)));
      CU(cudaMalloc(&ch->dvar, (size_t)chunk * ch->M * sizeof(double)));
      ch->ws_bytes += chunk * (int64_t)ch->M * 16;
      ch->cap_diag = chunk;
      g_ws_generation++;
    }
    for (int64_t s = 0; s < N; s += chunk) {
      const int64_t nn = std::min(chunk, N - s);
      const double* Xs = X + s * ch->p;
      for (size_t e = 0; e < ch->emus.size(); e++) {
        gpbt_emulator_t emu = ch->emus[e];
        if (int r = run_pc_predict(emu, Xs, nullptr, ch->z_mean + ch->q_off[e], ch->z_var + ch->q_off[e], ch->Q, nn, st))
          return r;
        if (int r = run_backtransform(emu, ch->z_mean + ch->q_off[e], ch->z_var + ch->q_off[e], ch->Q, ch->dmean,
                                      ch->M, nullptr, 0, ch->m_off[e], nn, st, ch->dvar))
          return r;
      }
      diag_loglike_kernel<<<(unsigned)((nn * 32 + 255) / 256), 256, 0, st>>>(
          Xs, ch->lo, ch->hi, ch->p, ch->dmean, ch->dvar, ch->y_exp, ch->cov_exp, ch->M, nn, oob_value, kSysConst,
          lp + s, n_notpd);
      LAUNCH_CHECK();
    }
    return scatter_result(lp, peers, n_peers, peer_off, N, st);
  }

  // dense path, PCA-mode chains with more than 80 observables: (a) -> fused (b)+(c), the covariance is
  // generated tile by tile inside the panel-synchronous Cholesky and never stored (chol_fused.cuh)
  const int chol_opt = g_opt.chol.load();
  if (ch->has_fused && (chol_opt == 'f' || (chol_opt == 0 && ch->M > 80 && N >= 64))) {
    const int64_t chunk = fused_chunk_rows(ch, N);
    if (int r = ensure_fused(ch, chunk)) return r;
    if (int r = ensure_rows(ch, chunk)) return r;
    for (int64_t s = 0; s < N; s += chunk) {
      const int64_t nn = std::min(chunk, N - s);
      const double* Xs = X + s * ch->p;
      bounds_mask_kernel<<<(unsigned)((nn + 127) / 128), 128, 0, st>>>(Xs, ch->lo, ch->hi, ch->p, nn, oob_value,
                                                                       ch->skip, lp + s);
      LAUNCH_CHECK();
      const FusedProduce produce = [&](int64_t w0, int64_t nw) {
        return chain_predict_rows(ch, Xs + w0 * ch->p, 0.0, ch->cf_mean + w0 * ch->M, nullptr, nw, st, false, w0, nn);
      };
      if (int r = run_chol_fused(ch, lp + s, n_notpd, oob_value, nn, st, &produce)) return r;
    }
    return scatter_result(lp, peers, n_peers, peer_off, N, st);
  }
  // dense path: (a) -> (b) with the covariance materialised in HBM -> (c), in row chunks
  const int64_t chunk = dense_chunk_rows(ch, N);
  if (int r = ensure_dense(ch, chunk)) return r;
  if (int r = ensure_rows(ch, chunk)) return r;
  for (int64_t s = 0; s < N; s += chunk) {
    const int64_t nn = std::min(chunk, N - s);
    const double* Xs = X + s * ch->p;
    bounds_mask_kernel<<<(unsigned)((nn + 127) / 128), 128, 0, st>>>(Xs, ch->lo, ch->hi, ch->p, nn, oob_value,
                                                                     ch->skip, lp + s);
    LAUNCH_CHECK();
    const bool pre = ch->exp_blockdiag;   // kernel (b) already adds the experimental covariance
    if (int r = chain_predict_rows(ch, Xs, 0.0, ch->mean, ch->cov, nn, st, pre)) return r;
    if (int r = run_chol(ch->mean, ch->y_exp, ch->cov, pre ? nullptr : ch->cov_exp, lp + s, n_notpd, ch->skip,
                         oob_value, kSysConst, nn, ch->M, st))
      return r;
  }
  return scatter_result(lp, peers, n_peers, peer_off, N, st);
}
}  // namespace

extern "C" int gpbt_log_posterior_host(gpbt_chain_t ch, const double* X_host, double oob_value, double* lp_host,
                                       int* n_notpd_host, int64_t N, int path) {
  if (!ch || !X_host || !lp_host || N < 0) return fail(GPBT_EINVAL, "gpbt_log_posterior_host: bad argument");
  if (N == 0) return 0;
  CU(cudaSetDevice(ch->device));
  cudaStream_t st = ch->stream;
  if (N <= kZeroCopyRows && !g_opt.no_zerocopy.load()) {
    if (!ch->zc_x_host) {
      CU(cudaHostAlloc(&ch->zc_x_host, (size_t)kZeroCopyRows * ch->p * sizeof(double), cudaHostAllocMapped));
      CU(cudaHostAlloc(&ch->zc_lp_host, (size_t)(kZeroCopyRows + 1) * sizeof(double), cudaHostAllocMapped));
      CU(cudaHostGetDevicePointer(&ch->zc_x_dev, ch->zc_x_host, 0));
      CU(cudaHostGetDevicePointer(&ch->zc_lp_dev, ch->zc_lp_host, 0));
    }
    memcpy(ch->zc_x_host, X_host, (size_t)N * ch->p * sizeof(double));
    int* cnt_host = reinterpret_cast<int*>(ch->zc_lp_host + kZeroCopyRows);
    int* cnt_dev = reinterpret_cast<int*>(ch->zc_lp_dev + kZeroCopyRows);
    *cnt_host = 0;   // mapped memory: the host clears the counter itself, no memset node on the stream
    if (int r = log_posterior_impl(ch, ch->zc_x_dev, oob_value, ch->zc_lp_dev, cnt_dev, N, path, st, nullptr, 0, 0,
                                   /*zero_counter=*/false))
      return r;
    CU(cudaStreamSynchronize(st));
    memcpy(lp_host, ch->zc_lp_host, (size_t)N * sizeof(double));
    if (n_notpd_host) *n_notpd_host = *cnt_host;
    return 0;
  }
  if (int r = ensure_io(ch, N)) return r;
  CU(cudaMemcpyAsync(ch->x_dev, X_host, (size_t)N * ch->p * sizeof(double), cudaMemcpyHostToDevice, st));
  if (int r = gpbt_log_posterior(ch, ch->x_dev, oob_value, ch->lp_dev, ch->notpd_dev, N, path, st)) return r;
  CU(cudaMemcpyAsync(lp_host, ch->lp_dev, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (n_notpd_host) CU(cudaMemcpyAsync(n_notpd_host, ch->notpd_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

// ---- device-resident ensemble sampler ------------------------------------------------------------
struct gpbt_ensemble {
  gpbt_chain_t ch = nullptr;
  int nw = 0, p = 0, n_half = 0, randomize = 1;
  double a = 2.0;
  uint64_t seed = 0;
  EnsembleCtl* ctl = nullptr;            // device
  double *x = nullptr, *lp = nullptr, *q = nullptr, *factor = nullptr, *u_acc = nullptr, *lp_new = nullptr;
  int *perm = nullptr, *notpd_call = nullptr;
  unsigned int* rec_done = nullptr;      // CTA arrival counter of ensemble_record_kernel
  unsigned long long* keys = nullptr;    // [nw] split keys of the current step
  long long *accepted = nullptr, *notpd_total = nullptr;
  double *hist_x = nullptr, *hist_lp = nullptr;
  int64_t hist_cap = 0, steps = 0;
  int64_t prepared_until = 0;            // steps the history / control block have room for (piecewise stepping)
  double* ru = nullptr;                  // device copies of the host random streams of the current run
  int *rp = nullptr, *perm_in = nullptr;
  cudaGraphExec_t graph = nullptr;
  int64_t graph_gen = -1, graph_launches = 0;   // kernels per replay (for gpbt_launch_count)
  bool has_state = false;
};

namespace {

EnsembleBuffers ensemble_buffers(const gpbt_ensemble* en) {
  EnsembleBuffers b;
  b.ctl = en->ctl; b.seed = en->seed; b.a = en->a;
  b.nw = en->nw; b.p = en->p; b.n_half = en->n_half; b.randomize = en->randomize;
  b.perm = en->perm; b.x = en->x; b.lp = en->lp; b.q = en->q; b.factor = en->factor; b.u_acc = en->u_acc;
  b.lp_new = en->lp_new; b.notpd_call = en->notpd_call; b.notpd_total = en->notpd_total; b.accepted = en->accepted;
  return b;
}

// log-posterior of the proposals of one half step; the counter of non-PD covariances is folded and
// reset by the accept that follows, not by a memset in front
int ensemble_log_posterior(gpbt_ensemble* en, int ns, cudaStream_t st) {
  return log_posterior_impl(en->ch, en->q, -INFINITY, en->lp_new, en->notpd_call, ns, GPBT_PATH_AUTO, st, nullptr, 0, 0,
                            /*zero_counter=*/false);
}

int ensemble_enqueue_step(gpbt_ensemble* en, cudaStream_t st) {
  const int nw = en->nw, p = en->p;
  const EnsembleBuffers b = ensemble_buffers(en);
  const size_t key_bytes = (size_t)nw * sizeof(unsigned long long);
  if (nw <= kEnsembleFusedMaxWalkers && !g_opt.ensemble_split_kernels.load()) {
    // small ensemble: launch latency is the cost, three single-CTA kernels around the two calls
    CU(launch_pdl(ensemble_begin_kernel, dim3(1), 1024, key_bytes, st, b));
    LAUNCH_CHECK();
    if (int r = ensemble_log_posterior(en, en->n_half, st)) return r;
    CU(launch_pdl(ensemble_mid_kernel, dim3(1), 1024, 0, st, b));
    LAUNCH_CHECK();
    if (nw - en->n_half > 0)
      if (int r = ensemble_log_posterior(en, nw - en->n_half, st)) return r;
    CU(launch_pdl(ensemble_end_kernel, dim3(1), 1024, 0, st, b, en->ctl));
    LAUNCH_CHECK();
    return 0;
  }
  ensemble_keys_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(en->ctl, en->seed, nw, en->randomize, en->keys,
                                                                    en->perm);
  LAUNCH_CHECK();
  if (nw <= kEnsembleRankMaxWalkers) {
    ensemble_rank_kernel<<<(unsigned)((nw + kRankThreads - 1) / kRankThreads), kRankThreads, 0, st>>>(
        en->ctl, nw, en->randomize, en->keys, en->perm);
    LAUNCH_CHECK();
  }
  for (int half = 0; half < 2; half++) {
    const int ns = half == 0 ? en->n_half : nw - en->n_half;
    if (ns == 0) continue;
    const unsigned grid = (unsigned)((ns + 127) / 128);
    ensemble_propose_kernel<<<grid, 128, 0, st>>>(b, half);
    LAUNCH_CHECK();
    if (int r = ensemble_log_posterior(en, ns, st)) return r;
    ensemble_accept_kernel<<<grid, 128, 0, st>>>(b, half);
    LAUNCH_CHECK();
  }
  const unsigned rec_grid = (unsigned)std::min<int64_t>(((int64_t)nw * p + 255) / 256, device_info().sm_count);
  ensemble_record_kernel<<<rec_grid, 256, 0, st>>>(en->ctl, nw, p, en->x, en->lp, en->rec_done);
  LAUNCH_CHECK();
  return 0;
}

// History capacity in steps: grown geometrically (or to an explicit reservation) because device
// allocation is the one slow call on this path (0.4 - 5 ms observed per cudaMalloc / cudaFree pair).
int ensemble_grow_history(gpbt_ensemble* en, int64_t need, cudaStream_t st, bool exact = false) {
  if (need <= en->hist_cap) return 0;
  const int64_t cap = exact ? need : std::max<int64_t>(need, 2 * en->hist_cap);
  double *hx = nullptr, *hl = nullptr;
  CU(cudaMalloc(&hx, (size_t)cap * en->nw * en->p * sizeof(double)));
  CU(cudaMalloc(&hl, (size_t)cap * en->nw * sizeof(double)));
  if (en->steps > 0) {
    CU(cudaMemcpyAsync(hx, en->hist_x, (size_t)en->steps * en->nw * en->p * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(hl, en->hist_lp, (size_t)en->steps * en->nw * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CU(cudaStreamSynchronize(st));
  }
  if (en->hist_x) { cudaFree(en->hist_x); cudaFree(en->hist_lp); }
  en->hist_x = hx; en->hist_lp = hl; en->hist_cap = cap;
  return 0;
}

template <typename T>
int ensemble_stage(T** dst, const T* src_host, size_t count, cudaStream_t st) {
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  if (!src_host) return 0;
  CU(cudaMalloc(dst, std::max<size_t>(count, 1) * sizeof(T)));
  CU(cudaMemcpyAsync(*dst, src_host, count * sizeof(T), cudaMemcpyHostToDevice, st));
  return 0;
}

}  // namespace

extern "C" int gpbt_ensemble_create(gpbt_ensemble_t* out, gpbt_chain_t chain, int n_walkers, double a,
                                    int randomize_split, uint64_t seed) {
  if (!out || !chain) return fail(GPBT_EINVAL, "gpbt_ensemble_create: null argument");
  if (n_walkers < 2) return fail(GPBT_EINVAL, "gpbt_ensemble_create: need at least 2 walkers, got %d", n_walkers);
  if (!(a > 1.0)) return fail(GPBT_EINVAL, "gpbt_ensemble_create: stretch scale a must be > 1, got %g", a);
  CU(cudaSetDevice(chain->device));
  gpbt_ensemble* en = new gpbt_ensemble();
  en->ch = chain; en->nw = n_walkers; en->p = chain->p; en->n_half = (n_walkers + 1) / 2;
  en->a = a; en->randomize = randomize_split ? 1 : 0; en->seed = seed;
  const size_t nw = n_walkers, p = chain->p, nh = en->n_half;
  cudaError_t e = cudaSuccess;
  auto A = [&](void** ptr, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(ptr, bytes); };
  A((void**)&en->ctl, sizeof(EnsembleCtl));
  A((void**)&en->x, nw * p * 8); A((void**)&en->lp, nw * 8);
  A((void**)&en->q, nh * p * 8); A((void**)&en->factor, nh * 8); A((void**)&en->u_acc, nh * 8);
  A((void**)&en->lp_new, nh * 8); A((void**)&en->perm, nw * 4); A((void**)&en->notpd_call, 4);
  A((void**)&en->accepted, nw * 8); A((void**)&en->notpd_total, 8); A((void**)&en->rec_done, 4);
  A((void**)&en->keys, nw * 8);
  if (e == cudaSuccess) e = cudaMemset(en->accepted, 0, nw * 8);
  if (e == cudaSuccess) e = cudaMemset(en->notpd_total, 0, 8);
  if (e == cudaSuccess) e = cudaMemset(en->notpd_call, 0, 4);
  if (e == cudaSuccess) e = cudaMemset(en->rec_done, 0, 4);
  if (e != cudaSuccess) {
    gpbt_ensemble_destroy(en);
    return fail((int)e, "gpbt_ensemble_create: %s", cudaGetErrorString(e));
  }
  *out = en;
  return 0;
}

extern "C" int gpbt_ensemble_destroy(gpbt_ensemble_t en) {
  if (!en) return 0;
  cudaSetDevice(en->ch->device);
  cudaStreamSynchronize(en->ch->stream);
  if (en->graph) cudaGraphExecDestroy(en->graph);
  void* bufs[] = {en->ctl, en->x, en->lp, en->q, en->factor, en->u_acc, en->lp_new, en->perm, en->notpd_call,
                  en->accepted, en->notpd_total, en->rec_done, en->keys, en->hist_x, en->hist_lp, en->ru, en->rp, en->perm_in};
  for (void* b : bufs)
    if (b) cudaFree(b);
  delete en;
  return 0;
}

extern "C" int gpbt_ensemble_set_state(gpbt_ensemble_t en, const double* X_host, const double* lp_host) {
  if (!en || !X_host) return fail(GPBT_EINVAL, "gpbt_ensemble_set_state: null argument");
  CU(cudaSetDevice(en->ch->device));
  cudaStream_t st = en->ch->stream;
  CU(cudaMemcpyAsync(en->x, X_host, (size_t)en->nw * en->p * sizeof(double), cudaMemcpyHostToDevice, st));
  if (lp_host) {
    CU(cudaMemcpyAsync(en->lp, lp_host, (size_t)en->nw * sizeof(double), cudaMemcpyHostToDevice, st));
  } else {
    if (int r = gpbt_log_posterior(en->ch, en->x, -INFINITY, en->lp, en->notpd_call, en->nw, GPBT_PATH_AUTO, st))
      return r;
  }
  CU(cudaStreamSynchronize(st));
  en->has_state = true;
  return 0;
}

extern "C" int gpbt_ensemble_get_state(gpbt_ensemble_t en, double* X_host, double* lp_host) {
  if (!en) return fail(GPBT_EINVAL, "gpbt_ensemble_get_state: null handle");
  if (!en->has_state) return fail(GPBT_EINVAL, "gpbt_ensemble_get_state: no state set");
  CU(cudaSetDevice(en->ch->device));
  cudaStream_t st = en->ch->stream;
  if (X_host) CU(cudaMemcpyAsync(X_host, en->x, (size_t)en->nw * en->p * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (lp_host) CU(cudaMemcpyAsync(lp_host, en->lp, (size_t)en->nw * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

namespace {
// room for n_steps more steps, the optional host draws on the device, and the control block
int ensemble_prepare(gpbt_ensemble* en, int64_t n_steps, const double* u_host, const int32_t* partner_host,
                     const int32_t* perm_host) {
  if (!en || n_steps < 0) return fail(GPBT_EINVAL, "gpbt_ensemble: bad argument");
  if (!en->has_state) return fail(GPBT_EINVAL, "gpbt_ensemble: call gpbt_ensemble_set_state first");
  if ((u_host == nullptr) != (partner_host == nullptr))
    return fail(GPBT_EINVAL, "gpbt_ensemble: u_host and partner_host come together");
  CU(cudaSetDevice(en->ch->device));
  cudaStream_t st = en->ch->stream;
  if (int r = ensemble_grow_history(en, en->steps + n_steps, st)) return r;
  const size_t rows = (size_t)n_steps * 2 * en->n_half;
  if (int r = ensemble_stage(&en->ru, u_host, rows * 2, st)) return r;
  if (int r = ensemble_stage(&en->rp, partner_host, rows, st)) return r;
  if (int r = ensemble_stage(&en->perm_in, perm_host, (size_t)n_steps * en->nw, st)) return r;
  EnsembleCtl h;
  h.step = en->steps; h.run_first = en->steps;
  h.ru = en->ru; h.rp = en->rp; h.perm_in = en->perm_in;
  h.hist_x = en->hist_x; h.hist_lp = en->hist_lp;
  CU(cudaMemcpyAsync(en->ctl, &h, sizeof h, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));   // h lives on this stack frame
  en->prepared_until = en->steps + n_steps;
  return 0;
}
}  // namespace

// ---- a sampler step in pieces, for callers that evaluate the proposals themselves -----------------
extern "C" int gpbt_ensemble_prepare(gpbt_ensemble_t en, int64_t n_steps) {
  return ensemble_prepare(en, n_steps, nullptr, nullptr, nullptr);
}

extern "C" int gpbt_ensemble_begin_half(gpbt_ensemble_t en, int half, void* stream) {
  if (!en || (half != 0 && half != 1)) return fail(GPBT_EINVAL, "gpbt_ensemble_begin_half: bad argument");
  if (en->steps >= en->prepared_until)
    return fail(GPBT_EINVAL, "gpbt_ensemble_begin_half: no prepared step left (call gpbt_ensemble_prepare)");
  cudaStream_t st = (cudaStream_t)stream;
  const int nw = en->nw;
  const EnsembleBuffers b = ensemble_buffers(en);
  if (half == 0) {
    ensemble_keys_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(en->ctl, en->seed, nw, en->randomize, en->keys,
                                                                      en->perm);
    LAUNCH_CHECK();
    if (nw <= kEnsembleRankMaxWalkers) {
      ensemble_rank_kernel<<<(unsigned)((nw + kRankThreads - 1) / kRankThreads), kRankThreads, 0, st>>>(
          en->ctl, nw, en->randomize, en->keys, en->perm);
      LAUNCH_CHECK();
    }
  }
  const int ns = half == 0 ? en->n_half : nw - en->n_half;
  if (ns > 0) {
    ensemble_propose_kernel<<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(b, half);
    LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int gpbt_ensemble_copy_proposals(gpbt_ensemble_t en, int half, int64_t first, int64_t n, double* dst_dev,
                                            void* stream) {
  if (!en || !dst_dev || first < 0 || n < 0) return fail(GPBT_EINVAL, "gpbt_ensemble_copy_proposals: bad argument");
  const int64_t ns = half == 0 ? en->n_half : en->nw - en->n_half;
  if (n == 0) return 0;
  if (ns == 0) return fail(GPBT_EINVAL, "gpbt_ensemble_copy_proposals: the active set is empty");
  ensemble_gather_rows_kernel<<<(unsigned)((n * en->p + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      en->q, ns, en->p, first, n, dst_dev);
  LAUNCH_CHECK();
  return 0;
}

extern "C" int gpbt_ensemble_end_half(gpbt_ensemble_t en, int half, const double* lp_new_dev, void* stream) {
  if (!en || !lp_new_dev || (half != 0 && half != 1)) return fail(GPBT_EINVAL, "gpbt_ensemble_end_half: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int nw = en->nw, p = en->p;
  EnsembleBuffers b = ensemble_buffers(en);
  b.lp_new = lp_new_dev;
  const int ns = half == 0 ? en->n_half : nw - en->n_half;
  if (ns > 0) {
    ensemble_accept_kernel<<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(b, half);
    LAUNCH_CHECK();
  }
  if (half == 1) {
    const unsigned rec_grid = (unsigned)std::min<int64_t>(((int64_t)nw * p + 255) / 256, device_info().sm_count);
    ensemble_record_kernel<<<rec_grid, 256, 0, st>>>(en->ctl, nw, p, en->x, en->lp, en->rec_done);
    LAUNCH_CHECK();
    en->steps += 1;
  }
  return 0;
}

extern "C" int gpbt_ensemble_run(gpbt_ensemble_t en, int64_t n_steps, const double* u_host,
                                 const int32_t* partner_host, const int32_t* perm_host, int use_graph) {
  if (int r = ensemble_prepare(en, n_steps, u_host, partner_host, perm_host)) return r;
  if (n_steps == 0) return 0;
  cudaStream_t st = en->ch->stream;

  int64_t done = 0;
  // graph replay pays off where a step is launch bound; large ensembles (separate-kernel shape) are
  // compute bound and measured 2 % faster with plain stream launches
  if (use_graph && en->nw <= kEnsembleFusedMaxWalkers) {
    if (en->graph && en->graph_gen != g_ws_generation.load()) {
      cudaGraphExecDestroy(en->graph);
      en->graph = nullptr;
    }
    if (!en->graph) {
      // the first step runs eagerly, so that every workspace the path needs exists before the capture
      if (int r = ensemble_enqueue_step(en, st)) return r;
      done = 1;
      if (n_steps > 1) {
        cudaGraph_t g = nullptr;
        CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
        const int64_t before = g_launches.load();
        const int r = ensemble_enqueue_step(en, st);
        en->graph_launches = g_launches.load() - before;
        g_launches.store(before);   // captured, not launched
        const cudaError_t ce = cudaStreamEndCapture(st, &g);
        if (r || ce != cudaSuccess) {
          if (g) cudaGraphDestroy(g);
          return r ? r : fail((int)ce, "gpbt_ensemble_run: graph capture failed: %s", cudaGetErrorString(ce));
        }
        const cudaError_t ie = cudaGraphInstantiate(&en->graph, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) {
          en->graph = nullptr;
          return fail((int)ie, "gpbt_ensemble_run: cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
        }
        en->graph_gen = g_ws_generation.load();
      }
    }
    for (; done < n_steps; done++) {
      CU(cudaGraphLaunch(en->graph, st));
      g_launches.fetch_add(en->graph_launches, std::memory_order_relaxed);
    }
  } else {
    for (; done < n_steps; done++)
      if (int r = ensemble_enqueue_step(en, st)) return r;
  }
  CU(cudaStreamSynchronize(st));
  en->steps += n_steps;
  return 0;
}

extern "C" int64_t gpbt_ensemble_steps(gpbt_ensemble_t en) { return en ? en->steps : 0; }

extern "C" int gpbt_ensemble_reserve(gpbt_ensemble_t en, int64_t n_steps) {
  if (!en || n_steps < 0) return fail(GPBT_EINVAL, "gpbt_ensemble_reserve: bad argument");
  CU(cudaSetDevice(en->ch->device));
  return ensemble_grow_history(en, en->steps + n_steps, en->ch->stream, /*exact=*/true);
}

extern "C" int gpbt_ensemble_read(gpbt_ensemble_t en, int64_t first, int64_t n, double* chain_host, double* lp_host,
                                  int64_t* accepted_host, int64_t* n_notpd_host) {
  if (!en || first < 0 || n < 0 || first + n > en->steps)
    return fail(GPBT_EINVAL, "gpbt_ensemble_read: rows [%lld, %lld) outside the %lld recorded steps",
                (long long)first, (long long)(first + n), (long long)(en ? en->steps : 0));
  CU(cudaSetDevice(en->ch->device));
  cudaStream_t st = en->ch->stream;
  const size_t row = (size_t)en->nw * en->p;
  if (chain_host && n > 0)
    CU(cudaMemcpyAsync(chain_host, en->hist_x + first * row, n * row * sizeof(double), cudaMemcpyDeviceToHost, st));
  if (lp_host && n > 0)
    CU(cudaMemcpyAsync(lp_host, en->hist_lp + first * en->nw, (size_t)n * en->nw * sizeof(double),
                       cudaMemcpyDeviceToHost, st));
  static_assert(sizeof(long long) == sizeof(int64_t), "accepted counters are copied as int64_t");
  if (accepted_host)
    CU(cudaMemcpyAsync(accepted_host, en->accepted, (size_t)en->nw * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  if (n_notpd_host) CU(cudaMemcpyAsync(n_notpd_host, en->notpd_total, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return 0;
}

extern "C" int gpbt_ensemble_reset(gpbt_ensemble_t en) {
  if (!en) return fail(GPBT_EINVAL, "gpbt_ensemble_reset: null handle");
  CU(cudaSetDevice(en->ch->device));
  CU(cudaMemsetAsync(en->accepted, 0, (size_t)en->nw * sizeof(long long), en->ch->stream));
  CU(cudaStreamSynchronize(en->ch->stream));
  en->steps = 0;
  en->prepared_until = 0;   // the device step counter is rewritten by the next prepare / run
  return 0;
}

// ---- host helper of the parallel-tempering driver ------------------------------------------------
// One sweep of Chain.tempexchange (src/mcmc.py:679-693): the swaps depend on each other, so the loop
// is sequential by nature; in Python it costs ~3 us per pick (120 ms per PTLMC iteration at 8192
// chains, ten times the GPU call it sits next to), here ~1 ns.  The random picks and log-uniform
// draws are made by the caller (NumPy's generator, in the reference's order).
extern "C" int gpbt_host_temp_exchange(const double* lp, const double* temps, int64_t n, const int64_t* picks,
                                       const double* log_u, int64_t n_picks, int64_t* order) {
  if (!lp || !temps || !picks || !log_u || !order || n < 0 || n_picks < 0)
    return fail(GPBT_EINVAL, "gpbt_host_temp_exchange: bad argument");
  for (int64_t i = 0; i < n_picks; i++) {
    const int64_t rt = picks[i];
    if (rt < 1 || rt >= n) return fail(GPBT_EINVAL, "gpbt_host_temp_exchange: pick %lld outside [1, n)", (long long)rt);
    const double gap = 1.0 / temps[rt - 1] - 1.0 / temps[rt];
    const volatile double diff = lp[order[rt]] - lp[order[rt - 1]];   // rounded before the product, as NumPy does
    if (diff * gap > log_u[i]) {
      const int64_t tmp = order[rt - 1];
      order[rt - 1] = order[rt];
      order[rt] = tmp;
    }
  }
  return 0;
}

#include "fanout.inl"
#include "ptlmc.inl"
