// Single-process multi-GPU fan-out behind the host boundary call (included at the end of gpbt_api.cu).
//
// The samplers of the reference are single-process and hand ONE host array X[N, p] to
// Chain.log_posterior / log_likelihood (src/mcmc.py:188-222, 261-299; pocoMC passes all active particles,
// :798-804; emcee through pool=self, :372-374).  Rows are independent, so a fan-out owns one replica of
// the chain per GPU and one worker thread per replica: each worker takes a contiguous row block, stages
// it through its own pinned buffer (H2D in row chunks that overlap the kernels of the chunk before), runs
// the log-posterior path on its GPU's stream and copies its block of lp straight into the caller's
// array.  There is no collective in this form -- every block travels host -> its GPU -> host.
#include <chrono>
#include <condition_variable>
#include <thread>

namespace {

constexpr int64_t kFanoutDefaultMinRows = 256;    // rows per GPU below which adding a GPU does not pay (measured:
                                                  // two GPUs beat one from 512 rows per call at config 2)
constexpr int64_t kFanoutChunkRows = 32768;       // staging granularity of a worker
// A sampler calls back within micro- to milliseconds: after a job a worker (and the caller, while it waits
// for the workers) polls for this long before it blocks on the condition variable, whose wake-up costs
// 20-50 us -- a third of an 8-GPU call of 4096 rows.
constexpr int64_t kFanoutSpinMicros = 300;

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#else
  std::this_thread::yield();
#endif
}

template <typename Pred>
bool spin_until(Pred pred, int64_t micros) {
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    for (int i = 0; i < 64; i++) {
      if (pred()) return true;
      cpu_relax();
    }
    if (std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() > micros)
      return pred();
  }
}

struct FanoutWorker {
  gpbt_chain_t ch = nullptr;
  std::thread th;
  std::mutex m;
  std::condition_variable cv;
  std::atomic<bool> has_job{false}, done{false}, quit{false};
  // job
  const double* X = nullptr;
  double* lp = nullptr;
  int64_t N = 0;
  double oob = 0.0;
  int path = 0;
  // result
  int rc = 0, notpd = 0;
  std::string err;
  // pinned staging: two X slots of kFanoutChunkRows rows, lp for the whole block
  double* pin_x[2] = {nullptr, nullptr};
  cudaEvent_t slot_free[2] = {nullptr, nullptr};
  double* pin_lp = nullptr;
  int* pin_cnt = nullptr;
  int64_t cap_lp = 0;
};

int fanout_worker_job(FanoutWorker* w) {
  gpbt_chain* ch = w->ch;
  CU(cudaSetDevice(ch->device));
  cudaStream_t st = ch->stream;
  const int64_t N = w->N;
  const int p = ch->p;
  if (N <= kZeroCopyRows && !g_opt.no_zerocopy.load())   // small block: the mapped-memory path, no explicit copies
    return gpbt_log_posterior_host(ch, w->X, w->oob, w->lp, &w->notpd, N, w->path);
  if (!w->pin_x[0]) {
    for (int s = 0; s < 2; s++) {
      CU(cudaHostAlloc(&w->pin_x[s], (size_t)kFanoutChunkRows * p * sizeof(double), cudaHostAllocDefault));
      CU(cudaEventCreateWithFlags(&w->slot_free[s], cudaEventDisableTiming));
    }
    CU(cudaHostAlloc(&w->pin_cnt, sizeof(int), cudaHostAllocDefault));
  }
  if (N > w->cap_lp) {
    const int64_t cap = std::max<int64_t>(N, 2 * w->cap_lp);
    if (w->pin_lp) cudaFreeHost(w->pin_lp);
    w->pin_lp = nullptr;
    w->cap_lp = 0;
    CU(cudaHostAlloc(&w->pin_lp, (size_t)cap * sizeof(double), cudaHostAllocDefault));
    w->cap_lp = cap;
  }
  if (int r = ensure_io(ch, N)) return r;
  CU(cudaMemsetAsync(ch->notpd_dev, 0, sizeof(int), st));
  int slot = 0;
  for (int64_t s = 0; s < N; s += kFanoutChunkRows, slot ^= 1) {
    const int64_t nn = std::min(kFanoutChunkRows, N - s);
    if (s >= 2 * kFanoutChunkRows) CU(cudaEventSynchronize(w->slot_free[slot]));   // its H2D has left the slot
    memcpy(w->pin_x[slot], w->X + s * p, (size_t)nn * p * sizeof(double));
    CU(cudaMemcpyAsync(ch->x_dev + s * p, w->pin_x[slot], (size_t)nn * p * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(w->slot_free[slot], st));
    if (int r = log_posterior_impl(ch, ch->x_dev + s * p, w->oob, ch->lp_dev + s, ch->notpd_dev, nn, w->path, st, nullptr,
                                   0, 0, /*zero_counter=*/false))
      return r;
  }
  CU(cudaMemcpyAsync(w->pin_lp, ch->lp_dev, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(w->pin_cnt, ch->notpd_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  memcpy(w->lp, w->pin_lp, (size_t)N * sizeof(double));
  w->notpd = *w->pin_cnt;
  return 0;
}

void fanout_worker_main(FanoutWorker* w) {
  for (;;) {
    if (!spin_until([&] { return w->has_job.load(std::memory_order_acquire) || w->quit.load(); }, kFanoutSpinMicros)) {
      std::unique_lock<std::mutex> lock(w->m);
      w->cv.wait(lock, [&] { return w->has_job.load() || w->quit.load(); });
    }
    if (w->quit.load()) return;
    const int rc = fanout_worker_job(w);
    {
      std::lock_guard<std::mutex> lock(w->m);
      w->rc = rc;
      w->err = rc ? g_err : std::string();
      w->has_job.store(false);
      w->done.store(true, std::memory_order_release);
    }
    w->cv.notify_all();
  }
}

}  // namespace

struct gpbt_fanout {
  std::vector<FanoutWorker*> workers;
  int p = 0;
};

extern "C" int gpbt_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int gpbt_set_device(int device) {
  CU(cudaSetDevice(device));
  return 0;
}

extern "C" int gpbt_get_device(void) { return current_device(); }

extern "C" int gpbt_set_option(const char* key, const char* value) {
  if (int r = set_option_value(key, value)) return fail(r, "gpbt_set_option: unknown key or bad value (%s = %s)", key ? key : "(null)", value ? value : "(null)");
  return 0;
}

extern "C" int gpbt_fanout_create(gpbt_fanout_t* out, const gpbt_chain_t* chains, int n_chains) {
  if (!out || !chains || n_chains < 1) return fail(GPBT_EINVAL, "gpbt_fanout_create: bad argument");
  for (int i = 0; i < n_chains; i++) {
    if (!chains[i]) return fail(GPBT_EINVAL, "gpbt_fanout_create: chain %d is null", i);
    if (chains[i]->p != chains[0]->p || chains[i]->M != chains[0]->M)
      return fail(GPBT_EINVAL, "gpbt_fanout_create: chain %d is not a replica of chain 0", i);
    for (int j = 0; j < i; j++)
      if (chains[j]->device == chains[i]->device)
        return fail(GPBT_EINVAL, "gpbt_fanout_create: chains %d and %d live on the same device", j, i);
  }
  gpbt_fanout* f = new gpbt_fanout();
  f->p = chains[0]->p;
  for (int i = 0; i < n_chains; i++) {
    FanoutWorker* w = new FanoutWorker();
    w->ch = chains[i];
    w->th = std::thread(fanout_worker_main, w);
    f->workers.push_back(w);
  }
  *out = f;
  return 0;
}

extern "C" int gpbt_fanout_destroy(gpbt_fanout_t f) {
  if (!f) return 0;
  const int prev = current_device();
  for (FanoutWorker* w : f->workers) {
    {
      std::lock_guard<std::mutex> lock(w->m);
      w->quit.store(true);
    }
    w->cv.notify_all();
    if (w->th.joinable()) w->th.join();
    cudaSetDevice(w->ch->device);
    for (int s = 0; s < 2; s++) {
      if (w->pin_x[s]) cudaFreeHost(w->pin_x[s]);
      if (w->slot_free[s]) cudaEventDestroy(w->slot_free[s]);
    }
    if (w->pin_lp) cudaFreeHost(w->pin_lp);
    if (w->pin_cnt) cudaFreeHost(w->pin_cnt);
    delete w;
  }
  cudaSetDevice(prev);
  delete f;
  return 0;
}

extern "C" int gpbt_fanout_size(gpbt_fanout_t f) { return f ? (int)f->workers.size() : 0; }

extern "C" int gpbt_fanout_log_posterior_host(gpbt_fanout_t f, const double* X_host, double oob_value, double* lp_host,
                                              int* n_notpd_host, int64_t N, int path, int max_devices,
                                              int* devices_used) {
  if (!f || !X_host || !lp_host || N < 0) return fail(GPBT_EINVAL, "gpbt_fanout_log_posterior_host: bad argument");
  if (devices_used) *devices_used = 0;
  if (n_notpd_host) *n_notpd_host = 0;
  if (N == 0) return 0;
  const int64_t opt_rows = g_opt.fanout_min_rows.load();
  const int64_t min_rows = opt_rows > 0 ? opt_rows : kFanoutDefaultMinRows;
  int G = (int)f->workers.size();
  if (max_devices > 0) {
    G = std::min(G, max_devices);
  } else {
    G = (int)std::max<int64_t>(1, std::min<int64_t>(G, N / min_rows));
  }
  if (devices_used) *devices_used = G;
  if (G == 1) {
    // (the caller's thread, the first replica: small batches keep the zero-copy latency path)
    const int prev = current_device();
    const int rc = gpbt_log_posterior_host(f->workers[0]->ch, X_host, oob_value, lp_host, n_notpd_host, N, path);
    cudaSetDevice(prev);
    return rc;
  }
  // contiguous row blocks, multiples of 16 walkers (kernel (a)'s tile) except the last
  int64_t per = (N + G - 1) / G;
  per = (per + 15) / 16 * 16;
  int used = 0;
  for (int i = 0; i < G; i++) {
    const int64_t lo = std::min<int64_t>(N, (int64_t)i * per), hi = std::min<int64_t>(N, lo + per);
    if (hi <= lo) break;
    FanoutWorker* w = f->workers[i];
    {
      std::lock_guard<std::mutex> lock(w->m);
      w->X = X_host + lo * f->p;
      w->lp = lp_host + lo;
      w->N = hi - lo;
      w->oob = oob_value;
      w->path = path;
      w->done.store(false);
      w->has_job.store(true, std::memory_order_release);
    }
    w->cv.notify_all();
    used++;
  }
  int rc = 0, notpd = 0, bad_device = -1;
  std::string err;
  for (int i = 0; i < used; i++) {   // every worker is waited for, also after a failure
    FanoutWorker* w = f->workers[i];
    if (!spin_until([&] { return w->done.load(std::memory_order_acquire); }, 20 * kFanoutSpinMicros)) {
      std::unique_lock<std::mutex> lock(w->m);
      w->cv.wait(lock, [&] { return w->done.load(); });
    }
    std::lock_guard<std::mutex> lock(w->m);   // (the worker publishes rc / err under the lock)
    if (w->rc && !rc) {
      rc = w->rc;
      err = w->err;
      bad_device = w->ch->device;
    }
    notpd += w->notpd;
  }
  if (devices_used) *devices_used = used;
  if (rc) return fail(rc, "gpbt_fanout_log_posterior_host: device %d: %s", bad_device, err.c_str());
  if (n_notpd_host) *n_notpd_host = notpd;
  return 0;
}
