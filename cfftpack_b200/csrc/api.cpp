/*
 * api.cpp -- the FFTPACK 5.1 entry points (include/cfftpack_b200.h) over the CUDA drivers.
 *
 * Each routine performs the argument checks of its reference twin with the reference's own expressions, so
 * that ier agrees value for value (cfft1f_ fftpack.c:2218-2228, cfftmf_ :2577-2590, cfft2f_ :2392-2405,
 * rfft1f_ :13053-13065, rfftmf_ :14057-14069, cost1f_ :6071-6084, costmf_ :6510-6530, sint1f_ :14640-14651,
 * sintmf_ :15025-15045, cosq1f_ :5480-5485, cosqmf_ :5870-5890), then hands DEVICE pointers to dispatch.cu.
 * Host arrays are staged through HBM here.  Nothing in this file computes a transform on the CPU.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/cfftpack_b200.h"
#include "internal.h"
#include "plan.h"
#include "tma.cuh"

namespace cfb {

/* ---------------- runtime plumbing ---------------- */
static thread_local char t_err[512] = "";
static thread_local cudaStream_t t_stream = 0;
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}
const char *last_error() { return t_err; }
bool cuda_ok(cudaError_t e, const char *what) {
  if (e == cudaSuccess) return true;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return false;
}
cudaStream_t current_stream() { return t_stream; }
void set_current_stream(cudaStream_t s) { t_stream = s; }
void count_launch(unsigned long long k) { g_launches += k; }
unsigned long long launch_count() { return g_launches.load(); }

bool device_ready() {
  static std::once_flag once;
  static bool ok = false;
  static char why[256] = "";
  std::call_once(once, [] {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
      snprintf(why, sizeof(why), "no CUDA device (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
      fprintf(stderr,
              "libcfftpack_b200: FATAL: %s. This library has no CPU path; transforms will fail with ier = -1.\n", why);
      return;
    }
    ok = true;
  });
  if (!ok) set_error("%s", why);
  return ok;
}

bool kernel_attrs_ready(const void *kernel, size_t smem) {
  static std::mutex mu;
  static std::vector<std::pair<std::pair<const void *, int>, bool>> done;  // few dozen entries: linear search is fine
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  for (auto &e : done)
    if (e.first.first == kernel && e.first.second == dev) return e.second;
  const bool ok = cuda_ok(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(max dynamic shared memory)") &&
                  cuda_ok(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100),
                          "cudaFuncSetAttribute(shared memory carveout)");
  done.push_back({{kernel, dev}, ok});
  return ok;
}

bool make_tensor_map3(TensorMap3 *tm, const void *base, unsigned long long d0, unsigned long long d1, unsigned long long d2,
                      unsigned long long s1, unsigned long long s2, unsigned b0, unsigned b1) {
#ifdef CFB_SIM
  tm->base = (const double *)base;
  tm->dim[0] = d0; tm->dim[1] = d1; tm->dim[2] = d2;
  tm->stride_bytes[0] = 8; tm->stride_bytes[1] = s1; tm->stride_bytes[2] = s2;
  tm->box[0] = b0; tm->box[1] = b1; tm->box[2] = 1;
  return true;
#else
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static encode_fn encode = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      encode = (encode_fn)fn;
  });
  if (!encode) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return false;
  }
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {s1, s2};  // bytes, dims 1 and 2
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<void *>(base), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return false;
  }
  return true;
#endif
}

static thread_local int t_tile_cap = 0;
void set_tile_cta_cap(int ctas) { t_tile_cap = ctas; }
int tile_cta_cap() { return t_tile_cap; }

int sm_count() {  // of the calling thread's current device
  static std::mutex mu;
  static int sms[64];
  int d = 0;
  cudaGetDevice(&d);
  if (d < 0 || d >= 64) d = 0;
  std::lock_guard<std::mutex> lk(mu);
  if (!sms[d]) {
    cudaDeviceProp p;
    sms[d] = cudaGetDeviceProperties(&p, d) == cudaSuccess && p.multiProcessorCount > 0 ? p.multiProcessorCount : 148;
  }
  return sms[d];
}

/* Plan tables: copy on the caller's stream and wait for the DMA itself.  A pageable cudaMemcpy only promises that the
 * source has been staged, and the non-blocking streams the transforms may run on do not order against the NULL stream. */
void *upload_table(const void *host, size_t bytes) {
  void *d = nullptr;
  if (!cuda_ok(cudaMalloc(&d, bytes ? bytes : 16), "cudaMalloc(plan table)")) return nullptr;
  if (bytes && !(cuda_ok(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, t_stream), "cudaMemcpyAsync(plan table)") &&
                 cuda_ok(cudaStreamSynchronize(t_stream), "cudaStreamSynchronize(plan table)"))) {
    cudaFree(d);
    return nullptr;
  }
  return d;
}

struct Scratch {
  void *p = nullptr;
  size_t cap = 0;
};
/* Scratch is private to a (host thread, device, stream): transforms issued by one thread on different streams -- the three
 * streams of the host-array pipeline below, or a caller alternating cfb200_set_stream -- run concurrently on the GPU, so
 * they must not share intermediates (four-step, chirp-z, long real).  Switching device or stream keeps every set alive. */
struct ScratchSet {
  int dev = -1;
  cudaStream_t stream = 0;
  Scratch s[10];  // 0 four-step, 1 staged host array, 2 chirp-z, 3 long real, 4-6 staging pipeline, 7 rfft2 pairs, 8 long 1-D
};
struct ScratchSets {
  std::vector<ScratchSet *> sets;
  void free_all() {
    for (auto *set : sets) {
      for (auto &x : set->s)
        if (x.p) cudaFree(x.p);
      delete set;
    }
    sets.clear();
  }
  ~ScratchSets() { free_all(); }
};
static thread_local ScratchSets t_scr;
void *scratch_get(int slot, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  ScratchSet *set = nullptr;
  for (auto *c : t_scr.sets)
    if (c->dev == dev && c->stream == t_stream) set = c;
  if (!set) {
    if (t_scr.sets.size() >= 16) {  // a caller cycling through short-lived streams: start over rather than grow without bound
      cudaDeviceSynchronize();
      t_scr.free_all();
    }
    set = new ScratchSet();
    set->dev = dev;
    set->stream = t_stream;
    t_scr.sets.push_back(set);
  }
  Scratch &s = set->s[slot];
  if (s.cap >= bytes && s.p) return s.p;
  if (s.p) {
    cudaStreamSynchronize(t_stream);
    cudaFree(s.p);
    s.p = nullptr;
    s.cap = 0;
  }
  size_t cap = bytes + bytes / 8 + 4096;
  if (!cuda_ok(cudaMalloc(&s.p, cap), "cudaMalloc(scratch)")) {
    s.p = nullptr;
    return nullptr;
  }
  s.cap = cap;
  return s.p;
}
void scratch_release_all() { t_scr.free_all(); }

/* Short host arrays (a single cfft1f_ of a few thousand points: config 1) skip the two copy-engine launches: the data
 * is placed in a per-thread pinned, device-mapped bounce buffer that the kernels read and write across PCIe directly. */
static const size_t BOUNCE_BYTES = 64 << 10;
struct Bounce {
  void *p = nullptr;
  bool tried = false;
  ~Bounce() {
    if (p) cudaFreeHost(p);
  }
};
static thread_local Bounce t_bounce;
static void *bounce_get() {
  if (!t_bounce.tried) {
    t_bounce.tried = true;
    static const bool off = getenv("CFB200_NO_BOUNCE") != nullptr;
    if (off || cudaHostAlloc(&t_bounce.p, BOUNCE_BYTES, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      t_bounce.p = nullptr;
    }
  }
  return t_bounce.p;
}

bool view_open(void *user, size_t bytes, DeviceView &v) {
  if (!device_ready()) return false;
  cudaPointerAttributes at;
  memset(&at, 0, sizeof(at));
  cudaError_t e = cudaPointerGetAttributes(&at, user);
  if (e != cudaSuccess) {
    cudaGetLastError();
    at.type = cudaMemoryTypeUnregistered;
  }
  v.bytes = bytes;
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) {
    v.dev = user;
    return true;
  }
  v.host = user;
  v.staged = true;
  if (bytes <= BOUNCE_BYTES && bounce_get()) {
    v.mapped = true;
    v.dev = t_bounce.p;  // unified addressing: the mapped buffer has the same address on the device
    memcpy(v.dev, user, bytes);
    return true;
  }
  v.dev = scratch_get(1, bytes);
  if (!v.dev) return false;
  return cuda_ok(cudaMemcpyAsync(v.dev, user, bytes, cudaMemcpyHostToDevice, t_stream), "cudaMemcpyAsync(H2D)");
}
bool view_close(DeviceView &v, bool ok) {
  if (!v.staged) return ok;
  if (v.mapped) {
    bool s = cuda_ok(cudaStreamSynchronize(t_stream), "cudaStreamSynchronize");
    if (ok && s) memcpy(v.host, v.dev, v.bytes);
    return ok && s;
  }
  if (ok) ok = cuda_ok(cudaMemcpyAsync(v.host, v.dev, v.bytes, cudaMemcpyDeviceToHost, t_stream), "cudaMemcpyAsync(D2H)");
  bool s = cuda_ok(cudaStreamSynchronize(t_stream), "cudaStreamSynchronize");
  return ok && s;
}

/* ---- batched host arrays: pipeline lot-chunks through HBM so that the H2D copy of chunk c+1, the transform of
 * chunk c and the D2H copy of chunk c-1 overlap (PCIe is full duplex; the copy engines are independent).
 * Applies when the sequences are not interleaved (jump >= extent of one sequence). ---- */
static const int PIPE_BUFS = 3;
struct PipeStreams {
  cudaStream_t st[PIPE_BUFS] = {0, 0, 0};
  bool ok = false;
  ~PipeStreams() {
    if (ok)
      for (auto s : st) cudaStreamDestroy(s);
  }
};
static thread_local PipeStreams t_pipe;

/* ---- several GPUs behind one C call (host arrays of batched transforms): the lot axis is cut into one contiguous
 * shard per device (independent sequences, no exchange) and each shard is staged and transformed by a
 * persistent worker thread bound to its device.  Everything a transform needs (stream, scratch, plans, kernel attributes,
 * SM count) is already keyed per host thread and device, so the workers run the ordinary single-GPU path.
 * Off unless asked for: cfb200_set_devices(n) or CFB200_DEVICES=n|all (one process per GPU under torchrun must not fan out). */
struct DevicePool {
  struct Worker {
    std::thread th;
    std::function<bool()> job;
    bool has_job = false, result = true, quit = false;
    char err[512] = "";
  };
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::vector<Worker *> workers;  // worker i serves device i + 1 (device 0 is the calling thread's share)
  int pending = 0;
  void ensure(int n) {
    std::lock_guard<std::mutex> lk(mu);
    while ((int)workers.size() < n) {
      Worker *w = new Worker();
      const int dev = (int)workers.size() + 1;
      workers.push_back(w);
      w->th = std::thread([this, w, dev] {
        cudaSetDevice(dev);
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
          cv_job.wait(lk, [w] { return w->has_job || w->quit; });
          if (w->quit) return;
          auto job = w->job;
          lk.unlock();
          t_err[0] = 0;
          const bool ok = job();
          lk.lock();
          w->result = ok;
          snprintf(w->err, sizeof(w->err), "%s", t_err);
          w->has_job = false;
          if (--pending == 0) cv_done.notify_all();
        }
      });
      w->th.detach();  // lives for the life of the process
    }
  }
};
// never destroyed: the detached workers wait on its condition variable for the life of the process, and destroying a
// condition variable with waiters (static destruction at exit) blocks forever
static DevicePool &g_pool = *new DevicePool();
static std::atomic<int> g_devices{0};  // 0 = not decided yet
static int devices_in_use() {
  int d = g_devices.load();
  if (d > 0) return d;
  d = 1;
  if (const char *e = getenv("CFB200_DEVICES")) {
    int n = 0;
    cudaGetDeviceCount(&n);
    d = (!strcmp(e, "all") || atoi(e) > n) ? n : atoi(e);
    if (d < 1) d = 1;
  }
  g_devices.store(d);
  return d;
}

template <class F>
static bool run_on_array_one(void *user, size_t esz, long long lot, long long jump, int n, long long inc, F &&fn, int mem_type);

template <class F>
static bool run_on_array(void *user, size_t esz, long long lot, long long jump, int n, long long inc, F &&fn) {
  const long long seq_span = inc * (long long)(n - 1) + 1;
  if (!device_ready()) return false;
  cudaPointerAttributes at;
  memset(&at, 0, sizeof(at));
  if (cudaPointerGetAttributes(&at, user) != cudaSuccess) {
    cudaGetLastError();
    at.type = cudaMemoryTypeUnregistered;
  }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return fn(user, lot);
  const int G = devices_in_use();
  int cur = 0;
  cudaGetDevice(&cur);
  if (G <= 1 || cur != 0 || lot < 2 * G || jump < seq_span)
    return run_on_array_one(user, esz, lot, jump, n, inc, fn, (int)at.type);
  // fan out: shard g = sequences [lot g / G, lot (g+1) / G).  One fan-out at a time per process: the workers have one
  // job slot each (concurrent callers queue up here; their transfers would share the same PCIe links anyway)
  static std::mutex fan_mu;
  std::lock_guard<std::mutex> fan_lock(fan_mu);
  g_pool.ensure(G - 1);
  {
    std::lock_guard<std::mutex> lk(g_pool.mu);
    g_pool.pending = G - 1;
    for (int g = 1; g < G; ++g) {
      const long long m0 = lot * g / G, m1 = lot * (g + 1) / G;
      char *p = (char *)user + (size_t)m0 * jump * esz;
      DevicePool::Worker *w = g_pool.workers[g - 1];
      w->job = [=, &fn] { return run_on_array_one(p, esz, m1 - m0, jump, n, inc, fn, (int)at.type); };
      w->has_job = true;
    }
  }
  g_pool.cv_job.notify_all();
  bool ok = run_on_array_one(user, esz, lot / G, jump, n, inc, fn, (int)at.type);
  std::unique_lock<std::mutex> lk(g_pool.mu);
  g_pool.cv_done.wait(lk, [] { return g_pool.pending == 0; });
  for (int g = 1; g < G; ++g)
    if (!g_pool.workers[g - 1]->result) {
      ok = false;
      set_error("device %d: %s", g, g_pool.workers[g - 1]->err);
    }
  return ok;
}

template <class F>
static bool run_on_array_one(void *user, size_t esz, long long lot, long long jump, int n, long long inc, F &&fn, int mem_type) {
  const long long seq_span = inc * (long long)(n - 1) + 1;
  const size_t total = (size_t)((lot - 1) * jump + seq_span) * esz;
  cudaPointerAttributes at;
  at.type = (cudaMemoryType)mem_type;
  static const size_t CHUNK = [] {  // bytes per pipeline stage; CFB200_PIPE_CHUNK_KB overrides (tests)
    const char *e = getenv("CFB200_PIPE_CHUNK_KB");
    return e ? (size_t)atoll(e) << 10 : (size_t)64 << 20;
  }();
  const bool pipelined = lot >= 4 && jump >= seq_span && total >= 2 * CHUNK && at.type == cudaMemoryTypeHost;
  if (!pipelined) {
    DeviceView v;
    bool ok = view_open(user, total, v);
    if (ok) ok = fn(v.dev, lot);
    return view_close(v, ok);
  }
  if (!t_pipe.ok) {
    for (auto &s : t_pipe.st)
      if (!cuda_ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    t_pipe.ok = true;
  }
  long long per = (long long)(CHUNK / ((size_t)jump * esz));
  if (per < 1) per = 1;
  const size_t buf_bytes = (size_t)((per - 1) * jump + seq_span) * esz;
  void *bufs[PIPE_BUFS];
  for (int i = 0; i < PIPE_BUFS; ++i)
    if (!(bufs[i] = scratch_get(4 + i, buf_bytes))) return false;
  cudaStream_t saved = t_stream;
  if (!cuda_ok(cudaStreamSynchronize(saved), "cudaStreamSynchronize")) return false;
  bool ok = true;
  int c = 0;
  for (long long m0 = 0; m0 < lot && ok; m0 += per, ++c) {
    const long long lc = (lot - m0) < per ? (lot - m0) : per;
    const size_t bytes = (size_t)((lc - 1) * jump + seq_span) * esz;
    char *h = (char *)user + (size_t)m0 * jump * esz;
    const int b = c % PIPE_BUFS;
    t_stream = t_pipe.st[b];
    ok = cuda_ok(cudaMemcpyAsync(bufs[b], h, bytes, cudaMemcpyHostToDevice, t_stream), "cudaMemcpyAsync(H2D)") &&
         fn(bufs[b], lc) &&
         cuda_ok(cudaMemcpyAsync(h, bufs[b], bytes, cudaMemcpyDeviceToHost, t_stream), "cudaMemcpyAsync(D2H)");
  }
  for (auto s : t_pipe.st) ok = cuda_ok(cudaStreamSynchronize(s), "cudaStreamSynchronize") && ok;
  t_stream = saved;
  return ok;
}

static inline long long span1(int n, int inc) { return (long long)inc * (n - 1) + 1; }
static inline long long spanm(int lot, int jump, int n, int inc) {
  return (long long)(lot - 1) * jump + (long long)inc * (n - 1) + 1;
}

/* ---------------- complex ---------------- */
static int complex_init(int *n, double *wsave, int *lensav, int *ier) {
  *ier = 0;
  if (*n < 1) return 0;
  if (*lensav < 2 * *n + log2_floor_ref(*n) + 4) {
    *ier = 2;
    return 0;
  }
  if (*n == 1) return 0;
  wsave_init_complex(*n, wsave);
  return 0;
}

static int complex_1d(int *n, int *inc, void *c, int *lenc, int *lensav, int *lenwrk, int *ier, int dir) {
  *ier = 0;
  if (*n < 1) return 0;  // undefined upstream (log of a non-positive length): nothing to transform
  if (*inc < 1) {  // undefined upstream; here: the array cannot hold the sequence
    *ier = 1;
    return 0;
  }
  if (*lenc < span1(*n, *inc)) *ier = 1;
  else if (*lensav < 2 * *n + log2_floor_ref(*n) + 4) *ier = 2;
  else if (*lenwrk < 2 * *n) *ier = 3;
  if (*ier || *n == 1) return 0;
  DeviceView v;
  bool ok = view_open(c, (size_t)span1(*n, *inc) * 16, v);
  if (ok) ok = run_c2c(*n, 1, *inc, (long long)*inc * *n, dir, v.dev);
  ok = view_close(v, ok);
  if (!ok) *ier = -1;
  return 0;
}

static int complex_multi(int *lot, int *jump, int *n, int *inc, void *c, int *lenc, int *lensav, int *lenwrk, int *ier,
                         int dir) {
  *ier = 0;
  if (*n < 1 || *lot < 1) return 0;  // empty batch: the reference's loops run zero times (lot) / undefined (n)
  if (*inc < 1 || *jump < 0) {  // undefined upstream; reported as inconsistent strides
    *ier = 4;
    return 0;
  }
  if (*lenc < spanm(*lot, *jump, *n, *inc)) *ier = 1;
  else if (*lensav < 2 * *n + log2_floor_ref(*n) + 4) *ier = 2;
  else if ((long long)*lenwrk < 2LL * *lot * *n) *ier = 3;
  else if (!strides_consistent(*inc, *jump, *n, *lot)) *ier = 4;
  if (*ier || *n == 1) return 0;
  const int nn = *n, ii = *inc, jj = *jump;
  bool ok = run_on_array(c, 16, *lot, jj, nn, ii, [&](void *dev, long long l) { return run_c2c(nn, l, ii, jj, dir, dev); });
  if (!ok) *ier = -1;
  return 0;
}

/* ---- cfft2f_/cfft2b_ on a HOST array over several GPUs of this process (cfb200_set_devices / CFB200_DEVICES): column
 * slabs, one per device, copied in concurrently; the two dimension sweeps are the fused-transpose phases of the sharded
 * transform (run_c2c_2d_sharded_phase: the last pass of each dimension stores into the peers' slabs over NVLink), with
 * event barriers between the phases; slabs copied back.  One host thread drives all devices (every launch is
 * asynchronous).  Needs l = ldim, power-of-two l and m in 2^12..2^20 divisible by the device count, peer access. */
#ifdef CFB_SIM
static bool multi2d_eligible(int, int, int, void *) { return false; }
static bool complex_2d_multi(int, int, void *, int) { return false; }
static void multi2d_release() {}
#else
struct Multi2d {
  int l = 0, m = 0, G = 0;
  cudaStream_t st[16];
  cudaEvent_t ev[16];
  void *C[16], *D[16];
  bool peers_on = false;
};
static Multi2d g_m2;
static std::mutex g_m2_mu;

static void multi2d_free() {
  for (int g = 0; g < g_m2.G; ++g) {
    cudaSetDevice(g);
    cudaFree(g_m2.C[g]);
    cudaFree(g_m2.D[g]);
    cudaStreamDestroy(g_m2.st[g]);
    cudaEventDestroy(g_m2.ev[g]);
  }
  g_m2.G = g_m2.l = g_m2.m = 0;
  cudaSetDevice(0);
}

static bool multi2d_prepare(int l, int m, int G) {
  if (g_m2.l == l && g_m2.m == m && g_m2.G == G) return true;
  if (g_m2.G) multi2d_free();
  if (!g_m2.peers_on) {
    for (int a = 0; a < G; ++a)
      for (int b = 0; b < G; ++b) {
        if (a == b) continue;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) {
          set_error("devices %d and %d cannot access each other's memory", a, b);
          return false;
        }
        cudaSetDevice(a);
        cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_ok(e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
      }
    g_m2.peers_on = true;
  }
  const size_t slab = (size_t)l * m / G * 16;
  bool ok = true;
  for (int g = 0; g < G && ok; ++g) {
    cudaSetDevice(g);
    ok = cuda_ok(cudaMalloc(&g_m2.C[g], slab), "cudaMalloc(slab)") && cuda_ok(cudaMalloc(&g_m2.D[g], slab), "cudaMalloc(slab)") &&
         cuda_ok(cudaStreamCreateWithFlags(&g_m2.st[g], cudaStreamNonBlocking), "cudaStreamCreate") &&
         cuda_ok(cudaEventCreateWithFlags(&g_m2.ev[g], cudaEventDisableTiming), "cudaEventCreate");
    if (ok) g_m2.G = g + 1;
  }
  cudaSetDevice(0);
  if (!ok) {
    multi2d_free();
    return false;
  }
  g_m2.l = l;
  g_m2.m = m;
  return true;
}

static bool multi2d_eligible(int ldim, int l, int m, void *c) {
  const int G = devices_in_use();
  if (G < 2 || G > 16 || (G & (G - 1)) || ldim != l) return false;
  int cur = 0;
  cudaGetDevice(&cur);
  if (cur != 0) return false;
  auto pow2_in_range = [](int v) { return v >= 4096 && v <= (1 << 20) && (v & (v - 1)) == 0; };
  if (!pow2_in_range(l) || !pow2_in_range(m) || l % G || m % G) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, c) != cudaSuccess) {
    cudaGetLastError();
    return true;  // unregistered host memory
  }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeUnregistered;
}

static bool complex_2d_multi(int l, int m, void *c, int dir) {
  std::lock_guard<std::mutex> lk(g_m2_mu);
  const int G = devices_in_use();
  if (!multi2d_prepare(l, m, G)) return false;
  const size_t slab = (size_t)l * m / G * 16;
  cudaStream_t saved = t_stream;
  bool ok = true;
  auto barrier = [&]() {  // every device's stream waits for all the others
    for (int g = 0; g < G && ok; ++g) {
      cudaSetDevice(g);
      ok = cuda_ok(cudaEventRecord(g_m2.ev[g], g_m2.st[g]), "cudaEventRecord");
    }
    for (int g = 0; g < G && ok; ++g) {
      cudaSetDevice(g);
      for (int h = 0; h < G && ok; ++h)
        if (h != g) ok = cuda_ok(cudaStreamWaitEvent(g_m2.st[g], g_m2.ev[h], 0), "cudaStreamWaitEvent");
    }
  };
  for (int g = 0; g < G && ok; ++g) {
    cudaSetDevice(g);
    ok = cuda_ok(cudaMemcpyAsync(g_m2.C[g], (char *)c + (size_t)g * slab, slab, cudaMemcpyHostToDevice, g_m2.st[g]), "cudaMemcpyAsync(H2D slab)");
  }
  for (int phase = 1; phase <= 2 && ok; ++phase) {
    barrier();  // phase 1: all slabs are in place (a peer's D must not be written while ... it is free here); phase 2: all D complete
    for (int g = 0; g < G && ok; ++g) {
      cudaSetDevice(g);
      t_stream = g_m2.st[g];
      ok = run_c2c_2d_sharded_phase(phase, dir, l, m, g, G, phase == 1 ? g_m2.C[g] : g_m2.D[g], phase == 1 ? g_m2.D : g_m2.C);
    }
  }
  barrier();  // all column slabs complete
  for (int g = 0; g < G && ok; ++g) {
    cudaSetDevice(g);
    ok = cuda_ok(cudaMemcpyAsync((char *)c + (size_t)g * slab, g_m2.C[g], slab, cudaMemcpyDeviceToHost, g_m2.st[g]), "cudaMemcpyAsync(D2H slab)");
  }
  for (int g = 0; g < G; ++g) {
    cudaSetDevice(g);
    ok = cuda_ok(cudaStreamSynchronize(g_m2.st[g]), "cudaStreamSynchronize") && ok;
  }
  cudaSetDevice(0);
  t_stream = saved;
  return ok;
}
static void multi2d_release() {
  std::lock_guard<std::mutex> lk(g_m2_mu);
  if (g_m2.G) multi2d_free();
}
#endif

static int complex_2d(int *ldim, int *l, int *m, void *c, int *lensav, int *lenwrk, int *ier, int dir) {
  *ier = 0;
  if (*l < 1 || *m < 1) return 0;
  if (*l > *ldim) *ier = 5;
  else if (*lensav < 2 * *l + log2_floor_ref(*l) + 2 * *m + log2_floor_ref(*m) + 8) *ier = 2;
  else if ((long long)*lenwrk < 2LL * *l * *m) *ier = 3;
  if (*ier) return 0;
  if (device_ready() && multi2d_eligible(*ldim, *l, *m, c)) {
    if (!complex_2d_multi(*l, *m, c, dir)) *ier = -1;
    return 0;
  }
  DeviceView v;
  bool ok = view_open(c, ((size_t)*ldim * (*m - 1) + *l) * 16, v);
  if (ok) ok = run_c2c_2d(*ldim, *l, *m, dir, v.dev);
  ok = view_close(v, ok);
  if (!ok) *ier = -1;
  return 0;
}

/* rfft2 (fftpack.c:13113-13508): wsave = [rfft plan of l | cfft plan of m | rfft plan of m] */
static void real_2d_sizes(int l, int m, int &lw, int &mw, int &mm) {
  lw = l + log2_floor_ref(l) + 4;
  mw = 2 * m + log2_floor_ref(m) + 4;
  mm = m + log2_floor_ref(m) + 4;
}
static int real_2d(int *ldim, int *l, int *m, double *r, int *lensav, int *lenwrk, int *ier, int dir) {
  int lw, mw, mm;
  *ier = 0;
  if (*l < 1 || *m < 1) return 0;
  real_2d_sizes(*l, *m, lw, mw, mm);
  if (*lensav < lw + mw + mm) *ier = 2;
  else if ((long long)*lenwrk < ((long long)*l + 1) * *m) *ier = 3;
  else if (*ldim < *l) *ier = 5;
  if (*ier) return 0;
  DeviceView v;
  bool ok = view_open(r, ((size_t)*ldim * (*m - 1) + *l) * 8, v);
  if (ok) ok = run_real_2d(*ldim, *l, *m, dir, (double *)v.dev);
  ok = view_close(v, ok);
  if (!ok) *ier = -1;
  return 0;
}

/* ---------------- real families ---------------- */
static int fam_lensav(int kind, int n) {
  switch (kind) {
    case K_RFFT: return n + log2_floor_ref(n) + 4;
    case K_SINT: return n / 2 + n + log2_floor_ref(n) + 4;
    default: return 2 * n + log2_floor_ref(n) + 4;
  }
}
static long long fam_lenwrk(int kind, int n, long long lot, bool multi) {
  switch (kind) {
    case K_RFFT: return multi ? lot * n : n;
    case K_COST: return multi ? lot * (n + 1) : n - 1;
    case K_SINT: return multi ? lot * (2LL * n + 4) : 2 * n + 2;
    default: return multi ? lot * n : n;
  }
}

static int real_init(int kind, int *n, double *wsave, int *lensav, int *ier) {
  *ier = 0;
  if (*n < 1) return 0;
  if (*lensav < fam_lensav(kind, *n)) {
    *ier = 2;
    return 0;
  }
  switch (kind) {
    case K_RFFT:
      if (*n > 1) wsave_init_real(*n, wsave);
      break;
    case K_COST: wsave_init_cost(*n, wsave); break;
    case K_SINT: wsave_init_sint(*n, wsave); break;
    default: wsave_init_cosq(*n, wsave); break;
  }
  return 0;
}

static int real_1d(int kind, int *n, int *inc, double *x, int *lenx, int *lensav, int *lenwrk, int *ier, int dir) {
  *ier = 0;
  if (*n < 1) return 0;
  if (*inc < 1) {
    *ier = 1;
    return 0;
  }
  if (*lenx < span1(*n, *inc)) *ier = 1;
  else if (*lensav < fam_lensav(kind, *n)) *ier = 2;
  else if (*lenwrk < fam_lenwrk(kind, *n, 1, false)) *ier = 3;
  // sinq1b_ falls through its checks into cosq1b_ and reports that failure as 20 (fftpack.c:14151-14179)
  if (*ier && kind == K_SINQ && dir > 0 && *n > 1) *ier = 20;
  if (*ier || *n == 1) return 0;
  DeviceView v;
  bool ok = view_open(x, (size_t)span1(*n, *inc) * 8, v);
  if (ok) ok = run_real(kind, *n, 1, *inc, (long long)*inc * *n, dir, (double *)v.dev);
  ok = view_close(v, ok);
  if (!ok) *ier = -1;
  return 0;
}

static int real_multi(int kind, int *lot, int *jump, int *n, int *inc, double *x, int *lenx, int *lensav, int *lenwrk,
                      int *ier, int dir) {
  *ier = 0;
  if (*n < 1 || *lot < 1) return 0;
  if (*inc < 1 || *jump < 0) {
    *ier = 4;
    return 0;
  }
  if (*lenx < spanm(*lot, *jump, *n, *inc)) *ier = 1;
  else if (*lensav < fam_lensav(kind, *n)) *ier = 2;
  else if ((long long)*lenwrk < fam_lenwrk(kind, *n, *lot, true)) *ier = 3;
  else if (!strides_consistent(*inc, *jump, *n, *lot)) *ier = 4;
  if (*ier && kind == K_SINQ && dir > 0 && *n > 1) *ier = 20;  // sinqmb_, same fall-through
  if (*ier || *n == 1) return 0;
  const int nn = *n, ii = *inc, jj = *jump;
  bool ok = run_on_array(x, 8, *lot, jj, nn, ii,
                         [&](void *dev, long long l) { return run_real(kind, nn, l, ii, jj, dir, (double *)dev); });
  if (!ok) *ier = -1;
  return 0;
}

}  // namespace cfb

using namespace cfb;

#pragma GCC visibility push(default)
extern "C" {

int cfft1i_(int *n, double *wsave, int *lensav, int *ier) { return complex_init(n, wsave, lensav, ier); }
int cfftmi_(int *n, double *wsave, int *lensav, int *ier) { return complex_init(n, wsave, lensav, ier); }
int cfft1f_(int *n, int *inc, fft_complex_t *c, int *lenc, double *, int *lensav, double *, int *lenwrk, int *ier) {
  return complex_1d(n, inc, c, lenc, lensav, lenwrk, ier, -1);
}
int cfft1b_(int *n, int *inc, fft_complex_t *c, int *lenc, double *, int *lensav, double *, int *lenwrk, int *ier) {
  return complex_1d(n, inc, c, lenc, lensav, lenwrk, ier, +1);
}
int cfftmf_(int *lot, int *jump, int *n, int *inc, fft_complex_t *c, int *lenc, double *, int *lensav, double *,
            int *lenwrk, int *ier) {
  return complex_multi(lot, jump, n, inc, c, lenc, lensav, lenwrk, ier, -1);
}
int cfftmb_(int *lot, int *jump, int *n, int *inc, fft_complex_t *c, int *lenc, double *, int *lensav, double *,
            int *lenwrk, int *ier) {
  return complex_multi(lot, jump, n, inc, c, lenc, lensav, lenwrk, ier, +1);
}

int cfft2i_(int *l, int *m, double *wsave, int *lensav, int *ier) {
  *ier = 0;
  if (*lensav < 2 * *l + log2_floor_ref(*l) + 2 * *m + log2_floor_ref(*m) + 8) {
    *ier = 2;
    return 0;
  }
  int ier1 = 0, ls = 2 * *l + log2_floor_ref(*l) + 4;
  complex_init(l, wsave, &ls, &ier1);
  if (ier1) {
    *ier = 20;
    return 0;
  }
  ls = 2 * *m + log2_floor_ref(*m) + 4;
  complex_init(m, wsave + 2 * *l + log2_floor_ref(*l) + 2, &ls, &ier1);
  if (ier1) *ier = 20;
  return 0;
}
int cfft2f_(int *ldim, int *l, int *m, fft_complex_t *c, double *, int *lensav, double *, int *lenwrk, int *ier) {
  return complex_2d(ldim, l, m, c, lensav, lenwrk, ier, -1);
}
int cfft2b_(int *ldim, int *l, int *m, fft_complex_t *c, double *, int *lensav, double *, int *lenwrk, int *ier) {
  return complex_2d(ldim, l, m, c, lensav, lenwrk, ier, +1);
}

int rfft2i_(int *l, int *m, double *wsave, int *lensav, int *ier) {
  int lw, mw, mm, ier1 = 0;
  *ier = 0;
  real_2d_sizes(*l, *m, lw, mw, mm);
  if (*lensav < lw + mw + mm) {
    *ier = 2;
    return 0;
  }
  real_init(K_RFFT, l, wsave, &lw, &ier1);
  if (!ier1) complex_init(m, wsave + lw, &mw, &ier1);
  if (!ier1) real_init(K_RFFT, m, wsave + lw + mw, &mm, &ier1);
  if (ier1) *ier = 20;
  return 0;
}
int rfft2f_(int *ldim, int *l, int *m, double *r, double *, int *lensav, double *, int *lenwrk, int *ier) {
  return real_2d(ldim, l, m, r, lensav, lenwrk, ier, -1);
}
int rfft2b_(int *ldim, int *l, int *m, double *r, double *, int *lensav, double *, int *lenwrk, int *ier) {
  return real_2d(ldim, l, m, r, lensav, lenwrk, ier, +1);
}

#define CFB_DEF_REAL(name, K)                                                                                          \
  int name##1i_(int *n, double *wsave, int *lensav, int *ier) { return real_init(K, n, wsave, lensav, ier); }          \
  int name##mi_(int *n, double *wsave, int *lensav, int *ier) { return real_init(K, n, wsave, lensav, ier); }          \
  int name##1f_(int *n, int *inc, double *x, int *lenx, double *, int *lensav, double *, int *lenwrk, int *ier) {      \
    return real_1d(K, n, inc, x, lenx, lensav, lenwrk, ier, -1);                                                       \
  }                                                                                                                    \
  int name##1b_(int *n, int *inc, double *x, int *lenx, double *, int *lensav, double *, int *lenwrk, int *ier) {      \
    return real_1d(K, n, inc, x, lenx, lensav, lenwrk, ier, +1);                                                       \
  }                                                                                                                    \
  int name##mf_(int *lot, int *jump, int *n, int *inc, double *x, int *lenx, double *, int *lensav, double *,          \
                int *lenwrk, int *ier) {                                                                               \
    return real_multi(K, lot, jump, n, inc, x, lenx, lensav, lenwrk, ier, -1);                                         \
  }                                                                                                                    \
  int name##mb_(int *lot, int *jump, int *n, int *inc, double *x, int *lenx, double *, int *lensav, double *,          \
                int *lenwrk, int *ier) {                                                                               \
    return real_multi(K, lot, jump, n, inc, x, lenx, lensav, lenwrk, ier, +1);                                         \
  }
CFB_DEF_REAL(rfft, K_RFFT)
CFB_DEF_REAL(cost, K_COST)
CFB_DEF_REAL(sint, K_SINT)
CFB_DEF_REAL(cosq, K_COSQ)
CFB_DEF_REAL(sinq, K_SINQ)

/* test/vargamma.c:42-106 for `lot` options at once; see include/cfftpack_b200.h */
int cfb200_option_convolution(int lot, int n, const double *S, const double *K, const double *sigma, const double *theta,
                              const double *kappa, const double *t, const double *r, const int *flags, double *value,
                              int *ier) {
  *ier = 0;
  if (lot <= 0 || n <= 0 || !S || !K || !sigma || !theta || !kappa || !t || !r || !flags || !value) {
    *ier = 1;
    return 0;
  }
  const int N = next_fast_even_size(n);
  if (!device_ready()) {
    *ier = -1;
    return N;
  }
  std::vector<double> par((size_t)8 * lot);
  const double *cols[7] = {S, K, sigma, theta, kappa, t, r};
  for (int k = 0; k < 7; ++k) memcpy(&par[(size_t)k * lot], cols[k], (size_t)lot * sizeof(double));
  for (int o = 0; o < lot; ++o) par[(size_t)7 * lot + o] = (double)(flags[o] & 3);
  if (!run_option_convolution(lot, N, par.data(), value)) *ier = -1;
  return N;
}

int cfb200_cfft2_sharded_phase(int phase, int direction, int l, int m, int rank, int nranks, void *local_src,
                               void *const *peer_dst, int *ier) {
  *ier = 0;
  if (phase < 1 || phase > 2 || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks || l < 1 || m < 1) {
    *ier = 1;
    return 0;
  }
  if (!device_ready() || !run_c2c_2d_sharded_phase(phase, direction < 0 ? -1 : +1, l, m, rank, nranks, local_src, peer_dst))
    *ier = -1;
  return 0;
}

int cfb200_cfft1_sharded_phase(int phase, int direction, int log2n, int rank, int nranks, void *local_src,
                               void *const *peer_dst, int *ier) {
  *ier = 0;
  if (phase < 0 || phase > 2 || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks || !local_src || !peer_dst) {
    *ier = 1;
    return 0;
  }
  if (!device_ready() || !run_c2c_1d_sharded_phase(phase, direction < 0 ? -1 : +1, log2n, rank, nranks, local_src, peer_dst)) *ier = -1;
  return 0;
}

void *cfb200_host_alloc(size_t bytes) {
  void *p = nullptr;
  if (!device_ready() || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return p;
}
void cfb200_host_free(void *p) {
  if (p) cudaFreeHost(p);
}

int cfb200_set_devices(int n) {
  int have = 0;
  if (!device_ready() || cudaGetDeviceCount(&have) != cudaSuccess || have < 1) return 0;
  g_devices.store(n <= 0 || n > have ? have : n);
  return g_devices.load();
}

int cfb200_set_stream(void *s) {
  set_current_stream((cudaStream_t)s);
  return 0;
}
int cfb200_synchronize(void) {
  if (!device_ready()) return -1;
  return cuda_ok(cudaStreamSynchronize(current_stream()), "cudaStreamSynchronize") ? 0 : -1;
}
unsigned long long cfb200_launch_count(void) { return launch_count(); }
const char *cfb200_last_error(void) { return last_error(); }
void cfb200_release(void) {
  multi2d_release();
  release_plans();
  scratch_release_all();
  if (t_bounce.p) {
    cudaFreeHost(t_bounce.p);
    t_bounce.p = nullptr;
    t_bounce.tried = false;
  }
}
const char *cfb200_version(void) {
#ifdef CFB_SIM
  return "cfftpack_b200 0.2 SIM (CPU thread emulator, tests only)";
#else
  return "cfftpack_b200 0.2 sm_100a";
#endif
}
int cfb200_max_onchip_complex(void) { return engine_max_c2c(); }
int cfb200_max_onchip_real(void) { return engine_max_real(); }
}
#pragma GCC visibility pop
