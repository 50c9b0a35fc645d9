/* tools/sim/cuda_sim.cpp -- fiber scheduler behind cuda_sim.h.  TEST INFRASTRUCTURE ONLY. */
#include "cuda_sim.h"

#include <map>
#include <mutex>

namespace cfbsim {
Ctx *cur = nullptr;
char *dyn_smem = nullptr;

namespace {
struct Fiber {
  ucontext_t uc;
  Ctx ctx;
  char *stack = nullptr;
  bool done = false;
  unsigned long bar_gen = 0;   // block barriers passed
  unsigned long wbar_gen = 0;  // warp barriers passed
};
ucontext_t sched_uc;
std::vector<Fiber> fibers;
int running = -1;
const std::function<void()> *body_fn = nullptr;
double shfl_slot[2048];
const size_t STACK = 256 * 1024;

void trampoline() {
  (*body_fn)();
  fibers[running].done = true;
  swapcontext(&fibers[running].uc, &sched_uc);
}
void yield() { swapcontext(&fibers[running].uc, &sched_uc); }
}  // namespace

void yield_once() { yield(); }

void yield_barrier() {
  Fiber &f = fibers[running];
  f.bar_gen++;
  // wait until every live fiber of the block has reached the same generation
  for (;;) {
    bool ok = true;
    for (auto &g : fibers)
      if (!g.done && g.bar_gen < f.bar_gen) {
        ok = false;
        break;
      }
    if (ok) break;
    yield();
  }
}

void warp_barrier() {
  int me = running;
  Fiber &f = fibers[me];
  f.wbar_gen++;
  int w0 = me / 32 * 32, w1 = w0 + 32;
  if (w1 > (int)fibers.size()) w1 = (int)fibers.size();
  for (;;) {
    bool ok = true;
    for (int i = w0; i < w1; ++i)
      if (!fibers[i].done && fibers[i].wbar_gen < f.wbar_gen) {
        ok = false;
        break;
      }
    if (ok) break;
    yield();
  }
}

double shfl(double v, int src_lane) {
  int me = running, w0 = me / 32 * 32;
  shfl_slot[me] = v;
  warp_barrier();
  double r = shfl_slot[w0 + (src_lane & 31)];
  warp_barrier();
  return r;
}

void run(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
  static std::mutex mu;  // one simulated device
  std::lock_guard<std::mutex> lk(mu);
  int nthreads = (int)(block.x * block.y * block.z);
  if (nthreads > 2048) {
    fprintf(stderr, "cfbsim: block too large\n");
    abort();
  }
  std::vector<char> smem_buf(smem + 64);
  dyn_smem = (char *)(((uintptr_t)smem_buf.data() + 15) & ~(uintptr_t)15);
  body_fn = &body;
  fibers.assign(nthreads, Fiber());
  for (auto &f : fibers) f.stack = (char *)malloc(STACK);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        memset(dyn_smem, 0xFF, smem);  // poison (all-ones = NaN for doubles): catches reads of unwritten shared memory
        for (int t = 0; t < nthreads; ++t) {
          Fiber &f = fibers[t];
          f.done = false;
          f.bar_gen = f.wbar_gen = 0;
          f.ctx.bidx = dim3(bx, by, bz);
          f.ctx.gdim = grid;
          f.ctx.bdim = block;
          f.ctx.tidx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
          getcontext(&f.uc);
          f.uc.uc_stack.ss_sp = f.stack;
          f.uc.uc_stack.ss_size = STACK;
          f.uc.uc_link = &sched_uc;
          makecontext(&f.uc, trampoline, 0);
        }
        int live = nthreads;
        while (live > 0) {
          live = 0;
          for (int t = 0; t < nthreads; ++t) {
            if (fibers[t].done) continue;
            running = t;
            cur = &fibers[t].ctx;
            swapcontext(&sched_uc, &fibers[t].uc);
            if (!fibers[t].done) ++live;
          }
        }
      }
  for (auto &f : fibers) free(f.stack);
  fibers.clear();
  running = -1;
  cur = nullptr;
}
}  // namespace cfbsim

static std::map<uintptr_t, size_t> g_dev;
static std::map<uintptr_t, size_t> g_pinned;
static std::mutex g_dev_mu;
void cfbsim_mark_device(const void *p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (bytes == 0) g_dev.erase((uintptr_t)p);
  else g_dev[(uintptr_t)p] = bytes;
}
int cfbsim_is_device(const void *p) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  auto it = g_dev.upper_bound((uintptr_t)p);
  if (it == g_dev.begin()) return 0;
  --it;
  return (uintptr_t)p < it->first + it->second;
}
int cfbsim_is_pinned(const void *p) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  auto it = g_pinned.upper_bound((uintptr_t)p);
  if (it == g_pinned.begin()) return 0;
  --it;
  return (uintptr_t)p < it->first + it->second;
}
extern "C" void cfb200_sim_mark_device(const void *p, size_t bytes) { cfbsim_mark_device(p, bytes); }
extern "C" void cfb200_sim_mark_pinned(const void *p, size_t bytes) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (bytes == 0) g_pinned.erase((uintptr_t)p);
  else g_pinned[(uintptr_t)p] = bytes;
}
