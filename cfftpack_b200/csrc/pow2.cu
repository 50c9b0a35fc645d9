/* pow2.cu -- host side of the power-of-two register kernels: twiddle tables, attributes, launches. */
#include "pow2.cuh"

#include <map>
#include <vector>

#include "plan.h"

namespace cfb {

namespace {
std::mutex g_mu;
std::map<std::pair<int, int>, cpx *> g_tw;  // (device, log2n) -> table

template <int LOG2N>
const cpx *pow2_table() {
  typedef Pow2Cfg<LOG2N> C;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_pair(dev, LOG2N);
  auto it = g_tw.find(key);
  if (it != g_tw.end()) return it->second;
  std::vector<cpx> h((size_t)C::TW_COUNT + 1);
  size_t o = 0;
  for (int st = 0; st < C::NFULL; ++st) {
    const int m = C::N >> (C::LP * (st + 1));
    const long long ncur = (long long)m * C::P;
    if (st == C::NFULL - 1 && C::REM == 0) break;
    for (int k = 1; k < C::P; ++k)
      for (int p = 0; p < m; ++p) {
        unit_root((long long)p * k, ncur, &h[o].x, &h[o].y);
        ++o;
      }
  }
  cpx *d = nullptr;
  if (!cuda_ok(cudaMalloc((void **)&d, h.size() * sizeof(cpx)), "cudaMalloc(pow2 twiddles)")) return nullptr;
  if (!cuda_ok(cudaMemcpy(d, h.data(), h.size() * sizeof(cpx), cudaMemcpyHostToDevice), "cudaMemcpy(pow2 twiddles)")) {
    cudaFree(d);
    return nullptr;
  }
  g_tw[key] = d;
  return d;
}

template <class K>
bool set_smem_once(K kernel, size_t smem, std::once_flag &once, bool &ok) {
  std::call_once(once, [&] {
    ok = smem <= 48 * 1024 ||
         cuda_ok(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                 "cudaFuncSetAttribute(pow2 kernel)");
  });
  return ok;
}

template <int LOG2N, int DIR>
bool launch_c2c(long long lot, long long jump, cpx *c) {
  typedef Pow2Cfg<LOG2N> C;
  const cpx *tw = pow2_table<LOG2N>();
  if (!tw) return false;
  static std::once_flag once;
  static bool ok = true;
  auto kern = pow2_c2c_kernel<LOG2N, DIR>;
  if (!set_smem_once(kern, C::SMEM, once, ok)) return false;
  const long long grid = (lot + C::TPB - 1) / C::TPB;
  const double scale = DIR < 0 ? 1.0 / (double)C::N : 1.0;
  CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, C::SMEM, current_stream(), c, lot, jump, tw, scale);
  count_launch();
  return cuda_ok(cudaGetLastError(), "pow2_c2c_kernel launch");
}

template <int LOG2N, int DIR>
bool launch_r2c(long long lot, long long jump, double *r) {
  typedef Pow2Cfg<LOG2N> C;
  const cpx *tw = pow2_table<LOG2N>();
  if (!tw) return false;
  static std::once_flag once;
  static bool ok = true;
  auto kern = pow2_r2c_kernel<LOG2N, DIR>;
  if (!set_smem_once(kern, C::SMEM, once, ok)) return false;
  const long long pairs = (lot + 1) / 2;
  const long long grid = (pairs + C::TPB - 1) / C::TPB;
  CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, C::SMEM, current_stream(), r, lot, jump, tw);
  count_launch();
  return cuda_ok(cudaGetLastError(), "pow2_r2c_kernel launch");
}

int ilog2_exact(int n) {
  int l = 0;
  while ((1 << l) < n) ++l;
  return (1 << l) == n ? l : -1;
}
}  // namespace

static const int POW2_MIN_LOG = 6, POW2_MAX_LOG = 13;

bool pow2_c2c_supported(int n, long long inc, long long jump, int aligned16) {
  int l = ilog2_exact(n);
  return l >= POW2_MIN_LOG && l <= POW2_MAX_LOG && inc == 1 && jump >= n && aligned16;
}
bool pow2_r2c_supported(int n, long long inc, long long jump, int) {
  int l = ilog2_exact(n);
  return l >= POW2_MIN_LOG && l <= POW2_MAX_LOG && inc == 1 && jump >= n;
}

#define CFB_POW2_CASES(FN, ...)                                       \
  switch (ilog2_exact(n)) {                                           \
    case 6: return dir < 0 ? FN<6, -1>(__VA_ARGS__) : FN<6, 1>(__VA_ARGS__);    \
    case 7: return dir < 0 ? FN<7, -1>(__VA_ARGS__) : FN<7, 1>(__VA_ARGS__);    \
    case 8: return dir < 0 ? FN<8, -1>(__VA_ARGS__) : FN<8, 1>(__VA_ARGS__);    \
    case 9: return dir < 0 ? FN<9, -1>(__VA_ARGS__) : FN<9, 1>(__VA_ARGS__);    \
    case 10: return dir < 0 ? FN<10, -1>(__VA_ARGS__) : FN<10, 1>(__VA_ARGS__); \
    case 11: return dir < 0 ? FN<11, -1>(__VA_ARGS__) : FN<11, 1>(__VA_ARGS__); \
    case 12: return dir < 0 ? FN<12, -1>(__VA_ARGS__) : FN<12, 1>(__VA_ARGS__); \
    case 13: return dir < 0 ? FN<13, -1>(__VA_ARGS__) : FN<13, 1>(__VA_ARGS__); \
    default: break;                                                   \
  }

bool pow2_c2c_launch(int n, long long lot, long long jump, int dir, cpx *c) {
  CFB_POW2_CASES(launch_c2c, lot, jump, c)
  set_error("pow2_c2c_launch: unsupported length %d", n);
  return false;
}
bool pow2_r2c_launch(int n, long long lot, long long jump, int dir, double *r) {
  CFB_POW2_CASES(launch_r2c, lot, jump, r)
  set_error("pow2_r2c_launch: unsupported length %d", n);
  return false;
}

}  // namespace cfb
