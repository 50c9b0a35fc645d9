/* radix10.cu -- host side of the N = 10^K register kernels: twiddle rows, attributes, launches. */
#include "radix10.cuh"

#include <map>
#include <mutex>
#include <vector>

#include "plan.h"

namespace cfb {

namespace {
const size_t SMEM_LIMIT = 227 * 1024;
std::mutex g_mu;
std::map<std::pair<int, int>, cpx *> g_tw;  // (device, K) -> rows w^p, w^4p of every non-last stage

template <int K>
const cpx *r10_table() {
  typedef R10Cfg<K> C;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_pair(dev, K);
  auto it = g_tw.find(key);
  if (it != g_tw.end()) return it->second;
  std::vector<cpx> h((size_t)C::TWS_COUNT + 1);
  size_t o = 0;
  for (int st = 0; st < K - 1; ++st) {
    const int m = C::stage_m(st);
    const long long ncur = (long long)m * C::P;
    for (int e = 1; e <= 4; e += 3)
      for (int p = 0; p < m; ++p) {
        unit_root((long long)p * e, ncur, &h[o].x, &h[o].y);
        ++o;
      }
  }
  cpx *d = (cpx *)upload_table(h.data(), h.size() * sizeof(cpx));
  if (!d) return nullptr;
  g_tw[key] = d;
  return d;
}

template <class Kern>
bool attr_once(Kern kern, size_t smem) {
  return kernel_attrs_ready((const void *)kern, smem);
}

template <int K>
long long grid_for(long long ntiles) {
  typedef R10Cfg<K> C;
  long long per_sm = (long long)((SMEM_LIMIT + 1024) / (C::BYTES + 1024));
  const long long reg_cap = C::THREADS <= 128 ? 5 : 3;
  if (per_sm > reg_cap) per_sm = reg_cap;
  if (per_sm < 1) per_sm = 1;
  const long long cap = per_sm * sm_count();
  return ntiles < cap ? ntiles : cap;
}

template <int K, int DIR>
bool launch_c2c(long long lot, long long jump, cpx *c, double scale) {
  typedef R10Cfg<K> C;
  const cpx *tw = r10_table<K>();
  if (!tw) return false;
  auto kern = r10_c2c_stream_kernel<K, DIR>;
  if (!attr_once(kern, C::BYTES)) return false;
  const long long ntiles = (lot + C::TPB - 1) / C::TPB;
  CFB_LAUNCH(kern, (unsigned)grid_for<K>(ntiles), C::THREADS, C::BYTES, current_stream(), c, lot, jump, tw, scale, ntiles);
  count_launch();
  return cuda_ok(cudaGetLastError(), "r10_c2c_stream_kernel launch");
}

template <int K, int KIND, int DIR>
bool launch_r2c(long long lot, long long jump, double *r, const double *trig) {
  typedef R10Cfg<K> C;
  const cpx *tw = r10_table<K>();
  if (!tw) return false;
  auto kern = r10_r2c_stream_kernel<K, KIND, DIR>;
  const size_t smem = KIND == K_COSQ ? C::BYTES_TRIG : C::BYTES;
  if (!attr_once(kern, smem)) return false;
  const long long pairs = (lot + 1) / 2;
  const long long ntiles = (pairs + C::TPB - 1) / C::TPB;
  CFB_LAUNCH(kern, (unsigned)grid_for<K>(ntiles), C::THREADS, smem, current_stream(), r, lot, jump, tw, trig, ntiles);
  count_launch();
  return cuda_ok(cudaGetLastError(), "r10_r2c_stream_kernel launch");
}
}  // namespace

void r10_release_tables() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto &kv : g_tw) cudaFree(kv.second);
  g_tw.clear();
}

bool r10_cost_launch(long long npairs, int dir, double *x, const double *trig) {
  typedef R10Cfg<3> C;
  const cpx *tw = r10_table<3>();
  if (!tw) return false;
  const long long ntiles = (npairs + C::TPB - 1) / C::TPB;
  long long per_sm = (long long)((SMEM_LIMIT + 1024) / (R10Cost::BYTES + 1024));
  if (per_sm > C::MINB) per_sm = C::MINB;
  long long grid = per_sm * sm_count();
  if (grid > ntiles) grid = ntiles;
  if (dir < 0) {
    auto kern = r10_cost_stream_kernel<-1>;
    if (!attr_once(kern, R10Cost::BYTES)) return false;
    CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, R10Cost::BYTES, current_stream(), x, npairs, tw, trig, ntiles);
  } else {
    auto kern = r10_cost_stream_kernel<1>;
    if (!attr_once(kern, R10Cost::BYTES)) return false;
    CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, R10Cost::BYTES, current_stream(), x, npairs, tw, trig, ntiles);
  }
  count_launch();
  return cuda_ok(cudaGetLastError(), "r10_cost_stream_kernel launch");
}

bool r10_supported(int n) { return n == 100 || n == 1000; }

bool r10_c2c_launch(int n, long long lot, long long jump, int dir, cpx *c, double scale) {
  if (n == 100) return dir < 0 ? launch_c2c<2, -1>(lot, jump, c, scale) : launch_c2c<2, 1>(lot, jump, c, scale);
  if (n == 1000) return dir < 0 ? launch_c2c<3, -1>(lot, jump, c, scale) : launch_c2c<3, 1>(lot, jump, c, scale);
  set_error("r10_c2c_launch: unsupported length %d", n);
  return false;
}
bool r10_r2c_launch(int n, long long lot, long long jump, int dir, double *r) {
  if (n == 100) return dir < 0 ? launch_r2c<2, K_RFFT, -1>(lot, jump, r, nullptr) : launch_r2c<2, K_RFFT, 1>(lot, jump, r, nullptr);
  if (n == 1000) return dir < 0 ? launch_r2c<3, K_RFFT, -1>(lot, jump, r, nullptr) : launch_r2c<3, K_RFFT, 1>(lot, jump, r, nullptr);
  set_error("r10_r2c_launch: unsupported length %d", n);
  return false;
}
bool r10_cosq_launch(int n, long long lot, long long jump, int dir, double *x, const double *trig) {
  if (n == 100) return dir < 0 ? launch_r2c<2, K_COSQ, -1>(lot, jump, x, trig) : launch_r2c<2, K_COSQ, 1>(lot, jump, x, trig);
  if (n == 1000) return dir < 0 ? launch_r2c<3, K_COSQ, -1>(lot, jump, x, trig) : launch_r2c<3, K_COSQ, 1>(lot, jump, x, trig);
  set_error("r10_cosq_launch: unsupported length %d", n);
  return false;
}

}  // namespace cfb
