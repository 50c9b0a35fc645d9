/* mixed3.cu -- host side of the 13*11*7 streaming kernel: twiddle rows, attributes, launches. */
#include "mixed3.cuh"

#include <stdlib.h>

#include <map>
#include <mutex>
#include <vector>

#include "plan.h"

namespace cfb {

namespace {
const size_t SMEM_LIMIT = 227 * 1024;
typedef M3Cfg<13, 11, 7> C1001;
const int M3_THREADS = 96;
std::mutex g_mu;
std::map<int, cpx *> g_tab;  // device -> table

template <class C>
const cpx *m3_table() {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_tab.find(dev);
  if (it != g_tab.end()) return it->second;
  std::vector<cpx> h(C::TAB);
  for (int e = 0; e < 2; ++e)
    for (int p = 0; p < C::NB0; ++p) unit_root((long long)p * (e ? 4 : 1), C::M, &h[C::T0 + e * C::NB0 + p].x, &h[C::T0 + e * C::NB0 + p].y);
  for (int e = 0; e < 2; ++e)
    for (int p = 0; p < C::MM1; ++p)
      unit_root((long long)p * (e ? 4 : 1), (long long)C::MM1 * C::R1, &h[C::T1 + e * C::MM1 + p].x, &h[C::T1 + e * C::MM1 + p].y);
  for (int j = 0; j < C::R0; ++j) unit_root(j, C::R0, &h[C::RT0 + j].x, &h[C::RT0 + j].y);
  for (int j = 0; j < C::R1; ++j) unit_root(j, C::R1, &h[C::RT1 + j].x, &h[C::RT1 + j].y);
  for (int j = 0; j < C::R2; ++j) unit_root(j, C::R2, &h[C::RT2 + j].x, &h[C::RT2 + j].y);
  cpx *d = (cpx *)upload_table(h.data(), h.size() * sizeof(cpx));
  if (!d) return nullptr;
  g_tab[dev] = d;
  return d;
}

template <int KIND, int DIR, int THREADS>
bool launch_w(long long npairs, double *x, const double *trig) {
  typedef C1001 C;
  const int M3_THREADS = THREADS;
  const cpx *tab = m3_table<C>();
  if (!tab) return false;
  auto kern = m3_stream_kernel<C, KIND, DIR, THREADS>;
  if (!kernel_attrs_ready((const void *)kern, C::BYTES)) return false;
  long long per_sm = (long long)((SMEM_LIMIT + 1024) / (C::BYTES + 1024));
  if (per_sm > 4) per_sm = 4;  // register budget of the launch bounds
  long long grid = per_sm * sm_count();
  if (grid > npairs) grid = npairs;
  CFB_LAUNCH(kern, (unsigned)grid, M3_THREADS, C::BYTES, current_stream(), x, npairs, tab, trig);
  count_launch();
  return cuda_ok(cudaGetLastError(), "m3_stream_kernel launch");
}
template <int KIND, int DIR>
bool launch(long long npairs, double *x, const double *trig) {
#ifdef CFB_M3_EXPERIMENT
  static const int w = getenv("CFB200_M3_THREADS") ? atoi(getenv("CFB200_M3_THREADS")) : 0;
  if (w == 64) return launch_w<KIND, DIR, 64>(npairs, x, trig);
  if (w == 128) return launch_w<KIND, DIR, 128>(npairs, x, trig);
  if (w == 160) return launch_w<KIND, DIR, 160>(npairs, x, trig);
#endif
  return launch_w<KIND, DIR, M3_THREADS>(npairs, x, trig);
}
}  // namespace

void m3_release_tables() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto &kv : g_tab) cudaFree(kv.second);
  g_tab.clear();
}

bool m3_supported(int kind, int n) {
  return (kind == K_RFFT && n == 1001) || (kind == K_COSQ && n == 1001) || (kind == K_SINT && n == 1000);
}

bool m3_launch(int kind, int n, long long npairs, int dir, double *x, const double *trig) {
  if (npairs <= 0) return true;
  if (!m3_supported(kind, n)) {
    set_error("m3_launch: unsupported (kind %d, n %d)", kind, n);
    return false;
  }
  switch (kind) {
    case K_RFFT: return dir < 0 ? launch<K_RFFT, -1>(npairs, x, trig) : launch<K_RFFT, 1>(npairs, x, trig);
    case K_COSQ: return dir < 0 ? launch<K_COSQ, -1>(npairs, x, trig) : launch<K_COSQ, 1>(npairs, x, trig);
    default: return dir < 0 ? launch<K_SINT, -1>(npairs, x, trig) : launch<K_SINT, 1>(npairs, x, trig);
  }
}

}  // namespace cfb
