/* mixed_impl.cuh -- host side of one length of the mixed-radix streaming kernel (included by mixed_<M>.cu). */
#ifndef CFB_MIXED_IMPL_CUH
#define CFB_MIXED_IMPL_CUH
#include <map>
#include <mutex>
#include <vector>

#include "mixed.cuh"
#include "plan.h"

namespace cfb {

constexpr int MIX_THREADS = 96;

template <class C>
struct MixTables {
  static std::mutex &mu() {
    static std::mutex m;
    return m;
  }
  static std::map<int, cpx *> &tabs() {
    static std::map<int, cpx *> t;
    return t;
  }
  static const cpx *get() {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu());
    auto it = tabs().find(dev);
    if (it != tabs().end()) return it->second;
    std::vector<cpx> h(C::TAB);
    auto root = [&](int at, long long num, long long den) { unit_root(num, den, &h[at].x, &h[at].y); };
    for (int p = 0; p < C::NB0; ++p) {
      root(C::T0 + p, p, C::M);
      if (C::W4_0) root(C::T0 + C::NB0 + p, 4LL * p, C::M);
    }
    if (C::R1 > 1)
      for (int p = 0; p < C::MM1; ++p) {
        root(C::T1 + p, p, (long long)C::MM1 * C::R1);
        if (C::W4_1) root(C::T1 + C::MM1 + p, 4LL * p, (long long)C::MM1 * C::R1);
      }
    for (int j = 0; j < C::R0; ++j) root(C::RT0 + j, j, C::R0);
    for (int j = 0; j < C::R1; ++j) root(C::RT1 + j, j, C::R1);
    for (int j = 0; j < C::R2; ++j) root(C::RT2 + j, j, C::R2);
    cpx *d = (cpx *)upload_table(h.data(), h.size() * sizeof(cpx));
    if (!d) return nullptr;
    tabs()[dev] = d;
    return d;
  }
  static void release() {
    std::lock_guard<std::mutex> lk(mu());
    for (auto &kv : tabs()) cudaFree(kv.second);
    tabs().clear();
  }
};

template <class C, int KIND, int DIR>
bool mix_launch_kd(long long npairs, double *x, const double *trig) {
  const cpx *tab = MixTables<C>::get();
  if (!tab) return false;
  auto kern = mix_stream_kernel<C, KIND, DIR, MIX_THREADS>;
  if (!kernel_attrs_ready((const void *)kern, C::BYTES)) return false;
  long long per_sm = (long long)((227 * 1024 + 1024) / (C::BYTES + 1024));
  if (per_sm > 4) per_sm = 4;  // register budget of the launch bounds
  long long grid = per_sm * sm_count();
  if (grid > npairs) grid = npairs;
  CFB_LAUNCH(kern, (unsigned)grid, MIX_THREADS, C::BYTES, current_stream(), x, npairs, tab, trig);
  count_launch();
  return cuda_ok(cudaGetLastError(), "mix_stream_kernel launch");
}

template <class C>
bool mix_launch_cfg(int kind, int dir, long long npairs, double *x, const double *trig) {
  switch (kind) {
    case K_RFFT: return dir < 0 ? mix_launch_kd<C, K_RFFT, -1>(npairs, x, trig) : mix_launch_kd<C, K_RFFT, 1>(npairs, x, trig);
    case K_COSQ: return dir < 0 ? mix_launch_kd<C, K_COSQ, -1>(npairs, x, trig) : mix_launch_kd<C, K_COSQ, 1>(npairs, x, trig);
    case K_SINT: return dir < 0 ? mix_launch_kd<C, K_SINT, -1>(npairs, x, trig) : mix_launch_kd<C, K_SINT, 1>(npairs, x, trig);
    case K_COST: return dir < 0 ? mix_launch_kd<C, K_COST, -1>(npairs, x, trig) : mix_launch_kd<C, K_COST, 1>(npairs, x, trig);
    default: break;
  }
  set_error("mixed-radix kernel: unknown family %d", kind);
  return false;
}

}  // namespace cfb
#endif
