/* pow2.cu -- host side of the power-of-two register kernels: twiddle tables, attributes, launches. */
#include "pow2.cuh"

#include <stdlib.h>

#include <map>
#include <tuple>
#include <vector>

#include "plan.h"

namespace cfb {

namespace {
const size_t SMEM_LIMIT = 227 * 1024;
std::mutex g_mu;
std::map<std::tuple<int, int, int, int>, cpx *> g_tw;  // (device, log2n, lp, tw) -> table

template <class C>
const cpx *pow2_table() {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  int l2 = 0;
  while ((1 << l2) < C::N) ++l2;
  auto key = std::make_tuple(dev, l2, C::LP, C::TW);
  auto it = g_tw.find(key);
  if (it != g_tw.end()) return it->second;
  std::vector<cpx> h((size_t)C::TW_COUNT + 1);
  size_t o = 0;
  for (int st = 0; st < C::NFULL; ++st) {
    if (C::stage_last(st)) break;
    const int m = C::stage_m(st);
    const long long ncur = (long long)m * C::P;
    if (C::stage_computed(st)) {
      for (int e = 1; e <= 4; e += 3)  // w^p then w^(4p)
        for (int p = 0; p < m; ++p) {
          unit_root((long long)p * e, ncur, &h[o].x, &h[o].y);
          ++o;
        }
    } else {
      for (int k = 1; k < C::P; ++k)
        for (int p = 0; p < m; ++p) {
          unit_root((long long)p * k, ncur, &h[o].x, &h[o].y);
          ++o;
        }
    }
  }
  cpx *d = (cpx *)upload_table(h.data(), h.size() * sizeof(cpx));
  if (!d) return nullptr;
  g_tw[key] = d;
  return d;
}

/* streaming kernels: per non-last stage the rows w^p and w^(4p) (copied to shared memory by each CTA) */
template <class C>
const cpx *pow2_stream_table() {
  typedef StreamSmem<C> S;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(g_mu);
  int l2 = 0;
  while ((1 << l2) < C::N) ++l2;
  auto key = std::make_tuple(dev, l2, C::LP, 100);
  auto it = g_tw.find(key);
  if (it != g_tw.end()) return it->second;
  std::vector<cpx> h((size_t)S::TWS_COUNT + 1);
  size_t o = 0;
  for (int st = 0; st < C::NFULL; ++st) {
    if (C::stage_last(st)) continue;
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

template <class K>
bool set_smem_once(K kernel, size_t smem) {
  return kernel_attrs_ready((const void *)kernel, smem);
}

template <class C, int MINB, int DIR>
bool launch_r2c_cfg(long long lot, long long jump, double *r) {
  const cpx *tw = pow2_table<C>();
  if (!tw) return false;
  auto kern = pow2_r2c_kernel<C, MINB, DIR>;
  if (!set_smem_once(kern, C::SMEM)) return false;
  const long long pairs = (lot + 1) / 2;
  const long long grid = (pairs + C::TPB - 1) / C::TPB;
  CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, C::SMEM, current_stream(), r, lot, jump, tw);
  count_launch();
  return cuda_ok(cudaGetLastError(), "pow2_r2c_kernel launch");
}

template <class C, int MINB, int DIR>
bool launch_c2c_stream(long long lot, long long jump, cpx *c, double scale) {
  const cpx *tw = pow2_stream_table<C>();
  if (!tw) return false;
  auto kern = pow2_c2c_stream_kernel<C, MINB, DIR>;
  if (!set_smem_once(kern, StreamSmem<C>::BYTES)) return false;
  const long long ntiles = (lot + C::TPB - 1) / C::TPB;
  const long long cap = (long long)MINB * sm_count();
  const long long grid = ntiles < cap ? ntiles : cap;
  CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, StreamSmem<C>::BYTES, current_stream(), c, lot, jump, tw, scale, ntiles);
  count_launch();
  return cuda_ok(cudaGetLastError(), "pow2_c2c_stream_kernel launch");
}

template <class C, int MINB, int DIR, bool BULK>
bool launch_r2c_stream_v(long long lot, long long jump, double *r);
template <class C, int MINB, int DIR>
bool launch_r2c_stream(long long lot, long long jump, double *r) {
  static const int bulk = getenv("CFB200_R2C_BULK") ? atoi(getenv("CFB200_R2C_BULK")) : 1;  // 0: per-thread stores (A/B)
  return bulk ? launch_r2c_stream_v<C, MINB, DIR, true>(lot, jump, r) : launch_r2c_stream_v<C, MINB, DIR, false>(lot, jump, r);
}
template <class C, int MINB, int DIR, bool BULK>
bool launch_r2c_stream_v(long long lot, long long jump, double *r) {
  const cpx *tw = pow2_stream_table<C>();
  if (!tw) return false;
  auto kern = pow2_r2c_stream_kernel<C, MINB, DIR, BULK>;
  if (!set_smem_once(kern, StreamSmem<C>::BYTES)) return false;
  const long long pairs = (lot + 1) / 2;
  const long long ntiles = (pairs + C::TPB - 1) / C::TPB;
  const long long cap = (long long)MINB * sm_count();
  const long long grid = ntiles < cap ? ntiles : cap;
  CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, StreamSmem<C>::BYTES, current_stream(), r, lot, jump, tw, ntiles);
  count_launch();
  return cuda_ok(cudaGetLastError(), "pow2_r2c_stream_kernel launch");
}

template <int LOG2N>
struct StreamMinB {
  static constexpr int value = LOG2N >= 13 ? 1 : (Pow2Cfg<LOG2N>::LP == 3 ? 4 : 2);
};

template <int LOG2N, int DIR>
bool launch_c2c(long long lot, long long jump, cpx *c, double scale) {
  return launch_c2c_stream<Pow2Cfg<LOG2N>, StreamMinB<LOG2N>::value, DIR>(lot, jump, c, scale);
}
template <int LOG2N, int DIR>
bool launch_r2c(long long lot, long long jump, double *r) {
  // the bulk-copy engine needs 16-byte aligned rows: odd jumps (or an odd base) take the direct-load kernel
  const bool tma_ok = (jump % 2 == 0) && (((uintptr_t)r & 15) == 0);
  if (!tma_ok) return launch_r2c_cfg<Pow2Cfg<LOG2N>, (LOG2N >= 13 ? 1 : 2), DIR>(lot, jump, r);
  return launch_r2c_stream<Pow2Cfg<LOG2N>, StreamMinB<LOG2N>::value, DIR>(lot, jump, r);
}

/* streaming tile kernel (default); CFB200_TILE_DIRECT=1 selects the direct-load one for comparison */
template <int LOG2N, int DIR, int THREADS = 256>
bool launch_tile_stream(TileParams &P) {
  typedef Pow2Cfg<LOG2N, 4, 1, THREADS> C;
  P.tw = pow2_stream_table<C>();
  if (!P.tw) return false;
  auto kern = pow2_tile_stream_kernel<C, DIR>;
  if (!set_smem_once(kern, SMEM_LIMIT)) return false;
  const size_t smem = TileStreamSmem<C>::bytes(P.fs_count);
  if (smem > SMEM_LIMIT) return false;
  const long long ntiles = (P.lot + C::TPB - 1) / C::TPB;
  long long per_sm = (long long)((SMEM_LIMIT + 1024) / (smem + 1024));
  if (per_sm > (C::THREADS > 256 ? 1 : 2)) per_sm = (C::THREADS > 256 ? 1 : 2);
  if (per_sm < 1) per_sm = 1;
  long long grid = per_sm * sm_count();
  if (tile_cta_cap() > 0 && grid > tile_cta_cap()) grid = tile_cta_cap();  // share of the CTA slots (pipelined sweeps)
  if (grid > ntiles) grid = ntiles;
  CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, smem, current_stream(), P, ntiles);
  count_launch();
  return cuda_ok(cudaGetLastError(), "pow2_tile_stream_kernel launch");
}

/* TMA tensor-box variant: rows contiguous on the input side (jump_lo = 1), tiles never straddle an inner batch group */
template <int LOG2N, int DIR, bool STAGED, int THREADS = 256>
bool launch_tile_tma(TileParams &P, bool *declined) {  // *declined: the tensor map could not be built, use another kernel
  typedef Pow2Cfg<LOG2N, 4, 1, THREADS> C;
  P.tw = pow2_stream_table<C>();
  if (!P.tw) return false;
  auto kern = pow2_tile_tma_kernel<C, DIR, STAGED>;
  const size_t smem = TileTmaSmem<C, STAGED>::bytes(P.fs_count);
  if (!set_smem_once(kern, SMEM_LIMIT)) return false;
  TensorMap3 tm;
  memset(&tm, 0, sizeof(tm));
  const unsigned long long nlo = (unsigned long long)P.ain.nlo;
  const unsigned long long nhi = (unsigned long long)((P.lot + P.ain.nlo - 1) / P.ain.nlo);
  if (!STAGED && !make_tensor_map3(&tm, P.in, 2 * nlo, (unsigned long long)C::N, nhi, (unsigned long long)P.ain.inc * 16,
                        (unsigned long long)(P.ain.jump_hi ? P.ain.jump_hi : 1) * 16, 2 * C::TPB, C::N)) {
    *declined = true;
    return false;
  }
  const long long ntiles = (P.lot + C::TPB - 1) / C::TPB;
  long long per_sm = (long long)((SMEM_LIMIT + 1024) / (smem + 1024));
  const long long reg_cap = C::THREADS <= 128 ? 4 : 2;
  if (per_sm > reg_cap) per_sm = reg_cap;
  if (per_sm < 1) per_sm = 1;
  long long grid = per_sm * sm_count();
  if (tile_cta_cap() > 0 && grid > tile_cta_cap()) grid = tile_cta_cap();  // share of the CTA slots (pipelined sweeps)
  if (grid > ntiles) grid = ntiles;
  CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, smem, current_stream(), P, tm, ntiles);
  count_launch();
  return cuda_ok(cudaGetLastError(), "pow2_tile_tma_kernel launch");
}

template <int LOG2N, int DIR>
bool launch_tile(TileParams &P) {
  {
    typedef Pow2Cfg<LOG2N, 4, 1> C;
    static const bool no_tma = getenv("CFB200_TILE_NO_TMA") != nullptr;
    bool declined = false, ok = false;
    const bool box_ok = C::N <= 256 && 2 * C::TPB <= 256;
    const bool layout_ok = !P.in_staged && P.ain.jump_lo == 1 && P.ain.nlo % C::TPB == 0 && P.lot % P.ain.nlo == 0 &&
                           (((uintptr_t)P.in) & 15) == 0 && P.ain.inc > 0 && P.ain.jump_hi >= 0 &&
                           (unsigned long long)P.ain.inc * 16 < (1ULL << 40) && (unsigned long long)P.ain.jump_hi * 16 < (1ULL << 40);
    if (!no_tma && box_ok && layout_ok && TileTmaSmem<C, false>::bytes(P.fs_count) <= SMEM_LIMIT) {
      ok = launch_tile_tma<LOG2N, DIR, false>(P, &declined);
      if (!declined) return ok;
    }
    const bool rows_ok = P.in_staged && P.ain.inc == 1 && (((uintptr_t)P.in) & 15) == 0;
    if (!no_tma && rows_ok && TileTmaSmem<C, true>::bytes(P.fs_count) <= SMEM_LIMIT) return launch_tile_tma<LOG2N, DIR, true>(P, &declined);
  }
  static const bool direct = getenv("CFB200_TILE_DIRECT") != nullptr;
  if (!direct && TileStreamSmem<Pow2Cfg<LOG2N, 4, 1>>::bytes(P.fs_count) <= SMEM_LIMIT) return launch_tile_stream<LOG2N, DIR>(P);
  typedef Pow2Cfg<LOG2N, 4, 0> C;
  P.tw = pow2_table<C>();
  if (!P.tw) return false;
  auto kern = pow2_tile_kernel<C, DIR>;
  if (!set_smem_once(kern, SMEM_LIMIT)) return false;
  const size_t smem = TileSmem<C>::bytes(P.fs_count);
  if (smem > SMEM_LIMIT) {
    set_error("pow2_tile_kernel: %zu bytes of shared memory", smem);
    return false;
  }
  const long long grid = (P.lot + C::TPB - 1) / C::TPB;
  if (grid > 2147483647LL) {
    set_error("batch too large");
    return false;
  }
  CFB_LAUNCH(kern, (unsigned)grid, C::THREADS, smem, current_stream(), P);
  count_launch();
  return cuda_ok(cudaGetLastError(), "pow2_tile_kernel launch");
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

void pow2_release_tables() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto &kv : g_tw) cudaFree(kv.second);
  g_tw.clear();
}

int pow2_tile_min_log2() { return 6; }
int pow2_tile_max_log2() { return 10; }
bool pow2_tile_launch(int log2n, int dir, TileParams &P) {
  switch (log2n) {
    case 6: return dir < 0 ? launch_tile<6, -1>(P) : launch_tile<6, 1>(P);
    case 7: return dir < 0 ? launch_tile<7, -1>(P) : launch_tile<7, 1>(P);
    case 8: return dir < 0 ? launch_tile<8, -1>(P) : launch_tile<8, 1>(P);
    case 9: return dir < 0 ? launch_tile<9, -1>(P) : launch_tile<9, 1>(P);
    case 10: return dir < 0 ? launch_tile<10, -1>(P) : launch_tile<10, 1>(P);
    default: break;
  }
  set_error("pow2_tile_launch: unsupported row length 2^%d", log2n);
  return false;
}

bool pow2_c2c_launch(int n, long long lot, long long jump, int dir, cpx *c, double scale) {
  CFB_POW2_CASES(launch_c2c, lot, jump, c, scale)
  set_error("pow2_c2c_launch: unsupported length %d", n);
  return false;
}
bool pow2_r2c_launch(int n, long long lot, long long jump, int dir, double *r) {
  CFB_POW2_CASES(launch_r2c, lot, jump, r)
  set_error("pow2_r2c_launch: unsupported length %d", n);
  return false;
}

}  // namespace cfb
