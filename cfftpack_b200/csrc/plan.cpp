/* plan.cpp -- builds and caches the device-resident plans declared in plan.h. */
#include "plan.h"

#include <math.h>

#include <map>
#include <mutex>
#include <vector>

#include "internal.h"

namespace cfb {

void unit_root(long long num, long long den, double *re, double *im) {
  // reduce to the first octant so that cosl/sinl see a small argument; exact symmetries elsewhere
  num %= den;
  if (num < 0) num += den;
  // angle = 2 pi num/den; work with eighths of a turn: t = 8 num / den
  long long n8 = 8 * num;
  int oct = (int)(n8 / den);       // 0..7
  long long rem = n8 - oct * den;  // angle within the octant = (pi/4) rem/den
  const long double PI4 = 0.78539816339744830961566084581987572L;
  long double c, s;
  if (oct & 1) {  // measure back from the next octant boundary to stay in [0, pi/4]
    long double a = PI4 * (long double)(den - rem) / (long double)den;
    c = cosl(a);
    s = sinl(a);
  } else {
    long double a = PI4 * (long double)rem / (long double)den;
    c = cosl(a);
    s = sinl(a);
  }
  long double cr, sr;  // cos, sin of the full angle
  switch (oct) {
    case 0: cr = c; sr = s; break;
    case 1: cr = s; sr = c; break;
    case 2: cr = -s; sr = c; break;
    case 3: cr = -c; sr = s; break;
    case 4: cr = -c; sr = -s; break;
    case 5: cr = -s; sr = -c; break;
    case 6: cr = s; sr = -c; break;
    default: cr = c; sr = -s; break;
  }
  *re = (double)cr;
  *im = (double)(-sr);
}

int engine_factor(int M, int *radix) {
  // prime-power counts of the small primes
  int c2 = 0, c3 = 0, c5 = 0;
  while (M % 2 == 0) { M /= 2; ++c2; }
  while (M % 3 == 0) { M /= 3; ++c3; }
  while (M % 5 == 0) { M /= 5; ++c5; }
  int nf = 0;
  // the power of two that is NOT paired into radix 10 / 6 below goes first as 8s (the first pass's scattered stores
  // are conflict-free with one pad slot per 8 elements), then 4s / one 2 -- the engine keeps <= 13 points per thread
  int pair10 = c2 < c5 ? c2 : c5;          // (2,5) -> 10: one pass instead of two
  int a = c2 - pair10;
  int f5 = c5 - pair10;
  int pair9 = c3 / 2;                      // (3,3) -> 9
  int f3 = c3 - 2 * pair9;
  int pair6 = (a > 0 && f3 > 0 && a % 3 == 1) ? 1 : 0;  // a lone 2 joins a lone 3
  a -= pair6;
  f3 -= pair6;
  int tail[2], ntail = 0;
  switch (a % 3) {
    case 1:
      if (a >= 4) {
        tail[ntail++] = 4;
        tail[ntail++] = 4;
        a -= 4;
      } else {
        tail[ntail++] = 2;
        a -= 1;
      }
      break;
    case 2: tail[ntail++] = 4; a -= 2; break;
    default: break;
  }
  while (a >= 3) {
    radix[nf++] = 8;
    a -= 3;
  }
  for (int i = 0; i < ntail; ++i) radix[nf++] = tail[i];
  for (int i = 0; i < pair10; ++i) radix[nf++] = 10;
  for (int i = 0; i < pair9; ++i) radix[nf++] = 9;
  for (int i = 0; i < pair6; ++i) radix[nf++] = 6;
  for (int i = 0; i < f5; ++i) radix[nf++] = 5;
  for (int i = 0; i < f3; ++i) radix[nf++] = 3;
  for (int p = 7; (long long)p * p <= M; p += 2)
    while (M % p == 0) {
      radix[nf++] = p;
      M /= p;
    }
  if (M > 1) radix[nf++] = M;
  return nf;
}

namespace {
std::mutex g_mu;
struct Key {
  int dev, a, b;
  bool operator<(const Key &o) const { return dev != o.dev ? dev < o.dev : a != o.a ? a < o.a : b < o.b; }
};
std::map<Key, CorePlan *> g_core;
std::map<Key, TrigPlan *> g_trig;
std::map<Key, RootPlan *> g_root;
std::map<Key, ChirpPlan *> g_chirp;

int cur_dev() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}

template <class T>
T *upload(const std::vector<T> &h) {
  return (T *)upload_table(h.data(), h.size() * sizeof(T));
}
}  // namespace

const CorePlan *get_core_plan(int M) {
  std::lock_guard<std::mutex> lk(g_mu);
  Key k{cur_dev(), M, 0};
  auto it = g_core.find(k);
  if (it != g_core.end()) return it->second;
  CorePlan *pl = new CorePlan();
  pl->M = M;
  int radix[64];
  pl->nf = engine_factor(M, radix);
  if (pl->nf > CFB_MAXPASS) {
    set_error("length %d needs %d passes (max %d)", M, pl->nf, CFB_MAXPASS);
    delete pl;
    return nullptr;
  }
  std::vector<cpx> tw;
  int s = 1, cur = M;
  for (int i = 0; i < pl->nf; ++i) {
    int r = radix[i], m = cur / r;
    PassDesc &pd = pl->pass[i];
    pd.radix = r;
    pd.s = s;
    pd.m = m;
    pd.twoff = (int)tw.size();
    pd.rtoff = 0;
    {
      auto magic = [](long long d) -> unsigned { return d <= 1 ? 0u : (unsigned)(((1ULL << 32) + d - 1) / d); };
      const long long nb = M / r;
      pd.mag_s = magic(s);
      pd.mag_nb = magic(nb);
      pd.mag_per = magic(nb * generic_items(r));
    }
    if (m > 1)
      for (int kk = 1; kk < r; ++kk)
        for (int p = 0; p < m; ++p) {
          cpx w;
          unit_root((long long)p * kk, cur, &w.x, &w.y);
          tw.push_back(w);
        }
    if (r > 5 && r != 8) {
      pd.rtoff = (int)tw.size();
      for (int j = 0; j < r; ++j) {
        cpx w;
        unit_root(j, r, &w.x, &w.y);
        tw.push_back(w);
      }
    }
    if (r > pl->max_radix) pl->max_radix = r;
    s *= r;
    cur = m;
  }
  pl->tw_count = tw.size();
  pl->d_tw = upload(tw);
  if (!pl->d_tw) {
    delete pl;
    return nullptr;
  }
  g_core[k] = pl;
  return pl;
}

const TrigPlan *get_trig_plan(int kind, int n) {
  std::lock_guard<std::mutex> lk(g_mu);
  Key k{cur_dev(), kind, n};
  auto it = g_trig.find(k);
  if (it != g_trig.end()) return it->second;
  TrigPlan *pl = new TrigPlan();
  pl->kind = kind;
  pl->n = n;
  std::vector<double> t;
  const long double PI = 3.14159265358979323846264338327950288L;
  if (kind == K_COST) {
    // t[j] = 2 sin(j pi/M), t[M + j] = 2 cos(j pi/M), j < M   (cost1i_, fftpack.c:6145-6150)
    int M = n - 1;
    pl->M = M;
    t.resize(2 * (size_t)M);
    for (int j = 0; j < M; ++j) {
      double c, s;  // exp(-2 pi i j/(2M)) = (cos, -sin)(j pi/M)
      unit_root(j, 2LL * M, &c, &s);
      t[j] = -2.0 * s;
      t[M + j] = 2.0 * c;
    }
  } else if (kind == K_SINT) {
    // t[k-1] = 2 sin(k pi/(n+1))   (sint1i_, fftpack.c:14703-14705)
    int M = n + 1;
    pl->M = M;
    t.resize(n / 2 + 1);
    for (int kk = 1; kk <= n / 2; ++kk) {
      double c, s;
      unit_root(kk, 2LL * M, &c, &s);
      t[kk - 1] = -2.0 * s;
    }
  } else {
    // t[i] = cos((i+1) pi/(2n))   (cosq1i_, fftpack.c:5550-5557)
    pl->M = n;
    t.resize(n);
    for (int i = 0; i < n; ++i) {
      double c, s;
      unit_root(i + 1, 4LL * n, &c, &s);
      t[i] = c;
    }
  }
  (void)PI;
  pl->d_trig = upload(t);
  if (!pl->d_trig) {
    delete pl;
    return nullptr;
  }
  g_trig[k] = pl;
  return pl;
}

const RootPlan *get_root_plan(int n) {
  std::lock_guard<std::mutex> lk(g_mu);
  Key k{cur_dev(), n, 0};
  auto it = g_root.find(k);
  if (it != g_root.end()) return it->second;
  RootPlan *pl = new RootPlan();
  pl->n = n;
  int shift = 0;
  while ((1LL << (2 * shift)) < n) ++shift;
  pl->shift = shift;
  const int B = 1 << shift, H = (n + B - 1) / B;
  std::vector<cpx> w((size_t)B + H);
  for (int j = 0; j < B; ++j) unit_root(j, n, &w[j].x, &w[j].y);
  for (int j = 0; j < H; ++j) unit_root((long long)j * B, n, &w[B + j].x, &w[B + j].y);
  pl->d_w = upload(w);
  if (!pl->d_w) {
    delete pl;
    return nullptr;
  }
  g_root[k] = pl;
  return pl;
}

const ChirpPlan *get_chirp_plan(int n) {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    Key k{cur_dev(), n, 0};
    auto it = g_chirp.find(k);
    if (it != g_chirp.end()) return it->second;
  }
  int L = 1;
  while (L < 2 * n - 1) L *= 2;
  if (L < 64) L = 64;
  std::vector<cpx> ch(n), b((size_t)L);
  for (int j = 0; j < n; ++j) {
    // exp(-pi i j^2 / n) = exp(-2 pi i (j^2 mod 2n) / (2n)), reduced exactly in integers
    const long long q = ((long long)j * j) % (2LL * n);
    unit_root(q, 2LL * n, &ch[j].x, &ch[j].y);
  }
  for (auto &v : b) v = make_double2(0.0, 0.0);
  for (int j = 0; j < n; ++j) {
    const cpx cj = make_double2(ch[j].x, -ch[j].y);  // conjugate chirp
    b[j] = cj;
    if (j > 0) b[L - j] = cj;
  }
  ChirpPlan *pl = new ChirpPlan();
  pl->n = n;
  pl->L = L;
  pl->d_chirp = upload(ch);
  pl->d_bhat = upload(b);
  if (!pl->d_chirp || !pl->d_bhat) {
    delete pl;
    return nullptr;
  }
  // transform the kernel once on the device (unscaled forward transform of length L)
  if (!run_c2c_scaled(L, 1, 1, L, -1, pl->d_bhat, 1.0) || !cuda_ok(cudaStreamSynchronize(current_stream()), "chirp plan")) {
    delete pl;
    return nullptr;
  }
  std::lock_guard<std::mutex> lk(g_mu);
  Key k{cur_dev(), n, 0};
  g_chirp[k] = pl;
  return pl;
}

void pow2_release_tables();
void r10_release_tables();
void mix_release_tables();

/* The caller must be quiescent: no other host thread inside a transform (plans are handed out as raw pointers). */
void release_plans() {
  cudaDeviceSynchronize();  // kernels in flight may still read the tables
  pow2_release_tables();
  r10_release_tables();
  mix_release_tables();
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto &kv : g_core) {
    cudaFree(kv.second->d_tw);
    delete kv.second;
  }
  for (auto &kv : g_trig) {
    cudaFree(kv.second->d_trig);
    delete kv.second;
  }
  for (auto &kv : g_root) {
    cudaFree(kv.second->d_w);
    delete kv.second;
  }
  for (auto &kv : g_chirp) {
    cudaFree(kv.second->d_chirp);
    cudaFree(kv.second->d_bhat);
    delete kv.second;
  }
  g_chirp.clear();
  g_core.clear();
  g_trig.clear();
  g_root.clear();
}

}  // namespace cfb
