/*
 * radix10.cuh -- register-resident streaming kernels for N = 10^K (100, 1000): the radix-2/5 lengths of
 * BASELINE config 4 ("N = 1000, radix 2.2.2.5.5.5").  Same design as the power-of-two stream kernels (pow2.cuh):
 * persistent CTAs, TMA bulk loads one tile ahead, N/10 threads per sequence holding 10 points each, one in-register
 * radix-10 DFT + twiddle + shared-memory exchange per stage (1000 = 10.10.10 -> two exchanges), real and imaginary
 * parts exchanged in two rounds through one double-per-element tile.  Replaces for these lengths the passes
 * c1f2kf_/c1f5kf_ + c1f4kf_ (cfftpack/fftpack.c:195, :1145, :752) of cfftmf_/cfftmb_ and, with the pair packing of
 * pow2.cuh, the mradf2/mradf4/mradf5 passes (:8600, :8880, :9099) of rfftmf_/rfftmb_.
 * Threads per sequence are padded to a multiple of 16/128 (idle threads only take part in the barriers).
 */
#ifndef CFB_RADIX10_CUH
#define CFB_RADIX10_CUH
#include "pow2.cuh"

#ifndef CFB_R10_THREADS
#define CFB_R10_THREADS 128  // one sequence per CTA, 4-5 CTAs per SM: measured 3-11% faster than 256 (two per CTA)
#endif

namespace cfb {

/* 10-point DFT in registers: two 5-point DFTs (even / odd inputs) and a radix-2 level with the 10th roots */
template <int DIR>
__device__ __forceinline__ void dft10(cpx (&a)[10]) {
  const double C36 = 0.80901699437494742410229341718282, S36 = 0.58778525229247312916870595463907;
  const double C72 = 0.30901699437494742410229341718282, S72 = 0.95105651629515357211643933337938;
  cpx e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], e4 = a[8];
  cpx o0 = a[1], o1 = a[3], o2 = a[5], o3 = a[7], o4 = a[9];
  dft5<DIR>(e0, e1, e2, e3, e4);
  dft5<DIR>(o0, o1, o2, o3, o4);
  // X[k] = E[k] + w10^k O[k], X[k+5] = E[k] - w10^k O[k];  forward w10^k = (cos 36k, -sin 36k)
  o1 = ctw<DIR>(o1, make_double2(C36, -S36));
  o2 = ctw<DIR>(o2, make_double2(C72, -S72));
  o3 = ctw<DIR>(o3, make_double2(-C72, -S72));
  o4 = ctw<DIR>(o4, make_double2(-C36, -S36));
  a[0] = cadd(e0, o0);
  a[5] = csub(e0, o0);
  a[1] = cadd(e1, o1);
  a[6] = csub(e1, o1);
  a[2] = cadd(e2, o2);
  a[7] = csub(e2, o2);
  a[3] = cadd(e3, o3);
  a[8] = csub(e3, o3);
  a[4] = cadd(e4, o4);
  a[9] = csub(e4, o4);
}

/* a[k] *= w^k, k = 1..9, from w1 = w and w4 = w^4 */
template <int DIR>
__device__ __forceinline__ void twiddle_powers10(cpx (&a)[10], cpx w1, cpx w4) {
  cpx w2 = cmul(w1, w1), w3 = cmul(w2, w1);
  a[1] = ctw<DIR>(a[1], w1);
  a[2] = ctw<DIR>(a[2], w2);
  a[3] = ctw<DIR>(a[3], w3);
  a[4] = ctw<DIR>(a[4], w4);
  a[5] = ctw<DIR>(a[5], cmul(w4, w1));
  a[6] = ctw<DIR>(a[6], cmul(w4, w2));
  a[7] = ctw<DIR>(a[7], cmul(w4, w3));
  cpx w8 = cmul(w4, w4);
  a[8] = ctw<DIR>(a[8], w8);
  a[9] = ctw<DIR>(a[9], cmul(w8, w1));
}

template <int K>
struct R10Cfg {
  static constexpr int P = 10;
  static constexpr int N = (K == 2) ? 100 : 1000;
  static constexpr int NT = N / P;                       // working threads per sequence
  static constexpr int NTP = (NT <= 16) ? 16 : 128;      // threads reserved per sequence (padded)
  static constexpr int THREADS = (NTP > CFB_R10_THREADS) ? NTP : CFB_R10_THREADS;
  static constexpr int MINB = THREADS <= 128 ? 4 : 2;    // CTAs per SM the register budget is cut for
  static constexpr int TPB = THREADS / NTP;              // sequences (pairs) per CTA
  static constexpr int XT = N + 4 * (N / P);             // exchange row: 4 pad doubles per 10 (14-double pitch: rows stay
                                                         // 16-byte aligned and the 16-byte stores of stage 0 conflict-free)
  static constexpr int stage_s(int st) { return st == 0 ? 1 : st == 1 ? 10 : 100; }
  static constexpr int stage_m(int st) { return N / (stage_s(st) * P); }
  static constexpr int tws_offset(int st) {
    int off = 0;
    for (int i = 0; i < st; ++i) off += 2 * stage_m(i);
    return off;
  }
  static constexpr int TWS_COUNT = tws_offset(K - 1);    // the last stage has no twiddles
  static constexpr size_t LAND = (size_t)TPB * N * sizeof(cpx);
  static constexpr size_t XCH = (size_t)TPB * XT * sizeof(double);
  static constexpr size_t TWS = (size_t)(TWS_COUNT > 0 ? TWS_COUNT : 1) * sizeof(cpx);
  static constexpr size_t BYTES = LAND + XCH + TWS + 16;
  static constexpr size_t BYTES_TRIG = BYTES + (size_t)N * sizeof(double);  // + cosq table
  static __device__ __forceinline__ int xpad(int e) { return e + 4 * (e / P); }
};

/* all K stages; a[i] <-> element t + NT*i on entry and exit.  act = false: idle thread (barriers only) */
template <class C, int K, int DIR>
__device__ __forceinline__ void r10_core(cpx (&a)[10], double *__restrict__ xr, const int t, const bool act,
                                         const cpx *__restrict__ tws) {
  constexpr int P = 10, NT = C::NT;
#pragma unroll
  for (int st = 0; st < K; ++st) {
    const int s = C::stage_s(st), m = C::stage_m(st);
    dft10<DIR>(a);
    if (st < K - 1) {
      const int p = t / s, q = t % s;
      const cpx *twp = tws + C::tws_offset(st) + (act ? p : 0);
      twiddle_powers10<DIR>(a, twp[0], twp[m]);
      const int base = q + s * P * p;
      if (act) {
        if (st == 0) {
          double2 *row = (double2 *)(xr + C::xpad(base));
#pragma unroll
          for (int k = 0; k < P; k += 2) row[k / 2] = make_double2(a[k].x, a[k + 1].x);
        } else {
#pragma unroll
          for (int k = 0; k < P; ++k) xr[C::xpad(base + s * k)] = a[k].x;
        }
      }
      __syncthreads();
      if (act) {
#pragma unroll
        for (int i = 0; i < P; ++i) a[i].x = xr[C::xpad(t + NT * i)];
      }
      __syncthreads();
      if (act) {
        if (st == 0) {
          double2 *row = (double2 *)(xr + C::xpad(base));
#pragma unroll
          for (int k = 0; k < P; k += 2) row[k / 2] = make_double2(a[k].y, a[k + 1].y);
        } else {
#pragma unroll
          for (int k = 0; k < P; ++k) xr[C::xpad(base + s * k)] = a[k].y;
        }
      }
      __syncthreads();
      if (act) {
#pragma unroll
        for (int i = 0; i < P; ++i) a[i].y = xr[C::xpad(t + NT * i)];
      }
      __syncthreads();
    }
  }
}

template <int K, int DIR>
__global__ void __launch_bounds__(R10Cfg<K>::THREADS, R10Cfg<K>::MINB) r10_c2c_stream_kernel(cpx *__restrict__ c, long long lot, long long jump,
                                                                const cpx *__restrict__ tw, double scale, long long ntiles) {
  typedef R10Cfg<K> C;
  CFB_DYN_SMEM(smem_raw);
  constexpr int N = C::N, P = 10, NT = C::NT;
  cpx *land = (cpx *)smem_raw;
  double *xch = (double *)(smem_raw + C::LAND);
  cpx *tws = (cpx *)(smem_raw + C::LAND + C::XCH);
  uint64_t *bar = (uint64_t *)(smem_raw + C::LAND + C::XCH + C::TWS);
  const int tid = threadIdx.x, tl = tid / C::NTP, t = tid % C::NTP;
  const bool act = t < NT;
  if (tid == 0) mbar_init(bar, 1);
  for (int i = tid; i < C::TWS_COUNT; i += C::THREADS) tws[i] = __ldg(tw + i);
  __syncthreads();
  long long tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) stream_issue<C::TPB>((char *)land, (const char *)c, lot, jump * 16, tile, N * 16, bar);
  unsigned parity = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    mbar_wait(bar, parity);
    parity ^= 1;
    const long long g = tile * C::TPB + tl;
    const bool live = act && g < lot;
    cpx a[P];
#pragma unroll
    for (int i = 0; i < P; ++i) a[i] = act ? land[(size_t)tl * N + t + NT * i] : make_double2(0.0, 0.0);
    landing_reads_done<P, true>(a, (volatile unsigned *)(bar + 1));
    const long long next = tile + gridDim.x;
    __syncthreads();  // (a split arrive/sync barrier, as in pow2_c2c_stream_kernel, measured slower here: 1.82 vs 1.68 ms)
    if (tid == 0 && next < ntiles) stream_issue<C::TPB>((char *)land, (const char *)c, lot, jump * 16, next, N * 16, bar);
    r10_core<C, K, DIR>(a, xch + (size_t)tl * C::XT, t, act, tws);
    if (live) {
      cpx *x = c + g * jump + t;
#pragma unroll
      for (int i = 0; i < P; ++i) x[NT * i] = make_double2(a[i].x * scale, a[i].y * scale);
    }
  }
}

/* two real rows per complex transform (see pow2_r2c_stream_kernel).  KIND = K_RFFT: rfftmf_ / rfftmb_;
 * KIND = K_COSQ: cosqmf_ / cosqmb_ with the fold and the pairwise post-processing of cosqf1_/cosqb1_
 * (fftpack.c:5693-5738, :5604-5652) fused into the register load and the store (trig[i] = cos((i+1) pi / 2n) in
 * shared memory).  DIR = -1 forward, +1 backward (user-level direction). */
template <int K, int KIND, int DIR>
__global__ void __launch_bounds__(R10Cfg<K>::THREADS, R10Cfg<K>::MINB) r10_r2c_stream_kernel(double *__restrict__ r, long long lot, long long jump,
                                                                const cpx *__restrict__ tw, const double *__restrict__ trig_g,
                                                                long long ntiles) {
  typedef R10Cfg<K> C;
  CFB_DYN_SMEM(smem_raw);
  constexpr int N = C::N, P = 10, NT = C::NT;
  double *land = (double *)smem_raw;  // [TPB][2][N]
  double *xch = (double *)(smem_raw + C::LAND);
  cpx *tws = (cpx *)(smem_raw + C::LAND + C::XCH);
  uint64_t *bar = (uint64_t *)(smem_raw + C::LAND + C::XCH + C::TWS);
  double *W = (double *)(bar + 2);  // cosq: trig table, N doubles
  const int tid = threadIdx.x, tl = tid / C::NTP, t = tid % C::NTP;
  const bool act = t < NT;
  if (tid == 0) mbar_init(bar, 1);
  for (int i = tid; i < C::TWS_COUNT; i += C::THREADS) tws[i] = __ldg(tw + i);
  if (KIND == K_COSQ)
    for (int i = tid; i < N; i += C::THREADS) W[i] = __ldg(trig_g + i);
  __syncthreads();
  constexpr int ROWS = 2 * C::TPB;
  long long tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) stream_issue<ROWS>((char *)land, (const char *)r, lot, jump * 8, tile, N * 8, bar);
  unsigned parity = 0;
  const double *la = land + (size_t)(2 * tl) * N, *lb = la + N;
  double *xq = xch + (size_t)tl * C::XT;
  const int lane = tid & 31;
  for (; tile < ntiles; tile += gridDim.x) {
    mbar_wait(bar, parity);
    parity ^= 1;
    const long long ga = 2 * (tile * C::TPB + tl), gb = ga + 1;
    const bool va = act && ga < lot, vb = act && gb < lot;
    double *xa = r + (va ? ga : 0) * jump, *xb = r + (vb ? gb : 0) * jump;
    cpx a[P];
    if (DIR < 0) {
      if (KIND == K_COSQ) {
        // cosqf1_ fold: u[0] = x[0]; u[j] = W[j-1] (x[j]-x[n-j]) + W[n-j-1] (x[j]+x[n-j]);
        //               u[n-j] = W[j-1] (x[j]+x[n-j]) - W[n-j-1] (x[j]-x[n-j]);  u[n/2] = W[n/2-1] 2 x[n/2]
#pragma unroll
        for (int i = 0; i < P; ++i) {
          const int e = t + NT * i;
          a[i] = make_double2(0.0, 0.0);
          if (act) {
            if (e == 0) a[i] = make_double2(va ? la[0] : 0.0, vb ? lb[0] : 0.0);
            else if (e == N / 2) {
              const double w = 2.0 * W[N / 2 - 1];
              a[i] = make_double2(va ? w * la[e] : 0.0, vb ? w * lb[e] : 0.0);
            } else {
              const int j = e < N / 2 ? e : N - e, jc = N - j;
              const double wj = W[j - 1], wc = W[jc - 1];
              const double sa = va ? la[j] + la[jc] : 0.0, da = va ? la[j] - la[jc] : 0.0;
              const double sb = vb ? lb[j] + lb[jc] : 0.0, db = vb ? lb[j] - lb[jc] : 0.0;
              a[i] = e < N / 2 ? make_double2(fma(wj, da, wc * sa), fma(wj, db, wc * sb))
                               : make_double2(fma(wj, sa, -(wc * da)), fma(wj, sb, -(wc * db)));
            }
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < P; ++i) a[i] = make_double2(va ? la[t + NT * i] : 0.0, vb ? lb[t + NT * i] : 0.0);
      }
    } else {
      // half-complex row h -> spectrum Z (rfftb1_ convention).  cosqb1_ first forms h from x:
      //   h[0] = x[0]/2, h[2f-1] = (x[2f-1]+x[2f])/2, h[2f] = (x[2f-1]-x[2f])/2, h[n-1] = x[n-1]/2
#pragma unroll
      for (int i = 0; i < P; ++i) {
        const int e = t + NT * i;
        a[i] = make_double2(0.0, 0.0);
        if (act) {
          const double edge = (KIND == K_COSQ) ? 0.5 : 1.0;
          if (e == 0) a[i] = make_double2(va ? edge * la[0] : 0.0, vb ? edge * lb[0] : 0.0);
          else if (e == N / 2) a[i] = make_double2(va ? edge * la[N - 1] : 0.0, vb ? edge * lb[N - 1] : 0.0);
          else {
            const int f = e < N / 2 ? e : N - e;
            double h1a = va ? la[2 * f - 1] : 0.0, h2a = va ? la[2 * f] : 0.0;
            double h1b = vb ? lb[2 * f - 1] : 0.0, h2b = vb ? lb[2 * f] : 0.0;
            if (KIND == K_COSQ) {
              const double s1 = 0.5 * (h1a + h2a), d1 = 0.5 * (h1a - h2a), s2 = 0.5 * (h1b + h2b), d2 = 0.5 * (h1b - h2b);
              h1a = s1;
              h2a = d1;
              h1b = s2;
              h2b = d2;
            }
            const double a1 = 0.5 * h1a, a2 = 0.5 * h2a, b1 = 0.5 * h1b, b2 = 0.5 * h2b;
            a[i] = e < N / 2 ? make_double2(a1 + b2, b1 - a2) : make_double2(a1 - b2, b1 + a2);
          }
        }
      }
    }
    landing_reads_done(a, (volatile unsigned *)(bar + 1));
    __syncthreads();
    const long long next = tile + gridDim.x;
    if (tid == 0 && next < ntiles) stream_issue<ROWS>((char *)land, (const char *)r, lot, jump * 8, next, N * 8, bar);
    r10_core<C, K, DIR>(a, xq, t, act, tws);
    if (DIR < 0) {
      cpx *zq = (cpx *)xq;  // N/2 complex slots for the upper half of the spectrum
      if (act) {
#pragma unroll
        for (int i = P / 2; i < P; ++i) zq[t + NT * (i - P / 2)] = a[i];
      }
      __syncthreads();
      const double sc = 1.0 / (double)N;
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        const int f = t + NT * i;
        cpx u = a[i], v = act ? zq[f == 0 ? 0 : N / 2 - f] : make_double2(0.0, 0.0);
        double Aa, Ba, Ab, Bb;
        if (f == 0) {
          Aa = 0.0;
          Ab = 0.0;
          Ba = u.x * sc;
          Bb = u.y * sc;
        } else {
          Aa = (u.x + v.x) * sc;
          Ba = (v.y - u.y) * sc;
          Ab = (u.y + v.y) * sc;
          Bb = (u.x - v.x) * sc;
          if (KIND == K_COSQ) {  // cosqf1_ post: y[2f-1] = (A_f + B_f)/2, y[2f] = (A_f - B_f)/2
            const double s1 = 0.5 * (Aa + Ba), d1 = 0.5 * (Aa - Ba), s2 = 0.5 * (Ab + Bb), d2 = 0.5 * (Ab - Bb);
            Aa = s1;
            Ba = d1;
            Ab = s2;
            Bb = d2;
          }
        }
        const double Aa_n = __shfl_down_sync(0xffffffffu, Aa, 1), Ab_n = __shfl_down_sync(0xffffffffu, Ab, 1);
        if (lane < 31 && t != NT - 1) {
          if (va) *(double2 *)(xa + 2 * f) = make_double2(Ba, Aa_n);
          if (vb) *(double2 *)(xb + 2 * f) = make_double2(Bb, Ab_n);
        } else {
          if (va) xa[2 * f] = Ba;
          if (vb) xb[2 * f] = Bb;
        }
        if ((lane == 0 || t == 0) && f != 0) {
          if (va) xa[2 * f - 1] = Aa;
          if (vb) xb[2 * f - 1] = Ab;
        }
        if (f == 0) {
          if (va) xa[N - 1] = v.x * sc;
          if (vb) xb[N - 1] = v.y * sc;
        }
      }
    } else if (KIND == K_COSQ) {
      // cosqb1_ post: y[0] = 2 u[0]; y[j] = p + q, y[n-j] = p - q with p = W[j-1] u[n-j] + W[n-j-1] u[j],
      // q = W[j-1] u[j] - W[n-j-1] u[n-j]; y[n/2] = W[n/2-1] 2 u[n/2].  u[n-j] of the upper half goes through zq.
      cpx *zq = (cpx *)xq;
      if (act) {
#pragma unroll
        for (int i = P / 2; i < P; ++i) zq[t + NT * (i - P / 2)] = a[i];
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        const int j = t + NT * i;
        if (!act) continue;
        if (j == 0) {
          const cpx mid = zq[0];
          const double w = 2.0 * W[N / 2 - 1];
          if (va) {
            xa[0] = a[0].x + a[0].x;
            xa[N / 2] = w * mid.x;
          }
          if (vb) {
            xb[0] = a[0].y + a[0].y;
            xb[N / 2] = w * mid.y;
          }
        } else {
          const cpx uj = a[i], uc = zq[N / 2 - j];
          const double wj = W[j - 1], wc = W[N - j - 1];
          const double pa = fma(wj, uc.x, wc * uj.x), qa = fma(wj, uj.x, -(wc * uc.x));
          const double pb = fma(wj, uc.y, wc * uj.y), qb = fma(wj, uj.y, -(wc * uc.y));
          if (va) {
            xa[j] = pa + qa;
            xa[N - j] = pa - qa;
          }
          if (vb) {
            xb[j] = pb + qb;
            xb[N - j] = pb - qb;
          }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < P; ++i) {
        if (va) xa[t + NT * i] = a[i].x;
        if (vb) xb[t + NT * i] = a[i].y;
      }
    }
  }
}

/* DCT-I of length n = N + 1 = 1001 (costmf_ / costmb_, fftpack.c:6485, :6419 -> mcstf1_ :7150): the real transform
 * underneath has length N = 1000.  Rows are contiguous (jump = n, odd), so a PAIR of rows is one 16-byte aligned run
 * of 2n doubles -> one bulk copy per pair.  The pre-fold (:6355-6377) happens on the way from the landing buffer to
 * the registers; dsum is a block reduction; the serial post recurrence (:6386-6400) is a two-level scan over the
 * threads of the sequence.  Both directions use the forward real transform, like the reference. */
struct R10Cost {
  typedef R10Cfg<3> C;
  static constexpr int N = 1000, n = 1001, NT = 100, NW = 4;  // NW warps per sequence
  static constexpr size_t LAND = (size_t)C::TPB * 2 * n * sizeof(double);  // 32032 B (multiple of 16)
  static constexpr size_t OFF_XCH = LAND;
  static constexpr size_t OFF_TWS = OFF_XCH + C::XCH;
  static constexpr size_t OFF_BAR = OFF_TWS + C::TWS;
  static constexpr size_t OFF_TRIG = OFF_BAR + 16;                          // S[500], C[500]
  static constexpr size_t OFF_RED = OFF_TRIG + (size_t)N * sizeof(double);  // per sequence: dsum partials [2][NW] + scan totals [2][5][NW]
  static constexpr int RED_PER_SEQ = 2 * NW + 2 * 5 * NW;
  static constexpr size_t BYTES = OFF_RED + (size_t)C::TPB * RED_PER_SEQ * sizeof(double) + 16;
};

template <int DIR>
__global__ void __launch_bounds__(R10Cfg<3>::THREADS, R10Cfg<3>::MINB) r10_cost_stream_kernel(double *__restrict__ r, long long npairs,
                                                                 const cpx *__restrict__ tw, const double *__restrict__ trig_g,
                                                                 long long ntiles) {
  typedef R10Cfg<3> C;
  typedef R10Cost R;
  CFB_DYN_SMEM(smem_raw);
  constexpr int N = R::N, n = R::n, P = 10, NT = R::NT, NW = R::NW;
  double *land = (double *)smem_raw;
  double *xch = (double *)(smem_raw + R::OFF_XCH);
  cpx *tws = (cpx *)(smem_raw + R::OFF_TWS);
  uint64_t *bar = (uint64_t *)(smem_raw + R::OFF_BAR);
  double *Ssm = (double *)(smem_raw + R::OFF_TRIG), *Csm = Ssm + N / 2;
  double *red = (double *)(smem_raw + R::OFF_RED);
  const int tid = threadIdx.x, tl = tid / C::NTP, t = tid % C::NTP, lane = tid & 31, w = t >> 5;
  const bool act = t < NT;
  if (tid == 0) mbar_init(bar, 1);
  for (int i = tid; i < C::TWS_COUNT; i += C::THREADS) tws[i] = __ldg(tw + i);
  for (int i = tid; i < N / 2; i += C::THREADS) {
    Ssm[i] = __ldg(trig_g + i);       // 2 sin(i pi / N)
    Csm[i] = __ldg(trig_g + N + i);   // 2 cos(i pi / N)
  }
  __syncthreads();
  auto issue = [&](long long tile) {  // thread 0: one bulk copy for the live pairs of the tile (contiguous)
    const long long p0 = tile * C::TPB;
    const int live = (int)((npairs - p0) < C::TPB ? (npairs - p0) : C::TPB);
    const unsigned bytes = (unsigned)live * 2 * n * 8;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(land, r + p0 * 2 * n, bytes, bar);
  };
  long long tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) issue(tile);
  unsigned parity = 0;
  const double *la = land + (size_t)tl * 2 * n, *lb = la + n;
  double *xq = xch + (size_t)tl * C::XT;
  double *rd = red + (size_t)tl * R::RED_PER_SEQ;  // [2][NW] dsum partials, then [2][5][NW] scan totals
  const double ends = DIR > 0 ? 2.0 : 1.0;         // the backward transform doubles the end points first
  for (; tile < ntiles; tile += gridDim.x) {
    mbar_wait(bar, parity);
    parity ^= 1;
    const long long pr = tile * C::TPB + tl;
    const bool live = act && pr < npairs;
    double *xa = r + (live ? pr : 0) * 2 * n, *xb = xa + n;
    cpx a[P];
    double pa = 0.0, pb = 0.0;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      const int e = t + NT * i;
      a[i] = make_double2(0.0, 0.0);
      if (live) {
        if (e == 0) a[i] = make_double2(ends * (la[0] + la[n - 1]), ends * (lb[0] + lb[n - 1]));
        else if (e == N / 2) a[i] = make_double2(la[e] + la[e], lb[e] + lb[e]);
        else {
          const int j = e < N / 2 ? e : N - e, jc = N - j;
          const double t1a = la[j] + la[jc], t2a = la[j] - la[jc], t1b = lb[j] + lb[jc], t2b = lb[j] - lb[jc];
          const double sj = Ssm[j];
          if (e < N / 2) {
            const double cj = Csm[j];
            pa = fma(cj, t2a, pa);
            pb = fma(cj, t2b, pb);
            a[i] = make_double2(fma(-sj, t2a, t1a), fma(-sj, t2b, t1b));
          } else {
            a[i] = make_double2(fma(sj, t2a, t1a), fma(sj, t2b, t1b));
          }
        }
      }
    }
    const double x0a = live ? ends * la[0] : 0.0, xna = live ? ends * la[n - 1] : 0.0;
    const double x0b = live ? ends * lb[0] : 0.0, xnb = live ? ends * lb[n - 1] : 0.0;
    pa = warp_sum(pa);
    pb = warp_sum(pb);
    if (lane == 0) {
      rd[w] = pa;
      rd[NW + w] = pb;
    }
    {
      const cpx ends4[2] = {make_double2(x0a, xna), make_double2(x0b, xnb)};
      landing_reads_done(a, (volatile unsigned *)(bar + 1));
      landing_reads_done(ends4, (volatile unsigned *)(bar + 1));
    }
    __syncthreads();  // landing buffer consumed; dsum partials visible
    const long long next = tile + gridDim.x;
    if (tid == 0 && next < ntiles) issue(next);
    double dsa = x0a - xna, dsb = x0b - xnb;
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      dsa += rd[k];
      dsb += rd[NW + k];
    }
    r10_core<C, 3, -1>(a, xq, t, act, tws);  // forward real transform in both directions (costb1_ calls rfft1f_ too)
    cpx *zq = (cpx *)xq;
    if (act) {
#pragma unroll
      for (int i = P / 2; i < P; ++i) zq[t + NT * (i - P / 2)] = a[i];
    }
    __syncthreads();
    // h[2f-1] = A_f, h[2f] = B_f (scaled like rfftf1_);  y[2f] = c1 A_f, y[2f-1] = D + sum_{m<f} c1 B_m, y[0] = c0 h[0]
    const double sc = 1.0 / (double)N;
    const double c0 = DIR < 0 ? 0.5 : 0.5 * (double)N, c1 = DIR < 0 ? 0.5 : 0.25 * (double)N;
    const double Da = DIR < 0 ? dsa * sc : 0.5 * dsa, Db = DIR < 0 ? dsb * sc : 0.5 * dsb;
    double Aa[P / 2], Ab[P / 2], inca[P / 2], incb[P / 2], va_[P / 2], vb_[P / 2];
    double *tot = rd + 2 * NW;  // [2][5][NW]
#pragma unroll
    for (int i = 0; i < P / 2; ++i) {
      const int f = t + NT * i;
      const cpx u = a[i], v = act ? zq[f == 0 ? 0 : N / 2 - f] : make_double2(0.0, 0.0);
      double Ba, Bb;
      if (f == 0 || !act) {
        Aa[i] = Ab[i] = 0.0;
        Ba = Bb = 0.0;  // slot 0 is h[0]: not part of the running sum
      } else {
        Aa[i] = (u.x + v.x) * sc;
        Ba = (v.y - u.y) * sc;
        Ab[i] = (u.y + v.y) * sc;
        Bb = (u.x - v.x) * sc;
      }
      va_[i] = c1 * Ba;
      vb_[i] = c1 * Bb;
      // inclusive scan over the lanes of the warp
      double sa = va_[i], sb = vb_[i];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double ua = __shfl_up_sync(0xffffffffu, sa, o), ub = __shfl_up_sync(0xffffffffu, sb, o);
        if (lane >= o) {
          sa += ua;
          sb += ub;
        }
      }
      inca[i] = sa;
      incb[i] = sb;
      if (lane == 31) {
        tot[i * NW + w] = sa;
        tot[5 * NW + i * NW + w] = sb;
      }
    }
    __syncthreads();
    {
      double carrya = Da, carryb = Db;  // prefix of everything before block i, warp w
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        double ca = carrya, cb = carryb;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
          const double ta = tot[i * NW + k], tb = tot[5 * NW + i * NW + k];
          if (k < w) {
            ca += ta;
            cb += tb;
          }
          carrya += ta;
          carryb += tb;
        }
        const int f = t + NT * i;
        if (live) {
          if (f == 0) {
            xa[0] = c0 * (a[0].x * sc);
            xb[0] = c0 * (a[0].y * sc);
          } else {
            xa[2 * f - 1] = ca + inca[i] - va_[i];
            xa[2 * f] = c1 * Aa[i];
            xb[2 * f - 1] = cb + incb[i] - vb_[i];
            xb[2 * f] = c1 * Ab[i];
          }
        }
      }
      if (live && t == 0) {  // f = N/2: y[n-2] = D + all of the sum, y[n-1] from X_{N/2}
        const cpx mid = zq[0];
        const double lf = DIR < 0 ? c1 : 2.0 * c1;
        xa[n - 2] = carrya;
        xa[n - 1] = lf * (mid.x * sc);
        xb[n - 2] = carryb;
        xb[n - 1] = lf * (mid.y * sc);
      }
    }
    // zq / tot are rewritten by the next tile only after its landing barrier
  }
}

/* host side (radix10.cu) */
bool r10_supported(int n);
bool r10_c2c_launch(int n, long long lot, long long jump, int dir, cpx *c, double scale);
bool r10_r2c_launch(int n, long long lot, long long jump, int dir, double *r);
bool r10_cosq_launch(int n, long long lot, long long jump, int dir, double *x, const double *trig);
/* costmf_/costmb_ n = 1001, contiguous rows, an even number of rows */
bool r10_cost_launch(long long npairs, int dir, double *x, const double *trig);
void r10_release_tables();

}  // namespace cfb
#endif
