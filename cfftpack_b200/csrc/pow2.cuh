/*
 * pow2.cuh -- register-resident kernels for contiguous power-of-two batches: the headline shape of
 * cfftmf_/cfftmb_ and rfftmf_/rfftmb_ (N = 4096, lot = 65536, inc = 1, jump >= N).
 *
 * They replace, for those shapes, the whole pass loop of cmfm1f_/cmfm1b_ (cfftpack/fftpack.c:5262, :5150)
 * with its cmf4kf_ sweeps (:3560) and of mrftf1_/mrftb1_ (:10149, :9946) with mradf4/mradb4 (:8880, :7603):
 * the reference makes log4(N) read+write sweeps over the whole array; here every sequence is read from
 * HBM once, transformed on chip and written once.
 *
 *   - N/P threads own one sequence, P = 16 (8 for N < 256) points per thread in registers;
 *   - thread t always holds elements t + (N/P) i, so global loads/stores are 16-byte, unit-stride per warp;
 *   - a stage is an in-register radix-P DFT (butterfly.cuh), a twiddle multiply and one exchange through a
 *     padded shared-memory tile (pitch 17/16 resp. 9/8 -> bank-conflict free for 16-byte elements);
 *   - the last stage (radix P or the leftover 2/4/8) leaves natural order in the same registers;
 *   - real sequences go through two at a time as z = x_a + i x_b and are separated with the Hermitian
 *     symmetry (forward) or merged (backward) using half an exchange.
 */
#ifndef CFB_POW2_CUH
#define CFB_POW2_CUH
#include <mutex>

#include "butterfly.cuh"
#include "internal.h"

namespace cfb {

template <int LOG2N>
struct Pow2Cfg {
  static constexpr int N = 1 << LOG2N;
  static constexpr int LP = (LOG2N >= 8) ? 4 : 3;  // log2 of points per thread
  static constexpr int P = 1 << LP;
  static constexpr int NT = N / P;                 // threads per sequence
  static constexpr int THREADS = (NT > 256) ? NT : 256;
  static constexpr int TPB = THREADS / NT;         // sequences (or pairs) per CTA
  static constexpr int NFULL = LOG2N / LP;
  static constexpr int REM = LOG2N % LP;
  static constexpr int TILE = N + (N >> LP);       // padded elements per sequence
  static constexpr size_t SMEM = (size_t)TPB * TILE * sizeof(cpx);
  // twiddle table: for full stage st (not last): (P-1) * m_st entries laid out [k-1][p]
  static constexpr int tw_offset(int st) {
    int off = 0;
    for (int i = 0; i < st; ++i) off += (P - 1) * (N >> (LP * (i + 1)));
    return off;
  }
  static constexpr int TW_COUNT = tw_offset(NFULL);
};

template <int LP>
__device__ __forceinline__ int pad(int e) {
  return e + (e >> LP);
}

/* all stages of one length-N transform; a[i] <-> element t + NT*i on entry and on exit (natural order) */
template <int LOG2N, int DIR>
__device__ __forceinline__ void pow2_core(cpx (&a)[Pow2Cfg<LOG2N>::P], cpx *__restrict__ sm, const int t,
                                          const cpx *__restrict__ tw) {
  typedef Pow2Cfg<LOG2N> C;
  constexpr int P = C::P, LP = C::LP, NT = C::NT;
#pragma unroll
  for (int st = 0; st < C::NFULL; ++st) {
    const int s = 1 << (LP * st);            // product of earlier radices
    const int m = C::N >> (LP * (st + 1));   // remaining length / P
    const bool last = (st == C::NFULL - 1) && (C::REM == 0);
    Dft<P, DIR>::run(a);
    if (!last) {
      const int p = t >> (LP * st), q = t & (s - 1);
      const cpx *twp = tw + C::tw_offset(st) + p;
      if (m > 1) {
#pragma unroll
        for (int k = 1; k < P; ++k) a[k] = ctw<DIR>(a[k], __ldg(twp + (k - 1) * m));
      }
      const int base = q + s * P * p;
#pragma unroll
      for (int k = 0; k < P; ++k) sm[pad<LP>(base + s * k)] = a[k];
      __syncthreads();
#pragma unroll
      for (int i = 0; i < P; ++i) a[i] = sm[pad<LP>(t + NT * i)];
      __syncthreads();
    }
  }
  if (C::REM > 0) {
    constexpr int R = 1 << (C::REM > 0 ? C::REM : 1), G = P / R;
#pragma unroll
    for (int u = 0; u < G; ++u) {
      cpx b[R];
#pragma unroll
      for (int j = 0; j < R; ++j) b[j] = a[u + G * j];
      Dft<R, DIR>::run(b);
#pragma unroll
      for (int j = 0; j < R; ++j) a[u + G * j] = b[j];
    }
  }
}

template <int LOG2N, int DIR>
__global__ void __launch_bounds__(Pow2Cfg<LOG2N>::THREADS) pow2_c2c_kernel(cpx *__restrict__ c, long long lot,
                                                                             long long jump,
                                                                             const cpx *__restrict__ tw, double scale) {
  typedef Pow2Cfg<LOG2N> C;
  CFB_DYN_SMEM(smem_raw);
  const int tl = threadIdx.x / C::NT, t = threadIdx.x % C::NT;
  const long long g = (long long)blockIdx.x * C::TPB + tl;
  const bool live = g < lot;
  cpx *sm = (cpx *)smem_raw + (size_t)tl * C::TILE;
  cpx *x = c + (live ? g : 0) * jump + t;
  cpx a[C::P];
#pragma unroll
  for (int i = 0; i < C::P; ++i) a[i] = live ? x[C::NT * i] : make_double2(0.0, 0.0);
  pow2_core<LOG2N, DIR>(a, sm, t, tw);
  if (live) {
#pragma unroll
    for (int i = 0; i < C::P; ++i) x[C::NT * i] = make_double2(a[i].x * scale, a[i].y * scale);
  }
}

/* two real sequences per complex transform.  DIR = -1: rfftmf_ (x -> scaled half-complex, fftpack.c:10281-10349),
 * DIR = +1: rfftmb_ (half-complex -> x). */
template <int LOG2N, int DIR>
__global__ void __launch_bounds__(Pow2Cfg<LOG2N>::THREADS) pow2_r2c_kernel(double *__restrict__ r, long long lot,
                                                                             long long jump,
                                                                             const cpx *__restrict__ tw) {
  typedef Pow2Cfg<LOG2N> C;
  constexpr int N = C::N, P = C::P, NT = C::NT, LP = C::LP;
  CFB_DYN_SMEM(smem_raw);
  const int tl = threadIdx.x / NT, t = threadIdx.x % NT;
  const long long pr = (long long)blockIdx.x * C::TPB + tl;  // pair index
  const long long ga = 2 * pr, gb = 2 * pr + 1;
  const bool la = ga < lot, lb = gb < lot;
  cpx *sm = (cpx *)smem_raw + (size_t)tl * C::TILE;
  double *xa = r + (la ? ga : 0) * jump, *xb = r + (lb ? gb : 0) * jump;
  cpx a[P];
  if (DIR < 0) {
#pragma unroll
    for (int i = 0; i < P; ++i) a[i] = make_double2(la ? xa[t + NT * i] : 0.0, lb ? xb[t + NT * i] : 0.0);
    pow2_core<LOG2N, DIR>(a, sm, t, tw);
    // separate X_a, X_b: needs Z[N-f]; the upper half of the registers goes through shared memory
#pragma unroll
    for (int i = P / 2; i < P; ++i) sm[pad<LP>(t + NT * i)] = a[i];
    __syncthreads();
    const double sc = 1.0 / (double)N;
#pragma unroll
    for (int i = 0; i < P / 2; ++i) {
      const int f = t + NT * i;
      if (f == 0) {
        cpx v = sm[pad<LP>(N / 2)];
        if (la) {
          xa[0] = a[0].x * sc;
          xa[N - 1] = v.x * sc;
        }
        if (lb) {
          xb[0] = a[0].y * sc;
          xb[N - 1] = v.y * sc;
        }
      } else {
        cpx u = a[i], v = sm[pad<LP>(N - f)];
        if (la) {
          xa[2 * f - 1] = (u.x + v.x) * sc;
          xa[2 * f] = (v.y - u.y) * sc;
        }
        if (lb) {
          xb[2 * f - 1] = (u.y + v.y) * sc;
          xb[2 * f] = (u.x - v.x) * sc;
        }
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < P / 2; ++i) {
      const int f = t + NT * i;
      if (f == 0) {
        a[0] = make_double2(la ? xa[0] : 0.0, lb ? xb[0] : 0.0);
        sm[pad<LP>(N / 2)] = make_double2(la ? xa[N - 1] : 0.0, lb ? xb[N - 1] : 0.0);
      } else {
        double a1 = la ? 0.5 * xa[2 * f - 1] : 0.0, a2 = la ? 0.5 * xa[2 * f] : 0.0;
        double b1 = lb ? 0.5 * xb[2 * f - 1] : 0.0, b2 = lb ? 0.5 * xb[2 * f] : 0.0;
        a[i] = make_double2(a1 + b2, b1 - a2);
        sm[pad<LP>(N - f)] = make_double2(a1 - b2, b1 + a2);
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = P / 2; i < P; ++i) a[i] = sm[pad<LP>(t + NT * i)];
    __syncthreads();
    pow2_core<LOG2N, DIR>(a, sm, t, tw);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      if (la) xa[t + NT * i] = a[i].x;
      if (lb) xb[t + NT * i] = a[i].y;
    }
  }
}

/* ---- host side ---- */
bool pow2_c2c_supported(int n, long long inc, long long jump, int aligned16);
bool pow2_r2c_supported(int n, long long inc, long long jump, int aligned16);
bool pow2_c2c_launch(int n, long long lot, long long jump, int dir, cpx *c);
bool pow2_r2c_launch(int n, long long lot, long long jump, int dir, double *r);

}  // namespace cfb
#endif
