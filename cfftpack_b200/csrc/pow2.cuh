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
#include "engine_types.h"
#include "internal.h"
#include "tma.cuh"

namespace cfb {

/* LP = log2(points per thread); TW = 1: the twiddles of a stage with many distinct values are rebuilt in
 * registers from two table entries (w^p, w^4p) instead of P-1 loads (keeps the tables L1-resident) */
template <int LOG2N, int LP_ = ((LOG2N >= 8) ? 4 : 3), int TW_ = 1, int MINTHREADS_ = 256>
struct Pow2Cfg {
  static constexpr int N = 1 << LOG2N;
  static constexpr int LP = LP_;
  static constexpr int TW = TW_;
  static constexpr int P = 1 << LP;
  static constexpr int NT = N / P;                 // threads per sequence
  static constexpr int THREADS = (NT > MINTHREADS_) ? NT : MINTHREADS_;
  static constexpr int TPB = THREADS / NT;         // sequences (or pairs) per CTA
  static constexpr int NFULL = LOG2N / LP;
  static constexpr int REM = LOG2N % LP;
  static constexpr int TILE = N + (N >> LP);       // padded elements per sequence
  static constexpr size_t SMEM = (size_t)TPB * TILE * sizeof(cpx);
  static constexpr int stage_m(int st) { return N >> (LP * (st + 1)); }
  static constexpr bool stage_last(int st) { return st == NFULL - 1 && REM == 0; }
  static constexpr bool stage_computed(int st) { return TW == 1 && stage_m(st) >= 64; }
  // table entries of stage st: 2*m (w^p, w^4p) when rebuilt in registers, else (P-1)*m laid out [k-1][p]
  static constexpr int stage_count(int st) {
    return stage_last(st) ? 0 : stage_computed(st) ? 2 * stage_m(st) : (P - 1) * stage_m(st);
  }
  static constexpr int tw_offset(int st) {
    int off = 0;
    for (int i = 0; i < st; ++i) off += stage_count(i);
    return off;
  }
  static constexpr int TW_COUNT = tw_offset(NFULL);
};

/* a[k] *= w^k (k = 1..P-1) with w^2, w^3, ... formed from w1 = w and w4 = w^4 (at most three products deep) */
template <int DIR>
__device__ __forceinline__ void twiddle_powers(cpx (&a)[8], cpx w1, cpx w4) {
  cpx w2 = cmul(w1, w1), w3 = cmul(w2, w1);
  a[1] = ctw<DIR>(a[1], w1);
  a[2] = ctw<DIR>(a[2], w2);
  a[3] = ctw<DIR>(a[3], w3);
  a[4] = ctw<DIR>(a[4], w4);
  a[5] = ctw<DIR>(a[5], cmul(w4, w1));
  a[6] = ctw<DIR>(a[6], cmul(w4, w2));
  a[7] = ctw<DIR>(a[7], cmul(w4, w3));
}
template <int DIR>
__device__ __forceinline__ void twiddle_powers(cpx (&a)[16], cpx w1, cpx w4) {
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
  a[10] = ctw<DIR>(a[10], cmul(w8, w2));
  a[11] = ctw<DIR>(a[11], cmul(w8, w3));
  cpx w12 = cmul(w8, w4);
  a[12] = ctw<DIR>(a[12], w12);
  a[13] = ctw<DIR>(a[13], cmul(w12, w1));
  a[14] = ctw<DIR>(a[14], cmul(w12, w2));
  a[15] = ctw<DIR>(a[15], cmul(w12, w3));
}

template <int LP>
__device__ __forceinline__ int pad(int e) {
  return e + (e >> LP);
}

/* all stages of one length-N transform; a[i] <-> element t + NT*i on entry and on exit (natural order) */
template <class C, int DIR>
__device__ __forceinline__ void pow2_core(cpx (&a)[C::P], cpx *__restrict__ sm, const int t,
                                          const cpx *__restrict__ tw) {
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
      if (C::stage_computed(st)) {
        twiddle_powers<DIR>(a, __ldg(twp), __ldg(twp + m));
      } else if (m > 1) {
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

/* two real sequences per complex transform.  DIR = -1: rfftmf_ (x -> scaled half-complex, fftpack.c:10281-10349),
 * DIR = +1: rfftmb_ (half-complex -> x). */
template <class C, int MINB, int DIR>
__global__ void __launch_bounds__(C::THREADS, MINB) pow2_r2c_kernel(double *__restrict__ r, long long lot, long long jump,
                                                                    const cpx *__restrict__ tw) {
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
    pow2_core<C, DIR>(a, sm, t, tw);
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
    pow2_core<C, DIR>(a, sm, t, tw);
#pragma unroll
    for (int i = 0; i < P; ++i) {
      if (la) xa[t + NT * i] = a[i].x;
      if (lb) xb[t + NT * i] = a[i].y;
    }
  }
}

/* ---------------------------------------------------------------------------------------------------------
 * Streaming variants.  A persistent CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the next tile is
 * brought into a shared-memory landing buffer by the bulk-copy engine (TMA, one cp.async.bulk per sequence)
 * while the current one is being transformed, so HBM reads never wait for registers or for the butterflies.
 * To make room for the landing buffer the exchange tile holds one double per element: real and imaginary
 * parts are exchanged in two rounds through the same 8-byte padded tile.
 * --------------------------------------------------------------------------------------------------------- */
template <class C>
struct StreamSmem {
  static constexpr int PADW = (C::LP == 4) ? 2 : 1;  // doubles of padding per P elements
  static constexpr int XTILE = C::N + PADW * (C::N >> C::LP);
  // twiddles of the streaming kernels: per non-last stage the two rows w^p and w^(4p), p < m, kept in shared memory
  static constexpr int tws_offset(int st) {
    int off = 0;
    for (int i = 0; i < st; ++i) off += C::stage_last(i) ? 0 : 2 * C::stage_m(i);
    return off;
  }
  static constexpr int TWS_COUNT = tws_offset(C::NFULL);
  static constexpr size_t LAND = (size_t)C::TPB * C::N * sizeof(cpx);
  static constexpr size_t XCH = (size_t)C::TPB * XTILE * sizeof(double);
  static constexpr size_t TWS = (size_t)(TWS_COUNT > 0 ? TWS_COUNT : 1) * sizeof(cpx);
  static constexpr size_t BYTES = LAND + XCH + TWS + 16;
};

template <class C>
__device__ __forceinline__ int xpad(int e) {
  return e + StreamSmem<C>::PADW * (e >> C::LP);
}

template <class C, int DIR, bool VEC0 = true>
__device__ __forceinline__ void pow2_core_split(cpx (&a)[C::P], double *__restrict__ xr, const int t,
                                                const cpx *__restrict__ tws) {
  constexpr int P = C::P, LP = C::LP, NT = C::NT;
  typedef StreamSmem<C> S;
#pragma unroll
  for (int st = 0; st < C::NFULL; ++st) {
    const int s = 1 << (LP * st);
    const int m = C::N >> (LP * (st + 1));
    const bool last = (st == C::NFULL - 1) && (C::REM == 0);
    Dft<P, DIR>::run(a);
    if (!last) {
      const int p = t >> (LP * st), q = t & (s - 1);
      const cpx *twp = tws + S::tws_offset(st) + p;
      twiddle_powers<DIR>(a, twp[0], twp[m]);
      const int base = q + s * P * p;
      if (VEC0 && st == 0 && S::PADW == 2) {
        // first exchange: a thread's P outputs are adjacent -> 16-byte stores (row pitch P+2 keeps them aligned)
        double2 *row = (double2 *)(xr + xpad<C>(base));
#pragma unroll
        for (int k = 0; k < P; k += 2) row[k / 2] = make_double2(a[k].x, a[k + 1].x);
      } else {
#pragma unroll
        for (int k = 0; k < P; ++k) xr[xpad<C>(base + s * k)] = a[k].x;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < P; ++i) a[i].x = xr[xpad<C>(t + NT * i)];
      __syncthreads();
      if (VEC0 && st == 0 && S::PADW == 2) {
        double2 *row = (double2 *)(xr + xpad<C>(base));
#pragma unroll
        for (int k = 0; k < P; k += 2) row[k / 2] = make_double2(a[k].y, a[k + 1].y);
      } else {
#pragma unroll
        for (int k = 0; k < P; ++k) xr[xpad<C>(base + s * k)] = a[k].y;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < P; ++i) a[i].y = xr[xpad<C>(t + NT * i)];
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

/* thread 0: arm the barrier and start the bulk copies of one tile (TPB sequences of `bytes` each) */
template <int TPB, bool L2_AHEAD = false>
__device__ __forceinline__ void stream_issue(char *land, const char *gbase, long long lot, long long jump_bytes,
                                             long long tile, unsigned seq_bytes, uint64_t *bar) {
  const long long g0 = tile * TPB;
  const int live = (int)((lot - g0) < TPB ? (lot - g0) : TPB);
  mbar_expect_tx(bar, (unsigned)live * seq_bytes);
  for (int i = 0; i < live; ++i) bulk_g2s(land + (size_t)i * seq_bytes, gbase + (g0 + i) * jump_bytes, seq_bytes, bar);
  if (L2_AHEAD) {  // also pull the tile after this one into L2 (measured: real kernel +1.6 %, complex kernel -15 %)
    const long long g2 = (tile + gridDim.x) * TPB;
    for (int i = 0; i < TPB && g2 + i < lot; ++i) bulk_prefetch_l2(gbase + (g2 + i) * jump_bytes, seq_bytes);
  }
}

template <class C, int MINB, int DIR>
__global__ void __launch_bounds__(C::THREADS, MINB) pow2_c2c_stream_kernel(cpx *__restrict__ c, long long lot,
                                                                           long long jump, const cpx *__restrict__ tw,
                                                                           double scale, long long ntiles) {
  CFB_DYN_SMEM(smem_raw);
  constexpr int N = C::N, P = C::P, NT = C::NT;
  typedef StreamSmem<C> S;
  cpx *land = (cpx *)smem_raw;
  double *xch = (double *)(smem_raw + S::LAND);
  cpx *tws = (cpx *)(smem_raw + S::LAND + S::XCH);
  uint64_t *bar = (uint64_t *)(smem_raw + S::LAND + S::XCH + S::TWS);
  const int tid = threadIdx.x, tl = tid / NT, t = tid % NT;
  if (tid == 0) mbar_init(bar, 1);
  for (int i = tid; i < S::TWS_COUNT; i += C::THREADS) tws[i] = __ldg(tw + i);
  __syncthreads();
  long long tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) stream_issue<C::TPB, (C::N >= 8192)>((char *)land, (const char *)c, lot, jump * 16, tile, N * 16, bar);
  unsigned parity = 0;
  for (; tile < ntiles; tile += gridDim.x) {
    mbar_wait(bar, parity);
    parity ^= 1;
    const long long g = tile * C::TPB + tl;
    const bool live = g < lot;
    cpx a[P];
#pragma unroll
    for (int i = 0; i < P; ++i) a[i] = land[(size_t)tl * N + t + NT * i];
    landing_reads_done<P, true>(a, (volatile unsigned *)(bar + 1));
    // the landing buffer has been consumed: warp 0 waits for everybody's report and refills it with the next tile while
    // the other warps are already computing
    const long long next = tile + gridDim.x;
    if (tid < 32) {
      named_sync(1, C::THREADS);
      if (tid == 0 && next < ntiles) stream_issue<C::TPB, (C::N >= 8192)>((char *)land, (const char *)c, lot, jump * 16, next, N * 16, bar);
    } else {
      named_arrive(1, C::THREADS);
    }
    pow2_core_split<C, DIR>(a, xch + (size_t)tl * S::XTILE, t, tws);
    if (live) {
      cpx *x = c + g * jump + t;
#pragma unroll
      for (int i = 0; i < P; ++i) x[NT * i] = make_double2(a[i].x * scale, a[i].y * scale);
    }
  }
}

/* real pairs, streaming: the landing buffer holds the two rows x_a, x_b of each pair back to back */
/* BULK: finished rows are written into the (then idle) exchange tile in their final layout and drained by one
 * cp.async.bulk shared->global per row instead of per-thread global stores (the split epilogue with its lane shuffles and
 * divergent edge stores was 34 % of the kernel's stall samples, profiles/r2_ncu_r2c4096.txt) */
template <class C, int MINB, int DIR, bool BULK = false>
__global__ void __launch_bounds__(C::THREADS, MINB) pow2_r2c_stream_kernel(double *__restrict__ r, long long lot,
                                                                           long long jump, const cpx *__restrict__ tw,
                                                                           long long ntiles) {
  CFB_DYN_SMEM(smem_raw);
  constexpr int N = C::N, P = C::P, NT = C::NT;
  typedef StreamSmem<C> S;
  double *land = (double *)smem_raw;  // [TPB][2][N]
  double *xch = (double *)(smem_raw + S::LAND);
  cpx *tws = (cpx *)(smem_raw + S::LAND + S::XCH);
  uint64_t *bar = (uint64_t *)(smem_raw + S::LAND + S::XCH + S::TWS);
  const int tid = threadIdx.x, tl = tid / NT, t = tid % NT;
  if (tid == 0) mbar_init(bar, 1);
  for (int i = tid; i < S::TWS_COUNT; i += C::THREADS) tws[i] = __ldg(tw + i);
  __syncthreads();
  // a tile = TPB pairs = 2*TPB consecutive sequences; reuse stream_issue with a "sequence" = one real row
  constexpr int ROWS = 2 * C::TPB;
  long long tile = blockIdx.x;
  if (tid == 0 && tile < ntiles) stream_issue<ROWS, true>((char *)land, (const char *)r, lot, jump * 8, tile, N * 8, bar);
  unsigned parity = 0;
  const double *la = land + (size_t)(2 * tl) * N, *lb = la + N;
  double *xq = xch + (size_t)tl * S::XTILE;
  for (; tile < ntiles; tile += gridDim.x) {
    mbar_wait(bar, parity);
    parity ^= 1;
    const long long ga = 2 * (tile * C::TPB + tl), gb = ga + 1;
    const bool va = ga < lot, vb = gb < lot;
    double *xa = r + (va ? ga : 0) * jump, *xb = r + (vb ? gb : 0) * jump;
    cpx a[P];
    if (DIR < 0) {
#pragma unroll
      for (int i = 0; i < P; ++i) a[i] = make_double2(la[t + NT * i], vb ? lb[t + NT * i] : 0.0);
    } else {
      // Z[e] from the half-complex rows (rfftb1_ convention): e < N/2: (a1 + b2, b1 - a2); e > N/2 uses f = N - e
#pragma unroll
      for (int i = 0; i < P; ++i) {
        const int e = t + NT * i;
        if (e == 0) a[i] = make_double2(la[0], vb ? lb[0] : 0.0);
        else if (e == N / 2) a[i] = make_double2(la[N - 1], vb ? lb[N - 1] : 0.0);
        else {
          const int f = e < N / 2 ? e : N - e;
          double a1 = 0.5 * la[2 * f - 1], a2 = 0.5 * la[2 * f];
          double b1 = vb ? 0.5 * lb[2 * f - 1] : 0.0, b2 = vb ? 0.5 * lb[2 * f] : 0.0;
          a[i] = e < N / 2 ? make_double2(a1 + b2, b1 - a2) : make_double2(a1 - b2, b1 + a2);
        }
      }
    }
    landing_reads_done(a, (volatile unsigned *)(bar + 1));
    const long long next = tile + gridDim.x;
    if (BULK) {  // full barrier: the previous tile's rows must also have left the exchange tile before anybody writes it
      if (t == 0) bulk_wait_read();
      __syncthreads();
      if (tid == 0 && next < ntiles) stream_issue<ROWS, true>((char *)land, (const char *)r, lot, jump * 8, next, N * 8, bar);
    } else if (tid < 32) {
      named_sync(1, C::THREADS);
      if (tid == 0 && next < ntiles) stream_issue<ROWS, true>((char *)land, (const char *)r, lot, jump * 8, next, N * 8, bar);
    } else {
      named_arrive(1, C::THREADS);
    }
    pow2_core_split<C, DIR>(a, xq, t, tws);
    if (DIR < 0) {
      // separate X_a, X_b: Z[N-f] of the upper half goes through the exchange tile, viewed as N/2 complex slots
      cpx *zq = (cpx *)xq;
#pragma unroll
      for (int i = P / 2; i < P; ++i) zq[t + NT * (i - P / 2)] = a[i];
      __syncthreads();
      // FFTPACK's half-complex row is [X0, A1, B1, A2, B2, ..., A_{N/2}] with A_f = 2 Re X_f / N, B_f = -2 Im X_f / N
      // (rfftf1_ epilogue, fftpack.c:13818-13853).  16-byte aligned pairs are (B_f, A_{f+1}): A_{f+1} comes from the
      // next lane by shuffle, so that all but the warp-edge lanes issue full 16-byte stores.
      const double sc = 1.0 / (double)N;
      const int lane = tid & 31;
      cpx nyq = make_double2(0.0, 0.0);
      // three passes over the 8 slots so that the shared-memory loads, the arithmetic and the shuffles of different
      // slots overlap (one fused loop exposed each latency 8 times: 39 % of the stall samples, profiles/r1_ncu_r2c_src.txt).
      // The upper half of a[] is free after the exchange: it receives the partners, then the results of row b.
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        const int f = t + NT * i;
        a[P / 2 + i] = zq[f == 0 ? 0 : N / 2 - f];
      }
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        const int f = t + NT * i;
        const cpx u = a[i], v = a[P / 2 + i];
        if (f == 0) {  // slot 0 holds X0 itself; A_{N/2} = X_{N/2} / N closes the row
          nyq = make_double2(v.x * sc, v.y * sc);
          if (!BULK && va) xa[N - 1] = v.x * sc;
          if (!BULK && vb) xb[N - 1] = v.y * sc;
          a[i] = make_double2(0.0, u.x * sc);
          a[P / 2 + i] = make_double2(0.0, u.y * sc);
        } else {
          a[i] = make_double2((u.x + v.x) * sc, (v.y - u.y) * sc);          // (A, B) of row a
          a[P / 2 + i] = make_double2((u.y + v.y) * sc, (u.x - v.x) * sc);  // (A, B) of row b
        }
      }
      if (BULK) {
        // a[i] = (A_f, B_f) of row a, a[P/2 + i] of row b (slot f = 0: (-, X0/N); X_{N/2}/N already stored below)
#pragma unroll
        for (int row = 0; row < 2; ++row) {
          // row 0: every partner read of zq is done.  row 1: thread 0 gets here only after the bulk engine has read row 0.
          __syncthreads();
#pragma unroll
          for (int i = 0; i < P / 2; ++i) {
            const int f = t + NT * i;
            const cpx ab = a[row * (P / 2) + i];
            if (f == 0) xq[0] = ab.y;
            else {
              xq[2 * f - 1] = ab.x;
              xq[2 * f] = ab.y;
            }
          }
          if (t == 0) xq[N - 1] = row == 0 ? nyq.x : nyq.y;
          fence_async_smem();
          __syncthreads();
          if (t == 0 && (row == 0 ? va : vb)) {
            bulk_s2g(row == 0 ? xa : xb, xq, (unsigned)(N * sizeof(double)));
            bulk_commit();
            if (row == 0) bulk_wait_read();  // the tile is reused for row b right away
          }
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < P / 2; ++i) {
        const int f = t + NT * i;
        const double Aa = a[i].x, Ba = a[i].y, Ab = a[P / 2 + i].x, Bb = a[P / 2 + i].y;
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
      }
      // no barrier needed here: the next tile's first write to this exchange tile comes after its landing-read barrier,
      // which every thread only reaches once it has finished reading zq above
    } else if (BULK) {
      // pow2_core_split ends with a barrier after its last exchange read: the tile is free
#pragma unroll
      for (int row = 0; row < 2; ++row) {
        if (row == 1) {
          if (t == 0 && va) bulk_wait_read();
          __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < P; ++i) xq[t + NT * i] = row == 0 ? a[i].x : a[i].y;
        fence_async_smem();
        __syncthreads();
        if (t == 0 && (row == 0 ? va : vb)) {
          bulk_s2g(row == 0 ? xa : xb, xq, (unsigned)(N * sizeof(double)));
          bulk_commit();
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
  if (BULK && t == 0) bulk_wait_all();
}

/* ---------------------------------------------------------------------------------------------------------
 * Tile kernel for the four-step decomposition of long power-of-two transforms (cfft2f_ 16384^2, cfftmf_ N >= 2^13
 * with strides): one CTA transforms TPB short sequences ("rows", N = 64..1024) whose elements are strided in
 * memory.  Threads are laid out rows-fastest, so for the layouts the four-step driver produces (the row index is the
 * contiguous axis) every warp-wide load/store of one register slot covers consecutive rows = full 128-byte runs.
 * The one side of the decomposition whose rows are contiguous along the ELEMENT axis is staged through the
 * shared-memory tile instead (that is where the transposition of the four-step algorithm happens).
 * Twiddles W_n^(j e) of the split are applied on store from two short shared-memory tables.
 * --------------------------------------------------------------------------------------------------------- */
struct TileParams {
  const cpx *in;
  cpx *out;
  Addr ain, aout;
  long long lot;  // rows
  double scale;
  const cpx *tw;  // pow2_table of the row length
  const cpx *fs;  // split twiddle tables of the long length (nullptr: none)
  int fs_shift, fs_from_hi, fs_count;
  int fs_nmask;   // long length - 1 (a power of two)
  int in_staged;  // 1: rows are contiguous along the element axis on the input side
  // sharded 2-D transform: the transform axis of the OUTPUT is split over npeers GPUs in chunks of 2^peer_shift
  // elements; element e goes to peers[e >> peer_shift] (a peer-mapped pointer, NVLink store) at local index
  // e & (2^peer_shift - 1).  npeers = 0: everything goes to `out`.
  cpx *peers[16];
  int npeers, peer_shift;
  long long out_base;
  // long 1-D transforms N = L*Mm (six-step): the results of the length-Mm transforms (this sweep computes elements
  // b = tw2_k0(row) + tw2_n1 * e of sequence i) are multiplied by W_N^(i b) on store.  tw2 = RootPlan table of N in global
  // memory (2^tw2_shift low entries, then the high ones); the sequence index i is the row's lo (tw2_seq_lo) or hi part
  // plus tw2_off (first sequence of this GPU / chunk).  nullptr: none.
  const cpx *tw2;
  int tw2_shift, tw2_seq_lo, tw2_n1;
  long long tw2_mask, tw2_off;
};

/* the last step of every tile kernel: scale, optional twiddles, store to `out` or straight into the peers' slabs */
template <class C, int DIR>
__device__ __forceinline__ void tile_store(const TileParams &P, cpx (&a)[C::P], const long long g, const int t,
                                           const cpx *__restrict__ fss) {
  constexpr int PP = C::P, NT = C::NT;
  const long long hi = g / P.aout.nlo, lo = g - hi * P.aout.nlo;
  const long long oout = hi * P.aout.jump_hi + lo * P.aout.jump_lo;
  const double scale = P.scale;
#pragma unroll
  for (int i = 0; i < PP; ++i) a[i] = make_double2(a[i].x * scale, a[i].y * scale);
  if (P.fs_count > 0) {
    // four-step twiddles W_n^(j (t + NT i)) = base * step^i: base, step and step^4 from the split tables in shared
    // memory (six reads instead of two per element), the powers as in twiddle_powers
    const int j = (int)(P.fs_from_hi ? hi : lo);
    const int mask = (1 << P.fs_shift) - 1, sh = P.fs_shift, nmask = P.fs_nmask;
    auto root = [&](int x) { return cmul(fss[x & mask], fss[mask + 1 + (x >> sh)]); };
    // a[i] *= b0 w1^i in place: a table of the 16 products would cost 64 registers (it spilled: the sweeps' top stall
    // was long_scoreboard on local memory)
    const cpx b0 = root(j * t), w1 = root((j * NT) & nmask), w4 = root((4 * j * NT) & nmask);
#pragma unroll
    for (int i = 0; i < PP; ++i) a[i] = ctw<DIR>(a[i], b0);
    twiddle_powers<DIR>(a, w1, w4);
  }
  if (P.tw2) {
    const int seq_lo = P.tw2_seq_lo;
    const long long seq = (seq_lo ? lo : hi) + P.tw2_off, k0 = seq_lo ? hi : lo;
    const long long mask = P.tw2_mask, step = (seq * P.tw2_n1) & mask;
    const long long lmask = (1LL << P.tw2_shift) - 1;
    const int sh = P.tw2_shift;
    const cpx *tb = P.tw2;
    auto root = [&](long long x) { return cmul(__ldg(tb + (x & lmask)), __ldg(tb + lmask + 1 + (x >> sh))); };
    const cpx b0 = root((seq * k0 + step * t) & mask), w1 = root((step * NT) & mask), w4 = root((4 * step * NT) & mask);
#pragma unroll
    for (int i = 0; i < PP; ++i) a[i] = ctw<DIR>(a[i], b0);
    twiddle_powers<DIR>(a, w1, w4);
  }
  if (P.npeers > 0) {
    // fused transpose: each element is stored straight into the memory of the GPU that owns its slab
    const int emask = (1 << P.peer_shift) - 1;
#pragma unroll
    for (int i = 0; i < PP; ++i) {
      const int e = t + NT * i;
      P.peers[e >> P.peer_shift][P.out_base + oout + (long long)(e & emask) * P.aout.inc] = a[i];
    }
  } else {
    cpx *y = P.out + oout + (long long)t * P.aout.inc;
    const long long st = (long long)NT * P.aout.inc;
#pragma unroll
    for (int i = 0; i < PP; ++i) y[i * st] = a[i];
  }
}

template <class C>
struct TileSmem {
  static constexpr int PITCH = C::TILE | 1;  // odd pitch: rows-fastest threads hit different banks
  static constexpr size_t TILE_BYTES = (size_t)C::TPB * PITCH * sizeof(cpx);
  static constexpr size_t bytes(int fs_count) { return TILE_BYTES + (size_t)C::TPB * 8 + (size_t)fs_count * sizeof(cpx) + 16; }
};

__device__ __forceinline__ long long tile_batch_off(const Addr &a, long long g) {
  long long hi = g / a.nlo, lo = g - hi * a.nlo;
  return hi * a.jump_hi + lo * a.jump_lo;
}

template <class C, int DIR>
__global__ void __launch_bounds__(C::THREADS, 2) pow2_tile_kernel(const TileParams P) {
  CFB_DYN_SMEM(smem_raw);
  constexpr int N = C::N, PP = C::P, NT = C::NT, TPB = C::TPB, LP = C::LP, PITCH = TileSmem<C>::PITCH;
  cpx *tile = (cpx *)smem_raw;
  long long *rowoff = (long long *)(smem_raw + TileSmem<C>::TILE_BYTES);
  cpx *fss = (cpx *)(smem_raw + TileSmem<C>::TILE_BYTES + (size_t)TPB * 8);
  const int tid = threadIdx.x, tl = tid % TPB, t = tid / TPB;
  const long long g = (long long)blockIdx.x * TPB + tl;
  const bool live = g < P.lot;
  const long long oin = live ? tile_batch_off(P.ain, g) : 0;
  for (int i = tid; i < P.fs_count; i += C::THREADS) fss[i] = __ldg(P.fs + i);
  cpx *sm = tile + (size_t)tl * PITCH;
  cpx a[PP];
  if (!P.in_staged) {
    const cpx *x = P.in + oin + (long long)t * P.ain.inc;
    const long long st = (long long)NT * P.ain.inc;
#pragma unroll
    for (int i = 0; i < PP; ++i) a[i] = live ? x[i * st] : make_double2(0.0, 0.0);
    if (P.fs_count > 0) __syncthreads();  // twiddle tables visible
  } else {
    // rows contiguous along the element axis: cooperative row-major loads into the tile, then pick up registers
    if (t == 0) rowoff[tl] = live ? oin : -1;
    __syncthreads();
    constexpr int TOTAL = TPB * N;
    const long long inc = P.ain.inc;
#pragma unroll 1
    for (int base = 0; base < TOTAL; base += 4 * C::THREADS) {
      cpx v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * C::THREADS + tid;
        const int r = idx / N, e = idx % N;
        const long long o = idx < TOTAL ? rowoff[r] : -1;
        v[u] = o >= 0 ? P.in[o + e * inc] : make_double2(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * C::THREADS + tid;
        const int r = idx / N, e = idx % N;
        if (idx < TOTAL) tile[(size_t)r * PITCH + pad<LP>(e)] = v[u];
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PP; ++i) a[i] = sm[pad<LP>(t + NT * i)];
    __syncthreads();
  }
  pow2_core<C, DIR>(a, sm, t, P.tw);
  if (live) tile_store<C, DIR>(P, a, g, t, fss);
}

/* Streaming form of the tile kernel: persistent CTAs; the rows of tile k+1 are gathered with per-thread cp.async
 * (16 bytes per element, whichever thread order is coalesced for the layout) into a landing buffer while tile k is
 * transformed.  The landing buffer doubles as the transposition stage, so both row layouts take the same path. */
template <class C>
struct TileStreamSmem {
  typedef StreamSmem<C> S;
  static constexpr int LPITCH = C::N + 1;          // landing rows, complex; odd pitch for rows-fastest readers
  static constexpr int XPITCH = S::XTILE | 1;      // exchange rows, doubles
  static constexpr size_t LAND = (size_t)C::TPB * LPITCH * sizeof(cpx);
  static constexpr size_t XCH = (size_t)C::TPB * XPITCH * sizeof(double);
  static constexpr size_t TWS = S::TWS;
  static constexpr size_t bytes(int fs_count) {
    return LAND + XCH + TWS + (size_t)2 * C::TPB * 8 + (size_t)fs_count * sizeof(cpx) + 64;
  }
};

template <class C>
__device__ __forceinline__ void tile_issue_loads(const TileParams &P, cpx *land, long long *rowoff, long long tile,
                                                 int tid) {
  constexpr int N = C::N, NT = C::NT, TPB = C::TPB, PP = C::P, LPITCH = TileStreamSmem<C>::LPITCH;
  if (!P.in_staged) {  // rows are the contiguous axis: thread (row tl, slot t) fetches its own P elements
    const int tl = tid % TPB, t = tid / TPB;
    const long long g = tile * TPB + tl;
    if (g < P.lot) {
      const cpx *x = P.in + tile_batch_off(P.ain, g) + (long long)t * P.ain.inc;
      const long long st = (long long)NT * P.ain.inc;
      cpx *row = land + (size_t)tl * LPITCH + t;
#pragma unroll
      for (int i = 0; i < PP; ++i) cp_async16(row + NT * i, x + i * st);
    }
  } else {  // rows are contiguous along the element axis: consecutive threads walk one row
    for (int idx = tid; idx < TPB * N; idx += C::THREADS) {
      const int r = idx / N, e = idx % N;
      const long long o = rowoff[r];
      if (o >= 0) cp_async16(land + (size_t)r * LPITCH + e, P.in + o + (long long)e * P.ain.inc);
    }
  }
  cp_async_commit();
}

template <class C, int DIR>
__global__ void __launch_bounds__(C::THREADS, (C::THREADS > 256 ? 1 : 2)) pow2_tile_stream_kernel(const TileParams P, long long ntiles) {
  CFB_DYN_SMEM(smem_raw);
  typedef TileStreamSmem<C> TS;
  typedef StreamSmem<C> S;
  constexpr int PP = C::P, NT = C::NT, TPB = C::TPB;
  cpx *land = (cpx *)smem_raw;
  double *xch = (double *)(smem_raw + TS::LAND);
  cpx *tws = (cpx *)(smem_raw + TS::LAND + TS::XCH);
  long long *rowoff = (long long *)(smem_raw + TS::LAND + TS::XCH + TS::TWS);  // [2][TPB] (staged input only)
  cpx *fss = (cpx *)(rowoff + 2 * TPB);
  const int tid = threadIdx.x, tl = tid % TPB, t = tid / TPB;
  for (int i = tid; i < S::TWS_COUNT; i += C::THREADS) tws[i] = __ldg(P.tw + i);
  for (int i = tid; i < P.fs_count; i += C::THREADS) fss[i] = __ldg(P.fs + i);
  auto fill_rowoff = [&](long long tile, int slot) {
    if (P.in_staged && tid < TPB) {
      const long long g = tile * TPB + tid;
      rowoff[slot * TPB + tid] = (tile < ntiles && g < P.lot) ? tile_batch_off(P.ain, g) : -1;
    }
  };
  long long tile = blockIdx.x;
  fill_rowoff(tile, 0);
  __syncthreads();
  if (tile < ntiles) tile_issue_loads<C>(P, land, rowoff, tile, tid);
  else cp_async_commit();
  double *xr = xch + (size_t)tl * TS::XPITCH;
  const cpx *lrow = land + (size_t)tl * TS::LPITCH;
  int it = 0;
  for (; tile < ntiles; tile += gridDim.x, ++it) {
    const long long g = tile * TPB + tl;
    const bool live = g < P.lot;
    cp_async_wait<0>();
    __syncthreads();  // the tile has landed for every thread
    cpx a[PP];
#pragma unroll
    for (int i = 0; i < PP; ++i) a[i] = lrow[t + NT * i];
    const long long next = tile + gridDim.x;
    fill_rowoff(next, (it + 1) & 1);
    __syncthreads();  // landing buffer consumed (and next row offsets visible): refill it while we compute
    if (next < ntiles) tile_issue_loads<C>(P, land, rowoff + ((it + 1) & 1) * TPB, next, tid);
    else cp_async_commit();
    pow2_core_split<C, DIR, false>(a, xr, t, tws);
    if (live) tile_store<C, DIR>(P, a, g, t, fss);
  }
  cp_async_wait<0>();
}

/* TMA form of the streaming tile kernel for the layouts whose ROWS are the contiguous axis (three of the four sweeps of
 * cfft2f_, all sweeps of batch-contiguous layouts): the whole strided tile -- TPB adjacent rows x N elements -- is one
 * 3-D tensor box, so a single cp.async.bulk.tensor instruction (UTMALDG) per tile replaces 16 LDGSTS per thread.
 * The box lands dense as [N][TPB] complex; thread (row tl, slot t) reads element e at land[e*TPB + tl]. */
template <class C, bool STAGED>
struct TileTmaSmem {
  typedef StreamSmem<C> S;
  static constexpr int XPITCH = S::XTILE | 1;
  static constexpr int LPITCH = C::N + 1;  // STAGED: one padded landing row per sequence
  static constexpr size_t LAND = (size_t)C::TPB * (STAGED ? LPITCH : C::N) * sizeof(cpx);
  static constexpr size_t XCH = (size_t)C::TPB * XPITCH * sizeof(double);
  static constexpr size_t bytes(int fs_count) { return 128 + LAND + XCH + S::TWS + 16 + (size_t)fs_count * sizeof(cpx) + 64; }
};

template <class C, int DIR, bool STAGED>
__global__ void __launch_bounds__(C::THREADS, (C::THREADS <= 128 ? 4 : 2)) pow2_tile_tma_kernel(const TileParams P, const CFB_GRID_CONSTANT TensorMap3 tmap,
                                                                      long long ntiles) {
  CFB_DYN_SMEM(smem_raw);
  typedef TileTmaSmem<C, STAGED> TS;
  typedef StreamSmem<C> S;
  constexpr int PP = C::P, NT = C::NT, TPB = C::TPB, N = C::N;
  // the tensor box must land 128-byte aligned.  The offset is added to the shared array itself: rounding the pointer
  // through an integer made the compiler lose the address space, and every exchange access became a generic LD/ST
  // (the sweeps' top stalls were lg_throttle / long_scoreboard, profiles/r2_ncu_tile_cfft2.txt)
#ifdef CFB_SIM
  char *base = (char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
#else
  char *base = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
#endif
  cpx *land = (cpx *)base;
  double *xch = (double *)(base + TS::LAND);
  cpx *tws = (cpx *)(base + TS::LAND + TS::XCH);
  uint64_t *bar = (uint64_t *)(base + TS::LAND + TS::XCH + S::TWS);
  cpx *fss = (cpx *)(bar + 2);
  const int tid = threadIdx.x, tl = tid % TPB, t = tid / TPB;
  if (tid == 0) mbar_init(bar, 1);
  for (int i = tid; i < S::TWS_COUNT; i += C::THREADS) tws[i] = __ldg(P.tw + i);
  for (int i = tid; i < P.fs_count; i += C::THREADS) fss[i] = __ldg(P.fs + i);
  __syncthreads();
  const int nlo = P.ain.nlo;
  auto issue = [&](long long tile) {  // warp 0
    const long long g0 = tile * TPB;
    if (!STAGED) {
      if (tid == 0) {
        const long long hi = g0 / nlo, lo = g0 - hi * nlo;
        mbar_expect_tx(bar, (unsigned)(TPB * N * sizeof(cpx)));
        tma_load_3d(land, &tmap, bar, (int)(2 * lo), 0, (int)hi);
      }
    } else {  // every sequence is one contiguous run: a 1-D bulk copy per row into its padded landing row
      const long long rows = P.lot - g0 < TPB ? P.lot - g0 : TPB;
      if (tid == 0) mbar_expect_tx(bar, (unsigned)(rows * N * sizeof(cpx)));
      __syncwarp();
      for (int r = tid; r < rows; r += 32)
        bulk_g2s(land + (size_t)r * TS::LPITCH, P.in + tile_batch_off(P.ain, g0 + r), (unsigned)(N * sizeof(cpx)), bar);
    }
  };
  long long tile = blockIdx.x;
  if (tid < 32 && tile < ntiles) issue(tile);
  unsigned parity = 0;
  double *xr = xch + (size_t)tl * TS::XPITCH;
  for (; tile < ntiles; tile += gridDim.x) {
    mbar_wait(bar, parity);
    parity ^= 1;
    const long long g = tile * TPB + tl;
    const bool live = g < P.lot;
    cpx a[PP];
#pragma unroll
    for (int i = 0; i < PP; ++i) a[i] = STAGED ? land[(size_t)tl * TS::LPITCH + t + NT * i] : land[(size_t)(t + NT * i) * TPB + tl];
    landing_reads_done<PP, true>(a, (volatile unsigned *)(bar + 1));
    __syncthreads();  // landing buffer consumed: refill it while we compute (split barrier: no gain here)
    const long long next = tile + gridDim.x;
    if (tid < 32 && next < ntiles) issue(next);
    pow2_core_split<C, DIR, false>(a, xr, t, tws);
    if (live) tile_store<C, DIR>(P, a, g, t, fss);
  }
}

/* ---- host side ---- */
bool pow2_c2c_supported(int n, long long inc, long long jump, int aligned16);
bool pow2_r2c_supported(int n, long long inc, long long jump, int aligned16);
bool pow2_c2c_launch(int n, long long lot, long long jump, int dir, cpx *c, double scale);
bool pow2_r2c_launch(int n, long long lot, long long jump, int dir, double *r);
/* four-step rows: log2 of the row length must be within [pow2_tile_min_log2, pow2_tile_max_log2] */
int pow2_tile_min_log2();
int pow2_tile_max_log2();
bool pow2_tile_launch(int log2n, int dir, TileParams &P);
void pow2_release_tables();

}  // namespace cfb
#endif
