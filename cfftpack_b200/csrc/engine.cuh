/*
 * engine.cuh -- the general shared-memory transform engine (any length, any stride).
 *
 * One CTA keeps T complex sequences of length M in shared memory, runs every radix pass of the
 * transform there (ping-pong between two buffers, one __syncthreads per pass) and touches global
 * memory exactly once on the way in and once on the way out.  It is the GPU counterpart of the
 * reference's pass drivers c1fm1f_/cmfm1f_ (cfftpack/fftpack.c:2041, :5262), rfftf1_/mrftf1_
 * (:13695, :10149) and of the pre/post loops of costf1_ (:6294), sintf1_ (:14828), cosqf1_ (:5665)
 * and their *b1_ / m* twins -- but not a translation of them:
 *
 *  - passes are decimation-in-frequency Stockham autosort passes: every pass reads
 *    x[b + (M/r) j] (unit stride across threads) and the last one leaves natural order;
 *  - real transforms are done two sequences at a time as ONE complex transform
 *    (z = x_a + i x_b), then separated through the Hermitian symmetry; this replaces the
 *    r1f*kf_/mradf* half-complex passes and works for odd lengths too;
 *  - cost/sint/cosq/sinq fold their pre/post-processing around that core inside the same CTA;
 *    the serial dsum recurrences of the reference (:6393-6400, :14909-14915) become warp scans;
 *  - lot/jump/inc (and a second batch level used by the four-step decomposition of long
 *    transforms) are resolved in the loader/storer, which walk whichever of the element axis or
 *    the batch axis is contiguous in memory so that global accesses stay coalesced.
 */
#ifndef CFB_ENGINE_CUH
#define CFB_ENGINE_CUH
#include "butterfly.cuh"
#include "engine_types.h"
#include "tma.cuh"

namespace cfb {

/* ------------------------------------------------------------------------------------------ */
__device__ __forceinline__ int fast_div(int n, int d, unsigned mag) { return d == 1 ? n : (int)__umulhi((unsigned)n, mag); }
/* padded position of element e inside a shared-memory row: one extra slot every 2^ps elements (ps = 31: none).
 * With ps = log2 of the first radix the scattered stores of the first pass are bank-conflict free. */
__device__ __forceinline__ int padx(int e, int ps) { return e + (e >> ps); }
/* compile-time switch: rows without padding skip the shift/add altogether */
template <bool PAD>
__device__ __forceinline__ int padq(int e, int ps) { return PAD ? e + (e >> ps) : e; }

/* one radix-R pass over T sequences: src, dst are [T][ldz] complex arrays in shared memory       */
template <int R, int DIR, bool PAD>
__device__ __forceinline__ void pass_fixed(const cpx *__restrict__ src, cpx *__restrict__ dst, int T, int ldz, int M,
                                           const PassDesc &pd, const cpx *__restrict__ tw, int tid, int nthr, int ps) {
  const int nb = M / R;  // butterflies per sequence
  const int total = nb * T;
  const int s = pd.s, m = pd.m;
  const cpx *twp = tw + pd.twoff;
  for (int idx = tid; idx < total; idx += nthr) {
    int t = fast_div(idx, nb, pd.mag_nb), b = idx - t * nb;
    int p = fast_div(b, s, pd.mag_s), q = b - p * s;
    cpx a[R];
    const cpx *sp = src + t * ldz;
#pragma unroll
    for (int j = 0; j < R; ++j) a[j] = sp[padq<PAD>(b + j * nb, ps)];
    DftRt<R, DIR>::run(a, tw + pd.rtoff);
    cpx *dp = dst + t * ldz;
    const int o0 = q + s * R * p;
    dp[padq<PAD>(o0, ps)] = a[0];
    if (m > 1) {
#pragma unroll
      for (int k = 1; k < R; ++k) dp[padq<PAD>(o0 + k * s, ps)] = ctw<DIR>(a[k], twp[(k - 1) * m + p]);
    } else {
#pragma unroll
      for (int k = 1; k < R; ++k) dp[padq<PAD>(o0 + k * s, ps)] = a[k];
    }
  }
}

/* generic odd radix r (the reference's c1fgkf_/c1fgkb_, fftpack.c:1650/:1410): O(r) work per output from the symmetric
 * sums.  A work item is one butterfly x a block of CFB_GENERIC_KB output pairs (k, r-k): the inputs a_j, a_{r-j} are
 * read from shared memory once per block instead of once per pair, which is what bounds this pass for large r. */
template <int DIR, bool PAD>
__device__ __forceinline__ void pass_generic(const cpx *__restrict__ src, cpx *__restrict__ dst, int T, int ldz, int M,
                                             const PassDesc &pd, const cpx *__restrict__ tw, int tid, int nthr, int ps) {
  constexpr int KB = CFB_GENERIC_KB;
  const int r = pd.radix, nb = M / r, s = pd.s, m = pd.m, half = (r + 1) / 2;
  const int per = nb * (1 + (half - 1 + KB - 1) / KB);  // items per sequence: (k = 0 | k-blocks) x butterflies, butterflies fastest
  const int total = per * T;
  const cpx *twp = tw + pd.twoff;
  const cpx *rt = tw + pd.rtoff;
  for (int idx = tid; idx < total; idx += nthr) {
    int t = fast_div(idx, per, pd.mag_per), rem = idx - t * per;
    int kb = fast_div(rem, nb, pd.mag_nb), b = rem - kb * nb;
    int p = fast_div(b, s, pd.mag_s), q = b - p * s;
    const cpx *sp = src + t * ldz;
    cpx *dp = dst + t * ldz;
    const int o0 = q + s * r * p;
    cpx a0 = sp[padq<PAD>(b, ps)];
    if (kb == 0) {
      cpx acc = a0;
      for (int j = 1; j < r; ++j) acc = cadd(acc, sp[padq<PAD>(b + j * nb, ps)]);
      dp[padq<PAD>(o0, ps)] = acc;
    } else {
      // X_k = a0 + sum_{j=1}^{half-1} [ c_jk (a_j + a_{r-j}) + DIR*i * s_jk (a_j - a_{r-j}) ],  X_{r-k}: minus sign
      const int k0 = 1 + KB * (kb - 1);
      double ar[KB], ai[KB], br[KB], bi[KB];
      int kk[KB], jk[KB];
#pragma unroll
      for (int c = 0; c < KB; ++c) {
        kk[c] = k0 + c < half ? k0 + c : half - 1;  // the tail block recomputes its last pair (not stored twice)
        jk[c] = 0;
        ar[c] = a0.x;
        ai[c] = a0.y;
        br[c] = 0.0;
        bi[c] = 0.0;
      }
      for (int j = 1; j < half; ++j) {
        const cpx u = sp[padq<PAD>(b + j * nb, ps)], v = sp[padq<PAD>(b + (r - j) * nb, ps)];
        const double pr = u.x + v.x, pi = u.y + v.y, mr = u.x - v.x, mi = u.y - v.y;
#pragma unroll
        for (int c = 0; c < KB; ++c) {
          jk[c] += kk[c];
          if (jk[c] >= r) jk[c] -= r;
          const cpx w = rt[jk[c]];  // (cos, -sin)(2 pi jk / r)
          ar[c] = fma(w.x, pr, ar[c]);
          ai[c] = fma(w.x, pi, ai[c]);
          br[c] = fma(-w.y, mr, br[c]);  // sin * (a_j - a_{r-j})
          bi[c] = fma(-w.y, mi, bi[c]);
        }
      }
#pragma unroll
      for (int c = 0; c < KB; ++c) {
        const int k = k0 + c;
        if (k < half) {
          // DIR*i*(br + i bi) = DIR*(-bi + i br)
          cpx xk, xc;
          if (DIR < 0) {
            xk = make_double2(ar[c] + bi[c], ai[c] - br[c]);
            xc = make_double2(ar[c] - bi[c], ai[c] + br[c]);
          } else {
            xk = make_double2(ar[c] - bi[c], ai[c] + br[c]);
            xc = make_double2(ar[c] + bi[c], ai[c] - br[c]);
          }
          if (m > 1) {
            xk = ctw<DIR>(xk, twp[(k - 1) * m + p]);
            xc = ctw<DIR>(xc, twp[(r - k - 1) * m + p]);
          }
          dp[padq<PAD>(o0 + k * s, ps)] = xk;
          dp[padq<PAD>(o0 + (r - k) * s, ps)] = xc;
        }
      }
    }
  }
}

template <int DIR, bool PAD>
__device__ __forceinline__ void run_passes_impl(cpx *&cur, cpx *&oth, const EngineParams &P, const cpx *__restrict__ tw,
                                                int tid, int nthr) {
  const int ps = P.padshift;
  for (int ip = 0; ip < P.nf; ++ip) {
    const PassDesc &pd = P.pass[ip];
    switch (pd.radix) {
      case 2: pass_fixed<2, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 3: pass_fixed<3, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 4: pass_fixed<4, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 5: pass_fixed<5, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 6: pass_fixed<6, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 7: pass_fixed<7, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 8: pass_fixed<8, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 9: pass_fixed<9, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 10: pass_fixed<10, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 11: pass_fixed<11, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      case 13: pass_fixed<13, DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
      default: pass_generic<DIR, PAD>(cur, oth, P.T, P.ldz, P.M, pd, tw, tid, nthr, ps); break;
    }
    __syncthreads();
    cpx *t = cur;
    cur = oth;
    oth = t;
  }
}
/* PAD = false is the only variant the real-family kernel needs (its rows are never padded) */
template <int DIR, bool ALLOW_PAD = true>
__device__ __forceinline__ void run_passes(cpx *&cur, cpx *&oth, const EngineParams &P, const cpx *__restrict__ tw,
                                           int tid, int nthr) {
  if (ALLOW_PAD && P.padshift != 31) run_passes_impl<DIR, true>(cur, oth, P, tw, tid, nthr);
  else run_passes_impl<DIR, false>(cur, oth, P, tw, tid, nthr);
}

__device__ __forceinline__ long long batch_off(const Addr &a, long long g) {
  if (a.jump_hi == 0) return g * a.jump_lo;  // single-level batch (no 64-bit division)
  long long hi = g / a.nlo, lo = g - hi * a.nlo;
  return hi * a.jump_hi + lo * a.jump_lo;
}

/* ------------------------------------------------------------------------------------------ */
/* forward real core, output side: Z = FFT(x_a + i x_b) (unscaled) -> FFTPACK half-complex rows ha, hb
 * scaled like rfftf1_ (fftpack.c:13818-13853): h[0]=X0/M, h[2f-1]=2Re X_f/M, h[2f]=-2Im X_f/M, h[M-1]=X_{M/2}/M */
__device__ __forceinline__ void split_pair(const cpx *__restrict__ z, double *__restrict__ ha, double *__restrict__ hb,
                                           int M, int f) {
  const double sc = 1.0 / (double)M;
  if (f == 0) {
    cpx z0 = z[0];
    ha[0] = z0.x * sc;
    hb[0] = z0.y * sc;
  } else if (2 * f < M) {
    cpx u = z[f], v = z[M - f];
    ha[2 * f - 1] = (u.x + v.x) * sc;
    ha[2 * f] = -(u.y - v.y) * sc;
    hb[2 * f - 1] = (u.y + v.y) * sc;
    hb[2 * f] = (u.x - v.x) * sc;
  } else if (2 * f == M) {
    cpx u = z[f];
    ha[M - 1] = u.x * sc;
    hb[M - 1] = u.y * sc;
  }
}

/* backward real core, input side: half-complex rows ha, hb -> Z with z = x_a + i x_b after the backward FFT
 * (rfftb1_, fftpack.c:13517: x_t = h0 + sum h[2f-1] cos + h[2f] sin (+ (-1)^t h[M-1])) */
__device__ __forceinline__ void build_pair(cpx *__restrict__ z, const double *__restrict__ ha,
                                           const double *__restrict__ hb, int M, int f) {
  if (f == 0) {
    z[0] = make_double2(ha[0], hb[0]);
  } else if (2 * f < M) {
    double a1 = 0.5 * ha[2 * f - 1], a2 = 0.5 * ha[2 * f], b1 = 0.5 * hb[2 * f - 1], b2 = 0.5 * hb[2 * f];
    z[f] = make_double2(a1 + b2, b1 - a2);
    z[M - f] = make_double2(a1 - b2, b1 + a2);
  } else if (2 * f == M) {
    z[f] = make_double2(ha[M - 1], hb[M - 1]);
  }
}

/* ---- kind-specific pre/post-processing of ONE real sequence by a GROUP of gs threads (gs = 32, 64, 128 or 256;
 * gl = this thread's index in the group).  Groups of more than one warp combine their partial sums / scan carries
 * through gsc (8 doubles of shared memory per row) and a block barrier, so every thread of the block must make the
 * call (the engine runs all groups in lockstep; the long-sequence kernels use gs = 32, where no barrier is needed). */
__device__ __forceinline__ double group_sum(double v, int gl, int gs, double *gsc) {
  v = warp_sum(v);
  if (gs > 32) {
    if ((gl & 31) == 0) gsc[gl >> 5] = v;
    __syncthreads();
    v = 0.0;
    for (int w = 0; w < (gs >> 5); ++w) v += gsc[w];
  }
  return v;
}
/* x: the loaded sequence (length n, unit stride).  Forward-core kinds write u (length M) into the re or im
 * lane of the complex row (zc, stride 2).  Backward-core kinds write the half-complex row h (unit stride). */
__device__ __forceinline__ void pre_forward_core(int kind, int dir, int n, int M, const double *__restrict__ x,
                                                 double *__restrict__ zc, const double *__restrict__ trig, double *dsum,
                                                 int gl, int gs, double *gsc) {
  if (kind == K_RFFT) {
    for (int j = gl; j < n; j += gs) zc[2 * j] = x[j];
  } else if (kind == K_COST) {
    // costf1_/costb1_ pre-fold (fftpack.c:6355-6377, :6222-6244); trig[j] = 2 sin(j pi/M), trig[M + j] = 2 cos(j pi/M)
    const int ns2 = n / 2;
    const double e = (dir > 0) ? 2.0 : 1.0;  // backward doubles the end points first
    double part = 0.0;
    for (int j = 1 + gl; j < ns2; j += gs) {
      int jc = n - 1 - j;
      double t1 = x[j] + x[jc], t2 = x[j] - x[jc];
      part = fma(trig[M + j], t2, part);
      t2 = trig[j] * t2;
      zc[2 * j] = t1 - t2;
      if (jc < M) zc[2 * jc] = t1 + t2;
    }
    part = group_sum(part, gl, gs, gsc);
    if (gl == 0) {
      double x0 = e * x[0], xn = e * x[n - 1];
      *dsum = (x0 - xn) + part;
      zc[0] = x0 + xn;
      if (n & 1) zc[2 * ns2] = x[ns2] + x[ns2];
    }
  } else if (kind == K_SINT) {
    // sintf1_ pre (fftpack.c:14873-14888); trig[k-1] = 2 sin(k pi/(n+1))
    const int ns2 = n / 2;
    for (int k = 1 + gl; k <= ns2; k += gs) {
      int kc = n + 1 - k;
      double t1 = x[k - 1] - x[kc - 1], t2 = trig[k - 1] * (x[k - 1] + x[kc - 1]);
      zc[2 * k] = t1 + t2;
      zc[2 * kc] = t2 - t1;
    }
    if (gl == 0) {
      zc[0] = 0.0;
      if (n & 1) zc[2 * (ns2 + 1)] = 4.0 * x[ns2];
    }
  } else {  // K_COSQ / K_SINQ forward: cosqf1_ pre (fftpack.c:5693-5717); trig[i] = cos((i+1) pi/(2n))
    const int ns2 = (n + 1) / 2;
    for (int j = 1 + gl; j < ns2; j += gs) {
      int jc = n - j;
      double a = x[j] + x[jc], b = x[j] - x[jc];
      zc[2 * j] = fma(trig[j - 1], b, trig[jc - 1] * a);
      zc[2 * jc] = fma(trig[j - 1], a, -(trig[jc - 1] * b));
    }
    if (gl == 0) {
      zc[0] = x[0];
      if (!(n & 1)) zc[2 * ns2] = trig[ns2 - 1] * (x[ns2] + x[ns2]);
    }
  }
}

__device__ __forceinline__ void pre_backward_core(int kind, int n, const double *__restrict__ x, double *__restrict__ h,
                                                  int gl, int gs) {
  if (kind == K_RFFT) {
    for (int j = gl; j < n; j += gs) h[j] = x[j];
  } else {  // cosqb1_ pre (fftpack.c:5604-5616); sinqb1_ first negates the odd entries (fftpack.c:14160-14166)
    const double so = (kind == K_SINQ) ? -1.0 : 1.0;
    for (int i0 = 2 + 2 * gl; i0 < n; i0 += 2 * gs) {
      double a = so * x[i0 - 1], b = x[i0];
      h[i0 - 1] = 0.5 * (a + b);
      h[i0] = 0.5 * (a - b);
    }
    if (gl == 0) {
      h[0] = 0.5 * x[0];
      if (!(n & 1)) h[n - 1] = 0.5 * so * x[n - 1];
    }
  }
}

/* s -> y (both unit stride).  s is the half-complex row h (forward core, length M) or the real sequence u
 * (backward core). */
__device__ __forceinline__ void post_sequence(int kind, int dir, int n, int M, const double *__restrict__ s,
                                              double *__restrict__ y, const double *__restrict__ trig, double dsum,
                                              int gl, int gs, double *gsc) {
  if (kind == K_RFFT) {
    for (int j = gl; j < n; j += gs) y[j] = s[j];
  } else if (kind == K_COST) {
    // costf1_ post (fftpack.c:6386-6407) / costb1_ post (:6253-6283):
    //   y[0] = c0 h[0]; y[2m] = c1 h'[2m-1]; y[2m-1] = D + sum_{m'<m} c1 h'[2m'];  h' = h with h[M-1] doubled if M even
    const double c0 = dir < 0 ? 0.5 : 0.5 * (double)M, c1 = dir < 0 ? 0.5 : 0.25 * (double)M;
    const double D = dir < 0 ? dsum / (double)M : 0.5 * dsum;
    const int last = (M % 2 == 0) ? M - 1 : -1;
    const int cnt = n / 2;  // odd output indices 2m-1, m = 1..cnt
    // exclusive prefix over m of v(m) = c1 h'[2m]: each warp of the group owns a contiguous segment of m and walks it
    // 32 at a time (lane-interleaved: conflict-free shared-memory access); segment totals give the carry-in
    auto val = [&](int mm) -> double {
      const int i = 2 * mm;
      return (mm <= cnt && i < M) ? c1 * (i == last ? 2.0 * s[i] : s[i]) : 0.0;
    };
    const int nw = gs >> 5, w = gl >> 5, lane = gl & 31;
    const int seg = ((cnt + nw - 1) / nw + 31) & ~31;  // segment length per warp, a multiple of 32
    const int m0 = 1 + w * seg;
    double carry = D;
    if (nw > 1) {
      double tot = 0.0;
      for (int mm = m0 + lane; mm < m0 + seg; mm += 32) tot += val(mm);
      tot = warp_sum(tot);
      if (lane == 0) gsc[w] = tot;
      __syncthreads();
      for (int w2 = 0; w2 < w; ++w2) carry += gsc[w2];
    }
    for (int base = m0; base < m0 + seg && base <= cnt; base += 32) {
      const int mm = base + lane;
      const double v = val(mm);
      const double ex = warp_excl_scan(v, lane);
      if (mm <= cnt) {
        const int i = 2 * mm;
        y[i - 1] = carry + ex;
        if (i < n) y[i] = c1 * ((i - 1) == last ? 2.0 * s[i - 1] : s[i - 1]);
        if (dir < 0 && mm == cnt) y[n - 1] *= 0.5;  // y[n-1] is y[2cnt-1] (n even) or y[2cnt] (n odd): mine
      }
      carry += __shfl_sync(0xffffffffu, ex + v, 31);
    }
    if (gl == 0) y[0] = c0 * s[0];
  } else if (kind == K_SINT) {
    // sintf1_ post (fftpack.c:14898-14919): y[2m] = sc h[0] + sum_{m'<=m} sc h[2m'-1]; y[2m-1] = sc h[2m]
    const double sc = dir < 0 ? 0.5 : 0.25 * (double)M;
    const int cnt = (n - 1) / 2;  // even output indices 2m, m = 1..cnt
    auto val = [&](int mm) -> double { return mm <= cnt ? sc * s[2 * mm - 1] : 0.0; };
    const int nw = gs >> 5, w = gl >> 5, lane = gl & 31;
    const int seg = ((cnt + nw - 1) / nw + 31) & ~31;
    const int m0 = 1 + w * seg;
    double carry = sc * s[0];
    if (nw > 1) {
      double tot = 0.0;
      for (int mm = m0 + lane; mm < m0 + seg; mm += 32) tot += val(mm);
      tot = warp_sum(tot);
      if (lane == 0) gsc[w] = tot;
      __syncthreads();
      for (int w2 = 0; w2 < w; ++w2) carry += gsc[w2];
    }
    for (int base = m0; base < m0 + seg && base <= cnt; base += 32) {
      const int mm = base + lane;
      const double v = val(mm);
      const double ex = warp_excl_scan(v, lane);
      if (mm <= cnt) {
        y[2 * mm] = carry + ex + v;
        y[2 * mm - 1] = sc * s[2 * mm];
      }
      carry += __shfl_sync(0xffffffffu, ex + v, 31);
    }
    if (gl == 0) {
      y[0] = sc * s[0];
      if (!(n & 1)) y[n - 1] = sc * s[n];
    }
  } else if (dir < 0) {  // cosqf1_ post (fftpack.c:5731-5738)
    for (int i0 = 2 + 2 * gl; i0 < n; i0 += 2 * gs) {
      double a = s[i0 - 1], b = s[i0];
      y[i0 - 1] = 0.5 * (a + b);
      y[i0] = 0.5 * (a - b);
    }
    if (gl == 0) {
      y[0] = s[0];
      if (!(n & 1)) y[n - 1] = s[n - 1];
    }
  } else {  // cosqb1_ post (fftpack.c:5625-5652)
    const int ns2 = (n + 1) / 2;
    for (int j = 1 + gl; j < ns2; j += gs) {
      int jc = n - j;
      double p = fma(trig[j - 1], s[jc], trig[jc - 1] * s[j]);
      double q = fma(trig[j - 1], s[j], -(trig[jc - 1] * s[jc]));
      y[j] = p + q;
      y[jc] = p - q;
    }
    if (gl == 0) {
      y[0] = s[0] + s[0];
      if (!(n & 1)) y[ns2] = trig[ns2 - 1] * (s[ns2] + s[ns2]);
    }
  }
}

/* complex sequences, software-pipelined: while tile k is transformed, tile k+1 is gathered from global memory into
 * a second landing buffer with per-thread asynchronous copies (cp.async / LDGSTS), so the long global-load latency
 * is off the critical path.  Buffers: L[0], L[1] (landing, alternate) and W (ping-pong partner of the passes).
 * Loader and storer keep a running global pointer per thread (no per-element 64-bit multiplies or table reads). */
struct TileWalk {
  int rs, rn, es, en;  // this thread's first row / row step, first element / element step
};
__device__ __forceinline__ TileWalk tile_walk(int tid, int nthr, int txl, int lanes_t) {
  const int tx = tid & ((1 << txl) - 1), ty = tid >> txl, nx = 1 << txl, ny = nthr >> txl;
  TileWalk w;
  if (lanes_t) {  // consecutive threads walk the rows (batch axis contiguous in memory)
    w.rs = tx; w.rn = nx; w.es = ty; w.en = ny;
  } else {        // consecutive threads walk the elements
    w.rs = ty; w.rn = ny; w.es = tx; w.en = nx;
  }
  return w;
}

__device__ __forceinline__ void c2c_issue_loads(const EngineParams &P, cpx *L, const long long *off_in, int tid, int nthr) {
  const cpx *in = (const cpx *)P.in;
  const int T = P.T, M = P.M, ldz = P.ldz, ps = P.padshift, al = P.aligned16;
  const TileWalk w = tile_walk(tid, nthr, P.tx_in_log2, P.ain.lanes_t);
  const long long step = (long long)w.en * P.ain.inc;
  for (int r = w.rs; r < T; r += w.rn) {
    const long long o = off_in[r];
    if (o < 0) continue;
    const cpx *p = in + o + (long long)w.es * P.ain.inc;
    cpx *row = L + r * ldz;
    if (al) {
      for (int e = w.es; e < M; e += w.en, p += step) cp_async16(row + padx(e, ps), p);
    } else {
      for (int e = w.es; e < M; e += w.en, p += step) {
        cpx *d = row + padx(e, ps);
        cp_async8(d, p);
        cp_async8((double *)d + 1, (const double *)p + 1);
      }
    }
  }
  cp_async_commit();
}

#define CFB_FS_SMEM_MAX 512  /* four-step twiddle tables up to this many entries are kept in shared memory */

__global__ void __launch_bounds__(CFB_ENGINE_THREADS, 3) engine_c2c_kernel(const EngineParams P) {
  CFB_DYN_SMEM(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int T = P.T, ldz = P.ldz, M = P.M;
  cpx *L0 = (cpx *)smem_raw, *L1 = L0 + T * ldz, *W = L1 + T * ldz;
  // row tables, three slots: slot (it+1)%3 is written for the next tile while slow threads may still be storing
  // tile it-1 from slot (it-1)%3
  long long *off_in = (long long *)(W + T * ldz);  // [3][T]
  long long *off_out = off_in + 3 * T;             // [3][T]
  int *row_lo = (int *)(off_out + 3 * T);          // [3][T]
  cpx *tws = (cpx *)(((uintptr_t)(row_lo + 3 * T) + 15) & ~(uintptr_t)15);
  cpx *fss = tws + P.tw_smem;
  const cpx *tw = P.tw;
  if (P.tw_smem > 0) {
    for (int i = tid; i < P.tw_smem; i += nthr) tws[i] = __ldg(P.tw + i);
    tw = tws;
  }
  const cpx *fs = P.fs_tw;
  if (fs && P.fs_smem > 0) {
    for (int i = tid; i < P.fs_smem; i += nthr) fss[i] = __ldg(P.fs_tw + i);
    fs = fss;
  }
  auto fill_offsets = [&](long long tile, int slot) {
    const long long row0 = tile * T;
    for (int r = tid; r < T; r += nthr) {
      long long g = row0 + r;
      const bool ok = tile < P.ntiles && g < P.lot;
      off_in[slot * T + r] = ok ? batch_off(P.ain, g) : -1;
      off_out[slot * T + r] = ok ? batch_off(P.aout, g) : -1;
      row_lo[slot * T + r] = (int)(P.fs_from_hi ? g / P.aout.nlo : g % P.aout.nlo);
    }
  };
  long long tile = blockIdx.x;
  fill_offsets(tile, 0);
  __syncthreads();
  c2c_issue_loads(P, L0, off_in, tid, nthr);
  const TileWalk ws = tile_walk(tid, nthr, P.tx_out_log2, P.aout.lanes_t);
  const long long ostep = (long long)ws.en * P.aout.inc;
  const int fs_mask = (1 << P.fs_shift) - 1, fs_shift = P.fs_shift, ps = P.padshift, al = P.aligned16;
  const double scale = P.scale;
  int it = 0;
  for (; tile < P.ntiles; tile += gridDim.x, ++it) {
    const int slot = it % 3, nslot = (it + 1) % 3;
    cpx *L = (it & 1) ? L1 : L0, *Ln = (it & 1) ? L0 : L1;
    fill_offsets(tile + gridDim.x, nslot);
    __syncthreads();  // offsets of the next tile are visible; everybody is done with the other landing buffer
    c2c_issue_loads(P, Ln, off_in + nslot * T, tid, nthr);  // (empty group past the last tile)
    cp_async_wait<1>();
    __syncthreads();  // this tile has landed for all threads
    cpx *cur = L, *oth = W;
    if (P.dir < 0) run_passes<-1>(cur, oth, P, tw, tid, nthr);
    else run_passes<1>(cur, oth, P, tw, tid, nthr);
    /* ---- store: cur -> global (scale, four-step twiddle) ---- */
    const long long *oo = off_out + slot * T;
    const int *rl = row_lo + slot * T;
    for (int r = ws.rs; r < T; r += ws.rn) {
      const long long o = oo[r];
      if (o < 0) continue;
      cpx *p = (cpx *)P.out + o + (long long)ws.es * P.aout.inc;
      const cpx *row = cur + r * ldz;
      const int j = rl[r];
      int x = j * ws.es;
      const int xstep = j * ws.en;
      for (int e = ws.es; e < M; e += ws.en, p += ostep, x += xstep) {
        cpx v = row[padx(e, ps)];
        v.x *= scale;
        v.y *= scale;
        if (fs) {  // W_n^(j*e), j*e < n: two short tables and one product instead of an n-entry gather
          const cpx w = cmul(fs[x & fs_mask], fs[fs_mask + 1 + (x >> fs_shift)]);
          v = (P.dir < 0) ? cmul(v, w) : cmulc(v, w);
        }
        if (al) *p = v;
        else {
          ((double *)p)[0] = v.x;
          ((double *)p)[1] = v.y;
        }
      }
    }
  }
  cp_async_wait<0>();
}

/* real families (rfft / cost / sint / cosq / sinq), T pairs of sequences per tile, software-pipelined like the
 * complex kernel: the rows of tile k+1 are gathered with 8-byte cp.async while tile k is transformed.
 * Buffers (each T*ldz complex = 2T rows of ldz doubles): B0, B1 alternate as landing buffer; the landing buffer of the
 * current tile becomes the ping-pong partner of A once the pre-processing has consumed the rows. */
template <bool REV>
__device__ __forceinline__ void real_issue_loads(const EngineParams &P, double *rowsL, const long long *off_in, int tid,
                                                 int nthr) {
  const double *in = (const double *)P.in;
  const int rows = 2 * P.T, n = P.n, ldz = P.ldz;
  const bool rev = REV;  // sinqf1_ reverses the sequence first (fftpack.c:14247-14256)
  const TileWalk w = tile_walk(tid, nthr, P.tx_in_log2, P.ain.lanes_t);
  const long long step = (long long)w.en * P.ain.inc;
  for (int r = w.rs; r < rows; r += w.rn) {
    const long long o = off_in[r];
    double *row = rowsL + r * ldz;
    if (o < 0) {  // a missing partner row must be exactly zero: it shares a complex transform with a live row
      for (int e = w.es; e < n; e += w.en) row[e] = 0.0;
      continue;
    }
    const double *p = in + o + (long long)w.es * P.ain.inc;
    for (int e = w.es; e < n; e += w.en, p += step) cp_async8(row + (rev ? n - 1 - e : e), p);
  }
  cp_async_commit();
}

template <int KIND, int DIR>
__global__ void __launch_bounds__(CFB_ENGINE_REAL_MAXTHREADS, 2) engine_kernel(const EngineParams P) {
  CFB_DYN_SMEM(smem_raw);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int T = P.T, ldz = P.ldz, M = P.M, n = P.n;
  cpx *B0 = (cpx *)smem_raw, *B1 = B0 + T * ldz, *A = B1 + T * ldz;
  const int rows = 2 * T;
  long long *off_in = (long long *)(A + T * ldz);  // [3][rows]
  long long *off_out = off_in + 3 * rows;           // [3][rows]
  double *dsum = (double *)(off_out + 3 * rows);    // [rows]
  double *gsc = dsum + rows;                        // [rows][8] partial sums / scan carries of multi-warp groups
  cpx *tws = (cpx *)(((uintptr_t)(gsc + 8 * rows) + 15) & ~(uintptr_t)15);
  const cpx *tw = P.tw;
  if (P.tw_smem > 0) {
    for (int i = tid; i < P.tw_smem; i += nthr) tws[i] = __ldg(P.tw + i);
    tw = tws;
  }
  auto fill_offsets = [&](long long tile, int slot) {
    const long long row0 = tile * rows;
    for (int r = tid; r < rows; r += nthr) {
      long long g = row0 + r;
      const bool ok = tile < P.ntiles && g < P.lot;
      off_in[slot * rows + r] = ok ? batch_off(P.ain, g) : -1;
      off_out[slot * rows + r] = ok ? batch_off(P.aout, g) : -1;
    }
  };
  constexpr int kind = KIND, dir = DIR;  // compile-time: the family-specific branches below fold away
  constexpr bool fwd_core = !((kind == K_RFFT || kind == K_COSQ || kind == K_SINQ) && dir > 0);
  // pre/post-processing: the block splits into `groups` groups of gs threads, one row each at a time (rows and the
  // block size are powers of two, so every group makes the same number of trips -- required by the barriers inside)
  const int gs = (nthr / rows) < 32 ? 32 : (nthr / rows), groups = nthr / gs, gid = tid / gs, gl = tid % gs;
  const TileWalk ws = tile_walk(tid, nthr, P.tx_out_log2, P.aout.lanes_t);
  const long long ostep = (long long)ws.en * P.aout.inc;
  long long tile = blockIdx.x;
  fill_offsets(tile, 0);
  __syncthreads();
  real_issue_loads<(KIND == K_SINQ && DIR < 0)>(P, (double *)B0, off_in, tid, nthr);
  int it = 0;
  for (; tile < P.ntiles; tile += gridDim.x, ++it) {
    const int slot = it % 3, nslot = (it + 1) % 3;
    cpx *zB = (it & 1) ? B1 : B0, *Bn = (it & 1) ? B0 : B1;
    fill_offsets(tile + gridDim.x, nslot);
    __syncthreads();  // next offsets visible; the previous tile's storers are done with the other landing buffer
    real_issue_loads<(KIND == K_SINQ && DIR < 0)>(P, (double *)Bn, off_in + nslot * rows, tid, nthr);
    cp_async_wait<1>();
    __syncthreads();  // rows of this tile have landed
    double *rowsB = (double *)zB, *rowsA = (double *)A;
    cpx *cur, *oth;
    double *ys;
    if (fwd_core) {
      for (int r = gid; r < rows; r += groups)
        pre_forward_core(kind, dir, n, M, rowsB + r * ldz, (double *)(A + (r >> 1) * ldz) + (r & 1), P.trig, dsum + r, gl, gs,
                         gsc + 8 * r);
      __syncthreads();
      cur = A;
      oth = zB;
      run_passes<-1, false>(cur, oth, P, tw, tid, nthr);
      /* split: cur -> half-complex rows in oth */
      double *hs = (double *)oth;
      const int nfq = M / 2 + 1;
      for (int t = 0; t < T; ++t)
        for (int f = tid; f < nfq; f += nthr) split_pair(cur + t * ldz, hs + (2 * t) * ldz, hs + (2 * t + 1) * ldz, M, f);
      __syncthreads();
      /* post: half-complex rows (oth) -> result rows (cur, no longer needed) */
      ys = (double *)cur;
      for (int r = gid; r < rows; r += groups)
        post_sequence(kind, dir, n, M, hs + r * ldz, ys + r * ldz, P.trig, dsum[r], gl, gs, gsc + 8 * r);
    } else {
      for (int r = gid; r < rows; r += groups) pre_backward_core(kind, n, rowsB + r * ldz, rowsA + r * ldz, gl, gs);
      __syncthreads();
      const int nfq = M / 2 + 1;
      for (int t = 0; t < T; ++t)
        for (int f = tid; f < nfq; f += nthr) build_pair(zB + t * ldz, rowsA + (2 * t) * ldz, rowsA + (2 * t + 1) * ldz, M, f);
      __syncthreads();
      cur = zB;
      oth = A;
      run_passes<1, false>(cur, oth, P, tw, tid, nthr);
      /* extract: re/im of cur -> real rows in oth */
      double *us = (double *)oth;
      for (int t = 0; t < T; ++t)
        for (int e = tid; e < M; e += nthr) {
          cpx v = cur[t * ldz + e];
          us[(2 * t) * ldz + e] = v.x;
          us[(2 * t + 1) * ldz + e] = v.y;
        }
      __syncthreads();
      ys = (double *)cur;
      for (int r = gid; r < rows; r += groups)
        post_sequence(kind, dir, n, M, us + r * ldz, ys + r * ldz, P.trig, 0.0, gl, gs, gsc + 8 * r);
    }
    __syncthreads();
    /* store: result rows -> global.  sinq: forward negates the odd entries, backward reverses (fftpack.c:14257-14266) */
    {
      double *out = (double *)P.out;
      const long long *oo = off_out + slot * rows;
      constexpr bool neg_odd = (kind == K_SINQ && dir < 0), rev = (kind == K_SINQ && dir > 0);
      for (int r = ws.rs; r < rows; r += ws.rn) {
        const long long o = oo[r];
        if (o < 0) continue;
        double *p = out + o + (long long)ws.es * P.aout.inc;
        const double *row = ys + r * ldz;
        for (int e = ws.es; e < n; e += ws.en, p += ostep) {
          double v = row[rev ? n - 1 - e : e];
          if (neg_odd && (e & 1)) v = -v;
          *p = v;
        }
      }
    }
    // note: the buffers this tile wrote last (ys in A or zB) are read by slow storers until the next top-of-loop barrier;
    // the next tile's prefetch only targets the other landing buffer, and A / zB are rewritten after that barrier
  }
  cp_async_wait<0>();
}

/* ------------------------------------------------------------------------------------------
 * Long real-family sequences (longer than one CTA's shared memory): the same pre / pair-packing / split / post
 * steps as engine_kernel, run from global scratch arrays around the long complex transform of dispatch.cu.
 * One warp per row; rows are contiguous in the scratch arrays (pitch ld doubles).
 * ------------------------------------------------------------------------------------------ */
struct LongRealParams {
  int kind, dir, n, M;
  long long lot;      // rows
  long long row0;     // first row of this launch (gridDim.y is limited to 65535)
  long long ld;       // pitch of the real scratch rows (doubles) and of the complex rows (cpx)
  Addr a;             // user layout
  double *user;       // caller's array
  double *xs, *ys;    // [lot][ld] real scratch rows
  cpx *z;             // [(lot+1)/2][ld] complex scratch rows
  double *dsum;       // [lot]
  const double *trig;
};

/* user rows -> xs (sinq: reversed on the forward side) */
__global__ void __launch_bounds__(256) long_gather_kernel(const LongRealParams P) {
  const long long row = P.row0 + blockIdx.y;
  const long long off = batch_off(P.a, row);
  const bool rev = (P.kind == K_SINQ && P.dir < 0);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < P.n; e += (long long)gridDim.x * blockDim.x)
    P.xs[row * P.ld + (rev ? P.n - 1 - e : e)] = P.user[off + e * P.a.inc];
}
/* ys -> user rows (sinq: sign / reversal on the way out) */
__global__ void __launch_bounds__(256) long_scatter_kernel(const LongRealParams P) {
  const long long row = P.row0 + blockIdx.y;
  const long long off = batch_off(P.a, row);
  const bool neg_odd = (P.kind == K_SINQ && P.dir < 0), rev = (P.kind == K_SINQ && P.dir > 0);
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < P.n; e += (long long)gridDim.x * blockDim.x) {
    double v = P.ys[row * P.ld + (rev ? P.n - 1 - e : e)];
    if (neg_odd && (e & 1)) v = -v;
    P.user[off + e * P.a.inc] = v;
  }
}
/* stage 1 (one warp per row): forward core: xs -> re/im lane of z; backward core: xs -> half-complex row in ys */
__global__ void __launch_bounds__(32) long_pre_kernel(const LongRealParams P, int fwd_core) {
  const long long row = blockIdx.x;
  const int lane = threadIdx.x;
  if (fwd_core) {
    double *zc = (double *)(P.z + (row >> 1) * P.ld) + (row & 1);
    pre_forward_core(P.kind, P.dir, P.n, P.M, P.xs + row * P.ld, zc, P.trig, P.dsum + row, lane, 32, nullptr);
    if ((row == P.lot - 1) && !(row & 1))  // odd lot: the missing partner row is zero
      for (int j = lane; j < P.M; j += 32) zc[2 * j + 1] = 0.0;
  } else {
    pre_backward_core(P.kind, P.n, P.xs + row * P.ld, P.ys + row * P.ld, lane, 32);
    if ((row == P.lot - 1) && !(row & 1))
      for (int j = lane; j < P.M; j += 32) P.ys[(row + 1) * P.ld + j] = 0.0;
  }
}
/* backward core: half-complex rows (ys) -> spectrum z of the pair */
__global__ void __launch_bounds__(256) long_build_kernel(const LongRealParams P) {
  const long long pr = P.row0 + blockIdx.y;
  const int nfq = P.M / 2 + 1;
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nfq; f += gridDim.x * blockDim.x)
    build_pair(P.z + pr * P.ld, P.ys + (2 * pr) * P.ld, P.ys + (2 * pr + 1) * P.ld, P.M, f);
}
/* forward core: z -> half-complex rows (xs);  backward core: z -> real rows (xs) */
__global__ void __launch_bounds__(256) long_split_kernel(const LongRealParams P, int fwd_core) {
  const long long pr = P.row0 + blockIdx.y;
  if (fwd_core) {
    const int nfq = P.M / 2 + 1;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nfq; f += gridDim.x * blockDim.x)
      split_pair(P.z + pr * P.ld, P.xs + (2 * pr) * P.ld, P.xs + (2 * pr + 1) * P.ld, P.M, f);
  } else {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < P.M; e += gridDim.x * blockDim.x) {
      cpx v = P.z[pr * P.ld + e];
      P.xs[(2 * pr) * P.ld + e] = v.x;
      P.xs[(2 * pr + 1) * P.ld + e] = v.y;
    }
  }
}
/* last stage (one warp per row): xs -> ys */
__global__ void __launch_bounds__(32) long_post_kernel(const LongRealParams P, int fwd_core) {
  const long long row = blockIdx.x;
  post_sequence(P.kind, P.dir, P.n, P.M, P.xs + row * P.ld, P.ys + row * P.ld, P.trig, fwd_core ? P.dsum[row] : 0.0,
                threadIdx.x, 32, nullptr);
}

/* ------------------------------------------------------------------------------------------
 * Bluestein (chirp-z) kernels for lengths whose prime factors exceed what one CTA or the four-step split can hold:
 * X_k = c_k * sum_j (x_j c_j) conj(c)_{k-j},  c_j = exp(-+ pi i j^2 / n).  The convolution runs as power-of-two
 * transforms of length L >= 2n-1 on a scratch array z[lot][L].
 * ------------------------------------------------------------------------------------------ */
struct ChirpParams {
  int n, L, dir;
  long long lot, row0;
  Addr a;             // user layout
  cpx *user;
  cpx *z;             // [lot][L]
  const cpx *chirp;   // [n]  forward chirp exp(-pi i j^2 / n)
  const cpx *bhat;    // [L]  transform of the wrapped conjugate chirp (forward) -- backward uses its conjugate symmetry
  double scale;
};
/* z[row][j] = x[row][j] * c_j (j < n), 0 (n <= j < L) */
__global__ void __launch_bounds__(256) chirp_pre_kernel(const ChirpParams P) {
  const long long row = P.row0 + blockIdx.y;
  const long long off = batch_off(P.a, row);
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.L; j += gridDim.x * blockDim.x) {
    cpx v = make_double2(0.0, 0.0);
    if (j < P.n) {
      const cpx x = P.user[off + j * P.a.inc], c = P.chirp[j];
      v = P.dir < 0 ? cmul(x, c) : cmulc(x, c);
    }
    P.z[row * P.L + j] = v;
  }
}
/* z[row][j] *= bhat[j]  (backward: the kernel of the convolution is the conjugate chirp, whose transform is
 * conj(bhat[(L - j) % L])) */
__global__ void __launch_bounds__(256) chirp_mul_kernel(const ChirpParams P) {
  const long long row = P.row0 + blockIdx.y;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < P.L; j += gridDim.x * blockDim.x) {
    const cpx b = P.dir < 0 ? P.bhat[j] : P.bhat[(P.L - j) % P.L];
    cpx *q = P.z + row * P.L + j;
    *q = P.dir < 0 ? cmul(*q, b) : cmulc(*q, b);
  }
}
/* x[row][k] = z[row][k] * c_k * scale */
__global__ void __launch_bounds__(256) chirp_post_kernel(const ChirpParams P) {
  const long long row = P.row0 + blockIdx.y;
  const long long off = batch_off(P.a, row);
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < P.n; k += gridDim.x * blockDim.x) {
    const cpx v = P.z[row * P.L + k], c = P.chirp[k];
    cpx y = P.dir < 0 ? cmul(v, c) : cmulc(v, c);
    P.user[off + k * P.a.inc] = make_double2(y.x * P.scale, y.y * P.scale);
  }
}

/* closed forms for the lengths the reference special-cases (costf1_ n=2,3 fftpack.c:6339-6353; sintf1_ n=2
 * :14858-14866; cosqf1_ n=2 :5498-5502; the backward twins) -- one thread per sequence */
struct TinyParams {
  int kind, dir, n;
  long long lot;
  Addr a;
  double *x;
};
__global__ void __launch_bounds__(128) tiny_kernel(const TinyParams P) {
  long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= P.lot) return;
  double *x = P.x + batch_off(P.a, g);
  const long long inc = P.a.inc;
  const bool fwd = P.dir < 0;
  if (P.kind == K_COST && P.n == 2) {
    double a = x[0], b = x[inc];
    if (fwd) {
      x[0] = (a + b) * 0.5;
      x[inc] = (a - b) * 0.5;
    } else {
      x[0] = a + b;
      x[inc] = a - b;
    }
  } else if (P.kind == K_COST && P.n == 3) {
    double a = x[0], b = x[inc], c = x[2 * inc], s = a + c;
    if (fwd) {
      double tb = b + b;
      x[inc] = (a - c) * 0.5;
      x[0] = (s + tb) * 0.25;
      x[2 * inc] = (s - tb) * 0.25;
    } else {
      x[inc] = a - c;
      x[0] = s + b;
      x[2 * inc] = s - b;
    }
  } else if (P.kind == K_SINT && P.n == 2) {
    const double c = fwd ? 0.57735026918962576450914878050196 : 0.86602540378443864676372317075294;
    double a = x[0], b = x[inc];
    x[0] = c * (a + b);
    x[inc] = c * (a - b);
  } else if ((P.kind == K_COSQ || P.kind == K_SINQ) && P.n == 2) {
    const double h = 0.70710678118654752440084436210485;
    double a = x[0], b = x[inc];
    if (P.kind == K_SINQ) {
      if (fwd) {  // reverse, cosq, negate odd entry
        double t = a;
        a = b;
        b = t;
      } else {
        b = -b;
      }
    }
    double y0, y1;
    if (fwd) {
      y0 = a * 0.5 + h * b;
      y1 = a * 0.5 - h * b;
    } else {
      y0 = a + b;
      y1 = h * (a - b);
    }
    if (P.kind == K_SINQ) {
      if (fwd) y1 = -y1;
      else {
        double t = y0;
        y0 = y1;
        y1 = t;
      }
    }
    x[0] = y0;
    x[inc] = y1;
  }
}

}  // namespace cfb
#endif
