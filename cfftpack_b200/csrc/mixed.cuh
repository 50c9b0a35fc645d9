/*
 * mixed.cuh -- streaming kernel for the real families (rfftm*, cosqm*, sintm*, costm*) whose underlying real FFT has a
 * length M ~ 1000 that is not a power of two: the lengths of BASELINE config 4.
 *
 *     M = 1001 = 13*11*7   cosq N=1001, sint N=1000, rfft N=1001, cost N=1002      (three register passes)
 *     M = 1000 = 10*10*10  cosq N=1000, cost N=1001, rfft N=1000, sint N= 999      (three register passes)
 *     M =  999 = 9*3*37    cost N=1000, ...                  (two register passes + a generic radix-37 pass)
 *     M = 1002 = 6*167     sint N=1001, ...                  (one register pass  + a generic radix-167 pass)
 *
 * It replaces, for those shapes, the pass loops mrftf1_/mrftb1_ (cfftpack/fftpack.c:10149, :9946) with their passes
 * mradf2..5/mradfg (:8600-9443; the generic one is the O(ip) per output scheme of r1fgkf_ :12564) and the pre/post sweeps
 * of mcsqf1_/mcsqb1_ (:6839, :6740), msntf1_/msntb1_ (:10636, :10530) and mcstf1_/mcstb1_ (:7150, :7045).
 *
 * Design (one HBM read and one HBM write per element, both by the bulk-copy engine):
 *   - one PAIR of contiguous rows per tile, z = x_a + i x_b, one complex transform of length M per pair;
 *   - persistent CTAs; the next pair (2n doubles, one cp.async.bulk) lands in shared memory while the current one is
 *     transformed (mbarrier complete_tx);
 *   - Stockham passes ping-ponging between two shared-memory buffers, one __syncthreads per pass.  Radices up to 13 are
 *     in-register butterflies (butterfly.cuh).  A large prime radix r (37, 167) is the LAST pass: symmetric sums in place,
 *     then every thread accumulates a block of KB output pairs (k, r-k) of one butterfly over j = 1..(r-1)/2 -- 4 KB FP64
 *     FMAs per two 16-byte shared loads, the roots looked up with a running index (j k mod r);
 *   - the family's pre-processing (cosqf1_ fold :5693-5712, sintf1_ fold :14873-14890, costf1_ fold :6355-6377,
 *     half-complex -> spectrum of rfftb1_) is computed on the way from the landing buffer into the first pass;
 *   - the Hermitian split and the family's post-processing write the finished rows, in their final layout, into the
 *     free buffer, and ONE cp.async.bulk shared->global per pair drains it (no per-thread global stores);
 *   - the serial running sums of sintf1_ (:14905-14915) and costf1_ (:6386-6400) are warp scans over a padded copy.
 * Twiddles w^k (k < radix) are rebuilt from the table rows w^p (and w^4p) kept in shared memory.
 */
#ifndef CFB_MIXED_CUH
#define CFB_MIXED_CUH
#include "butterfly.cuh"
#include "engine_types.h"
#include "internal.h"
#include "tma.cuh"

namespace cfb {

constexpr int MIX_GENERIC_MIN = 17;  // radices from here on take the generic (shared-memory) pass
constexpr int MIX_KB = 6;            // output pairs per thread in the generic pass

/* R1 = 1: no middle pass.  R2 >= MIX_GENERIC_MIN: generic last pass. */
template <int R0_, int R1_, int R2_>
struct MixCfg {
  static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;
  static constexpr int M = R0 * R1 * R2, HF = M / 2;
  static constexpr bool EVEN = (M % 2) == 0;
  // generic last pass: butterflies per work item.  2 halves the root look-ups per FMA but leaves only 42 of the 96
  // threads busy for M = 1002 (measured 0.700 ms against 0.619 ms with 1), so 1 it is.
  static constexpr int QB = 1;
  static constexpr int NB0 = M / R0;                 // butterflies of pass 0 = its twiddle period
  static constexpr int MM1 = M / (R0 * R1);          // twiddle period of pass 1
  static constexpr bool W4_0 = R0 > 6, W4_1 = R1 > 6;  // a w^4p row only where the radix needs powers beyond 5
  // table (cpx): pass 0 rows w_M^p [, w_M^4p] (p < NB0); pass 1 rows (p < MM1); roots of unity of R0, R1, R2
  static constexpr int T0 = 0, T1 = (W4_0 ? 2 : 1) * NB0, RT0 = T1 + (R1 > 1 ? (W4_1 ? 2 : 1) * MM1 : 0), RT1 = RT0 + R0,
                       RT2 = RT1 + R1, TAB = RT2 + R2;
  static constexpr int SCAN_PITCH = (HF + 2) + (HF + 2) / 16 + 8;  // padded copy of one row's running-sum terms
  static constexpr size_t BUF = ((size_t)(M + 1) * sizeof(cpx) + 15) / 16 * 16;  // >= 2n doubles for n <= M + 1
  static constexpr size_t OFF_L = 0, OFF_P = BUF, OFF_Q = 2 * BUF, OFF_TAB = 3 * BUF,
                          OFF_RED = OFF_TAB + (size_t)TAB * sizeof(cpx), OFF_BAR = OFF_RED + 16 * sizeof(double);
  static constexpr size_t BYTES = OFF_BAR + 16;
};

/* a[k] *= w^k (k = 1..R-1); w^k from w1 = w and, for R > 6, w4 = w^4 (few products deep) */
template <int R, int DIR>
__device__ __forceinline__ void mix_twiddle(cpx (&a)[R], const cpx w1, const cpx w4_in) {
  cpx lo[4];
  lo[1] = w1;
  lo[2] = cmul(w1, w1);
  lo[3] = cmul(lo[2], w1);
  const cpx w4 = (R > 6) ? w4_in : cmul(lo[2], lo[2]);
  cpx hi = w4;  // w^(4j)
#pragma unroll
  for (int k = 1; k < R; ++k) {
    if (k < 4) a[k] = ctw<DIR>(a[k], lo[k]);
    else {
      if (k % 4 == 0 && k > 4) hi = cmul(hi, w4);
      a[k] = ctw<DIR>(a[k], (k % 4 == 0) ? hi : cmul(hi, lo[k % 4]));
    }
  }
}

/* one register pass: butterfly b = q + S p reads elements b + NB i (i < R) through `load`, writes q + S R p + S k */
template <int R, int S, int MM, bool W4, int W, int DIR, class Load>
__device__ __forceinline__ void mix_pass(const Load &load, cpx *__restrict__ dst, const cpx *__restrict__ tw,
                                         const cpx *__restrict__ rt, const int t) {
  constexpr int NB = S * MM;
#pragma unroll 1
  for (int b = t; b < NB; b += W) {
    const int p = b / S, q = b - p * S;
    cpx a[R];
#pragma unroll
    for (int i = 0; i < R; ++i) a[i] = load(b + NB * i);
    DftRt<R, DIR>::run(a, rt);
    if (MM > 1) mix_twiddle<R, DIR>(a, tw[p], W4 ? tw[MM + p] : tw[p]);
    cpx *d = dst + q + S * R * p;
#pragma unroll
    for (int k = 0; k < R; ++k) d[S * k] = a[k];
  }
}

/* generic last pass, radix R (odd prime), S butterflies: src[q + S i] -> dst[q + S k].  src is overwritten.
 * An item is a block of KB output pairs (k, R-k) of QB butterflies: every root of unity fetched from shared memory feeds
 * 4 QB FMAs, every pair of data loads 4 KB FMAs, so for QB = 2 the loop is bound by the FP64 pipe, not by loads. */
template <int R, int S, int W, int DIR, int QB>
__device__ __forceinline__ void mix_pass_generic(cpx *__restrict__ src, cpx *__restrict__ dst, const cpx *__restrict__ rt, const int t) {
  constexpr int H = (R - 1) / 2, KB = MIX_KB, NBLK = (H + KB - 1) / KB, SQ = S / QB;
  static_assert(S % QB == 0, "butterflies per item must divide their number");
  // 1. symmetric sums in place: src[q + S j] <- x_j + x_{R-j}, src[q + S (R-j)] <- x_j - x_{R-j}
  for (int it = t; it < S * H; it += W) {
    const int q = it % S, j = 1 + it / S;
    const cpx u = src[q + S * j], v = src[q + S * (R - j)];
    src[q + S * j] = cadd(u, v);
    src[q + S * (R - j)] = csub(u, v);
  }
  __syncthreads();
  // 2. item = (block of KB output pairs, group of QB butterflies), groups fastest: lanes of one block share the root look-ups
  for (int it = t; it < SQ * NBLK; it += W) {
    const int q = (it % SQ) * QB, k0 = 1 + (it / SQ) * KB;
    double ar[QB][KB], ai[QB][KB], br[QB][KB], bi[QB][KB], s0x[QB], s0y[QB];
    int idx[KB];
#pragma unroll
    for (int b = 0; b < QB; ++b) {
      const cpx x0 = src[q + b];
      s0x[b] = x0.x;  // X_0 = x_0 + sum_j p_j (stored by the first block)
      s0y[b] = x0.y;
#pragma unroll
      for (int u = 0; u < KB; ++u) {
        ar[b][u] = x0.x;
        ai[b][u] = x0.y;
        br[b][u] = bi[b][u] = 0.0;
      }
    }
    // root look-ups by BYTE offset (j k mod R) * 16 kept as a running sum: add, compare, conditional subtract per root
    // (indexing rt[] with an element index cost two more integer instructions each: 79 integer vs 52 FP64 per two j)
    const char *rtb = (const char *)rt;
    constexpr unsigned LIM = R * (unsigned)sizeof(cpx);
#pragma unroll
    for (int u = 0; u < KB; ++u) idx[u] = 0;
    const cpx *pp = src + q, *mp = src + q + S * R;
#pragma unroll 2
    for (int j = 1; j <= H; ++j) {
      cpx p[QB], m[QB];
      pp += S;
      mp -= S;
#pragma unroll
      for (int b = 0; b < QB; ++b) {
        p[b] = pp[b];
        m[b] = mp[b];
        s0x[b] += p[b].x;
        s0y[b] += p[b].y;
      }
#pragma unroll
      for (int u = 0; u < KB; ++u) {
        unsigned i2 = (unsigned)idx[u] + (unsigned)(k0 + u) * (unsigned)sizeof(cpx);  // < 2 LIM
        i2 -= (i2 >= LIM) ? LIM : 0u;
        idx[u] = (int)i2;
        const cpx w = *(const cpx *)(rtb + i2);  // (cos, -sin)(2 pi j k / R)
#pragma unroll
        for (int b = 0; b < QB; ++b) {
          ar[b][u] = fma(w.x, p[b].x, ar[b][u]);
          ai[b][u] = fma(w.x, p[b].y, ai[b][u]);
          br[b][u] = fma(-w.y, m[b].x, br[b][u]);
          bi[b][u] = fma(-w.y, m[b].y, bi[b][u]);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < QB; ++b) {
      if (k0 == 1) dst[q + b] = make_double2(s0x[b], s0y[b]);
#pragma unroll
      for (int u = 0; u < KB; ++u) {
        const int k = k0 + u;
        if (k <= H) {  // X_k = A + DIR i B, X_{R-k} = A - DIR i B with B = (br, bi)
          if (DIR < 0) {
            dst[q + b + S * k] = make_double2(ar[b][u] + bi[b][u], ai[b][u] - br[b][u]);
            dst[q + b + S * (R - k)] = make_double2(ar[b][u] - bi[b][u], ai[b][u] + br[b][u]);
          } else {
            dst[q + b + S * k] = make_double2(ar[b][u] - bi[b][u], ai[b][u] + br[b][u]);
            dst[q + b + S * (R - k)] = make_double2(ar[b][u] + bi[b][u], ai[b][u] - br[b][u]);
          }
        }
      }
    }
  }
}

/* inclusive running sums of `count` terms per row (two rows), terms at s[g + (g >> 4)]: warp 0 row a, warp 1 row b */
__device__ __forceinline__ void mix_scan_rows(double *sa, double *sb, const int count, const int t) {
  if (t < 64) {
    const int lane = t & 31;
    double *s = (t < 32 ? sa : sb);
    double carry = 0.0;
    for (int base = 0; base < count; base += 512) {  // 16 terms per lane per round (pitch 17: conflict-free)
      const int g0 = base + 16 * lane;
      double *sp = s + g0 + (g0 >> 4);
      double v[16], run = 0.0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        v[i] = (g0 + i < count) ? sp[i] : 0.0;
        run += v[i];
        v[i] = run;
      }
      const double off = warp_excl_scan(run, lane) + carry;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (g0 + i < count) sp[i] = v[i] + off;
      carry = __shfl_sync(0xffffffffu, off + run, 31);
    }
  }
}

/* KIND: K_RFFT (n = M), K_COSQ (n = M), K_SINT (n = M - 1), K_COST (n = M + 1).  DIR: -1 forward, +1 backward. */
template <class C, int KIND, int DIR, int W>
__global__ void __launch_bounds__(W, (W <= 96 ? 4 : W <= 128 ? 3 : 2)) mix_stream_kernel(double *__restrict__ x, long long npairs,
                                                                                       const cpx *__restrict__ tab_g,
                                                                                       const double *__restrict__ trig) {
  CFB_DYN_SMEM(smem_raw);
  constexpr int M = C::M, HF = C::HF, n = (KIND == K_SINT) ? M - 1 : (KIND == K_COST) ? M + 1 : M;
  constexpr bool EVEN = C::EVEN;
  // direction of the complex transform: sintb1_ and costb1_ run the forward real transform too
  constexpr int CD = (KIND == K_SINT || KIND == K_COST) ? -1 : DIR;
  constexpr unsigned PAIR_BYTES = 2u * n * sizeof(double);
  constexpr bool GENERIC = C::R2 >= MIX_GENERIC_MIN;
  static_assert(PAIR_BYTES % 16 == 0 && PAIR_BYTES <= C::BUF, "a pair of rows is one aligned bulk copy");
  static_assert(W >= 64 && W % 32 == 0, "the running sums of the two rows take one warp each");
  double *land = (double *)(smem_raw + C::OFF_L);
  cpx *P = (cpx *)(smem_raw + C::OFF_P), *Q = (cpx *)(smem_raw + C::OFF_Q);
  cpx *tab = (cpx *)(smem_raw + C::OFF_TAB);
  double *red = (double *)(smem_raw + C::OFF_RED);  // cost: dsum partials [2][W / 32]
  uint64_t *bar = (uint64_t *)(smem_raw + C::OFF_BAR);
  const int t = threadIdx.x;
  if (t == 0) mbar_init(bar, 1);
  for (int i = t; i < C::TAB; i += W) tab[i] = __ldg(tab_g + i);
  __syncthreads();
  long long tile = blockIdx.x;
  if (t == 0 && tile < npairs) {
    mbar_expect_tx(bar, PAIR_BYTES);
    bulk_g2s(land, x + tile * 2 * n, PAIR_BYTES, bar);
  }
  unsigned parity = 0;
  const double *la = land, *lb = land + n;
  double *qa = (double *)Q, *qb = qa + n;
  for (; tile < npairs; tile += gridDim.x) {
    mbar_wait(bar, parity);
    parity ^= 1;
    if constexpr (C::R1 == 1) {
      // two-pass lengths: pass 0 writes Q, which the previous pair's bulk store may still be reading
      if (t == 0) bulk_wait_read();
      __syncthreads();
    }
    /* ---- pass 0 (radix R0): the family's pre-processing happens in the loader ---- */
    double dsa = 0.0, dsb = 0.0;                     // cost: this thread's share of dsum (costf1_ :6361-6371)
    const double ends = (KIND == K_COST && DIR > 0) ? 2.0 : 1.0;  // costb1_ doubles the end points first
    auto load0 = [&](int e) -> cpx {
      if (KIND == K_RFFT && DIR < 0) return make_double2(la[e], lb[e]);
      if (KIND == K_SINT) {  // sintf1_: xh[0] = 0, xh[k] = t1 + t2, xh[M-k] = t2 - t1, middle (M even) 4 x[n/2]
        if (e == 0) return make_double2(0.0, 0.0);
        if (EVEN && e == HF) return make_double2(4.0 * la[n / 2], 4.0 * lb[n / 2]);
        const int k = e < M - e ? e : M - e;
        const double s = __ldg(trig + k - 1);
        const double t1a = la[k - 1] - la[n - k], t2a = s * (la[k - 1] + la[n - k]);
        const double t1b = lb[k - 1] - lb[n - k], t2b = s * (lb[k - 1] + lb[n - k]);
        return e < M - e ? make_double2(t1a + t2a, t1b + t2b) : make_double2(t2a - t1a, t2b - t1b);
      }
      if (KIND == K_COST) {  // costf1_: u[0] = x[0] + x[n-1], u[j] = t1 - S t2, u[M-j] = t1 + S t2, middle (M even) 2 x[M/2]
        if (e == 0) {
          dsa += ends * (la[0] - la[n - 1]);
          dsb += ends * (lb[0] - lb[n - 1]);
          return make_double2(ends * (la[0] + la[n - 1]), ends * (lb[0] + lb[n - 1]));
        }
        if (EVEN && e == HF) return make_double2(la[e] + la[e], lb[e] + lb[e]);
        const int j = e < M - e ? e : M - e, jc = M - j;
        const double t1a = la[j] + la[jc], t2a = la[j] - la[jc], t1b = lb[j] + lb[jc], t2b = lb[j] - lb[jc];
        const double sj = __ldg(trig + j);
        if (e < M - e) {
          const double cj = __ldg(trig + M + j);
          dsa = fma(cj, t2a, dsa);
          dsb = fma(cj, t2b, dsb);
          return make_double2(fma(-sj, t2a, t1a), fma(-sj, t2b, t1b));
        }
        return make_double2(fma(sj, t2a, t1a), fma(sj, t2b, t1b));
      }
      if (KIND == K_COSQ && DIR < 0) {  // cosqf1_ fold, middle (n even) 2 W[n/2-1] x[n/2]
        if (e == 0) return make_double2(la[0], lb[0]);
        if (EVEN && e == HF) {
          const double w = 2.0 * __ldg(trig + HF - 1);
          return make_double2(w * la[e], w * lb[e]);
        }
        const int j = e < M - e ? e : M - e, jc = M - j;
        const double wj = __ldg(trig + j - 1), wc = __ldg(trig + jc - 1);
        const double sa = la[j] + la[jc], da = la[j] - la[jc], sb = lb[j] + lb[jc], db = lb[j] - lb[jc];
        return e < M - e ? make_double2(fma(wj, da, wc * sa), fma(wj, db, wc * sb))
                         : make_double2(fma(wj, sa, -(wc * da)), fma(wj, sb, -(wc * db)));
      }
      // backward: half-complex rows h_a, h_b -> spectrum Z of z = x_a + i x_b (rfftb1_ convention); cosqb1_ first forms
      // h[0] = x[0]/2, h[2f-1] = (x[2f-1] + x[2f])/2, h[2f] = (x[2f-1] - x[2f])/2, h[n-1] = x[n-1]/2 (n even)
      const double edge = (KIND == K_COSQ) ? 0.5 : 1.0;
      if (e == 0) return make_double2(edge * la[0], edge * lb[0]);
      if (EVEN && e == HF) return make_double2(edge * la[M - 1], edge * lb[M - 1]);
      const int f = e < M - e ? e : M - e;
      double h1a = la[2 * f - 1], h2a = la[2 * f], h1b = lb[2 * f - 1], h2b = lb[2 * f];
      if (KIND == K_COSQ) {
        const double s1 = 0.5 * (h1a + h2a), d1 = 0.5 * (h1a - h2a), s2 = 0.5 * (h1b + h2b), d2 = 0.5 * (h1b - h2b);
        h1a = s1;
        h2a = d1;
        h1b = s2;
        h2b = d2;
      }
      const double a1 = 0.5 * h1a, a2 = 0.5 * h2a, b1 = 0.5 * h1b, b2 = 0.5 * h2b;
      return e < M - e ? make_double2(a1 + b2, b1 - a2) : make_double2(a1 - b2, b1 + a2);
    };
    cpx *const dst0 = (C::R1 > 1) ? P : Q;  // without a middle pass, pass 0 writes where the last pass reads
    mix_pass<C::R0, 1, C::NB0, C::W4_0, W, CD>(load0, dst0, tab + C::T0, tab + C::RT0, t);
    if (KIND == K_COST) {
      dsa = warp_sum(dsa);
      dsb = warp_sum(dsb);
      if ((t & 31) == 0) {
        red[t >> 5] = dsa;
        red[W / 32 + (t >> 5)] = dsb;
      }
    }
    // the previous pair's bulk store must have finished reading Q before this pair's passes overwrite it
    if (t == 0) bulk_wait_read();
    __syncthreads();  // landing buffer consumed (pass 0 has turned every read into results stored in shared memory)
    const long long next = tile + gridDim.x;
    if (t == 0 && next < npairs) {
      mbar_expect_tx(bar, PAIR_BYTES);
      bulk_g2s(land, x + next * 2 * n, PAIR_BYTES, bar);
    }
    /* ---- pass 1 (radix R1): P -> Q;  last pass (radix R2): Q -> P, natural order ---- */
    if constexpr (C::R1 > 1) {
      mix_pass<C::R1, C::R0, C::MM1, C::W4_1, W, CD>([&](int e) -> cpx { return P[e]; }, Q, tab + C::T1, tab + C::RT1, t);
      __syncthreads();
    }
    if constexpr (GENERIC) mix_pass_generic<C::R2, C::R0 * C::R1, W, CD, C::QB>(Q, P, tab + C::RT2, t);
    else mix_pass<C::R2, C::R0 * C::R1, 1, false, W, CD>([&](int e) -> cpx { return Q[e]; }, P, nullptr, tab + C::RT2, t);
    __syncthreads();
    /* ---- split / post-processing: P -> finished rows in Q ---- */
    if (CD < 0) {
      // FFTPACK's half-complex row [X0, A1, B1, ...]: A_f = 2 Re X_f / M, B_f = -2 Im X_f / M (rfftf1_ epilogue :13818-13853)
      constexpr bool SCAN = (KIND == K_SINT || KIND == K_COST);
      constexpr int NTERM = (KIND == K_SINT) ? (n + 1) / 2 : n / 2;  // running-sum terms g = 0 .. NTERM-1
      constexpr int ITERS = (HF + W) / W;
      const double sc = 1.0 / (double)M;
      // sint: output factor of sintf1_ / sintb1_; cost: c0, c1 of costf1_ / costb1_ (:6386-6407, :6230-6250)
      const double ss = DIR < 0 ? 0.5 : 0.25 * (double)(KIND == K_SINT ? M : 0);
      const double c0 = DIR < 0 ? 0.5 : 0.5 * (double)M, c1 = DIR < 0 ? 0.5 : 0.25 * (double)M;
      double keep_a[ITERS], keep_b[ITERS];  // running-sum terms, stored after the barrier
      double Da = 0.0, Db = 0.0;
      if (KIND == K_COST) {
#pragma unroll
        for (int k = 0; k < W / 32; ++k) {
          Da += red[k];
          Db += red[W / 32 + k];
        }
        Da = DIR < 0 ? Da * sc : 0.5 * Da;
        Db = DIR < 0 ? Db * sc : 0.5 * Db;
      }
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int f = t + it * W;
        keep_a[it] = keep_b[it] = 0.0;
        if (f > HF) continue;
        const cpx u = P[f];
        if (f == 0) {
          if (KIND == K_SINT) {
            keep_a[it] = ss * (u.x * sc);
            keep_b[it] = ss * (u.y * sc);
          } else if (KIND == K_COST) {
            keep_a[it] = Da;
            keep_b[it] = Db;
            qa[0] = c0 * (u.x * sc);
            qb[0] = c0 * (u.y * sc);
          } else {
            qa[0] = u.x * sc;
            qb[0] = u.y * sc;
          }
          continue;
        }
        if (EVEN && f == HF) {  // Nyquist term X_{M/2} (real for both rows)
          if (KIND == K_RFFT || KIND == K_COSQ) {
            qa[n - 1] = u.x * sc;
            qb[n - 1] = u.y * sc;
          } else if (KIND == K_COST) {  // y[n-1] = c1 X_{M/2}/M (forward), twice that (backward)
            const double lf = DIR < 0 ? c1 : 2.0 * c1;
            qa[n - 1] = lf * (u.x * sc);
            qb[n - 1] = lf * (u.y * sc);
          }
          continue;
        }
        const cpx v = P[M - f];
        const double Aa = (u.x + v.x) * sc, Ba = (v.y - u.y) * sc, Ab = (u.y + v.y) * sc, Bb = (u.x - v.x) * sc;
        if (KIND == K_RFFT) {
          qa[2 * f - 1] = Aa;
          qa[2 * f] = Ba;
          qb[2 * f - 1] = Ab;
          qb[2 * f] = Bb;
        } else if (KIND == K_COSQ) {  // cosqf1_ post :5728-5738
          qa[2 * f - 1] = 0.5 * (Aa + Ba);
          qa[2 * f] = 0.5 * (Aa - Ba);
          qb[2 * f - 1] = 0.5 * (Ab + Bb);
          qb[2 * f] = 0.5 * (Ab - Bb);
        } else if (KIND == K_SINT) {  // sintf1_ post :14897-14919: y[2f-1] = ss B_f, y[2f] = ss (h0 + A_1 + ... + A_f)
          qa[2 * f - 1] = ss * Ba;
          qb[2 * f - 1] = ss * Bb;
          keep_a[it] = ss * Aa;
          keep_b[it] = ss * Ab;
        } else {  // costf1_ post: y[2f] = c1 A_f, y[2f+1] = D + c1 (B_1 + ... + B_f)
          qa[2 * f] = c1 * Aa;
          qb[2 * f] = c1 * Ab;
          keep_a[it] = c1 * Ba;
          keep_b[it] = c1 * Bb;
        }
      }
      if (SCAN) {
        constexpr int SP = C::SCAN_PITCH;
        double *sa = (double *)P, *sb = sa + SP;
        __syncthreads();  // every thread has read its spectrum values: P becomes the scan scratch
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int g = t + it * W;
          if (g < NTERM) {
            sa[g + (g >> 4)] = keep_a[it];
            sb[g + (g >> 4)] = keep_b[it];
          }
        }
        __syncthreads();
        mix_scan_rows(sa, sb, NTERM, t);
        __syncthreads();
        // sint: y[2g] = sum_{<= g};  cost: y[2g+1] = sum_{<= g}, and for n even the last one is y[n-1] = dsum (halved forward)
        constexpr int POS = (KIND == K_SINT) ? 0 : 1;
        const double last = (KIND == K_COST && !EVEN && DIR < 0) ? 0.5 : 1.0;
        for (int g = t; g < NTERM; g += W) {
          const double f = (g == NTERM - 1) ? last : 1.0;
          qa[2 * g + POS] = f * sa[g + (g >> 4)];
          qb[2 * g + POS] = f * sb[g + (g >> 4)];
        }
      }
    } else if (KIND == K_COSQ) {
      // cosqb1_ post :5626-5652: y[0] = 2 u[0]; y[j] = p + q, y[n-j] = p - q,
      //   p = W[j-1] u[n-j] + W[n-j-1] u[j], q = W[j-1] u[j] - W[n-j-1] u[n-j]; middle (n even) 2 W[n/2-1] u[n/2]
      for (int j = t; j <= HF; j += W) {
        const cpx uj = P[j];
        if (j == 0) {
          qa[0] = uj.x + uj.x;
          qb[0] = uj.y + uj.y;
          continue;
        }
        if (EVEN && j == HF) {
          const double w = 2.0 * __ldg(trig + HF - 1);
          qa[j] = w * uj.x;
          qb[j] = w * uj.y;
          continue;
        }
        const cpx uc = P[M - j];
        const double wj = __ldg(trig + j - 1), wc = __ldg(trig + M - j - 1);
        const double pa = fma(wj, uc.x, wc * uj.x), ra = fma(wj, uj.x, -(wc * uc.x));
        const double pb = fma(wj, uc.y, wc * uj.y), rb = fma(wj, uj.y, -(wc * uc.y));
        qa[j] = pa + ra;
        qa[M - j] = pa - ra;
        qb[j] = pb + rb;
        qb[M - j] = pb - rb;
      }
    } else {  // rfftmb_: the inverse transform of Z is x_a + i x_b
      for (int j = t; j < M; j += W) {
        const cpx z = P[j];
        qa[j] = z.x;
        qb[j] = z.y;
      }
    }
    fence_async_smem();
    __syncthreads();
    if (t == 0) {
      bulk_s2g(x + tile * 2 * n, Q, PAIR_BYTES);
      bulk_commit();
    }
  }
  if (t == 0) bulk_wait_all();
}

/* host side (mixed.cu + one translation unit per length) */
bool mix_supported(int kind, int n);
/* `npairs` pairs of contiguous rows (jump = n, 16-byte aligned base); trig = the family's table (nullptr for rfft) */
bool mix_launch(int kind, int n, long long npairs, int dir, double *x, const double *trig);
void mix_release_tables();

}  // namespace cfb
#endif
