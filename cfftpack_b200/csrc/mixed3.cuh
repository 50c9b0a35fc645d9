/*
 * mixed3.cuh -- streaming kernel for real-family transforms whose underlying real FFT has length M = R0*R1*R2 with three
 * odd prime radices: 1001 = 13*11*7, the "generic radix 7.11.13" lengths of BASELINE config 4 (cosqmf_/cosqmb_ N = 1001,
 * sintmf_/sintmb_ N = 1000, rfftmf_/rfftmb_ N = 1001).  It replaces, for those shapes, the pass loops mrftf1_/mrftb1_
 * (cfftpack/fftpack.c:10149, :9946) with their generic passes mradfg/mradbg (:9443, :8106 -- the O(ip) per output
 * r1fgkf_ scheme, :12564) and the pre/post sweeps of mcsqf1_/mcsqb1_ (:6839, :6740) and msntf1_/msntb1_ (:10636, :10530).
 *
 * Design (one HBM read and one HBM write per element, both by the bulk-copy engine):
 *   - one PAIR of contiguous rows per tile, z = x_a + i x_b, one complex transform of length M per pair;
 *   - persistent CTAs; the next pair (2n doubles, one cp.async.bulk) lands in shared memory while the current one is
 *     transformed (mbarrier complete_tx);
 *   - three Stockham passes 13, 11, 7 with the butterflies in registers (dft_odd, butterfly.cuh) ping-ponging between
 *     two shared-memory buffers, one __syncthreads per pass; 16-byte elements at odd strides: bank-conflict free;
 *   - the family's pre-processing (cosqf1_ fold :5693-5712, sintf1_ fold :14873-14890, half-complex -> spectrum of
 *     rfftb1_) is computed on the way from the landing buffer into the registers of the first pass;
 *   - the Hermitian split and the family's post-processing write the finished rows, in their final layout, into the
 *     free buffer, and ONE cp.async.bulk shared->global per pair drains it (no per-thread global stores);
 *   - sintf1_'s serial running sum (:14905-14915) is a warp scan over a padded copy of the 500 terms.
 * Twiddles w^k (k < radix) are rebuilt from the two table rows w^p, w^4p kept in shared memory.
 */
#ifndef CFB_MIXED3_CUH
#define CFB_MIXED3_CUH
#include "butterfly.cuh"
#include "engine_types.h"
#include "internal.h"
#include "tma.cuh"

namespace cfb {

template <int R0_, int R1_, int R2_>
struct M3Cfg {
  static constexpr int R0 = R0_, R1 = R1_, R2 = R2_;
  static constexpr int M = R0 * R1 * R2, H = (M - 1) / 2;
  static constexpr int NB0 = M / R0, NB1 = M / R1, NB2 = M / R2;  // butterflies per pass
  static constexpr int MM1 = R2;                                  // sub-transform count of pass 1 (twiddle period)
  // table (cpx): pass 0 rows w_M^p, w_M^4p (p < NB0); pass 1 rows w_{R1 R2}^p, ^4p (p < R2); roots of unity of R0, R1, R2
  static constexpr int T0 = 0, T1 = 2 * NB0, RT0 = T1 + 2 * MM1, RT1 = RT0 + R0, RT2 = RT1 + R1, TAB = RT2 + R2;
  static constexpr int SCAN_PITCH = (H + 1) + (H + 1) / 16 + 8;  // sint: padded copy of one row's 500 running-sum terms
  static constexpr size_t BUF = ((size_t)M * sizeof(cpx) + 15) / 16 * 16;  // >= 2n doubles for n <= M
  static constexpr size_t OFF_L = 0, OFF_P = BUF, OFF_Q = 2 * BUF, OFF_TAB = 3 * BUF, OFF_BAR = OFF_TAB + (size_t)TAB * sizeof(cpx);
  static constexpr size_t BYTES = OFF_BAR + 16;
};

/* a[k] *= w^k (k = 1..R-1), w^k from w1 = w and w4 = w^4: at most three products deep */
template <int R, int DIR>
__device__ __forceinline__ void twiddle_powers_r(cpx (&a)[R], cpx w1, cpx w4) {
  cpx lo[4];
  lo[1] = w1;
  lo[2] = cmul(w1, w1);
  lo[3] = cmul(lo[2], w1);
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

/* one Stockham pass: butterfly b = q + S p reads elements b + NB i (i < R) through `load`, writes q + S R p + S k */
template <int R, int S, int MM, int W, int DIR, class Load>
__device__ __forceinline__ void m3_pass(const Load &load, cpx *__restrict__ dst, const cpx *__restrict__ tw,
                                        const cpx *__restrict__ rt, const int t) {
  constexpr int NB = S * MM;
#pragma unroll 1
  for (int b = t; b < NB; b += W) {
    const int p = b / S, q = b - p * S;
    cpx a[R];
#pragma unroll
    for (int i = 0; i < R; ++i) a[i] = load(b + NB * i);
    dft_odd<R, DIR>(a, rt);
    if (MM > 1) twiddle_powers_r<R, DIR>(a, tw[p], tw[MM + p]);
    cpx *d = dst + q + S * R * p;
#pragma unroll
    for (int k = 0; k < R; ++k) d[S * k] = a[k];
  }
}

/* KIND: K_RFFT (n = M), K_COSQ (n = M), K_SINT (n = M - 1).  DIR: -1 forward, +1 backward (user-level direction). */
template <class C, int KIND, int DIR, int W>
__global__ void __launch_bounds__(W, (W <= 96 ? 4 : W <= 128 ? 3 : 2)) m3_stream_kernel(double *__restrict__ x, long long npairs, const cpx *__restrict__ tab_g,
                                                         const double *__restrict__ trig) {
  CFB_DYN_SMEM(smem_raw);
  constexpr int M = C::M, H = C::H, n = (KIND == K_SINT) ? M - 1 : M;
  constexpr int CD = (KIND == K_SINT) ? -1 : DIR;  // direction of the complex transform (sintb1_ runs rfftf too)
  constexpr unsigned PAIR_BYTES = 2u * n * sizeof(double);
  static_assert(PAIR_BYTES % 16 == 0 && PAIR_BYTES <= C::BUF, "a pair of rows is one aligned bulk copy");
  static_assert(KIND != K_SINT || W >= 64, "the running sums of the two rows take one warp each");
  double *land = (double *)(smem_raw + C::OFF_L);
  cpx *P = (cpx *)(smem_raw + C::OFF_P), *Q = (cpx *)(smem_raw + C::OFF_Q);
  cpx *tab = (cpx *)(smem_raw + C::OFF_TAB);
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
    /* ---- pass 0 (radix R0): the family's pre-processing happens in the loader ---- */
    auto load0 = [&](int e) -> cpx {
      if (KIND == K_RFFT && DIR < 0) return make_double2(la[e], lb[e]);
      if (KIND == K_SINT) {  // sintf1_: xh[0] = 0, xh[k] = t1 + t2, xh[M-k] = t2 - t1 (k = 1..n/2)
        if (e == 0) return make_double2(0.0, 0.0);
        const int k = e <= H ? e : M - e;
        const double s = __ldg(trig + k - 1);
        const double t1a = la[k - 1] - la[n - k], t2a = s * (la[k - 1] + la[n - k]);
        const double t1b = lb[k - 1] - lb[n - k], t2b = s * (lb[k - 1] + lb[n - k]);
        return e <= H ? make_double2(t1a + t2a, t1b + t2b) : make_double2(t2a - t1a, t2b - t1b);
      }
      if (KIND == K_COSQ && DIR < 0) {  // cosqf1_ fold
        if (e == 0) return make_double2(la[0], lb[0]);
        const int j = e <= H ? e : M - e, jc = M - j;
        const double wj = __ldg(trig + j - 1), wc = __ldg(trig + jc - 1);
        const double sa = la[j] + la[jc], da = la[j] - la[jc], sb = lb[j] + lb[jc], db = lb[j] - lb[jc];
        return e <= H ? make_double2(fma(wj, da, wc * sa), fma(wj, db, wc * sb))
                      : make_double2(fma(wj, sa, -(wc * da)), fma(wj, sb, -(wc * db)));
      }
      // backward: half-complex rows h_a, h_b -> spectrum Z of z = x_a + i x_b (rfftb1_ convention); cosqb1_ first forms
      // h[0] = x[0]/2, h[2f-1] = (x[2f-1] + x[2f])/2, h[2f] = (x[2f-1] - x[2f])/2
      const double edge = (KIND == K_COSQ) ? 0.5 : 1.0;
      if (e == 0) return make_double2(edge * la[0], edge * lb[0]);
      const int f = e <= H ? e : M - e;
      double h1a = la[2 * f - 1], h2a = la[2 * f], h1b = lb[2 * f - 1], h2b = lb[2 * f];
      if (KIND == K_COSQ) {
        const double s1 = 0.5 * (h1a + h2a), d1 = 0.5 * (h1a - h2a), s2 = 0.5 * (h1b + h2b), d2 = 0.5 * (h1b - h2b);
        h1a = s1;
        h2a = d1;
        h1b = s2;
        h2b = d2;
      }
      const double a1 = 0.5 * h1a, a2 = 0.5 * h2a, b1 = 0.5 * h1b, b2 = 0.5 * h2b;
      return e <= H ? make_double2(a1 + b2, b1 - a2) : make_double2(a1 - b2, b1 + a2);
    };
    m3_pass<C::R0, 1, C::NB0, W, CD>(load0, P, tab + C::T0, tab + C::RT0, t);
    // the previous pair's bulk store must have finished reading Q before pass 1 overwrites it
    if (t == 0) bulk_wait_read();
    __syncthreads();  // landing buffer consumed (pass 0 has turned every read into results stored in P): refill it with the next pair while this one is transformed
    const long long next = tile + gridDim.x;
    if (t == 0 && next < npairs) {
      mbar_expect_tx(bar, PAIR_BYTES);
      bulk_g2s(land, x + next * 2 * n, PAIR_BYTES, bar);
    }
    /* ---- pass 1 (radix R1): P -> Q;  pass 2 (radix R2): Q -> P, natural order ---- */
    m3_pass<C::R1, C::R0, C::MM1, W, CD>([&](int e) -> cpx { return P[e]; }, Q, tab + C::T1, tab + C::RT1, t);
    __syncthreads();
    m3_pass<C::R2, C::R0 * C::R1, 1, W, CD>([&](int e) -> cpx { return Q[e]; }, P, nullptr, tab + C::RT2, t);
    __syncthreads();
    /* ---- split / post-processing: P -> finished rows in Q ---- */
    if (CD < 0) {
      // FFTPACK's half-complex row [X0, A1, B1, ...]: A_f = 2 Re X_f / M, B_f = -2 Im X_f / M (rfftf1_ epilogue :13818-13853)
      const double sc = 1.0 / (double)M;
      const double ss = DIR < 0 ? 0.5 : 0.25 * (double)M;  // sint only: sintf1_ / sintb1_ output factor
      double keepA_a[(H + W) / W], keepA_b[(H + W) / W];   // sint: running-sum terms, stored after the barrier
#pragma unroll
      for (int it = 0; it < (H + W) / W; ++it) {
        const int f = t + it * W;
        keepA_a[it] = keepA_b[it] = 0.0;
        if (f > H) continue;
        const cpx u = P[f];
        if (f == 0) {
          if (KIND == K_SINT) {
            keepA_a[it] = ss * (u.x * sc);
            keepA_b[it] = ss * (u.y * sc);
          } else {
            qa[0] = u.x * sc;
            qb[0] = u.y * sc;
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
        } else {  // sintf1_ post :14897-14919: y[2f-1] = ss B_f, y[2f] = ss (h0 + A_1 + ... + A_f)
          qa[2 * f - 1] = ss * Ba;
          qb[2 * f - 1] = ss * Bb;
          keepA_a[it] = ss * Aa;
          keepA_b[it] = ss * Ab;
        }
      }
      if (KIND == K_SINT) {
        constexpr int SP = C::SCAN_PITCH, NT = n / 2;  // terms g = 0..NT-1 -> y[2g]
        double *sa = (double *)P, *sb = sa + SP;
        __syncthreads();  // every thread has read its spectrum values: P becomes the scan scratch
#pragma unroll
        for (int it = 0; it < (H + W) / W; ++it) {
          const int g = t + it * W;
          if (g < NT) {
            sa[g + (g >> 4)] = keepA_a[it];
            sb[g + (g >> 4)] = keepA_b[it];
          }
        }
        __syncthreads();
        if (t < 64) {  // warp 0: row a, warp 1: row b; lane l owns terms 16 l .. 16 l + 15 (pitch 17: conflict-free)
          double *s = (t < 32 ? sa : sb) + 17 * (t & 31);
          const int g0 = 16 * (t & 31);
          double v[16], run = 0.0;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] = (g0 + i < NT) ? s[i] : 0.0;
            run += v[i];
            v[i] = run;
          }
          const double off = warp_excl_scan(run, t & 31);
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (g0 + i < NT) s[i] = v[i] + off;
        }
        __syncthreads();
        for (int g = t; g < NT; g += W) {
          qa[2 * g] = sa[g + (g >> 4)];
          qb[2 * g] = sb[g + (g >> 4)];
        }
      }
    } else if (KIND == K_COSQ) {
      // cosqb1_ post :5626-5652: y[0] = 2 u[0]; y[j] = p + q, y[n-j] = p - q,
      //   p = W[j-1] u[n-j] + W[n-j-1] u[j], q = W[j-1] u[j] - W[n-j-1] u[n-j]
      for (int j = t; j <= H; j += W) {
        const cpx uj = P[j];
        if (j == 0) {
          qa[0] = uj.x + uj.x;
          qb[0] = uj.y + uj.y;
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

/* host side (mixed3.cu) */
bool m3_supported(int kind, int n);
/* `npairs` pairs of contiguous rows (jump = n, 16-byte aligned base); trig = the family's table (nullptr for rfft) */
bool m3_launch(int kind, int n, long long npairs, int dir, double *x, const double *trig);
void m3_release_tables();

}  // namespace cfb
#endif
