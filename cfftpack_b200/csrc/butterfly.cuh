/*
 * butterfly.cuh -- in-register DFT butterflies (FP64), sizes 2, 3, 4, 5, 8, 16.
 *
 * These take the place of the radix kernels of the reference's complex passes
 * (cfftpack/fftpack.c: c1f2kf_:195, c1f3kf_:443, c1f4kf_:752, c1f5kf_:1145 and the *kb_ twins);
 * 8 and 16 are two fused radix-2/4 passes kept in registers.  DIR = -1 is the forward
 * transform (e^{-i...}), +1 the backward one.  Outputs are in natural order, unscaled.
 */
#ifndef CFB_BUTTERFLY_CUH
#define CFB_BUTTERFLY_CUH
#include "cfb_rt.h"

namespace cfb {

typedef double2 cpx;

__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return make_double2(a.x - b.x, a.y - b.y); }
/* a * w */
__device__ __forceinline__ cpx cmul(cpx a, cpx w) {
  return make_double2(fma(a.x, w.x, -(a.y * w.y)), fma(a.x, w.y, a.y * w.x));
}
/* a * conj(w) */
__device__ __forceinline__ cpx cmulc(cpx a, cpx w) {
  return make_double2(fma(a.x, w.x, a.y * w.y), fma(a.y, w.x, -(a.x * w.y)));
}
/* tables hold the FORWARD twiddle w = exp(-i theta); the backward transform uses its conjugate */
template <int DIR>
__device__ __forceinline__ cpx ctw(cpx a, cpx w) {
  return DIR < 0 ? cmul(a, w) : cmulc(a, w);
}
/* multiply by DIR * i  (forward: -i, backward: +i) */
template <int DIR>
__device__ __forceinline__ cpx mul_dir_i(cpx a) {
  return DIR < 0 ? make_double2(a.y, -a.x) : make_double2(-a.y, a.x);
}

template <int DIR>
__device__ __forceinline__ void dft2(cpx &a0, cpx &a1) {
  cpx t = a0;
  a0 = cadd(t, a1);
  a1 = csub(t, a1);
}

template <int DIR>
__device__ __forceinline__ void dft3(cpx &a0, cpx &a1, cpx &a2) {
  const double C = 0.86602540378443864676372317075294;  // sin(pi/3)
  cpx s = cadd(a1, a2), d = csub(a1, a2);
  cpx m = make_double2(fma(-0.5, s.x, a0.x), fma(-0.5, s.y, a0.y));
  cpx e = mul_dir_i<DIR>(make_double2(C * d.x, C * d.y));
  a0 = cadd(a0, s);
  a1 = cadd(m, e);
  a2 = csub(m, e);
}

template <int DIR>
__device__ __forceinline__ void dft4(cpx &a0, cpx &a1, cpx &a2, cpx &a3) {
  cpx t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_dir_i<DIR>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, t3);
  a3 = csub(t1, t3);
}

template <int DIR>
__device__ __forceinline__ void dft5(cpx &a0, cpx &a1, cpx &a2, cpx &a3, cpx &a4) {
  const double C1 = 0.30901699437494742410229341718282;   // cos(2pi/5)
  const double C2 = -0.80901699437494742410229341718282;  // cos(4pi/5)
  const double S1 = 0.95105651629515357211643933337938;   // sin(2pi/5)
  const double S2 = 0.58778525229247312916870595463907;   // sin(4pi/5)
  cpx p1 = cadd(a1, a4), m1 = csub(a1, a4), p2 = cadd(a2, a3), m2 = csub(a2, a3);
  cpx c1 = make_double2(fma(C2, p2.x, fma(C1, p1.x, a0.x)), fma(C2, p2.y, fma(C1, p1.y, a0.y)));
  cpx c2 = make_double2(fma(C1, p2.x, fma(C2, p1.x, a0.x)), fma(C1, p2.y, fma(C2, p1.y, a0.y)));
  cpx s1 = mul_dir_i<DIR>(make_double2(fma(S2, m2.x, S1 * m1.x), fma(S2, m2.y, S1 * m1.y)));
  cpx s2 = mul_dir_i<DIR>(make_double2(fma(-S1, m2.x, S2 * m1.x), fma(-S1, m2.y, S2 * m1.y)));
  a0 = make_double2(a0.x + p1.x + p2.x, a0.y + p1.y + p2.y);
  a1 = cadd(c1, s1);
  a4 = csub(c1, s1);
  a2 = cadd(c2, s2);
  a3 = csub(c2, s2);
}

/* multiply by exp(DIR * i * pi/4 * k), k = 1, 3 */
template <int DIR>
__device__ __forceinline__ cpx mul_w8_1(cpx a) {
  const double H = 0.70710678118654752440084436210485;
  return DIR < 0 ? make_double2(H * (a.x + a.y), H * (a.y - a.x)) : make_double2(H * (a.x - a.y), H * (a.y + a.x));
}
template <int DIR>
__device__ __forceinline__ cpx mul_w8_3(cpx a) {
  const double H = 0.70710678118654752440084436210485;
  return DIR < 0 ? make_double2(H * (a.y - a.x), -H * (a.x + a.y)) : make_double2(-H * (a.x + a.y), H * (a.x - a.y));
}

/* 8-point DFT, natural order in and out: 2 x dft4 (even/odd inputs) + one radix-2 level */
template <int DIR>
__device__ __forceinline__ void dft8(cpx (&a)[8]) {
  dft4<DIR>(a[0], a[2], a[4], a[6]);
  dft4<DIR>(a[1], a[3], a[5], a[7]);
  cpx o1 = mul_w8_1<DIR>(a[3]), o2 = mul_dir_i<DIR>(a[5]), o3 = mul_w8_3<DIR>(a[7]);
  cpx e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1];
  a[0] = cadd(e0, o0);
  a[4] = csub(e0, o0);
  a[1] = cadd(e1, o1);
  a[5] = csub(e1, o1);
  a[2] = cadd(e2, o2);
  a[6] = csub(e2, o2);
  a[3] = cadd(e3, o3);
  a[7] = csub(e3, o3);
}

/* 16-point DFT, natural order in and out: 4 x dft4 over stride-4 inputs, w16 twiddles, 4 x dft4 */
template <int DIR>
__device__ __forceinline__ void dft16(cpx (&a)[16]) {
  const double C1 = 0.92387953251128675612818318939679;  // cos(pi/8)
  const double S1 = 0.38268343236508977172845998403040;  // sin(pi/8)
  // forward twiddles w16^m = (cos(m pi/8), -sin(m pi/8))
  const cpx w1 = make_double2(C1, -S1), w3 = make_double2(S1, -C1);
#pragma unroll
  for (int i = 0; i < 4; ++i) dft4<DIR>(a[i], a[i + 4], a[i + 8], a[i + 12]);
  // after this, a[i + 4*k1] = B_i[k1]; multiply by w16^{i*k1}
  a[5] = ctw<DIR>(a[5], w1);            // i=1,k1=1
  a[9] = mul_w8_1<DIR>(a[9]);           // i=1,k1=2 -> w16^2 = w8^1
  a[13] = ctw<DIR>(a[13], w3);          // i=1,k1=3
  a[6] = mul_w8_1<DIR>(a[6]);           // i=2,k1=1 -> w16^2
  a[10] = mul_dir_i<DIR>(a[10]);        // i=2,k1=2 -> w16^4
  a[14] = mul_w8_3<DIR>(a[14]);         // i=2,k1=3 -> w16^6 = w8^3
  a[7] = ctw<DIR>(a[7], w3);            // i=3,k1=1 -> w16^3
  a[11] = mul_w8_3<DIR>(a[11]);         // i=3,k1=2 -> w16^6
  {                                     // i=3,k1=3 -> w16^9 = -w16^1
    cpx t = ctw<DIR>(a[15], w1);
    a[15] = make_double2(-t.x, -t.y);
  }
  // X[k1 + 4*k2] = sum_i a[i + 4*k1] w4^{i*k2}
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4<DIR>(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
  // now a[4*k1 + k2] = X[k1 + 4*k2]: transpose the 4x4 to natural order
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
    for (int k2 = k1 + 1; k2 < 4; ++k2) {
      cpx t = a[4 * k1 + k2];
      a[4 * k1 + k2] = a[4 * k2 + k1];
      a[4 * k2 + k1] = t;
    }
}

template <int R, int DIR>
struct Dft;
template <int DIR>
struct Dft<2, DIR> {
  static __device__ __forceinline__ void run(cpx (&a)[2]) { dft2<DIR>(a[0], a[1]); }
};
template <int DIR>
struct Dft<3, DIR> {
  static __device__ __forceinline__ void run(cpx (&a)[3]) { dft3<DIR>(a[0], a[1], a[2]); }
};
template <int DIR>
struct Dft<4, DIR> {
  static __device__ __forceinline__ void run(cpx (&a)[4]) { dft4<DIR>(a[0], a[1], a[2], a[3]); }
};
template <int DIR>
struct Dft<5, DIR> {
  static __device__ __forceinline__ void run(cpx (&a)[5]) { dft5<DIR>(a[0], a[1], a[2], a[3], a[4]); }
};
template <int DIR>
struct Dft<8, DIR> {
  static __device__ __forceinline__ void run(cpx (&a)[8]) { dft8<DIR>(a); }
};
template <int DIR>
struct Dft<16, DIR> {
  static __device__ __forceinline__ void run(cpx (&a)[16]) { dft16<DIR>(a); }
};

/* odd prime R (7, 11, 13) in registers: symmetric sums, (R-1)^2/2 complex-by-real products.
 * rt[j] = exp(-2 pi i j / R) (forward convention), j < R, typically in shared memory */
template <int R, int DIR>
__device__ __forceinline__ void dft_odd(cpx (&a)[R], const cpx *__restrict__ rt) {
  constexpr int H = (R - 1) / 2;
  double c[H + 1], sn[H + 1];
#pragma unroll
  for (int j = 1; j <= H; ++j) {
    const cpx w = rt[j];
    c[j] = w.x;
    sn[j] = -w.y;  // sin(2 pi j / R)
  }
  cpx p[H + 1], m[H + 1];
  cpx x0 = a[0];
#pragma unroll
  for (int j = 1; j <= H; ++j) {
    p[j] = cadd(a[j], a[R - j]);
    m[j] = csub(a[j], a[R - j]);
    x0 = cadd(x0, p[j]);
  }
  const cpx a0 = a[0];
#pragma unroll
  for (int k = 1; k <= H; ++k) {
    double ar = a0.x, ai = a0.y, br = 0.0, bi = 0.0;
#pragma unroll
    for (int j = 1; j <= H; ++j) {
      constexpr int dummy = 0;
      (void)dummy;
      const int idx = (j * k) % R;                 // compile-time after unrolling
      const int h = idx <= H ? idx : R - idx;      // cos is even, sin is odd about R/2
      const double cc = c[h], ss = idx <= H ? sn[h] : -sn[h];
      ar = fma(cc, p[j].x, ar);
      ai = fma(cc, p[j].y, ai);
      br = fma(ss, m[j].x, br);
      bi = fma(ss, m[j].y, bi);
    }
    // X_k = A + DIR*i*B, X_{R-k} = A - DIR*i*B with B = (br, bi)
    if (DIR < 0) {
      a[k] = make_double2(ar + bi, ai - br);
      a[R - k] = make_double2(ar - bi, ai + br);
    } else {
      a[k] = make_double2(ar - bi, ai + br);
      a[R - k] = make_double2(ar + bi, ai - br);
    }
  }
  a[0] = x0;
}

/* composite radix R = A*B in registers (6 = 2*3, 9 = 3*3, 10 = 2*5): B-point... two levels with the R-th roots from
 * rt[j] = exp(-2 pi i j / R).  Natural order in and out. */
template <int A, int B, int DIR>
__device__ __forceinline__ void dft_comp(cpx (&a)[A * B], const cpx *__restrict__ rt) {
  constexpr int R = A * B;
  // X[k1 + A*k2] = sum_{i<B} w_R^{i k1} w_B^{i k2} [ sum_{j<A} x[i + B j] w_A^{j k1} ],  k1 < A, k2 < B
  cpx t[R];
#pragma unroll
  for (int i = 0; i < B; ++i) {
    cpx u[A];
#pragma unroll
    for (int j = 0; j < A; ++j) u[j] = a[i + B * j];
    Dft<A, DIR>::run(u);
#pragma unroll
    for (int k1 = 0; k1 < A; ++k1) t[k1 * B + i] = (i * k1 == 0) ? u[k1] : ctw<DIR>(u[k1], rt[(i * k1) % R]);
  }
#pragma unroll
  for (int k1 = 0; k1 < A; ++k1) {
    cpx v[B];
#pragma unroll
    for (int i = 0; i < B; ++i) v[i] = t[k1 * B + i];
    Dft<B, DIR>::run(v);
#pragma unroll
    for (int k2 = 0; k2 < B; ++k2) a[k1 + A * k2] = v[k2];
  }
}

/* ---- warp helpers (one warp works on one real sequence in the pre/post phases) ---- */
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
/* exclusive prefix over the lanes of a warp */
__device__ __forceinline__ double warp_excl_scan(double v, int lane) {
  double inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    double u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  return inc - v;
}

/* uniform entry point for the engine: radices that need the table of R-th roots take it from rt */
template <int R, int DIR>
struct DftRt {
  static __device__ __forceinline__ void run(cpx (&a)[R], const cpx *) { Dft<R, DIR>::run(a); }
};
#define CFB_DFT_ODD(R)                                                                                  \
  template <int DIR>                                                                                     \
  struct DftRt<R, DIR> {                                                                                 \
    static __device__ __forceinline__ void run(cpx (&a)[R], const cpx *rt) { dft_odd<R, DIR>(a, rt); }   \
  };
CFB_DFT_ODD(7)
CFB_DFT_ODD(11)
CFB_DFT_ODD(13)
#define CFB_DFT_COMP(A, B)                                                                                        \
  template <int DIR>                                                                                               \
  struct DftRt<(A) * (B), DIR> {                                                                                   \
    static __device__ __forceinline__ void run(cpx (&a)[(A) * (B)], const cpx *rt) { dft_comp<A, B, DIR>(a, rt); } \
  };
CFB_DFT_COMP(2, 3)
CFB_DFT_COMP(3, 3)
CFB_DFT_COMP(2, 5)

}  // namespace cfb
#endif
