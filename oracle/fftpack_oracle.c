/*
 * oracle/fftpack_oracle.c -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 * (see fftpack_oracle.h for the rules on who may call this and the pin status)
 *
 * Restates, in plain C, the FFTPACK 5.1 algorithms behind the hot path of
 * zywina/cfftpack.  file:line citations are into /root/reference/cfftpack/fftpack.c.
 *
 *  - factorisation, wsave layouts and twiddle tables follow the reference
 *    expression by expression (they are compared BITWISE with oracle/_ref);
 *  - the complex passes are the reference's decimation-in-frequency Stockham
 *    passes cc(l1,ido,ip) -> ch(l1,ip,ido) with the radix-2/3/4/5 butterflies
 *    written out and a symmetric O(ip^2) generic-prime pass;
 *  - the real passes keep the reference's pass order, ping-pong and
 *    half-complex layout cc(ido,l1,ip) <-> ch(ido,ip,l1) but evaluate each
 *    pass from its definition (combine ip half-complex spectra of length ido
 *    into one of length ido*ip), which is the same arithmetic contract as
 *    r1f{2,3,4,5,g}k{f,b} without their hand-unrolled forms;
 *  - cost/sint/cosq/sinq pre- and post-processing follow the reference loops.
 */
#include "fftpack_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef orc_complex_t cpx;

/* the literal lensav expression, fftpack.c:2221 */
static int il2(int n) { return (int)(log((double)n) / log(2.0)); }

/* ------------------------------------------------------------------ */
/* factorisation: fftpack.c:6613-6657 (factor_)                        */
int orc_factor(int n, int *fac) {
  static const int ntryh[4] = {4, 2, 3, 5};
  int nl = n, nf = 0, j = 0, ntry = 0;
  while (nl != 1) {
    ntry = (j < 4) ? ntryh[j] : ntry + 2;
    ++j;
    while (nl % ntry == 0) {
      fac[nf++] = ntry;
      nl /= ntry;
    }
  }
  return nf;
}

/* real variant: a factor 2 is moved to the front, fftpack.c:13892-13935 */
int orc_rfactor(int n, int *fac) {
  static const int ntryh[4] = {4, 2, 3, 5};
  int nl = n, nf = 0, j = 0, ntry = 0;
  while (nl != 1) {
    ntry = (j < 4) ? ntryh[j] : ntry + 2;
    ++j;
    while (nl % ntry == 0) {
      fac[nf++] = ntry;
      nl /= ntry;
      if (ntry == 2 && nf != 1) {
        for (int i = nf - 1; i > 0; --i) fac[i] = fac[i - 1];
        fac[0] = 2;
      }
    }
  }
  return nf;
}

/* fftpack.c:15210-15258 (xercon_): 1 = consistent */
int orc_xercon(int inc, int jump, int n, int lot) {
  int i = inc, j = jump;
  while (j != 0) {
    int jnew = i % j;
    i = j;
    j = jnew;
  }
  int lcm = inc * jump / i;
  return !(lcm <= (n - 1) * inc && lcm <= (lot - 1) * jump);
}

/* ------------------------------------------------------------------ */
/* complex init: fftpack.c:6666-6697 (mcfti1_) + :15124-15166 (tables_) */
static void c_init(int n, double *wsave) {
  int fac[64];
  int nf = orc_factor(n, fac);
  double *wa = wsave;
  wsave[2 * n] = (double)nf;
  for (int k = 0; k < nf; ++k) wsave[2 * n + 1 + k] = (double)fac[k];
  double tpi = atan(1.0) * 8.0;
  int iw = 0, l1 = 1;
  for (int k1 = 0; k1 < nf; ++k1) {
    int ip = fac[k1], l2 = l1 * ip, ido = n / l2;
    double argz = tpi / (double)ip;
    double arg1 = tpi / (double)(ido * ip);
    for (int j = 1; j < ip; ++j) {
      double arg2 = (double)j * arg1;
      for (int i = 0; i < ido; ++i) {
        double arg3 = (double)i * arg2;
        wa[iw + i + (j - 1) * ido] = cos(arg3);
        wa[iw + i + (j - 1) * ido + (ip - 1) * ido] = sin(arg3);
      }
      if (ip > 5) {
        double arg4 = (double)j * argz;
        wa[iw + (j - 1) * ido] = cos(arg4);
        wa[iw + (j - 1) * ido + (ip - 1) * ido] = sin(arg4);
      }
    }
    iw += (ip - 1) * (ido + ido);
    l1 = l2;
  }
}

/* one complex Stockham DIF pass, fftpack.c:96-1930 (c1f{2,3,4,5,g}k{f,b}).
 * cc(l1,ido,ip) stride ics -> ch(l1,ip,ido) stride ihs.  sgn=-1 forward,
 * +1 backward.  scale!=0: multiply outputs by sn (forward last pass). */
#define CCX(k, i, j) cc[(size_t)((k) + l1 * ((i) + ido * (j))) * ics]
#define CHX(k, j, i) ch[(size_t)((k) + l1 * ((j) + ip * (i))) * ihs]
static inline cpx twid(const double *wa, int ido, int ip, int i, int j, int sgn, cpx t) {
  /* forward multiplies by conj(w), backward by w (fftpack.c:301-304) */
  double wr = wa[i + (j - 1) * ido], wi = wa[i + (j - 1) * ido + (ip - 1) * ido];
  cpx o;
  if (sgn < 0) {
    o.r = wr * t.r + wi * t.i;
    o.i = wr * t.i - wi * t.r;
  } else {
    o.r = wr * t.r - wi * t.i;
    o.i = wr * t.i + wi * t.r;
  }
  return o;
}

static void c_pass(int ido, int l1, int ip, const cpx *cc, int ics, cpx *ch, int ihs, const double *wa, int sgn,
                   double sn) {
  const int last = (ido == 1);
  cpx o[5];
  if (ip == 2) {
    for (int i = 0; i < ido; ++i)
      for (int k = 0; k < l1; ++k) {
        cpx a = CCX(k, i, 0), b = CCX(k, i, 1);
        o[0].r = a.r + b.r; o[0].i = a.i + b.i;
        o[1].r = a.r - b.r; o[1].i = a.i - b.i;
        if (last) {
          if (sgn < 0) { o[0].r *= sn; o[0].i *= sn; o[1].r *= sn; o[1].i *= sn; }
        } else if (i > 0) {
          o[1] = twid(wa, ido, ip, i, 1, sgn, o[1]);
        }
        CHX(k, 0, i) = o[0]; CHX(k, 1, i) = o[1];
      }
  } else if (ip == 3) {
    const double taur = -.5, taui = (sgn < 0) ? -.866025403784439 : .866025403784439; /* :318-319, :448-449 */
    for (int i = 0; i < ido; ++i)
      for (int k = 0; k < l1; ++k) {
        cpx c1 = CCX(k, i, 0), c2 = CCX(k, i, 1), c3 = CCX(k, i, 2);
        double tr2 = c2.r + c3.r, cr2 = c1.r + taur * tr2;
        double ti2 = c2.i + c3.i, ci2 = c1.i + taur * ti2;
        double cr3 = taui * (c2.r - c3.r), ci3 = taui * (c2.i - c3.i);
        o[0].r = c1.r + tr2; o[0].i = c1.i + ti2;
        o[1].r = cr2 - ci3; o[2].r = cr2 + ci3;
        o[1].i = ci2 + cr3; o[2].i = ci2 - cr3;
        for (int j = 0; j < 3; ++j) {
          if (last) { if (sgn < 0) { o[j].r *= sn; o[j].i *= sn; } }
          else if (i > 0 && j > 0) o[j] = twid(wa, ido, ip, i, j, sgn, o[j]);
          CHX(k, j, i) = o[j];
        }
      }
  } else if (ip == 4) {
    for (int i = 0; i < ido; ++i)
      for (int k = 0; k < l1; ++k) {
        cpx c1 = CCX(k, i, 0), c2 = CCX(k, i, 1), c3 = CCX(k, i, 2), c4 = CCX(k, i, 3);
        double ti1 = c1.i - c3.i, ti2 = c1.i + c3.i, ti3 = c2.i + c4.i;
        double tr1 = c1.r - c3.r, tr2 = c1.r + c3.r, tr3 = c2.r + c4.r;
        double tr4 = (sgn < 0) ? c2.i - c4.i : c4.i - c2.i; /* :752 vs :604 */
        double ti4 = (sgn < 0) ? c4.r - c2.r : c2.r - c4.r;
        o[0].r = tr2 + tr3; o[0].i = ti2 + ti3;
        o[2].r = tr2 - tr3; o[2].i = ti2 - ti3;
        o[1].r = tr1 + tr4; o[1].i = ti1 + ti4;
        o[3].r = tr1 - tr4; o[3].i = ti1 - ti4;
        for (int j = 0; j < 4; ++j) {
          if (last) { if (sgn < 0) { o[j].r *= sn; o[j].i *= sn; } }
          else if (i > 0 && j > 0) o[j] = twid(wa, ido, ip, i, j, sgn, o[j]);
          CHX(k, j, i) = o[j];
        }
      }
  } else if (ip == 5) {
    const double tr11 = .3090169943749474, tr12 = -.8090169943749474; /* :942-945, :1150-1153 */
    const double ti11 = (sgn < 0) ? -.9510565162951536 : .9510565162951536;
    const double ti12 = (sgn < 0) ? -.5877852522924731 : .5877852522924731;
    for (int i = 0; i < ido; ++i)
      for (int k = 0; k < l1; ++k) {
        cpx c1 = CCX(k, i, 0), c2 = CCX(k, i, 1), c3 = CCX(k, i, 2), c4 = CCX(k, i, 3), c5 = CCX(k, i, 4);
        double ti5 = c2.i - c5.i, ti2 = c2.i + c5.i, ti4 = c3.i - c4.i, ti3 = c3.i + c4.i;
        double tr5 = c2.r - c5.r, tr2 = c2.r + c5.r, tr4 = c3.r - c4.r, tr3 = c3.r + c4.r;
        double cr2 = c1.r + tr11 * tr2 + tr12 * tr3, ci2 = c1.i + tr11 * ti2 + tr12 * ti3;
        double cr3 = c1.r + tr12 * tr2 + tr11 * tr3, ci3 = c1.i + tr12 * ti2 + tr11 * ti3;
        double cr5 = ti11 * tr5 + ti12 * tr4, ci5 = ti11 * ti5 + ti12 * ti4;
        double cr4 = ti12 * tr5 - ti11 * tr4, ci4 = ti12 * ti5 - ti11 * ti4;
        o[0].r = c1.r + tr2 + tr3; o[0].i = c1.i + ti2 + ti3;
        o[1].r = cr2 - ci5; o[1].i = ci2 + cr5;
        o[2].r = cr3 - ci4; o[2].i = ci3 + cr4;
        o[3].r = cr3 + ci4; o[3].i = ci3 - cr4;
        o[4].r = cr2 + ci5; o[4].i = ci2 - cr5;
        for (int j = 0; j < 5; ++j) {
          if (last) { if (sgn < 0) { o[j].r *= sn; o[j].i *= sn; } }
          else if (i > 0 && j > 0) o[j] = twid(wa, ido, ip, i, j, sgn, o[j]);
          CHX(k, j, i) = o[j];
        }
      }
  } else {
    /* generic odd factor, fftpack.c:1410-1930: symmetric half sums, cosine
     * part and sine part from the roots stored at wa(1,j,.) (idlj = l*j mod ip, :1758) */
    int ipph = (ip + 1) / 2;
    cpx *tp = (cpx *)malloc(sizeof(cpx) * 2 * ip), *tm = tp + ip;
    for (int i = 0; i < ido; ++i)
      for (int k = 0; k < l1; ++k) {
        cpx c0 = CCX(k, i, 0);
        cpx s0 = c0;
        for (int j = 1; j < ipph; ++j) {
          cpx a = CCX(k, i, j), b = CCX(k, i, ip - j);
          tp[j].r = a.r + b.r; tp[j].i = a.i + b.i;
          tm[j].r = a.r - b.r; tm[j].i = a.i - b.i;
          s0.r += tp[j].r; s0.i += tp[j].i;
        }
        if (last && sgn < 0) { s0.r *= sn; s0.i *= sn; }
        CHX(k, 0, i) = s0;
        for (int l = 1; l < ipph; ++l) {
          double ar = c0.r, ai = c0.i, br = 0.0, bi = 0.0;
          for (int j = 1; j < ipph; ++j) {
            int idlj = (l * j) % ip; /* root index, 1..ip-1 */
            double wr = wa[(idlj - 1) * ido], wi = wa[(idlj - 1) * ido + (ip - 1) * ido];
            ar += wr * tp[j].r; ai += wr * tp[j].i;
            br += wi * tm[j].r; bi += wi * tm[j].i;
          }
          /* X_l = A + sgn*i*B ; X_{ip-l} = A - sgn*i*B, with B = sum sin * (x_j - x_{ip-j}) */
          cpx ol, oc;
          if (sgn < 0) { ol.r = ar + bi; ol.i = ai - br; oc.r = ar - bi; oc.i = ai + br; }
          else         { ol.r = ar - bi; ol.i = ai + br; oc.r = ar + bi; oc.i = ai - br; }
          if (last) { if (sgn < 0) { ol.r *= sn; ol.i *= sn; oc.r *= sn; oc.i *= sn; } }
          else if (i > 0) {
            ol = twid(wa, ido, ip, i, l, sgn, ol);
            oc = twid(wa, ido, ip, i, ip - l, sgn, oc);
          }
          CHX(k, l, i) = ol; CHX(k, ip - l, i) = oc;
        }
      }
    free(tp);
  }
}

/* driver: fftpack.c:2041-2141 (c1fm1f_), :1931-2031 (c1fm1b_).  The result
 * always ends in c (stride inc); work is the ping-pong buffer (stride 1). */
static void c_fft(int n, int inc, cpx *c, const double *wsave, double *work, int sgn) {
  int nf = (int)wsave[2 * n];
  const double *fac = wsave + 2 * n + 1, *wa = wsave;
  cpx *w = (cpx *)work;
  int na = 0, l1 = 1, iw = 0;
  for (int k1 = 0; k1 < nf; ++k1) {
    int ip = (int)fac[k1], l2 = ip * l1, ido = n / l2;
    double sn = 1.0 / (double)(ip * l1);
    if (na == 0) c_pass(ido, l1, ip, c, inc, w, 1, wa + iw, sgn, sn);
    else         c_pass(ido, l1, ip, w, 1, c, inc, wa + iw, sgn, sn);
    na = 1 - na;
    l1 = l2;
    iw += (ip - 1) * (ido + ido);
  }
  if (na == 1)
    for (int i = 0; i < n; ++i) c[(size_t)i * inc] = w[i];
}

/* ------------------------------------------------------------------ */
/* real init: fftpack.c:13863-13975 (rffti1_) == :10360-10472 (mrfti1_) */
static void r_init(int n, double *wsave) {
  int fac[64];
  int nf = orc_rfactor(n, fac);
  double *wa = wsave, *f = wsave + n;
  f[0] = (double)n;
  f[1] = (double)nf;
  for (int k = 0; k < nf; ++k) f[2 + k] = (double)fac[k];
  double tpi = atan(1.) * 8.;
  double argh = tpi / (double)n;
  int is = 0, l1 = 1;
  for (int k1 = 0; k1 < nf - 1; ++k1) {
    int ip = fac[k1], ld = 0, l2 = l1 * ip, ido = n / l2;
    for (int j = 1; j < ip; ++j) {
      ld += l1;
      int i = is;
      double argld = (double)ld * argh, fi = 0.0;
      for (int ii = 3; ii <= ido; ii += 2) {
        i += 2;
        fi += 1.0;
        double arg = fi * argld;
        wa[i - 2] = cos(arg);
        wa[i - 1] = sin(arg);
      }
      is += ido;
    }
    l1 = l2;
  }
}

/* element g (0<=g<len) of the Hermitian spectrum held in half-complex order
 * [a0, re1, im1, re2, im2, ..., (re_{len/2})] */
static inline cpx hc_get(const double *s, int ss, int len, int g) {
  cpx v;
  if (g == 0) { v.r = s[0]; v.i = 0.0; return v; }
  if (2 * g < len) { v.r = s[(size_t)(2 * g - 1) * ss]; v.i = s[(size_t)(2 * g) * ss]; return v; }
  if (2 * g == len) { v.r = s[(size_t)(len - 1) * ss]; v.i = 0.0; return v; }
  g = len - g;
  v.r = s[(size_t)(2 * g - 1) * ss]; v.i = -s[(size_t)(2 * g) * ss];
  return v;
}
static inline void hc_put(double *s, int ss, int len, int f, cpx v) {
  if (f == 0) s[0] = v.r;
  else if (2 * f < len) { s[(size_t)(2 * f - 1) * ss] = v.r; s[(size_t)(2 * f) * ss] = v.i; }
  else if (2 * f == len) s[(size_t)(len - 1) * ss] = v.r;
}

/* roots of unity for one pass: rt[m] = exp(-2*pi*i*m/L) */
static cpx *roots(int L) {
  cpx *rt = (cpx *)malloc(sizeof(cpx) * L);
  double tpi = atan(1.0) * 8.0;
  for (int m = 0; m < L; ++m) {
    double a = tpi * (double)m / (double)L;
    rt[m].r = cos(a); rt[m].i = -sin(a);
  }
  return rt;
}

/* forward real pass: combine, for each k<l1, ip half-complex spectra of
 * length ido (of the subsequences offset k+l1*j) into one of length ido*ip.
 * cc(ido,l1,ip) -> ch(ido,ip,l1); contract of r1f{2,3,4,5,g}kf_ (fftpack.c:10886-12983) */
static void r_pass_f(int ido, int l1, int ip, const double *cc, int ics, double *ch, int ihs) {
  int L = ido * ip;
  cpx *rt = roots(L);
  for (int k = 0; k < l1; ++k)
    for (int f = 0; 2 * f <= L; ++f) {
      double xr = 0.0, xi = 0.0;
      int g = f % ido;
      for (int j = 0; j < ip; ++j) {
        cpx s = hc_get(cc + (size_t)ido * (k + l1 * j) * ics, ics, ido, g);
        cpx w = rt[(int)(((long long)j * f) % L)];
        xr += w.r * s.r - w.i * s.i;
        xi += w.r * s.i + w.i * s.r;
      }
      cpx v = {xr, xi};
      hc_put(ch + (size_t)L * k * ihs, ihs, L, f, v);
    }
  free(rt);
}

/* backward real pass: split, for each k<l1, one spectrum of length ido*ip
 * into the ip spectra of its decimated subsequences.  cc(ido,ip,l1) ->
 * ch(ido,l1,ip); contract of r1f{2,3,4,5,g}kb_ (fftpack.c:10791-12563) */
static void r_pass_b(int ido, int l1, int ip, const double *cc, int ics, double *ch, int ihs) {
  int L = ido * ip;
  cpx *rt = roots(L);
  for (int k = 0; k < l1; ++k)
    for (int j = 0; j < ip; ++j)
      for (int g = 0; 2 * g <= ido; ++g) {
        double sr = 0.0, si = 0.0;
        for (int r = 0; r < ip; ++r) {
          int f = g + ido * r;
          cpx y = hc_get(cc + (size_t)L * k * ics, ics, L, f);
          cpx w = rt[(int)(((long long)j * f) % L)]; /* conj -> e^{+} */
          sr += w.r * y.r + w.i * y.i;
          si += w.r * y.i - w.i * y.r;
        }
        cpx v = {sr, si};
        hc_put(ch + (size_t)ido * (k + l1 * j) * ihs, ihs, ido, g, v);
      }
  free(rt);
}

/* fftpack.c:13695-13853 (rfftf1_): factors walked in reverse, then the
 * sn / tsn / tsnm epilogue */
static void r_fftf(int n, int inc, double *r, const double *wsave, double *work) {
  const double *f = wsave + n;
  int nf = (int)f[1];
  int na = 0, l2 = n; /* na: 0 = data in r, 1 = data in work */
  for (int k1 = 0; k1 < nf; ++k1) {
    int ip = (int)f[2 + nf - 1 - k1], l1 = l2 / ip, ido = n / l2;
    if (na == 0) r_pass_f(ido, l1, ip, r, inc, work, 1);
    else         r_pass_f(ido, l1, ip, work, 1, r, inc);
    na = 1 - na;
    l2 = l1;
  }
  const double *src = na ? work : r;
  int ss = na ? 1 : inc;
  double sn = 1.0 / n, tsn = 2.0 / n, tsnm = -tsn;
  int nl = (n % 2) ? n - 1 : n - 2;
  r[0] = sn * src[0];
  for (int j = 1; j < nl; j += 2) {
    r[(size_t)j * inc] = tsn * src[(size_t)j * ss];
    r[(size_t)(j + 1) * inc] = tsnm * src[(size_t)(j + 1) * ss];
  }
  if (n % 2 == 0) r[(size_t)(n - 1) * inc] = sn * src[(size_t)(n - 1) * ss];
}

/* fftpack.c:13517-13685 (rfftb1_): half / halfm prescale, then the passes */
static void r_fftb(int n, int inc, double *r, const double *wsave, double *work) {
  const double *f = wsave + n;
  int nf = (int)f[1];
  int nl = (n % 2) ? n - 1 : n - 2;
  for (int j = 1; j < nl; j += 2) {
    r[(size_t)j * inc] = .5 * r[(size_t)j * inc];
    r[(size_t)(j + 1) * inc] = -.5 * r[(size_t)(j + 1) * inc];
  }
  int na = 0, l1 = 1;
  for (int k1 = 0; k1 < nf; ++k1) {
    int ip = (int)f[2 + k1], l2 = ip * l1, ido = n / l2;
    if (na == 0) r_pass_b(ido, l1, ip, r, inc, work, 1);
    else         r_pass_b(ido, l1, ip, work, 1, r, inc);
    na = 1 - na;
    l1 = l2;
  }
  if (na == 1)
    for (int i = 0; i < n; ++i) r[(size_t)i * inc] = work[i];
}

/* ------------------------------------------------------------------ */
/* public complex API: fftpack.c:2151-2276, 2499-2639 */
static int c_check1(int n, int inc, int lenc, int lensav, int lenwrk) {
  if (lenc < inc * (n - 1) + 1) return 1;
  if (lensav < 2 * n + il2(n) + 4) return 2;
  if (lenwrk < 2 * n) return 3;
  return 0;
}
int orc_cfft1i_(int *n, double *wsave, int *lensav, int *ier) {
  *ier = 0;
  if (*lensav < 2 * *n + il2(*n) + 4) { *ier = 2; return 0; } /* oracle stops; reference continues (:2263) */
  if (*n == 1) return 0;
  c_init(*n, wsave);
  return 0;
}
int orc_cfftmi_(int *n, double *wsave, int *lensav, int *ier) { return orc_cfft1i_(n, wsave, lensav, ier); }

static int c_fft1(int *n, int *inc, cpx *c, int *lenc, double *wsave, int *lensav, double *work, int *lenwrk, int *ier,
                  int sgn) {
  *ier = c_check1(*n, *inc, *lenc, *lensav, *lenwrk);
  if (*ier) return 0; /* oracle never touches data on error (documented deviation) */
  if (*n == 1) return 0;
  c_fft(*n, *inc, c, wsave, work, sgn);
  return 0;
}
int orc_cfft1f_(int *n, int *inc, cpx *c, int *lenc, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return c_fft1(n, inc, c, lenc, wsave, lensav, work, lenwrk, ier, -1);
}
int orc_cfft1b_(int *n, int *inc, cpx *c, int *lenc, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return c_fft1(n, inc, c, lenc, wsave, lensav, work, lenwrk, ier, +1);
}

static int c_fftm(int *lot, int *jump, int *n, int *inc, cpx *c, int *lenc, double *wsave, int *lensav, double *work,
                  int *lenwrk, int *ier, int sgn) {
  *ier = 0;
  if (*lenc < (*lot - 1) * *jump + *inc * (*n - 1) + 1) *ier = 1;
  else if (*lensav < 2 * *n + il2(*n) + 4) *ier = 2;
  else if (*lenwrk < 2 * *lot * *n) *ier = 3;
  else if (!orc_xercon(*inc, *jump, *n, *lot)) *ier = 4;
  if (*ier) return 0;
  if (*n == 1) return 0;
  /* cmfm1f_ (:5262) is bit-identical to looping c1fm1f_ over the lot */
  for (int m = 0; m < *lot; ++m) c_fft(*n, *inc, c + (size_t)m * *jump, wsave, work, sgn);
  return 0;
}
int orc_cfftmf_(int *lot, int *jump, int *n, int *inc, cpx *c, int *lenc, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return c_fftm(lot, jump, n, inc, c, lenc, wsave, lensav, work, lenwrk, ier, -1);
}
int orc_cfftmb_(int *lot, int *jump, int *n, int *inc, cpx *c, int *lenc, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return c_fftm(lot, jump, n, inc, c, lenc, wsave, lensav, work, lenwrk, ier, +1);
}

/* 2-D: fftpack.c:2285-2490 */
int orc_cfft2i_(int *l, int *m, double *wsave, int *lensav, int *ier) {
  int ier1, ls;
  *ier = 0;
  if (*lensav < 2 * *l + il2(*l) + 2 * *m + il2(*m) + 8) { *ier = 2; return 0; }
  ls = 2 * *l + il2(*l) + 4;
  orc_cfftmi_(l, wsave, &ls, &ier1);
  if (ier1) { *ier = 20; return 0; }
  ls = 2 * *m + il2(*m) + 4;
  orc_cfftmi_(m, wsave + 2 * *l + il2(*l) + 2, &ls, &ier1);
  if (ier1) *ier = 20;
  return 0;
}
static int c_fft2(int *ldim, int *l, int *m, cpx *c, double *wsave, int *lensav, double *work, int *lenwrk, int *ier,
                  int sgn) {
  int one = 1, ier1, lenc, ls, lw;
  *ier = 0;
  if (*l > *ldim) { *ier = 5; return 0; }
  if (*lensav < 2 * *l + il2(*l) + 2 * *m + il2(*m) + 8) { *ier = 2; return 0; }
  if (*lenwrk < 2 * *l * *m) { *ier = 3; return 0; }
  /* X lines: lot=l, jump=1, n=m, inc=ldim (:2408-2413) */
  lenc = *l - 1 + *ldim * (*m - 1) + 1; ls = 2 * *m + il2(*m) + 4; lw = 2 * *l * *m;
  c_fftm(l, &one, m, ldim, c, &lenc, wsave + 2 * *l + il2(*l) + 2, &ls, work, &lw, &ier1, sgn);
  if (ier1) { *ier = 20; return 0; }
  /* Y lines: lot=m, jump=ldim, n=l, inc=1 (:2421-2426) */
  lenc = (*m - 1) * *ldim + *l; ls = 2 * *l + il2(*l) + 4; lw = 2 * *m * *l;
  c_fftm(m, ldim, l, &one, c, &lenc, wsave, &ls, work, &lw, &ier1, sgn);
  if (ier1) *ier = 20;
  return 0;
}
int orc_cfft2f_(int *ldim, int *l, int *m, cpx *c, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return c_fft2(ldim, l, m, c, wsave, lensav, work, lenwrk, ier, -1);
}
int orc_cfft2b_(int *ldim, int *l, int *m, cpx *c, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return c_fft2(ldim, l, m, c, wsave, lensav, work, lenwrk, ier, +1);
}

/* 2-D real: fftpack.c:13113-13508 (rfft2b_, rfft2f_, rfft2i_) with the copies r2w_/w2r_ (:12949, :15175).
 * wsave = [rfft plan of l | cfft plan of m | rfft plan of m].  The array is real column-major r(ldim, m); each column
 * becomes a half-complex vector along i, then rows i=0 (and i=l-1 when l is even) get a real transform along j and
 * the (Re, Im) row pairs in between a complex one, run on a copy w(2*((l+1)/2), m) so that pairs are complex-aligned. */
static void rfft2_sizes(int l, int m, int *lw, int *mw, int *mm) {
  *lw = l + il2(l) + 4;
  *mw = 2 * m + il2(m) + 4;
  *mm = m + il2(m) + 4;
}
int orc_rfft2i_(int *l, int *m, double *wsave, int *lensav, int *ier) {
  int lw, mw, mm, ier1;
  *ier = 0;
  rfft2_sizes(*l, *m, &lw, &mw, &mm);
  if (*lensav < lw + mw + mm) { *ier = 2; return 0; }
  orc_rfftmi_(l, wsave, &lw, &ier1);
  if (ier1) { *ier = 20; return 0; }
  orc_cfftmi_(m, wsave + lw, &mw, &ier1);
  if (ier1) { *ier = 20; return 0; }
  orc_rfftmi_(m, wsave + lw + mw, &mm, &ier1);
  if (ier1) *ier = 20;
  return 0;
}
/* half-complex (FFTPACK 2/N cos, 2/N sin) <-> plain (Re, Im)/N along one strided line of length len */
static void hc_halve(double *x, int stride, int len) {
  int top = 2 * ((len + 1) / 2) - 1, k;
  for (k = 1; k < top; ++k) x[(long)k * stride] *= 0.5;
  for (k = 2; k < len; k += 2) x[(long)k * stride] = -x[(long)k * stride];
}
static void hc_double(double *x, int stride, int len) {
  int top = 2 * ((len + 1) / 2) - 1, k;
  for (k = 1; k < top; ++k) x[(long)k * stride] += x[(long)k * stride];
  for (k = 2; k < len; k += 2) x[(long)k * stride] = -x[(long)k * stride];
}
static int r_fft2(int *ldim, int *l, int *m, double *r, double *wsave, int *lensav, double *work, int *lenwrk, int *ier,
                  int fwd) {
  int lw, mw, mm, one = 1, ier1 = 0, lenr = *m * *ldim, ldh = (*l + 1) / 2, ldw = 2 * ldh, i, j;
  *ier = 0;
  rfft2_sizes(*l, *m, &lw, &mw, &mm);
  if (*lensav < lw + mw + mm) { *ier = 2; return 0; }
  if (*lenwrk < (*l + 1) * *m) { *ier = 3; return 0; }
  if (*ldim < *l) { *ier = 5; return 0; }
  if (fwd) {
    orc_rfftmf_(m, ldim, l, &one, r, &lenr, wsave, &lw, work, lenwrk, &ier1);
    if (ier1) { *ier = 20; return 0; }
    for (j = 0; j < *m; ++j) hc_halve(r + (long)j * *ldim, 1, *l);
    orc_rfftmf_(&one, &one, m, ldim, r, &lenr, wsave + lw + mw, &mm, work, lenwrk, &ier1);
    hc_halve(r, *ldim, *m);
  } else {
    hc_double(r, *ldim, *m);
    orc_rfftmb_(&one, &one, m, ldim, r, &lenr, wsave + lw + mw, &mm, work, lenwrk, &ier1);
  }
  if (ldh > 1) {
    int lot = ldh - 1, lenc = ldh * *m, lwk = *l * *m;
    for (j = 0; j < *m; ++j)
      for (i = 0; i < *l; ++i) work[i + (long)j * ldw] = r[i + (long)j * *ldim];
    if (fwd) orc_cfftmf_(&lot, &one, m, &ldh, (cpx *)(work + 1), &lenc, wsave + lw, &mw, r, &lwk, &ier1);
    else orc_cfftmb_(&lot, &one, m, &ldh, (cpx *)(work + 1), &lenc, wsave + lw, &mw, r, &lwk, &ier1);
    if (ier1) { *ier = 20; return 0; }
    for (j = 0; j < *m; ++j)
      for (i = 0; i < *l; ++i) r[i + (long)j * *ldim] = work[i + (long)j * ldw];
  }
  if (*l % 2 == 0) {
    double *ny = r + (*l - 1);
    if (fwd) {
      orc_rfftmf_(&one, &one, m, ldim, ny, &lenr, wsave + lw + mw, &mm, work, lenwrk, &ier1);
      hc_halve(ny, *ldim, *m);
    } else {
      hc_double(ny, *ldim, *m);
      orc_rfftmb_(&one, &one, m, ldim, ny, &lenr, wsave + lw + mw, &mm, work, lenwrk, &ier1);
    }
  }
  if (!fwd) {
    for (j = 0; j < *m; ++j) hc_double(r + (long)j * *ldim, 1, *l);
    orc_rfftmb_(m, ldim, l, &one, r, &lenr, wsave, &lw, work, lenwrk, &ier1);
  }
  if (ier1) *ier = 20;
  return 0;
}
int orc_rfft2f_(int *ldim, int *l, int *m, double *r, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return r_fft2(ldim, l, m, r, wsave, lensav, work, lenwrk, ier, 1);
}
int orc_rfft2b_(int *ldim, int *l, int *m, double *r, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return r_fft2(ldim, l, m, r, wsave, lensav, work, lenwrk, ier, 0);
}

/* ------------------------------------------------------------------ */
/* public real API: fftpack.c:12984-13112, 13984-14122 */
int orc_rfft1i_(int *n, double *wsave, int *lensav, int *ier) {
  *ier = 0;
  if (*lensav < *n + il2(*n) + 4) { *ier = 2; return 0; }
  if (*n == 1) return 0;
  r_init(*n, wsave);
  return 0;
}
int orc_rfftmi_(int *n, double *wsave, int *lensav, int *ier) { return orc_rfft1i_(n, wsave, lensav, ier); }

static int r_fft1(int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier, int fwd) {
  *ier = 0;
  if (*lenr < *inc * (*n - 1) + 1) *ier = 1;
  else if (*lensav < *n + il2(*n) + 4) *ier = 2;
  else if (*lenwrk < *n) *ier = 3;
  if (*ier) return 0;
  if (*n == 1) return 0;
  if (fwd) r_fftf(*n, *inc, r, wsave, work); else r_fftb(*n, *inc, r, wsave, work);
  return 0;
}
int orc_rfft1f_(int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return r_fft1(n, inc, r, lenr, wsave, lensav, work, lenwrk, ier, 1);
}
int orc_rfft1b_(int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return r_fft1(n, inc, r, lenr, wsave, lensav, work, lenwrk, ier, 0);
}
static int r_fftm(int *lot, int *jump, int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier, int fwd) {
  *ier = 0;
  if (*lenr < (*lot - 1) * *jump + *inc * (*n - 1) + 1) *ier = 1;
  else if (*lensav < *n + il2(*n) + 4) *ier = 2;
  else if (*lenwrk < *lot * *n) *ier = 3;
  else if (!orc_xercon(*inc, *jump, *n, *lot)) *ier = 4;
  if (*ier) return 0;
  if (*n == 1) return 0;
  for (int m = 0; m < *lot; ++m) {
    if (fwd) r_fftf(*n, *inc, r + (size_t)m * *jump, wsave, work);
    else     r_fftb(*n, *inc, r + (size_t)m * *jump, wsave, work);
  }
  return 0;
}
int orc_rfftmf_(int *lot, int *jump, int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return r_fftm(lot, jump, n, inc, r, lenr, wsave, lensav, work, lenwrk, ier, 1);
}
int orc_rfftmb_(int *lot, int *jump, int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) {
  return r_fftm(lot, jump, n, inc, r, lenr, wsave, lensav, work, lenwrk, ier, 0);
}

/* ------------------------------------------------------------------ */
/* DCT-I: fftpack.c:6107-6160 (cost1i_), :6294-6408 (costf1_), :6169-6284 (costb1_) */
static void cost_init(int n, double *wsave) {
  if (n <= 3) return;
  int nm1 = n - 1, ns2 = n / 2;
  double pi = atan(1.0) * 4.0, dt = pi / (double)nm1, fk = 0.0;
  for (int k = 2; k <= ns2; ++k) {
    int kc = n + 1 - k;
    fk += 1.0;
    wsave[k - 1] = sin(fk * dt) * 2.0;
    wsave[kc - 1] = cos(fk * dt) * 2.0;
  }
  r_init(nm1, wsave + n);
}
#define X(k) x[(size_t)((k) - 1) * inc] /* 1-based like the reference */
static void cost_core(int n, int inc, double *x, const double *wsave, double *work, int fwd) {
  int nm1 = n - 1, np1 = n + 1, ns2 = n / 2;
  if (n < 2) return;
  if (n == 2) {
    double x1h = X(1) + X(2);
    if (fwd) { X(2) = (X(1) - X(2)) * .5; X(1) = x1h * .5; }
    else     { X(2) = X(1) - X(2); X(1) = x1h; }
    return;
  }
  if (n == 3) {
    double x1p3 = X(1) + X(3);
    if (fwd) {
      double tx2 = X(2) + X(2);
      X(2) = (X(1) - X(3)) * .5; X(1) = (x1p3 + tx2) * .25; X(3) = (x1p3 - tx2) * .25;
    } else {
      double x2 = X(2);
      X(2) = X(1) - X(3); X(1) = x1p3 + x2; X(3) = x1p3 - x2;
    }
    return;
  }
  if (!fwd) { X(1) += X(1); X(n) += X(n); }
  double dsum = X(1) - X(n);
  X(1) += X(n);
  for (int k = 2; k <= ns2; ++k) {
    int kc = np1 - k;
    double t1 = X(k) + X(kc), t2 = X(k) - X(kc);
    dsum += wsave[kc - 1] * t2;
    t2 = wsave[k - 1] * t2;
    X(k) = t1 - t2; X(kc) = t1 + t2;
  }
  if (n % 2) X(ns2 + 1) += X(ns2 + 1);
  r_fftf(nm1, inc, x, wsave + n, work);
  if (fwd) {
    dsum = (1.0 / (double)nm1) * dsum;
    if (nm1 % 2 == 0) X(nm1) += X(nm1);
    for (int i = 3; i <= n; i += 2) {
      double xi = X(i) * .5;
      X(i) = X(i - 1) * .5; X(i - 1) = dsum; dsum += xi;
    }
    if (n % 2 == 0) X(n) = dsum;
    X(1) *= .5; X(n) *= .5;
  } else {
    double fnm1s2 = (double)nm1 / 2.0, fnm1s4 = (double)nm1 / 4.0;
    dsum *= .5;
    X(1) = fnm1s2 * X(1);
    if (nm1 % 2 == 0) X(nm1) += X(nm1);
    for (int i = 3; i <= n; i += 2) {
      double xi = fnm1s4 * X(i);
      X(i) = fnm1s4 * X(i - 1); X(i - 1) = dsum; dsum += xi;
    }
    if (n % 2 == 0) X(n) = dsum;
  }
}

/* DST-I: fftpack.c:14667-14715 (sint1i_), :14828-14923 (sintf1_), :14725-14820 (sintb1_) */
static void sint_init(int n, double *wsave) {
  if (n <= 1) return;
  int ns2 = n / 2, np1 = n + 1;
  double pi = atan(1.0) * 4.0, dt = pi / (double)np1;
  for (int k = 1; k <= ns2; ++k) wsave[k - 1] = sin(k * dt) * 2.0;
  r_init(np1, wsave + ns2);
}
static void sint_core(int n, int inc, double *x, const double *wsave, double *work, int fwd) {
  if (n < 2) return;
  if (n == 2) {
    double c = fwd ? 1.0 / sqrt(3.0) : sqrt(3.0) / 2.0;
    double xhold = c * (X(1) + X(2));
    X(2) = c * (X(1) - X(2)); X(1) = xhold;
    return;
  }
  int np1 = n + 1, ns2 = n / 2;
  double *xh = work, *rwork = work + np1; /* xh(1..np1) 1-based in the reference */
  for (int k = 1; k <= ns2; ++k) {
    int kc = np1 - k;
    double t1 = X(k) - X(kc), t2 = wsave[k - 1] * (X(k) + X(kc));
    xh[k] = t1 + t2; xh[kc] = t2 - t1;
  }
  if (n % 2) xh[ns2 + 1] = X(ns2 + 1) * 4.0;
  xh[0] = 0.0;
  r_fftf(np1, 1, xh, wsave + ns2, rwork);
  if (np1 % 2 == 0) xh[np1 - 1] += xh[np1 - 1];
  double sc = fwd ? .5 : (double)np1 / 4.0;
  X(1) = sc * xh[0];
  double dsum = X(1);
  for (int i = 3; i <= n; i += 2) {
    X(i - 1) = sc * xh[i - 1];
    dsum += sc * xh[i - 2];
    X(i) = dsum;
  }
  if (n % 2 == 0) X(n) = sc * xh[n];
}

/* quarter-wave cosine: fftpack.c:5523-5566 (cosq1i_), :5665-5741 (cosqf1_), :5576-5655 (cosqb1_) */
static void cosq_init(int n, double *wsave) {
  double pih = atan(1.0) * 2.0, dt = pih / (double)n, fk = 0.0;
  for (int k = 1; k <= n; ++k) { fk += 1.0; wsave[k - 1] = cos(fk * dt); }
  if (n > 1) r_init(n, wsave + n);
}
static void cosq_core(int n, int inc, double *x, const double *wsave, double *work, int fwd) {
  if (n < 2) return;
  if (n == 2) {
    double ssqrt2 = 1.0 / sqrt(2.0);
    if (fwd) { double tsqx = ssqrt2 * X(2); X(2) = X(1) * .5 - tsqx; X(1) = X(1) * .5 + tsqx; }
    else     { double x1 = X(1) + X(2); X(2) = ssqrt2 * (X(1) - X(2)); X(1) = x1; }
    return;
  }
  int ns2 = (n + 1) / 2, np2 = n + 2;
  double *w = work - 1; /* 1-based */
  if (fwd) {
    for (int k = 2; k <= ns2; ++k) { int kc = np2 - k; w[k] = X(k) + X(kc); w[kc] = X(k) - X(kc); }
    if (n % 2 == 0) w[ns2 + 1] = X(ns2 + 1) + X(ns2 + 1);
    for (int k = 2; k <= ns2; ++k) {
      int kc = np2 - k;
      X(k) = wsave[k - 2] * w[kc] + wsave[kc - 2] * w[k];
      X(kc) = wsave[k - 2] * w[k] - wsave[kc - 2] * w[kc];
    }
    if (n % 2 == 0) X(ns2 + 1) = wsave[ns2 - 1] * w[ns2 + 1];
    r_fftf(n, inc, x, wsave + n, work);
    for (int i = 3; i <= n; i += 2) {
      double xim1 = (X(i - 1) + X(i)) * .5;
      X(i) = (X(i - 1) - X(i)) * .5; X(i - 1) = xim1;
    }
  } else {
    for (int i = 3; i <= n; i += 2) {
      double xim1 = X(i - 1) + X(i);
      X(i) = (X(i - 1) - X(i)) * .5; X(i - 1) = xim1 * .5;
    }
    X(1) *= .5;
    if (n % 2 == 0) X(n) *= .5;
    r_fftb(n, inc, x, wsave + n, work);
    for (int k = 2; k <= ns2; ++k) {
      int kc = np2 - k;
      w[k] = wsave[k - 2] * X(kc) + wsave[kc - 2] * X(k);
      w[kc] = wsave[k - 2] * X(k) - wsave[kc - 2] * X(kc);
    }
    if (n % 2 == 0) X(ns2 + 1) = wsave[ns2 - 1] * (X(ns2 + 1) + X(ns2 + 1));
    for (int k = 2; k <= ns2; ++k) { int kc = np2 - k; X(k) = w[k] + w[kc]; X(kc) = w[k] - w[kc]; }
    X(1) += X(1);
  }
}

/* quarter-wave sine: fftpack.c:14123-14266: index reversal / sign flips around cosq */
static void sinq_core(int n, int inc, double *x, const double *wsave, double *work, int fwd) {
  if (n < 2) return;
  int ns2 = n / 2;
  if (fwd) {
    for (int k = 1; k <= ns2; ++k) { int kc = n - k; double t = X(k); X(k) = X(kc + 1); X(kc + 1) = t; }
    cosq_core(n, inc, x, wsave, work, 1);
    for (int k = 2; k <= n; k += 2) X(k) = -X(k);
  } else {
    for (int k = 2; k <= n; k += 2) X(k) = -X(k);
    cosq_core(n, inc, x, wsave, work, 0);
    for (int k = 1; k <= ns2; ++k) { int kc = n - k; double t = X(k); X(k) = X(kc + 1); X(kc + 1) = t; }
  }
}
#undef X

/* argument checks of the trig families (each returns early on error):
 * cost :6071-6084/:6522, sint :14640-14651/:15037, cosq :5480-5485, sinq :14247 */
enum { K_COST, K_SINT, K_COSQ, K_SINQ };
static int trig_lensav(int kind, int n) { return kind == K_SINT ? n / 2 + n + il2(n) + 4 : 2 * n + il2(n) + 4; }
static int trig_lenwrk1(int kind, int n) { return kind == K_COST ? n - 1 : kind == K_SINT ? 2 * n + 2 : n; }
static int trig_lenwrkm(int kind, int n, int lot) {
  return kind == K_COST ? lot * (n + 1) : kind == K_SINT ? lot * (2 * n + 4) : lot * n;
}
static int trig_i(int kind, int *n, double *wsave, int *lensav, int *ier) {
  *ier = 0;
  if (*lensav < trig_lensav(kind, *n)) { *ier = 2; return 0; }
  if (kind == K_COST) cost_init(*n, wsave);
  else if (kind == K_SINT) sint_init(*n, wsave);
  else cosq_init(*n, wsave);
  return 0;
}
static void trig_core(int kind, int n, int inc, double *x, const double *wsave, double *work, int fwd) {
  if (kind == K_COST) cost_core(n, inc, x, wsave, work, fwd);
  else if (kind == K_SINT) sint_core(n, inc, x, wsave, work, fwd);
  else if (kind == K_COSQ) cosq_core(n, inc, x, wsave, work, fwd);
  else sinq_core(n, inc, x, wsave, work, fwd);
}
static int trig_1(int kind, int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier, int fwd) {
  *ier = 0;
  if (*lenx < *inc * (*n - 1) + 1) *ier = 1;
  else if (*lensav < trig_lensav(kind, *n)) *ier = 2;
  else if (*lenwrk < trig_lenwrk1(kind, *n)) *ier = 3;
  /* sinq1b_ falls through its argument checks into cosq1b_, whose failure it reports as 20 (fftpack.c:14151-14179) */
  if (*ier && kind == K_SINQ && !fwd && *n > 1) *ier = 20;
  if (*ier) return 0;
  /* private scratch: the oracle's generic real passes need n+1 extra doubles */
  double *scratch = (double *)malloc(sizeof(double) * (size_t)(3 * *n + 8));
  trig_core(kind, *n, *inc, x, wsave, scratch, fwd);
  free(scratch);
  return 0;
}
static int trig_m(int kind, int *lot, int *jump, int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier, int fwd) {
  *ier = 0;
  if (*lenx < (*lot - 1) * *jump + *inc * (*n - 1) + 1) *ier = 1;
  else if (*lensav < trig_lensav(kind, *n)) *ier = 2;
  else if (*lenwrk < trig_lenwrkm(kind, *n, *lot)) *ier = 3;
  else if (!orc_xercon(*inc, *jump, *n, *lot)) *ier = 4;
  if (*ier && kind == K_SINQ && !fwd && *n > 1) *ier = 20;  /* sinqmb_ :14319-14409, same fall-through as sinq1b_ */
  if (*ier) return 0;
  double *scratch = (double *)malloc(sizeof(double) * (size_t)(3 * *n + 8));
  /* the batched drivers (mcstf1_ :7150, msntf1_ :10636, mcsqf1_ :6839) are
   * bit-identical to looping the single-sequence routine */
  for (int m = 0; m < *lot; ++m) trig_core(kind, *n, *inc, x + (size_t)m * *jump, wsave, scratch, fwd);
  free(scratch);
  return 0;
}
#define ORC_DEF_TRIG(name, K)                                                                                          \
  int orc_##name##1i_(int *n, double *wsave, int *lensav, int *ier) { return trig_i(K, n, wsave, lensav, ier); }      \
  int orc_##name##mi_(int *n, double *wsave, int *lensav, int *ier) { return trig_i(K, n, wsave, lensav, ier); }      \
  int orc_##name##1f_(int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) { \
    return trig_1(K, n, inc, x, lenx, wsave, lensav, work, lenwrk, ier, 1); }                                          \
  int orc_##name##1b_(int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) { \
    return trig_1(K, n, inc, x, lenx, wsave, lensav, work, lenwrk, ier, 0); }                                          \
  int orc_##name##mf_(int *lot, int *jump, int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) { \
    return trig_m(K, lot, jump, n, inc, x, lenx, wsave, lensav, work, lenwrk, ier, 1); }                               \
  int orc_##name##mb_(int *lot, int *jump, int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier) { \
    return trig_m(K, lot, jump, n, inc, x, lenx, wsave, lensav, work, lenwrk, ier, 0); }
ORC_DEF_TRIG(cost, K_COST)
ORC_DEF_TRIG(sint, K_SINT)
ORC_DEF_TRIG(cosq, K_COSQ)
ORC_DEF_TRIG(sinq, K_SINQ)

/* ------------------------------------------------------------------ */
/* O(N^2) definitions with FFTPACK scaling (test/naivepack.c:12-228).   */
/* ------------------------------------------------------------------ */
/* Application path: option value by convolution with the risk-neutral density in the frequency
 * domain, test/vargamma.c:42-106 (conv_bsvg_option) on top of cfftpack.c:446-492 (rfft_forward / rfft_inverse are a
 * shift of the half-complex vector, so the characteristic function multiplies the pairs (r[2i-1], r[2i]) directly; the
 * imaginary parts of the products at i = 0 and i = N/2 are dropped by the shift back).  Sizes: cfftextra.c:20-46. */
#include <complex.h>
int orc_next_fast_even_size(int n) {
  if (n <= 2) return 2;
  if (n & 1) ++n;
  for (;; n += 2) {
    int m = n;
    while (m % 5 == 0) m /= 5;
    while (m % 3 == 0) m /= 3;
    while (m % 2 == 0) m /= 2;
    if (m == 1) return n;
  }
}
static double _Complex option_charfn(double u, double sigma, double theta, double kappa, double t, double drift, int bs) {
  if (bs) return cexp(-0.5 * sigma * sigma * u * u * t + I * u * t * drift);
  double _Complex base = 1.0 + sigma * sigma * kappa * u * u / 2.0 - I * theta * kappa * u;
  return cpow(base, -t / kappa) * cexp(I * drift * u * t);
}
double orc_conv_bsvg_option(int n, double S, double K, double sigma, double theta, double kappa, double t, double r,
                            int is_call, int is_bs) {
  int N = orc_next_fast_even_size(n), N2 = N / 2, one = 1, ier = 0, i;
  int lensav = N + il2(N) + 4;
  double *V = (double *)calloc((size_t)N, sizeof(double)), *ws = (double *)calloc((size_t)lensav, sizeof(double));
  double *wk = (double *)calloc((size_t)N, sizeof(double));
  double L = 2 * 10 * sigma * sqrt(t), ds = L / N, du = 2 * M_PI / (ds * N), lS = log(S), value;
  double drift = is_bs ? r - 0.5 * sigma * sigma : r + (1.0 / kappa) * log(1.0 - sigma * sigma * kappa / 2.0 - theta * kappa);
  for (i = 0; i < N; ++i) {
    double e = exp(lS + (N2 - i) * ds);
    V[i] = is_call ? (e - K > 0.0 ? e - K : 0.0) : (K - e > 0.0 ? K - e : 0.0);
  }
  orc_rfft1i_(&N, ws, &lensav, &ier);
  orc_rfft1f_(&N, &one, V, &N, ws, &lensav, wk, &N, &ier);
  for (i = 0; i <= N2; ++i) {
    double _Complex phi = option_charfn(i * du, sigma, theta, kappa, t, drift, is_bs);
    if (i == 0) V[0] = creal(V[0] * phi);
    else if (i == N2) V[N - 1] = creal(V[N - 1] * phi);
    else {
      double _Complex v = (V[2 * i - 1] + I * V[2 * i]) * phi;
      V[2 * i - 1] = creal(v);
      V[2 * i] = cimag(v);
    }
  }
  orc_rfft1b_(&N, &one, V, &N, ws, &lensav, wk, &N, &ier);
  value = V[N2] * exp(-r * t);
  free(V); free(ws); free(wk);
  return value;
}

/* ------------------------------------------------------------------ */
/* L2 object wrapper, cfftpack/cfftpack.c: one handle per (algorithm, n), the reference's scaling
 * conventions with and without fft_ortho.  algo numbers follow cfftintern.h's enum (CFFT 1, RFFT 2, CFFT2 3, DCT1 4,
 * DCT 5, DST1 7, DST 8).  Quirks kept on purpose: the length passed down is n (n*inc only for DCT, :186/:203), so a
 * stride > 1 yields ier = 1 everywhere else; the ortho factors of fft_forward/inverse are 1/sqrt(n) and sqrt(n) ON
 * TOP of FFTPACK's 1/n (:69-76, :90-97); ortho loops ignore the stride except in dct_forward/inverse. */
struct orc_l2 {
  int algo, n, m, ortho, inc, lensav, lenwork;
  double *save, *work;
};
void orc_l2_free(orc_l2_t *f) {
  if (!f) return;
  free(f->save);
  free(f->work);
  free(f);
}
orc_l2_t *orc_l2_create(int algo, int n, int m) {
  orc_l2_t *f;
  int ier = 0;
  if (n <= 0 || (algo == 3 && m <= 0) || (algo == 4 && n <= 1)) return NULL;
  f = (orc_l2_t *)calloc(1, sizeof(*f));
  f->algo = algo; f->n = n; f->m = m; f->inc = 1;
  switch (algo) {
    case 1: f->lensav = 2 * n + il2(n) + 4; f->lenwork = 2 * n; break;                          /* :8-29 */
    case 2: f->lensav = n + il2(n) + 4; f->lenwork = n; break;                                  /* :427-444 */
    case 3: f->lensav = 2 * n + il2(n) + 2 * m + il2(m) + 8; f->lenwork = 2 * n * m; break;     /* :102-127 */
    case 7: f->lensav = n / 2 + n + il2(n) + 4; f->lenwork = 2 * n + 2; break;                  /* :374-392 */
    default: f->lensav = 2 * n + il2(n) + 4; f->lenwork = n; break;                             /* :155, :223, :302 */
  }
  f->save = (double *)calloc((size_t)f->lensav + 8, sizeof(double));
  f->work = (double *)calloc((size_t)f->lenwork + 8, sizeof(double));
  switch (algo) {
    case 1: orc_cfft1i_(&n, f->save, &f->lensav, &ier); break;
    case 2: orc_rfft1i_(&n, f->save, &f->lensav, &ier); break;
    case 3: orc_cfft2i_(&n, &m, f->save, &f->lensav, &ier); break;
    case 4: orc_cost1i_(&n, f->save, &f->lensav, &ier); break;
    case 5: orc_cosq1i_(&n, f->save, &f->lensav, &ier); break;
    case 7: orc_sint1i_(&n, f->save, &f->lensav, &ier); break;
    case 8: orc_sinq1i_(&n, f->save, &f->lensav, &ier); break;
    default: ier = 1;
  }
  if (ier) { orc_l2_free(f); return NULL; }
  return f;
}
void orc_l2_ortho(orc_l2_t *f, int ortho) { if (f) f->ortho = ortho; }
void orc_l2_stride(orc_l2_t *f, int stride) { if (f) f->inc = stride > 0 ? stride : 1; }
static void scale_first_rest(double *x, int n, int inc, double first, double rest) {
  int i;
  x[0] *= first;
  for (i = 1; i < n; ++i) x[(long)i * inc] *= rest;
}
/* cfftpack.c:245-275: both directions of the orthonormal DCT-I go through the unscaled backward transform */
static int l2_dct1_ortho(orc_l2_t *f, double *x) {
  const double r2 = 1.0 / sqrt(2.0), m = sqrt(2.0 / (f->n - 1.0));
  double ev = (x[0] + x[f->n - 1]) * (-1.0 + r2), od = (x[0] - x[f->n - 1]) * (-1.0 + r2);
  int ier = 0, i;
  orc_cost1b_(&f->n, &f->inc, x, &f->n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
  if (ier) return ier;
  for (i = 0; i < f->n; ++i) x[i] = (x[i] + (i % 2 == 0 ? ev : od)) * m;
  x[0] *= r2;
  x[f->n - 1] *= r2;
  return 0;
}
static int l2_run(orc_l2_t *f, void *data, int fwd) {
  int ier = 0, n, len, i;
  double *x = (double *)data;
  if (!f || !data) return -1;
  n = f->n;
  switch (f->algo) {
    case 1: {  /* fft_forward :57-78, fft_inverse :81-99 */
      cpx *c = (cpx *)data;
      double mul = fwd ? 1.0 / sqrt(n) : sqrt(n);
      if (fwd) orc_cfft1f_(&n, &f->inc, c, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
      else orc_cfft1b_(&n, &f->inc, c, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
      if (ier) return ier;
      if (f->ortho) for (i = 0; i < n; ++i) { c[i].r *= mul; c[i].i *= mul; }
      return 0;
    }
    case 3:  /* fft2_forward :131-141, fft2_inverse :143-152: ldim = l */
      if (fwd) orc_cfft2f_(&n, &n, &f->m, (cpx *)data, f->save, &f->lensav, f->work, &f->lenwork, &ier);
      else orc_cfft2b_(&n, &n, &f->m, (cpx *)data, f->save, &f->lensav, f->work, &f->lenwork, &ier);
      return ier;
    case 5:  /* dct_forward :176-195 (DCT-III), dct_inverse :198-218 (DCT-II) */
      len = n * f->inc;
      if (fwd) {
        if (f->ortho) scale_first_rest(x, n, f->inc, sqrt(n), sqrt(0.5 * n));
        orc_cosq1f_(&n, &f->inc, x, &len, f->save, &f->lensav, f->work, &f->lenwork, &ier);
      } else {
        orc_cosq1b_(&n, &f->inc, x, &len, f->save, &f->lensav, f->work, &f->lenwork, &ier);
        if (f->ortho) scale_first_rest(x, n, f->inc, 1.0 / sqrt(n), sqrt(2.0 / n));
      }
      return ier;
    case 4:  /* dct1_forward :277-292, dct1_inverse :294-308 */
      if (f->ortho) return l2_dct1_ortho(f, x);
      if (fwd) orc_cost1f_(&n, &f->inc, x, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
      else orc_cost1b_(&n, &f->inc, x, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
      return ier;
    case 8:  /* dst_forward :330-352, dst_inverse :354-371 */
      if (fwd) {
        if (f->ortho) scale_first_rest(x, n, 1, sqrt(1.0 / n), sqrt(0.5 / n));
        orc_sinq1f_(&n, &f->inc, x, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
        if (f->ortho) for (i = 0; i < n; ++i) x[i] *= (double)n;
        return ier;
      }
      orc_sinq1b_(&n, &f->inc, x, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
      if (ier) return ier;
      if (f->ortho) scale_first_rest(x, n, 1, sqrt(1.0 / n), sqrt(2.0 / n));
      return 0;
    case 7:  /* dst1_forward :394-408 (ortho: same as inverse), dst1_inverse :410-426 */
      if (fwd && !f->ortho) {
        orc_sint1f_(&n, &f->inc, x, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
        return ier;
      }
      orc_sint1b_(&n, &f->inc, x, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
      if (ier) return ier;
      if (f->ortho) for (i = 0; i < n; ++i) x[i] *= sqrt(2.0 / (n + 1));
      return 0;
    default: return -2;
  }
}
int orc_l2_forward(orc_l2_t *f, void *data) { return l2_run(f, data, 1); }
int orc_l2_inverse(orc_l2_t *f, void *data) { return l2_run(f, data, 0); }
/* rfft_forward :446-469: half-complex vector shifted by one slot into complex[n/2+1] with zero imaginary ends */
int orc_l2_rfft_forward(orc_l2_t *f, const double *in, void *out) {
  double *d = (double *)out;
  int ier = 0, one = 1, i, n;
  if (!f || !in || !out) return -1;
  if (f->algo != 2) return -2;
  n = f->n;
  if (in != d) memcpy(d, in, (size_t)n * sizeof(double));
  orc_rfft1f_(&n, &one, d, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
  for (i = n; i > 1; --i) d[i] = d[i - 1];
  d[1] = 0;
  if (n % 2 == 0) d[n + 1] = 0;
  return ier;
}
/* rfft_inverse :471-490 */
int orc_l2_rfft_inverse(orc_l2_t *f, const void *in, double *out) {
  const double *d = (const double *)in;
  int ier = 0, one = 1, i, n;
  if (!f || !in || !out) return -1;
  if (f->algo != 2) return -2;
  n = f->n;
  out[0] = d[0];
  for (i = 1; i < n; ++i) out[i] = d[i + 1];
  orc_rfft1b_(&n, &one, out, &n, f->save, &f->lensav, f->work, &f->lenwork, &ier);
  return ier;
}

static const double PI_ = 3.14159265358979323846;
void orc_naive_cfft(int n, const cpx *x, cpx *y, int forward) {
  for (int k = 0; k < n; ++k) {
    long double sr = 0, si = 0;
    for (int t = 0; t < n; ++t) {
      long long m = ((long long)k * t) % n;
      long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / n;
      long double c = cosl(a), s = forward ? -sinl(a) : sinl(a);
      sr += c * x[t].r - s * x[t].i;
      si += c * x[t].i + s * x[t].r;
    }
    if (forward) { sr /= n; si /= n; }
    y[k].r = (double)sr; y[k].i = (double)si;
  }
}
void orc_naive_rfftf(int n, const double *x, double *y) {
  for (int f = 0; 2 * f <= n; ++f) {
    long double a = 0, b = 0;
    for (int t = 0; t < n; ++t) {
      long long m = ((long long)f * t) % n;
      long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / n;
      a += x[t] * cosl(ang); b += x[t] * sinl(ang);
    }
    if (f == 0) y[0] = (double)(a / n);
    else if (2 * f < n) { y[2 * f - 1] = (double)(2 * a / n); y[2 * f] = (double)(2 * b / n); }
    else y[n - 1] = (double)(a / n);
  }
}
void orc_naive_rfftb(int n, const double *r, double *y) {
  for (int t = 0; t < n; ++t) {
    long double s = r[0];
    for (int f = 1; 2 * f < n; ++f) {
      long long m = ((long long)f * t) % n;
      long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)m / n;
      s += r[2 * f - 1] * cosl(ang) + r[2 * f] * sinl(ang);
    }
    if (n % 2 == 0) s += (t % 2 ? -1.0L : 1.0L) * r[n - 1];
    y[t] = (double)s;
  }
}
/* naive_dct1 mode +1 / -1 (naivepack.c:12-40) */
void orc_naive_cost(int N, const double *x, double *y, int forward) {
  double M = N - 1, m0 = forward ? 0.5 : 1.0, m = forward ? 2.0 / M : 1.0;
  for (int k = 0; k < N; ++k) {
    double s = 0;
    for (int n = 1; n < N - 1; ++n) s += x[n] * cos(n * (double)k * PI_ / M);
    s += m0 * x[0];
    s += m0 * x[N - 1] * (k % 2 == 0 ? 1 : -1);
    y[k] = s * m;
  }
  y[0] *= m0; y[N - 1] *= m0;
}
/* naive_dst1 (naivepack.c:138-160): y_k = m * sum x_n sin((n+1)(k+1)pi/(N+1)) */
void orc_naive_sint(int N, const double *x, double *y, int forward) {
  double m = forward ? 2.0 / (N + 1) : 1.0;
  for (int k = 0; k < N; ++k) {
    double s = 0;
    for (int n = 0; n < N; ++n) s += x[n] * sin((n + 1) * (double)(k + 1) * PI_ / (N + 1));
    y[k] = s * m;
  }
}
/* forward = naive_dct3 non-ortho (naivepack.c:61-80); backward = naive_dct2 (:43-59) */
void orc_naive_cosq(int N, const double *x, double *y, int forward) {
  if (forward) {
    for (int k = 0; k < N; ++k) {
      double s = 0.5 * x[0];
      for (int n = 1; n < N; ++n) s += x[n] * cos(n * (k + 0.5) * PI_ / N);
      y[k] = s * 2.0 / N;
    }
  } else {
    for (int k = 0; k < N; ++k) {
      double s = 0;
      for (int n = 0; n < N; ++n) s += x[n] * cos((n + 0.5) * k * PI_ / N);
      y[k] = s;
    }
  }
}
/* forward = naive_dst3 non-ortho, backward = naive_dst2 (naivepack.c:162-228) */
void orc_naive_sinq(int N, const double *x, double *y, int forward) {
  if (forward) {
    for (int k = 0; k < N; ++k) {
      double s = 0.5 * x[N - 1] * (k % 2 == 0 ? 1 : -1);
      for (int n = 0; n < N - 1; ++n) s += x[n] * sin((n + 1) * (k + 0.5) * PI_ / N);
      y[k] = s * 2.0 / N;
    }
  } else {
    for (int k = 0; k < N; ++k) {
      double s = 0;
      for (int n = 0; n < N; ++n) s += x[n] * sin((n + 0.5) * (k + 1) * PI_ / N);
      y[k] = s;
    }
  }
}

/* ------------------------------------------------------------------ */
/* lot-parallel CPU baseline driver     */
typedef struct {
  orc_fft1_fn fn; int is_complex, n, lot0, lot1, lensav, reps; void *data; double *wsave;
} lp_arg;
static void *lp_thread(void *p) {
  lp_arg *a = (lp_arg *)p;
  int n = a->n, inc = 1, lenx = n, lenwrk = 2 * n + 8, ier = 0;
  double *work = (double *)malloc(sizeof(double) * (size_t)lenwrk);
  size_t esz = a->is_complex ? 16 : 8;
  for (int rep = 0; rep < a->reps; ++rep)
    for (int m = a->lot0; m < a->lot1; ++m)
      a->fn(&n, &inc, (char *)a->data + (size_t)m * n * esz, &lenx, a->wsave, &a->lensav, work, &lenwrk, &ier);
  free(work);
  return NULL;
}
double orc_lot_parallel(orc_fft1_fn fn, int is_complex, int lot, int n, void *data, double *wsave, int lensav,
                        int nthreads, int reps) {
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
  lp_arg *args = (lp_arg *)malloc(sizeof(lp_arg) * nthreads);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < nthreads; ++t) {
    lp_arg a = {fn, is_complex, n, (int)((long long)lot * t / nthreads), (int)((long long)lot * (t + 1) / nthreads), lensav, reps, data, wsave};
    args[t] = a;
    pthread_create(&th[t], NULL, lp_thread, &args[t]);
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  free(th); free(args);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
