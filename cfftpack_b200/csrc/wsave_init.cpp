/*
 * wsave_init.cpp -- host initialisers for the caller-owned wsave arrays.
 *
 * Drop-in contract: after cfft1i_/cfftmi_/rfft1i_/cost1i_/... the caller's wsave must hold
 * what the reference would have written, value for value, so that code which mixes this library with the
 * reference's wrapper layer (cfftpack/cfftpack.c) or inspects the factor list keeps working.  The layouts
 * and the order of the floating-point operations therefore follow cfftpack/fftpack.c exactly:
 *   complex: mcfti1_ :6666-6697, factor_ :6613-6657, tables_ :15124-15166
 *   real   : rffti1_ :13863-13975 (== mrfti1_ :10360-10472)
 *   cost   : cost1i_ :6107-6160      sint: sint1i_ :14667-14715      cosq: cosq1i_ :5523-5566
 * The GPU transforms do not read these tables (they use the device plans of plan.cpp).
 */
#include <math.h>

#include "internal.h"

namespace cfb {

int log2_floor_ref(int n) { return (int)(log((double)n) / log(2.0)); }

bool strides_consistent(int inc, int jump, int n, int lot) {
  // gcd by Euclid, then the least common multiple; the sequences overlap iff the lcm is reachable
  // along both axes (xercon_, fftpack.c:15210-15258)
  int a = inc, b = jump;
  while (b != 0) {
    int t = a % b;
    a = b;
    b = t;
  }
  int lcm = inc * jump / a;
  return !(lcm <= (n - 1) * inc && lcm <= (lot - 1) * jump);
}

/* trial divisors 4, 2, 3, 5, then 7, 9, 11, ... ; front2: a factor 2 found after the first factor moves to the front */
static int factor_list(int n, int *fac, bool front2) {
  int left = n, count = 0, trial = 0;
  for (int step = 0; left != 1; ++step) {
    trial = step == 0 ? 4 : step == 1 ? 2 : step == 2 ? 3 : step == 3 ? 5 : trial + 2;
    while (left % trial == 0) {
      fac[count++] = trial;
      left /= trial;
      if (front2 && trial == 2 && count != 1) {
        for (int i = count - 1; i > 0; --i) fac[i] = fac[i - 1];
        fac[0] = 2;
      }
    }
  }
  return count;
}

void wsave_init_complex(int n, double *wsave) {
  int fac[64];
  const int nf = factor_list(n, fac, false);
  wsave[2 * n] = (double)nf;
  for (int k = 0; k < nf; ++k) wsave[2 * n + 1 + k] = (double)fac[k];
  const double tpi = atan(1.0) * 8.0;
  double *wa = wsave;
  int l1 = 1;
  for (int k = 0; k < nf; ++k) {
    const int ip = fac[k], l2 = l1 * ip, ido = n / l2;
    const double argz = tpi / (double)ip;
    const double arg1 = tpi / (double)(ido * ip);
    double *wc = wa, *ws = wa + (ip - 1) * ido;  // cosine block, sine block of this stage
    for (int j = 1; j < ip; ++j) {
      const double arg2 = (double)j * arg1;
      for (int i = 0; i < ido; ++i) {
        const double arg3 = (double)i * arg2;
        wc[(j - 1) * ido + i] = cos(arg3);
        ws[(j - 1) * ido + i] = sin(arg3);
      }
      if (ip > 5) {  // generic radix: row i = 0 carries the roots of unity of order ip
        const double arg4 = (double)j * argz;
        wc[(j - 1) * ido] = cos(arg4);
        ws[(j - 1) * ido] = sin(arg4);
      }
    }
    wa += (ip - 1) * 2 * ido;
    l1 = l2;
  }
}

void wsave_init_real(int n, double *wsave) {
  int fac[64];
  const int nf = factor_list(n, fac, true);
  double *tail = wsave + n;
  tail[0] = (double)n;
  tail[1] = (double)nf;
  for (int k = 0; k < nf; ++k) tail[2 + k] = (double)fac[k];
  const double tpi = atan(1.) * 8.;
  const double argh = tpi / (double)n;
  int is = 0, l1 = 1;
  for (int k = 0; k + 1 < nf; ++k) {
    const int ip = fac[k], l2 = l1 * ip, ido = n / l2;
    int ld = 0;
    for (int j = 1; j < ip; ++j) {
      ld += l1;
      const double argld = (double)ld * argh;
      double fi = 0.0;
      int i = is;
      for (int ii = 3; ii <= ido; ii += 2) {
        fi += 1.0;
        const double arg = fi * argld;
        wsave[i] = cos(arg);
        wsave[i + 1] = sin(arg);
        i += 2;
      }
      is += ido;
    }
    l1 = l2;
  }
}

void wsave_init_cost(int n, double *wsave) {
  if (n <= 3) return;
  const int nm1 = n - 1, ns2 = n / 2;
  const double pi = atan(1.0) * 4.0, dt = pi / (double)nm1;
  double fk = 0.0;
  for (int k = 2; k <= ns2; ++k) {
    fk += 1.0;
    wsave[k - 1] = sin(fk * dt) * 2.0;
    wsave[n - k] = cos(fk * dt) * 2.0;
  }
  wsave_init_real(nm1, wsave + n);
}

void wsave_init_sint(int n, double *wsave) {
  if (n <= 1) return;
  const int ns2 = n / 2, np1 = n + 1;
  const double pi = atan(1.0) * 4.0, dt = pi / (double)np1;
  for (int k = 1; k <= ns2; ++k) wsave[k - 1] = sin(k * dt) * 2.0;
  wsave_init_real(np1, wsave + ns2);
}

void wsave_init_cosq(int n, double *wsave) {
  const double pih = atan(1.0) * 2.0, dt = pih / (double)n;
  double fk = 0.0;
  for (int k = 0; k < n; ++k) {
    fk += 1.0;
    wsave[k] = cos(fk * dt);
  }
  if (n > 1) wsave_init_real(n, wsave + n);
}

}  // namespace cfb
