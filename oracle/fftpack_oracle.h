/*
 * oracle/fftpack_oracle.h -- CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the FFTPACK 5.1 algorithms on the hot path of
 * zywina/cfftpack (cfftpack/fftpack.c).  It exists so that the CUDA library
 * can be checked against something that runs on a machine where
 * /root/reference does not exist.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * library (cfftpack_b200/csrc) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_pinned.py compares every routine
 * below with the unmodified reference compiled into oracle/_ref/ (wsave
 * tables bitwise, transforms to <= 4 ulp-ish relative L2) and with the
 * reference's own O(N^2) definitions in test/naivepack.c, on the input
 * vectors the reference's tests use (test/testall.c:79, test/test1.c:21).
 *
 * All entry points carry the reference's Fortran-style signatures with an
 * `orc_` prefix: scalars by pointer, status through *ier, return value 0.
 */
#ifndef FFTPACK_ORACLE_H
#define FFTPACK_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { double r, i; } orc_complex_t;

/* factorisation helpers (fftpack.c:6613 factor_, :13892 inline copy in rffti1_) */
int orc_factor(int n, int *fac);
int orc_rfactor(int n, int *fac);
int orc_xercon(int inc, int jump, int n, int lot);

/* complex 1-D / multi / 2-D  (fftpack.c:2151-2639) */
int orc_cfft1i_(int *n, double *wsave, int *lensav, int *ier);
int orc_cfft1f_(int *n, int *inc, orc_complex_t *c, int *lenc, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
int orc_cfft1b_(int *n, int *inc, orc_complex_t *c, int *lenc, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
int orc_cfftmi_(int *n, double *wsave, int *lensav, int *ier);
int orc_cfftmf_(int *lot, int *jump, int *n, int *inc, orc_complex_t *c, int *lenc, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
int orc_cfftmb_(int *lot, int *jump, int *n, int *inc, orc_complex_t *c, int *lenc, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
int orc_cfft2i_(int *l, int *m, double *wsave, int *lensav, int *ier);
int orc_cfft2f_(int *ldim, int *l, int *m, orc_complex_t *c, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
int orc_cfft2b_(int *ldim, int *l, int *m, orc_complex_t *c, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
/* fftpack.c:13454, :13282, :13113 */
int orc_rfft2i_(int *l, int *m, double *wsave, int *lensav, int *ier);
int orc_rfft2f_(int *ldim, int *l, int *m, double *r, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
int orc_rfft2b_(int *ldim, int *l, int *m, double *r, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);

/* real 1-D / multi (fftpack.c:12984-13112, 13984-14122) */
int orc_rfft1i_(int *n, double *wsave, int *lensav, int *ier);
int orc_rfft1f_(int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
int orc_rfft1b_(int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
int orc_rfftmi_(int *n, double *wsave, int *lensav, int *ier);
int orc_rfftmf_(int *lot, int *jump, int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
int orc_rfftmb_(int *lot, int *jump, int *n, int *inc, double *r, int *lenr, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);

/* DCT-I (cost), DST-I (sint), quarter-wave cosine (cosq), quarter-wave sine (sinq) */
#define ORC_DECL_TRIG(name)                                                                                          \
  int orc_##name##1i_(int *n, double *wsave, int *lensav, int *ier);                                                  \
  int orc_##name##1f_(int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier); \
  int orc_##name##1b_(int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier); \
  int orc_##name##mi_(int *n, double *wsave, int *lensav, int *ier);                                                  \
  int orc_##name##mf_(int *lot, int *jump, int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier); \
  int orc_##name##mb_(int *lot, int *jump, int *n, int *inc, double *x, int *lenx, double *wsave, int *lensav, double *work, int *lenwrk, int *ier);
ORC_DECL_TRIG(cost)
ORC_DECL_TRIG(sint)
ORC_DECL_TRIG(cosq)
ORC_DECL_TRIG(sinq)

/* O(N^2) definitions, restating test/naivepack.c:12-228 with FFTPACK scaling */
void orc_naive_cfft(int n, const orc_complex_t *x, orc_complex_t *y, int forward);
void orc_naive_rfftf(int n, const double *x, double *y);
void orc_naive_rfftb(int n, const double *x, double *y);
void orc_naive_cost(int n, const double *x, double *y, int forward);
void orc_naive_sint(int n, const double *x, double *y, int forward);
void orc_naive_cosq(int n, const double *x, double *y, int forward);
void orc_naive_sinq(int n, const double *x, double *y, int forward);

/* lot-parallel CPU baseline helper for bench.py (pthreads over lot chunks,
 * each thread looping the single-transform routine with a private work
 * array).  kind: 0 cfft1f, 1 rfft1f.
 * fn is the address of the single-transform routine to loop (oracle or
 * oracle/_ref symbol), so the same driver times either arm. */
typedef int (*orc_fft1_fn)(int *, int *, void *, int *, double *, int *, double *, int *, int *);
double orc_lot_parallel(orc_fft1_fn fn, int is_complex, int lot, int n, void *data, double *wsave, int lensav,
                        int nthreads, int reps);

#ifdef __cplusplus
}
#endif
/* test/vargamma.c:42-106, cfftextra.c:42-46 */
int orc_next_fast_even_size(int n);
double orc_conv_bsvg_option(int n, double S, double K, double sigma, double theta, double kappa, double t, double r,
                            int is_call, int is_bs);
/* cfftpack/cfftpack.c (L2 object wrapper); algo numbers as in cfftintern.h */
typedef struct orc_l2 orc_l2_t;
orc_l2_t *orc_l2_create(int algo, int n, int m);
void orc_l2_free(orc_l2_t *f);
void orc_l2_ortho(orc_l2_t *f, int ortho);
void orc_l2_stride(orc_l2_t *f, int stride);
int orc_l2_forward(orc_l2_t *f, void *data);
int orc_l2_inverse(orc_l2_t *f, void *data);
int orc_l2_rfft_forward(orc_l2_t *f, const double *in, void *out);
int orc_l2_rfft_inverse(orc_l2_t *f, const void *in, double *out);
#endif
