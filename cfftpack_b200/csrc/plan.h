/*
 * plan.h -- device-resident factor plans.
 *
 * The reference keeps its plan in the caller's wsave array: factor list + per-stage twiddles written by
 * factor_/tables_/mcfti1_/rffti1_ (cfftpack/fftpack.c:6613, :15124, :6666, :13863) and the cost/sint/cosq
 * tables of cost1i_/sint1i_/cosq1i_ (:6107, :14667, :5523).  Here the *i routines still fill wsave exactly
 * like the reference (wsave_init.cpp), but the transforms run from plans that live in HBM, are built once
 * per (device, length) and are shared by every thread of the process.
 */
#ifndef CFB_PLAN_H
#define CFB_PLAN_H
#include "engine_types.h"

namespace cfb {

/* complex core plan for length M */
struct CorePlan {
  int M = 0, nf = 0;
  PassDesc pass[CFB_MAXPASS];
  cpx *d_tw = nullptr;  // all pass twiddles + generic-radix root tables
  size_t tw_count = 0;
  int max_radix = 1;
};

/* tables of the trigonometric kinds, by (kind, n) */
struct TrigPlan {
  int kind = 0, n = 0, M = 0;
  double *d_trig = nullptr;
};

/* four-step twiddles W_n^x, x < n, as two short tables: d_w[x & (B-1)] * d_w[B + (x >> shift)], B = 2^shift ~ sqrt(n) */
struct RootPlan {
  int n = 0, shift = 0;
  cpx *d_w = nullptr;
};

/* Bluestein plan of length n: chirp c_j = exp(-pi i j^2/n) and the length-L transform of its wrapped conjugate */
struct ChirpPlan {
  int n = 0, L = 0;
  cpx *d_chirp = nullptr, *d_bhat = nullptr;
};
const ChirpPlan *get_chirp_plan(int n);

/* Stockham radix schedule used by the engine: 16/8/4/2 for the power of two, then 3, 5, odd primes */
int engine_factor(int M, int *radix);

/* all three return nullptr on a CUDA failure (cfb_last_error() says why) */
const CorePlan *get_core_plan(int M);
const TrigPlan *get_trig_plan(int kind, int n);
const RootPlan *get_root_plan(int n);
void release_plans();

/* exp(-2 pi i num / den) to correctly rounded-ish double precision (evaluated in long double) */
void unit_root(long long num, long long den, double *re, double *im);

}  // namespace cfb
#endif
