/* Config 1: latency of a single cfft1f_ + cfft1b_ round trip at N = 1024, measured from C so that no
 * interpreter overhead is in the number.  Three caller situations: device-resident data (two async calls + one
 * synchronize), pageable host array, pinned host array (each call stages in and out and returns synchronously).
 * Build: gcc -O2 -o c1_latency c1_latency.c -I../../include -I/usr/local/cuda/include -L../../cfftpack_b200
 *        -lcfftpack_b200 -L/usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,'$ORIGIN/../../cfftpack_b200' */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "cfftpack_b200.h"

static double now_us(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec * 1e6 + t.tv_nsec * 1e-3;
}
static int cmp(const void *a, const void *b) { return (*(const double *)a > *(const double *)b) - (*(const double *)a < *(const double *)b); }

static void run(const char *label, fft_complex_t *c, int sync, int n, double *ws, int lensav) {
  int inc = 1, lenwrk = 2 * n, ier = 0, reps = 2000, i;
  double dummy = 0, *t = (double *)malloc(sizeof(double) * reps);
  for (i = -200; i < reps; ++i) {
    double t0 = now_us();
    cfft1f_(&n, &inc, c, &n, ws, &lensav, &dummy, &lenwrk, &ier);
    cfft1b_(&n, &inc, c, &n, ws, &lensav, &dummy, &lenwrk, &ier);
    if (sync) cfb200_synchronize();
    if (i >= 0) t[i] = now_us() - t0;
    if (ier) { printf("ier=%d %s\n", ier, cfb200_last_error()); exit(1); }
  }
  qsort(t, reps, sizeof(double), cmp);
  printf("cfft1f+cfft1b N=%d round trip, %-28s median %7.2f us  p10 %7.2f  p90 %7.2f\n", n, label, t[reps / 2], t[reps / 10], t[reps * 9 / 10]);
  free(t);
}

int main(int argc, char **argv) {
  int n = argc > 1 ? atoi(argv[1]) : 1024, lensav = 2 * n + (int)(log((double)n) / log(2.0)) + 4, ier = 0, i;
  double *ws = (double *)malloc(sizeof(double) * lensav);
  fft_complex_t *h = (fft_complex_t *)malloc(sizeof(fft_complex_t) * n), *p, *d;
  cfft1i_(&n, ws, &lensav, &ier);
  for (i = 0; i < n; ++i) { h[i].r = i + 1.0; h[i].i = 0.0; }
  cudaMalloc((void **)&d, sizeof(fft_complex_t) * n);
  cudaMemcpy(d, h, sizeof(fft_complex_t) * n, cudaMemcpyHostToDevice);
  cudaMallocHost((void **)&p, sizeof(fft_complex_t) * n);
  for (i = 0; i < n; ++i) p[i] = h[i];
  run("device pointer (2 calls+sync)", d, 1, n, ws, lensav);
  run("pinned host array", p, 0, n, ws, lensav);
  run("pageable host array", h, 0, n, ws, lensav);
  {
    double e = 0;
    for (i = 0; i < n; ++i) e = fmax(e, fabs(p[i].r - (i + 1.0)));
    printf("max drift after 2200 round trips: %.3e\n", e);
  }
  return 0;
}
