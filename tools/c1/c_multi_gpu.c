/* Plain C caller (no CUDA headers): one cfftmf_ call on a host array fans out over all visible GPUs.
 *   gcc -O2 -I include tools/c1/c_multi_gpu.c -L cfftpack_b200 -lcfftpack_b200 -Wl,-rpath,$PWD/cfftpack_b200 -lm -o tools/c1/c_multi_gpu
 *   ./c_multi_gpu [lot] [devices]     (N = 4096; lot = 65536 -> 4 GiB, the BASELINE config 2 batch per GPU)
 * Prints the end-to-end rate (host array in, host array out) for 1 device and for all of them, and checks the round trip. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "cfftpack_b200.h"

static double now(void) {
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return t.tv_sec + 1e-9 * t.tv_nsec;
}

int main(int argc, char **argv) {
  int n = 4096, lot = argc > 1 ? atoi(argv[1]) : 65536, want = argc > 2 ? atoi(argv[2]) : 0;
  int inc = 1, jump = n, lenc = lot * n, lensav = 2 * n + (int)(log((double)n) / log(2.0)) + 4, lenwrk = 2 * n * 8, ier = 0;
  size_t count = (size_t)lot * n;
  fft_complex_t *c = (fft_complex_t *)cfb200_host_alloc(count * sizeof(fft_complex_t));
  double *wsave = (double *)malloc(sizeof(double) * lensav), work[8];
  if (!c || !wsave) {
    fprintf(stderr, "allocation failed (%s)\n", cfb200_last_error());
    return 1;
  }
  lenwrk = 2147483647;  /* the library never touches work; the reference would need 2*lot*n doubles */
  for (size_t i = 0; i < count; ++i) {
    c[i].r = (double)((i * 2654435761u) % 1000003) / 1000003.0 - 0.5;
    c[i].i = (double)((i * 40503u) % 999983) / 999983.0 - 0.5;
  }
  cfftmi_(&n, wsave, &lensav, &ier);
  if (ier) return 2;
  for (int pass = 0; pass < 2; ++pass) {
    int dev = cfb200_set_devices(pass == 0 ? 1 : want);
    cfftmf_(&lot, &jump, &n, &inc, c, &lenc, wsave, &lensav, work, &lenwrk, &ier); /* warm-up: plans, staging buffers */
    if (ier) { fprintf(stderr, "cfftmf_ ier=%d: %s\n", ier, cfb200_last_error()); return 3; }
    cfftmb_(&lot, &jump, &n, &inc, c, &lenc, wsave, &lensav, work, &lenwrk, &ier);
    double t0 = now();
    int reps = 3;
    for (int r = 0; r < reps; ++r) {
      cfftmf_(&lot, &jump, &n, &inc, c, &lenc, wsave, &lensav, work, &lenwrk, &ier);
      cfftmb_(&lot, &jump, &n, &inc, c, &lenc, wsave, &lensav, work, &lenwrk, &ier);
    }
    double dt = (now() - t0) / (2 * reps);
    if (ier) return 4;
    printf("devices=%d: cfftm N=%d lot=%d host array -> host array: %.1f ms per call, %.1f GB/s algorithmic (2 x 16 B x N x lot)\n",
           dev, n, lot, dt * 1e3, 2.0 * 16.0 * count / dt / 1e9);
  }
  /* cfft2f_/cfft2b_ on a host matrix: one GPU, then all of them (column slabs, fused P2P transposes) */
  {
    int l2 = 8192, ld = l2, ls2 = 2 * (2 * l2 + (int)(log((double)l2) / log(2.0)) + 4), lw2 = 2147483647, ier2 = 0;
    size_t cnt2 = (size_t)l2 * l2;
    if (cnt2 <= count) {
      double *ws2 = (double *)malloc(sizeof(double) * ls2);
      cfft2i_(&l2, &l2, ws2, &ls2, &ier2);
      for (int pass = 0; pass < 2 && !ier2; ++pass) {
        int dev = cfb200_set_devices(pass == 0 ? 1 : want);
        cfft2f_(&ld, &l2, &l2, c, ws2, &ls2, work, &lw2, &ier2);
        cfft2b_(&ld, &l2, &l2, c, ws2, &ls2, work, &lw2, &ier2);
        double t0 = now();
        cfft2f_(&ld, &l2, &l2, c, ws2, &ls2, work, &lw2, &ier2);
        cfft2b_(&ld, &l2, &l2, c, ws2, &ls2, work, &lw2, &ier2);
        double dt = (now() - t0) / 2;
        printf("devices=%d: cfft2 %dx%d host matrix -> host matrix: %.1f ms per call (ier %d)\n", dev, l2, l2, dt * 1e3, ier2);
      }
      if (ier2) { fprintf(stderr, "cfft2 ier=%d: %s\n", ier2, cfb200_last_error()); return 6; }
      free(ws2);
    }
  }
  double err = 0, ref = 0;
  for (size_t i = 0; i < count; ++i) {
    double xr = (double)((i * 2654435761u) % 1000003) / 1000003.0 - 0.5, xi = (double)((i * 40503u) % 999983) / 999983.0 - 0.5;
    err += (c[i].r - xr) * (c[i].r - xr) + (c[i].i - xi) * (c[i].i - xi);
    ref += xr * xr + xi * xi;
  }
  printf("round trip after all forward/backward pairs (1-D batched and 2-D): relative L2 error %.2e\n", sqrt(err / ref));
  cfb200_host_free(c);
  free(wsave);
  return sqrt(err / ref) < 1e-12 ? 0 : 5;
}
