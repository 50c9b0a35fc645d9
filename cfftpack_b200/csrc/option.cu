/* option.cu -- batched option valuation by frequency-domain convolution, the application behind config 3
 * (reference: test/vargamma.c:42-106 conv_bsvg_option, one option per call, on the host).
 * Here `lot` options are valued at once, device-resident from payoff to price:
 *   payoff grid V(o, i) -> rfftmf_ -> multiply by the characteristic function on the half-complex pairs -> rfftmb_
 *   -> value(o) = V(o, N/2) exp(-r t).
 * The reference goes through rfft_forward/rfft_inverse (cfftpack.c:446-492), which only shift the half-complex
 * vector by one slot; the shift is folded into the indexing of the multiply kernel. */
#include <math.h>
#include <string.h>

#include "engine_types.h"
#include "internal.h"

namespace cfb {

struct OptionParams {
  double *V;            // [lot][N]
  const double *par;    // [8][lot]: S K sigma theta kappa t r flags
  double *value;        // [lot]
  int lot, N;
};

__global__ void __launch_bounds__(256) option_payoff_kernel(const OptionParams P) {
  const int N2 = P.N / 2;
  for (int o = blockIdx.y; o < P.lot; o += gridDim.y) {
    const double S = P.par[o], K = P.par[P.lot + o], sigma = P.par[2 * P.lot + o], t = P.par[5 * P.lot + o];
    const bool call = ((int)P.par[7 * P.lot + o]) & 1;
    const double ds = 2 * 10 * sigma * sqrt(t) / P.N, lS = log(S);
    double *row = P.V + (size_t)o * P.N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.N; i += gridDim.x * blockDim.x) {
      const double e = exp(lS + (N2 - i) * ds);
      const double v = call ? e - K : K - e;
      row[i] = v > 0.0 ? v : 0.0;
    }
  }
}

/* half-complex pairs (r[2i-1], r[2i]) *= phi(i du); the ends i = 0, N/2 are real slots and keep the real part */
__global__ void __launch_bounds__(256) option_charfn_kernel(const OptionParams P) {
  const int N2 = P.N / 2;
  for (int o = blockIdx.y; o < P.lot; o += gridDim.y) {
    const double sigma = P.par[2 * P.lot + o], theta = P.par[3 * P.lot + o], kappa = P.par[4 * P.lot + o];
    const double t = P.par[5 * P.lot + o], r = P.par[6 * P.lot + o];
    const bool bs = (((int)P.par[7 * P.lot + o]) >> 1) & 1;
    const double ds = 2 * 10 * sigma * sqrt(t) / P.N, du = 2 * M_PI / (ds * P.N);
    const double drift = bs ? r - 0.5 * sigma * sigma : r + (1.0 / kappa) * log(1.0 - sigma * sigma * kappa / 2.0 - theta * kappa);
    double *row = P.V + (size_t)o * P.N;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= N2; i += gridDim.x * blockDim.x) {
      const double u = i * du;
      double mag, ang;
      if (bs) {
        mag = exp(-0.5 * sigma * sigma * u * u * t);
        ang = u * t * drift;
      } else {  // (1 + sigma^2 kappa u^2 / 2 - i theta kappa u)^(-t/kappa) e^(i drift u t)
        const double zr = 1.0 + sigma * sigma * kappa * u * u / 2.0, zi = -theta * kappa * u, p = -t / kappa;
        mag = exp(p * log(hypot(zr, zi)));
        ang = p * atan2(zi, zr) + drift * u * t;
      }
      double sn, cs;
      sincos(ang, &sn, &cs);
      const double pr = mag * cs, pi = mag * sn;
      if (i == 0) row[0] *= pr;
      else if (i == N2) row[P.N - 1] *= pr;
      else {
        const double a = row[2 * i - 1], b = row[2 * i];
        row[2 * i - 1] = a * pr - b * pi;
        row[2 * i] = a * pi + b * pr;
      }
    }
  }
}

__global__ void __launch_bounds__(256) option_value_kernel(const OptionParams P) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= P.lot) return;
  const double t = P.par[5 * P.lot + o], r = P.par[6 * P.lot + o];
  P.value[o] = P.V[(size_t)o * P.N + P.N / 2] * exp(-r * t);
}

int next_fast_even_size(int n) {  // cfftextra.c:42-46
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

/* par: host [8][lot]; value: host [lot] */
bool run_option_convolution(int lot, int N, const double *par_host, double *value_host) {
  cudaStream_t st = current_stream();
  long long chunk = ((1LL << 27) / N);  // at most 1 GiB of grid values in flight
  if (chunk < 1) chunk = 1;
  if (chunk > lot) chunk = lot;
  char *base = (char *)scratch_get(7, (size_t)chunk * N * 8 + (size_t)chunk * 9 * 8 + 64);
  if (!base) return false;
  OptionParams P;
  P.V = (double *)base;
  double *par = P.V + (size_t)chunk * N;
  P.par = par;
  P.value = par + 8 * chunk;
  P.N = N;
  for (long long o0 = 0; o0 < lot; o0 += chunk) {
    const int lc = (int)(lot - o0 < chunk ? lot - o0 : chunk);
    P.lot = lc;
    for (int k = 0; k < 8; ++k)
      CFB_CUDA(cudaMemcpyAsync(par + (size_t)k * lc, par_host + (size_t)k * lot + o0, (size_t)lc * 8, cudaMemcpyHostToDevice, st));
    const unsigned gx = (unsigned)((N + 255) / 256 < 64 ? (N + 255) / 256 : 64), gy = (unsigned)(lc < 8192 ? lc : 8192);
    CFB_LAUNCH(option_payoff_kernel, dim3(gx, gy), 256, 0, st, P);
    count_launch();
    if (!run_real(K_RFFT, N, lc, 1, N, -1, P.V)) return false;
    CFB_LAUNCH(option_charfn_kernel, dim3(gx, gy), 256, 0, st, P);
    count_launch();
    if (!run_real(K_RFFT, N, lc, 1, N, +1, P.V)) return false;
    CFB_LAUNCH(option_value_kernel, (unsigned)((lc + 255) / 256), 256, 0, st, P);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "option kernels")) return false;
    CFB_CUDA(cudaMemcpyAsync(value_host + o0, P.value, (size_t)lc * 8, cudaMemcpyDeviceToHost, st));
    CFB_CUDA(cudaStreamSynchronize(st));
  }
  return true;
}

}  // namespace cfb
