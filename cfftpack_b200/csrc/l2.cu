/* l2.cu -- the reference's object wrapper (cfftpack/cfftpack.c) on the device; see include/cfftpack_b200_l2.h.
 * Every call resolves `data` to device memory once, runs the scaling / repack kernels and the transform driver on
 * the current stream, and (for host arrays) copies the result back.  No wsave/work: plans are cached per device. */
#include <math.h>
#include <string.h>

#include <new>

#include "../../include/cfftpack_b200_l2.h"
#include "engine_types.h"
#include "internal.h"

struct FFT_ {
  int algo, n, m, ortho, inc, lot;
};

namespace cfb {

enum { A_CFFT = 1, A_RFFT, A_CFFT2, A_DCT1, A_DCT, A_DCT4, A_DST1, A_DST };  // numbering of cfftintern.h

struct L2Scale {
  double *x;
  long long inc, jump;
  int n, lot;
  double first, rest, last;
  const double *aux;  // dct1: (even, odd) addend per sequence, applied before the factors
};
/* x(o, i) = (x(o, i) + aux) * (first | rest | last) */
__global__ void __launch_bounds__(256) l2_scale_kernel(const L2Scale P) {
  for (int o = blockIdx.y; o < P.lot; o += gridDim.y) {
    double *row = P.x + (long long)o * P.jump;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
      double v = row[(long long)i * P.inc];
      if (P.aux) v += P.aux[2 * o + (i & 1)];
      // the factors compose as in the reference: the general factor first, then the end factors on top of it
      v *= (i == 0 ? P.first : (i == P.n - 1 ? P.last : P.rest));
      row[(long long)i * P.inc] = v;
    }
  }
}
/* cfftpack.c:252-253: addends of the orthonormal DCT-I, from the untransformed ends */
__global__ void __launch_bounds__(256) l2_dct1_ends_kernel(const double *x, long long jump, int n, int lot, double *aux) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= lot) return;
  const double a = x[(long long)o * jump], b = x[(long long)o * jump + n - 1], c = -1.0 + 1.0 / sqrt(2.0);
  aux[2 * o] = (a + b) * c;
  aux[2 * o + 1] = (a - b) * c;
}
/* half-complex row s(n) -> complex[n/2+1] row (cfftpack.c:461-466) and back (:483-486) */
__global__ void __launch_bounds__(256) l2_rfft_unpack_kernel(const double *s, double *out, int n, int ldo, int lot) {
  for (int o = blockIdx.y; o < lot; o += gridDim.y) {
    const double *src = s + (long long)o * n;
    double *dst = out + (long long)o * ldo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ldo; i += gridDim.x * blockDim.x)
      dst[i] = i == 0 ? src[0] : ((i == 1 || i > n) ? 0.0 : src[i - 1]);
  }
}
__global__ void __launch_bounds__(256) l2_rfft_pack_kernel(const double *in, double *s, int n, int ldo, int lot) {
  for (int o = blockIdx.y; o < lot; o += gridDim.y) {
    const double *src = in + (long long)o * ldo;
    double *dst = s + (long long)o * n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = i == 0 ? src[0] : src[i + 1];
  }
}

static dim3 grid2(int n, int lot) {
  return dim3((unsigned)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64), (unsigned)(lot < 16384 ? lot : 16384));
}
static bool scale(double *x, const FFT_ *f, long long inc, double first, double rest, double last, const double *aux = nullptr) {
  L2Scale P;
  P.x = x;
  P.inc = inc;
  P.jump = (long long)f->n * f->inc;
  P.n = f->n;
  P.lot = f->lot;
  P.first = first;
  P.rest = rest;
  P.last = last;
  P.aux = aux;
  CFB_LAUNCH(l2_scale_kernel, grid2(f->n, f->lot), 256, 0, current_stream(), P);
  count_launch();
  return cuda_ok(cudaGetLastError(), "l2_scale_kernel");
}

static fft_t *make(int algo, int n, int m) {
  fft_t *f = new (std::nothrow) FFT_;
  if (!f) return nullptr;
  f->algo = algo;
  f->n = n;
  f->m = m;
  f->ortho = 0;
  f->inc = 1;
  f->lot = 1;
  return f;
}

/* in-place families: the reference's length checks (lenx = n, or n*inc for the DCT), then the device pipeline */
static int run_inplace(fft_t *f, void *data, int fwd, int want_algo, bool check_algo) {
  if (!f || !data) return -1;
  if (check_algo && f->algo != want_algo) return -2;
  const int n = f->n, inc = f->inc, lot = f->lot, algo = f->algo;
  const long long jump = (long long)n * inc;
  const int lenx = algo == A_DCT ? n * inc : n;
  if (lenx < (long long)inc * (n - 1) + 1) return (algo == A_DST && !fwd && n > 1) ? 20 : 1;  // ier of the inner routine
  const size_t esz = algo == A_CFFT ? 16 : 8;
  DeviceView v;
  bool ok = view_open(data, (size_t)((lot - 1) * jump + (long long)inc * (n - 1) + 1) * esz, v);
  double *x = (double *)v.dev;
  const double dn = (double)n;
  if (ok) switch (algo) {
    case A_CFFT:
      ok = run_c2c_scaled(n, lot, inc, jump, fwd ? -1 : +1, x,
                          fwd ? (f->ortho ? 1.0 / dn / sqrt(dn) : 1.0 / dn) : (f->ortho ? sqrt(dn) : 1.0));
      break;
    case A_DCT:
      if (fwd) ok = (!f->ortho || scale(x, f, inc, sqrt(dn), sqrt(0.5 * dn), sqrt(0.5 * dn))) && run_real(K_COSQ, n, lot, inc, jump, -1, x);
      else ok = run_real(K_COSQ, n, lot, inc, jump, +1, x) && (!f->ortho || scale(x, f, inc, 1.0 / sqrt(dn), sqrt(2.0 / dn), sqrt(2.0 / dn)));
      break;
    case A_DCT1:
      if (!f->ortho) ok = run_real(K_COST, n, lot, inc, jump, fwd ? -1 : +1, x);
      else {  // cfftpack.c:245-275: both directions through the unscaled backward transform
        double *aux = (double *)scratch_get(7, (size_t)lot * 16);
        const double m = sqrt(2.0 / (dn - 1.0)), r2 = 1.0 / sqrt(2.0);
        ok = aux != nullptr;
        if (ok) {
          CFB_LAUNCH(l2_dct1_ends_kernel, (unsigned)((lot + 255) / 256), 256, 0, current_stream(), x, jump, n, lot, aux);
          count_launch();
          ok = run_real(K_COST, n, lot, inc, jump, +1, x) && scale(x, f, inc, m * r2, m, m * r2, aux);
        }
      }
      break;
    case A_DST:
      if (fwd) ok = (!f->ortho || scale(x, f, 1, sqrt(1.0 / dn), sqrt(0.5 / dn), sqrt(0.5 / dn))) && run_real(K_SINQ, n, lot, inc, jump, -1, x) &&
                    (!f->ortho || scale(x, f, 1, dn, dn, dn));
      else ok = run_real(K_SINQ, n, lot, inc, jump, +1, x) && (!f->ortho || scale(x, f, 1, sqrt(1.0 / dn), sqrt(2.0 / dn), sqrt(2.0 / dn)));
      break;
    case A_DST1: {
      const bool unscaled = !fwd || f->ortho;  // dst1_forward with ortho is dst1_inverse (cfftpack.c:399-401)
      const double m = sqrt(2.0 / (dn + 1.0));
      ok = run_real(K_SINT, n, lot, inc, jump, unscaled ? +1 : -1, x) && (!f->ortho || scale(x, f, 1, m, m, m));
      break;
    }
    default: ok = false; set_error("l2: unknown algorithm %d", algo);
  }
  ok = view_close(v, ok);
  return ok ? 0 : -1;
}

static int run_2d(fft_t *f, fft_complex_t *data, int dir) {
  if (!f || !data) return -1;
  const size_t plane = (size_t)f->n * f->m;
  DeviceView v;
  bool ok = view_open(data, plane * f->lot * 16, v);
  for (int o = 0; ok && o < f->lot; ++o) ok = run_c2c_2d(f->n, f->n, f->m, dir, (char *)v.dev + plane * o * 16);
  ok = view_close(v, ok);
  return ok ? 0 : -1;
}

}  // namespace cfb

using namespace cfb;

#pragma GCC visibility push(default)
extern "C" {

void fft_free(fft_t *f) { delete f; }
void fft_ortho(fft_t *f, bool ortho) {
  if (f) f->ortho = ortho;
}
void fft_stride(fft_t *f, int stride) {
  if (f) f->inc = stride > 0 ? stride : 1;
}
void cfb200_fft_batch(fft_t *f, int lot) {
  if (f) f->lot = lot > 0 ? lot : 1;
}

fft_t *fft_create(int size) { return size <= 0 ? nullptr : make(A_CFFT, size, 0); }
int fft_forward(fft_t *f, void *data) { return run_inplace(f, data, 1, A_CFFT, true); }
int fft_inverse(fft_t *f, void *data) { return run_inplace(f, data, 0, A_CFFT, true); }

fft_t *fft2_create(int M, int N) { return (M <= 0 || N <= 0) ? nullptr : make(A_CFFT2, M, N); }
int fft2_forward(fft_t *f, fft_complex_t *data) { return run_2d(f, data, -1); }
int fft2_inverse(fft_t *f, fft_complex_t *data) { return run_2d(f, data, +1); }

fft_t *dct_create(int size) { return size <= 0 ? nullptr : make(A_DCT, size, 0); }
int dct_forward(fft_t *f, fft_real_t *data) { return run_inplace(f, data, 1, A_DCT, true); }
int dct_inverse(fft_t *f, fft_real_t *data) { return run_inplace(f, data, 0, A_DCT, true); }

fft_t *dct1_create(int size) { return size <= 1 ? nullptr : make(A_DCT1, size, 0); }
int dct1_forward(fft_t *f, fft_real_t *data) { return run_inplace(f, data, 1, A_DCT1, true); }
int dct1_inverse(fft_t *f, fft_real_t *data) { return run_inplace(f, data, 0, A_DCT1, true); }

/* dst_* do not check the handle's algorithm upstream (cfftpack.c:330-371); a handle of another family is an error here */
fft_t *dst_create(int size) { return size <= 0 ? nullptr : make(A_DST, size, 0); }
int dst_forward(fft_t *f, fft_real_t *data) { return run_inplace(f, data, 1, A_DST, true); }
int dst_inverse(fft_t *f, fft_real_t *data) { return run_inplace(f, data, 0, A_DST, true); }

fft_t *dst1_create(int size) { return size <= 0 ? nullptr : make(A_DST1, size, 0); }
int dst1_forward(fft_t *f, fft_real_t *data) { return run_inplace(f, data, 1, A_DST1, true); }
int dst1_inverse(fft_t *f, fft_real_t *data) { return run_inplace(f, data, 0, A_DST1, true); }

fft_t *rfft_create(int size) { return size <= 0 ? nullptr : make(A_RFFT, size, 0); }
int rfft_forward(fft_t *f, const fft_real_t *inp, void *outp) {
  if (!f || !inp || !outp) return -1;
  if (f->algo != A_RFFT) return -2;
  const int n = f->n, lot = f->lot;
  if (!device_ready()) return -1;
  double *s = (double *)scratch_get(7, (size_t)lot * n * 8);
  if (!s) return -1;
  cudaStream_t st = current_stream();
  bool ok = cuda_ok(cudaMemcpyAsync(s, inp, (size_t)lot * n * 8, cudaMemcpyDefault, st), "cudaMemcpyAsync(rfft input)") &&
            run_real(K_RFFT, n, lot, 1, n, -1, s);
  DeviceView v;
  // odd n: the reference writes n + 1 doubles per row ((n+1)/2 complex); even n: n + 2
  const int row = n % 2 ? n + 1 : n + 2;
  ok = ok && view_open(outp, (size_t)lot * row * 8, v);
  if (ok) {
    CFB_LAUNCH(l2_rfft_unpack_kernel, grid2(row, lot), 256, 0, st, (const double *)s, (double *)v.dev, n, row, lot);
    count_launch();
    ok = cuda_ok(cudaGetLastError(), "l2_rfft_unpack_kernel");
  }
  ok = view_close(v, ok);
  return ok ? 0 : -1;
}
int rfft_inverse(fft_t *f, const void *inp, fft_real_t *outp) {
  if (!f || !inp || !outp) return -1;
  if (f->algo != A_RFFT) return -2;
  const int n = f->n, lot = f->lot, row = n % 2 ? n + 1 : n + 2;
  if (!device_ready()) return -1;
  double *s = (double *)scratch_get(7, (size_t)lot * n * 8);
  if (!s) return -1;
  cudaStream_t st = current_stream();
  DeviceView v;
  bool ok = view_open(const_cast<void *>(inp), (size_t)lot * row * 8, v);
  if (ok) {
    CFB_LAUNCH(l2_rfft_pack_kernel, grid2(n, lot), 256, 0, st, (const double *)v.dev, s, n, row, lot);
    count_launch();
    ok = cuda_ok(cudaGetLastError(), "l2_rfft_pack_kernel") && run_real(K_RFFT, n, lot, 1, n, +1, s) &&
         cuda_ok(cudaMemcpyAsync(outp, s, (size_t)lot * n * 8, cudaMemcpyDefault, st), "cudaMemcpyAsync(rfft output)");
  }
  // the input view is read-only (no copy back); wait only when a host array is involved on either side
  cudaPointerAttributes at;
  memset(&at, 0, sizeof(at));
  const bool out_on_device = cudaPointerGetAttributes(&at, outp) == cudaSuccess &&
                             (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged);
  cudaGetLastError();
  if (v.staged || !out_on_device) ok = cuda_ok(cudaStreamSynchronize(st), "cudaStreamSynchronize") && ok;
  return ok ? 0 : -1;
}

}  // extern "C"
#pragma GCC visibility pop
