/* internal.h -- host-side plumbing shared by the translation units of libcfftpack_b200. */
#ifndef CFB_INTERNAL_H
#define CFB_INTERNAL_H
#include "cfb_rt.h"

namespace cfb {

void set_error(const char *fmt, ...);
const char *last_error();
bool cuda_ok(cudaError_t e, const char *what);
#define CFB_CUDA(call)                     \
  do {                                     \
    if (!cfb::cuda_ok((call), #call)) return false; \
  } while (0)

cudaStream_t current_stream();
void set_current_stream(cudaStream_t s);
void count_launch(unsigned long long k = 1);
unsigned long long launch_count();

/* opt a kernel into `smem` bytes of dynamic shared memory and the full shared-memory carveout, once per (kernel, device) */
bool kernel_attrs_ready(const void *kernel, size_t smem);

/* true once a CUDA device is usable; otherwise prints one loud line to stderr (there is no CPU path) */
bool device_ready();
int sm_count();
/* cap on the grid (CTAs) of the tile kernels launched by the calling thread (0 = none): lets the two persistent sweeps
 * of a pipelined sharded transform share the SMs' CTA slots */
void set_tile_cta_cap(int ctas);
int tile_cta_cap();
/* device copy of a host table, complete (not merely staged) on return; nullptr + last_error on failure */
void *upload_table(const void *host, size_t bytes);

/* grow-only per-thread device scratch (four-step intermediates, staging of host arrays) */
void *scratch_get(int slot, size_t bytes);
void scratch_release_all();

/* a caller array, resolved to device memory: device pointers pass through, host arrays are staged (scratch slot 1)
 * on the current stream and copied back + synchronised by view_close */
struct DeviceView {
  void *dev = nullptr;
  void *host = nullptr;
  size_t bytes = 0;
  bool staged = false;
  bool mapped = false;  // short host array placed in the pinned, device-mapped bounce buffer
};
bool view_open(void *user, size_t bytes, DeviceView &v);
bool view_close(DeviceView &v, bool ok);

/* ---- transform drivers (dispatch.cu).  Pointers are DEVICE pointers; strides in elements. ---- */
bool run_c2c(int n, long long lot, long long inc, long long jump, int dir, void *c);
bool run_c2c_scaled(int n, long long lot, long long inc, long long jump, int dir, void *c, double scale);
bool run_real(int kind, int n, long long lot, long long inc, long long jump, int dir, double *x);
bool run_c2c_2d(int ldim, int l, int m, int dir, void *c);
bool run_real_2d(int ldim, int l, int m, int dir, double *r);
bool run_c2c_2d_sharded_phase(int phase, int dir, int l, int m, int rank, int nranks, void *src, void *const *peers);
bool run_c2c_1d_sharded_phase(int phase, int dir, int log2n, int rank, int nranks, void *src, void *const *peers);

/* batched option valuation (option.cu): par = host [8][lot] (S K sigma theta kappa t r flags), value = host [lot] */
int next_fast_even_size(int n);
bool run_option_convolution(int lot, int N, const double *par_host, double *value_host);

/* largest core length the single-kernel paths take (for tests and docs) */
int engine_max_c2c();
int engine_max_real();

/* ---- host wsave initialisers (wsave_init.cpp), bit-compatible with the reference ---- */
int log2_floor_ref(int n);  // the literal (int)(log((double)n)/log(2.0)) of fftpack.c:2221
void wsave_init_complex(int n, double *wsave);
void wsave_init_real(int n, double *wsave);
void wsave_init_cost(int n, double *wsave);
void wsave_init_sint(int n, double *wsave);
void wsave_init_cosq(int n, double *wsave);
bool strides_consistent(int inc, int jump, int n, int lot);  // xercon_, fftpack.c:15210

}  // namespace cfb
#endif
