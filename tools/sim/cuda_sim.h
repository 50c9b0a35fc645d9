/*
 * tools/sim/cuda_sim.h -- CUDA-thread emulator for CPU-side tests.  TEST INFRASTRUCTURE ONLY.
 *
 * The kernel sources under cfftpack_b200/csrc are written once, in CUDA C++.  This header lets
 * the SAME translation units be compiled by g++ (-DCFB_SIM) so that `-m "not gpu"` tests can
 * exercise the host logic (argument checks, plans, launch geometry, stride handling) and the
 * kernels' index arithmetic without a GPU.  Every CUDA thread of a block becomes a ucontext
 * fiber; __syncthreads() and the warp shuffles yield to a round-robin scheduler.  Blocks run
 * one after another.  It is slow (a debugging aid), it is never built into or loaded by the
 * product library libcfftpack_b200.so, and nothing in the product falls back to it.
 */
#ifndef CFB_CUDA_SIM_H
#define CFB_CUDA_SIM_H
#ifndef CFB_SIM
#error "cuda_sim.h is only for -DCFB_SIM builds"
#endif
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(x) __attribute__((aligned(x)))
#define __constant__ static

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct alignas(16) double2 {
  double x, y;
};
static inline double2 make_double2(double a, double b) {
  double2 r;
  r.x = a;
  r.y = b;
  return r;
}

namespace cfbsim {
struct Ctx {
  dim3 tidx, bidx, bdim, gdim;
};
extern Ctx *cur;              // context of the running fiber
extern char *dyn_smem;        // dynamic shared memory of the running block
void yield_barrier();         // __syncthreads
void warp_barrier();          // internal, for shuffles
void yield_once();            // let the other fibers run (spin-wait emulation)
double shfl(double v, int src_lane);
void run(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body);
}  // namespace cfbsim

#define threadIdx (cfbsim::cur->tidx)
#define blockIdx (cfbsim::cur->bidx)
#define blockDim (cfbsim::cur->bdim)
#define gridDim (cfbsim::cur->gdim)
#define __syncthreads() cfbsim::yield_barrier()
#define __syncwarp(...) cfbsim::warp_barrier()
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
template <class T>
static inline T __ldg(const T *p) {
  return *p;
}
static inline double __shfl_sync(unsigned, double v, int lane) { return cfbsim::shfl(v, lane); }
static inline double __shfl_xor_sync(unsigned, double v, int m) {
  return cfbsim::shfl(v, (int)((threadIdx.x & 31) ^ m));
}
static inline double __shfl_up_sync(unsigned, double v, int d) {
  int l = (int)(threadIdx.x & 31);
  return cfbsim::shfl(v, l >= d ? l - d : l);
}
static inline double __shfl_down_sync(unsigned, double v, int d) {
  int l = (int)(threadIdx.x & 31);
  return cfbsim::shfl(v, l + d < 32 ? l + d : l);
}
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }

/* ---- the sliver of the runtime API the host code uses ---- */
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
  cudaMemoryType type;
  int device;
};
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
static inline const char *cudaGetErrorString(cudaError_t) { return "sim"; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaMalloc(void **p, size_t n) {
  *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
  return *p ? 0 : 2;
}
static inline cudaError_t cudaFree(void *p) {
  free(p);
  return 0;
}
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
enum { cudaHostAllocPortable = 1, cudaHostAllocMapped = 2 };
static inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) {
  memmove(d, s, n);
  return 0;
}
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
  memmove(d, s, n);
  return 0;
}
static inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind,
                                            cudaStream_t = 0) {
  for (size_t i = 0; i < h; ++i) memmove((char *)d + i * dp, (const char *)s + i * sp, w);
  return 0;
}
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) {
  memset(d, v, n);
  return 0;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
  *s = 0;
  return 0;
}
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
#define cudaStreamNonBlocking 1
#define cudaEventDisableTiming 2
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) {
  *e = 0;
  return 0;
}
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaGetDevice(int *d) {
  *d = 0;
  return 0;
}
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) {
  *n = 1;
  return 0;
}
/* every pointer looks like pageable host memory unless registered with cfbsim_mark_device() */
void cfbsim_mark_device(const void *p, size_t bytes);
int cfbsim_is_device(const void *p);
int cfbsim_is_pinned(const void *p);
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
  a->type = cfbsim_is_device(p) ? cudaMemoryTypeDevice : cfbsim_is_pinned(p) ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
  a->device = 0;
  return 0;
}
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
  return 0;
}
struct cudaDeviceProp {
  int multiProcessorCount;
  size_t sharedMemPerBlockOptin;
  int major, minor;
  char name[64];
};
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  p->multiProcessorCount = 148;
  p->sharedMemPerBlockOptin = 232448;
  p->major = 10;
  p->minor = 0;
  strcpy(p->name, "cfb-sim");
  return 0;
}

#define CFB_DYN_SMEM(name) char *name = cfbsim::dyn_smem
#define CFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  cfbsim::run(dim3(grid), dim3(block), (size_t)(smem), [&]() { kernel(__VA_ARGS__); })

#endif
