// Hardware probe: FP64 FMA peak, FP64 add peak, HBM copy bandwidth, launch latency.
// Test/measurement tool only; not part of the product library.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__global__ void dfma_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void dadd_kernel(double *out, int iters, double a) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
    x0 += a; x1 += a; x2 += a; x3 += a; x4 += a; x5 += a; x6 += a; x7 += a;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
__global__ void copy_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = in[i];
}
__global__ void inplace_kernel(double2 *__restrict__ io, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) { double2 v = io[i]; v.x += 1.0; io[i] = v; }
}
int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sms %d smem/blk optin %zu l2 %d clock %d kHz\n", p.name, p.multiProcessorCount, p.sharedMemPerBlockOptin, p.l2CacheSize, p.clockRate);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double *out; CK(cudaMalloc(&out, 148 * 8 * 1024 * sizeof(double)));
  float ms;
  for (int rep = 0; rep < 3; ++rep) {
    int iters = 20000;
    CK(cudaEventRecord(e0)); dfma_kernel<<<148 * 8, 256>>>(out, iters, 1.0000001, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("dfma: %.3f ms -> %.2f TFLOP/s (FMA=2)\n", ms, 148.0 * 8 * 256 * 8 * iters * 2 / ms / 1e9);
    CK(cudaEventRecord(e0)); dadd_kernel<<<148 * 8, 256>>>(out, iters, 1e-9); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("dadd: %.3f ms -> %.2f Tinstr-lane/s\n", ms, 148.0 * 8 * 256 * 8 * iters / ms / 1e9);
  }
  size_t n = (size_t)1 << 28;  // 4 GiB of double2
  double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
  CK(cudaMemset(a, 0, n * 16)); CK(cudaMemset(b, 0, n * 16));
  for (int blocks = 148 * 4; blocks <= 148 * 32; blocks *= 2) {
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0)); copy_kernel<<<blocks, 512>>>(a, b, n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 2) printf("copy  blocks=%d: %.3f ms -> %.1f GB/s\n", blocks, ms, 2.0 * n * 16 / ms / 1e6);
    }
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0)); inplace_kernel<<<blocks, 512>>>(a, n); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep == 2) printf("inplace blocks=%d: %.3f ms -> %.1f GB/s\n", blocks, ms, 2.0 * n * 16 / ms / 1e6);
    }
  }
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(b, a, n * 16, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("memcpy d2d: %.3f ms -> %.1f GB/s\n", ms, 2.0 * n * 16 / ms / 1e6);
  }
  // pinned host <-> device bandwidth
  void *h; size_t hb = (size_t)1 << 30; CK(cudaMallocHost(&h, hb));
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(a, h, hb, cudaMemcpyHostToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1)); printf("h2d pinned: %.1f GB/s\n", hb / ms / 1e6);
    CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(h, a, hb, cudaMemcpyDeviceToHost)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1)); printf("d2h pinned: %.1f GB/s\n", hb / ms / 1e6);
  }
  printf("done\n");
  return 0;
}
