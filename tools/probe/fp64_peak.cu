/* FP64 vector peak of the device (DFMA throughput), the denominator for the FP64-pipe utilisation of the generic-radix
 * passes ("FP64 vector peak not measured by the driver -- measure before quoting pipe utilisation").
 *   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/probe/fp64_peak.cu -o tools/probe/fp64_peak && ./tools/probe/fp64_peak
 * Each thread runs 16 independent FMA chains; CUDA events around the launch; prints TFLOP/s (2 flops per FMA) and
 * FMA warp-instructions per clock per SM. */
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
  double v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double *out;
  cudaMalloc(&out, 8);
  const int iters = 20000, threads = 256;
  for (int per_sm = 1; per_sm <= 8; per_sm *= 2) {
    const int blocks = p.multiProcessorCount * per_sm;
    dfma_kernel<<<blocks, threads>>>(out, 100, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)blocks * threads * 16.0 * iters;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s: %d SMs, %d CTAs/SM x %d threads: %.3f ms, %.2f TFLOP/s FP64 (FMA = 2 flops), %.2f FMA warp-instr/clk/SM at %d MHz nominal\n",
           p.name, p.multiProcessorCount, per_sm, threads, ms, 2.0 * fma / (ms * 1e-3) / 1e12,
           fma / 32.0 / (ms * 1e-3) / p.multiProcessorCount / (clk * 1e3), clk / 1000);
  }
  return 0;
}
