// Access-pattern probe for the four-step tile sweeps: copy 4 GiB of double2 where each CTA moves a tile of
// ROWS consecutive "rows" (16 B each, contiguous) x 128 elements at element stride ESTRIDE (in double2), i.e.
// ROWS*16-byte chunks at a large stride -- the pattern of pow2_tile_kernel's loads and stores.  Measurement tool only.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)
template <int ROWS>
__global__ void __launch_bounds__(256) tile_copy(const double2 *__restrict__ in, double2 *__restrict__ out, long long estride,
                                                 long long ntiles, long long tiles_per_group) {
  // tile id -> (group, chunk): group selects a 128*estride block, chunk selects ROWS consecutive columns inside it
  const int tl = threadIdx.x % ROWS, t = threadIdx.x / ROWS;  // rows fastest
  constexpr int NT = 256 / ROWS, PER = 128 / NT;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long grp = tile / tiles_per_group, ch = tile % tiles_per_group;
    const long long base = grp * 128 * estride + ch * ROWS + tl;
    double2 v[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = in[base + (long long)(t + NT * i) * estride];
#pragma unroll
    for (int i = 0; i < PER; ++i) out[base + (long long)(t + NT * i) * estride] = v[i];
  }
}
int main() {
  const size_t n = (size_t)1 << 28;
  double2 *a, *b; CK(cudaMalloc(&a, n * 16)); CK(cudaMalloc(&b, n * 16));
  CK(cudaMemset(a, 0, n * 16)); CK(cudaMemset(b, 0, n * 16));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms;
  const long long estrides[] = {128, 16384, 2097152};  // 2 KB, 256 KB, 32 MB between elements of a row
  for (long long es : estrides) {
    for (int rows : {16, 32, 64}) {
      const long long tiles_per_group = es / rows;  // chunks of `rows` columns inside one group of 128*es elements
      const long long ntiles = (long long)(n / 128 / rows);
      for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0));
        if (rows == 16) tile_copy<16><<<148 * 8, 256>>>(a, b, es, ntiles, tiles_per_group);
        if (rows == 32) tile_copy<32><<<148 * 8, 256>>>(a, b, es, ntiles, tiles_per_group);
        if (rows == 64) tile_copy<64><<<148 * 8, 256>>>(a, b, es, ntiles, tiles_per_group);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
      }
      printf("estride %8lld (x16 B)  chunk %4d B: %.3f ms -> %.1f GB/s\n", es, rows * 16, ms, 2.0 * n * 16 / ms / 1e6);
    }
  }
  CK(cudaGetLastError());
  return 0;
}
