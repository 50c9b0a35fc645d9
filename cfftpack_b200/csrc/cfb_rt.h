/* cfb_rt.h -- one include for the CUDA runtime (or, in -DCFB_SIM test builds only, its emulator). */
#ifndef CFB_RT_H
#define CFB_RT_H
#ifdef CFB_SIM
#include "cuda_sim.h"
#else
#include <cuda_runtime.h>
#define CFB_DYN_SMEM(name) extern __shared__ __align__(16) char name[]
#define CFB_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#endif
#ifdef CFB_SIM
#define CFB_GRID_CONSTANT
#else
#define CFB_GRID_CONSTANT __grid_constant__
#endif
#include <stdint.h>
#endif
