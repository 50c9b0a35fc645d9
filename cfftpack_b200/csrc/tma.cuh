/*
 * tma.cuh -- the bulk-asynchronous copy engine (TMA, cp.async.bulk) and mbarrier primitives used to stream
 * whole sequences from HBM into shared memory without staging through registers.  sm_90+ PTX; in SASS these
 * appear as UBLKCP / SYNCS (B200_PROFILING.md).  Under -DCFB_SIM (CPU tests) the copy is a memcpy and the
 * barrier is a no-op, because the emulator runs the issuing thread's copy to completion at once.
 */
#ifndef CFB_TMA_CUH
#define CFB_TMA_CUH
#include "cfb_rt.h"

#ifndef CFB_SIM
#include <cuda.h>
#endif

namespace cfb {

/* 3-D tiled tensor map (TMA descriptor) over an array of doubles: dims/strides outermost last.  In CUDA builds this is
 * the driver's CUtensorMap; the emulator keeps the geometry and copies the box itself. */
#ifdef CFB_SIM
struct TensorMap3 {
  const double *base;
  unsigned long long dim[3], stride_bytes[3];  // stride_bytes[0] unused (contiguous)
  unsigned box[3];
};
#else
typedef CUtensorMap TensorMap3;
#endif
/* host: describe `base` as dim0 x dim1 x dim2 doubles with byte strides s1, s2 (multiples of 16) and a box b0 x b1 x 1 */
bool make_tensor_map3(TensorMap3 *tm, const void *base, unsigned long long d0, unsigned long long d1,
                      unsigned long long d2, unsigned long long s1, unsigned long long s2, unsigned b0, unsigned b1);

#ifdef CFB_SIM
/* emulated transaction barrier: low word = completed phases, high word = bytes still expected */
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) { *bar = 0; }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) { *bar += (uint64_t)bytes << 32; }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  while (((unsigned)(*bar) & 1u) == parity) cfbsim::yield_once();
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar) {
  if ((((uintptr_t)smem_dst) | ((uintptr_t)gsrc) | bytes) & 15) {
    fprintf(stderr, "cfbsim: cp.async.bulk needs 16-byte aligned addresses and size (%p <- %p, %u)\n", smem_dst, gsrc, bytes);
    abort();
  }
  memcpy(smem_dst, gsrc, bytes);
  *bar -= (uint64_t)bytes << 32;
  if ((*bar >> 32) == 0) *bar += 1;  // all bytes of this phase have landed
}
__device__ __forceinline__ void bulk_g2s_stream(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar) {
  bulk_g2s(smem_dst, gsrc, bytes, bar);
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *, unsigned) {}  // L2 prefetch: nothing to emulate
/* shared -> global bulk store (emulated: an immediate copy, so the group waits are no-ops) */
__device__ __forceinline__ void fence_async_smem() {}
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, unsigned bytes) {
  if ((((uintptr_t)smem_src) | ((uintptr_t)gdst) | bytes) & 15) {
    fprintf(stderr, "cfbsim: cp.async.bulk (store) needs 16-byte aligned addresses and size (%p <- %p, %u)\n", gdst, smem_src, bytes);
    abort();
  }
  memcpy(gdst, smem_src, bytes);
}
__device__ __forceinline__ void bulk_commit() {}
__device__ __forceinline__ void bulk_wait_read() {}
__device__ __forceinline__ void bulk_wait_all() {}
/* box load: dense [box1][box0] doubles at smem_dst, zero fill outside the tensor */
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const TensorMap3 *tm, uint64_t *bar, int c0, int c1, int c2) {
  double *d = (double *)smem_dst;
  unsigned bytes = 0;
  for (unsigned y = 0; y < tm->box[1]; ++y)
    for (unsigned x = 0; x < tm->box[0]; ++x) {
      const unsigned long long i0 = (unsigned long long)c0 + x, i1 = (unsigned long long)c1 + y, i2 = (unsigned long long)c2;
      double v = 0.0;
      if (i0 < tm->dim[0] && i1 < tm->dim[1] && i2 < tm->dim[2])
        v = *(const double *)((const char *)tm->base + i0 * 8 + i1 * tm->stride_bytes[1] + i2 * tm->stride_bytes[2]);
      d[y * tm->box[0] + x] = v;
      bytes += 8;
    }
  *bar -= (uint64_t)bytes << 32;
  if ((*bar >> 32) == 0) *bar += 1;
}
/* Ampere-style per-thread asynchronous copies (LDGSTS): gathers with arbitrary addresses, no register staging */
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 16); }
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 8); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
#else
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
/* one arrival + the number of bytes the bulk copies of this phase will deliver */
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
/* global -> shared bulk copy; bytes and both addresses are multiples of 16; completion is signalled on bar */
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
/* TMA tensor box load (SASS UTMALDG): one instruction moves the whole strided tile */
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const TensorMap3 *tm, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  // .ca: the two 16-byte halves of a 32-byte sector requested by neighbouring lanes merge in L1 (with .cg every
  // 16-byte request fetched its own sector from L2: 2x read amplification measured)
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
/* same, tagging the lines evict-first in L2: a streamed sequence is read exactly once */
__device__ __forceinline__ void bulk_g2s_stream(void *smem_dst, const void *gsrc, unsigned bytes, uint64_t *bar) {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *gsrc, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(gsrc), "r"(bytes) : "memory");
}
/* ---- shared -> global bulk stores (SASS UBLKCP.G.S): one instruction drains a finished row from shared memory ----
 * every thread that wrote the source executes fence_async_smem() before the barrier that precedes the store */
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_s2g(void *gdst, const void *smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
/* the issuing thread: all of its committed stores have finished READING shared memory (the source may be reused) */
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
/* ... have completed entirely (writes visible) */
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
#endif

/* Split barrier (PTX bar.arrive / bar.sync on a named barrier): the warps that only have to REPORT that they are done
 * with the landing buffer arrive and run on; only the warp that issues the next bulk copy waits.  Under the emulator
 * both are a full barrier (its bulk copies complete at once, so nobody may run ahead). */
#ifdef CFB_SIM
__device__ __forceinline__ void named_arrive(int, int) { __syncthreads(); }
__device__ __forceinline__ void named_sync(int, int) { __syncthreads(); }
#else
__device__ __forceinline__ void named_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
#endif

/* Handing a landing buffer back to the bulk-copy engine.  The engine writes shared memory through the async proxy and
 * may overtake ordinary shared-memory reads that are still queued in the SM's memory pipeline when the CTA passes the
 * barrier that precedes the next bulk copy (observed as rare corrupt tiles).  Every kernel therefore calls this before
 * that barrier with the registers that were filled from the landing buffer: an integer instruction chain that depends on
 * all of them feeds a (never taken, harmless) predicated shared-memory store, so the reads have delivered their values
 * before the barrier -- without the cost of fence.proxy.async, which also drains the previous tile's global stores
 * (measured: rfftmf 0.82 -> 0.88 ms with the fence).  `spare` is any 4-byte shared-memory word the kernel does not use. */
/* WHOLE: every a[i] was filled by ONE 16-byte load, so one of its four registers stands for the load */
template <int P, bool WHOLE = false>
__device__ __forceinline__ void landing_reads_done(const double2 (&a)[P], volatile unsigned *spare) {
#ifndef CFB_SIM
  unsigned acc = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    acc ^= (unsigned)__double2loint(a[i].x);
    if (!WHOLE) acc ^= (unsigned)__double2hiint(a[i].y);
  }
  if (acc == 0x9e3779b9u) *spare = acc;
#endif
}

}  // namespace cfb
#endif
