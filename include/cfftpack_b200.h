/*
 * cfftpack_b200.h -- C ABI of libcfftpack_b200.so, the B200 (sm_100a) implementation of the FFTPACK 5.1
 * transform path of zywina/cfftpack.
 *
 * The entry points below are the ones the reference exports from cfftpack/fftpack.c and declares in
 * cfftpack/fftpack.h:80-169 (plus the batched *m* routines it exports without declaring).  Names, argument
 * order, argument meaning and ier codes are the reference's: Fortran convention, every scalar by pointer,
 * arrays by pointer, status through *ier, return value always 0.  A program written against fftpack.h
 * relinks against this library unchanged.
 *
 * Data arrays (c / r / x) may be HOST pointers (the library stages them through HBM: copy in, transform,
 * copy out, synchronous like the reference) or DEVICE pointers (transformed in place, asynchronously on the
 * stream set with cfb200_set_stream, default stream otherwise).  wsave is always a host array; work is
 * accepted for length checking and never touched.
 *
 * ier: 0 ok; 1 data array too short; 2 lensav too short; 3 lenwrk too short; 4 (inc,jump,n,lot)
 * inconsistent; 5 l > ldim (2-D); 20 failure in an inner call; -1 CUDA failure (see cfb200_last_error()).
 * Deliberate deviation: on ier != 0 no data is touched, for every routine.
 */
#ifndef CFFTPACK_B200_H
#define CFFTPACK_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fft_real_t; /* cfftpack/fftpack.h:59-64 (FP64 build) */
typedef struct {
  fft_real_t r, i;
} fft_complex_t; /* cfftpack/fftpack.h:72-75 */

/* ---- complex 1-D: cfftpack/fftpack.h:83-94, fftpack.c:2151 (cfft1b_), :2199 (cfft1f_), :2247 (cfft1i_) ---- */
int cfft1i_(int *n, fft_real_t *wsave, int *lensav, int *ier);
int cfft1f_(int *n, int *inc, fft_complex_t *c, int *lenc, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);
int cfft1b_(int *n, int *inc, fft_complex_t *c, int *lenc, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);

/* ---- complex batched: fftpack.c:2609 (cfftmi_), :2554 (cfftmf_), :2499 (cfftmb_); exported, not in fftpack.h ---- */
int cfftmi_(int *n, fft_real_t *wsave, int *lensav, int *ier);
int cfftmf_(int *lot, int *jump, int *n, int *inc, fft_complex_t *c, int *lenc, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);
int cfftmb_(int *lot, int *jump, int *n, int *inc, fft_complex_t *c, int *lenc, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);

/* ---- complex 2-D: cfftpack/fftpack.h:97-109, fftpack.c:2443 (cfft2i_), :2363 (cfft2f_), :2285 (cfft2b_) ---- */
int cfft2i_(int *l, int *m, fft_real_t *wsave, int *lensav, int *ier);
int cfft2f_(int *ldim, int *l, int *m, fft_complex_t *c, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);
int cfft2b_(int *ldim, int *l, int *m, fft_complex_t *c, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);

/* ---- real 2-D: fftpack.c:13454 (rfft2i_), :13282 (rfft2f_), :13113 (rfft2b_); exported, not in fftpack.h.
 * r(ldim, m) real column-major; result along i in half-complex order [Re0, Re1, Im1, ...] of X/(l*m), rows 0 and
 * (l even) l-1 half-complex along j as well.  lensav >= l+L2(l)+4 + 2m+L2(m)+4 + m+L2(m)+4, lenwrk >= (l+1)*m.
 * Unlike the reference (which uses r as scratch, :13407) rows l..ldim-1 of r are left untouched. ---- */
int rfft2i_(int *l, int *m, fft_real_t *wsave, int *lensav, int *ier);
int rfft2f_(int *ldim, int *l, int *m, fft_real_t *r, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);
int rfft2b_(int *ldim, int *l, int *m, fft_real_t *r, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);

/* ---- real 1-D: cfftpack/fftpack.h:150-157, fftpack.c:13076 (rfft1i_), :13030 (rfft1f_), :12984 (rfft1b_) ---- */
int rfft1i_(int *n, fft_real_t *wsave, int *lensav, int *ier);
int rfft1f_(int *n, int *inc, fft_real_t *r, int *lenr, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);
int rfft1b_(int *n, int *inc, fft_real_t *r, int *lenr, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);

/* ---- real batched: fftpack.c:14086 (rfftmi_), :14035 (rfftmf_), :13984 (rfftmb_); exported, not in fftpack.h ---- */
int rfftmi_(int *n, fft_real_t *wsave, int *lensav, int *ier);
int rfftmf_(int *lot, int *jump, int *n, int *inc, fft_real_t *r, int *lenr, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);
int rfftmb_(int *lot, int *jump, int *n, int *inc, fft_real_t *r, int *lenr, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);

/* ---- DCT-I (cost), DST-I (sint), quarter-wave cosine (cosq = DCT-III/II) and sine (sinq):
 *      cfftpack/fftpack.h:112-147, 160-168; fftpack.c:6107/:6046/:5985 (cost1i_/f_/b_), :6551/:6485/:6419 (costm*),
 *      :14667/:14611/:14553 (sint1*), :15066/:14999/:14931 (sintm*), :5523/:5448/:5374 (cosq1*),
 *      :5931/:5845/:5750 (cosqm*), :14123-14513 (sinq1*, sinqm*) ---- */
#define CFB200_DECL_TRIG(name)                                                                                        \
  int name##1i_(int *n, fft_real_t *wsave, int *lensav, int *ier);                                                    \
  int name##1f_(int *n, int *inc, fft_real_t *x, int *lenx, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier); \
  int name##1b_(int *n, int *inc, fft_real_t *x, int *lenx, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier); \
  int name##mi_(int *n, fft_real_t *wsave, int *lensav, int *ier);                                                    \
  int name##mf_(int *lot, int *jump, int *n, int *inc, fft_real_t *x, int *lenx, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier); \
  int name##mb_(int *lot, int *jump, int *n, int *inc, fft_real_t *x, int *lenx, fft_real_t *wsave, int *lensav, fft_real_t *work, int *lenwrk, int *ier);
CFB200_DECL_TRIG(cost)
CFB200_DECL_TRIG(sint)
CFB200_DECL_TRIG(cosq)
CFB200_DECL_TRIG(sinq)

/* ---- extensions (no counterpart in the reference) ---- */
/* Batched option valuation by frequency-domain convolution -- the reference's test/vargamma.c:42-106
 * (conv_bsvg_option: payoff grid -> rfft -> characteristic function -> inverse rfft -> V[N/2] e^{-rt}) for `lot`
 * options in one call, device-resident between the steps.  All arrays are HOST arrays of length lot; flags[o] bit 0:
 * call (else put), bit 1: Black-Scholes (else variance gamma).  The grid has N = fft_next_fast_even_size(n) points
 * (cfftextra.c:42-46); N is the return value.  ier: 0, 1 (bad argument), -1 (CUDA failure). */
int cfb200_option_convolution(int lot, int n, const double *S, const double *K, const double *sigma, const double *theta,
                              const double *kappa, const double *t, const double *r, const int *flags, double *value,
                              int *ier);

/* Sharded cfft2f_/cfft2b_: matrix c(l, m) column-major distributed over `nranks` GPUs of one node.
 * phase 1: local_src = this rank's column slab C[m/nranks][l]; every rank's row slab D[m][l/nranks] is given by
 *          peer_dst[r] (peer-mapped device pointers).  The length-l transforms of the slab are computed and each
 *          result is stored straight into the owning GPU's D (NVLink P2P stores fused into the last pass).
 * phase 2: local_src = this rank's D; peer_dst[r] = every rank's C; length-m transforms, results back into C.
 * The caller synchronises the ranks between the phases (all writes into D/C must have landed).
 * direction < 0: forward (each phase scaled by 1/n like cfftmf_), > 0: backward.  l, m powers of two in 2^12..2^20,
 * divisible by nranks.  ier: 0 ok, 1 bad argument, -1 CUDA failure / unsupported size (cfb200_last_error()). */
int cfb200_cfft2_sharded_phase(int phase, int direction, int l, int m, int rank, int nranks, void *local_src,
                               void *const *peer_dst, int *ier);
/* Very long 1-D complex transform (cfft1f_/cfft1b_ semantics, cfftpack/fftpack.c:2199, :2151) of N = 2^log2n points
 * (24 <= log2n <= 30) distributed in natural order over `nranks` GPUs of one node: rank r holds x[r N/G .. (r+1) N/G).
 * Four-step decomposition N = L * Mm (L = 2^(log2n/2)) with the exchanges fused into the transform kernels as P2P stores
 * (SURVEY 8(e) row 3).  Two symmetric buffers of N/nranks complex elements per rank, X (the data) and Y:
 *   phase 0: local_src = my X, peer_dst = every rank's Y   (transpose only)
 *   phase 1: local_src = my Y, peer_dst = every rank's X   (length-Mm transforms, twiddles W_N^(i b))
 *   phase 2: local_src = my X, peer_dst = every rank's Y   (length-L transforms, natural-order scatter)
 * with a barrier among the ranks after each phase (caller).  The result is in Y, natural order, scaled by 1/N when
 * direction < 0.  ier as for cfb200_cfft2_sharded_phase. */
int cfb200_cfft1_sharded_phase(int phase, int direction, int log2n, int rank, int nranks, void *local_src,
                               void *const *peer_dst, int *ier);
/* Number of GPUs (devices 0 .. n-1 of this process) that batched calls on HOST arrays -- cfftmf_/cfftmb_, rfftm*, costm*,
 * sintm*, cosqm*, sinqm* (cfftpack/fftpack.c:2554, :14035, ...) -- fan out over: the lot axis is cut into one contiguous
 * shard per GPU, each staged and transformed concurrently (no exchange: the sequences are independent).
 * n <= 0: all visible devices.  Default 1, or the environment variable CFB200_DEVICES (a number or "all").  Applies
 * when the calling thread's current device is 0 and the sequences do not interleave (jump >= span of one sequence).
 * Returns the number now in effect (0 without a CUDA device). */
int cfb200_set_devices(int n);
/* Page-locked host memory, usable from every GPU of the process: arrays allocated here are staged by DMA at full PCIe
 * rate (and in overlapping lot-chunks); ordinary malloc'ed arrays work too, through the driver's bounce buffers. */
void *cfb200_host_alloc(size_t bytes);
void cfb200_host_free(void *p);
/* CUDA stream (cudaStream_t) used by THIS host thread for device-pointer calls; NULL = default stream */
int cfb200_set_stream(void *cuda_stream);
/* block until the calling thread's stream is idle; returns 0 or -1 */
int cfb200_synchronize(void);
/* number of kernels this library has launched since it was loaded (all threads) */
unsigned long long cfb200_launch_count(void);
/* message of the last failure on this thread ("" if none) */
const char *cfb200_last_error(void);
/* free cached device plans (all threads) and the calling thread's scratch buffers.  The process must be quiescent: no
 * other host thread inside a transform; the call waits for the device to go idle first.  The worker threads of
 * cfb200_set_devices keep their staging buffers until the process exits. */
void cfb200_release(void);
/* "cfftpack_b200 <version> sm_100a" */
const char *cfb200_version(void);
/* longest complex / real-family sequence that a single CTA transforms on chip */
int cfb200_max_onchip_complex(void);
int cfb200_max_onchip_real(void);

#ifdef __cplusplus
}
#endif
#endif
