/* cfftpack_b200_l2.h -- the reference's object wrapper (cfftpack/cfftpack.h:44-281, cfftpack/cfftpack.c) served by the
 * CUDA library directly: same names, arguments, return codes and scaling conventions, so a program
 * written against cfftpack.h links against libcfftpack_b200.so alone.
 *
 * What is different underneath: the handle owns no wsave/work (plans live in the library's device-side cache); `data`
 * may be a host array (staged through HBM, call returns when the result is back) or a DEVICE pointer (asynchronous on
 * the stream of cfb200_set_stream); the orthonormal scalings (cfftpack.c:69-76, 186-193, 210-217, 245-275, ...) and
 * the half-complex <-> complex[n/2+1] repack of rfft_forward/rfft_inverse (cfftpack.c:461-466, 483-486) run on the
 * device, folded into the transform's own scale factor where it has one.
 *
 * Extension: cfb200_fft_batch(f, lot) makes every call on `f` transform `lot` sequences stored back to back (sequence
 * o starts at element o*n*stride; rfft: input rows of n reals, output rows of n/2+1 complex).  lot = 1 is the reference.
 *
 * Return codes as in the reference: 0, -1 (NULL argument), -2 (handle of another algorithm), or the FFTPACK ier
 * (e.g. 1 when a stride > 1 makes the length n passed down too short, as upstream); -1 also reports a CUDA failure
 * (cfb200_last_error()).  On any non-zero code the data is untouched.
 */
#ifndef CFFTPACK_B200_L2_H
#define CFFTPACK_B200_L2_H
#include <stdbool.h>

#include "cfftpack_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct FFT_ fft_t;                                   /* cfftpack.h:44 */
void fft_free(fft_t *f);                                      /* cfftpack.h:52,  cfftpack.c:32 */
void fft_ortho(fft_t *f, bool ortho);                         /* cfftpack.h:67,  cfftpack.c:45 */
void fft_stride(fft_t *f, int stride);                        /* cfftpack.h:78,  cfftpack.c:50 */
void cfb200_fft_batch(fft_t *f, int lot);                     /* extension, see above */

fft_t *fft_create(int size);                                  /* cfftpack.h:91,  cfftpack.c:9 */
int fft_forward(fft_t *f, void *data);                        /* cfftpack.h:104, cfftpack.c:59: (1/n) sum x e^{-i}; ortho: x 1/sqrt(n) more */
int fft_inverse(fft_t *f, void *data);                        /* cfftpack.h:115, cfftpack.c:81: sum x e^{+i}; ortho: x sqrt(n) */
fft_t *fft2_create(int M, int N);                             /* cfftpack.h:124, cfftpack.c:102: column-major M x N, ldim = M */
int fft2_forward(fft_t *f, fft_complex_t *data);              /* cfftpack.h:129 */
int fft2_inverse(fft_t *f, fft_complex_t *data);              /* cfftpack.h:133 */
fft_t *dct_create(int size);                                  /* cfftpack.h:147: cosq */
int dct_forward(fft_t *f, fft_real_t *data);                  /* cfftpack.h:161, cfftpack.c:176 (DCT-III) */
int dct_inverse(fft_t *f, fft_real_t *data);                  /* cfftpack.h:175, cfftpack.c:198 (DCT-II) */
fft_t *dct1_create(int size);                                 /* cfftpack.h:188: cost, size >= 2 */
int dct1_forward(fft_t *f, fft_real_t *data);                 /* cfftpack.h:198, cfftpack.c:277 */
int dct1_inverse(fft_t *f, fft_real_t *data);                 /* cfftpack.h:208, cfftpack.c:294 */
fft_t *dst_create(int size);                                  /* cfftpack.h:219: sinq */
int dst_forward(fft_t *f, fft_real_t *data);                  /* cfftpack.h:222, cfftpack.c:330 */
int dst_inverse(fft_t *f, fft_real_t *data);                  /* cfftpack.h:224, cfftpack.c:354 */
fft_t *dst1_create(int size);                                 /* cfftpack.h:234: sint */
int dst1_forward(fft_t *f, fft_real_t *data);                 /* cfftpack.h:236, cfftpack.c:394 */
int dst1_inverse(fft_t *f, fft_real_t *data);                 /* cfftpack.h:238, cfftpack.c:410 */
fft_t *rfft_create(int size);                                 /* cfftpack.h:265 */
int rfft_forward(fft_t *f, const fft_real_t *inp, void *outp);/* cfftpack.h:273, cfftpack.c:446: out = complex[n/2+1] */
int rfft_inverse(fft_t *f, const void *inp, fft_real_t *outp);/* cfftpack.h:281, cfftpack.c:471 */

#ifdef __cplusplus
}
#endif
#endif
