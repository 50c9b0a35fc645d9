/* engine_types.h -- plain-data descriptions shared by the plans (host) and the engine kernel (device). */
#ifndef CFB_ENGINE_TYPES_H
#define CFB_ENGINE_TYPES_H
#include "cfb_rt.h"

namespace cfb {

typedef double2 cpx;

enum Kind { K_C2C = 0, K_RFFT = 1, K_COST = 2, K_SINT = 3, K_COSQ = 4, K_SINQ = 5 };

#define CFB_MAXPASS 24
#define CFB_ENGINE_THREADS 256
#define CFB_ENGINE_REAL_MAXTHREADS 384  /* real-family engine kernel: 256, or 384 when a tile holds four rows */

/* generic odd radix: output pairs (k, r-k) one thread computes together from each pair of inputs it loads */
#define CFB_GENERIC_KB 4
static inline int generic_items(int r) { return 1 + ((r + 1) / 2 - 1 + CFB_GENERIC_KB - 1) / CFB_GENERIC_KB; }

struct PassDesc {
  int radix;  // 2,3,4,5,8 or a generic odd prime
  int s;      // product of the radices of earlier passes
  int m;      // remaining length / radix
  int twoff;  // offset of this pass's twiddles in the plan table, laid out [k-1][p], p < m
  int rtoff;  // generic radix only: offset of the table exp(-2 pi i j / radix), j < radix
  unsigned mag_s;   // ceil(2^32 / s):  b / s == __umulhi(b, mag_s) for b * s < 2^32 (s > 1)
  unsigned mag_nb;  // ceil(2^32 / nb), nb = M / radix butterflies per sequence
  unsigned mag_per; // generic radix: ceil(2^32 / (nb * generic_items(radix)))
};

/* element (g, e) of a batch lives at  (g / nlo) * jump_hi + (g % nlo) * jump_lo + e * inc  (units: elements) */
struct Addr {
  long long inc, jump_lo, jump_hi;
  int nlo;
  int lanes_t;  // 1: consecutive threads walk the batch axis (jump_lo is the small stride)
};

struct EngineParams {
  int kind, dir;  // dir: -1 forward, +1 backward (user-level direction)
  int n;          // user sequence length
  int M;          // length of the complex core transform (n, n-1 for cost, n+1 for sint)
  int nf;
  int T;    // complex sequences (c2c) or PAIRS of real sequences (other kinds) per CTA
  int ldz;  // row pitch of the complex buffers (odd, >= max(M, n)); real rows use the same pitch in doubles
  int tx_in_log2, tx_out_log2;  // loader / storer thread tiling: 2^tx threads walk the contiguous axis
  int padshift;                 // shared-memory rows get one pad slot every 2^padshift elements (31 = none)
  int tw_smem;                  // > 0: number of plan twiddles copied to shared memory by each CTA
  long long ntiles;             // tiles of T sequences (pairs); CTAs are persistent and stride over them
  int aligned16;  // c2c: in and out are 16-byte aligned
  long long lot;  // sequences in the batch
  Addr ain, aout;
  const void *in;
  void *out;
  const cpx *tw;       // twiddles of all passes, forward convention
  const double *trig;  // kind tables (see plan.cpp)
  double scale;        // c2c: factor applied on store
  // four-step: element e of row g is multiplied on store by W_n^(j*e), j = the row's index along the split axis
  // (g % nlo or g / nlo, fs_from_hi).  W_n^x = fs_tw[x & (2^fs_shift - 1)] * fs_tw[2^fs_shift + (x >> fs_shift)]
  const cpx *fs_tw;
  int fs_n, fs_shift, fs_from_hi;
  int fs_smem;  // > 0: entries of fs_tw copied to shared memory by each CTA
  PassDesc pass[CFB_MAXPASS];
};

}  // namespace cfb
#endif
