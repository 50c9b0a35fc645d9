/*
 * dispatch.cu -- chooses a kernel and a launch geometry for each transform request.
 *
 *   contiguous power-of-two batches (cfftmf_/rfftmf_ headline shape)  -> pow2.cuh register kernels
 *   everything else that fits one CTA                                   -> engine.cuh
 *   long complex transforms                                             -> four-step: two engine sweeps
 *                                                                          through a device scratch array
 * There is no CPU path: a failure to launch is reported to the caller through ier.
 */
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "engine.cuh"
#include "internal.h"
#include "plan.h"
#include "pow2.cuh"
#include "radix10.cuh"
#include "mixed.cuh"

namespace cfb {

static const size_t SMEM_MAX = 227 * 1024;      // opt-in limit per CTA on sm_100
static const size_t SMEM_TARGET = 72 * 1024;    // aim: three CTAs per SM for the streaming kernels

int engine_max_c2c() { return (int)((SMEM_MAX - 1024) / 48) - 4; }
int engine_max_real() { return (int)((SMEM_MAX - 1024) / 48) - 4; }

static bool engine_attr_once() { return kernel_attrs_ready((const void *)engine_c2c_kernel, SMEM_MAX); }

static const size_t TW_SMEM_MAX = 16 * 1024;  // plan twiddle tables up to this size are copied to shared memory

static void fill_passes(EngineParams &P, const CorePlan *cp) {
  P.M = cp->M;
  P.nf = cp->nf;
  for (int i = 0; i < cp->nf; ++i) P.pass[i] = cp->pass[i];
  P.tw = cp->d_tw;
  P.tw_smem = (cp->tw_count > 0 && cp->tw_count * sizeof(cpx) <= TW_SMEM_MAX) ? (int)cp->tw_count : 0;
}

static Addr make_addr(long long inc, long long jump_lo, long long jump_hi, long long nlo) {
  Addr a;
  a.inc = inc;
  a.jump_lo = jump_lo;
  a.jump_hi = jump_hi;
  a.nlo = (int)nlo;
  long long ai = inc < 0 ? -inc : inc, aj = jump_lo < 0 ? -jump_lo : jump_lo;
  a.lanes_t = (aj < ai) ? 1 : 0;
  return a;
}

/* split n = n1 * n2 with both factors as close to sqrt(n) as the limit allows; 0 if impossible */
static int four_step_split(int n, int limit) {
  int best = 0;
  for (int d = 1; (long long)d * d <= n; ++d)
    if (n % d == 0 && n / d <= limit) {
      best = d;  // largest d <= sqrt(n); n/d shrinks as d grows
    }
  return best;  // n1 = best (<= n2 = n / best)
}

/* the real-family kernel is instantiated per (family, direction): the family-specific code folds at compile time */
template <int KIND, int DIR>
static bool launch_real_kernel(unsigned grid, unsigned threads, size_t smem, const EngineParams &P) {
  auto kern = engine_kernel<KIND, DIR>;
  if (!kernel_attrs_ready((const void *)kern, SMEM_MAX)) return false;
  CFB_LAUNCH(kern, grid, threads, smem, current_stream(), P);
  return true;
}
static bool launch_real_dispatch(unsigned grid, unsigned threads, size_t smem, const EngineParams &P) {
#define CFB_REAL_CASE(K) \
  case K: return P.dir < 0 ? launch_real_kernel<K, -1>(grid, threads, smem, P) : launch_real_kernel<K, 1>(grid, threads, smem, P);
  switch (P.kind) {
    CFB_REAL_CASE(K_RFFT)
    CFB_REAL_CASE(K_COST)
    CFB_REAL_CASE(K_SINT)
    CFB_REAL_CASE(K_COSQ)
    CFB_REAL_CASE(K_SINQ)
    default: break;
  }
#undef CFB_REAL_CASE
  set_error("unknown real family %d", P.kind);
  return false;
}

static int log2_ceil_capped(long long v, int cap) {
  int l = 0;
  while ((1LL << l) < v && l < cap) ++l;
  return l;
}

/* one engine launch; P has kind/dir/n/addresses/plan filled in */
static bool launch_engine(EngineParams &P) {
  if (!engine_attr_once()) return false;
  const bool real = P.kind != K_C2C;
  const int len = real ? P.n : P.M;  // elements per row that cross the global-memory boundary
  // complex rows: one pad slot per 2^padshift elements when the first radix is a power of two
  P.padshift = 31;
  if (!real && P.nf > 0) {
    const int r1 = P.pass[0].radix;
    if (r1 == 2) P.padshift = 1;
    if (r1 == 4) P.padshift = 2;
    if (r1 == 8) P.padshift = 3;
  }
  const int longest = P.M > P.n ? P.M : P.n;
  P.ldz = (longest + ((longest - 1) >> P.padshift) + 1) | 1;
  if (!real && (size_t)P.ldz * 48 + 64 + 64 + (size_t)(P.tw_smem + P.fs_smem) * sizeof(cpx) > SMEM_MAX) {
    P.padshift = 31;  // a long sequence that only fits without the padding slots
    P.ldz = (longest + 1) | 1;
  }
  // bytes per sequence (c2c: three complex rows -- two landing buffers + the ping-pong partner -- and double-buffered
  // row tables) or per pair (real kinds: two complex rows and the row tables)
  const size_t per = real ? (size_t)P.ldz * 48 + 256 : (size_t)P.ldz * 48 + 64;
  const size_t fixed = 64 + (size_t)(P.tw_smem + P.fs_smem) * sizeof(cpx);
  const long long units = real ? (P.lot + 1) / 2 : P.lot;
  if (per + fixed > SMEM_MAX) {
    set_error("length %d does not fit one CTA (%zu bytes)", P.n, per + fixed);
    return false;
  }
  long long T = (long long)((SMEM_TARGET - (fixed < SMEM_TARGET / 2 ? fixed : 0)) / per);
  const bool strided = P.ain.lanes_t || P.aout.lanes_t;
  // batch-contiguous layouts want at least 8 sequences side by side (128-byte runs of 16-byte elements)
  const long long want = strided ? (real ? 4 : 8) : 1;
  const long long tmax = (long long)((SMEM_MAX - fixed) / per);
  if (T < want) T = tmax < want ? tmax : want;
  if (T < 1) T = 1;
  if (T > 32) T = 32;
  if (strided && T >= 8) {  // whole 128-byte runs along the batch axis: a power of two of rows
    long long p2 = 8;
    while (p2 * 2 <= T) p2 *= 2;
    T = p2;
  }
  if (real) {  // rows = 2T must be a power of two (thread groups of the pre/post-processing); prefer >= 4 rows per CTA
    if (T < 2 && (long long)((SMEM_MAX - fixed) / per) >= 2 && units >= 2) T = 2;
    long long p2 = 1;
    while (p2 * 2 <= T) p2 *= 2;
    T = p2;
  }
  {
    static const int force_T = getenv("CFB200_ENGINE_T") ? atoi(getenv("CFB200_ENGINE_T")) : 0;  // experiments
    if (force_T > 0 && (size_t)force_T * per + fixed <= SMEM_MAX) T = force_T;
  }
  if (T > units) T = units;
  if (real && T > 1) {  // keep the power of two after clamping to the batch
    long long p2 = 1;
    while (p2 * 2 <= T) p2 *= 2;
    T = p2;
  }
  P.T = (int)T;
  const long long rows = real ? 2 * T : T;
  // thread tiling of the loader/storer: threads along the contiguous axis first
  P.tx_in_log2 = log2_ceil_capped(P.ain.lanes_t ? rows : len, 7);
  P.tx_out_log2 = log2_ceil_capped(P.aout.lanes_t ? rows : len, 7);
  const size_t smem = per * (size_t)T + fixed;
  P.ntiles = (units + T - 1) / T;
  // block size: 128 threads when a tile has too little work per pass for 256 (small T x M): more CTAs per SM instead
  static const int force_threads = getenv("CFB200_ENGINE_THREADS") ? atoi(getenv("CFB200_ENGINE_THREADS")) : 0;
  unsigned threads = CFB_ENGINE_THREADS;
  if (!real && (long long)T * P.M <= 1280) threads = 128;  // measured: helps the complex kernel slightly, hurts the real one
  // two pairs (four rows) per tile, i.e. lengths around 1000: 384 threads = three warps per row group make every radix
  // pass of M ~ 1000 a single trip (154-286 butterflies) and shorten the pre/post loops; measured 5-8 % faster than 256
  if (real && T == 2 && P.M >= 512) threads = 384;
  if (force_threads == 128 || force_threads == 256) threads = (unsigned)force_threads;
  long long per_sm = (long long)((SMEM_MAX + 1024) / (smem + 1024));
  const long long reg_limit = threads == 128 ? 6 : (threads > 256 ? 2 : 3);  // 80 registers per thread (launch bounds)
  if (per_sm > reg_limit) per_sm = reg_limit;
  if (per_sm < 1) per_sm = 1;
  long long grid = per_sm * sm_count();
  if (grid > P.ntiles) grid = P.ntiles;
  if (real) {
    if (!launch_real_dispatch((unsigned)grid, threads, smem, P)) return false;
  } else {
    CFB_LAUNCH(engine_c2c_kernel, (unsigned)grid, threads, smem, current_stream(), P);
  }
  count_launch();
  return cuda_ok(cudaGetLastError(), "engine kernel launch");
}

/* long power-of-two transforms: both sweeps of the four-step split run in the register tile kernel (pow2.cuh) */
/* destination of the second sweep when the output is spread over peer GPUs (sharded 2-D transform) */
struct PeerOut {
  int npeers = 0;
  cpx *const *peers = nullptr;
  long long chunk = 0;   // elements of the transform axis owned by each peer
  long long base = 0;    // offset added on every peer
  long long row_inc = 0, row_jump = 0, elem_inc = 0;  // layout on the peers: index = base + m*row_jump + (e % chunk)*elem_inc
};

/* optional second twiddle of the long 1-D decomposition (TileParams::tw2): results of sequence i, output index b, times W_N^(i b) */
struct Tw2 {
  const RootPlan *rp = nullptr;  // roots of N
  long long n = 0, off = 0;      // N, index of the first sequence
};
/* out-of-place form: sequences (inc, jump) of `cin` -> (inc_out, jump_out) of `cout` */
struct FourStepIO {
  const cpx *cin;
  long long inc, jump;
  cpx *cout;
  long long inc_out, jump_out;
};
static bool run_c2c_pow2_four_step_chunk(int n, int a1, int a2, long long lot, const FourStepIO &io, int dir, double scale,
                                         const PeerOut *po, const Tw2 *tw2, int sweep = 0, cpx *scr_in = nullptr);

/* Batches much larger than the L2 cache (126 MB) are walked in lot-chunks whose intermediate (the scratch array
 * between the two sweeps) stays L2-resident: sweep 1 of a chunk reads HBM and writes L2, sweep 2 reads L2 and writes
 * HBM, so the whole transform moves each element over the HBM pins once each way instead of twice.
 * CFB200_FS_CHUNK_MB sets the chunk size (0 = one chunk). */
/* Sharded transforms: the second sweep of a phase is bound by NVLink (its stores go to the peers), the first one by
 * HBM.  The slab is cut into chunks; sweep 1 of chunk c+1 runs on a second stream while sweep 2 of chunk c drains over
 * the links, each kernel holding one CTA per SM, with two intermediate buffers.  CFB200_P2P_CHUNKS sets the number of
 * chunks (default 8; 1 = the two sweeps back to back). */
struct OverlapStreams {
  cudaStream_t st[2] = {0, 0};
  cudaEvent_t ev_start = 0, ev1[16], ev2[16], ev_end[2];
  int dev = -1;
  bool ok = false;
};
static thread_local OverlapStreams t_ov_all[16];  // per device: one host thread may drive several GPUs (cfft2f_ fan-out)
static OverlapStreams &ov_current() {  // the calling thread's set for its current device
  int dev = 0;
  cudaGetDevice(&dev);
  return t_ov_all[dev >= 0 && dev < 16 ? dev : 0];
}
#define t_ov (ov_current())
static bool overlap_ready() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return false;
  if (t_ov.ok && t_ov.dev == dev) return true;
  for (auto &s : t_ov.st) CFB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CFB_CUDA(cudaEventCreateWithFlags(&t_ov.ev_start, cudaEventDisableTiming));
  for (auto &e : t_ov.ev1) CFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : t_ov.ev2) CFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : t_ov.ev_end) CFB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  t_ov.dev = dev;
  t_ov.ok = true;
  return true;
}

static bool run_c2c_pow2_four_step_pipelined(int n, int a1, int a2, long long lot, long long inc, long long jump, int dir, cpx *c,
                                             double scale, const PeerOut *po, const Tw2 *tw2, int chunks) {
  long long per = (lot + chunks - 1) / chunks;
  per = (per + 255) / 256 * 256;
  const int nch = (int)((lot + per - 1) / per);
  if (nch < 2 || nch > 16 || !overlap_ready()) return false;
  cudaStream_t user = current_stream();
  cpx *scr = (cpx *)scratch_get(0, (size_t)2 * per * n * sizeof(cpx));
  if (!scr || !get_root_plan(n)) return false;
  bool ok = cuda_ok(cudaEventRecord(t_ov.ev_start, user), "cudaEventRecord");
  for (int k = 0; k < 2 && ok; ++k) ok = cuda_ok(cudaStreamWaitEvent(t_ov.st[k], t_ov.ev_start, 0), "cudaStreamWaitEvent");
  // the two sweeps share the 2 CTA slots per SM: `split` eighths of them go to sweep 1 (HBM bound, quick), the rest to
  // sweep 2 (NVLink bound).  CFB200_P2P_SPLIT=1..7, default 4 (one slot per SM each).
  static const int split = [] {
    const int v = getenv("CFB200_P2P_SPLIT") ? atoi(getenv("CFB200_P2P_SPLIT")) : 4;
    return v >= 1 && v <= 7 ? v : 4;
  }();
  const int slots = 2 * sm_count(), cap1 = slots * split / 8, cap2 = slots - cap1;
  for (int ch = 0; ch < nch && ok; ++ch) {
    const long long m0 = (long long)ch * per, lc = lot - m0 < per ? lot - m0 : per;
    PeerOut sub = *po;
    sub.base += m0 * po->row_jump;
    Tw2 t2;
    if (tw2) {
      t2 = *tw2;
      t2.off += m0;
    }
    const FourStepIO io = {c + m0 * jump, inc, jump, c + m0 * jump, inc, jump};
    cpx *buf = scr + (size_t)(ch & 1) * per * n;
    // sweep 1 (HBM bound) on stream 0; its buffer was last read by sweep 2 of chunk ch - 2
    set_current_stream(t_ov.st[0]);
    set_tile_cta_cap(cap1);
    if (ch >= 2) ok = cuda_ok(cudaStreamWaitEvent(t_ov.st[0], t_ov.ev2[ch - 2], 0), "cudaStreamWaitEvent");
    ok = ok && run_c2c_pow2_four_step_chunk(n, a1, a2, lc, io, dir, scale, &sub, tw2 ? &t2 : nullptr, 1, buf) &&
         cuda_ok(cudaEventRecord(t_ov.ev1[ch], t_ov.st[0]), "cudaEventRecord");
    // sweep 2 (NVLink bound) on stream 1
    set_current_stream(t_ov.st[1]);
    set_tile_cta_cap(cap2);
    ok = ok && cuda_ok(cudaStreamWaitEvent(t_ov.st[1], t_ov.ev1[ch], 0), "cudaStreamWaitEvent") &&
         run_c2c_pow2_four_step_chunk(n, a1, a2, lc, io, dir, scale, &sub, tw2 ? &t2 : nullptr, 2, buf) &&
         cuda_ok(cudaEventRecord(t_ov.ev2[ch], t_ov.st[1]), "cudaEventRecord");
  }
  set_tile_cta_cap(0);
  set_current_stream(user);
  for (int k = 0; k < 2; ++k)
    ok = cuda_ok(cudaEventRecord(t_ov.ev_end[k], t_ov.st[k]), "cudaEventRecord") &&
         cuda_ok(cudaStreamWaitEvent(user, t_ov.ev_end[k], 0), "cudaStreamWaitEvent") && ok;
  return ok;
}

static bool run_c2c_pow2_four_step(int n, int a1, int a2, long long lot, long long inc, long long jump, int dir, cpx *c,
                                   double scale, const PeerOut *po = nullptr, const Tw2 *tw2 = nullptr) {
  if (po && po->npeers > 1) {
    static const int chunks = getenv("CFB200_P2P_CHUNKS") ? atoi(getenv("CFB200_P2P_CHUNKS")) : 8;
    if (chunks > 1 && lot >= 512 && run_c2c_pow2_four_step_pipelined(n, a1, a2, lot, inc, jump, dir, c, scale, po, tw2, chunks))
      return true;
  }
  static const long long chunk_mb = getenv("CFB200_FS_CHUNK_MB") ? atoll(getenv("CFB200_FS_CHUNK_MB")) : 0;
  long long per = chunk_mb > 0 ? (chunk_mb << 20) / ((long long)n * (long long)sizeof(cpx)) : lot;
  per -= per % 64;  // whole tiles of the row lengths the sweeps use
  if (per < 64 || per >= lot) {
    const FourStepIO io = {c, inc, jump, c, inc, jump};
    return run_c2c_pow2_four_step_chunk(n, a1, a2, lot, io, dir, scale, po, tw2);
  }
  for (long long m0 = 0; m0 < lot; m0 += per) {
    const long long lc = lot - m0 < per ? lot - m0 : per;
    PeerOut sub;
    if (po) {
      sub = *po;
      sub.base += m0 * po->row_jump;
    }
    Tw2 t2;
    if (tw2) {
      t2 = *tw2;
      t2.off += m0;
    }
    const FourStepIO io = {c + m0 * jump, inc, jump, c + m0 * jump, inc, jump};
    if (!run_c2c_pow2_four_step_chunk(n, a1, a2, lc, io, dir, scale, po ? &sub : nullptr, tw2 ? &t2 : nullptr)) return false;
  }
  return true;
}

static void set_tw2(TileParams &P, const Tw2 *tw2, int seq_lo, int n1) {
  if (!tw2 || !tw2->rp) return;
  P.tw2 = tw2->rp->d_w;
  P.tw2_shift = tw2->rp->shift;
  P.tw2_mask = tw2->n - 1;
  P.tw2_off = tw2->off;
  P.tw2_seq_lo = seq_lo;
  P.tw2_n1 = n1;
}

/* sweep = 1, 2: only that sweep (the pipelined sharded path issues them on different streams); 0: both.  scr_in: the
 * intermediate array to use (nullptr: the thread's scratch slot 0) */
static bool run_c2c_pow2_four_step_chunk(int n, int a1, int a2, long long lot, const FourStepIO &io, int dir, double scale,
                                         const PeerOut *po, const Tw2 *tw2, int sweep, cpx *scr_in) {
  const long long inc = io.inc, jump = io.jump;
  const cpx *c = io.cin;
  const int n1 = 1 << a1, n2 = 1 << a2;
  const RootPlan *rp = get_root_plan(n);
  if (!rp) return false;
  cpx *scr = scr_in ? scr_in : (cpx *)scratch_get(0, (size_t)lot * n * sizeof(cpx));
  if (!scr) return false;
  const long long ainc = inc < 0 ? -inc : inc, ajump = jump < 0 ? -jump : jump;
  const bool batch_fast = ajump < ainc && lot > 1;
  TileParams P;
  memset(&P, 0, sizeof(P));
  // step 1: rows (m, j2), transform over j1 (length n1), twiddle W_n^(j2*k1) on store
  P.in = c;
  P.out = scr;
  P.lot = lot * n2;
  P.scale = 1.0;
  P.fs = rp->d_w;
  P.fs_shift = rp->shift;
  P.fs_count = (1 << rp->shift) + (n + (1 << rp->shift) - 1) / (1 << rp->shift);
  P.fs_nmask = n - 1;
  if (!batch_fast) {  // row g = m*n2 + j2 (j2 fast); scratch S[m][k1][j2]
    P.ain = make_addr((long long)n2 * inc, inc, jump, n2);
    P.aout = make_addr(n2, 1, n, n2);
    P.fs_from_hi = 0;
  } else {  // row g = j2*lot + m (m fast); scratch S[k1][j2][m]
    P.ain = make_addr((long long)n2 * inc, jump, inc, lot);
    P.aout = make_addr((long long)n2 * lot, 1, lot, lot);
    P.fs_from_hi = 1;
  }
  P.in_staged = 0;
  if (sweep != 2 && !pow2_tile_launch(a1, dir, P)) return false;
  if (sweep == 1) return true;
  // step 2: rows (m, k1), transform over j2 (length n2), output element k2 goes to index k1 + n1*k2
  const long long oinc = io.inc_out, ojump = io.jump_out;
  P.in = scr;
  P.out = io.cout;
  P.lot = lot * n1;
  P.scale = scale;
  P.fs = nullptr;
  P.fs_count = 0;
  // destination strides of (sequence m, output index k): on peers they come from the PeerOut description
  const long long d_inc = (po && po->npeers > 0) ? po->elem_inc : oinc, d_jump = (po && po->npeers > 0) ? po->row_jump : ojump;
  const long long ad_inc = d_inc < 0 ? -d_inc : d_inc, ad_jump = d_jump < 0 ? -d_jump : d_jump;
  // out of place with the SEQUENCE index as the contiguous axis of the destination (the transposed write of a long 1-D
  // transform): enumerate the rows sequence-fastest, so that a tile's stores are runs along m instead of stride-d_inc
  const bool seq_fast_out = !batch_fast && ad_jump < ad_inc && lot >= 32;
  if (!batch_fast && !seq_fast_out) {  // row g = m*n1 + k1: scratch rows are contiguous along j2 -> staged load
    P.ain = make_addr(1, n2, n, n1);
    P.aout = make_addr((long long)n1 * oinc, oinc, ojump, n1);
    P.in_staged = 1;
    set_tw2(P, tw2, 0, n1);
  } else if (!batch_fast) {  // row g = k1*lot + m: scratch row (m, k1) at m*n + k1*n2, still contiguous along j2
    P.ain = make_addr(1, n, n2, lot);
    P.aout = make_addr((long long)n1 * oinc, ojump, oinc, lot);
    P.in_staged = 1;
    set_tw2(P, tw2, 1, n1);
  } else {  // row g = k1*lot + m
    P.ain = make_addr(lot, 1, (long long)n2 * lot, lot);
    P.aout = make_addr((long long)n1 * oinc, ojump, oinc, lot);
    P.in_staged = 0;
    set_tw2(P, tw2, 1, n1);
  }
  if (po && po->npeers > 0) {
    // output element k = k1 + n1*e of sequence m lives on peer k / chunk at base + m*row_jump + (k % chunk)*elem_inc
    if (po->chunk % n1 || po->npeers > 16) {
      set_error("sharded transform: slab of %lld elements is not a multiple of %d", po->chunk, n1);
      return false;
    }
    int sh = 0;
    while ((1LL << sh) < po->chunk / n1) ++sh;
    P.npeers = po->npeers;
    P.peer_shift = sh;
    P.out_base = po->base;
    for (int i = 0; i < po->npeers; ++i) P.peers[i] = po->peers[i];
    if (!batch_fast && !seq_fast_out) P.aout = make_addr((long long)n1 * po->elem_inc, po->elem_inc, po->row_jump, n1);
    else P.aout = make_addr((long long)n1 * po->elem_inc, po->row_jump, po->elem_inc, lot);
  }
  return pow2_tile_launch(a2, dir, P);
}

/* one sweep of the tile kernel: `lot` sequences of length 2^a (a in 6..10), any in/out layout, optional second twiddle */
static bool run_c2c_pow2_single_sweep(int a, long long lot, const FourStepIO &io, int dir, double scale, const Tw2 *tw2) {
  if (lot > 2147483647LL) {
    set_error("batch too large");
    return false;
  }
  TileParams P;
  memset(&P, 0, sizeof(P));
  P.in = io.cin;
  P.out = io.cout;
  P.lot = lot;
  P.scale = scale;
  P.ain = make_addr(io.inc, io.jump, 0, lot);
  P.aout = make_addr(io.inc_out, io.jump_out, 0, lot);
  P.in_staged = io.inc == 1 ? 1 : 0;
  set_tw2(P, tw2, 1, 1);  // rows are (hi = 0, lo = sequence): output index b = element
  return pow2_tile_launch(a, dir, P);
}

/* `lot` sequences of length n = 2^a, a in {6..10} (one sweep) or {12..20} (four-step), out of place */
static bool run_c2c_pow2_sub(int a, long long lot, const FourStepIO &io, int dir, double scale, const Tw2 *tw2) {
  if (a <= pow2_tile_max_log2()) return run_c2c_pow2_single_sweep(a, lot, io, dir, scale, tw2);
  const int a1 = a / 2, a2 = a - a1;
  return run_c2c_pow2_four_step_chunk(1 << a, a1, a2, lot, io, dir, scale, nullptr, tw2);
}

/* split of a long power-of-two length 2^a into 2^aL * 2^aM with both parts served by run_c2c_pow2_sub */
static bool long_pow2_split(int a, int *aL, int *aM) {
  if (a < 21 || a > 30) return false;
  *aL = a >= 24 ? a / 2 : (a == 21 ? 9 : 10);
  *aM = a - *aL;
  return true;
}

/* Power-of-two lengths beyond the four-step range (2^21 .. 2^30): N = L * Mm, x[i + L j] (i < L, j < Mm)
 *   A. for every i: transform over j (length Mm, stride L), results times W_N^(i b)   -> T[i + L b]
 *   B. for every b: transform over i (length L, contiguous)                            -> X[Mm a + b]
 * Each part is itself a four-step pair of sweeps (or one sweep when short), so a 2^28-point transform is four sweeps.
 * The reference does any N in core with log(N) sweeps (c1fm1f_, cfftpack/fftpack.c:2041-2141). */
static bool run_c2c_long_pow2(int a, long long lot, long long inc, long long jump, int dir, cpx *c, double scale) {
  int aL, aM;
  if (!long_pow2_split(a, &aL, &aM)) {
    set_error("power-of-two length 2^%d is outside the supported range (<= 2^30)", a);
    return false;
  }
  const long long N = 1LL << a, L = 1LL << aL, Mm = 1LL << aM;
  const RootPlan *rp = get_root_plan((int)N);
  if (!rp) return false;
  cpx *T = (cpx *)scratch_get(8, (size_t)N * sizeof(cpx));
  if (!T) return false;
  Tw2 tw2;
  tw2.rp = rp;
  tw2.n = N;
  for (long long m = 0; m < lot; ++m) {
    cpx *x = c + m * jump;
    const FourStepIO ioA = {x, L * inc, inc, T, L, 1};  // sequences i (jump = inc), elements j (stride L inc) -> T[i + L b]
    if (!run_c2c_pow2_sub(aM, L, ioA, dir, 1.0, &tw2)) return false;
    const FourStepIO ioB = {T, 1, L, x, Mm * inc, inc};  // sequences b (jump L), elements i -> x[(Mm a + b) inc]
    if (!run_c2c_pow2_sub(aL, Mm, ioB, dir, scale, nullptr)) return false;
  }
  return true;
}

/* lengths with a prime factor beyond the four-step split: chirp-z through power-of-two transforms */
static bool run_c2c_bluestein(int n, long long lot, long long inc, long long jump, int dir, cpx *c, double scale) {
  const ChirpPlan *cp = get_chirp_plan(n);
  if (!cp) return false;
  const int L = cp->L;
  cpx *z = (cpx *)scratch_get(2, (size_t)lot * L * sizeof(cpx));
  if (!z) return false;
  ChirpParams P;
  memset(&P, 0, sizeof(P));
  P.n = n;
  P.L = L;
  P.dir = dir;
  P.lot = lot;
  P.a = make_addr(inc, jump, 0, 1LL << 30);
  P.user = c;
  P.z = z;
  P.chirp = cp->d_chirp;
  P.bhat = cp->d_bhat;
  P.scale = scale / (double)L;
  cudaStream_t st = current_stream();
  const unsigned gx = (unsigned)((L + 255) / 256 < 256 ? (L + 255) / 256 : 256);
  const long long YMAX = 65535;
  int launches = 0;
  for (P.row0 = 0; P.row0 < lot; P.row0 += YMAX, ++launches)
    CFB_LAUNCH(chirp_pre_kernel, dim3(gx, (unsigned)(lot - P.row0 < YMAX ? lot - P.row0 : YMAX)), 256, 0, st, P);
  count_launch(launches);
  if (!cuda_ok(cudaGetLastError(), "chirp_pre_kernel")) return false;
  if (!run_c2c_scaled(L, lot, 1, L, -1, z, 1.0)) return false;
  launches = 0;
  for (P.row0 = 0; P.row0 < lot; P.row0 += YMAX, ++launches)
    CFB_LAUNCH(chirp_mul_kernel, dim3(gx, (unsigned)(lot - P.row0 < YMAX ? lot - P.row0 : YMAX)), 256, 0, st, P);
  count_launch(launches);
  if (!run_c2c_scaled(L, lot, 1, L, +1, z, 1.0)) return false;
  launches = 0;
  for (P.row0 = 0; P.row0 < lot; P.row0 += YMAX, ++launches)
    CFB_LAUNCH(chirp_post_kernel, dim3(gx, (unsigned)(lot - P.row0 < YMAX ? lot - P.row0 : YMAX)), 256, 0, st, P);
  count_launch(launches);
  return cuda_ok(cudaGetLastError(), "chirp kernels");
}

bool run_c2c(int n, long long lot, long long inc, long long jump, int dir, void *c) {
  return run_c2c_scaled(n, lot, inc, jump, dir, c, dir < 0 ? 1.0 / (double)n : 1.0);
}

bool run_c2c_scaled(int n, long long lot, long long inc, long long jump, int dir, void *c, double scale) {
  if (n <= 1 || lot <= 0) return true;
  const int aligned = (((uintptr_t)c) & 15) == 0;
  if (pow2_c2c_supported(n, inc, jump, aligned)) return pow2_c2c_launch(n, lot, jump, dir, (cpx *)c, scale);
  if (r10_supported(n) && inc == 1 && jump >= n && aligned) return r10_c2c_launch(n, lot, jump, dir, (cpx *)c, scale);
  EngineParams P;
  memset(&P, 0, sizeof(P));
  P.kind = K_C2C;
  P.dir = dir;
  P.n = n;
  P.aligned16 = aligned;
  // batch-contiguous layouts need >= 8 rows side by side in one CTA for full 128-byte runs; when a sequence is too
  // long for that, the four-step split (short sub-transforms, many rows per CTA) is the faster route
  const long long ainc0 = inc < 0 ? -inc : inc, ajump0 = jump < 0 ? -jump : jump;
  const bool want_rows = ajump0 < ainc0 && lot >= 8;
  const bool split_for_rows = want_rows && (SMEM_MAX - 64) / ((size_t)(n | 1) * 32 + 32) < 8 && n >= 256 &&
                              four_step_split(n, engine_max_c2c()) > 1;
  if (n <= engine_max_c2c() && !split_for_rows) {
    const CorePlan *cp = get_core_plan(n);
    if (!cp) return false;
    fill_passes(P, cp);
    P.lot = lot;
    P.ain = P.aout = make_addr(inc, jump, 0, 1LL << 30);
    P.in = c;
    P.out = c;
    P.scale = scale;
    return launch_engine(P);
  }
  /* four-step: x[j1*n2 + j2] -> (FFT over j1) * W_n^{j2 k1} -> scratch[k1*n2 + j2] -> (FFT over j2) -> X[k1 + n1 k2] */
  if (aligned && (n & (n - 1)) == 0) {
    int a = 0;
    while ((1 << a) < n) ++a;
    const int a1 = a / 2, a2 = a - a1;
    if (a1 >= pow2_tile_min_log2() && a2 <= pow2_tile_max_log2()) return run_c2c_pow2_four_step(n, a1, a2, lot, inc, jump, dir, (cpx *)c, scale);
    if (a > 2 * pow2_tile_max_log2()) return run_c2c_long_pow2(a, lot, inc, jump, dir, (cpx *)c, scale);
  }
  const int n1 = four_step_split(n, engine_max_c2c());
  if (n1 <= 1) return run_c2c_bluestein(n, lot, inc, jump, dir, (cpx *)c, scale);  // a prime factor too large to split
  const int n2 = n / n1;
  if (lot * (long long)n2 > 2147483647LL * 16 || lot * (long long)n1 > 2147483647LL * 16) {
    set_error("batch too large");
    return false;
  }
  const CorePlan *p1 = get_core_plan(n1), *p2 = get_core_plan(n2);
  const RootPlan *rp = get_root_plan(n);
  if (!p1 || !p2 || !rp) return false;
  cpx *scr = (cpx *)scratch_get(0, (size_t)lot * n * sizeof(cpx));
  if (!scr) return false;
  const long long ainc = inc < 0 ? -inc : inc, ajump = jump < 0 ? -jump : jump;
  const bool batch_fast = ajump < ainc && lot > 1;  // interleaved layouts: the batch index is the contiguous axis
  P.fs_tw = rp->d_w;
  P.fs_n = n;
  P.fs_shift = rp->shift;
  {
    const int entries = (1 << rp->shift) + (n + (1 << rp->shift) - 1) / (1 << rp->shift);
    P.fs_smem = entries <= CFB_FS_SMEM_MAX ? entries : 0;
  }
  // step 1: rows (m, j2), transform over j1, twiddle W_n^(j2*k1) on store
  fill_passes(P, p1);
  P.n = n1;
  P.lot = lot * n2;
  if (!batch_fast) {  // row g = m*n2 + j2 (j2 fast); scratch S[m][k1][j2]
    P.ain = make_addr((long long)n2 * inc, inc, jump, n2);
    P.aout = make_addr(n2, 1, n, n2);
    P.fs_from_hi = 0;
  } else {  // row g = j2*lot + m (m fast); scratch S[k1][j2][m]
    P.ain = make_addr((long long)n2 * inc, jump, inc, lot);
    P.aout = make_addr((long long)n2 * lot, 1, lot, lot);
    P.fs_from_hi = 1;
  }
  P.in = c;
  P.out = scr;
  P.scale = 1.0;
  if (!launch_engine(P)) return false;
  // step 2: rows (m, k1), transform over j2, output element k2 goes to index k1 + n1*k2
  fill_passes(P, p2);
  P.n = n2;
  P.lot = lot * n1;
  if (!batch_fast) {  // row g = m*n1 + k1
    P.ain = make_addr(1, n2, n, n1);
    P.aout = make_addr((long long)n1 * inc, inc, jump, n1);
  } else {  // row g = k1*lot + m
    P.ain = make_addr(lot, 1, (long long)n2 * lot, lot);
    P.aout = make_addr((long long)n1 * inc, jump, inc, lot);
  }
  P.in = scr;
  P.out = c;
  P.scale = scale;
  P.fs_tw = nullptr;
  P.fs_n = 0;
  P.fs_smem = 0;
  return launch_engine(P);
}

/* One phase of the sharded 2-D transform, fused with its transpose: the local length-n transforms of
 * this rank's slab are computed and their results stored directly into the slabs of the GPUs that own them.
 *   phase 1: src = my column slab C[m_loc][l] (sequences contiguous);  result element (i, j) -> rank i / l_loc,
 *            D_r[j * l_loc + i % l_loc]                     (D = row slab, column-major (l_loc, m))
 *   phase 2: src = my row slab D[m][l_loc] (sequences along m, stride l_loc); result element (i, j) -> rank j / m_loc,
 *            C_r[(j % m_loc) * l + i]
 * The forward direction scales each phase by 1/n like cfftmf_ (total 1/(l m), as cfft2f_). */
bool run_c2c_2d_sharded_phase(int phase, int dir, int l, int m, int rank, int nranks, void *src, void *const *peers) {
  const int l_loc = l / nranks, m_loc = m / nranks;
  const int n = phase == 1 ? l : m;
  if ((n & (n - 1)) != 0 || l % nranks || m % nranks) {
    set_error("sharded 2-D transform needs power-of-two l, m divisible by the number of ranks");
    return false;
  }
  int a = 0;
  while ((1 << a) < n) ++a;
  const int a1 = a / 2, a2 = a - a1;
  if (a1 < pow2_tile_min_log2() || a2 > pow2_tile_max_log2()) {
    set_error("sharded 2-D transform: length %d outside the fused path (2^12 .. 2^20)", n);
    return false;
  }
  PeerOut po;
  po.npeers = nranks;
  po.peers = (cpx *const *)peers;
  if (phase == 1) {
    po.chunk = l_loc;
    po.base = (long long)rank * m_loc * l_loc;
    po.row_jump = l_loc;  // sequence j_loc -> column (rank*m_loc + j_loc) of D
    po.elem_inc = 1;
    return run_c2c_pow2_four_step(l, a1, a2, m_loc, 1, l, dir, (cpx *)src, dir < 0 ? 1.0 / (double)l : 1.0, &po);
  }
  po.chunk = m_loc;
  po.base = (long long)rank * l_loc;
  po.row_jump = 1;    // sequence i_loc -> row rank*l_loc + i_loc of C
  po.elem_inc = l;    // output index j_loc -> column j_loc of C (pitch l)
  return run_c2c_pow2_four_step(m, a1, a2, l_loc, l_loc, 1, dir, (cpx *)src, dir < 0 ? 1.0 / (double)m : 1.0, &po);
}

/* ---- very long 1-D transforms across GPUs: N = L * Mm, the array distributed in natural order
 * (rank r owns x[r N/G .. (r+1) N/G) = the columns j of x[i + L j] in its slab).  Three phases, each followed by a
 * barrier among the ranks (the caller's job, as for the 2-D transform):
 *   phase 0: transpose only -- my column slab C[m_loc][L] goes to the row slabs D_r[j][i_loc] of the GPUs owning i;
 *   phase 1: on my row slab D: length-Mm transforms over j, results times W_N^(i b), stored into the column slabs
 *            C_r[(b % m_loc) L + i] of the GPUs owning b (the exchange rides on the last pass, as in the 2-D case);
 *   phase 2: on my column slab C: length-L transforms over i, result X[Mm a + b] stored into natural order on the GPU
 *            owning a: D_r[(a % l_loc) Mm + b].
 * The result therefore ends in the D buffers, in natural order.  Forward scales by 1/N in total. */
struct TransposeParams {
  const cpx *src;
  cpx *peers[16];
  long long L, l_loc, m_loc, base;  // base = rank * m_loc * l_loc
  int shift;                        // log2(l_loc)
};
__global__ void __launch_bounds__(256) p2p_transpose_kernel(const TransposeParams P) {
  const long long total = P.m_loc * P.L, stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const long long j = idx / P.L, i = idx - j * P.L;
    P.peers[i >> P.shift][P.base + j * P.l_loc + (i & (P.l_loc - 1))] = P.src[idx];
  }
}

bool run_c2c_1d_sharded_phase(int phase, int dir, int log2n, int rank, int nranks, void *src, void *const *peers) {
  const int aL = log2n / 2, aM = log2n - aL;
  if (log2n < 24 || log2n > 30 || nranks < 1 || nranks > 16 || (nranks & (nranks - 1)) || aM > 2 * pow2_tile_max_log2()) {
    set_error("sharded 1-D transform: needs N = 2^24 .. 2^30 and a power-of-two number of ranks <= 16");
    return false;
  }
  const long long N = 1LL << log2n, L = 1LL << aL, Mm = 1LL << aM, l_loc = L / nranks, m_loc = Mm / nranks;
  if (phase == 0) {
    TransposeParams P;
    memset(&P, 0, sizeof(P));
    P.src = (const cpx *)src;
    for (int i = 0; i < nranks; ++i) P.peers[i] = (cpx *)peers[i];
    P.L = L;
    P.l_loc = l_loc;
    P.m_loc = m_loc;
    P.base = (long long)rank * m_loc * l_loc;
    while ((1LL << P.shift) < l_loc) ++P.shift;
    CFB_LAUNCH(p2p_transpose_kernel, (unsigned)(8 * sm_count()), 256, 0, current_stream(), P);
    count_launch();
    return cuda_ok(cudaGetLastError(), "p2p_transpose_kernel launch");
  }
  PeerOut po;
  po.npeers = nranks;
  po.peers = (cpx *const *)peers;
  if (phase == 1) {
    const RootPlan *rp = get_root_plan((int)N);
    if (!rp) return false;
    Tw2 tw2;
    tw2.rp = rp;
    tw2.n = N;
    tw2.off = (long long)rank * l_loc;
    po.chunk = m_loc;
    po.base = (long long)rank * l_loc;
    po.row_jump = 1;
    po.elem_inc = L;
    return run_c2c_pow2_four_step((int)Mm, aM / 2, aM - aM / 2, l_loc, l_loc, 1, dir, (cpx *)src, dir < 0 ? 1.0 / (double)Mm : 1.0, &po,
                                  &tw2);
  }
  po.chunk = l_loc;
  po.base = (long long)rank * m_loc;
  po.row_jump = 1;
  po.elem_inc = Mm;
  return run_c2c_pow2_four_step((int)L, aL / 2, aL - aL / 2, m_loc, 1, L, dir, (cpx *)src, dir < 0 ? 1.0 / (double)L : 1.0, &po);
}

bool run_c2c_2d(int ldim, int l, int m, int dir, void *c) {
  // cfft2f_ order (fftpack.c:2408-2426): lines along the second index first, then along the first
  if (!run_c2c(m, l, ldim, 1, dir, c)) return false;
  return run_c2c(l, m, 1, ldim, dir, c);
}

/* ---- rfft2f_/rfft2b_ (fftpack.c:13282, :13113).  Columns become half-complex vectors along i; the (Re, Im) row pairs
 * f = 1..(l-1)/2 are gathered into a compact complex array W(f-1, j) -- the conversion between FFTPACK's (2/N cos, 2/N
 * sin) and plain (Re, Im)/N, a factor 1/2 and a sign, rides on that copy -- transformed along j as one batched complex
 * transform, and scattered back; rows 0 and (l even) l-1 are real along j.  The reference instead copies the whole
 * array into `work` and sweeps r three more times for the conversions (:13383-13395, r2w_/w2r_). ---- */
struct Real2dParams {
  double *r;
  cpx *w;
  long long ldim;
  int lotc, m;
  double sre, sim;
};
__global__ void __launch_bounds__(256) rfft2_gather_kernel(const Real2dParams P) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;  // pair index f+1
  if (f >= P.lotc) return;
  for (int j = blockIdx.y; j < P.m; j += gridDim.y) {
    const double *col = P.r + (long long)j * P.ldim + 2 * f + 1;
    P.w[(long long)j * P.lotc + f] = make_double2(col[0] * P.sre, col[1] * P.sim);
  }
}
__global__ void __launch_bounds__(256) rfft2_scatter_kernel(const Real2dParams P) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= P.lotc) return;
  for (int j = blockIdx.y; j < P.m; j += gridDim.y) {
    const cpx v = P.w[(long long)j * P.lotc + f];
    double *col = P.r + (long long)j * P.ldim + 2 * f + 1;
    col[0] = v.x * P.sre;
    col[1] = v.y * P.sim;
  }
}
/* one strided line of length len: x[k] *= s for 0 < k < 2*((len+1)/2) - 1, sign flipped at even k >= 2 */
__global__ void __launch_bounds__(256) rfft2_line_kernel(double *x, long long stride, int len, double s) {
  const int top = 2 * ((len + 1) / 2) - 1;
  for (int k = 1 + blockIdx.x * blockDim.x + threadIdx.x; k < top; k += gridDim.x * blockDim.x)
    x[(long long)k * stride] *= (k & 1) ? s : -s;
}

bool run_real_2d(int ldim, int l, int m, int dir, double *r) {
  cudaStream_t st = current_stream();
  const int lotc = (l + 1) / 2 - 1;
  auto line = [&](double *row, double s) {
    if (2 * ((m + 1) / 2) - 1 <= 1) return true;  // nothing between the mean and the Nyquist term
    const unsigned g = (unsigned)((m + 255) / 256 < 64 ? (m + 255) / 256 : 64);
    CFB_LAUNCH(rfft2_line_kernel, g, 256, 0, st, row, (long long)ldim, m, s);
    count_launch();
    return cuda_ok(cudaGetLastError(), "rfft2_line_kernel");
  };
  auto rows_along_j = [&](double *row) {  // one real transform of length m, stride ldim
    if (dir < 0) return (m == 1 || run_real(K_RFFT, m, 1, ldim, 1, -1, row)) && line(row, 0.5);
    return line(row, 2.0) && (m == 1 || run_real(K_RFFT, m, 1, ldim, 1, +1, row));
  };
  Real2dParams P;
  memset(&P, 0, sizeof(P));
  P.r = r;
  P.ldim = ldim;
  P.lotc = lotc;
  P.m = m;
  const dim3 grid((unsigned)((lotc + 255) / 256), (unsigned)(m < 4096 ? m : 4096));
  if (dir < 0 && l > 1 && !run_real(K_RFFT, l, m, 1, ldim, -1, r)) return false;
  if (!rows_along_j(r)) return false;
  if (l % 2 == 0 && !rows_along_j(r + (l - 1))) return false;
  if (lotc > 0) {
    P.w = (cpx *)scratch_get(7, (size_t)lotc * m * sizeof(cpx));
    if (!P.w) return false;
    P.sre = dir < 0 ? 0.5 : 1.0;
    P.sim = dir < 0 ? -0.5 : 1.0;
    CFB_LAUNCH(rfft2_gather_kernel, grid, 256, 0, st, P);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "rfft2_gather_kernel")) return false;
    if (m > 1 && !run_c2c(m, lotc, lotc, 1, dir, P.w)) return false;
    P.sre = dir < 0 ? 1.0 : 2.0;
    P.sim = dir < 0 ? 1.0 : -2.0;
    CFB_LAUNCH(rfft2_scatter_kernel, grid, 256, 0, st, P);
    count_launch();
    if (!cuda_ok(cudaGetLastError(), "rfft2_scatter_kernel")) return false;
  }
  if (dir > 0 && l > 1 && !run_real(K_RFFT, l, m, 1, ldim, +1, r)) return false;
  return true;
}

/* real-family sequences too long for one CTA: global-memory pipeline around the long complex transform */
static int long_real_threshold() {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("CFB200_LONG_REAL_MIN");  // tests lower it to exercise this path with short sequences
    v = e ? atoi(e) : engine_max_real() + 1;
  }
  return v;
}

static bool run_real_long(int kind, int n, int M, long long lot, long long inc, long long jump, int dir, double *x) {
  const bool fwd_core = !((kind == K_RFFT || kind == K_COSQ || kind == K_SINQ) && dir > 0);
  LongRealParams P;
  memset(&P, 0, sizeof(P));
  P.kind = kind;
  P.dir = dir;
  P.n = n;
  P.M = M;
  P.lot = lot;
  P.ld = (long long)(M > n ? M : n) + 2;
  P.a = make_addr(inc, jump, 0, 1LL << 30);
  P.user = x;
  const long long pairs = (lot + 1) / 2, rows2 = 2 * pairs;
  char *base = (char *)scratch_get(3, (size_t)rows2 * P.ld * 8 * 2 + (size_t)pairs * P.ld * 16 + (size_t)rows2 * 8 + 64);
  if (!base) return false;
  P.xs = (double *)base;
  P.ys = P.xs + rows2 * P.ld;
  P.z = (cpx *)(P.ys + rows2 * P.ld);
  P.dsum = (double *)(P.z + pairs * P.ld);
  if (kind != K_RFFT) {
    const TrigPlan *tp = get_trig_plan(kind == K_SINQ ? K_COSQ : kind, n);
    if (!tp) return false;
    P.trig = tp->d_trig;
  }
  if (lot > 2147483647LL) {
    set_error("long real path: batch too large");
    return false;
  }
  cudaStream_t st = current_stream();
  const unsigned gx = (unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024);
  const unsigned gq = (unsigned)((M / 2 + 256) / 256 < 1024 ? (M / 2 + 256) / 256 : 1024);
  const unsigned gm = (unsigned)((M + 255) / 256 < 1024 ? (M + 255) / 256 : 1024);
  const long long YMAX = 65535;
  int launches = 0;
  for (P.row0 = 0; P.row0 < lot; P.row0 += YMAX, ++launches)
    CFB_LAUNCH(long_gather_kernel, dim3(gx, (unsigned)(lot - P.row0 < YMAX ? lot - P.row0 : YMAX)), 256, 0, st, P);
  CFB_LAUNCH(long_pre_kernel, (unsigned)lot, 32, 0, st, P, fwd_core ? 1 : 0);
  ++launches;
  if (!fwd_core)
    for (P.row0 = 0; P.row0 < pairs; P.row0 += YMAX, ++launches)
      CFB_LAUNCH(long_build_kernel, dim3(gq, (unsigned)(pairs - P.row0 < YMAX ? pairs - P.row0 : YMAX)), 256, 0, st, P);
  count_launch(launches);
  if (!cuda_ok(cudaGetLastError(), "long real pre kernels")) return false;
  // the real core wants the UNSCALED complex transform (split_pair applies 1/M itself)
  if (!run_c2c_scaled(M, pairs, 1, P.ld, fwd_core ? -1 : +1, P.z, 1.0)) return false;
  launches = 0;
  for (P.row0 = 0; P.row0 < pairs; P.row0 += YMAX, ++launches)
    CFB_LAUNCH(long_split_kernel, dim3(fwd_core ? gq : gm, (unsigned)(pairs - P.row0 < YMAX ? pairs - P.row0 : YMAX)), 256, 0,
               st, P, fwd_core ? 1 : 0);
  CFB_LAUNCH(long_post_kernel, (unsigned)lot, 32, 0, st, P, fwd_core ? 1 : 0);
  ++launches;
  for (P.row0 = 0; P.row0 < lot; P.row0 += YMAX, ++launches)
    CFB_LAUNCH(long_scatter_kernel, dim3(gx, (unsigned)(lot - P.row0 < YMAX ? lot - P.row0 : YMAX)), 256, 0, st, P);
  count_launch(launches);
  return cuda_ok(cudaGetLastError(), "long real post kernels");
}

bool run_real(int kind, int n, long long lot, long long inc, long long jump, int dir, double *x) {
  if (n <= 1 || lot <= 0) return true;
  const bool tiny = (kind == K_COST && n <= 3) || (kind != K_RFFT && n == 2);
  if (tiny) {
    TinyParams tp;
    tp.kind = kind;
    tp.dir = dir;
    tp.n = n;
    tp.lot = lot;
    tp.a = make_addr(inc, jump, 0, 1LL << 30);
    tp.x = x;
    long long grid = (lot + 127) / 128;
    CFB_LAUNCH(tiny_kernel, (unsigned)grid, 128, 0, current_stream(), tp);
    count_launch();
    return cuda_ok(cudaGetLastError(), "tiny_kernel launch");
  }
  // A/B switch: CFB200_NO_MIX=1 the older kernels everywhere, =2 only for M = 1000 (radix10.cuh)
  static const int no_mix = getenv("CFB200_NO_MIX") ? atoi(getenv("CFB200_NO_MIX")) : 0;
  const int Mu = kind == K_COST ? n - 1 : kind == K_SINT ? n + 1 : n;
  if (no_mix != 1 && !(no_mix == 2 && Mu == 1000) && mix_supported(kind, n) && inc == 1 && jump == n && lot >= 2 &&
      (((uintptr_t)x) & 15) == 0) {
    // mixed-radix streaming kernel on the pairs of rows; a last odd row goes through the general engine
    const double *trig = nullptr;
    if (kind != K_RFFT) {
      const TrigPlan *tp = get_trig_plan(kind, n);
      if (!tp) return false;
      trig = tp->d_trig;
    }
    if (!mix_launch(kind, n, lot / 2, dir, x, trig)) return false;
    if (lot % 2 == 0) return true;
    x += (lot - 1) * jump;
    lot = 1;
  }
  if (kind == K_RFFT && pow2_r2c_supported(n, inc, jump, (((uintptr_t)x) & 15) == 0))
    return pow2_r2c_launch(n, lot, jump, dir, x);
  if (kind == K_RFFT && r10_supported(n) && inc == 1 && jump >= n && jump % 2 == 0 && (((uintptr_t)x) & 15) == 0)
    return r10_r2c_launch(n, lot, jump, dir, x);
  if (kind == K_COSQ && r10_supported(n) && inc == 1 && jump >= n && jump % 2 == 0 && (((uintptr_t)x) & 15) == 0) {
    const TrigPlan *tp = get_trig_plan(K_COSQ, n);
    if (!tp) return false;
    return r10_cosq_launch(n, lot, jump, dir, x, tp->d_trig);
  }
  if (kind == K_COST && n == 1001 && inc == 1 && jump == n && lot >= 2 && (((uintptr_t)x) & 15) == 0) {
    // radix-10 register kernel on the even part of the batch; a last odd row goes through the general engine
    const TrigPlan *tp = get_trig_plan(K_COST, n);
    if (!tp) return false;
    if (!r10_cost_launch(lot / 2, dir, x, tp->d_trig)) return false;
    if (lot % 2 == 0) return true;
    x += (lot - 1) * jump;
    lot = 1;
  }
  const int M = kind == K_COST ? n - 1 : kind == K_SINT ? n + 1 : n;
  if (M >= long_real_threshold() || n >= long_real_threshold()) return run_real_long(kind, n, M, lot, inc, jump, dir, x);
  EngineParams P;
  memset(&P, 0, sizeof(P));
  P.kind = kind;
  P.dir = dir;
  P.n = n;
  const CorePlan *cp = get_core_plan(M);
  if (!cp) return false;
  fill_passes(P, cp);
  if (kind != K_RFFT) {
    const TrigPlan *tp = get_trig_plan(kind == K_SINQ ? K_COSQ : kind, n);
    if (!tp) return false;
    P.trig = tp->d_trig;
  }
  P.lot = lot;
  P.ain = P.aout = make_addr(inc, jump, 0, 1LL << 30);
  P.in = x;
  P.out = x;
  P.scale = 1.0;
  return launch_engine(P);
}

}  // namespace cfb
