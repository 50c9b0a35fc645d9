#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json): batched cfftmf_ FP64, N=4096, lot=65536 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfftm|rfftm|cfft2]

One "step" = one forward pass of the hot path (cfftmf_, in place) over one batch of synthetic data.
  value     : algorithmic GB/s (one read + one write of the payload) of the whole job, data resident
              in HBM, timed on the device with CUDA events on the launching stream, max over ranks.
  e2e       : the same metric through the C ABI with HOST (pinned) arrays: H2D copy + transform + D2H copy timed.
  roofline  : achieved GB/s of the dominant kernel vs the measured HBM copy peak (MEASURED_PEAKS.json).
  configs   : (N = 1) every other BASELINE.json config on the same box in the same run -- C1 latency, C3 rfftmf/rfftmb,
              the C4 DCT/DST matrix of SURVEY 8(a) note 2, C5 cfft2f 16384^2 -- each with ms, GB/s, roofline fraction
              and its own cpu_baseline from the unmodified reference (oracle/_ref).
  cfft2     : (N > 1) BASELINE configs[4]: cfft2f 16384^2 sharded over the N GPUs as column slabs with the transposes
              fused into the FFT kernels (P2P stores over NVLink), strong scaling, checked in the same run against
              the CPU oracle on sampled rows and columns of a non-separable random matrix.
  cpu_baseline / --impl reference : the unmodified reference (oracle/_ref, compiled from /root/reference) or, if
              that is absent, the oracle port, lot-parallel over all host cores.
Multi-GPU: the lot axis is sharded, one process per GPU, no data-path collective (weak scaling: each GPU
transforms its own 65536 sequences).  Inputs (4 GiB per GPU) are far larger than L2 (126 MB), so no L2 flush.
"""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DEFAULT, LOT_DEFAULT = 4096, 65536
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent
NVLINK_GBS = 900.0         # NVLink 5 per direction per GPU


def algorithmic_bytes(fam, n, lot):
    return 2 * (16 if fam == "cfft" else 8) * n * lot


def underlying(fam, n):
    return {"cost": n - 1, "sint": n + 1}.get(fam, n)


def nominal_flops(fam, n, lot):
    m = underlying(fam, n)
    return (5.0 if fam == "cfft" else 2.5) * m * math.log2(m) * lot


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def workload_config(fam, n, lot):
    """identical for the b200 arm and the reference arm (the driver compares them)"""
    bytes_rank = algorithmic_bytes(fam, n, lot)
    return {"workload": f"{fam}mf N={n} lot={lot} per GPU, inc=1, jump=N, in place, forward (BASELINE configs[{1 if fam == 'cfft' else 2}])",
            "sharding": "by lot, one process per GPU, no collective",
            "l2": f"inputs {bytes_rank // 2 >> 20} MiB per GPU >> 126 MB L2, no flush needed"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "10"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def mark(self):
        """start of the timed region: earlier samples (warm-up, idle) are dropped"""
        self.t0 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t0 = getattr(self, "t0", 0.0)
        for ts, ln in self.lines:
            if ts < t0:
                continue
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/_ref) -- or the oracle port -- looped lot-parallel over the host cores.
# The only place bench.py executes anything under oracle/ besides the cfft2 parity check.
# ------------------------------------------------------------------------------------------------------------------
class CpuRef:
    def __init__(self):
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import fftlibs as fl
        self.fl = fl
        self.orc = fl.oracle()
        if self.orc is None:
            raise RuntimeError("oracle/liboracle.so missing: run python __graft_entry__.py")
        ref = fl.ref()
        self.kind = "reference" if ref is not None else "port"
        self.lib = ref if ref is not None else self.orc
        self.prefix = "" if ref is not None else "orc_"
        self.L = fl.Lib(self.lib, self.prefix)
        self.orc.orc_lot_parallel.restype = ctypes.c_double
        self.orc.orc_lot_parallel.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                              ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        self.threads = os.cpu_count() or 1

    def plan(self, fam, n, d="f"):
        ws, ier = self.L.init(fam, n)
        assert ier == 0, (fam, n, ier)
        fn = ctypes.cast(getattr(self.lib, f"{self.prefix}{fam}1{d}_"), ctypes.c_void_p)
        return ws, fn

    def passes(self, fam, n, lot, x, ws, fn, reps=1, threads=None):
        """seconds for `reps` lot-parallel passes of looped <fam>1f_ over x[lot][n] (in place)"""
        fl = self.fl
        return self.orc.orc_lot_parallel(fn, 1 if fam == "cfft" else 0, lot, n, fl.P(x), fl.P(ws), fl.lensav(fam, n),
                                         threads or self.threads, reps)

    def baseline(self, fam, n, d="f", lot=None, budget_s=3.0, shipped=False):
        """bounded lot-parallel sample -> dict in the cpu_baseline format (GB/s algorithmic)"""
        fl = self.fl
        lot = lot or 256 * self.threads
        ws, fn = self.plan(fam, n, d)
        x = fl.rand_input(fam, lot * n, 11)
        t1 = self.passes(fam, n, lot, x, ws, fn)  # warm + calibrate
        reps = max(1, min(64, int(budget_s / max(t1, 1e-3))))
        x = fl.rand_input(fam, lot * n, 11)  # repeated scaled-forward passes must not underflow into denormals
        per = self.passes(fam, n, lot, x, ws, fn, reps) / reps
        out = {"value": algorithmic_bytes(fam, n, lot) / per / 1e9, "unit": "GB/s", "cores": self.threads, "kind": self.kind,
               "sample": f"{lot} sequences x {reps} passes of looped {fam}1{d}_ N={n}, lot-parallel over {self.threads} "
                         f"threads ({per / lot * self.threads * 1e6:.1f} us per sequence per thread)",
               "gflops": nominal_flops(fam, n, lot) / per / 1e9, "seconds_per_pass": per, "sample_lot": lot}
        if shipped:  # the as-shipped batched routine on the contiguous layout, 1 thread, small lot (cache-hostile)
            lot_m = 64
            xm = fl.rand_input(fam, lot_m * n, 12)
            t0 = time.perf_counter()
            _, ier = self.L.runm(fam, d, lot_m, n, n, 1, xm)
            tm = time.perf_counter() - t0
            out["sample"] += f"; as-shipped {fam}m{d}_ lot={lot_m} 1 thread: {tm / lot_m * 1e6:.0f} us/sequence"
        return out


def run_reference(args, fam, n, lot, rank, world):
    """--impl reference: the reference's own CPU implementation of the path on the identical workload (full lot)."""
    if rank != 0:
        return
    cpu = CpuRef()
    fl = cpu.fl
    import numpy as np
    ws, fn = cpu.plan(fam, n)
    rng = np.random.default_rng(11)
    x = rng.uniform(-1, 1, lot * n * (2 if fam == "cfft" else 1))  # the whole lot, as on the GPU arm
    for _ in range(args.warmup):
        cpu.passes(fam, n, lot, x, ws, fn)
    t = 0.0
    for i in range(args.steps):
        if (i + args.warmup) % 64 == 63:  # in-place 1/N-scaled passes: refresh before the values turn denormal
            x[:] = rng.uniform(-1, 1, x.size)
        t += cpu.passes(fam, n, lot, x, ws, fn)
    per = t / max(args.steps, 1)
    val = algorithmic_bytes(fam, n, lot) / per / 1e9
    out = {"impl": "reference", "metric": f"batched {fam}mf FP64 N={n} algorithmic HBM GB/s", "value": val, "unit": "GB/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "gflops_5nlogn": nominal_flops(fam, n, lot) / per / 1e9,
           "config": workload_config(fam, n, lot),
           "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cpu.threads, "kind": cpu.kind,
                            "sample": f"each step = the full lot of {lot} sequences, looped {fam}1f_ lot-parallel over "
                                      f"{cpu.threads} threads (bit-identical to {fam}mf_, 9x faster on this layout)"},
           "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# GPU side helpers
# ------------------------------------------------------------------------------------------------------------------
def time_launches(torch, fn, reps, warm, stream=None):
    """average and median ms of `fn` over `reps` launches, CUDA events on the launching stream"""
    stream = stream or torch.cuda.current_stream()
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record(stream)
    for i in range(reps):
        fn()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ev[0].elapsed_time(ev[-1]) / reps, ts[len(ts) // 2]


def measure_family(torch, cb, fam, d, n, lot, reps=10, warm=3, seed=5):
    """device-resident <fam>m<d>_ on a contiguous batch (inc=1, jump=n): dict with ms / GB/s / fractions"""
    esz = 2 if fam == "cfft" else 1
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.rand(lot * n * esz, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    plan = cb.Plan(fam, n)

    def step():
        ier = plan.multi(d, x.data_ptr(), lot, n, 1, lot * n)
        if ier != 0:
            raise RuntimeError(f"{fam}m{d}_ n={n} ier={ier}: {cb.last_error()}")

    l0 = cb.launch_count()
    step()
    per_call = cb.launch_count() - l0
    avg, med = time_launches(torch, step, reps, warm)
    peak, _ = hbm_peak()
    by = algorithmic_bytes(fam, n, lot)
    gbs = by / (avg * 1e-3) / 1e9
    del x
    return {"routine": f"{fam}m{d}_", "n": n, "lot": lot, "underlying_fft_length": underlying(fam, n), "ms": avg, "median_ms": med,
            "gbs": gbs, "frac": gbs / peak, "frac_of_8TBps": gbs / 8000.0, "algorithmic_bytes": by,
            "gflops_nominal": nominal_flops(fam, n, lot) / (avg * 1e-3) / 1e9, "kernels_per_call": int(per_call)}


def measure_c1(torch, cb, cpu):
    """BASELINE configs[0]: one cfft1f_ + cfft1b_ round trip at N=1024 (latency-bound: microseconds, not GB/s)"""
    import numpy as np
    fl = cpu.fl
    n = 1024
    I = ctypes.c_int
    out = {"routine": "cfft1f_ + cfft1b_", "n": n, "unit": "us per round trip (2 calls through ctypes)"}
    P = fl.Lib(fl.product())
    ws, _ = P.init("cfft", n)
    ier, dummy = I(-1), np.zeros(2 * n + 8)
    xd = torch.rand(n, 2, device="cuda", dtype=torch.float64)
    xh = np.random.default_rng(3).uniform(-1, 1, 2 * n)

    def call(lib, pre, ptr, name, w):
        getattr(lib, pre + name)(ctypes.byref(I(n)), ctypes.byref(I(1)), ctypes.c_void_p(ptr), ctypes.byref(I(n)), fl.P(w),
                                 ctypes.byref(I(fl.lensav("cfft", n))), fl.P(dummy), ctypes.byref(I(2 * n)), ctypes.byref(ier))
        assert ier.value == 0

    cb.set_stream(0)
    for label, ptr, sync in (("device_pointer_us", xd.data_ptr(), True), ("host_array_us", xh.ctypes.data, False)):
        for _ in range(20):
            call(fl.product(), "", ptr, "cfft1f_", ws)
            call(fl.product(), "", ptr, "cfft1b_", ws)
        torch.cuda.synchronize()
        reps = 300
        t0 = time.perf_counter()
        for _ in range(reps):
            call(fl.product(), "", ptr, "cfft1f_", ws)
            call(fl.product(), "", ptr, "cfft1b_", ws)
            if sync:
                torch.cuda.synchronize()
        out[label] = (time.perf_counter() - t0) / reps * 1e6
    wsr, _ = cpu.L.init("cfft", n)
    xr = xh.copy()
    for _ in range(50):
        call(cpu.lib, cpu.prefix, xr.ctypes.data, "cfft1f_", wsr)
        call(cpu.lib, cpu.prefix, xr.ctypes.data, "cfft1b_", wsr)
    reps = 2000
    t0 = time.perf_counter()
    for _ in range(reps):
        call(cpu.lib, cpu.prefix, xr.ctypes.data, "cfft1f_", wsr)
        call(cpu.lib, cpu.prefix, xr.ctypes.data, "cfft1b_", wsr)
    out["cpu_baseline"] = {"value": (time.perf_counter() - t0) / reps * 1e6, "unit": "us per round trip", "cores": 1,
                           "kind": cpu.kind, "sample": f"{reps} round trips of cfft1f_+cfft1b_ N={n} on one host array"}
    cb.set_stream(torch.cuda.current_stream().cuda_stream)
    return out


def measure_cfft2_single(torch, cb, cpu, l, reps=5):
    """BASELINE configs[4] on one GPU: cfft2f_ l x l through the C ABI, device resident"""
    fl = cpu.fl
    I = ctypes.c_int
    c = torch.rand(l * l * 2, device="cuda", dtype=torch.float64) - 0.5
    P = fl.Lib(fl.product())
    ws, ls, ier0 = P.init2(l, l)
    ier, dummy = I(-1), ctypes.c_double(0)
    argv = (ctypes.byref(I(l)), ctypes.byref(I(l)), ctypes.byref(I(l)), ctypes.c_void_p(c.data_ptr()), fl.P(ws),
            ctypes.byref(I(ls)), ctypes.byref(dummy), ctypes.byref(I(min(2 * l * l, 2**31 - 1))), ctypes.byref(ier))
    fn = fl.product().cfft2f_

    def step():
        fn(*argv)
        if ier.value != 0:
            raise RuntimeError(f"cfft2f_ ier={ier.value}: {cb.last_error()}")

    l0 = cb.launch_count()
    step()
    per_call = cb.launch_count() - l0
    avg, med = time_launches(torch, step, reps, 2)
    peak, _ = hbm_peak()
    by = 2 * 2 * 16 * l * l  # two read+write passes minimum
    gbs = by / (avg * 1e-3) / 1e9
    out = {"routine": "cfft2f_", "l": l, "m": l, "ms": avg, "median_ms": med, "gbs": gbs, "frac": gbs / peak,
           "frac_of_8TBps": gbs / 8000.0, "algorithmic_bytes": by, "kernels_per_call": int(per_call),
           "gflops_nominal": 5.0 * l * l * math.log2(l * l) / (avg * 1e-3) / 1e9,
           "note": "GB/s and frac are against the 2-pass minimum (one read + one write per dimension)"}
    del c
    # CPU: the two sweeps of cfft2f_ are 2*l transforms of length l; time a bounded lot-parallel sample of them
    try:
        base = cpu.baseline("cfft", l, lot=8 * cpu.threads, budget_s=2.0)
        sec = base["seconds_per_pass"] / base["sample_lot"] * 2 * l
        out["cpu_baseline"] = {"value": by / sec / 1e9, "unit": "GB/s", "cores": cpu.threads, "kind": cpu.kind,
                               "sample": f"extrapolated: 2*{l} length-{l} transforms at the measured lot-parallel rate of looped "
                                         f"cfft1f_ ({base['sample']}); the strided sweep of the real cfft2f_ is slower",
                               "seconds_extrapolated": sec}
    except Exception as ex:
        out["cpu_baseline"] = {"value": None, "error": str(ex)}
    return out


def measure_configs(torch, cb, args):
    """every BASELINE config other than the headline one, same box, same run (device resident, CUDA events)"""
    cpu = CpuRef()
    out = {}

    def guarded(name, fn):
        try:
            out[name] = fn()
        except Exception as ex:  # report, never hide
            out[name] = {"error": f"{type(ex).__name__}: {ex}"}

    guarded("C1_cfft1_n1024", lambda: measure_c1(torch, cb, cpu))
    for d in "fb":
        def c3(d=d):
            r = measure_family(torch, cb, "rfft", d, 4096, 65536, reps=args.steps)
            r["cpu_baseline"] = cpu.baseline("rfft", 4096, d, budget_s=2.0)
            return r
        guarded(f"C3_rfftm{d}_n4096", c3)
    # SURVEY 8(a) note 2: the radices BASELINE names belong to the underlying real transform (N-1 cost, N+1 sint, N cosq)
    for fam, n, what in (("cost", 1001, "rfft 1000 = 2.4.5.5.5"), ("cosq", 1000, "rfft 1000 = 2.4.5.5.5"),
                         ("sint", 1000, "rfft 1001 = 7.11.13"), ("cosq", 1001, "rfft 1001 = 7.11.13"),
                         ("cost", 1000, "rfft 999 = 3.3.3.37 (generic-radix stress)"),
                         ("sint", 1001, "rfft 1002 = 2.3.167 (generic-radix stress)")):
        def c4(fam=fam, n=n, what=what):
            r = measure_family(torch, cb, fam, "f", n, 32768, reps=args.steps)
            r["radices"] = what
            r["cpu_baseline"] = cpu.baseline(fam, n, budget_s=1.0)
            return r
        guarded(f"C4_{fam}mf_n{n}", c4)
    guarded(f"C5_cfft2f_{args.l2d}", lambda: measure_cfft2_single(torch, cb, cpu, args.l2d))
    return out


# ------------------------------------------------------------------------------------------------------------------
# BASELINE configs[4] across GPUs: sharded cfft2f with the transposes fused into the FFT kernels + oracle parity
# ------------------------------------------------------------------------------------------------------------------
def slab_input(torch, l, m_loc, rank):
    """this rank's columns of a NON-separable random matrix (i.i.d. uniform(-0.5, 0.5), seed by rank)"""
    g = torch.Generator(device="cuda").manual_seed(99 + rank)
    return torch.view_as_complex(torch.rand(m_loc, l, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)


def cfft2_parity(torch, dist, x_in, y_out, l, m, rank, world, samples=32):
    """Oracle check of the sharded forward transform on `samples` full output columns and `samples` full output rows.

    Y[k1, k2] = 1/(l m) sum_{i,j} X[i, j] w_l^{i k1} w_m^{j k2}  (cfft2f_, fftpack.c:2363-2434: two cfftmf_ sweeps).
    Column k2 of Y is the oracle's cfft1f_ (length l, scaled 1/l) of v[i] = 1/m sum_j X[i, j] w_m^{j k2}; row k1 is the
    oracle's cfft1f_ (length m) of u[j] = 1/l sum_i X[i, j] w_l^{i k1}.  The single-bin DFT sums v and u are formed
    with plain FP64 dot products (16384 terms: error ~1e-15 relative), the length-16384 transforms by the CPU oracle.
    Slabs are [m_loc][l] (column j contiguous).  Returns max relative L2 over all sampled lines (max over ranks)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fftlibs as fl
    ORC = fl.Lib(fl.oracle(), "orc_")
    m_loc = m // world
    dev = x_in.device
    rng = np.random.default_rng(2024)  # same sample on every rank
    cols = sorted(set(int(v) for v in rng.integers(0, m, samples)) | {0, m - 1, m_loc - 1, m_loc % m})
    rows = sorted(set(int(v) for v in rng.integers(0, l, samples)) | {0, l - 1, l // 2 + 1})
    j_glob = torch.arange(rank * m_loc, (rank + 1) * m_loc, device=dev, dtype=torch.float64)
    i_glob = torch.arange(l, device=dev, dtype=torch.float64)
    worst = 0.0
    ws_l, ws_m = ORC.init("cfft", l)[0], ORC.init("cfft", m)[0]
    # ---- output columns k2: v = (1/m) sum_j X[:, j] w_m^{j k2}, partial over my columns, summed over ranks
    for k2 in cols:
        ph = torch.remainder(j_glob * k2, m) * (-2.0 * math.pi / m)
        w = torch.complex(torch.cos(ph), torch.sin(ph))
        v = (w.unsqueeze(0) @ x_in).squeeze(0) / m  # [l]
        if world > 1:
            vr = torch.view_as_real(v).contiguous()
            dist.all_reduce(vr)
            v = torch.view_as_complex(vr)
        owner = k2 // m_loc
        if rank == owner:
            want, ier = ORC.run1("cfft", "f", l, v.cpu().numpy(), ws=ws_l)
            got = y_out[k2 - owner * m_loc].cpu().numpy()
            worst = max(worst, fl.rel_l2(got, want))
    # ---- output rows k1: u[j] = (1/l) sum_i X[i, j] w_l^{i k1} for my columns, gathered over ranks
    for k1 in rows:
        ph = torch.remainder(i_glob * k1, l) * (-2.0 * math.pi / l)
        w = torch.complex(torch.cos(ph), torch.sin(ph))
        u = (x_in @ w) / l  # [m_loc]
        got = y_out[:, k1].contiguous()
        if world > 1:
            ug = [torch.empty(m_loc, 2, device=dev, dtype=torch.float64) for _ in range(world)]
            gg = [torch.empty(m_loc, 2, device=dev, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(ug, torch.view_as_real(u).contiguous())
            dist.all_gather(gg, torch.view_as_real(got).contiguous())
            u, got = torch.view_as_complex(torch.cat(ug)), torch.view_as_complex(torch.cat(gg))
        if rank == (k1 % world):
            want, ier = ORC.run1("cfft", "f", m, u.cpu().numpy(), ws=ws_m)
            worst = max(worst, fl.rel_l2(got.cpu().numpy(), want))
    t = torch.tensor([worst], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), len(cols), len(rows)


def nvlink_tx_kib(index):
    """sum over the links of GPU `index` of the NVLink data-transmit counter (KiB), or None (nvidia-smi nvlink -gt d)"""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
        tot, seen = 0, False
        for ln in out.splitlines():
            if "Data Tx" in ln:
                tot += int(ln.split(":")[-1].strip().split()[0])
                seen = True
        return tot if seen else None
    except Exception:
        return None


def run_cfft2_sharded(args, torch, dist, cb, rank, world, barrier, t1_ms=None):
    """sharded cfft2f l x l over `world` GPUs (strong scaling); returns the `cfft2` object of the JSON line"""
    from cfftpack_b200.dist import Cfft2Sharded, Cfft2ShardedP2P
    l = m = args.l2d
    m_loc = m // world
    x_in = slab_input(torch, l, m_loc, rank)
    out = {"workload": f"cfft2f {l}x{m} c128 column slabs over {world} GPUs, 2 transposes (BASELINE configs[4])", "n_gpus": world}
    fused = True
    try:  # fused path: FFT kernels store straight into the peers' slabs (NVLink P2P on symmetric memory)
        p2p = Cfft2ShardedP2P(l, m)
        slab = p2p.slab
        forward = p2p.forward
        out["exchange"] = "transposes fused into the FFT kernels (P2P stores over NVLink, symmetric memory)"
    except Exception as ex:  # never silent: the line says which path ran and why
        fused = False
        plan = Cfft2Sharded(l, m)
        slab = torch.empty_like(x_in)
        forward = lambda: plan.forward(slab)
        out["exchange"] = "NCCL all_to_all_single (baseline path)"
        out["fallback"] = True
        out["fallback_reason"] = f"{type(ex).__name__}: {ex}"
        if rank == 0:
            print(f"bench.py: fused P2P cfft2 path unavailable, NCCL baseline used: {ex}", file=sys.stderr, flush=True)
    # ---- parity first: one forward transform of the random matrix against the CPU oracle
    slab.copy_(x_in)
    barrier()
    forward()
    barrier()
    err, ncols, nrows = cfft2_parity(torch, dist, x_in, slab, l, m, rank, world)
    out["parity_rel_l2"] = err
    out["parity_bar"] = 1e-12 * math.log2(l * m)
    out["parity_ok"] = bool(err <= out["parity_bar"])
    out["parity_sample"] = f"{ncols} full output columns + {nrows} full output rows of a non-separable random matrix vs oracle cfft1f_"
    # ---- timing (values shrink by 1/(l m) per step; refresh the slab between steps outside the timed region is not
    # needed for <= 10 steps: 2^-28 per step stays far inside the FP64 normal range)
    slab.copy_(x_in)
    steps = max(3, min(args.steps, 10))
    for _ in range(3):
        forward()
    barrier()
    l0 = cb.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tx0 = nvlink_tx_kib(torch.cuda.current_device()) if rank == 0 else None
    barrier()
    e0.record()
    for _ in range(steps):
        forward()
    e1.record()
    barrier()
    tx1 = nvlink_tx_kib(torch.cuda.current_device()) if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    link_bytes_one = 16 * l * m * (world - 1) // (world * world)  # per GPU per exchange
    if tx0 is not None and tx1 is not None:
        out["nvlink_tx_bytes_per_step_counted"] = (tx1 - tx0) * 1024 // steps
        out["nvlink_counter_source"] = "nvidia-smi nvlink -gt d on rank 0's GPU, before and after the timed loop (includes barrier traffic)"
    out.update({"ms": ms, "steps": steps, "kernels_per_step": int((cb.launch_count() - l0) // steps),
                "hbm_gbs_2pass_minimum": 2 * 2 * 16 * l * m / (ms * 1e-3) / 1e9,
                "gflops_nominal": 5.0 * l * m * math.log2(l * m) / (ms * 1e-3) / 1e9,
                "nvlink_bytes_per_gpu": 2 * link_bytes_one, "nvlink_gbs": 2 * link_bytes_one / (ms * 1e-3) / 1e9,
                "frac_of_900": 2 * link_bytes_one / (ms * 1e-3) / 1e9 / NVLINK_GBS,
                "nvlink_floor_ms": 2 * link_bytes_one / (NVLINK_GBS * 1e9) * 1e3,
                "note": "nvlink_gbs = bytes each GPU sends in both exchanges / whole-transform time (local FFT time included)"})
    if t1_ms:
        out["single_gpu_ms"] = t1_ms
        out["strong_scaling_eff"] = t1_ms / (world * ms)
    del x_in
    return out


def run_long1d(args, torch, dist, cb, rank, world, barrier, log2n=28, nbins=64):
    """one complex transform of 2^log2n points (cfft1f_ semantics), on one GPU through the C ABI
    (six-step over the four-step sweeps) or in natural order across the GPUs with the exchanges fused into the kernels as
    P2P stores.  Parity: sampled output bins against the direct O(N) DFT sums of the definition (test/naivepack.c)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from run_dist1d import dft_bins
    n = 1 << log2n
    n_loc = n // world
    g = torch.Generator(device="cuda").manual_seed(11 + rank)
    x0 = torch.view_as_complex(torch.rand(n_loc, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    out = {"workload": f"cfft1f 2^{log2n} complex points in natural order over {world} GPU(s)", "n_gpus": world}
    if world > 1:
        from cfftpack_b200.dist import Cfft1ShardedP2P
        plan = Cfft1ShardedP2P(log2n)
        src = plan.x
        forward = plan.forward
        out["exchange"] = "3 exchanges fused into the kernels (P2P stores over NVLink): transpose, twiddled length-Mm pass, length-L pass"
    else:
        p1 = cb.Plan("cfft", n)
        src = x0.clone()

        def forward():
            ier = p1.multi("f", src.data_ptr(), 1, n, 1, n)
            if ier != 0:
                raise RuntimeError(f"cfft1 2^{log2n}: ier={ier}: {cb.last_error()}")
            return src
    src.copy_(x0)
    barrier()
    y = forward()
    barrier()
    gen = torch.Generator().manual_seed(3)
    bins = sorted(set([0, 1, n // 2, n - 1, n_loc - 1, n_loc % n] + torch.randint(0, n, (nbins,), generator=gen).tolist()))
    want = dft_bins(x0, rank * n_loc, n, bins, -1.0, world) / n
    got = torch.zeros(len(bins), 2, device="cuda", dtype=torch.float64)
    for bi, k in enumerate(bins):
        if k // n_loc == rank:
            v = y[k - rank * n_loc]
            got[bi, 0], got[bi, 1] = v.real, v.imag
    ms2 = torch.tensor([float((y.abs() ** 2).sum())], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(got)
        dist.all_reduce(ms2)
    rms = math.sqrt(float(ms2.item()) / n)
    err = float((torch.view_as_complex(got) - want).abs().max()) / rms
    out.update({"parity_max_bin_err_over_rms": err, "parity_bar": 1e-12 * log2n, "parity_ok": bool(err <= 1e-12 * log2n),
                "parity_sample": f"{len(bins)} output bins vs direct DFT sums over all 2^{log2n} inputs"})
    src.copy_(x0)
    steps = max(3, min(args.steps, 5))
    for _ in range(2):
        forward()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = cb.launch_count()
    e0.record()
    for _ in range(steps):
        forward()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    link = 3 * 16 * n * (world - 1) // (world * world)
    out.update({"ms": ms, "steps": steps, "kernels_per_step": int((cb.launch_count() - l0) // steps),
                "hbm_gbs_algorithmic": 2 * 16 * n / (ms * 1e-3) / 1e9, "gflops_nominal": 5.0 * n * log2n / (ms * 1e-3) / 1e9})
    if world > 1:
        out.update({"nvlink_bytes_per_gpu": link, "nvlink_gbs": link / (ms * 1e-3) / 1e9,
                    "frac_of_900": link / (ms * 1e-3) / 1e9 / NVLINK_GBS, "nvlink_floor_ms": link / (NVLINK_GBS * 1e9) * 1e3})
    else:
        peak, _ = hbm_peak()
        out.update({"sweeps": 4, "frac_per_sweep": 4 * 2 * 16 * n / (ms * 1e-3) / 1e9 / peak})
    return out


def run_cfft2(args, torch, dist, cb, rank, local_rank, world, barrier):
    """--workload cfft2: the cfft2 measurement as the main line (strong scaling)"""
    l = m = args.l2d
    if world > 1:
        obj = run_cfft2_sharded(args, torch, dist, cb, rank, world, barrier)
        ms = obj["ms"]
        roofline = {"bound": "nvlink", "achieved": obj["nvlink_gbs"], "peak": NVLINK_GBS, "unit": "GB/s", "frac": obj["frac_of_900"],
                    "traffic": None}
    else:
        obj = measure_cfft2_single(torch, cb, CpuRef(), l, reps=min(args.steps, 10))
        ms = obj["ms"]
        peak, src = hbm_peak()
        sweeps = obj["kernels_per_call"]
        sweep = 2 * 16 * l * m  # every sweep kernel reads and writes the whole array once
        ach = sweep / (ms / max(sweeps, 1) * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "peak_source": src, "kernel": f"pow2_tile_tma_kernel ({sweeps} sweeps per cfft2f_)",
                    "algorithmic_bytes_per_launch": sweep, "avg_launch_ms": ms / max(sweeps, 1)}
    if rank == 0:
        print(json.dumps({
            "metric": f"cfft2f FP64 {l}x{m} algorithmic HBM GB/s (2-pass minimum)", "value": 2 * 2 * 16 * l * m / (ms * 1e-3) / 1e9,
            "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cfft2f {l}x{m} c128 column slabs, transpose x2 (BASELINE configs[4])"},
            "roofline": roofline, "cfft2": obj, "gpu_launches": int(obj.get("kernels_per_step", obj.get("kernels_per_call", 0)) * args.steps)}),
            flush=True)


def bind_to_gpu_numa_node(torch, local_rank):
    """Run this rank's host threads on the CPUs of the NUMA node its GPU hangs off, so that the pinned staging buffers
    of the end-to-end leg are first-touched (and therefore placed) next to the GPU's PCIe root instead of all ranks
    sharing node 0.  Best effort: returns a description, never raises."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        if hasattr(pr, "pci_bus_id") and hasattr(pr, "pci_device_id"):
            bdf = f"{getattr(pr, 'pci_domain_id', 0):04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:
            bdf = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().lower()
            if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
                bdf = bdf[4:]  # sysfs uses a 4-digit PCI domain
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return {"numa_node": node, "bound": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"numa_node": node, "bound": False}
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "bound": True, "cpus": len(cpus)}
    except Exception as ex:
        return {"numa_node": None, "bound": False, "why": str(ex)[:80]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfftm", choices=["cfftm", "rfftm", "cfft2"])
    ap.add_argument("--l2d", type=int, default=16384, help="cfft2: matrix is l2d x l2d (BASELINE configs[4])")
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--lot", type=int, default=LOT_DEFAULT)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-array leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE configs (profiling runs)")
    ap.add_argument("--no-cfft2", action="store_true", help="N > 1: skip the sharded cfft2f measurement")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    fam = "cfft" if args.workload == "cfftm" else "rfft"
    n, lot = args.n, args.lot
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, fam, n, lot, rank, world)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    if not os.path.exists(os.path.join(ROOT, "cfftpack_b200", "libcfftpack_b200.so")):
        ge.build()
    import cfftpack_b200 as cb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: cfftpack_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    cb.set_stream(stream.cuda_stream)
    if args.workload == "cfft2":
        run_cfft2(args, torch, dist, cb, rank, local_rank, world, barrier)
        if world > 1:
            dist.destroy_process_group()
        return
    esz = 2 if fam == "cfft" else 1
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    x = torch.rand(lot * n * esz, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    plan = cb.Plan(fam, n)

    def step():
        ier = plan.multi("f", x.data_ptr(), lot, n, 1, lot * n)
        if ier != 0:
            raise RuntimeError(f"{fam}mf_ ier={ier}: {cb.last_error()}")

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi come up before the timed region
    # in-place forward transforms scale by 1/N each step, so values shrink towards 0 without ever leaving FP64
    # normal range for K <= 20 (4096^-20 ~ 1e-72); timing is data independent.
    launches0 = cb.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    sampler.mark()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    launches = cb.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    bytes_rank = algorithmic_bytes(fam, n, lot)
    value = world * bytes_rank / (ms_per_step * 1e-3) / 1e9
    gflops = world * nominal_flops(fam, n, lot) / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C ABI with host arrays (pinned): H2D + transform + D2H inside the timed region
    e2e = None
    try:
        if args.no_e2e:
            raise RuntimeError('skipped (--no-e2e)')
        h = torch.empty(lot * n * esz, dtype=torch.float64, pin_memory=True)
        h.uniform_(-1, 1)
        hplan = cb.Plan(fam, n)

        def e2e_step():
            ier = hplan.multi("f", h.data_ptr(), lot, n, 1, lot * n)
            if ier != 0:
                raise RuntimeError(f"host-array {fam}mf_ ier={ier}: {cb.last_error()}")

        e2e_step()  # warm-up: allocates the staging buffer
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()  # synchronous for host arrays, like the reference
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.e2e_steps
        tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        nbytes = lot * n * esz * 8
        # ceiling of this leg: the same pinned array copied to the device and back with NO transform, the two directions
        # on separate streams and in 64 MiB pieces like the library's staging pipeline (PCIe is full duplex)
        ceil_gbs = None
        try:
            dbuf = torch.empty_like(h, device="cuda")
            s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
            pieces = list(zip(h.split(1 << 23), dbuf.split(1 << 23)))
            barrier()
            t0 = time.perf_counter()
            for _ in range(2):
                for hp, dp in pieces:
                    with torch.cuda.stream(s_in):
                        dp.copy_(hp, non_blocking=True)
                        ev_ = torch.cuda.Event()
                        ev_.record()
                    with torch.cuda.stream(s_out):
                        s_out.wait_event(ev_)
                        hp.copy_(dp, non_blocking=True)
                torch.cuda.synchronize()
            dtc = (time.perf_counter() - t0) / 2
            ttc = torch.tensor([dtc], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ttc, op=dist.ReduceOp.MAX)
            ceil_gbs = world * bytes_rank / float(ttc.item()) / 1e9
            del dbuf
        except Exception as ex:
            ceil_gbs = f"unavailable: {ex}"
        e2e = {"value": world * bytes_rank / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": nbytes,
               "copy_only_ceiling_gbs": ceil_gbs,
               "d2h_bytes_per_step": nbytes, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "note": f"host pinned array -> {fam}mf_ C ABI -> host; copies inside the timed region", "host_numa": numa}
        del h
    except Exception as ex:  # report, never hide
        e2e = {"value": None, "unit": "GB/s", "error": str(ex)}

    try:
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline below uses every host core again
    except Exception:
        pass
    peak, peak_src = hbm_peak()
    kern_ms = sorted(per_launch_ms)[len(per_launch_ms) // 2]
    avg_ms = sum(per_launch_ms) / len(per_launch_ms)
    achieved = bytes_rank / (avg_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(f"{fam}m_{n}")
            traffic_src = "constant from profiles/traffic.json (one earlier `ncu --set full` capture of this kernel), not measured in this run"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel": f"pow2_{'c2c' if fam == 'cfft' else 'r2c'}_stream_kernel<Pow2Cfg<{int(math.log2(n))}>, DIR=-1>",
                "algorithmic_bytes_per_launch": bytes_rank, "avg_launch_ms": avg_ms, "median_launch_ms": kern_ms,
                "frac_of_8TBps_nominal": achieved / 8000.0}
    del x
    torch.cuda.empty_cache()

    # ---- N > 1: BASELINE configs[4], the one path with an exchange step, in the same driver-run line
    cfft2 = None
    if world > 1 and not args.no_cfft2:
        try:
            t1 = None
            try:
                t1 = float(json.load(open(os.path.join(ROOT, "profiles", "cfft2_single_gpu.json")))["ms"])
            except Exception:
                pass
            cfft2 = run_cfft2_sharded(args, torch, dist, cb, rank, world, barrier, t1_ms=t1)
            if t1:
                cfft2["single_gpu_ms_source"] = "profiles/cfft2_single_gpu.json (1-GPU run of this bench)"
        except Exception as ex:
            cfft2 = {"error": f"{type(ex).__name__}: {ex}"}

    long1d = None
    if not args.no_cfft2:
        try:
            long1d = run_long1d(args, torch, dist, cb, rank, world, barrier)
        except Exception as ex:
            long1d = {"error": f"{type(ex).__name__}: {ex}"}
        torch.cuda.empty_cache()

    configs = None
    cpu = None
    if rank == 0:
        if not args.no_cpu_baseline:
            try:
                cpu = CpuRef().baseline(fam, n, budget_s=10.0, shipped=True)
            except Exception as ex:
                cpu = {"value": None, "unit": "GB/s", "error": str(ex)}
        if world == 1 and not args.no_configs:
            try:
                configs = measure_configs(torch, cb, args)
            except Exception as ex:
                configs = {"error": f"{type(ex).__name__}: {ex}"}
        out = {"metric": f"batched {fam}mf FP64 N={n} algorithmic HBM GB/s", "value": value, "unit": "GB/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "gflops_5nlogn": gflops, "frac_of_8TBps": value / world / 8000.0,
               "config": workload_config(fam, n, lot),
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if cfft2 is not None:
            out["cfft2"] = cfft2
        if long1d is not None:
            out["cfft1_long"] = long1d
        if configs is not None:
            out["configs"] = configs
        print(json.dumps(out), flush=True)
    if world > 1:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
