#!/usr/bin/env python
"""bench.py -- the headline measurement (BASELINE.json): batched cfftmf_ FP64, N=4096, lot=65536 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfftm|rfftm]

One "step" = one forward pass of the hot path (cfftmf_, in place) over one batch of synthetic data.
  value     : algorithmic GB/s (one read + one write of the payload) of the whole job, data resident
              in HBM, timed on the device with CUDA events on the launching stream, max over ranks.
  e2e       : the same metric through the C ABI with HOST (pinned) arrays: H2D copy + transform + D2H copy timed.
  roofline  : achieved GB/s of the dominant kernel vs the measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference : the unmodified reference (oracle/_ref, compiled from /root/reference) or, if
              that is absent, the oracle port, lot-parallel over all host cores on a bounded sample.
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


def algorithmic_bytes(fam, n, lot):
    return 2 * (16 if fam == "cfft" else 8) * n * lot


def nominal_flops(fam, n, lot):
    return (5.0 if fam == "cfft" else 2.5) * n * math.log2(n) * lot


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


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


def cpu_reference(fam, n, threads, budget_s=12.0):
    """lot-parallel CPU run of the reference (or the oracle port) on a bounded sample; returns a dict."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fftlibs as fl
    orc = fl.oracle()
    if orc is None:
        raise RuntimeError("oracle/liboracle.so missing: run python __graft_entry__.py")
    ref = fl.ref()
    kind = "reference" if ref is not None else "port"
    lib = ref if ref is not None else orc
    prefix = "" if ref is not None else "orc_"
    L = fl.Lib(lib, prefix)
    ws, ier = L.init(fam, n)
    assert ier == 0
    fn = ctypes.cast(getattr(lib, prefix + fam + "1f_"), ctypes.c_void_p)
    orc.orc_lot_parallel.restype = ctypes.c_double
    orc.orc_lot_parallel.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lot = 256 * threads
    x = fl.rand_input(fam, lot * n, 11)
    is_c = 1 if fam == "cfft" else 0
    t1 = orc.orc_lot_parallel(fn, is_c, lot, n, fl.P(x), fl.P(ws), fl.lensav(fam, n), threads, 1)  # warm + calibrate
    reps = max(1, min(64, int(budget_s / max(t1, 1e-3))))
    t = orc.orc_lot_parallel(fn, is_c, lot, n, fl.P(x), fl.P(ws), fl.lensav(fam, n), threads, reps)
    per_pass = t / reps
    gbs = algorithmic_bytes(fam, n, lot) / per_pass / 1e9
    # the as-shipped batched routine on the contiguous layout, one thread, small lot (cache-hostile)
    lot_m = 64
    xm = fl.rand_input(fam, lot_m * n, 12)
    t0 = time.perf_counter()
    _, ier = L.runm(fam, "f", lot_m, n, n, 1, xm)
    tm = time.perf_counter() - t0
    return {"value": gbs, "unit": "GB/s", "cores": threads, "kind": kind,
            "sample": f"{lot} sequences x {reps} passes of looped {fam}1f_ N={n}, lot-parallel over {threads} threads "
                      f"({per_pass / lot * 1e6:.1f} us/sequence/thread-group); as-shipped {fam}mf_ lot={lot_m} 1 thread: "
                      f"{tm / lot_m * 1e6:.0f} us/sequence",
            "gflops": nominal_flops(fam, n, lot) / per_pass / 1e9, "seconds_per_pass": per_pass, "sample_lot": lot}


def run_reference(args, fam, n, lot, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # each step = one lot-parallel pass over a bounded sample
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fftlibs as fl
    orc = fl.oracle()
    ref = fl.ref()
    lib, prefix, kind = (ref, "", "reference") if ref is not None else (orc, "orc_", "port")
    L = fl.Lib(lib, prefix)
    ws, ier = L.init(fam, n)
    fn = ctypes.cast(getattr(lib, prefix + fam + "1f_"), ctypes.c_void_p)
    orc.orc_lot_parallel.restype = ctypes.c_double
    orc.orc_lot_parallel.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    slot = 1024 * threads  # sequences per step: large enough that thread start-up is amortised (~1 GiB at 16 threads)
    x = fl.rand_input(fam, slot * n, 11)
    is_c = 1 if fam == "cfft" else 0
    for _ in range(args.warmup):
        orc.orc_lot_parallel(fn, is_c, slot, n, fl.P(x), fl.P(ws), fl.lensav(fam, n), threads, 1)
    t = 0.0
    for _ in range(args.steps):
        t += orc.orc_lot_parallel(fn, is_c, slot, n, fl.P(x), fl.P(ws), fl.lensav(fam, n), threads, 1)
    per = t / args.steps
    val = algorithmic_bytes(fam, n, slot) / per / 1e9
    out = {"impl": "reference", "metric": f"batched {fam}mf FP64 N={n} algorithmic HBM GB/s", "value": val, "unit": "GB/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": per * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "gflops_5nlogn": nominal_flops(fam, n, slot) / per / 1e9,
           "config": {"workload": f"{fam}mf N={n} lot={lot} per GPU, inc=1, jump=N (BASELINE configs[1])",
                      "sample": f"each step = {slot} sequences (bounded sample), looped {fam}1f_ lot-parallel"},
           "cpu_baseline": {"value": val, "unit": "GB/s", "cores": threads, "kind": kind,
                            "sample": f"{slot} sequences per step, {threads} threads"},
           "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def run_cfft2(args, torch, dist, cb, rank, local_rank, world, barrier):
    """BASELINE configs[4]: cfft2f 2-D c2c FP64 l x l, column slabs over the ranks, all-to-all transposes (strong scaling)."""
    from cfftpack_b200.dist import Cfft2Sharded, Cfft2ShardedP2P
    l = m = args.l2d
    g = torch.Generator(device="cuda").manual_seed(99 + rank)
    slab = torch.view_as_complex(torch.rand(m // world, l, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    mode = "single GPU"
    if world > 1:
        try:  # fused path: FFT kernels store straight into the peers' slabs (NVLink P2P on symmetric memory)
            p2p = Cfft2ShardedP2P(l, m)
            p2p.slab.copy_(slab)
            plan = type("P", (), {"forward": staticmethod(lambda s: p2p.forward())})
            mode = "transposes fused into the FFT kernels (P2P stores over NVLink, symmetric memory)"
        except Exception as ex:  # e.g. sizes outside the fused path
            plan = Cfft2Sharded(l, m)
            mode = f"NCCL all-to-all ({ex})"
    else:
        plan = Cfft2Sharded(l, m)
    for _ in range(max(args.warmup, 2)):
        plan.forward(slab)
    barrier()
    steps = min(args.steps, 10)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = cb.launch_count()
    barrier()
    e0.record()
    for _ in range(steps):
        plan.forward(slab)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    hbm_bytes = 2 * 2 * 16 * l * m  # two read+write passes minimum
    link_bytes = 2 * 16 * l * m * (world - 1) // (world * world)  # per GPU, both exchanges
    roofline = None
    if world == 1:  # dominant kernel: the four-step sweep (pow2_tile_tma_kernel), 4 launches per transform, each one
        peak, src = hbm_peak()  # one read + one write of the whole array per sweep
        sweep = 2 * 16 * l * m
        ach = sweep / (ms / 4 * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": 8544292864 if l == 16384 else None, "peak_source": src,
                    "kernel": "pow2_tile_tma_kernel (4 sweeps per cfft2f_)", "algorithmic_bytes_per_launch": sweep,
                    "avg_launch_ms": ms / 4, "traffic_source": "profiles/r1_ncu_tile_tma_v1.txt"}
    if rank == 0:
        print(json.dumps({
            "roofline": roofline,
            "metric": f"cfft2f FP64 {l}x{m} algorithmic HBM GB/s (2-pass minimum)", "value": hbm_bytes / (ms * 1e-3) / 1e9,
            "unit": "GB/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 2), "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "gflops_5nlogn": 5.0 * l * m * math.log2(l * m) / (ms * 1e-3) / 1e9,
            "config": {"workload": f"cfft2f {l}x{m} c128 column slabs, transpose x2 (BASELINE configs[4])", "exchange": mode},
            "nvlink_bytes_per_gpu_per_step": link_bytes,
            "nvlink_floor_ms_at_770GBps": link_bytes / 770e9 * 1e3,
            "gpu_launches": int(cb.launch_count() - launches0)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfftm", choices=["cfftm", "rfftm", "cfft2"])
    ap.add_argument("--l2d", type=int, default=16384, help="cfft2 workload: matrix is l2d x l2d (BASELINE configs[4])")
    ap.add_argument("--n", type=int, default=N_DEFAULT)
    ap.add_argument("--lot", type=int, default=LOT_DEFAULT)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-array leg")
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
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.workload == "cfft2":
        run_cfft2(args, torch, dist, cb, rank, local_rank, world, barrier)
        if world > 1:
            dist.destroy_process_group()
        return
    esz = 2 if fam == "cfft" else 1
    stream = torch.cuda.current_stream()
    cb.set_stream(stream.cuda_stream)
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
        e2e = {"value": world * bytes_rank / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
               "note": f"host pinned array -> {fam}mf_ C ABI -> host; copies inside the timed region"}
        del h
    except Exception as ex:  # report, never hide
        e2e = {"value": None, "unit": "GB/s", "error": str(ex)}

    peak, peak_src = hbm_peak()
    kern_ms = sorted(per_launch_ms)[len(per_launch_ms) // 2]
    avg_ms = sum(per_launch_ms) / len(per_launch_ms)
    achieved = bytes_rank / (avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"{fam}m_{n}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "kernel": f"pow2_{'c2c' if fam == 'cfft' else 'r2c'}_stream_kernel<Pow2Cfg<{int(math.log2(n))}>, DIR=-1>",
                "algorithmic_bytes_per_launch": bytes_rank, "avg_launch_ms": avg_ms, "median_launch_ms": kern_ms,
                "frac_of_8TBps_nominal": achieved / 8000.0}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_reference(fam, n, os.cpu_count() or 1)
            except Exception as ex:
                cpu = {"value": None, "unit": "GB/s", "error": str(ex)}
        out = {"metric": f"batched {fam}mf FP64 N={n} algorithmic HBM GB/s", "value": value, "unit": "GB/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "gflops_5nlogn": gflops, "frac_of_8TBps": value / world / 8000.0,
               "config": {"workload": f"{fam}mf N={n} lot={lot} per GPU, inc=1, jump=N, in place, forward (BASELINE configs[1])",
                          "sharding": "by lot, one process per GPU, no collective",
                          "l2": f"inputs {bytes_rank // 2 >> 20} MiB per GPU >> 126 MB L2, no flush needed"},
               "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
