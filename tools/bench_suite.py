"""Developer tool: device-resident timings of every BASELINE.json config shape (CUDA events, median of reps).
Prints one line per case with algorithmic GB/s (one read + one write of the payload per pass)."""
import ctypes, json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import cfftpack_b200 as cb
import fftlibs as fl

def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]

def case(fam, n, lot, inc=None, jump=None, d="f"):
    inc = 1 if inc is None else inc
    jump = n if jump is None else jump
    esz = 2 if fam == "cfft" else 1
    span = (lot - 1) * jump + (n - 1) * inc + 1
    x = (torch.rand(span * esz, device="cuda", dtype=torch.float64) - 0.5)
    plan = cb.Plan(fam, n)
    def run():
        ier = plan.multi(d, x.data_ptr(), lot, jump, inc, span)
        assert ier == 0, (ier, cb.last_error())
    ms = timeit(run)
    by = 2 * 8 * esz * n * lot
    print(f"{fam}m{d} n={n:6d} lot={lot:6d} inc={inc:6d} jump={jump:6d}: {ms:8.3f} ms  {by / ms / 1e6:8.1f} GB/s  ({by / ms / 1e6 / 65.475:5.1f}% of measured HBM)", flush=True)

def case2d(l, m, d="f"):
    c = torch.rand(l * m * 2, device="cuda", dtype=torch.float64) - 0.5
    P = fl.Lib(fl.product())
    ws, ls, ier = P.init2(l, m)
    I = ctypes.c_int; ierc = I(-1); dummy = ctypes.c_double(0)
    args = (ctypes.byref(I(l)), ctypes.byref(I(l)), ctypes.byref(I(m)), ctypes.c_void_p(c.data_ptr()), fl.P(ws),
            ctypes.byref(I(ls)), ctypes.byref(dummy), ctypes.byref(I(2 * l * m)), ctypes.byref(ierc))
    fn = getattr(fl.product(), "cfft2" + d + "_")
    def run():
        fn(*args); assert ierc.value == 0, cb.last_error()
    ms = timeit(run, reps=5, warm=2)
    by = 2 * 2 * 16 * l * m  # two passes minimum
    print(f"cfft2{d} {l}x{m}: {ms:8.3f} ms  {by / ms / 1e6:8.1f} GB/s vs 2-pass minimum ({by / ms / 1e6 / 65.475:5.1f}% of measured HBM)", flush=True)

def case_c1():
    """config 1: single cfft1f_/cfft1b_ N=1024 (latency): device pointer (async launch + sync) and host array (staged)"""
    import numpy as np
    P = fl.Lib(fl.product())
    n = 1024
    ws, _ = P.init("cfft", n)
    I = ctypes.c_int
    ier = I(-1); dummy = np.zeros(2 * n + 8)
    xd = torch.rand(n, 2, device="cuda", dtype=torch.float64)
    xh = np.random.rand(n).astype(np.complex128)
    def call(ptr, name):
        getattr(fl.product(), name)(ctypes.byref(I(n)), ctypes.byref(I(1)), ctypes.c_void_p(ptr), ctypes.byref(I(n)), fl.P(ws),
                                    ctypes.byref(I(fl.lensav("cfft", n))), fl.P(dummy), ctypes.byref(I(2 * n)), ctypes.byref(ier))
        assert ier.value == 0
    for label, ptr, sync in (("device pointer", xd.data_ptr(), True), ("host array", xh.ctypes.data, False)):
        for _ in range(20):
            call(ptr, "cfft1f_"); call(ptr, "cfft1b_")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 200
        for _ in range(reps):
            call(ptr, "cfft1f_"); call(ptr, "cfft1b_")
            if sync: torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"cfft1f+cfft1b N=1024 round trip, {label}: {dt * 1e6:8.1f} us per round trip (2 calls, via ctypes)", flush=True)

def case_options(n=4096, lot=16384):
    """batched option valuation (test/vargamma.c) end to end from host parameters to host values,
    against the oracle's single-option CPU version on one core"""
    import numpy as np
    rng = np.random.default_rng(1)
    K = rng.uniform(80, 120, lot)
    cb.option_convolution(n, 100.0, K, 0.12, -0.14, 0.2, 1.0, 0.05)  # warm plans and scratch
    for bs in (True, False):
        t0 = time.perf_counter()
        v, N = cb.option_convolution(n, 100.0, K, 0.12, -0.14, 0.2, 1.0, 0.05, call=True, black_scholes=bs)
        dt = time.perf_counter() - t0
        t1 = time.perf_counter()
        w = np.array([fl.option_oracle(n, (100.0, float(k), 0.12, -0.14, 0.2, 1.0, 0.05, 1, int(bs))) for k in K[:64]])
        dc = (time.perf_counter() - t1) / 64
        err = float(np.max(np.abs(v[:64] - w) / np.abs(w)))
        print(f"option convolution {'BS' if bs else 'VG'} N={N} lot={lot}: {dt * 1e3:8.2f} ms = {lot / dt / 1e3:8.1f} k options/s "
              f"(CPU oracle, 1 core: {1 / dc / 1e3:6.2f} k options/s; max rel diff {err:.1e})", flush=True)


which = sys.argv[1].split(",") if len(sys.argv) > 1 else ["c1", "c2", "c3", "c4", "c5", "misc"]
if "c1" in which:
    case_c1()
if "c2" in which:
    case("cfft", 4096, 65536); case("cfft", 4096, 65536, d="b")
if "c3" in which:
    case("rfft", 4096, 65536); case("rfft", 4096, 65536, d="b")
if "c4" in which:
    for fam, n in (("cost", 1001), ("sint", 1000), ("cosq", 1000), ("cosq", 1001), ("cost", 1000), ("sint", 1001)):
        case(fam, n, 32768); case(fam, n, 32768, d="b")
if "c5" in which:
    case2d(16384, 16384); case2d(4096, 4096)
if "n4" in which:
    case_options()
if "misc" in which:
    case("cfft", 1024, 262144); case("cfft", 256, 1048576); case("cfft", 64, 4194304); case("cfft", 8192, 32768)
    case("cfft", 1000, 262144); case("cfft", 4096, 65536, inc=65536, jump=1); case("rfft", 1000, 262144)
    case("cfft", 16384, 16384); case("cfft", 16384, 16384, inc=16384, jump=1)
