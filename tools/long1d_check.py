"""Developer check: single-GPU long power-of-two cfft1f_/cfft1b_ (2^21 .. 2^28) against the CPU oracle (N <= 2^24) and
the direct DFT sums of sampled bins, with timings."""
import ctypes, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import cfftpack_b200 as cb
import fftlibs as fl
from run_dist1d import dft_bins

ORC = fl.Lib(fl.oracle(), "orc_")
for a in [int(v) for v in sys.argv[1:]] or [21, 22, 23, 24, 26, 28]:
    n = 1 << a
    g = torch.Generator(device="cuda").manual_seed(a)
    x0 = torch.view_as_complex(torch.rand(n, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    x = x0.clone()
    plan = cb.Plan("cfft", n)
    ier = plan.multi("f", x.data_ptr(), 1, n, 1, n)
    cb.synchronize()
    assert ier == 0, (a, ier, cb.last_error())
    gen = torch.Generator().manual_seed(3)
    bins = sorted(set([0, 1, n // 2, n - 1] + torch.randint(0, n, (64,), generator=gen).tolist()))
    want = dft_bins(x0, 0, n, bins, -1.0, 1) / n
    rms = float(torch.sqrt((x.abs() ** 2).mean()))
    err = float((x[torch.tensor(bins, device="cuda")] - want).abs().max() / rms)
    msg = f"cfft1f n=2^{a}: sampled-bin err/rms {err:.2e}"
    if a <= 24:
        want_all, ier2 = ORC.run1("cfft", "f", n, x0.cpu().numpy())
        msg += f", vs oracle rel-L2 {fl.rel_l2(x.cpu().numpy(), want_all):.2e}"
    ier = plan.multi("b", x.data_ptr(), 1, n, 1, n)
    cb.synchronize()
    rt = float((torch.view_as_real(x) - torch.view_as_real(x0)).norm() / torch.view_as_real(x0).norm())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    plan.multi("f", x.data_ptr(), 1, n, 1, n); torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        plan.multi("f", x.data_ptr(), 1, n, 1, n)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(msg + f", round trip {rt:.2e} (bar {1e-12 * a:.1e}); {ms:.3f} ms = {2 * 16 * n / ms / 1e6:.0f} GB/s algorithmic", flush=True)
    del x, x0
