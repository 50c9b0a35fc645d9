import sys, os, ctypes
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import cfftpack_b200 as cb
a = 24; n = 1 << a; L = Mm = 4096
g = torch.Generator(device="cuda").manual_seed(a)
y0 = torch.rand(n, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5
Y = y0.clone(); X = torch.zeros_like(y0)
I = ctypes.c_int
peers = (ctypes.c_void_p * 1)(X.data_ptr())
ref = None
for rep in range(12):
    Y.copy_(y0); X.zero_()
    ier = I(-1)
    cb.lib.cfb200_cfft1_sharded_phase(I(1), I(-1), I(a), I(0), I(1), ctypes.c_void_p(Y.data_ptr()), peers, ctypes.byref(ier))
    cb.synchronize()
    assert ier.value == 0, cb.last_error()
    o = torch.view_as_complex(X).cpu().numpy()
    if ref is None:
        ref = o.copy(); continue
    bad = np.nonzero(o != ref)[0]
    if len(bad) == 0:
        continue
    # X[(b) * L + i]: b = output index, i = sequence
    b, i = bad // L, bad % L
    k1, k2 = b % 64, b // 64
    print("rep", rep, "bad", len(bad), "| i range", i.min(), i.max(), "distinct i", len(np.unique(i)), "| k1", np.unique(k1), "| k2 distinct", len(np.unique(k2)), np.unique(k2)[:8])
    r = o[bad] / ref[bad]
    print("    ratio abs", np.abs(r[:4]), "phase/2pi*N", (np.angle(r[:6]) / (2 * np.pi) * n))
    # per (i) list of k2
    for ii in np.unique(i)[:3]:
        print("    i", ii, "k2:", np.unique(k2[i == ii])[:20])
