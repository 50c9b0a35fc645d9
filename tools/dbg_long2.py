import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import cfftpack_b200 as cb
a = int(sys.argv[1]); n = 1 << a
g = torch.Generator(device="cuda").manual_seed(a)
x0 = torch.view_as_complex(torch.rand(n, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
aL = a // 2 if a >= 24 else (9 if a == 21 else 10); aM = a - aL; L, Mm = 1 << aL, 1 << aM
plan = cb.Plan("cfft", n)
xs = x0.clone()
ref = None
for rep in range(6):
    xs.copy_(x0)
    assert plan.multi("f", xs.data_ptr(), 1, n, 1, n) == 0
    cb.synchronize()
    o = xs.cpu().numpy()
    if ref is None:
        ref = o.copy(); continue
    bad = np.nonzero(o != ref)[0]
    print("rep", rep, "bad", len(bad))
    if len(bad):
        # the twiddle acts on T[i + L b] before step B (an FFT over i): undo step B is hard; instead look at structure
        aa, bb = bad // Mm, bad % Mm
        print("  distinct b", len(np.unique(bb)), "b values(first 16)", np.unique(bb)[:16], " k1 = b%64:", np.unique(bb % 64))
        k1s = np.unique(bb % 64)
        for k1 in k1s[:2]:
            sel = bb % 64 == k1
            print("   k1", k1, "k2 values", np.unique(bb[sel] // 64)[:20], "count", len(np.unique(bb[sel] // 64)))
