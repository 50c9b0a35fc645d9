import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import cfftpack_b200 as cb
a = int(sys.argv[1]); n = 1 << a
g = torch.Generator(device="cuda").manual_seed(a)
x0 = torch.view_as_complex(torch.rand(n, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
aL = a // 2 if a >= 24 else (9 if a == 21 else 10); aM = a - aL; L, Mm = 1 << aL, 1 << aM
# 1. the pre-existing in-place path with the shapes of step A and step B
for name, nn, lot, jump, inc in (("stepA-like", Mm, L, 1, L), ("stepB-like", L, Mm, L, 1)):
    plan = cb.Plan("cfft", nn)
    outs = []
    for rep in range(4):
        x = x0.clone() if rep % 2 else x0.clone()
        assert plan.multi("f", x.data_ptr(), lot, jump, inc, n) == 0, cb.last_error()
        cb.synchronize()
        outs.append(x)
    print(name, [bool(torch.equal(torch.view_as_real(o), torch.view_as_real(outs[0]))) for o in outs], flush=True)
# 2. the long transform, same buffer and fresh buffers
plan = cb.Plan("cfft", n)
xs = x0.clone()
outs = []
for rep in range(4):
    xs.copy_(x0)
    assert plan.multi("f", xs.data_ptr(), 1, n, 1, n) == 0
    cb.synchronize()
    outs.append(xs.clone())
print("long same buffer", [int((torch.view_as_real(o) != torch.view_as_real(outs[0])).sum()) for o in outs], flush=True)
outs2 = []
keep = []
for rep in range(4):
    x = x0.clone(); keep.append(x)
    assert plan.multi("f", x.data_ptr(), 1, n, 1, n) == 0
    cb.synchronize()
    outs2.append(x)
print("long fresh buffers", [int((torch.view_as_real(o) != torch.view_as_real(outs[0])).sum()) for o in outs2], [hex(o.data_ptr()) for o in outs2], flush=True)
