"""Small run that touches every kernel once (for compute-sanitizer): checks results against the oracle too."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fftlibs as fl
P, O = fl.Lib(fl.product()), fl.Lib(fl.oracle(), "orc_")
cases = [("cfft", 256, 40, 256, 1), ("cfft", 4096, 5, 4096, 1), ("rfft", 512, 37, 512, 1), ("rfft", 256, 9, 259, 1),
         ("cfft", 100, 40, 100, 1), ("cfft", 60, 33, 1, 33), ("cfft", 77, 5, 80, 1), ("cosq", 100, 21, 100, 1),
         ("cost", 101, 8, 1, 8), ("sint", 64, 7, 64, 1), ("sinq", 50, 6, 50, 1), ("rfft", 99, 10, 99, 1),
         ("cfft", 16384, 2, 16384, 1), ("cfft", 4096, 9, 1, 9), ("cfft", 9009, 2, 9009, 1), ("rfft", 20000, 2, 20000, 1),
         ("cost", 3, 5, 3, 1), ("cosq", 2, 5, 2, 1)]
bad = 0
for fam, n, lot, jump, inc in cases:
    span = (lot - 1) * jump + (n - 1) * inc + 1
    x = fl.rand_input(fam, span, n + lot)
    for d in "fb":
        a, ia = P.runm(fam, d, lot, jump, n, inc, x, lenx=span, work=False)
        b, ib = O.runm(fam, d, lot, jump, n, inc, x, lenx=span)
        e = fl.rel_l2(a, b)
        ok = ia == ib == 0 and e <= fl.tol(n)
        bad += not ok
        print(fam, d, n, lot, jump, inc, ia, ib, f"{e:.2e}", "ok" if ok else "FAIL", flush=True)
c = fl.rand_input("cfft", 70 * 48, 3)
for d in "fb":
    a, ia = P.run2(d, 70, 64, 48, c); b, ib = O.run2(d, 70, 64, 48, c)
    ok = ia == ib == 0 and fl.rel_l2(a, b) <= fl.tol(64 * 48); bad += not ok
    print("cfft2", d, ia, ib, "ok" if ok else "FAIL")
print("BAD", bad)
sys.exit(1 if bad else 0)
