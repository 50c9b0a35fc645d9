"""Developer check: product sources under the CUDA-thread emulator vs the CPU oracle (tests only)."""
import sys, ctypes, numpy as np
sys.path.insert(0, '/root/repo/tests')
import fftlibs as fl
sim = fl.Lib(ctypes.CDLL('/root/repo/tools/sim/libcfftpack_sim.so'))
orc = fl.Lib(fl.oracle(), 'orc_')
fams = sys.argv[1].split(',') if len(sys.argv) > 1 else fl.FAMILIES
sizes = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [2,3,4,5,6,7,8,9,10,12,15,16,20,25,30,32,49,60,64,77,100,121,128,256,1000,1001]
bad = 0
for fam in fams:
    for n in sizes:
        wa, i1 = sim.init(fam, n); wb, i2 = orc.init(fam, n)
        if not (i1 == i2 == 0 and np.array_equal(wa, wb)):
            print("WSAVE MISMATCH", fam, n, i1, i2); bad += 1
        for d in 'fb':
            x = fl.rand_input(fam, n, 17 * n + 3)
            a, ia = sim.run1(fam, d, n, x)
            b, ib = orc.run1(fam, d, n, x)
            e = fl.rel_l2(a, b)
            ok = ia == ib == 0 and e <= fl.tol(n)
            if not ok:
                bad += 1
            print(f"{fam}1{d} n={n:5d} ier={ia},{ib} err={e:.2e} {'ok' if ok else 'FAIL'}")
            lot = 5
            x = fl.rand_input(fam, n * lot, 7 * n + 1)
            a, ia = sim.runm(fam, d, lot, n, n, 1, x)
            b, ib = orc.runm(fam, d, lot, n, n, 1, x)
            e = fl.rel_l2(a, b)
            ok = ia == ib == 0 and e <= fl.tol(n)
            if not ok: bad += 1; print(f"   {fam}m{d} cols n={n} lot={lot} ier={ia},{ib} err={e:.2e} FAIL")
            a, ia = sim.runm(fam, d, lot, 1, n, lot, x)
            b, ib = orc.runm(fam, d, lot, 1, n, lot, x)
            e = fl.rel_l2(a, b)
            ok = ia == ib == 0 and e <= fl.tol(n)
            if not ok: bad += 1; print(f"   {fam}m{d} rows n={n} lot={lot} ier={ia},{ib} err={e:.2e} FAIL")
print("BAD", bad)
