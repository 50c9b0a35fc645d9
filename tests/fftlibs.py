"""ctypes access to the three FFTPACK-ABI libraries used by the tests.

  product : cfftpack_b200/libcfftpack_b200.so   (CUDA, symbols `cfft1f_` ...)
  oracle  : oracle/liboracle.so                 (CPU restatement, `orc_cfft1f_` ...)
  ref     : oracle/_ref/libfftpack_ref.so       (unmodified reference, only if prebuilt)

All three share the reference's Fortran-style signatures, so one caller
serves them all.  TEST INFRASTRUCTURE ONLY.
"""
import ctypes
import math
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
I = ctypes.c_int
_cache = {}


def il2(n):
    return int(math.log(float(n)) / math.log(2.0))  # fftpack.c:2221 literal


def _load(path):
    if path not in _cache:
        _cache[path] = ctypes.CDLL(path) if os.path.exists(path) else None
    return _cache[path]


def oracle():
    return _load(os.path.join(ROOT, "oracle", "liboracle.so"))


def ref():
    return _load(os.path.join(ROOT, "oracle", "_ref", "libfftpack_ref.so"))


def naive_ref():
    return _load(os.path.join(ROOT, "oracle", "_ref", "libnaive_ref.so"))


def product():
    lib = _load(os.path.join(ROOT, "cfftpack_b200", "libcfftpack_b200.so"))
    if lib is not None:
        lib.cfb200_last_error.restype = ctypes.c_char_p
        lib.cfb200_launch_count.restype = ctypes.c_ulonglong
    return lib


def sim():
    """product sources compiled against the CUDA-thread emulator (tools/sim) -- CPU-side tests only"""
    return _load(os.path.join(ROOT, "tools", "sim", "libcfftpack_sim.so"))


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


FAMILIES = ("cfft", "rfft", "cost", "sint", "cosq", "sinq")


def lensav(fam, n):
    if fam == "cfft":
        return 2 * n + il2(n) + 4
    if fam == "rfft":
        return n + il2(n) + 4
    if fam == "sint":
        return n // 2 + n + il2(n) + 4
    return 2 * n + il2(n) + 4


def lenwrk(fam, n, lot=None):
    if lot is None:
        return {"cfft": 2 * n, "rfft": n, "cost": max(n - 1, 1), "sint": 2 * n + 2, "cosq": n, "sinq": n}[fam]
    return {"cfft": 2 * lot * n, "rfft": lot * n, "cost": lot * (n + 1), "sint": lot * (2 * n + 4),
            "cosq": lot * n, "sinq": lot * n}[fam]


class Lib:
    """Uniform caller. prefix '' for product/ref, 'orc_' for the oracle."""

    def __init__(self, lib, prefix=""):
        self.lib, self.prefix = lib, prefix

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def init(self, fam, n, multi=False, lensav_override=None):
        ls = lensav(fam, n) if lensav_override is None else lensav_override
        ws = np.zeros(max(ls, 1) + 8)
        ier = I(-1)
        self.fn(fam + ("mi_" if multi else "1i_"))(ctypes.byref(I(n)), P(ws), ctypes.byref(I(ls)), ctypes.byref(ier))
        return ws, ier.value

    def run1(self, fam, d, n, x, inc=1, ws=None, lenx=None, lensav_=None, lenwrk_=None):
        """single transform, in place on a copy; returns (y, ier)"""
        if ws is None:
            ws, ier0 = self.init(fam, n)
            assert ier0 == 0
        y = np.array(x, copy=True)
        lw = lenwrk(fam, n) if lenwrk_ is None else lenwrk_
        wk = np.zeros(max(lw, 1) + 8)
        ier = I(-1)
        self.fn(fam + "1" + d + "_")(
            ctypes.byref(I(n)), ctypes.byref(I(inc)), P(y), ctypes.byref(I(len(y) if lenx is None else lenx)), P(ws),
            ctypes.byref(I(lensav(fam, n) if lensav_ is None else lensav_)), P(wk), ctypes.byref(I(lw)),
            ctypes.byref(ier))
        return y, ier.value

    def runm(self, fam, d, lot, jump, n, inc, x, ws=None, lenx=None, lensav_=None, lenwrk_=None, work=True):
        """batched transform, in place on a copy; returns (y, ier)"""
        if ws is None:
            ws, ier0 = self.init(fam, n, multi=True)
            assert ier0 == 0
        y = np.array(x, copy=True)
        lw = lenwrk(fam, n, lot) if lenwrk_ is None else lenwrk_
        wk = np.zeros(max(lw, 1) + 8) if work else np.zeros(8)
        ier = I(-1)
        self.fn(fam + "m" + d + "_")(
            ctypes.byref(I(lot)), ctypes.byref(I(jump)), ctypes.byref(I(n)), ctypes.byref(I(inc)), P(y),
            ctypes.byref(I(len(y) if lenx is None else lenx)), P(ws),
            ctypes.byref(I(lensav(fam, n) if lensav_ is None else lensav_)), P(wk), ctypes.byref(I(lw)),
            ctypes.byref(ier))
        return y, ier.value

    def init2(self, l, m):
        ls = 2 * l + il2(l) + 2 * m + il2(m) + 8
        ws = np.zeros(ls + 8)
        ier = I(-1)
        self.fn("cfft2i_")(ctypes.byref(I(l)), ctypes.byref(I(m)), P(ws), ctypes.byref(I(ls)), ctypes.byref(ier))
        return ws, ls, ier.value

    def run2(self, d, ldim, l, m, c, lenwrk_=None):
        ws, ls, ier0 = self.init2(l, m)
        assert ier0 == 0
        y = np.array(c, copy=True)
        lw = 2 * l * m if lenwrk_ is None else lenwrk_
        wk = np.zeros(8)  # never touched by the product; oracle/ref need the real thing
        if self.prefix == "orc_" or self.lib is ref():
            wk = np.zeros(lw + 8)
        ier = I(-1)
        self.fn("cfft2" + d + "_")(ctypes.byref(I(ldim)), ctypes.byref(I(l)), ctypes.byref(I(m)), P(y), P(ws),
                                   ctypes.byref(I(ls)), P(wk), ctypes.byref(I(lw)), ctypes.byref(ier))
        return y, ier.value


    def init2r(self, l, m):
        """rfft2i_ (fftpack.c:13454): wsave = rfft plan of l, cfft plan of m, rfft plan of m"""
        ls = l + il2(l) + 4 + 2 * m + il2(m) + 4 + m + il2(m) + 4
        ws = np.zeros(ls + 8)
        ier = I(-1)
        self.fn("rfft2i_")(ctypes.byref(I(l)), ctypes.byref(I(m)), P(ws), ctypes.byref(I(ls)), ctypes.byref(ier))
        return ws, ls, ier.value

    def run2r(self, d, ldim, l, m, r, lenwrk_=None, lensav_=None):
        """rfft2f_/rfft2b_ on a real column-major r(ldim, m).  The reference uses r itself as the complex pass's
        scratch (fftpack.c:13407), so rows l..ldim-1 come back clobbered there: compare rows < l only."""
        ws, ls, ier0 = self.init2r(l, m)
        assert ier0 == 0
        y = np.array(r, dtype=np.float64, copy=True)
        lw = (l + 1) * m if lenwrk_ is None else lenwrk_
        wk = np.zeros(8)
        if self.prefix == "orc_" or self.lib is ref():
            wk = np.zeros(max(lw, (l + 2) * m) + 8)
        ier = I(-1)
        self.fn("rfft2" + d + "_")(ctypes.byref(I(ldim)), ctypes.byref(I(l)), ctypes.byref(I(m)), P(y), P(ws),
                                   ctypes.byref(I(ls if lensav_ is None else lensav_)), P(wk), ctypes.byref(I(lw)),
                                   ctypes.byref(ier))
        return y, ier.value


def rows2(a, ldim, l, m):
    """the addressed part r(0:l, 0:m) of a column-major array with leading dimension ldim"""
    a = np.asarray(a)
    full = np.zeros(ldim * m, a.dtype)
    full[: min(len(a), ldim * m)] = a[: ldim * m]
    return full.reshape(m, ldim)[:, :l]


def rand_input(fam, count, seed):
    rng = np.random.default_rng(seed)
    if fam == "cfft":
        return (rng.uniform(-1, 1, count) + 1j * rng.uniform(-1, 1, count)).astype(np.complex128)
    return rng.uniform(-1, 1, count)


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def tol(n):
    """north_star parity bar: relative L2 <= 1e-12 * log2(N) (FP64)"""
    return 1e-12 * max(math.log2(max(n, 2)), 1.0)


def vec(fam, kind, n):
    if kind == "ramp":
        v = np.arange(n, dtype=np.float64) + 1.0
    elif kind == "frac":
        v = np.arange(n, dtype=np.float64) / n + 1.0
    else:
        return rand_input(fam, n, 1000 + n)
    return v.astype(np.complex128) if fam == "cfft" else v


GOLDEN = os.path.join(ROOT, "tests", "golden", "fftpack_golden.npz")


def golden():
    if "golden" not in _cache:
        _cache["golden"] = np.load(GOLDEN)
    return _cache["golden"]


def underlying(fam, n):
    """length of the FFT that does the work: N-1 for cost, N+1 for sint, N otherwise (SURVEY 8(a) note 2)"""
    return {"cost": n - 1, "sint": n + 1}.get(fam, n)


def max_generic_factor(n):
    """largest factor > 5 in FFTPACK's factorisation (4, 2, 3, 5, then 7, 9, 11, ...) or 0"""
    best, p = 0, 2
    while n > 1 and p * p <= n:
        while n % p == 0:
            n //= p
            best = max(best, p)
        p += 1
    best = max(best, n if n > 1 else 0)
    return best if best > 5 else 0


def ref_noise(fam, n):
    """extra allowance when comparing against the *reference*: its generic-prime real pass builds the
    roots of unity by recurrence (fftpack.c:12784-12805), so its own error grows ~ 4e-15 * prime."""
    return 5e-15 * max_generic_factor(underlying(fam, n)) if fam != "cfft" else 0.0


def vargamma_ref():
    """the reference's test/vargamma.c as a library (oracle/Makefile); None on the GPU box"""
    return _load(os.path.join(ROOT, "oracle", "_ref", "libvargamma_ref.so"))


OPTION_CASES = [  # (S, K, sigma, theta, kappa, t, r, call, black_scholes); first row = test/vargamma.c:108-118
    (100.0, 98.0, 0.12, -0.14, 0.2, 1.0, 0.05, 1, 1), (100.0, 98.0, 0.12, -0.14, 0.2, 1.0, 0.05, 1, 0),
    (100.0, 98.0, 0.12, -0.14, 0.2, 1.0, 0.05, 0, 1), (100.0, 98.0, 0.12, -0.14, 0.2, 1.0, 0.05, 0, 0),
    (90.0, 100.0, 0.2, -0.1, 0.3, 0.5, 0.03, 1, 0), (110.0, 100.0, 0.3, -0.2, 0.1, 2.0, 0.01, 0, 0),
    (100.0, 105.0, 0.12, -0.14, 0.2, 1.0, 0.05, 1, 1), (100.0, 98.0, 0.25, 0.05, 0.5, 0.25, 0.0, 1, 0),
]


def option_oracle(n, case):
    lib = oracle()
    lib.orc_conv_bsvg_option.restype = ctypes.c_double
    lib.orc_conv_bsvg_option.argtypes = [ctypes.c_int] + [ctypes.c_double] * 7 + [ctypes.c_int] * 2
    return lib.orc_conv_bsvg_option(n, *case)


def option_product(lib, n, cases):
    cols = [np.ascontiguousarray([c[k] for c in cases], dtype=np.float64) for k in range(7)]
    flags = np.ascontiguousarray([c[7] | (c[8] << 1) for c in cases], dtype=np.int32)
    val = np.zeros(len(cases))
    ier = I(-9)
    N = lib.cfb200_option_convolution(I(len(cases)), I(n), *(P(c) for c in cols), P(flags), P(val), ctypes.byref(ier))
    return val, N, ier.value
