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


L2_ALGOS = {"fft": 1, "dct1": 4, "dct": 5, "dst1": 7, "dst": 8}  # cfftintern.h numbering (the oracle's orc_l2_create)


def bind_l2(lib):
    """argtypes for the reference's object API (cfftpack/cfftpack.h) as exported by `lib`"""
    vp = ctypes.c_void_p
    for name in ("fft_create", "rfft_create", "dct_create", "dct1_create", "dst_create", "dst1_create"):
        getattr(lib, name).restype, getattr(lib, name).argtypes = vp, [ctypes.c_int]
    lib.fft2_create.restype, lib.fft2_create.argtypes = vp, [ctypes.c_int] * 2
    for fam in ("fft", "fft2", "dct", "dct1", "dst", "dst1"):
        for d in ("forward", "inverse"):
            getattr(lib, f"{fam}_{d}").argtypes = [vp, vp]
    lib.fft_ortho.argtypes, lib.fft_stride.argtypes, lib.fft_free.argtypes = [vp, ctypes.c_bool], [vp, ctypes.c_int], [vp]
    lib.rfft_forward.argtypes = lib.rfft_inverse.argtypes = [vp, vp, vp]
    if hasattr(lib, "cfb200_fft_batch"):
        lib.cfb200_fft_batch.argtypes = [vp, ctypes.c_int]
    return lib


def bind_l2_oracle():
    lib, vp = oracle(), ctypes.c_void_p
    lib.orc_l2_create.restype, lib.orc_l2_create.argtypes = vp, [ctypes.c_int] * 3
    lib.orc_l2_forward.argtypes = lib.orc_l2_inverse.argtypes = [vp, vp]
    lib.orc_l2_ortho.argtypes = lib.orc_l2_stride.argtypes = [vp, ctypes.c_int]
    lib.orc_l2_free.argtypes = [vp]
    lib.orc_l2_rfft_forward.argtypes = lib.orc_l2_rfft_inverse.argtypes = [vp, vp, vp]
    return lib


L2_FAMILY = {"fft": "cfft", "dct": "cosq", "dct1": "cost", "dst": "sinq", "dst1": "sint"}


def l2_compare(lib, sizes, exact_codes=True, tol_=1e-13, noise=False):
    """`lib` (reference names) against the oracle's restatement of cfftpack.c: every family, both directions, with and
    without fft_ortho, stride 1 and 2 (the latter is an error upstream for all but the DCT); rfft repack; fft2."""
    S, O = bind_l2(lib), bind_l2_oracle()
    worst = 0.0
    for n in sizes:
        for nm, algo in L2_ALGOS.items():
            for ortho in (0, 1):
                for inc in (1, 2):
                    fs, fo = getattr(S, nm + "_create")(n), O.orc_l2_create(algo, n, 0)
                    assert (fs is None) == (fo is None), (nm, n)
                    if fs is None:
                        continue
                    S.fft_ortho(fs, bool(ortho)); O.orc_l2_ortho(fo, ortho)
                    S.fft_stride(fs, inc); O.orc_l2_stride(fo, inc)
                    x = np.random.default_rng(n + algo).uniform(-1, 1, n * inc * (2 if algo == 1 else 1))
                    for d in ("forward", "inverse"):
                        a, b = x.copy(), x.copy()
                        ra, rb = getattr(S, f"{nm}_{d}")(fs, P(a)), getattr(O, "orc_l2_" + d)(fo, P(b))
                        if exact_codes:
                            assert ra == rb, (nm, n, ortho, inc, d, ra, rb)
                        else:
                            assert (ra == 0) == (rb == 0), (nm, n, ortho, inc, d, ra, rb)
                        if ra == 0:
                            e = rel_l2(a, b)
                            assert e <= tol_ + (ref_noise(L2_FAMILY[nm], n) if noise else 0.0), (nm, n, ortho, inc, d, e)
                            worst = max(worst, e)
                    S.fft_free(fs); O.orc_l2_free(fo)
        fs, fo = S.rfft_create(n), O.orc_l2_create(2, n, 0)
        x = np.random.default_rng(n).uniform(-1, 1, n)
        a, b = np.full(n + 3, 7.0), np.full(n + 3, 7.0)
        assert S.rfft_forward(fs, P(x), P(a)) == O.orc_l2_rfft_forward(fo, P(x), P(b)) == 0
        used = n + 2 if n % 2 == 0 else n + 1
        assert rel_l2(a[:used], b[:used]) <= tol_ and np.array_equal(a[used:], b[used:]), n
        ya, yb = np.zeros(n), np.zeros(n)
        assert S.rfft_inverse(fs, P(a), P(ya)) == O.orc_l2_rfft_inverse(fo, P(b), P(yb)) == 0
        assert rel_l2(ya, yb) <= tol_ and rel_l2(ya, x) <= tol_, n
        S.fft_free(fs); O.orc_l2_free(fo)
    fs, fo = S.fft2_create(8, 6), O.orc_l2_create(3, 8, 6)
    c = np.random.default_rng(1).uniform(-1, 1, 96)
    for d in ("forward", "inverse"):
        a, b = c.copy(), c.copy()
        assert getattr(S, "fft2_" + d)(fs, P(a)) == getattr(O, "orc_l2_" + d)(fo, P(b)) == 0 and rel_l2(a, b) <= tol_
    S.fft_free(fs); O.orc_l2_free(fo)
    return worst


def l2_batch_check(lib, n=60, lot=5):
    """cfb200_fft_batch: lot sequences back to back == the same handle called per sequence"""
    S = bind_l2(lib)
    for nm, algo in L2_ALGOS.items():
        w = n * (2 if algo == 1 else 1)
        f = getattr(S, nm + "_create")(n)
        S.fft_ortho(f, True)
        x = np.random.default_rng(3).uniform(-1, 1, w * lot)
        a = x.copy()
        for o in range(lot):
            seg = a[o * w:(o + 1) * w].copy()
            assert getattr(S, nm + "_forward")(f, P(seg)) == 0
            a[o * w:(o + 1) * w] = seg
        S.cfb200_fft_batch(f, lot)
        b = x.copy()
        assert getattr(S, nm + "_forward")(f, P(b)) == 0 and rel_l2(a, b) <= 1e-14, nm
        S.fft_free(f)
    f = S.rfft_create(n)
    S.cfb200_fft_batch(f, lot)
    x = np.random.default_rng(4).uniform(-1, 1, n * lot)
    out, back = np.zeros(lot * (n + 2)), np.zeros(n * lot)
    assert S.rfft_forward(f, P(x), P(out)) == 0 and S.rfft_inverse(f, P(out), P(back)) == 0
    assert rel_l2(back, x) <= 1e-14
    S.fft_free(f)
