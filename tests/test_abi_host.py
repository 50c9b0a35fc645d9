"""CPU-side checks of the product library's C ABI and host logic (no kernel runs here).

 - libcfftpack_b200.so loads and exports every symbol include/cfftpack_b200.h and include/cfftpack_b200_l2.h declare;
 - the *i_ routines fill wsave bit-for-bit like the reference (golden fixture) -- host code;
 - ier codes agree with the oracle for every argument error (all return before any GPU work);
 - without a GPU the library fails loudly (ier = -1) instead of falling back to a CPU path.
"""
import ctypes
import os
import re

import numpy as np
import pytest

import fftlibs as fl

PROD = fl.Lib(fl.product())
ORC = fl.Lib(fl.oracle(), "orc_")
G = fl.golden()
I = ctypes.c_int


def _declared_symbols():
    src = open(os.path.join(fl.ROOT, "include", "cfftpack_b200.h")).read()
    names = set(re.findall(r"\b(?:int|void|const char \*|unsigned long long)\s*\*?\s*((?:cfft|rfft|cfb200)\w+)\s*\(", src))
    for fam in re.findall(r"CFB200_DECL_TRIG\((\w+)\)", src):
        if fam != "name":
            names |= {f"{fam}{v}_" for v in ("1i", "1f", "1b", "mi", "mf", "mb")}
    # the object API header (cfftpack.h names): every prototype "type name(" at the start of a line
    l2 = open(os.path.join(fl.ROOT, "include", "cfftpack_b200_l2.h")).read()
    names |= set(re.findall(r"^(?:int|void|fft_t \*)\s*(\w+)\(", l2, flags=re.M))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    names = _declared_symbols()
    assert len(names) >= 9 + 6 + 24 + 8 + 3 + 1 + 24, names
    for n in names:
        assert hasattr(fl.product(), n), n
    # the north_star list, explicitly
    for n in ("cfft1i_ cfft1f_ cfft1b_ cfftmi_ cfftmf_ cfftmb_ rfftmi_ rfftmf_ rfftmb_ cfft2i_ cfft2f_ cfft2b_ "
              "cost1i_ cost1f_ cost1b_ sint1i_ sint1f_ sint1b_ cosq1i_ cosq1f_ cosq1b_ rfft1i_ rfft1f_ rfft1b_ "
              "costmf_ sintmf_ cosqmf_ cosqmb_ cosqmi_ sinq1f_ sinqmf_").split():
        assert hasattr(fl.product(), n), n


def test_product_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", os.path.join(fl.ROOT, "cfftpack_b200", "libcfftpack_b200.so")], capture_output=True,
                         text=True).stdout
    assert "oracle" not in out and "fftpack_ref" not in out and "sim" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", os.path.join(fl.ROOT, "cfftpack_b200", "libcfftpack_b200.so")],
                          capture_output=True, text=True).stdout
    assert "orc_" not in syms and "cfbsim" not in syms


@pytest.mark.parametrize("fam", fl.FAMILIES)
def test_wsave_tables_bitwise_vs_reference(fam):
    for key in G.files:
        if key.startswith(f"wsave_{fam}_"):
            n = int(key.split("_")[2])
            for multi in (False, True):
                ws, ier = PROD.init(fam, n, multi=multi)
                assert ier == 0
                assert np.array_equal(ws[: fl.lensav(fam, n)], G[key]), (fam, n)


def test_cfft2i_matches_oracle():
    for (l, m) in ((8, 6), (16, 16), (5, 12), (1, 7), (64, 1)):
        a, ls, ia = PROD.init2(l, m)
        b, _, ib = ORC.init2(l, m)
        assert ia == ib == 0 and np.array_equal(a, b)


@pytest.mark.parametrize("fam", fl.FAMILIES)
def test_ier_codes_match_oracle(fam):
    n, lot = 12, 3
    x1 = fl.rand_input(fam, n, 1)
    xm = fl.rand_input(fam, n * lot, 2)
    for d in "fb":
        for kw in (dict(lenx=n - 1), dict(lensav_=fl.lensav(fam, n) - 1), dict(lenwrk_=fl.lenwrk(fam, n) - 1)):
            a, ia = PROD.run1(fam, d, n, x1, **kw)
            b, ib = ORC.run1(fam, d, n, x1, **kw)
            assert ia == ib != 0, (fam, d, kw, ia, ib)
            assert np.array_equal(a, x1), "data must be untouched on error"
        for kw in (dict(lenx=n * lot - 1), dict(lensav_=fl.lensav(fam, n) - 1),
                   dict(lenwrk_=fl.lenwrk(fam, n, lot) - 1)):
            a, ia = PROD.runm(fam, d, lot, n, n, 1, xm, **kw)
            b, ib = ORC.runm(fam, d, lot, n, n, 1, xm, **kw)
            assert ia == ib != 0, (fam, d, kw, ia, ib)
            assert np.array_equal(a, xm)
        # inconsistent strides -> 4 (xercon_, fftpack.c:15210): jump = 2, inc = 3 overlap for n = 12, lot = 3... use lcm rule
        big = fl.rand_input(fam, 200, 3)
        a, ia = PROD.runm(fam, d, 4, 2, 6, 3, big)
        b, ib = ORC.runm(fam, d, 4, 2, 6, 3, big)
        # sinqmb_ reports every argument error as 20 (upstream fall-through, fftpack.c:14151-14179)
        assert ia == ib == (20 if (fam, d) == ("sinq", "b") else 4), (fam, d, ia, ib)
    ws, ier = PROD.init(fam, n, lensav_override=fl.lensav(fam, n) - 1)
    assert ier == 2


def test_cfft2_ier_codes():
    c = fl.rand_input("cfft", 11 * 6, 4)
    for d in "fb":
        y, ier = PROD.run2(d, 7, 8, 6, c)  # l > ldim
        assert ier == 5 and np.array_equal(y, c)
        y, ier = PROD.run2(d, 8, 8, 6, c, lenwrk_=2 * 8 * 6 - 1)
        assert ier == 3 and np.array_equal(y, c)


def test_rfft2_init_and_ier_codes():
    for (l, m) in ((8, 6), (7, 5), (1, 4), (32, 32)):
        a, ls, ia = PROD.init2r(l, m)
        b, _, ib = ORC.init2r(l, m)
        assert ia == ib == 0 and np.array_equal(a, b)
    assert np.array_equal(PROD.init2r(8, 6)[0][: len(G["wsave_rfft2_8x6"])], G["wsave_rfft2_8x6"])
    r = fl.rand_input("rfft", 11 * 6, 4)
    for d in "fb":
        for kw, want in ((dict(lenwrk_=9 * 6 - 1), 3), (dict(lensav_=8 + 3 + 4 + 12 + 2 + 4 + 6 + 2 + 4 - 1), 2)):
            y, ier = PROD.run2r(d, 8, 8, 6, r, **kw)
            assert ier == want == ORC.run2r(d, 8, 8, 6, r, **kw)[1] and np.array_equal(y, r)
        y, ier = PROD.run2r(d, 7, 8, 6, r)  # ldim < l
        assert ier == 5 and np.array_equal(y, r)


def test_option_convolution_argument_checks():
    val, N, ier = fl.option_product(fl.product(), 0, fl.OPTION_CASES[:2])
    assert ier == 1 and not val.any()
    if not fl.has_gpu():
        val, N, ier = fl.option_product(fl.product(), 1001, fl.OPTION_CASES[:2])
        assert (N, ier) == (1024, -1) and not val.any()  # sizes like cfftextra.c:42-46; no CPU path


def test_degenerate_arguments_never_touch_data_or_crash():
    """non-positive n / lot (the reference loops zero times or is undefined) and non-positive strides"""
    lib = fl.product()
    x = np.random.default_rng(1).uniform(-1, 1, 256)
    ws, wk = np.zeros(600), np.zeros(8)
    for name, (lot, jump, n, inc), want in (("cfftmf_", (-1, 8, 8, 1), 0), ("cfftmf_", (0, 8, 8, 1), 0), ("cfftmb_", (2, 8, 0, 1), 0),
                                            ("rfftmb_", (0, 8, 8, 1), 0), ("costmf_", (2, 8, -1, 1), 0), ("sinqmb_", (-2, 8, 8, 1), 0),
                                            ("cfftmf_", (2, 8, 8, -1), 4), ("rfftmf_", (2, -8, 8, 1), 4), ("cosqmf_", (2, 8, 8, 0), 4)):
        y, ier = x.copy(), I(-7)
        getattr(lib, name)(ctypes.byref(I(lot)), ctypes.byref(I(jump)), ctypes.byref(I(n)), ctypes.byref(I(inc)), fl.P(y),
                           ctypes.byref(I(128)), fl.P(ws), ctypes.byref(I(600)), fl.P(wk), ctypes.byref(I(100000)), ctypes.byref(ier))
        assert ier.value == want and np.array_equal(y, x), (name, lot, jump, n, inc, ier.value)
    for name, n, inc, want in (("cfft1f_", 0, 1, 0), ("rfft1b_", -4, 1, 0), ("cost1f_", 0, 1, 0), ("cfft1b_", 8, -1, 1), ("sint1f_", 8, 0, 1)):
        y, ier = x.copy(), I(-7)
        getattr(lib, name)(ctypes.byref(I(n)), ctypes.byref(I(inc)), fl.P(y), ctypes.byref(I(128)), fl.P(ws), ctypes.byref(I(600)),
                           fl.P(wk), ctypes.byref(I(1000)), ctypes.byref(ier))
        assert ier.value == want and np.array_equal(y, x), (name, n, inc, ier.value)
    ier = I(-7)
    lib.cfft1i_(ctypes.byref(I(0)), fl.P(ws), ctypes.byref(I(600)), ctypes.byref(ier))
    assert ier.value == 0 and not ws.any()
    for fn in ("cfft2f_", "rfft2b_"):
        y, ier = x.copy(), I(-7)
        getattr(lib, fn)(ctypes.byref(I(4)), ctypes.byref(I(0)), ctypes.byref(I(4)), fl.P(y), fl.P(ws), ctypes.byref(I(600)),
                         fl.P(wk), ctypes.byref(I(1000)), ctypes.byref(ier))
        assert ier.value == 0 and np.array_equal(y, x)


def test_length_one_is_a_no_op():
    for fam in fl.FAMILIES:
        x = fl.rand_input(fam, 1, 5)
        for d in "fb":
            y, ier = PROD.run1(fam, d, 1, x)
            assert ier == 0 and np.array_equal(y, x)


@pytest.mark.skipif(fl.has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_cpu_fallback():
    x = fl.rand_input("cfft", 16, 6)
    y, ier = PROD.run1("cfft", "f", 16, x)
    assert ier == -1
    assert np.array_equal(y, x)
    assert b"no CUDA device" in fl.product().cfb200_last_error()
