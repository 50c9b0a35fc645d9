"""Pins the CPU oracle (oracle/fftpack_oracle.c) to the reference.

Three anchors:
  1. committed golden vectors produced by the unmodified reference
     (tests/golden/make_golden.py) -- runs everywhere, incl. the GPU box;
  2. the reference's own O(N^2) definitions (test/naivepack.c), stored in the
     same fixture, and the oracle's restatement of them;
  3. the compiled reference itself (oracle/_ref), when it is present.
No GPU needed.
"""
import numpy as np
import pytest

import fftlibs as fl

G = fl.golden()
ORC = fl.Lib(fl.oracle(), "orc_")


def _cases():
    for key in G.files:
        if key.endswith("/y") and key.split("_")[0] in fl.FAMILIES and key.count("_") == 3:
            fam, d, n, kind = key[:-2].split("_")
            yield fam, d, int(n), kind


@pytest.mark.parametrize("fam", fl.FAMILIES)
def test_wsave_tables_bitwise(fam):
    """*1i_ / *mi_ fill wsave exactly like the reference (factor lists, twiddles, trig tables)."""
    for key in G.files:
        if key.startswith(f"wsave_{fam}_"):
            n = int(key.split("_")[2])
            for multi in (False, True):
                ws, ier = ORC.init(fam, n, multi=multi)
                assert ier == 0
                assert np.array_equal(ws[: fl.lensav(fam, n)], G[key]), (fam, n)


@pytest.mark.parametrize("fam", fl.FAMILIES)
def test_oracle_vs_golden(fam):
    worst = 0.0
    for f, d, n, kind in _cases():
        if f != fam:
            continue
        x = fl.vec(fam, kind, n)
        y, ier = ORC.run1(fam, d, n, x)
        assert ier == 0
        want = G[f"{fam}_{d}_{n}_{kind}/y"]
        if fam == "cfft":
            # the complex passes restate the reference operation by operation
            assert np.array_equal(y, want), (fam, d, n, kind)
        else:
            e = fl.rel_l2(y, want)
            worst = max(worst, e)
            # 999 = 3^3*37, 1002 = 2*3*167, 211: the reference builds prime roots by
            # recurrence (fftpack.c:12784-12805), so it is the less accurate side there
            assert e <= 2e-14 + fl.ref_noise(fam, n), (fam, d, n, kind, e)
    print(fam, "worst rel-L2 vs golden", worst)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 16, 30, 32, 60, 64])
def test_oracle_vs_reference_naive(n):
    """test/testall.c:44-59 bar: abs error <= 1e-13 against the naive definitions (N = 2, 32, 60 there)."""
    x = fl.vec("rfft", "frac", n)
    for fam in ("cost", "sint", "cosq", "sinq"):
        for d in "fb":
            if fam == "cost" and n < 2:
                continue
            y, ier = ORC.run1(fam, d, n, x)
            assert ier == 0
            want = G[f"naive_{fam}_{d}_{n}/y"]
            scale = 1.0  # FFTPACK's backward transforms are the unscaled naive_dct2/dst2/dct1/dst1
            assert np.max(np.abs(y - scale * want)) <= 1e-13 * max(1.0, np.max(np.abs(want)) * scale), (fam, d, n)
    xc = G[f"naive_cfft_f_{n}/x"]
    yf, _ = ORC.run1("cfft", "f", n, xc)
    yb, _ = ORC.run1("cfft", "b", n, xc)
    assert np.max(np.abs(yf - G[f"naive_cfft_f_{n}/y"])) <= 1e-13 * np.max(np.abs(xc))
    assert np.max(np.abs(yb - G[f"naive_cfft_b_{n}/y"])) <= 1e-13 * np.max(np.abs(xc)) * n


@pytest.mark.parametrize("n", [2, 3, 5, 8, 12, 30, 60, 64, 100])
def test_oracle_own_naive(n):
    """the oracle's restatement of naivepack agrees with the oracle transforms"""
    lib = fl.oracle()
    import ctypes
    x = fl.rand_input("rfft", n, 5 + n)
    y = np.zeros(n)
    for fam in ("cost", "sint", "cosq", "sinq"):
        for d, fwd in (("f", 1), ("b", 0)):
            getattr(lib, "orc_naive_" + fam)(ctypes.c_int(n), fl.P(x), fl.P(y), ctypes.c_int(fwd))
            got, _ = ORC.run1(fam, d, n, x)
            scale = 1.0
            assert np.max(np.abs(got - scale * y)) <= 1e-13 * max(1.0, np.max(np.abs(y)) * scale), (fam, d, n)
    for d, fwd in (("f", 1), ("b", 0)):
        (lib.orc_naive_rfftf if fwd else lib.orc_naive_rfftb)(ctypes.c_int(n), fl.P(x), fl.P(y))
        got, _ = ORC.run1("rfft", d, n, x)
        assert np.max(np.abs(got - y)) <= 1e-13 * max(1.0, np.max(np.abs(y)))
    xc = fl.rand_input("cfft", n, 9 + n)
    yc = np.zeros(n, dtype=np.complex128)
    for d, fwd in (("f", 1), ("b", 0)):
        lib.orc_naive_cfft(ctypes.c_int(n), fl.P(xc), fl.P(yc), ctypes.c_int(fwd))
        got, _ = ORC.run1("cfft", d, n, xc)
        assert np.max(np.abs(got - yc)) <= 1e-13 * max(1.0, np.max(np.abs(yc)))


def test_oracle_batched_and_2d_vs_golden():
    for key in G.files:
        if not key.endswith("/y"):
            continue
        name = key[:-2]
        if name.startswith("cfft2_"):
            d = name.split("_")[1]
            ldim = 11 if name.endswith("ld11") else 8
            y, ier = ORC.run2(d, ldim, 8, 6, G[name + "/x"])
            assert ier == 0 and np.array_equal(y, G[key]), name
        elif name.startswith("rfft2_"):
            d, (l, m), ldim = name.split("_")[1], (int(v) for v in name.split("_")[2].split("x")), int(name.split("ld")[1])
            y, ier = ORC.run2r(d, ldim, l, m, G[name + "/x"])
            assert ier == 0
            assert fl.rel_l2(fl.rows2(y, ldim, l, m), fl.rows2(G[key], ldim, l, m)) <= 2e-15, name
        elif name[4:6] == "m_":
            fam, d = name[:4], name[6]
            lot, n = (int(v) for v in name.split("_")[2].split("x"))
            jump, inc = (n, 1) if name.endswith("cols") else (1, lot)
            y, ier = ORC.runm(fam, d, lot, jump, n, inc, G[name + "/x"])
            assert ier == 0
            assert fl.rel_l2(y, G[key]) <= 1e-14, name


@pytest.mark.skipif(fl.ref() is None, reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_vs_live_reference():
    R = fl.Lib(fl.ref())
    sizes = list(range(1, 40)) + [49, 77, 96, 125, 169, 256, 343, 360, 500, 625, 1001, 2048]
    for fam in fl.FAMILIES:
        for n in sizes:
            x = fl.rand_input(fam, n, 31 * n + 1)
            wa, _ = R.init(fam, n)
            wb, _ = ORC.init(fam, n)
            assert np.array_equal(wa, wb), (fam, n)
            for d in "fb":
                a, ia = R.run1(fam, d, n, x)
                b, ib = ORC.run1(fam, d, n, x)
                assert ia == ib == 0
                if fam == "cfft":
                    assert np.array_equal(a, b), (fam, d, n)
                else:
                    assert fl.rel_l2(b, a) <= 2e-14 + fl.ref_noise(fam, n), (fam, d, n, fl.rel_l2(b, a))


@pytest.mark.skipif(fl.ref() is None, reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_rfft2_vs_live_reference():
    R = fl.Lib(fl.ref())
    a, _, _ = R.init2r(12, 10)
    b, _, _ = ORC.init2r(12, 10)
    assert np.array_equal(a, b)
    for (ld, l, m) in ((1, 1, 1), (1, 1, 4), (4, 4, 1), (2, 2, 2), (3, 3, 3), (5, 5, 8), (16, 16, 16), (33, 30, 21), (100, 100, 60)):
        x = fl.rand_input("rfft", ld * m, l * 100 + m)
        for d in "fb":
            ya, ia = R.run2r(d, ld, l, m, x)
            yb, ib = ORC.run2r(d, ld, l, m, x)
            assert ia == ib == 0
            assert fl.rel_l2(fl.rows2(yb, ld, l, m), fl.rows2(ya, ld, l, m)) <= 2e-15, (d, ld, l, m)
        for kw in (dict(lenwrk_=(l + 1) * m - 1), dict(lensav_=10)):
            ya, ia = R.run2r("f", ld, l, m, x, **kw)
            yb, ib = ORC.run2r("f", ld, l, m, x, **kw)
            assert ia == ib != 0 and np.array_equal(ya, yb)
    x = fl.rand_input("rfft", 40, 3)
    assert R.run2r("f", 4, 5, 8, x)[1] == ORC.run2r("f", 4, 5, 8, x)[1] == 5


def test_option_convolution_oracle_vs_golden_and_closed_form():
    """test/vargamma.c: the oracle's restatement against values produced by the reference itself, and the
    Black-Scholes closed form the reference prints (8.779874623570; its own N = 4096 value is 8.779878465793)"""
    assert [fl.oracle().orc_next_fast_even_size(n) for n in (1, 2, 3, 7, 11, 13, 127, 1001, 4097)] == \
        [2, 2, 4, 8, 12, 16, 128, 1024, 4320]
    for n in (128, 1000, 4096, 5000):
        got = np.array([fl.option_oracle(n, c) for c in fl.OPTION_CASES])
        assert np.max(np.abs(got - G[f"option_{n}"]) / np.abs(G[f"option_{n}"])) <= 1e-13, n
    assert abs(fl.option_oracle(4096, fl.OPTION_CASES[0]) - 8.779878465793) < 1e-11
    assert abs(fl.option_oracle(1 << 16, fl.OPTION_CASES[0]) - 8.779874623570) < 2e-8


@pytest.mark.skipif(fl.ref() is None, reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_ier_codes_vs_live_reference():
    """every argument-error code, incl. the upstream quirk that sinq1b_/sinqmb_ report 20 (fftpack.c:14151-14179)"""
    R = fl.Lib(fl.ref())
    n, lot = 12, 3
    for fam in fl.FAMILIES:
        xm, x1, big = fl.rand_input(fam, n * lot, 1), fl.rand_input(fam, n, 1), fl.rand_input(fam, 200, 3)
        for d in "fb":
            for kw in (dict(lenx=n * lot - 1), dict(lensav_=fl.lensav(fam, n) - 1), dict(lenwrk_=fl.lenwrk(fam, n, lot) - 1)):
                assert R.runm(fam, d, lot, n, n, 1, xm, **kw)[1] == ORC.runm(fam, d, lot, n, n, 1, xm, **kw)[1] != 0, (fam, d, kw)
            for kw in (dict(lenx=n - 1), dict(lensav_=fl.lensav(fam, n) - 1), dict(lenwrk_=fl.lenwrk(fam, n) - 1)):
                assert R.run1(fam, d, n, x1, **kw)[1] == ORC.run1(fam, d, n, x1, **kw)[1] != 0, (fam, d, kw)
            assert R.runm(fam, d, 4, 2, 6, 3, big)[1] == ORC.runm(fam, d, 4, 2, 6, 3, big)[1] != 0, (fam, d)


@pytest.mark.skipif(fl.naive_ref() is None, reason="oracle/_ref not built (no /root/reference on this box)")
def test_oracle_l2_wrapper_vs_live_reference():
    """the oracle's restatement of cfftpack.c (orc_l2_*) against the reference's own object API, all families,
    ortho on/off, stride errors, rfft repack, fft2.  Return codes are compared as zero/non-zero only: with a bad
    stride the reference keeps transforming and reports whichever inner code it meets last."""
    worst = fl.l2_compare(fl.naive_ref(), (1, 2, 3, 4, 5, 8, 16, 30, 31, 60, 100, 1000, 1001), exact_codes=False, tol_=2e-14, noise=True)
    print("L2 oracle vs reference, worst rel-L2", worst)


def test_reference_is_the_noisy_side_for_large_primes():
    """rfft 998 = 2*499 (cost N=999): against the long-double definition the oracle is ~100x closer than
    the golden (reference) answer, which justifies fl.ref_noise()."""
    import ctypes
    n = 998
    x = fl.vec("rfft", "rand", n)
    want = np.zeros(n)
    fl.oracle().orc_naive_rfftf(ctypes.c_int(n), fl.P(x), fl.P(want))
    got, _ = ORC.run1("rfft", "f", n, x)
    e_orc = fl.rel_l2(got, want)
    assert e_orc <= 5e-15
    xc = fl.vec("cost", "rand", 999)
    yo, _ = ORC.run1("cost", "f", 999, xc)
    e_gold = fl.rel_l2(G["cost_f_999_rand/y"], yo)
    assert e_gold > 10 * e_orc
