"""GPU parity tests: libcfftpack_b200.so, called through its C ABI, against the CPU oracle, the golden
vectors produced by the unmodified reference, and size-independent properties at BASELINE.json's sizes.

Bar (north_star): relative L2 error per sequence <= 1e-12 * log2(N)  (fl.tol).
"""
import ctypes

import numpy as np
import pytest

import fftlibs as fl

pytestmark = pytest.mark.gpu

PROD = fl.Lib(fl.product())
ORC = fl.Lib(fl.oracle(), "orc_")
G = fl.golden()

SIZES = [2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 15, 16, 20, 25, 27, 30, 32, 35, 49, 60, 64, 77, 100, 121, 128, 169, 210, 256,
         343, 360, 500, 512, 625, 999, 1000, 1001, 1002, 1024, 2048, 4096]


def _rows(a, lot, jump, n, inc):
    return [a[m * jump + inc * np.arange(n)] for m in range(lot)]


def _check_batched(fam, d, lot, jump, n, inc, seed=0, extra=0.0):
    span = (lot - 1) * jump + (n - 1) * inc + 1
    x = fl.rand_input(fam, span + 2, seed + 13 * n + lot)
    a, ia = PROD.runm(fam, d, lot, jump, n, inc, x, lenx=span, work=False)
    b, ib = ORC.runm(fam, d, lot, jump, n, inc, x, lenx=span)
    assert ia == ib == 0, (fam, d, lot, jump, n, inc, ia, ib, fl.product().cfb200_last_error())
    touched = np.zeros(len(x), bool)
    for m in range(lot):
        touched[m * jump + inc * np.arange(n)] = True
    assert np.array_equal(a[~touched], x[~touched]), "elements outside the sequences must not change"
    worst = max(fl.rel_l2(ra, rb) for ra, rb in zip(_rows(a, lot, jump, n, inc), _rows(b, lot, jump, n, inc)))
    assert worst <= fl.tol(n) + extra, (fam, d, lot, jump, n, inc, worst)
    return worst


@pytest.mark.parametrize("fam", fl.FAMILIES)
def test_single_sequence_vs_oracle(fam):
    launches0 = fl.product().cfb200_launch_count()
    for n in SIZES:
        x = fl.rand_input(fam, n, 17 * n + 3)
        for d in "fb":
            a, ia = PROD.run1(fam, d, n, x)
            b, ib = ORC.run1(fam, d, n, x)
            assert ia == ib == 0, (fam, d, n, ia, ib, fl.product().cfb200_last_error())
            assert fl.rel_l2(a, b) <= fl.tol(n), (fam, d, n, fl.rel_l2(a, b))
    assert fl.product().cfb200_launch_count() > launches0, "no kernel was launched: the CUDA path did not run"


@pytest.mark.parametrize("fam", fl.FAMILIES)
def test_golden_vectors_from_the_reference(fam):
    """outputs of the unmodified reference (tests/golden/make_golden.py), incl. its tests' own input vectors"""
    for key in G.files:
        if key.endswith("/y") and key.split("_")[0] == fam and key.count("_") == 3:
            _, d, n, kind = key[:-2].split("_")
            n = int(n)
            if n == 1:
                continue
            y, ier = PROD.run1(fam, d, n, fl.vec(fam, kind, n))
            assert ier == 0
            e = fl.rel_l2(y, G[key])
            assert e <= fl.tol(n) + fl.ref_noise(fam, n), (key, e)


def test_golden_batched_and_2d():
    for key in G.files:
        if not key.endswith("/y"):
            continue
        name = key[:-2]
        if name.startswith("cfft2_"):
            d = name.split("_")[1]
            ldim = 11 if name.endswith("ld11") else 8
            y, ier = PROD.run2(d, ldim, 8, 6, G[name + "/x"])
            assert ier == 0 and fl.rel_l2(y, G[key]) <= fl.tol(48), name
        elif name.startswith("rfft2_"):
            d, (l, m), ldim = name.split("_")[1], (int(v) for v in name.split("_")[2].split("x")), int(name.split("ld")[1])
            y, ier = PROD.run2r(d, ldim, l, m, G[name + "/x"])
            assert ier == 0
            assert fl.rel_l2(fl.rows2(y, ldim, l, m), fl.rows2(G[key], ldim, l, m)) <= fl.tol(l * m), name
        elif name[4:6] == "m_":
            fam, d = name[:4], name[6]
            lot, n = (int(v) for v in name.split("_")[2].split("x"))
            jump, inc = (n, 1) if name.endswith("cols") else (1, lot)
            y, ier = PROD.runm(fam, d, lot, jump, n, inc, G[name + "/x"], work=False)
            assert ier == 0 and fl.rel_l2(y, G[key]) <= fl.tol(n), name


@pytest.mark.parametrize("fam", fl.FAMILIES)
def test_lot_jump_inc_layouts(fam):
    for (lot, n) in ((5, 12), (33, 64), (7, 30), (17, 128), (64, 100), (3, 1024), (40, 256)):
        for d in "fb":
            _check_batched(fam, d, lot, n, n, 1)
            _check_batched(fam, d, lot, 1, n, lot)
            _check_batched(fam, d, lot, n + 3, n, 1)
            _check_batched(fam, d, lot, 1, n, lot + 2)
            _check_batched(fam, d, lot, 2 * n + 1, n, 2)


def test_config4_dct_dst_lengths():
    """cost/sint/cosq at N = 1000 and 1001 (underlying real FFTs 999, 1000, 1001, 1002 -- SURVEY 8(a) note 2)"""
    for fam in ("cost", "sint", "cosq", "sinq"):
        for n in (1000, 1001):
            for d in "fb":
                _check_batched(fam, d, 65, n, n, 1, seed=5)
                _check_batched(fam, d, 16, 1, n, 16, seed=6)


def test_headline_shapes_reduced_lot():
    for fam in ("cfft", "rfft"):
        for d in "fb":
            _check_batched(fam, d, 33, 4096, 4096, 1)
            _check_batched(fam, d, 8, 1, 4096, 8)
            _check_batched(fam, d, 5, 8192, 8192, 1)


def test_long_complex_four_step():
    nmax = fl.product().cfb200_max_onchip_complex()
    for n in (16384, 7 * 11 * 13 * 9, 2 * nmax + 2, 65536):
        for d in "fb":
            _check_batched("cfft", d, 3, n, n, 1)
    _check_batched("cfft", "f", 4, 1, 16384, 4)


def test_long_power_of_two_1d_beyond_four_step():
    """cfft1f_/cfft1b_ of 2^21 .. 2^26 points on one GPU (six-step over the four-step sweeps; the reference does any N in
    core, fftpack.c:2199-2245): full comparison with the CPU oracle up to 2^24, sampled bins against the direct DFT sums
    of the definition (test/naivepack.c naive_fft) at 2^26, round trips, and run-to-run determinism (the first version
    of this path exposed a landing-buffer race between shared-memory reads and the next bulk copy)."""
    torch = _torch()
    import sys
    import cfftpack_b200 as cb
    sys.path.insert(0, fl.ROOT + "/tools")
    from run_dist1d import dft_bins
    for a in (21, 22, 23, 24, 26):
        n = 1 << a
        g = torch.Generator(device="cuda").manual_seed(a)
        x0 = torch.view_as_complex(torch.rand(n, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
        plan = cb.Plan("cfft", n)
        outs = []
        for rep in range(4):
            x = x0.clone()
            assert plan.multi("f", x.data_ptr(), 1, n, 1, n) == 0, cb.last_error()
            cb.synchronize()
            outs.append(x)
        for o in outs[1:]:
            assert torch.equal(torch.view_as_real(o), torch.view_as_real(outs[0])), (a, "run-to-run difference")
        x = outs[0]
        if a <= 24:
            want, ier = ORC.run1("cfft", "f", n, x0.cpu().numpy())
            assert fl.rel_l2(x.cpu().numpy(), want) <= fl.tol(n), a
        else:
            gen = torch.Generator().manual_seed(3)
            bins = sorted(set([0, 1, n // 2, n - 1] + torch.randint(0, n, (96,), generator=gen).tolist()))
            want = dft_bins(x0, 0, n, bins, -1.0, 1) / n
            rms = float(torch.sqrt((x.abs() ** 2).mean()))
            assert float((x[torch.tensor(bins, device="cuda")] - want).abs().max()) <= fl.tol(n) * rms, a
        assert plan.multi("b", x.data_ptr(), 1, n, 1, n) == 0
        cb.synchronize()
        back = float((torch.view_as_real(x) - torch.view_as_real(x0)).norm() / torch.view_as_real(x0).norm())
        assert back <= fl.tol(n), (a, back)
    # batched and strided long sequences: lot = 3 sequences of 2^21 interleaved (jump = 1, inc = 3)
    n, lot = 1 << 21, 3
    xh = fl.rand_input("cfft", n * lot, 5)
    a_, ia = PROD.runm("cfft", "f", lot, 1, n, lot, xh, work=False)
    b_, ib = ORC.runm("cfft", "f", lot, 1, n, lot, xh)
    assert ia == ib == 0
    assert max(fl.rel_l2(a_[m::lot], b_[m::lot]) for m in range(lot)) <= fl.tol(n)


def test_large_prime_factors_chirp_z():
    for n, lot in ((4831, 3), (2 * 4339, 2), (10007, 2), (65537, 1)):
        extra = 2e-15 * np.log2(n)  # the oracle's own O(p^2) sums are the noisier side for primes this large
        for d in "fb":
            _check_batched("cfft", d, lot, n, n, 1, extra=extra)
    _check_batched("cfft", "f", 3, 1, 4831, 3, extra=3e-14)
    _check_batched("cost", "f", 2, 8580, 8580, 1, extra=3e-14)  # rfft length 8579 = 23 * 373


def test_cfft2_vs_oracle():
    for (ldim, l, m) in ((8, 8, 6), (11, 8, 6), (64, 64, 64), (130, 128, 96), (1024, 1024, 512), (100, 100, 75)):
        c = fl.rand_input("cfft", ldim * m, l + m)
        for d in "fb":
            a, ia = PROD.run2(d, ldim, l, m, c)
            b, ib = ORC.run2(d, ldim, l, m, c)
            assert ia == ib == 0
            assert fl.rel_l2(a, b) <= fl.tol(l * m), (d, ldim, l, m, fl.rel_l2(a, b))


def test_rfft2_vs_oracle_and_numpy():
    """2-D real transforms (fftpack.c:13282, :13113) incl. odd sizes, ldim > l, and a size whose rows leave the chip"""
    for (ldim, l, m) in ((1, 1, 4), (4, 4, 1), (2, 2, 2), (3, 3, 3), (8, 8, 6), (11, 8, 6), (9, 7, 5), (64, 64, 48),
                         (33, 30, 21), (130, 128, 96), (100, 100, 75), (1024, 1024, 512), (1001, 999, 1000)):
        r = fl.rand_input("rfft", ldim * (m - 1) + l, 3 * l + m)
        for d in "fb":
            a, ia = PROD.run2r(d, ldim, l, m, r)
            b, ib = ORC.run2r(d, ldim, l, m, r)
            assert ia == ib == 0
            e = fl.rel_l2(fl.rows2(a, ldim, l, m), fl.rows2(b, ldim, l, m))
            assert e <= fl.tol(l * m), (d, ldim, l, m, e)
        if ldim > l:
            pad = np.ones(len(r), bool)
            for j in range(m):
                pad[j * ldim: j * ldim + l] = False
            assert np.array_equal(a[pad], r[pad])
    # definition: F(i, j) half-complex along i of fft2(x) / (l m); then the round trip
    l, m = 2048, 1536
    x = fl.rand_input("rfft", l * m, 77)
    f, ier = PROD.run2r("f", l, l, m, x)
    assert ier == 0
    X = np.fft.rfft2(x.reshape(m, l), axes=(0, 1)) / (l * m)  # X[j, f]: rfft along i (fast axis), fft along j
    F = f.reshape(m, l)
    k = np.arange(1, l // 2)
    assert fl.rel_l2(F[:, 2 * k - 1] + 1j * F[:, 2 * k], X[:, k]) <= fl.tol(l * m)
    g, ier = PROD.run2r("b", l, l, m, f)
    assert ier == 0 and fl.rel_l2(g, x) <= fl.tol(l * m)


def test_option_convolution_batched_vs_oracle_golden_and_black_scholes():
    """the reference's option-pricing application, batched on the device"""
    for n in (128, 1000, 4096, 5000):
        val, N, ier = fl.option_product(fl.product(), n, fl.OPTION_CASES)
        assert ier == 0
        assert np.max(np.abs(val - G[f"option_{n}"]) / np.abs(G[f"option_{n}"])) <= 1e-12, n
    # a larger batch: every option must equal its single-option oracle value; the grid size leaves the chip at 2^16
    rng = np.random.default_rng(9)
    cases = [(float(rng.uniform(80, 120)), float(rng.uniform(80, 120)), float(rng.uniform(0.1, 0.4)), float(rng.uniform(-0.3, 0.1)),
              float(rng.uniform(0.1, 0.5)), float(rng.uniform(0.2, 2.0)), float(rng.uniform(0.0, 0.06)), int(rng.integers(2)),
              int(rng.integers(2))) for _ in range(300)]
    val, N, ier = fl.option_product(fl.product(), 2000, cases)
    assert ier == 0 and N == 2000
    want = np.array([fl.option_oracle(2000, c) for c in cases])
    assert np.max(np.abs(val - want) / np.maximum(np.abs(want), 1e-3)) <= 1e-11
    val, N, ier = fl.option_product(fl.product(), 1 << 16, fl.OPTION_CASES[:2])
    assert ier == 0 and abs(val[0] - 8.779874623570) < 2e-8 and abs(val[1] - 9.3424659413582116) < 2e-5
    import cfftpack_b200 as cb
    v2, _ = cb.option_convolution(4096, 100.0, np.array([98.0, 105.0]), 0.12, -0.14, 0.2, 1.0, 0.05, call=True, black_scholes=True)
    assert abs(v2[0] - 8.779878465793) < 1e-9


def test_l2_object_api_vs_oracle():
    """fft_create/.../rfft_inverse (cfftpack.h) on the device, incl. orthonormal scalings, the rfft
    repack, return codes, the batch extension, and device-resident data"""
    fl.l2_compare(fl.product(), (1, 2, 3, 4, 5, 8, 16, 30, 31, 60, 100, 1000, 1001, 4096))
    fl.l2_batch_check(fl.product())
    fl.l2_batch_check(fl.product(), n=1000, lot=64)
    torch = _torch()
    S = fl.bind_l2(fl.product())
    n, lot = 4096, 256
    x = torch.rand(lot * n, device="cuda", dtype=torch.float64) - 0.5
    out = torch.zeros(lot * (n + 2), device="cuda", dtype=torch.float64)
    back = torch.zeros_like(x)
    f = S.rfft_create(n)
    S.cfb200_fft_batch(f, lot)
    assert S.rfft_forward(f, x.data_ptr(), out.data_ptr()) == 0 and S.rfft_inverse(f, out.data_ptr(), back.data_ptr()) == 0
    fl.product().cfb200_synchronize()
    X = torch.fft.rfft(x.view(lot, n), dim=1) / n * 2
    got = torch.view_as_complex(out.view(lot, n // 2 + 1, 2))
    # FFTPACK's half-complex convention: (2/n) Re, -(2/n) Im for 0 < k < n/2; (1/n) at the ends
    assert float((got[:, 1:n // 2] - torch.conj(X[:, 1:n // 2])).abs().max()) < 1e-13
    assert float((got[:, 0].real - X[:, 0].real / 2).abs().max()) < 1e-13
    assert float((back - x).abs().max()) < 1e-13
    S.fft_free(f)


def test_known_answers():
    """impulse, constant and single tone under the reference's scaling (forward 1/N e^{-i}, backward e^{+i})"""
    n = 4096
    x = np.zeros(n, np.complex128)
    x[1] = 1.0
    y, _ = PROD.run1("cfft", "f", n, x)
    k = np.arange(n)
    assert np.max(np.abs(y - np.exp(-2j * np.pi * k / n) / n)) < 1e-15
    y, _ = PROD.run1("cfft", "b", n, x)
    assert np.max(np.abs(y - np.exp(2j * np.pi * k / n))) < 1e-12
    y, _ = PROD.run1("cfft", "f", n, np.ones(n, np.complex128))
    assert abs(y[0] - 1) < 1e-15 and np.max(np.abs(y[1:])) < 1e-15
    # real tone: x_t = cos(2 pi 3 t/n) + 0.5 sin(2 pi 5 t/n)  ->  a_3 = 1, b_5 = 0.5 (rfftmf_ convention)
    t = np.arange(n)
    r, _ = PROD.run1("rfft", "f", n, np.cos(2 * np.pi * 3 * t / n) + 0.5 * np.sin(2 * np.pi * 5 * t / n))
    want = np.zeros(n)
    want[2 * 3 - 1] = 1.0
    want[2 * 5] = 0.5
    assert np.max(np.abs(r - want)) < 1e-14


def _sample_rows(lot, count):
    """>= `count` sequences: the first and last tiles plus an odd-stride walk over the whole lot"""
    if lot <= count:
        return np.arange(lot)
    rng = np.random.default_rng(lot)
    walk = (np.arange(count) * ((lot // count) | 1) + 3) % lot
    return np.unique(np.concatenate([np.arange(8), np.arange(lot - 8, lot), walk, rng.integers(0, lot, 64)]))


def _sampled_rows_vs_oracle(torch, fam, d, n, lot, x_in, x_out, count=1024, jump=None, inc=1):
    """>= count sampled sequences of a device-resident batch against the CPU oracle, each to the north_star bar.
    Arrays are float64 tensors (complex as interleaved pairs); layout (jump, inc) in elements, default contiguous."""
    jump = n if jump is None else jump
    rows = _sample_rows(lot, count)
    esz = 2 if fam == "cfft" else 1
    idx = torch.as_tensor(rows[:, None] * jump + inc * np.arange(n)[None, :], device=x_in.device)  # [rows][n] elements
    xi = x_in.view(-1, esz)[idx.reshape(-1)].cpu().numpy().reshape(-1)
    xo = x_out.view(-1, esz)[idx.reshape(-1)].cpu().numpy().reshape(-1)
    if fam == "cfft":
        xi, xo = xi.view(np.complex128), xo.view(np.complex128)
    want, ier = ORC.runm(fam, d, len(rows), n, n, 1, xi)
    assert ier == 0
    err = np.array([fl.rel_l2(xo[j * n:(j + 1) * n], want[j * n:(j + 1) * n]) for j in range(len(rows))])
    assert err.max() <= fl.tol(n), (fam, d, n, lot, int(rows[err.argmax()]), float(err.max()))
    return float(err.max())


def _torch():
    import torch
    return torch


def test_device_pointers_full_size_config2_properties():
    """BASELINE config 2: cfftmf_ N=4096, lot=65536 on device-resident data: forward then backward is the identity,
    a sample of sequences agrees with the oracle, and the library really launched kernels."""
    torch = _torch()
    import cfftpack_b200 as cb
    n, lot = 4096, 65536
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(lot * n, 2, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    x0 = x.clone()
    plan = cb.Plan("cfft", n)
    before = cb.launch_count()
    assert plan.multi("f", x.data_ptr(), lot, n, 1, lot * n) == 0
    cb.synchronize()
    _sampled_rows_vs_oracle(torch, "cfft", "f", n, lot, x0, x)
    assert plan.multi("b", x.data_ptr(), lot, n, 1, lot * n) == 0
    cb.synchronize()
    assert cb.launch_count() >= before + 2
    err = (x - x0).norm() / x0.norm()
    assert float(err) <= fl.tol(n), float(err)
    # Parseval on the whole batch (checksum of checksums): sum |X|^2 * N == sum |x|^2
    assert plan.multi("f", x.data_ptr(), lot, n, 1, lot * n) == 0
    cb.synchronize()
    e_in, e_out = float((x0 * x0).sum()), float((x * x).sum()) * n
    assert abs(e_in - e_out) <= 1e-11 * e_in


def test_device_pointers_full_size_config3_properties():
    """BASELINE config 3: rfftmf_ N=4096, lot=65536."""
    torch = _torch()
    import cfftpack_b200 as cb
    n, lot = 4096, 65536
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand(lot * n, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    x0 = x.clone()
    plan = cb.Plan("rfft", n)
    assert plan.multi("f", x.data_ptr(), lot, n, 1, lot * n) == 0
    cb.synchronize()
    _sampled_rows_vs_oracle(torch, "rfft", "f", n, lot, x0, x)
    xf = x.clone()
    assert plan.multi("b", x.data_ptr(), lot, n, 1, lot * n) == 0
    cb.synchronize()
    _sampled_rows_vs_oracle(torch, "rfft", "b", n, lot, xf, x)
    err = (x - x0).norm() / x0.norm()
    assert float(err) <= fl.tol(n), float(err)


def test_device_pointers_config4_roundtrip_and_sample():
    torch = _torch()
    import cfftpack_b200 as cb
    lot = 32768
    for fam, n in (("cost", 1001), ("sint", 1000), ("cosq", 1000), ("cosq", 1001), ("cost", 1000), ("sint", 1001)):
        g = torch.Generator(device="cuda").manual_seed(n)
        x = torch.rand(lot * n, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
        x0 = x.clone()
        plan = cb.Plan(fam, n)
        assert plan.multi("f", x.data_ptr(), lot, n, 1, lot * n) == 0, cb.last_error()
        cb.synchronize()
        _sampled_rows_vs_oracle(torch, fam, "f", n, lot, x0, x)
        xf = x.clone()
        assert plan.multi("b", x.data_ptr(), lot, n, 1, lot * n) == 0
        cb.synchronize()
        _sampled_rows_vs_oracle(torch, fam, "b", n, lot, xf, x, count=256)
        cb.synchronize()
        err = float((x - x0).norm() / x0.norm())
        assert err <= fl.tol(n), (fam, n, err)


def test_cfft2_16384_device_roundtrip_and_separability():
    """BASELINE config 5 shape on one GPU: 2-D forward of a separable input equals the outer product of 1-D
    transforms; forward then backward is the identity."""
    torch = _torch()
    import cfftpack_b200 as cb
    l = m = 16384
    g = torch.Generator(device="cuda").manual_seed(3)
    u = torch.view_as_complex(torch.rand(l, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    v = torch.view_as_complex(torch.rand(m, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    c = (v[:, None] * u[None, :]).contiguous()  # column-major c(l, m): c[j, i] = u[i] v[j]
    lib = fl.product()
    ws, ls, ier = PROD.init2(l, m)
    I = ctypes.c_int
    ierc = I(-1)
    dummy = ctypes.c_double(0)
    args = (ctypes.byref(I(l)), ctypes.byref(I(l)), ctypes.byref(I(m)), ctypes.c_void_p(c.data_ptr()), fl.P(ws),
            ctypes.byref(I(ls)), ctypes.byref(dummy), ctypes.byref(I(2 * l * m)), ctypes.byref(ierc))
    lib.cfft2f_(*args)
    cb.synchronize()
    assert ierc.value == 0, lib.cfb200_last_error()
    U, _ = ORC.run1("cfft", "f", l, u.cpu().numpy())
    V, _ = ORC.run1("cfft", "f", m, v.cpu().numpy())
    rows = [0, 1, 5000, 16383]
    for j in rows:
        got = c[j].cpu().numpy()
        assert fl.rel_l2(got, V[j] * U) <= fl.tol(l * m), j
    lib.cfft2b_(*args)
    cb.synchronize()
    assert ierc.value == 0
    back = (v[:, None] * u[None, :])
    err = float((torch.view_as_real(c) - torch.view_as_real(back)).norm() / torch.view_as_real(back).norm())
    assert err <= fl.tol(l * m), err


def test_interleaved_layout_full_size_config2():
    """BASELINE config 2 in the reference's own "vector" layout (jump=1, inc=lot; test/ftest.c:64):
    N=4096, lot=65536, >= 1024 sampled sequences against the oracle, forward and back."""
    torch = _torch()
    import cfftpack_b200 as cb
    n, lot = 4096, 65536
    g = torch.Generator(device="cuda").manual_seed(17)
    x = torch.rand(lot * n, 2, generator=g, device="cuda", dtype=torch.float64) * 2 - 1
    x0 = x.clone()
    plan = cb.Plan("cfft", n)
    assert plan.multi("f", x.data_ptr(), lot, 1, lot, lot * n) == 0, cb.last_error()
    cb.synchronize()
    _sampled_rows_vs_oracle(torch, "cfft", "f", n, lot, x0, x, jump=1, inc=lot)
    assert plan.multi("b", x.data_ptr(), lot, 1, lot, lot * n) == 0
    cb.synchronize()
    err = float((x - x0).norm() / x0.norm())
    assert err <= fl.tol(n), err


def test_cfft2_16384_nonseparable_sampled_rows_and_columns():
    """BASELINE config 5 at full size on a NON-separable random matrix: 36 full output columns and 35 full output rows
    against the oracle's 1-D transforms of the single-bin DFT sums along the other dimension (bench.cfft2_parity),
    then forward+backward is the identity.  Catches any misplaced element of the four-step transposes."""
    torch = _torch()
    import cfftpack_b200 as cb
    import bench
    l = m = 16384
    x_in = bench.slab_input(torch, l, m, 0)
    c = x_in.clone()
    lib = fl.product()
    ws, ls, ier = PROD.init2(l, m)
    I = ctypes.c_int
    ierc, dummy = I(-1), ctypes.c_double(0)
    args = (ctypes.byref(I(l)), ctypes.byref(I(l)), ctypes.byref(I(m)), ctypes.c_void_p(c.data_ptr()), fl.P(ws),
            ctypes.byref(I(ls)), ctypes.byref(dummy), ctypes.byref(I(2**31 - 1)), ctypes.byref(ierc))
    lib.cfft2f_(*args)
    cb.synchronize()
    assert ierc.value == 0, lib.cfb200_last_error()
    err, ncols, nrows = bench.cfft2_parity(torch, None, x_in, c, l, m, 0, 1)
    assert ncols >= 32 and nrows >= 32
    assert err <= fl.tol(l * m), err
    lib.cfft2b_(*args)
    cb.synchronize()
    back = float((torch.view_as_real(c) - torch.view_as_real(x_in)).norm() / torch.view_as_real(x_in).norm())
    assert back <= fl.tol(l * m), back


def test_pipelined_staging_with_scratch_using_lengths():
    """the host-array pipeline runs lot-chunks concurrently on three streams; lengths whose transform
    needs device scratch (four-step 16384, chirp-z prime, long real) must not share it between chunks.  The chunk size
    is lowered through CFB200_PIPE_CHUNK_KB (read once per process, hence the subprocess) so that small batches take
    the pipelined path with many chunks in flight; results must equal the device-pointer path (bit for bit for complex data)."""
    import os
    import subprocess
    import sys
    code = r"""
import sys, torch
sys.path.insert(0, %r)
import cfftpack_b200 as cb
bad = 0
for fam, n, lot in (("cfft", 16384, 96), ("cfft", 10007, 64), ("rfft", 32768, 96), ("cost", 10001, 64), ("cfft", 4096, 512)):
    esz = 2 if fam == "cfft" else 1
    h = torch.empty(lot * n * esz, dtype=torch.float64, pin_memory=True).uniform_(-1, 1)
    d = h.cuda()
    plan = cb.Plan(fam, n)
    for rep in range(3):
        hh = h.clone().pin_memory()
        dd = d.clone()
        assert plan.multi("f", dd.data_ptr(), lot, n, 1, lot * n) == 0, cb.last_error()
        cb.synchronize()
        assert plan.multi("f", hh.data_ptr(), lot, n, 1, lot * n) == 0, cb.last_error()
        # complex: bit for bit.  Real families pair rows (z = x_a + i x_b) inside a chunk, so a chunk boundary at an odd
        # row re-pairs them and moves the rounding (the running sums of cost/sint amplify it to ~1e-14 relative; the
        # inputs are unseeded, so 1e-14 was borderline); a scratch race would be a gross error.
        ref_ = dd.cpu()
        same = torch.equal(hh, ref_) if fam == "cfft" else float((hh - ref_).norm() / ref_.norm()) <= 2e-13
        if not same:
            bad += 1
            print("MISMATCH", fam, n, lot, rep, float((hh - ref_).abs().max()))
print("BAD", bad)
sys.exit(1 if bad else 0)
""" % fl.ROOT
    env = dict(os.environ, CFB200_PIPE_CHUNK_KB="1024")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-2000:])


def test_pinned_host_arrays_take_the_pipelined_path():
    """a pinned host batch >= 128 MiB is staged in lot-chunks on three streams; result must equal the device path"""
    torch = _torch()
    import cfftpack_b200 as cb
    for fam, n, lot, esz in (("cfft", 4096, 2500, 2), ("rfft", 4096, 4500, 1)):
        h = torch.empty(lot * n * esz, dtype=torch.float64, pin_memory=True).uniform_(-1, 1)
        d = h.cuda()
        plan = cb.Plan(fam, n)
        assert plan.multi("f", d.data_ptr(), lot, n, 1, lot * n) == 0
        cb.synchronize()
        assert plan.multi("f", h.data_ptr(), lot, n, 1, lot * n) == 0, cb.last_error()
        assert torch.equal(h, d.cpu()), fam


def test_reference_own_test_programs_relinked_against_the_cuda_library():
    """the reference's wrapper layer (cfftpack.c, cfftextra.c) and its own asserting test
    (test/testall.c: DCT/DST families vs test/naivepack.c, abs tol 1e-13, N = 2, 32, 60) linked against
    libcfftpack_b200.so instead of fftpack.c.  Built by oracle/Makefile where /root/reference exists."""
    import os
    import subprocess
    exe = os.path.join(fl.ROOT, "oracle", "_ref", "testall_b200")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/testall_b200 not prebuilt")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.stdout.count("DCT tests passed") == 3 and out.stdout.count("DST tests passed") == 2, out.stdout + out.stderr
    assert "Assertion failed" not in out.stdout + out.stderr
    # test/ftest.c (print-only in the reference): cosqmb_/cosqmf_ on a 10x10 grid in both (lot, jump, inc) orientations
    a = subprocess.run([os.path.join(fl.ROOT, "oracle", "_ref", "ftest_b200")], capture_output=True, text=True, timeout=120).stdout
    b = subprocess.run([os.path.join(fl.ROOT, "oracle", "_ref", "ftest_ref")], capture_output=True, text=True, timeout=120).stdout
    import re
    fa = [float(v) for v in re.findall(r"-?\d+\.\d+", a)]
    fb = [float(v) for v in re.findall(r"-?\d+\.\d+", b)]
    assert len(fa) == len(fb) > 100
    assert max(abs(u - v) for u, v in zip(fa, fb)) <= 0.0100001  # printed with two decimals


def test_long_real_family_sequences():
    """sequences longer than one CTA's shared memory (e.g. test/vargamma.c runs rfft up to N = 2^20)"""
    for fam, n, lot in (("rfft", 32768, 3), ("rfft", 1 << 20, 1), ("cost", 10001, 2), ("sint", 9999, 3), ("cosq", 16384, 2),
                        ("sinq", 12000, 1), ("rfft", 3 * 5 * 7 * 11 * 13, 2)):
        x = fl.rand_input(fam, n * lot, n % 1000)
        for d in "fb":
            a, ia = PROD.runm(fam, d, lot, n, n, 1, x, work=False)
            b, ib = ORC.runm(fam, d, lot, n, n, 1, x)
            assert ia == ib == 0, (fam, d, n, ia, ib, fl.product().cfb200_last_error())
            worst = max(fl.rel_l2(a[i * n:(i + 1) * n], b[i * n:(i + 1) * n]) for i in range(lot))
            assert worst <= fl.tol(n), (fam, d, n, worst)


def test_concurrent_host_threads_share_wsave():
    """the reference is re-entrant (SURVEY 8(b) threading): concurrent calls with disjoint data and a shared wsave"""
    import threading
    fam, n, lot = "cfft", 1000, 64
    ws, ier = PROD.init(fam, n, multi=True)
    xs = [fl.rand_input(fam, n * lot, 100 + i) for i in range(6)]
    want = [ORC.runm(fam, "f", lot, n, n, 1, x)[0] for x in xs]
    got = [None] * len(xs)

    def work(i):
        for _ in range(5):
            got[i], ier = PROD.runm(fam, "f", lot, n, n, 1, xs[i], ws=ws, work=False)
            assert ier == 0

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(xs))]
    [t.start() for t in th]
    [t.join() for t in th]
    for i in range(len(xs)):
        assert fl.rel_l2(got[i], want[i]) <= fl.tol(n), i


def test_user_stream_ordering():
    """device-pointer calls are asynchronous on the stream given to cfb200_set_stream"""
    torch = _torch()
    import cfftpack_b200 as cb
    n, lot = 4096, 4096
    s = torch.cuda.Stream()
    plan = cb.Plan("cfft", n)
    with torch.cuda.stream(s):
        cb.set_stream(s.cuda_stream)
        x = torch.rand(lot * n, 2, device="cuda", dtype=torch.float64) - 0.5
        x0 = x.clone()
        assert plan.multi("f", x.data_ptr(), lot, n, 1, lot * n) == 0
        assert plan.multi("b", x.data_ptr(), lot, n, 1, lot * n) == 0
        err = (x - x0).norm() / x0.norm()  # enqueued on the same stream: sees both transforms
    s.synchronize()
    cb.set_stream(0)
    assert float(err) <= fl.tol(n)


def test_repeatability_and_sampled_parity_at_full_occupancy():
    """every kernel family with enough tiles that each persistent CTA walks many of them: repeated runs must agree
    bit for bit (a race between pipeline stages shows up as run-to-run differences) and sampled rows match the oracle"""
    torch = _torch()
    import cfftpack_b200 as cb
    cases = [("cfft", 4096, 8192, 4096, 1), ("rfft", 4096, 8191, 4096, 1), ("cfft", 1000, 30000, 1000, 1),
             ("cfft", 360, 40000, 1, 40000), ("cosq", 1000, 8192, 1000, 1), ("cost", 1001, 4097, 1001, 1),
             ("sint", 1000, 4096, 1, 4096), ("cfft", 16384, 700, 16384, 1), ("cfft", 16384, 512, 1, 512),
             ("rfft", 1000, 30001, 1000, 1)]
    for fam, n, lot, jump, inc in cases:
        esz = 2 if fam == "cfft" else 1
        span = (lot - 1) * jump + (n - 1) * inc + 1
        g = torch.Generator(device="cuda").manual_seed(n + lot)
        x0 = torch.rand(span * esz, generator=g, device="cuda", dtype=torch.float64) - 0.5
        plan = cb.Plan(fam, n)
        outs = []
        for rep in range(4):
            x = x0.clone()
            assert plan.multi("f", x.data_ptr(), lot, jump, inc, span) == 0, cb.last_error()
            cb.synchronize()
            outs.append(x)
        for o in outs[1:]:
            assert torch.equal(o, outs[0]), (fam, n, lot, "run-to-run difference")
        host_in = x0.cpu().numpy()
        host_out = outs[0].cpu().numpy()
        if fam == "cfft":
            host_in = host_in.view(np.complex128)
            host_out = host_out.view(np.complex128)
        for mrow in (0, 1, lot // 2, lot - 2, lot - 1):
            idx = mrow * jump + inc * np.arange(n)
            want, ier = ORC.run1(fam, "f", n, host_in[idx])
            assert fl.rel_l2(host_out[idx], want) <= fl.tol(n), (fam, n, lot, mrow)


def test_reference_option_pricing_demo_relinked():
    """test/vargamma.c (Carr-Madan style rfft convolution pricing, N = 128 ... 2^20; the workload behind BASELINE
    config 3) built against libcfftpack_b200.so prints the same prices as the all-reference build, and reproduces the
    known answers of SURVEY 8(c): Black-Scholes 8.779874623570, QuantLib variance-gamma target 9.3424659413582116."""
    import os
    import re
    import subprocess
    a_exe = os.path.join(fl.ROOT, "oracle", "_ref", "vargamma_b200")
    b_exe = os.path.join(fl.ROOT, "oracle", "_ref", "vargamma_ref")
    if not (os.path.exists(a_exe) and os.path.exists(b_exe)):
        pytest.skip("oracle/_ref/vargamma_* not prebuilt")
    a = subprocess.run([a_exe], capture_output=True, text=True, timeout=600).stdout
    b = subprocess.run([b_exe], capture_output=True, text=True, timeout=600).stdout
    rows = lambda s: [(int(m.group(1)), float(m.group(2))) for m in re.finditer(r"^\s*(\d+)\s+(-?\d+\.\d{12})\s", s, re.M)]
    ra, rb = rows(a), rows(b)
    assert len(ra) == len(rb) == 28, (len(ra), len(rb), a[-500:])
    for (na, pa), (nb, pb) in zip(ra, rb):
        assert na == nb and abs(pa - pb) <= 1e-9 * max(1.0, abs(pb)), (na, pa, pb)
    assert "Black Scholes Formula: 8.779874623570" in a
    bs_4096 = [p for n, p in ra[:14] if n == 4096][0]
    assert abs(bs_4096 - 8.779878465793) < 1e-9
    vg_last = ra[-1][1]
    assert abs(vg_last - 9.3424659413582116) < 2e-5


def test_host_array_fans_out_over_gpus_from_the_c_boundary():
    """needs >= 2 GPUs (skipped otherwise): with CFB200_DEVICES=2 one cfftmf_/rfftmf_/cosqmf_ call on a host array is
    sharded by lot over both GPUs inside the library (worker threads); results equal the single-GPU path."""
    import os
    import subprocess
    import sys
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("single GPU")
    code = r"""
import sys, ctypes, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
import cfftpack_b200 as cb
import fftlibs as fl
ORC = fl.Lib(fl.oracle(), "orc_")
assert cb.lib.cfb200_set_devices(0) >= 2
bad = 0
for fam, n, lot in (("cfft", 4096, 3001), ("rfft", 4096, 5000), ("cosq", 1001, 4001), ("cfft", 360, 7), ("sint", 1000, 2048)):
    esz = 2 if fam == "cfft" else 1
    h = torch.empty(lot * n * esz, dtype=torch.float64, pin_memory=True).uniform_(-1, 1)
    x0 = h.clone().numpy()
    plan = cb.Plan(fam, n)
    assert plan.multi("f", h.data_ptr(), lot, n, 1, lot * n) == 0, cb.last_error()
    got = h.numpy()
    rows = sorted(set([0, 1, lot // 2 - 1, lot // 2, lot // 2 + 1, lot - 2, lot - 1]))
    for r in rows:
        a = x0[r * n * esz:(r + 1) * n * esz]; b = got[r * n * esz:(r + 1) * n * esz]
        if fam == "cfft": a, b = a.view(np.complex128), b.view(np.complex128)
        want, ier = ORC.run1(fam, "f", n, a)
        if fl.rel_l2(b, want) > fl.tol(n):
            bad += 1; print("MISMATCH", fam, n, lot, r, fl.rel_l2(b, want))
# cfft2f_ on a host matrix: one GPU vs two (column slabs + fused P2P transposes inside the library)
l = 4096
h = torch.empty(l * l * 2, dtype=torch.float64, pin_memory=True).uniform_(-1, 1)
P = fl.Lib(fl.product())
outs = []
for ndev in (1, 2):
    assert cb.lib.cfb200_set_devices(ndev) == ndev
    y, ier = P.run2("f", l, l, l, h.numpy().view(np.complex128))
    assert ier == 0, cb.last_error()
    outs.append(y)
e2 = fl.rel_l2(outs[1], outs[0])
if e2 > fl.tol(l * l):
    bad += 1; print("MISMATCH cfft2 multi-GPU", e2)
yb, ier = P.run2("b", l, l, l, outs[1])
if ier != 0 or fl.rel_l2(yb, h.numpy().view(np.complex128)) > fl.tol(l * l):
    bad += 1; print("MISMATCH cfft2 multi-GPU round trip", ier)
print("BAD", bad)
sys.exit(1 if bad else 0)
""" % (fl.ROOT, fl.ROOT + "/tests")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-2000:])


def test_sharded_cfft2_two_gpus_fused_p2p_vs_nccl():
    """needs >= 2 GPUs (skipped on single-GPU boxes): the FFT+transpose fused path (P2P stores into peer slabs) must
    agree with the NCCL all-to-all path (bit for bit only when both pick the same factorisation, e.g. at 16384) and
    invert itself"""
    import json
    import os
    import subprocess
    import sys
    torch = _torch()
    if torch.cuda.device_count() < 2:
        pytest.skip("single GPU")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547",
                          os.path.join(fl.ROOT, "tools", "run_dist2d.py"), "4096"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    j = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert j["p2p_vs_nccl_rel_err"] <= fl.tol(4096 * 4096) and j["p2p_roundtrip_rel_err"] <= fl.tol(4096 * 4096), j


def test_randomized_shapes_vs_oracle():
    """seeded random (family, n, lot, inc, jump, direction) including awkward strides and tile-boundary lots"""
    rng = np.random.default_rng(20261018)
    lengths = [2, 3, 4, 5, 6, 7, 9, 10, 12, 14, 15, 18, 21, 25, 27, 33, 36, 45, 48, 50, 63, 64, 70, 81, 90, 96, 98, 100, 105, 110,
               121, 126, 128, 130, 143, 150, 169, 180, 187, 200, 221, 243, 250, 256, 289, 300, 323, 343, 360, 400, 441, 500, 512,
               539, 600, 625, 686, 720, 729, 768, 800, 900, 1000, 1024, 1331, 1536, 2000, 2048, 2187, 2401, 3000, 3125, 4000,
               4096, 4199, 4283, 4289, 5000, 6561, 8192, 9000, 10000, 16384]
    worst = 0.0
    for case in range(160):
        fam = fl.FAMILIES[int(rng.integers(len(fl.FAMILIES)))]
        n = int(lengths[int(rng.integers(len(lengths)))])
        lot = int(rng.choice([1, 2, 3, 5, 8, 15, 16, 17, 31, 32, 33, 64, 100]))
        if n * lot > 400000:
            lot = max(1, 400000 // n)
        layout = int(rng.integers(4))
        if layout == 0:
            inc, jump = 1, n
        elif layout == 1:
            inc, jump = 1, n + int(rng.integers(1, 9))
        elif layout == 2:
            inc, jump = lot, 1
        else:
            inc = int(rng.integers(2, 5))
            jump = inc * (n - 1) + 1 + int(rng.integers(0, 7))
        d = "fb"[int(rng.integers(2))]
        extra = fl.ref_noise(fam, n) + (3e-15 * np.log2(n) if fl.max_generic_factor(fl.underlying(fam, n)) > 1000 else 0.0)
        worst = max(worst, _check_batched(fam, d, lot, jump, n, inc, seed=case, extra=extra))
    print("worst rel-L2 over the random shapes:", worst)
