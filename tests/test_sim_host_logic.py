"""Host logic + kernel index arithmetic on the CPU: the product's own sources compiled against the
CUDA-thread emulator (tools/sim), compared with the oracle.  This is how plans, launch geometry, the
lot/jump/inc handling, the pair packing of real sequences and the four-step split are tested where no GPU
exists.  The emulator is test infrastructure; libcfftpack_b200.so neither contains nor loads it.
"""
import subprocess
import os

import numpy as np
import pytest

import fftlibs as fl


@pytest.fixture(scope="module")
def libs():
    subprocess.check_call(["make", "-s", "-C", os.path.join(fl.ROOT, "tools", "sim")])
    fl._cache.pop(os.path.join(fl.ROOT, "tools", "sim", "libcfftpack_sim.so"), None)
    return fl.Lib(fl.sim()), fl.Lib(fl.oracle(), "orc_")


def _check(S, O, fam, d, lot, jump, n, inc, seed=0):
    span = (lot - 1) * jump + (n - 1) * inc + 1
    x = fl.rand_input(fam, span + 3, seed + 13 * n + lot)
    a, ia = S.runm(fam, d, lot, jump, n, inc, x, lenx=span)
    b, ib = O.runm(fam, d, lot, jump, n, inc, x, lenx=span)
    assert ia == ib == 0, (fam, d, lot, jump, n, inc, ia, ib)
    # untouched elements must stay bit-identical; touched ones within the parity bar
    touched = np.zeros(len(x), bool)
    for m in range(lot):
        touched[m * jump + inc * np.arange(n)] = True
    assert np.array_equal(a[~touched], x[~touched])
    for m in range(lot):
        idx = m * jump + inc * np.arange(n)
        e = fl.rel_l2(a[idx], b[idx])
        assert e <= fl.tol(n), (fam, d, lot, jump, n, inc, m, e)


@pytest.mark.parametrize("fam", fl.FAMILIES)
def test_single_sequence_sizes(libs, fam):
    S, O = libs
    for n in [2, 3, 4, 5, 6, 7, 8, 9, 11, 12, 16, 27, 30, 32, 35, 49, 60, 64, 100, 128, 210, 256]:
        x = fl.rand_input(fam, n, 17 * n + 3)
        for d in "fb":
            a, ia = S.run1(fam, d, n, x)
            b, ib = O.run1(fam, d, n, x)
            assert ia == ib == 0
            assert fl.rel_l2(a, b) <= fl.tol(n), (fam, d, n)


@pytest.mark.parametrize("fam", fl.FAMILIES)
def test_lot_jump_inc_layouts(libs, fam):
    S, O = libs
    for (lot, n) in ((5, 12), (4, 64), (7, 30), (3, 128), (9, 8)):
        for d in "fb":
            _check(S, O, fam, d, lot, n, n, 1)          # contiguous sequences (test/ftest.c:42-47)
            _check(S, O, fam, d, lot, 1, n, lot)        # interleaved (test/ftest.c:64), cfft2f_'s first sweep
            _check(S, O, fam, d, lot, n + 3, n, 1)      # padded rows
            _check(S, O, fam, d, lot, 1, n, lot + 2)    # padded interleave
            _check(S, O, fam, d, lot, 2 * n + 1, n, 2)  # both strides non-unit


def test_config4_lengths(libs):
    S, O = libs
    for fam, n in (("cost", 1001), ("sint", 1000), ("cosq", 1000), ("cosq", 1001), ("cost", 1000), ("sint", 1001)):
        for d in "fb":
            _check(S, O, fam, d, 3, n, n, 1, seed=5)


def test_headline_lengths(libs):
    S, O = libs
    for fam in ("cfft", "rfft"):
        for d in "fb":
            _check(S, O, fam, d, 3, 4096, 4096, 1)
            _check(S, O, fam, d, 2, 1024, 1024, 1)
            _check(S, O, fam, d, 2, 1, 4096, 2)  # same length through the strided engine


def test_four_step_long_complex(libs):
    S, O = libs
    nmax = fl.sim().cfb200_max_onchip_complex()
    for n in (8192 * 2, 7 * 11 * 13 * 9, 2 * nmax + 2):
        if n <= nmax:
            continue
        for d in "fb":
            _check(S, O, "cfft", d, 2, n, n, 1)
    _check(S, O, "cfft", "f", 2, 1, 8192, 2)  # interleaved long: pow2 kernel not applicable -> four-step
    for d in "fb":  # batch-contiguous four-step (rows enumerate the batch index fastest), cfft2f_'s first sweep
        _check(S, O, "cfft", d, 3, 1, 16384, 3)
        _check(S, O, "cfft", d, 5, 1, 9000, 7)


def test_large_prime_factor_uses_chirp_z(libs):
    """a prime factor beyond one CTA and beyond the four-step split (4831 is prime): Bluestein on power-of-two transforms"""
    S, O = libs
    nmax = fl.sim().cfb200_max_onchip_complex()
    assert 4831 > nmax
    for d in "fb":
        _check(S, O, "cfft", d, 2, 4831, 4831, 1)
        _check(S, O, "cfft", d, 2, 1, 4831, 2)
    _check(S, O, "rfft", "f", 2, 2 * 4831, 2 * 4831, 1)  # long real path -> complex length with the same prime


def test_longest_on_chip_complex_lengths(libs):
    """lengths just below the single-CTA limit (padded rows no longer fit and are dropped), incl. a prime"""
    S, O = libs
    nmax = fl.sim().cfb200_max_onchip_complex()
    for n in (4000, 4096 + 512, nmax - 1, nmax):
        inc = 2 if n == 4000 else 1
        _check(S, O, "cfft", "f", 2, n * inc + 3, n, inc)


def test_cfft2(libs):
    S, O = libs
    for (ldim, l, m) in ((8, 8, 6), (11, 8, 6), (64, 64, 64), (130, 128, 96)):
        c = fl.rand_input("cfft", ldim * m, l + m)
        for d in "fb":
            a, ia = S.run2(d, ldim, l, m, c)
            b, ib = O.run2(d, ldim, l, m, c)
            assert ia == ib == 0
            assert fl.rel_l2(a, b) <= fl.tol(l * m), (d, ldim, l, m)
            if ldim > l:
                pad = np.ones(ldim * m, bool)
                for j in range(m):
                    pad[j * ldim: j * ldim + l] = False
                assert np.array_equal(a[pad], c[pad])


def test_rfft2(libs):
    S, O = libs
    for (ldim, l, m) in ((1, 1, 1), (1, 1, 4), (4, 4, 1), (2, 2, 2), (3, 3, 3), (8, 8, 6), (11, 8, 6), (7, 7, 5), (9, 7, 5),
                         (64, 64, 48), (33, 30, 21), (130, 128, 96), (100, 100, 60)):
        r = fl.rand_input("rfft", ldim * (m - 1) + l, 3 * l + m)
        for d in "fb":
            a, ia = S.run2r(d, ldim, l, m, r)
            b, ib = O.run2r(d, ldim, l, m, r)
            assert ia == ib == 0
            assert fl.rel_l2(fl.rows2(a, ldim, l, m), fl.rows2(b, ldim, l, m)) <= fl.tol(l * m), (d, ldim, l, m)
            if ldim > l:  # rows l..ldim-1 are not ours to touch (the reference scribbles on them, fftpack.c:13407)
                pad = np.ones(len(r), bool)
                for j in range(m):
                    pad[j * ldim: j * ldim + l] = False
                assert np.array_equal(a[pad], r[pad])


def test_option_convolution_pipeline(libs):
    """payoff -> rfftmf -> characteristic function -> rfftmb -> value, batched (test/vargamma.c:42-106)"""
    for n in (128, 1000):
        val, N, ier = fl.option_product(fl.sim(), n, fl.OPTION_CASES)
        want = np.array([fl.option_oracle(n, c) for c in fl.OPTION_CASES])
        assert ier == 0 and N == fl.oracle().orc_next_fast_even_size(n)
        assert np.max(np.abs(val - want) / np.abs(want)) <= 1e-12


def test_l2_object_api(libs):
    """include/cfftpack_b200_l2.h: the reference's fft_t API served by the library, vs the oracle's cfftpack.c"""
    fl.l2_compare(fl.sim(), (1, 2, 3, 4, 5, 8, 16, 30, 31, 60, 100, 1000, 1001))
    fl.l2_batch_check(fl.sim())


def test_rfft_stream_kernel_both_output_paths():
    """CFB200_R2C_BULK=0 keeps the per-thread global stores of the real stream kernel, the default drains finished rows
    with bulk shared->global copies; env is read once per process, hence the subprocesses"""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import fftlibs as fl; "
            "S = fl.Lib(fl.sim()); O = fl.Lib(fl.oracle(), 'orc_'); "
            "x = fl.rand_input('rfft', 3 * 4102, 5); "
            "r = [fl.rel_l2(S.runm('rfft', d, 3, 4102, 4096, 1, x)[0], O.runm('rfft', d, 3, 4102, 4096, 1, x)[0]) for d in 'fb']; "
            "y = fl.rand_input('rfft', 5 * 256, 6); "
            "r += [fl.rel_l2(S.runm('rfft', d, 5, 256, 256, 1, y)[0], O.runm('rfft', d, 5, 256, 256, 1, y)[0]) for d in 'fb']; "
            "assert max(r) < 1e-14, r" % os.path.dirname(os.path.abspath(__file__)))
    for v in ("0", "1"):
        out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, CFB200_R2C_BULK=v), capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-2000:]


def _device_batch(S, O, fam, d, lot, n, seed):
    """contiguous, 16-byte aligned batch marked as DEVICE memory, so that the aligned fast paths (bulk copies) are taken"""
    import ctypes
    lib = S.lib
    lib.cfb200_sim_mark_device.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    x = fl.rand_input(fam, n * lot, seed)
    buf = np.zeros(n * lot + 2)
    off = (16 - buf.ctypes.data % 16) % 16 // 8
    y = buf[off:off + n * lot]
    y[:] = x
    lib.cfb200_sim_mark_device(ctypes.c_void_p(y.ctypes.data), y.nbytes)
    ws, _ = S.init(fam, n, multi=True)
    ier, I = ctypes.c_int(-1), ctypes.c_int
    getattr(lib, fam + "m" + d + "_")(ctypes.byref(I(lot)), ctypes.byref(I(n)), ctypes.byref(I(n)), ctypes.byref(I(1)), fl.P(y),
                                      ctypes.byref(I(n * lot)), fl.P(ws), ctypes.byref(I(fl.lensav(fam, n))), fl.P(np.zeros(8)),
                                      ctypes.byref(I(fl.lenwrk(fam, n, lot))), ctypes.byref(ier))
    want, ib = O.runm(fam, d, lot, n, n, 1, x)
    assert ier.value == 0 and ib == 0, (fam, d, n, ier.value, ib)
    return max(fl.rel_l2(y[i * n:(i + 1) * n], want[i * n:(i + 1) * n]) for i in range(lot)), buf


def test_mixed_radix_stream_kernel_all_families_and_lengths(libs):
    """mixed.cuh: underlying real lengths 999 = 9*3*37, 1000 = 10^3, 1001 = 13*11*7, 1002 = 6*167 (even and odd, register and
    generic last pass) for rfft / cosq / sint / cost, both directions, even and odd lots (the odd row takes the engine)"""
    S, O = libs
    for M in (999, 1000, 1001, 1002):
        for fam, n in (("rfft", M), ("cosq", M), ("sint", M - 1), ("cost", M + 1)):
            for d in "fb":
                for lot in (2, 3):
                    before = S.lib.cfb200_launch_count()
                    err, _ = _device_batch(S, O, fam, d, lot, n, 7 * n + lot)
                    assert err <= fl.tol(n), (fam, d, n, lot, err)
                    assert S.lib.cfb200_launch_count() - before == (1 if lot % 2 == 0 else 2), (fam, n, lot)


def test_long_power_of_two_beyond_the_four_step(libs):
    """2^21 points: six-step over one tile sweep and a four-step pair (dispatch.cu run_c2c_long_pow2)"""
    S, O = libs
    n = 1 << 21
    x = fl.rand_input("cfft", n, 3)
    for d in "fb":
        a, ia = S.run1("cfft", d, n, x)
        b, ib = O.run1("cfft", d, n, x)
        assert ia == ib == 0
        assert fl.rel_l2(a, b) <= fl.tol(n), d


def test_pipelined_host_staging(libs):
    """pinned host arrays go through HBM in lot-chunks on three streams; chunk size forced small here"""
    import ctypes
    import subprocess
    import sys
    code = """
import sys, ctypes, numpy as np
sys.path.insert(0, %r)
import fftlibs as fl
S, O = fl.Lib(fl.sim()), fl.Lib(fl.oracle(), 'orc_')
for fam, n, lot, jump in (('cfft', 64, 37, 64), ('rfft', 100, 29, 103), ('cost', 33, 50, 40), ('cfft', 128, 9, 128)):
    x = fl.rand_input(fam, (lot - 1) * jump + n, 5)
    fl.sim().cfb200_sim_mark_pinned(fl.P(x), ctypes.c_size_t(0))
    want, ib = O.runm(fam, 'f', lot, jump, n, 1, x, lenx=len(x))
    y = np.array(x, copy=True)
    fl.sim().cfb200_sim_mark_pinned(fl.P(y), ctypes.c_size_t(y.nbytes))
    ws, _ = S.init(fam, n, multi=True)
    ier = ctypes.c_int(-1); I = ctypes.c_int
    wk = np.zeros(8)
    getattr(fl.sim(), fam + 'mf_')(ctypes.byref(I(lot)), ctypes.byref(I(jump)), ctypes.byref(I(n)), ctypes.byref(I(1)), fl.P(y),
        ctypes.byref(I(len(y))), fl.P(ws), ctypes.byref(I(fl.lensav(fam, n))), fl.P(wk), ctypes.byref(I(fl.lenwrk(fam, n, lot))), ctypes.byref(ier))
    assert ier.value == 0 == ib, (fam, ier.value)
    assert fl.rel_l2(y, want) <= fl.tol(n), (fam, n, fl.rel_l2(y, want))
print('ok')
""" % os.path.join(fl.ROOT, "tests")
    env = dict(os.environ, CFB200_PIPE_CHUNK_KB="4")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_long_real_path_forced_on_short_sequences(libs):
    """real-family sequences longer than one CTA go through global scratch arrays around the long complex transform;
    CFB200_LONG_REAL_MIN lowers the switch-over so the emulator can cover it"""
    import subprocess
    import sys
    code = """
import sys, numpy as np
sys.path.insert(0, %r)
import fftlibs as fl
S, O = fl.Lib(fl.sim()), fl.Lib(fl.oracle(), 'orc_')
for fam in ('rfft', 'cost', 'sint', 'cosq', 'sinq'):
    for n, lot, jump, inc in ((24, 5, 24, 1), (45, 4, 1, 4), (64, 3, 70, 1), (100, 2, 203, 2)):
        span = (lot - 1) * jump + (n - 1) * inc + 1
        x = fl.rand_input(fam, span, 9)
        for d in 'fb':
            a, ia = S.runm(fam, d, lot, jump, n, inc, x, lenx=span)
            b, ib = O.runm(fam, d, lot, jump, n, inc, x, lenx=span)
            assert ia == ib == 0, (fam, d, n, ia, ib)
            assert fl.rel_l2(a, b) <= fl.tol(n), (fam, d, n, lot, fl.rel_l2(a, b))
print('ok')
""" % os.path.join(fl.ROOT, "tests")
    env = dict(os.environ, CFB200_LONG_REAL_MIN="20")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stdout[-500:] + out.stderr[-2000:]


def test_randomized_shapes_small(libs):
    """the same randomized sweep as the GPU test, at sizes the emulator handles quickly"""
    S, O = libs
    rng = np.random.default_rng(42)
    lengths = [2, 3, 4, 5, 6, 7, 9, 10, 12, 14, 15, 18, 21, 25, 27, 33, 36, 45, 49, 50, 63, 64, 70, 81, 90, 98, 100, 121, 128, 143,
               169, 180, 200, 243, 256, 300, 343, 360, 512, 625, 1000, 1024]
    for case in range(60):
        fam = fl.FAMILIES[int(rng.integers(len(fl.FAMILIES)))]
        n = int(lengths[int(rng.integers(len(lengths)))])
        lot = int(rng.choice([1, 2, 3, 5, 8, 9, 16, 17]))
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
        _check(S, O, fam, "fb"[int(rng.integers(2))], lot, jump, n, inc, seed=case)


def _device_array(lib, count):
    """16-byte aligned complex128 array registered as device memory with the emulator"""
    import ctypes
    lib.cfb200_sim_mark_device.argtypes = [ctypes.c_void_p, ctypes.c_size_t]
    buf = np.zeros(2 * count + 2)
    off = (16 - buf.ctypes.data % 16) % 16 // 8
    a = buf[off:off + 2 * count].view(np.complex128)
    lib.cfb200_sim_mark_device(ctypes.c_void_p(a.ctypes.data), a.nbytes)
    return a, buf


def test_sharded_phases_two_ranks_emulated_in_one_process(libs):
    """The multi-GPU phase entry points with G = 2 "ranks" run one after the other in this process: every
    rank's slabs are ordinary arrays, the peer pointers point at the other rank's arrays, and a barrier is simply the end
    of the loop over ranks.  Checks the slab/peer index arithmetic of the fused transposes, the W_N^(i b) twiddles and the
    natural-order scatter of the long 1-D transform against the oracle -- no GPU, no NCCL."""
    import ctypes
    S, O = libs
    lib = S.lib
    I, G = ctypes.c_int, 2

    def phase(fn, *args):
        ier = I(-1)
        fn(*args, ctypes.byref(ier))
        assert ier.value == 0, lib.cfb200_last_error()

    # ---- 2^24-point transform in natural order over two ranks: X = data chunks, Y = second buffer (result)
    a = 24
    n = 1 << a
    x = fl.rand_input("cfft", n, 99)
    X, Y, keep = [], [], []
    for r in range(G):
        xa, kb = _device_array(lib, n // G); keep.append(kb)
        ya, kb = _device_array(lib, n // G); keep.append(kb)
        xa[:] = x[r * (n // G):(r + 1) * (n // G)]
        X.append(xa); Y.append(ya)
    px = (ctypes.c_void_p * G)(*[v.ctypes.data for v in X])
    py = (ctypes.c_void_p * G)(*[v.ctypes.data for v in Y])
    f1 = lib.cfb200_cfft1_sharded_phase
    for ph, src, dst in ((0, X, py), (1, Y, px), (2, X, py)):
        for r in range(G):  # "barrier" = all ranks finish the phase before the next one starts
            phase(f1, I(ph), I(-1), I(a), I(r), I(G), ctypes.c_void_p(src[r].ctypes.data), dst)
    want, ier = O.run1("cfft", "f", n, x)
    got = np.concatenate(Y)
    assert fl.rel_l2(got, want) <= fl.tol(n)

    # ---- 4096 x 4096 cfft2f on column slabs C[m_loc][l], row slabs D[m][l_loc] (another minute of emulation: opt-in;
    # the same phases run on real GPUs in bench.py's `cfft2` object and in the 2-GPU tests)
    if os.environ.get("CFB200_SLOW_TESTS") != "1":
        return
    l = m = 4096
    c = fl.rand_input("cfft", l * m, 7)
    C, D = [], []
    for r in range(G):
        ca, kb = _device_array(lib, l * m // G); keep.append(kb)
        da, kb = _device_array(lib, l * m // G); keep.append(kb)
        ca[:] = c[r * (l * m // G):(r + 1) * (l * m // G)]
        C.append(ca); D.append(da)
    pc = (ctypes.c_void_p * G)(*[v.ctypes.data for v in C])
    pd = (ctypes.c_void_p * G)(*[v.ctypes.data for v in D])
    f2 = lib.cfb200_cfft2_sharded_phase
    for ph, src, dst in ((1, C, pd), (2, D, pc)):
        for r in range(G):
            phase(f2, I(ph), I(-1), I(l), I(m), I(r), I(G), ctypes.c_void_p(src[r].ctypes.data), dst)
    want2, ier = O.run2("f", l, l, m, c)
    assert fl.rel_l2(np.concatenate(C), want2) <= fl.tol(l * m)
