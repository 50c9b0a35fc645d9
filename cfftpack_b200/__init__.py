"""cfftpack_b200 -- thin Python loader for libcfftpack_b200.so (the product is the C-ABI library).

The library exports the reference's FFTPACK entry points (include/cfftpack_b200.h, mirroring
cfftpack/fftpack.h:80-169 of zywina/cfftpack).  This module only locates it and offers a small
convenience caller used by tests/ and bench.py; it contains no transform code and no fallback:
importing it without the built library raises.
"""
import ctypes
import math
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcfftpack_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C cfftpack_b200/csrc` (python __graft_entry__.py). "
        "cfftpack_b200 has no CPU or PyTorch fallback.")

lib = ctypes.CDLL(LIB_PATH)
lib.cfb200_launch_count.restype = ctypes.c_ulonglong
lib.cfb200_last_error.restype = ctypes.c_char_p
lib.cfb200_version.restype = ctypes.c_char_p
lib.cfb200_set_stream.argtypes = [ctypes.c_void_p]

_I = ctypes.c_int


def il2(n):
    """the reference's literal (int)(log((double)n)/log(2.0)) (fftpack.c:2221)"""
    return int(math.log(float(n)) / math.log(2.0))


def lensav(fam, n):
    if fam == "rfft":
        return n + il2(n) + 4
    if fam == "sint":
        return n // 2 + n + il2(n) + 4
    return 2 * n + il2(n) + 4


def lenwrk(fam, n, lot=None):
    if lot is None:
        return {"cfft": 2 * n, "rfft": n, "cost": max(n - 1, 1), "sint": 2 * n + 2}.get(fam, n)
    return {"cfft": 2 * lot * n, "rfft": lot * n, "cost": lot * (n + 1), "sint": lot * (2 * n + 4)}.get(fam, lot * n)


def version():
    return lib.cfb200_version().decode()


def launch_count():
    return int(lib.cfb200_launch_count())


def last_error():
    return lib.cfb200_last_error().decode()


def set_stream(cuda_stream_handle):
    """cuda_stream_handle: integer cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); 0 = default"""
    lib.cfb200_set_stream(ctypes.c_void_p(cuda_stream_handle))


def synchronize():
    return lib.cfb200_synchronize()


class Plan:
    """Caller-side state of one (family, n): the host wsave array, exactly as a C caller would keep it."""

    def __init__(self, fam, n):
        import numpy as np
        self.fam, self.n = fam, n
        self.lensav = lensav(fam, n)
        self.wsave = np.zeros(self.lensav + 8)
        ier = _I(-1)
        getattr(lib, fam + "mi_")(ctypes.byref(_I(n)), self.wsave.ctypes.data_as(ctypes.c_void_p),
                                  ctypes.byref(_I(self.lensav)), ctypes.byref(ier))
        if ier.value != 0:
            raise RuntimeError(f"{fam}mi_ n={n}: ier={ier.value}")
        self._wp = self.wsave.ctypes.data_as(ctypes.c_void_p)
        self._dummy = ctypes.c_double(0.0)

    def multi(self, direction, ptr, lot, jump, inc, lenx):
        """Call <fam>m<f|b>_ on the array at address `ptr` (host or device).  Returns ier."""
        ier = _I(-1)
        lw = lenwrk(self.fam, self.n, lot)
        getattr(lib, self.fam + "m" + direction + "_")(
            ctypes.byref(_I(lot)), ctypes.byref(_I(jump)), ctypes.byref(_I(self.n)), ctypes.byref(_I(inc)),
            ctypes.c_void_p(ptr), ctypes.byref(_I(lenx)), self._wp, ctypes.byref(_I(self.lensav)),
            ctypes.byref(self._dummy), ctypes.byref(_I(min(lw, 2**31 - 1))), ctypes.byref(ier))
        return ier.value


def option_convolution(n, S, K, sigma, theta, kappa, t, r, call=True, black_scholes=False):
    """Value a batch of options by frequency-domain convolution: the reference's test/vargamma.c:42-106
    (conv_bsvg_option) for all options in one device-resident pipeline (cfb200_option_convolution).
    Scalars broadcast against arrays.  Returns (values, N) with N the grid size actually used."""
    import numpy as np
    cols = np.broadcast_arrays(*(np.asarray(v, dtype=np.float64) for v in (S, K, sigma, theta, kappa, t, r)),
                               np.asarray(call, dtype=bool), np.asarray(black_scholes, dtype=bool))
    lot = max(1, cols[0].size)
    flat = [np.ascontiguousarray(c.reshape(-1), dtype=np.float64) for c in cols[:7]]
    flags = np.ascontiguousarray(cols[7].reshape(-1).astype(np.int32) | (cols[8].reshape(-1).astype(np.int32) << 1))
    values = np.zeros(lot)
    ier = _I(-1)
    vp = ctypes.c_void_p
    N = lib.cfb200_option_convolution(_I(lot), _I(n), *(vp(c.ctypes.data) for c in flat), vp(flags.ctypes.data),
                                      vp(values.ctypes.data), ctypes.byref(ier))
    if ier.value != 0:
        raise RuntimeError(f"cfb200_option_convolution: ier={ier.value} ({last_error()})")
    return values.reshape(cols[0].shape), N
