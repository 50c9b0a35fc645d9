"""One warm + two profiled cfft2f_ calls on a device-resident l x m array (for an ncu launch list of the sweeps)."""
import sys, ctypes
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, fftlibs as fl, cfftpack_b200 as cb
l, m = int(sys.argv[1]), int(sys.argv[2])
c = torch.rand(l * m * 2, device="cuda", dtype=torch.float64) - 0.5
P = fl.Lib(fl.product())
ws, ls, ier = P.init2(l, m)
I = ctypes.c_int; ierc = I(-1); dummy = ctypes.c_double(0)
for _ in range(3):
    fl.product().cfft2f_(ctypes.byref(I(l)), ctypes.byref(I(l)), ctypes.byref(I(m)), ctypes.c_void_p(c.data_ptr()), fl.P(ws),
                         ctypes.byref(I(ls)), ctypes.byref(dummy), ctypes.byref(I(2 * l * m)), ctypes.byref(ierc))
    assert ierc.value == 0, cb.last_error()
torch.cuda.synchronize()
