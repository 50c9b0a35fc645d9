"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

 * Batched transforms (`cfftm*`, `rfftm*`, `costm*`, ...) shard by lot with NO data-path collective: `shard_lot`.
 * `Cfft2Sharded`: the reference's 2-D transform `cfft2f_/cfft2b_` (cfftpack/fftpack.c:2363, :2285 -- two `cfftmf_`
   sweeps, :2408-2426) on a matrix distributed as column slabs, with ONE exchange per dimension switch
   (SURVEY 8(e)): local length-l transforms down the columns, all-to-all transpose (NCCL over NVLink on GPUs),
   local length-m transforms along the rows, all-to-all back to the caller's slab layout.

The local transforms are calls into the C ABI (`cfftmf_`/`cfftmb_`); nothing here computes a transform.
"""
import ctypes

_I = ctypes.c_int


def shard_lot(lot, rank, world):
    """contiguous lot range [m0, m1) owned by `rank`; with jump = N this is a contiguous byte range"""
    return lot * rank // world, lot * (rank + 1) // world


class Cfft2Sharded:
    """c(l, m) column-major, rank g owns columns [g*m/G, (g+1)*m/G) as a contiguous slab `c_loc[m/G][l]`."""

    def __init__(self, l, m, group=None, lib=None):
        import numpy as np
        import torch.distributed as dist
        if lib is None:
            from . import lib as product_lib
            lib = product_lib
        self.lib, self.group = lib, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        G = self.world
        if l % G or m % G:
            raise ValueError(f"l={l} and m={m} must be multiples of the number of ranks {G}")
        self.l, self.m, self.l_loc, self.m_loc = l, m, l // G, m // G
        from . import lensav
        self.ws = {}
        for n in {l, m}:
            ls = lensav("cfft", n)
            ws = np.zeros(ls + 8)
            ier = _I(-1)
            lib.cfftmi_(ctypes.byref(_I(n)), ws.ctypes.data_as(ctypes.c_void_p), ctypes.byref(_I(ls)), ctypes.byref(ier))
            if ier.value:
                raise RuntimeError(f"cfftmi_ n={n}: ier={ier.value}")
            self.ws[n] = (ws, ls)
        self._send = self._recv = None

    def _cfftm(self, direction, ptr, lot, jump, n, inc, lenc):
        ws, ls = self.ws[n]
        ier, dummy = _I(-1), ctypes.c_double(0.0)
        getattr(self.lib, "cfftm" + direction + "_")(
            ctypes.byref(_I(lot)), ctypes.byref(_I(jump)), ctypes.byref(_I(n)), ctypes.byref(_I(inc)),
            ctypes.c_void_p(ptr), ctypes.byref(_I(lenc)), ws.ctypes.data_as(ctypes.c_void_p), ctypes.byref(_I(ls)),
            ctypes.byref(dummy), ctypes.byref(_I(min(2 * lot * n, 2**31 - 1))), ctypes.byref(ier))
        if ier.value:
            raise RuntimeError(f"cfftm{direction}_ lot={lot} n={n} inc={inc} jump={jump}: ier={ier.value}")

    def transform(self, c_loc, direction):
        """in place on the complex128 slab c_loc of shape [m_loc, l] (contiguous); direction 'f' or 'b'"""
        import torch
        import torch.distributed as dist
        G, l, m, l_loc, m_loc = self.world, self.l, self.m, self.l_loc, self.m_loc
        assert c_loc.dtype == torch.complex128 and c_loc.is_contiguous() and tuple(c_loc.shape) == (m_loc, l)
        if c_loc.is_cuda:  # stream order: library kernels, torch copies and the collectives all follow torch's stream
            self.lib.cfb200_set_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        # 1. columns: lot = m_loc sequences of length l, contiguous
        self._cfftm(direction, c_loc.data_ptr(), m_loc, l, l, 1, m_loc * l)
        if G == 1:
            # rows on the same slab: lot = l, jump = 1, n = m, inc = l (the reference's first sweep, fftpack.c:2408)
            self._cfftm(direction, c_loc.data_ptr(), l, 1, m, l, m * l)
            return c_loc
        if self._send is None or self._send.device != c_loc.device:
            self._send = torch.empty(G, m_loc, l_loc, dtype=torch.complex128, device=c_loc.device)
            self._recv = torch.empty(G, m_loc, l_loc, dtype=torch.complex128, device=c_loc.device)
        send, recv = self._send, self._recv
        # 2. transpose exchange: block (row slab r, my columns) -> rank r
        send.copy_(c_loc.view(m_loc, G, l_loc).permute(1, 0, 2))
        dist.all_to_all_single(torch.view_as_real(recv).view(-1), torch.view_as_real(send).view(-1), group=self.group)
        # recv is d(l_loc, m) column-major: element (i_loc, j) at j*l_loc + i_loc
        # 3. rows: lot = l_loc, jump = 1, n = m, inc = l_loc
        self._cfftm(direction, recv.data_ptr(), l_loc, 1, m, l_loc, m * l_loc)
        # 4. back to column slabs
        dist.all_to_all_single(torch.view_as_real(send).view(-1), torch.view_as_real(recv).view(-1), group=self.group)
        c_loc.view(m_loc, G, l_loc).copy_(send.permute(1, 0, 2))
        return c_loc

    def forward(self, c_loc):
        return self.transform(c_loc, "f")

    def backward(self, c_loc):
        return self.transform(c_loc, "b")


class Cfft2ShardedP2P:
    """Same transform as Cfft2Sharded, with the transposes FUSED into the FFT kernels: the last pass of each dimension
    stores every result straight into the slab of the GPU that owns it (P2P stores over NVLink/NVSwitch on
    peer-mapped symmetric memory), so there is no pack kernel and no NCCL all-to-all.  One stream-ordered barrier per
    phase.  Needs power-of-two l, m in 2^12..2^20 and all ranks on one node.

    `slab` is this rank's column slab C[m_loc][l] (complex128) living in symmetric memory: fill it, call
    forward()/backward(), read the result from it."""

    def __init__(self, l, m, group=None, lib=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if lib is None:
            from . import lib as product_lib
            lib = product_lib
        self.lib = lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        G = self.world
        if l % G or m % G:
            raise ValueError(f"l={l} and m={m} must be multiples of the number of ranks {G}")
        self.l, self.m, self.l_loc, self.m_loc = l, m, l // G, m // G
        dev = torch.device("cuda", torch.cuda.current_device())
        # complex128 slabs as float64 pairs (symmetric memory allocations are dtype-agnostic byte buffers)
        self._c = symm_mem.empty(self.m_loc * l * 2, dtype=torch.float64, device=dev)
        self._d = symm_mem.empty(m * self.l_loc * 2, dtype=torch.float64, device=dev)
        self._hc = symm_mem.rendezvous(self._c, self.group)
        self._hd = symm_mem.rendezvous(self._d, self.group)
        self.slab = torch.view_as_complex(self._c.view(self.m_loc, l, 2))
        self._c_ptrs = (ctypes.c_void_p * G)(*[int(p) for p in self._hc.buffer_ptrs])
        self._d_ptrs = (ctypes.c_void_p * G)(*[int(p) for p in self._hd.buffer_ptrs])

    def _phase(self, phase, direction, src, peers):
        ier = _I(-1)
        self.lib.cfb200_cfft2_sharded_phase(_I(phase), _I(-1 if direction == "f" else 1), _I(self.l), _I(self.m),
                                            _I(self.rank), _I(self.world), ctypes.c_void_p(src), peers, ctypes.byref(ier))
        if ier.value:
            from . import last_error
            raise RuntimeError(f"cfb200_cfft2_sharded_phase({phase}): ier={ier.value}: {last_error()}")

    def transform(self, direction):
        import torch
        self.lib.cfb200_set_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._hd.barrier(channel=0)   # every rank is done reading its row slab from the previous call
        self._phase(1, direction, self._c.data_ptr(), self._d_ptrs)
        self._hd.barrier(channel=0)   # all row slabs are complete
        self._phase(2, direction, self._d.data_ptr(), self._c_ptrs)
        self._hc.barrier(channel=0)   # all column slabs are complete
        return self.slab

    def forward(self):
        return self.transform("f")

    def backward(self):
        return self.transform("b")


class Cfft1ShardedP2P:
    """Very long 1-D complex transform (cfft1f_/cfft1b_ semantics, cfftpack/fftpack.c:2199, :2151) of N = 2^log2n points
    distributed in natural order over the ranks of one node: four-step N = L * Mm with the three
    exchanges fused into the kernels as P2P stores (`cfb200_cfft1_sharded_phase`), one stream-ordered barrier per phase.

    `x` is this rank's chunk x[rank N/G : (rank+1) N/G] (complex128, symmetric memory): fill it, call forward() or
    backward(); the result -- natural order, this rank's chunk of X -- is returned as a view of the second buffer."""

    def __init__(self, log2n, group=None, lib=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if lib is None:
            from . import lib as product_lib
            lib = product_lib
        self.lib = lib
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.log2n, self.n = log2n, 1 << log2n
        if self.n % self.world:
            raise ValueError("N must be a multiple of the number of ranks")
        self.n_loc = self.n // self.world
        dev = torch.device("cuda", torch.cuda.current_device())
        self._x = symm_mem.empty(self.n_loc * 2, dtype=torch.float64, device=dev)
        self._y = symm_mem.empty(self.n_loc * 2, dtype=torch.float64, device=dev)
        self._hx = symm_mem.rendezvous(self._x, self.group)
        self._hy = symm_mem.rendezvous(self._y, self.group)
        self.x = torch.view_as_complex(self._x.view(self.n_loc, 2))
        self.y = torch.view_as_complex(self._y.view(self.n_loc, 2))
        G = self.world
        self._x_ptrs = (ctypes.c_void_p * G)(*[int(p) for p in self._hx.buffer_ptrs])
        self._y_ptrs = (ctypes.c_void_p * G)(*[int(p) for p in self._hy.buffer_ptrs])

    def _phase(self, phase, direction, src, peers):
        ier = _I(-1)
        self.lib.cfb200_cfft1_sharded_phase(_I(phase), _I(-1 if direction == "f" else 1), _I(self.log2n), _I(self.rank),
                                            _I(self.world), ctypes.c_void_p(src), peers, ctypes.byref(ier))
        if ier.value:
            from . import last_error
            raise RuntimeError(f"cfb200_cfft1_sharded_phase({phase}): ier={ier.value}: {last_error()}")

    def transform(self, direction):
        import torch
        self.lib.cfb200_set_stream(ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        self._hy.barrier(channel=0)   # every rank is done with Y from the previous call
        self._phase(0, direction, self._x.data_ptr(), self._y_ptrs)
        self._hy.barrier(channel=0)   # all row slabs are complete
        self._phase(1, direction, self._y.data_ptr(), self._x_ptrs)
        self._hx.barrier(channel=0)   # all column slabs are complete (and every Y has been consumed)
        self._phase(2, direction, self._x.data_ptr(), self._y_ptrs)
        self._hy.barrier(channel=0)   # the natural-order result is complete
        return self.y

    def forward(self):
        return self.transform("f")

    def backward(self):
        return self.transform("b")
