"""Worker for tests/test_dist_gloo.py: world_size ranks on CPU (gloo), product sources under the thread emulator."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fftlibs as fl  # noqa: E402
from cfftpack_b200.dist import Cfft2Sharded, shard_lot  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    sim, orc = fl.sim(), fl.Lib(fl.oracle(), "orc_")
    # ---- lot sharding: every rank transforms its own contiguous lot range; the union equals the unsharded result
    n, lot = 60, 13
    x = fl.rand_input("cfft", n * lot, 42)
    m0, m1 = shard_lot(lot, rank, world)
    mine, ier = fl.Lib(sim).runm("cfft", "f", m1 - m0, n, n, 1, x[m0 * n:m1 * n])
    assert ier == 0
    parts = [None] * world
    dist.all_gather_object(parts, (m0, m1, mine))
    whole = np.concatenate([p[2] for p in sorted(parts)])
    want, _ = orc.runm("cfft", "f", lot, n, n, 1, x)
    assert fl.rel_l2(whole, want) <= fl.tol(n)
    assert sorted((p[0], p[1]) for p in parts)[0][0] == 0 and sorted((p[0], p[1]) for p in parts)[-1][1] == lot
    # ---- 2-D: column slabs + all-to-all transposes
    for (l, m) in ((16, 12), (64, 32), (12, 20)):
        c = fl.rand_input("cfft", l * m, l * m)
        for d in "fb":
            want, ier = orc.run2(d, l, l, m, c)
            assert ier == 0
            m_loc = m // world
            slab = torch.from_numpy(c.reshape(m, l)[rank * m_loc:(rank + 1) * m_loc].copy())
            plan = Cfft2Sharded(l, m, lib=sim)
            plan.transform(slab, d)
            got = slab.numpy().ravel()
            ref = want.reshape(m, l)[rank * m_loc:(rank + 1) * m_loc].ravel()
            e = fl.rel_l2(got, ref)
            assert e <= fl.tol(l * m), (rank, d, l, m, e)
    dist.barrier()
    if rank == 0:
        print("DIST_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
