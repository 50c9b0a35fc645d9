"""Multi-GPU check + timing of the sharded 2-D transform: NCCL all-to-all path vs fused P2P-store path.
torchrun --nproc-per-node G tools/run_dist2d.py [l]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import cfftpack_b200 as cb
from cfftpack_b200.dist import Cfft2Sharded, Cfft2ShardedP2P

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
l = m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m_loc = m // world
g = torch.Generator(device="cuda").manual_seed(5 + rank)
x = torch.view_as_complex(torch.rand(m_loc, l, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
a = Cfft2Sharded(l, m)
ya = x.clone(); a.forward(ya)
b = Cfft2ShardedP2P(l, m)
b.slab.copy_(x); b.forward()
torch.cuda.synchronize()
err = float((torch.view_as_real(b.slab) - torch.view_as_real(ya)).norm() / torch.view_as_real(ya).norm())
b.backward(); torch.cuda.synchronize()
rt = float((torch.view_as_real(b.slab) - torch.view_as_real(x)).norm() / torch.view_as_real(x).norm())
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
t_nccl = timeit(lambda: a.forward(ya))
t_p2p = timeit(lambda: b.forward())
if rank == 0:
    print(json.dumps({"l": l, "m": m, "gpus": world, "p2p_vs_nccl_rel_err": err, "p2p_roundtrip_rel_err": rt,
                      "ms_nccl_alltoall": t_nccl, "ms_fused_p2p": t_p2p,
                      "nvlink_bytes_per_gpu_per_exchange": 16 * l * m * (world - 1) // (world * world)}), flush=True)
dist.destroy_process_group()
