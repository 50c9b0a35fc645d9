"""Multi-GPU check + timing of the very long 1-D transform: N = 2^log2n complex points in natural
order over the ranks, four-step with fused P2P exchanges (cfftpack_b200.dist.Cfft1ShardedP2P).  Parity: sampled output
bins against the direct O(N) DFT sums of the definition (test/naivepack.c naive_fft; for N <= 2^24 also the CPU oracle).
    torchrun --nproc-per-node G tools/run_dist1d.py [log2n] [bins]"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import cfftpack_b200 as cb
from cfftpack_b200.dist import Cfft1ShardedP2P


def dft_bins(x_loc, first, n, bins, sign, world):
    """sum_j x[j] exp(sign 2 pi i j k / n) for k in bins, over the distributed x (this rank holds x[first : first+len])"""
    out = torch.zeros(len(bins), 2, device=x_loc.device, dtype=torch.float64)
    idx = torch.arange(first, first + x_loc.numel(), device=x_loc.device, dtype=torch.int64)
    step = 1 << 22
    for bi, k in enumerate(bins):
        acc = torch.zeros((), device=x_loc.device, dtype=torch.complex128)
        for s in range(0, x_loc.numel(), step):
            ph = ((idx[s:s + step] * k) % n).to(torch.float64) * (sign * 2.0 * math.pi / n)
            acc = acc + (x_loc[s:s + step] * torch.complex(torch.cos(ph), torch.sin(ph))).sum()
        out[bi, 0], out[bi, 1] = acc.real, acc.imag
    if world > 1:
        dist.all_reduce(out)
    return torch.view_as_complex(out)


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0))))
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    n = 1 << log2n
    plan = Cfft1ShardedP2P(log2n)
    g = torch.Generator(device="cuda").manual_seed(11 + rank)
    x0 = torch.view_as_complex(torch.rand(plan.n_loc, 2, generator=g, device="cuda", dtype=torch.float64) - 0.5)
    plan.x.copy_(x0)
    y = plan.forward()
    torch.cuda.synchronize()
    gen = torch.Generator().manual_seed(3)
    bins = sorted(set([0, 1, n // 2, n - 1, plan.n_loc - 1, plan.n_loc % n] + torch.randint(0, n, (nb,), generator=gen).tolist()))
    want = dft_bins(x0, rank * plan.n_loc, n, bins, -1.0, world) / n
    got = torch.zeros(len(bins), 2, device="cuda", dtype=torch.float64)
    for bi, k in enumerate(bins):
        if k // plan.n_loc == rank:
            v = y[k - rank * plan.n_loc]
            got[bi, 0], got[bi, 1] = v.real, v.imag
    dist.all_reduce(got)
    got = torch.view_as_complex(got)
    # error of the sampled bins relative to the RMS magnitude of the outputs (single bins can be arbitrarily small)
    rms = float(torch.sqrt((y.abs() ** 2).mean()))
    err_bins = float(((got - want).abs().max()) / rms)
    # round trip: backward(forward(x)) == x
    plan.x.copy_(y)
    z = plan.backward()
    torch.cuda.synchronize()
    rt = torch.tensor([float((torch.view_as_real(z) - torch.view_as_real(x0)).norm()), float(torch.view_as_real(x0).norm())],
                      device="cuda", dtype=torch.float64) ** 2
    dist.all_reduce(rt)
    rt = float((rt[0] / rt[1]).sqrt())
    plan.x.copy_(x0)
    for _ in range(2):
        plan.forward()
    torch.cuda.synchronize(); dist.barrier()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        plan.forward()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    link = 3 * 16 * n * (world - 1) // (world * world)  # three exchanges
    if rank == 0:
        print(json.dumps({"log2n": log2n, "gpus": world, "bins": len(bins), "max_bin_err_over_rms": err_bins,
                          "roundtrip_rel_err": rt, "bar": 1e-12 * log2n, "ms": ms,
                          "hbm_gbs_algorithmic": 2 * 16 * n / (ms * 1e-3) / 1e9,
                          "gflops_5nlogn": 5.0 * n * log2n / (ms * 1e-3) / 1e9,
                          "nvlink_bytes_per_gpu": link, "nvlink_gbs": link / (ms * 1e-3) / 1e9,
                          "frac_of_900": link / (ms * 1e-3) / 1e9 / 900.0}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
