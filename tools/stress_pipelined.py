"""Developer stress: host-array pipeline (3 streams, small chunks) vs the device-pointer path, repeated (CFB200_PIPE_CHUNK_KB=1024)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cfftpack_b200 as cb
bad = 0
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for fam, n, lot in (("cfft", 16384, 96), ("cfft", 10007, 64), ("rfft", 32768, 96), ("cost", 10001, 64), ("cfft", 4096, 512), ("rfft", 4096, 1024), ("cfft", 1000, 2000), ("cosq", 1001, 4000)):
    esz = 2 if fam == "cfft" else 1
    h = torch.empty(lot * n * esz, dtype=torch.float64, pin_memory=True).uniform_(-1, 1)
    d = h.cuda()
    plan = cb.Plan(fam, n)
    for rep in range(reps):
        hh = h.clone().pin_memory()
        dd = d.clone()
        assert plan.multi("f", dd.data_ptr(), lot, n, 1, lot * n) == 0, cb.last_error()
        cb.synchronize()
        assert plan.multi("f", hh.data_ptr(), lot, n, 1, lot * n) == 0, cb.last_error()
        ref_ = dd.cpu()
        diff = (hh - ref_)
        nbad = int((diff != 0).sum())
        same = nbad == 0 if fam == "cfft" else float(diff.norm() / ref_.norm()) <= 2e-13
        if not same:
            bad += 1
            idx = torch.nonzero(diff != 0).flatten()
            print("MISMATCH", fam, n, lot, "rep", rep, "bad elements", nbad, "first", int(idx[0]) // esz, "last", int(idx[-1]) // esz,
                  "max abs", float(diff.abs().max()), "seq range", int(idx[0]) // (n * esz), int(idx[-1]) // (n * esz), flush=True)
print("BAD", bad)
