import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import torch, cfftpack_b200 as cb
fam, n, lot = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
x = torch.rand(lot * n, device='cuda', dtype=torch.float64) - 0.5
plan = cb.Plan(fam, n)
for _ in range(3):
    assert plan.multi('f', x.data_ptr(), lot, n, 1, lot * n) == 0
torch.cuda.synchronize()
