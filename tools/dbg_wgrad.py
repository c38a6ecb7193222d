import sys, torch
sys.path.insert(0, "/root/repo")
import torch.nn.functional as F
from vdm4cdm_b200 import ops
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(1)
for (b, ci, co, d, h, w) in [(1, 32, 32, 4, 16, 8), (1, 32, 32, 8, 16, 16), (1, 16, 32, 5, 16, 16)]:
    a = torch.randint(-2, 3, (b, ci, d, h, w), generator=gen).float().to(dev)
    g = torch.randint(-2, 3, (b, co, d, h, w), generator=gen).float().to(dev)
    wt = torch.zeros((co, ci, 3, 3, 3), device=dev, dtype=torch.float64, requires_grad=True)
    F.conv3d(a.double(), wt, padding=1).backward(g.double())
    ref = wt.grad.round().float()
    dw = ops.conv3d_wgrad(ops.to_planar(a, 16), ops.to_planar(g, 16), ci, co, 3)
    got = ops.wgrad_to_torch(dw, 3)
    torch.cuda.synchronize()
    err = (got - ref).abs()
    print((b, ci, co, d, h, w), "max err", err.max().item(), "ref max", ref.abs().max().item())
    print(" per (kd,kh,kw) max err:", [[ [round(err[:, :, kd, kh, kw].max().item()) for kw in range(3)] for kh in range(3)] for kd in range(3)])
    print(" per ci-block max err:", [round(err[:, c:c + 8].max().item()) for c in range(0, ci, 8)], " per co-block:", [round(err[c:c + 8].max().item()) for c in range(0, co, 8)])
    # is got a permutation of taps?
    for kw in range(3):
        for kw2 in range(3):
            if torch.equal(got[:, :, 1, 1, kw], ref[:, :, 1, 1, kw2]):
                print("  got kw", kw, "== ref kw", kw2)
