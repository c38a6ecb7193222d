import sys, torch
sys.path.insert(0, "/root/repo")
import torch.nn.functional as F
from vdm4cdm_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
b, ci, co, d, h, w = 1, 32, 32, 8, 16, 16
x = torch.randint(-2, 3, (b, ci, d, h, w), generator=g).float().to(dev)
wt = torch.randint(-1, 2, (co, ci, 3, 3, 3), generator=g).float().to(dev)
dy = torch.randint(-1, 2, (b, co, d, h, w), generator=g).float().to(dev)
xg = x.double().clone().requires_grad_(True)
F.conv3d(F.pad(xg, (1,)*6, mode="circular"), wt.double()).backward(dy.double())
ref = xg.grad.round().float()
dyp = ops.pad_circular(ops.to_planar(dy), co)
dx = ops.from_planar(ops.conv3d(dyp, ops.pack_conv_weight(wt, transpose_flip=True), ci, circular=True), ci)
err = (dx - ref.to(torch.bfloat16).float()).abs()
print("max err", err.max().item(), "ref max", ref.abs().max().item(), "n bad", (err > 0).sum().item(), "of", err.numel())
bad = (err > 0).nonzero()
print(bad[:10].tolist())
print("bad by d:", [(err[0, :, i] > 0).sum().item() for i in range(d)])
print("bad by h:", [(err[0, :, :, i] > 0).sum().item() for i in range(h)])
print("bad by w:", [(err[0, :, :, :, i] > 0).sum().item() for i in range(w)])
# zero-padded dgrad for comparison
xg2 = x.double().clone().requires_grad_(True)
F.conv3d(xg2, wt.double(), padding=1).backward(dy.double())
dx0 = ops.from_planar(ops.conv3d(ops.to_planar(dy), ops.pack_conv_weight(wt, transpose_flip=True), ci), ci)
print("zero-pad dgrad max err", (dx0 - xg2.grad.round().float().to(torch.bfloat16).float()).abs().max().item())
for (b, ci, co, d, h, w) in [(1, 32, 32, 8, 16, 16), (2, 32, 32, 8, 16, 16), (1, 64, 64, 6, 16, 8)]:
    x = torch.randint(-2, 3, (b, ci, d, h, w), generator=g).float().to(dev)
    dy = torch.randint(-1, 2, (b, co, d, h, w), generator=g).float().to(dev)
    wg = torch.zeros((co, ci, 3, 3, 3), device=dev, dtype=torch.float64, requires_grad=True)
    F.conv3d(F.pad(x.double(), (1,)*6, mode="circular"), wg).backward(dy.double())
    xpad = ops.pad_circular(ops.to_planar(x), ci)
    dw = ops.wgrad_to_torch(ops.conv3d_wgrad(xpad, ops.to_planar(dy, 16), ci, co, 3, a_padded=True), 3)
    err = (dw - wg.grad.round().float()).abs()
    print((b, ci, co, d, h, w), "wgrad max err", err.max().item(), "ref max", wg.grad.abs().max().item(),
          "per kd", [err[:, :, k].max().item() for k in range(3)], "per kh", [err[:, :, :, k].max().item() for k in range(3)],
          "per kw", [err[..., k].max().item() for k in range(3)])
