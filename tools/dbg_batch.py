import sys, torch
sys.path.insert(0, "/root/repo")
from vdm4cdm_b200.networks import CUNet
from vdm4cdm_b200 import ops, _C
torch.manual_seed(0)
shape, chs = (1, 16, 16, 16), (16, 32)
net = CUNet(shape=shape, chs=chs, s_conditioning_channels=1, v_conditioning_dims=[], t_conditioning=True).cuda().eval()
g = torch.Generator().manual_seed(1)
x1 = torch.randn((1,) + shape, generator=g).cuda()
c1 = torch.randn((1,) + shape, generator=g).cuda()
x2 = torch.cat([x1, x1]); c2 = torch.cat([c1, c1])
t = torch.tensor([0.7]).cuda()
with torch.no_grad():
    r1 = net.chan_add_rows(1, t, None, x1.device)
    r2 = net.chan_add_rows(2, t.expand(2), None, x1.device)
    for k in r1:
        d = (r2[k][1] - r1[k][0]).abs().max().item()
        if d > 0:
            print("row", k, d, r1[k].abs().max().item())
    o1 = net(x1, t=t, s_conditioning=c1).clone()
    o2 = net(x2, t=t.expand(2), s_conditioning=c2).clone()
print("identical samples in one batch: o2[0] vs o2[1]", (o2[0] - o2[1]).abs().max().item(), " o2[1] vs o1", (o2[1] - o1[0]).abs().max().item(),
      "rel L2", ((o2[1] - o1[0]).norm() / o1.norm()).item())
bufs = net._arena.bufs
names = sorted(set(k.rsplit(".", 1)[0] for k in bufs if k.endswith(".1")))
for k in bufs:
    if k.endswith(".1") or ".1x" in k:
        k2 = k.replace(".1x", ".2x") if ".1x" in k else k[:-2] + ".2"
        if k2 in bufs and bufs[k].dtype != torch.float64:
            a, b = bufs[k].float(), bufs[k2].float()
            print(f"{k:28s} B1 vs B2[1]: max diff {(b[1:] - a).abs().max().item():.4e}  B2[0] vs B2[1] {(b[0] - b[1]).abs().max().item():.4e}  scale {a.abs().max().item():.3f}")
