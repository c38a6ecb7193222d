"""Mode counts of vdm_pk against the integer wave-vector shells, repeated (determinism check)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vdm4cdm_b200 import utils
for N in (128, 64, 96):
    x = torch.zeros((1, 1, N, N, N), device="cuda"); x[0, 0, 3, 5, 7] = 1.0
    kk = torch.fft.fftfreq(N, 1.0 / N)
    kmag = torch.sqrt(kk[:, None, None] ** 2 + kk[None, :, None] ** 2 + kk[None, None, :] ** 2)
    want = torch.bincount(torch.ceil(kmag).long().flatten(), minlength=N // 2 + 1)[1:N // 2 + 1]
    bad = 0
    for it in range(30):
        k, p, n = utils.power(x)
        d = (n.cpu().long() - want)
        if d.abs().sum() != 0:
            bad += 1
            if bad <= 2: print(N, it, "diff at", d.nonzero().flatten().tolist(), d[d != 0].tolist())
    print(N, "bad runs", bad, "of 30")
    xb = torch.randn((4, 1, N, N, N), device="cuda")
    ref = None
    for it in range(10):
        k, p, n = utils.power(xb)
        if ref is None: ref = p.clone()
        elif not torch.equal(ref, p): print(N, "P differs run to run: max rel", float(((p - ref).abs() / ref.abs()).max()))
