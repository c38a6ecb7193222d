"""bf16 noise floor of the CUNet against the fp32 oracle: relative L2 error for zero / circular padding over a few seeds."""
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_unet import _models, _rel_l2
for shape, chs in [((1, 16, 32, 16), (16, 32, 64)), ((1, 24, 24, 24), (16, 32, 64))]:
    for pad in ("zeros", "circular"):
        errs = []
        for seed in range(4):
            torch.manual_seed(seed)
            ref, net = _models(shape, chs, seed=seed, padding=pad)
            g = torch.Generator().manual_seed(seed)
            x = torch.randn((2,) + shape, generator=g)
            cond = 0.7 * x + 0.3 * torch.randn((2,) + shape, generator=g)
            t, v = torch.rand(2, generator=g), [torch.rand(2, 6, generator=g)]
            with torch.no_grad():
                want = ref(x, t=t, s_conditioning=cond, v_conditionings=v)
                got = net(x.cuda(), t=t.cuda(), s_conditioning=cond.cuda(), v_conditionings=[v[0].cuda()]).cpu()
            errs.append(_rel_l2(got, want))
        print(shape, pad, " ".join(f"{e:.2e}" for e in errs))
