"""march vs tile schedule, output and statistics, on the epilogue modes the network uses (bring-up build)."""
import os, sys
os.environ["VDM4CDM_BRINGUP"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vdm4cdm_b200 import _C, ops
lib = _C.lib()
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
def rnd(*shape):
    return torch.randn(shape, device=dev, generator=g).to(torch.bfloat16)
for (b, ci, co, d, h, w) in [(2, 16, 16, 16, 32, 16), (2, 32, 32, 8, 16, 8), (2, 32, 16, 16, 32, 16), (1, 32, 32, 24, 40, 24), (3, 16, 32, 6, 20, 12)]:
    wt = ops.pack_conv_weight(torch.randn((co, ci, 3, 3, 3), device=dev) / (27 * ci) ** 0.5)
    cadd = torch.randn((b, co), device=dev)
    xw = rnd(b, ci // 8 + 3, d, h, w, 8)              # window: planes 2.. of a wider buffer
    modes = {
        "plain residual": dict(residual=rnd(b, co // 8, d, h, w, 8)),
        "residual window": dict(residual=rnd(b, co // 8 + 2, d, h, w, 8), residual_plane0=1),
        "coarse residual": dict(residual=rnd(b, co // 8, d // 2, h // 2, w // 2, 8), residual_upsample=True),
        "d2s residual": dict(residual=rnd(b, co, d // 2, h // 2, w // 2, 8), residual_upsample="d2s"),
        "no residual": dict(),
    }
    for name, kw in modes.items():
        for out_window in (False, True):
            res = []
            for no_march in (1, 0):
                lib.vdm_debug_set(7, no_march)
                stats = torch.zeros((b, co + 16, 2), dtype=torch.float64, device=dev)
                out = torch.zeros((b, co // 8 + 2, d, h, w, 8), dtype=torch.bfloat16, device=dev) if out_window else None
                y = ops.conv3d(xw, wt, co, x_plane0=2, c_in=ci, chan_add=cadd, stats=stats, stats_c0=8, out=out,
                               out_plane0=1 if out_window else 0, **kw)
                torch.cuda.synchronize()
                res.append((y.float().clone(), stats.clone()))
            lib.vdm_debug_set(7, 0)
            dy = float((res[0][0] - res[1][0]).abs().max())
            ds = float(((res[0][1] - res[1][1]).abs() / res[0][1].abs().clamp_min(1e-20)).max())
            flag = "" if dy == 0 and ds < 1e-6 else "   <<<<<<<<"
            print(f"B={b} {ci}->{co} {d}x{h}x{w} {name:16s} out_window={out_window}: y max|diff| {dy:.3g}  stats rel diff {ds:.3g}{flag}", flush=True)
    # fp32 single-channel output
    wo = ops.pack_conv_weight(torch.randn((1, ci, 3, 3, 3), device=dev) / (27 * ci) ** 0.5)
    res = []
    for no_march in (1, 0):
        lib.vdm_debug_set(7, no_march)
        res.append(ops.conv3d(xw, wo, 1, x_plane0=2, c_in=ci, chan_add=cadd[:, :1].contiguous(), out_fp32=True).clone())
    lib.vdm_debug_set(7, 0)
    print(f"B={b} {ci}->1 fp32 out: max|diff| {float((res[0] - res[1]).abs().max()):.3g}", flush=True)
