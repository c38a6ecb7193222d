#!/usr/bin/env python
"""Launch a few vdm_conv3d / wgrad cases once each (for an `ncu --set full` capture).
usage: python tools/ncu_conv_case.py [--cin 32 --cout 32 --grid 128 --batch 2]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vdm4cdm_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=32)
ap.add_argument("--cout", type=int, default=32)
ap.add_argument("--grid", type=int, default=128)
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--taps", type=int, default=27, choices=[1, 27])
ap.add_argument("--in-norm", action="store_true", help="also launch the conv with the fused GroupNorm + SiLU input transform")
args = ap.parse_args()
dev = torch.device("cuda:0")
b, ci, co, n = args.batch, args.cin, args.cout, args.grid
x = torch.randn((b, ci // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
res = torch.randn((b, co // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
k = 3 if args.taps == 27 else 1
taps = ops.TAPS_3X3X3 if k == 3 else ops.TAPS_1X1X1
w = ops.pack_conv_weight(torch.randn((co, ci, k, k, k), device=dev) / (args.taps * ci) ** 0.5)
res_c = torch.randn((b, co // 8, n // 2, n // 2, n // 2, 8), device=dev).to(torch.bfloat16)
out = torch.empty((b, co // 8, n, n, n, 8), dtype=torch.bfloat16, device=dev)
cadd = torch.randn((b, co), device=dev)
stats = torch.zeros((b, co, 2), dtype=torch.float64, device=dev)
for _ in range(2):
    ops.conv3d(x, w, co, out=out, taps=taps)                                                   # bare
    ops.conv3d(x, w, co, out=out, taps=taps, chan_add=cadd, residual=res, stats=stats)          # full epilogue
    ops.conv3d(x, w, co, out=out, taps=taps, chan_add=cadd, residual=res_c, residual_upsample=True)
    ops.conv3d_wgrad(x, out, ci, co, k)
    if args.in_norm:
        coef = torch.randn((b, ci, 2), device=dev) * 0.5
        ops.conv3d(x, w, co, out=out, taps=taps, chan_add=cadd, stats=stats, in_norm=coef)
torch.cuda.synchronize()
print("done")
