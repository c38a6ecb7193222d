#!/usr/bin/env python
"""Which part of the conv epilogue costs what: the full-epilogue 32->32 conv with pieces switched off
(vdm_debug_set key 5; results are wrong by construction).  usage: python tools/bench_epilogue.py [--taps 27|1]"""
import argparse
import os
os.environ["VDM4CDM_BRINGUP"] = "1"   # bring-up build of the library (make -C vdm4cdm_b200/csrc bringup)
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vdm4cdm_b200 import _C, ops  # noqa: E402
from bench_conv import timeit  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cin", type=int, default=32)
ap.add_argument("--cout", type=int, default=32)
ap.add_argument("--grid", type=int, default=128)
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--taps", type=int, default=27)
a = ap.parse_args()
dev = torch.device("cuda:0")
b, ci, co, n = a.batch, a.cin, a.cout, a.grid
k = 3 if a.taps == 27 else 1
taps = ops.TAPS_3X3X3 if k == 3 else ops.TAPS_1X1X1
x = torch.randn((b, ci // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
res = torch.randn((b, co // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
w = ops.pack_conv_weight(torch.randn((co, ci, k, k, k), device=dev) / (a.taps * ci) ** 0.5)
out = torch.empty((b, co // 8, n, n, n, 8), dtype=torch.bfloat16, device=dev)
cadd = torch.randn((b, co), device=dev)
stats = torch.zeros((b, co, 2), dtype=torch.float64, device=dev)
lib = _C.lib()
full = lambda: ops.conv3d(x, w, co, taps=taps, out=out, chan_add=cadd, residual=res, stats=stats)
nores = lambda: ops.conv3d(x, w, co, taps=taps, out=out, chan_add=cadd, stats=stats)
print(f"conv {ci}->{co} taps={a.taps} grid={n}^3 B={b}")
for flags, what in ((0, "nothing off"), (4, "no TMEM reads"), (8, "no output stores"), (16, "no stats transpose-reduction"),
                    (32, "no stats barrier+fold"), (48, "no stats reduction, barrier, fold"), (12, "no TMEM reads, no stores"),
                    (60, "all four off"), (1, "epilogue does nothing")):
    lib.vdm_debug_set(5, flags)
    print(f"  {what:40s}: bias+residual+stats {timeit(full):6.3f} ms   bias+stats {timeit(nores):6.3f} ms")
lib.vdm_debug_set(5, 0)
# fused input transform (GroupNorm + SiLU applied to the landed halo tile by warps 12..15): cost of the arithmetic vs the hand-off
coef = torch.randn((b, ci, 2), device=dev) * 0.5
xf = lambda: ops.conv3d(x, w, co, taps=taps, out=out, chan_add=cadd, stats=stats, in_norm=coef)
xf_res = lambda: ops.conv3d(x, w, co, taps=taps, out=out, chan_add=cadd, residual=res, stats=stats, in_norm=coef)
print(f"  fused input transform                   : bias+residual+stats {timeit(xf_res):6.3f} ms   bias+stats {timeit(xf):6.3f} ms")
for flags, what in ((64, "transform warps hand the stage on only"), (128, "transform copies through (LDS + STS, no math)"),
                    (256, "transform computes, no stores"), (384, "transform loads only")):
    lib.vdm_debug_set(5, flags)
    print(f"  {what:40s}: bias+residual+stats {timeit(xf_res):6.3f} ms   bias+stats {timeit(xf):6.3f} ms")
lib.vdm_debug_set(5, 0)
