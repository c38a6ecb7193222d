#!/usr/bin/env python
"""Sweep of the conv tile shape (MT, KC, n_split; bring-up library) for the wide layers of the 128^3 sampling step:
what the cost model of vdm_conv3d picks against the best forced configuration.  usage: python tools/bench_tiles.py"""
import os
os.environ["VDM4CDM_BRINGUP"] = "1"
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
from vdm4cdm_b200 import _C, ops
from bench_conv import timeit
dev = torch.device("cuda:0")
lib = _C.lib()
cases = [(256, 256, 16, 8, 27, False), (128, 128, 32, 8, 27, False), (256, 256, 16, 8, 27, True), (128, 128, 32, 8, 27, True),
         (64, 128, 32, 8, 27, False), (128, 256, 16, 8, 27, False), (64, 64, 64, 8, 1, False), (128, 128, 32, 8, 1, False)]
for ci, co, n, b, taps, res in cases:
    k = 3 if taps == 27 else 1
    x = torch.randn((b, ci // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
    r = torch.randn((b, co // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
    w = ops.pack_conv_weight(torch.randn((co, ci, k, k, k), device=dev) / (taps * ci) ** 0.5)
    out = torch.empty((b, co // 8, n, n, n, 8), dtype=torch.bfloat16, device=dev)
    cadd = torch.randn((b, co), device=dev)
    stats = torch.zeros((b, co, 2), dtype=torch.float64, device=dev)
    fn = lambda: ops.conv3d(x, w, co, taps=ops.TAPS_3X3X3 if k == 3 else ops.TAPS_1X1X1, out=out, chan_add=cadd, stats=stats,
                            residual=r if res else None)
    flops = 2.0 * taps * ci * co * b * n ** 3
    t0 = timeit(fn)
    print(f"conv {ci}->{co} taps={taps} at {n}^3 x {b} residual={res}: automatic {t0:.4f} ms ({flops / t0 / 1e9:.0f} TFLOP/s)")
    best = []
    for mt in (1, 2, 3, 4):
        for kc in (16, 32, 64):
            for ns in (1, 2, 4, 8):
                lib.vdm_debug_set(1, mt); lib.vdm_debug_set(2, kc); lib.vdm_debug_set(3, ns)
                try:
                    best.append((timeit(fn, iters=5), mt, kc, ns))
                except RuntimeError:
                    pass
    lib.vdm_debug_set(1, 0); lib.vdm_debug_set(2, 0); lib.vdm_debug_set(3, 0)
    for t, mt, kc, ns in sorted(best)[:5]:
        print(f"   MT={mt} KC={kc} n_split={ns}: {t:.4f} ms ({flops / t / 1e9:.0f} TFLOP/s)")
