#!/usr/bin/env python
"""Timing of the polyphase parity convs (8 taps on the coarse grid) of the three up blocks, with the tile shape forced
through the bring-up library (vdm_debug_set keys 1 = MT, 2 = KC, 3 = n_split).  usage: python tools/bench_poly.py"""
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
for cc, co, n, b in ((64, 32, 64, 8), (128, 64, 32, 8), (256, 128, 16, 8)):
    x = torch.randn((b, cc // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
    w = ops.pack_conv_weight(torch.randn((co, cc, 2, 2, 2), device=dev) / (8 * cc) ** 0.5)
    out = torch.empty((b, co, n, n, n, 8), dtype=torch.bfloat16, device=dev)
    taps = ops.polyphase_taps((0, 1, 0))
    fn = lambda: ops.conv3d(x, w, co, taps=taps, out=out, out_plane0=3 * (co // 8))
    flops = 2.0 * 8 * cc * co * b * n ** 3
    print(f"parity conv {cc}->{co} at {n}^3 x {b}: automatic {timeit(fn):.4f} ms ({flops / timeit(fn) / 1e9:.0f} TFLOP/s)")
    for mt in (1, 2, 3, 4):
        for kc in (16, 32, 64):
            for ns in (1, 2, 4):
                lib.vdm_debug_set(1, mt); lib.vdm_debug_set(2, kc); lib.vdm_debug_set(3, ns)
                try:
                    t = timeit(fn)
                    print(f"   MT={mt} KC={kc} n_split={ns}: {t:.4f} ms")
                except RuntimeError as e:
                    pass
    lib.vdm_debug_set(1, 0); lib.vdm_debug_set(2, 0); lib.vdm_debug_set(3, 0)
