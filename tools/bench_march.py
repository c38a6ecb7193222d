#!/usr/bin/env python
"""A/B timing of the d-marching schedule (conv3d_march.cuh) against the tile schedule on the level-0 layer classes.
usage: python tools/bench_march.py [--grid 128 --batch 8]   (bring-up build: make -C vdm4cdm_b200/csrc bringup)"""
import argparse
import os
os.environ["VDM4CDM_BRINGUP"] = "1"
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vdm4cdm_b200 import _C, ops  # noqa: E402
from tools.bench_conv import timeit  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--batch", type=int, default=8)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    b, n = args.batch, args.grid
    lib = _C.lib()
    g = torch.Generator(device="cuda").manual_seed(0)
    for ci, co, fp32 in ((32, 32, False), (16, 32, False), (32, 1, True), (32, 16, False), (16, 16, False)):
        x = torch.randn((b, ci // 8, n, n, n, 8), device=dev, generator=g).to(torch.bfloat16)
        cop = max(co, 8)
        res = torch.randn((b, (cop + 7) // 8, n, n, n, 8), device=dev, generator=g).to(torch.bfloat16)
        w = ops.pack_conv_weight(torch.randn((co, ci, 3, 3, 3), device=dev) / (27 * ci) ** 0.5)
        cadd = torch.randn((b, co), device=dev)
        stats = torch.zeros((b, co, 2), dtype=torch.float64, device=dev)
        flops = 2.0 * 27 * ci * co * b * n ** 3
        variants = [("bias", {}), ("bias+stats", {"stats": stats}), ("bias+res+stats", {"stats": stats, "residual": res})]
        if fp32:
            variants = [("bias, fp32 out", {"out_fp32": True})]
        for name, kw in variants:
            outs = []
            line = f"conv {ci}->{co} {n}^3 x {b} {name:16s}"
            for no_march in (1, 0):
                lib.vdm_debug_set(7, no_march)
                y = ops.conv3d(x, w, co, taps=ops.TAPS_3X3X3, chan_add=cadd, **kw)
                outs.append(y.float().clone())
                ms = timeit(lambda: ops.conv3d(x, w, co, taps=ops.TAPS_3X3X3, chan_add=cadd, **kw))
                line += f"  {'tile ' if no_march else 'march'} {ms:6.3f} ms {flops / ms / 1e9:6.0f} TF/s"
            lib.vdm_debug_set(7, 0)
            line += f"  max|diff| {float((outs[0] - outs[1]).abs().max()):.3g}"
            print(line, flush=True)
        if (ci, co) in ((32, 32), (16, 32)):
            for flags, what in ((1, "epilogue only drains TMEM"), (2, "no halo loads after the fill"), (3, "neither")):
                lib.vdm_debug_set(5, flags)
                ms = timeit(lambda: ops.conv3d(x, w, co, taps=ops.TAPS_3X3X3, chan_add=cadd, stats=stats))
                print(f"    EXPERIMENT march, {what}: {ms:6.3f} ms {flops / ms / 1e9:6.0f} TF/s", flush=True)
            lib.vdm_debug_set(5, 0)


if __name__ == "__main__":
    main()
