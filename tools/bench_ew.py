#!/usr/bin/env python
"""Achieved HBM bandwidth of the elementwise kernels at the level-0 shapes of the 128^3 network.
usage: python tools/bench_ew.py [--grid 128 --batch 2]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vdm4cdm_b200 import ops  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    b, n = args.batch, args.grid
    for c in (32, 96):
        x = torch.randn((b, c // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
        dy = torch.randn((b, c // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
        add = torch.randn((b, c // 8, n, n, n, 8), device=dev).to(torch.bfloat16)
        out = torch.empty_like(x)
        gamma, beta = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.1
        st = ops.channel_stats(x, c)
        tensor_bytes = x.numel() * 2
        sums = torch.zeros((b, c, 2), dtype=torch.float64, device=dev)
        ostats = torch.zeros((b, c, 2), dtype=torch.float64, device=dev)
        lib = ops._C.lib()
        import ctypes
        cases = [
            ("channel_stats        (1 pass )", 1, lambda: ops.channel_stats(x, c, stats=ostats)),
            ("gn_silu              (2 passes)", 2, lambda: ops.gn_silu(x, c, 8, st, gamma, beta, out=out)),
            ("gn_silu dropout 0.1  (2 passes)", 2, lambda: ops.gn_silu(x, c, 8, st, gamma, beta, out=out, dropout_p=0.1, seed=1, layer_tag=1)),
            ("gn_silu_bwd r+a      (5 passes)", 5, lambda: ops.gn_silu_bwd(x, dy, c, 8, st, gamma, beta, out=out, sums=sums, out_stats=ostats)),
            ("gn_silu_bwd r+a +add (6 passes)", 6, lambda: ops.gn_silu_bwd(x, dy, c, 8, st, gamma, beta, out=out, add=add, sums=sums, out_stats=ostats)),
        ]
        for name, passes, fn in cases:
            ms = timeit(fn)
            print(f"C={c:3d} {name}: {ms:7.3f} ms  {passes * tensor_bytes / ms / 1e6:7.0f} GB/s")
    c = 64
    coarse = torch.randn((b, c // 8, n // 2, n // 2, n // 2, 8), device=dev).to(torch.bfloat16)
    fine = torch.empty((b, 12, n, n, n, 8), dtype=torch.bfloat16, device=dev)
    st = torch.zeros((b, 96, 2), dtype=torch.float64, device=dev)
    ms = timeit(lambda: ops.upsample2(coarse, c, fine, stats=st))
    print(f"upsample2 64ch -> 128^3: {ms:7.3f} ms  {(coarse.numel() * 2 + coarse.numel() * 16) / ms / 1e6:7.0f} GB/s")
    x32 = torch.randn((b, 4, n, n, n, 8), device=dev).to(torch.bfloat16)
    ms = timeit(lambda: ops.avgpool2(x32, 32, stats=st))
    print(f"avgpool2 32ch 128^3: {ms:7.3f} ms  {(x32.numel() * 2 * 1.125) / ms / 1e6:7.0f} GB/s")
    t = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
    t2 = torch.empty_like(t)
    ms = timeit(lambda: t2.copy_(t))
    print(f"torch copy 2 GiB: {ms:7.3f} ms  {t.numel() * 4 / ms / 1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
