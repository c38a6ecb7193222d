#!/usr/bin/env python
"""Micro-benchmark of vdm_conv3d on one layer shape with epilogue features toggled.
usage: python tools/bench_conv.py [--cin 32 --cout 32 --grid 128 --batch 2]"""
import argparse
import itertools
import os
os.environ["VDM4CDM_BRINGUP"] = "1"   # bring-up build of the library (make -C vdm4cdm_b200/csrc bringup)
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vdm4cdm_b200 import _C, ops  # noqa: E402


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
    ap.add_argument("--cin", type=int, default=32)
    ap.add_argument("--cout", type=int, default=32)
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--taps", type=int, default=27)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    b, ci, co, n = args.batch, args.cin, args.cout, args.grid
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((b, ci // 8, n, n, n, 8), device=dev, generator=g).to(torch.bfloat16)
    small = (x.float() * 1e-5).to(torch.bfloat16)
    res = torch.randn((b, co // 8, n, n, n, 8), device=dev, generator=g).to(torch.bfloat16)
    k = 3 if args.taps == 27 else 1
    w = ops.pack_conv_weight(torch.randn((co, ci, k, k, k), device=dev) / (args.taps * ci) ** 0.5)
    taps = ops.TAPS_3X3X3 if k == 3 else ops.TAPS_1X1X1
    out = torch.empty((b, co // 8, n, n, n, 8), dtype=torch.bfloat16, device=dev)
    cadd = torch.randn((b, co), device=dev)
    stats = torch.zeros((b, co, 2), dtype=torch.float64, device=dev)
    flops = 2.0 * args.taps * ci * co * b * n ** 3
    lib = _C.lib()
    print(f"conv {ci}->{co} taps={args.taps} grid={n}^3 B={b}")
    for use_stats, use_cadd, use_res in itertools.product((0, 1), (0, 1), (0, 1)):
        ms = timeit(lambda: ops.conv3d(x, w, co, taps=taps, out=out, chan_add=cadd if use_cadd else None,
                                       residual=res if use_res else None, stats=stats if use_stats else None))
        print(f"  stats={use_stats} chan_add={use_cadd} residual={use_res}: {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")
    ms = timeit(lambda: ops.conv3d(small, w, co, taps=taps, out=out))
    print(f"  tiny-valued input (1e-5 scale), no epilogue extras: {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")
    zero = torch.zeros_like(x)
    ms = timeit(lambda: ops.conv3d(zero, w, co, taps=taps, out=out))
    print(f"  all-zero input, no epilogue extras: {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")
    for flags, what in ((1, "epilogue does nothing"), (2, "no halo loads after the fill"), (3, "neither")):
        lib.vdm_debug_set(5, flags)
        ms = timeit(lambda: ops.conv3d(x, w, co, taps=taps, out=out))
        print(f"  EXPERIMENT (automatic schedule) {what}: {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")
    lib.vdm_debug_set(5, 0)
    for mt in (1, 2, 3, 4):
        lib.vdm_debug_set(1, mt)
        try:
            ms = timeit(lambda: ops.conv3d(x, w, co, taps=taps, out=out))
            print(f"  forced MT={mt}, no epilogue extras: {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")
        except RuntimeError as e:
            print(f"  forced MT={mt}: {e}")
    lib.vdm_debug_set(1, 0)
    for flags, what in ((1, "epilogue does nothing"), (2, "no halo loads after the fill"), (3, "neither")):
        lib.vdm_debug_set(5, flags)
        for mt in (1, 4):
            lib.vdm_debug_set(1, mt)
            ms = timeit(lambda: ops.conv3d(x, w, co, taps=taps, out=out))
            print(f"  EXPERIMENT {what}, MT={mt}: {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")
    lib.vdm_debug_set(5, 0)
    lib.vdm_debug_set(1, 0)
    lib.vdm_debug_set(4, 1)
    ms = timeit(lambda: ops.conv3d(x, w, co, taps=taps, out=out))
    print(f"  weights streamed per tile (resident mode off), no epilogue extras: {ms:7.3f} ms  {flops / ms / 1e9:7.1f} TFLOP/s")
    lib.vdm_debug_set(4, 0)


if __name__ == "__main__":
    main()
