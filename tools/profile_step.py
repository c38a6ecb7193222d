#!/usr/bin/env python
"""Kernel-time breakdown of one training step and one sampler step at 128^3 (torch.profiler / CUPTI).
usage: python tools/profile_step.py [--grid 128] [--batch 2] > gpurun_out/<tag>_kernels.txt"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from bench import model_kwargs, synthetic_batch  # noqa: E402
from vdm4cdm_b200.networks import CUNet  # noqa: E402
from vdm4cdm_b200.trainer import Trainer  # noqa: E402
from vdm4cdm_b200.vdm_model import LightVDM  # noqa: E402


def table(prof, title, n_steps):
    rows = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(e.name[:90], [0.0, 0])
            r[0] += e.device_time / 1e3
            r[1] += 1
    tot = sum(r[0] for r in rows.values())
    print(f"## {title}: {tot / n_steps:.3f} ms of kernels per step ({n_steps} steps profiled)")
    for k, (ms, n) in sorted(rows.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{ms / n_steps:9.3f} ms {n // n_steps:5d}x  {k}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--chs", type=int, nargs="+", default=[32, 64, 128, 256])
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    net = CUNet(**model_kwargs(args.grid, args.chs))
    model = LightVDM(score_model=net).to(dev)
    x, cond, params = synthetic_batch(args.batch, args.grid, 42)
    batch = {"x": x.to(dev), "conditioning": cond.to(dev), "conditioning_values": [params.to(dev)]}
    # sampler
    model.eval()
    sess = model.model.session(args.batch, 1000, dev, seed=1, s_conditioning=batch["conditioning"],
                               v_conditionings=batch["conditioning_values"])
    for _ in range(3):
        sess.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            sess.step()
        torch.cuda.synchronize()
    table(prof, "sampler step (graph replay)", 3)
    model.model.__dict__.get("_sessions", {}).clear()
    # training
    tr = Trainer(model.train())
    for _ in range(3):
        tr.training_step(batch)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        tr.training_step(batch)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"## host: {(t1 - t0) / 3 * 1e3:.2f} ms/step to issue, {(t2 - t0) / 3 * 1e3:.2f} ms/step until the GPU is done")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            tr.training_step(batch)
        torch.cuda.synchronize()
    table(prof, "training step", 2)


if __name__ == "__main__":
    main()
