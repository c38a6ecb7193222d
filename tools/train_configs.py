#!/usr/bin/env python
"""Steady-state training step time and peak memory for the other configurations BASELINE.json lists (SFM 160^3 batch 4,
VDM 224^3 batch 2, VDM 64^3 batch 2) and one sampling step of the 256^3 circular-padding model.  Timing: CUDA events
around K graph-replayed steps after warm-up + capture; inputs are synthetic Gaussian fields (scripts/_common.py)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import torch

from _common import synthetic_batch
from vdm4cdm_b200.networks import CUNet
from vdm4cdm_b200.sfm_model import LightSFM
from vdm4cdm_b200.trainer import Trainer
from vdm4cdm_b200.vdm_model import LightVDM


def net_for(n, chs):
    return CUNet(shape=(1, n, n, n), chs=chs, s_conditioning_channels=1, v_conditioning_dims=[6], t_conditioning=True,
                 norm_groups=8, mid_attn=False, dropout_prob=0.1, conv_padding_mode="circular" if n == 256 else "zeros",
                 n_attention_heads=4)


def train_case(kind, n, chs, batch, steps):
    torch.manual_seed(42)
    torch.cuda.reset_peak_memory_stats()
    net = net_for(n, chs)
    model = (LightVDM(score_model=net, gamma_max=13.3) if kind == "VDM" else LightSFM(velocity_model=net)).cuda()
    trainer = Trainer(model, gradient_clip_val=0.5)
    raw = synthetic_batch(batch, n, 42, device="cuda")
    b = raw if kind == "VDM" else {"x0": raw["conditioning"], "x1": raw["x"], "conditioning_values": raw["conditioning_values"]}
    losses = [trainer.training_step(b).item() for _ in range(5)]            # 3 eager + capture + replay
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(steps):
        loss = trainer.training_step(b)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    print(f"[train] {kind} {n}^3 chs={chs} batch {batch}: {ms:.2f} ms/step, {batch / ms * 1e3:.1f} samples/s, "
          f"{batch * n ** 3 / ms * 1e3 / 1e6:.1f} Mvoxel/s, peak memory {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB, "
          f"cuda_graph={trainer._graph is not None}, loss {losses[0]:.3f} -> {loss.item():.3f}", flush=True)
    del trainer, model, net
    torch.cuda.empty_cache()


def sample_case(n, chs, batch, steps):
    torch.manual_seed(42)
    torch.cuda.reset_peak_memory_stats()
    model = LightVDM(score_model=net_for(n, chs), gamma_max=13.3).cuda().eval()
    raw = synthetic_batch(batch, n, 42, device="cuda")
    kw = dict(s_conditioning=raw["conditioning"], v_conditionings=raw["conditioning_values"], seed=1)
    model.draw_samples(batch_size=batch, n_sampling_steps=4, **kw)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    out = model.draw_samples(batch_size=batch, n_sampling_steps=steps, **kw)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / steps
    print(f"[sample] VDM {n}^3 chs={chs} ({'circular' if n == 256 else 'zeros'} padding) {batch} realisations: {ms:.2f} ms/step, "
          f"{batch * n ** 3 / ms * 1e3:.3e} voxel-steps/s, peak memory {torch.cuda.max_memory_allocated() / 2 ** 30:.1f} GiB, "
          f"finite={bool(torch.isfinite(out).all())}", flush=True)
    del model
    torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    train_case("VDM", 64, [16, 32, 64, 128], 2, args.steps)
    train_case("VDM", 128, [32, 64, 128, 256], 2, args.steps)
    train_case("SFM", 160, [32, 64, 128, 256], 4, args.steps)
    train_case("VDM", 224, [16, 32, 64, 128], 2, args.steps)
    train_case("VDM", 256, [16, 32, 64, 128], 1, args.steps)
    sample_case(256, [16, 32, 64, 128], 2, args.steps)
    sample_case(224, [16, 32, 64, 128], 2, args.steps)
