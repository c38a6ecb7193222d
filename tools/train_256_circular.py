#!/usr/bin/env python
"""Does the 256^3 circular-padding model learn?  (loss 13.87 -> 13.97 over 10 steps.)  Loss curves of
50 training steps at batch 1 for 256^3 circular, 256^3-sized zero padding is impossible (registry uses circular there), so the
controls are 128^3 circular vs 128^3 zeros with the SAME chs / batch / seeds, and 256^3 circular with a fixed batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
from _common import synthetic_batch
from vdm4cdm_b200.networks import CUNet
from vdm4cdm_b200.trainer import Trainer
from vdm4cdm_b200.vdm_model import LightVDM

def run(n, pad, batch, steps, fixed_batch):
    torch.manual_seed(42)
    net = CUNet(shape=(1, n, n, n), chs=[16, 32, 64, 128], s_conditioning_channels=1, v_conditioning_dims=[6], t_conditioning=True,
                norm_groups=8, dropout_prob=0.1, conv_padding_mode=pad)
    model = LightVDM(score_model=net, gamma_max=13.3).cuda()
    tr = Trainer(model, gradient_clip_val=0.5)
    losses = []
    for i in range(steps):
        b = synthetic_batch(batch, n, 42 if fixed_batch else 42 + i, device="cuda")
        losses.append(tr.training_step(b).item())
    k = max(1, steps // 10)
    avg = [sum(losses[i:i + k]) / k for i in range(0, steps, k)]
    print(f"{n}^3 {pad:8s} batch {batch} fixed_batch={fixed_batch}: mean loss per {k} steps: " + " ".join(f"{a:.3f}" for a in avg), flush=True)
    del tr, model, net
    torch.cuda.empty_cache()

run(128, "zeros", 1, 50, False)
run(128, "circular", 1, 50, False)
run(256, "circular", 1, 50, False)
run(256, "circular", 1, 50, True)
run(128, "circular", 2, 50, False)
