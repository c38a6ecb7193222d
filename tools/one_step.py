#!/usr/bin/env python
"""One eager 128^3 x 8 sampler step inside the NVTX range "profiled_step" (after warm-up steps), for ncu launch lists:
   ncu --nvtx --nvtx-include "profiled_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python tools/one_step.py"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from bench import model_kwargs, synthetic_batch
from vdm4cdm_b200.networks import CUNet
from vdm4cdm_b200.vdm_model import LightVDM

dev = torch.device("cuda:0")
torch.manual_seed(42)
batch, grid = int(os.environ.get("STEP_BATCH", "8")), 128
model = LightVDM(score_model=CUNet(**model_kwargs(grid, [32, 64, 128, 256])), gamma_max=13.3).to(dev).eval()
x, cond, params = synthetic_batch(batch, grid, 42)
vdm = model.model
vdm.use_cuda_graph = False
sess = vdm.session(batch, 1000, dev, seed=42, s_conditioning=cond.to(dev), v_conditionings=[params.to(dev)])
for _ in range(3):
    sess.step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("profiled_step")
sess.step()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done")
