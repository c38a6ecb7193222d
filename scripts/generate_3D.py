#!/usr/bin/env python
"""Ensemble sampling with a trained (or random-init) 3-D conditional VDM: the reference's ``generate_3D.py``
(same positional CLI: model_name save_path runtype; its loop is generate_3D.py:43-97) on the B200-native path.

Differences, all in how the work is scheduled, not in what is computed:
  * realisations are independent units: realisation r of a test field goes to rank r mod world_size
    (``torchrun --nproc-per-node N scripts/generate_3D.py ...``), several per GPU at a time (``--batch``), and its
    noise stream is keyed by (seed, r) -- the saved ensemble does not depend on N or on the batch size;
  * the reverse loop replays one captured CUDA graph per step (vdm4cdm_b200.vdm_model.SamplerSession);
  * ``--synthetic`` stands in for the CAMELS CV test set (see scripts/_common.py).
Output: ``save_path/gen_{count}.npy`` with shape (rep, 1, N, N, N), as the reference writes it.
"""
import argparse
import os

import numpy as np
import yaml

from _common import ROOT, init_distributed, synthetic_batch

import torch
import torch.distributed as dist

from vdm4cdm_b200 import utils
from vdm4cdm_b200.trainer import shard_indices


def main():
    ap = argparse.ArgumentParser(description="Generate 3D CDM")
    ap.add_argument("model_name", type=str, help="Model name (an entry of configs.yaml)")
    ap.add_argument("save_path", type=str, help="Save path")
    ap.add_argument("runtype", type=str, help="CV_12_12 (12 fields x 12 realisations) or CV_1_128 (1 field x 128)")
    ap.add_argument("--synthetic", action="store_true", help="synthetic conditioning fields instead of the CAMELS CV set")
    ap.add_argument("--n-sampling-steps", type=int, default=250, help="reference default (model_test.ipynb:667)")
    ap.add_argument("--batch", type=int, default=2, help="realisations sampled together on one GPU")
    ap.add_argument("--rep", type=int, default=None, help="override the number of realisations per field")
    ap.add_argument("--fields", type=int, default=None, help="override the number of test fields")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--configs", default=os.path.join(ROOT, "configs.yaml"))
    args = ap.parse_args()
    if "SFM" in args.model_name:
        raise NotImplementedError("This model is not implemented yet")       # generate_3D.py:16-17
    assert args.runtype in ["CV_12_12", "CV_1_128"]
    configs = yaml.safe_load(open(args.configs))
    assert args.model_name in configs, f"{args.model_name} is not in {args.configs}"
    config = configs[args.model_name]
    rank, world, device = init_distributed()
    if rank == 0:
        os.makedirs(args.save_path, exist_ok=True)
    model = utils.get_model(config, device=device).eval()
    grid = int(config.get("cropsize", 128))
    n_params = int(config.get("conditioning_values", 6))
    n_fields, rep = (12, 12) if args.runtype == "CV_12_12" else (1, 128)
    n_fields = args.fields or n_fields
    rep = args.rep or rep
    if not args.synthetic:
        raise NotImplementedError("the CAMELS AstroDataModule is not part of this package yet (SURVEY.md section 8f); "
                                  "run with --synthetic")
    for count in range(n_fields):
        field = synthetic_batch(1, grid, args.seed + 1000 + count, n_params)     # the same field on every rank
        mine = list(shard_indices(rep, rank, world))
        gens = torch.zeros((rep, 1, grid, grid, grid), dtype=torch.float32, device=device)
        for i0 in range(0, len(mine), args.batch):
            ids = mine[i0:i0 + args.batch]
            b = len(ids)
            cond = field["conditioning"].to(device).expand(b, -1, -1, -1, -1).contiguous()
            vals = [v.to(device).expand(b, -1).contiguous() for v in field["conditioning_values"]]
            gen = model.draw_samples(batch_size=b, n_sampling_steps=args.n_sampling_steps, s_conditioning=cond,
                                     v_conditionings=vals, verbose=(rank == 0), seed=args.seed + count,
                                     realisation_ids=ids)
            gens[ids] = gen
        if world > 1:
            dist.all_reduce(gens)                       # every realisation was written by exactly one rank
        if rank == 0:
            np.save(os.path.join(args.save_path, f"gen_{count}.npy"), gens.cpu().numpy())
            print(f"field {count}: saved {rep} realisations")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
