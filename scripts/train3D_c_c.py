#!/usr/bin/env python
"""Training entry point for the 3-D conditional models: the reference's
``trainVDM3D{,128,160,192,224}_c_c_from_field_name_thick_lowbatch.py`` and ``trainSFM3D*_c_c_..._lowbatch.py``
(positional CLI ``field_in field_out cropsize`` as there; hyper-parameters from those files: chs, batch_size,
dropout 0.1, norm_groups 8, gamma_max 13.3, lr 3e-4, gradient_clip_val 0.5, seed 42) with the Lightning /
Comet orchestration replaced by ``vdm4cdm_b200.trainer.Trainer``:

    torchrun --nproc-per-node 8 scripts/train3D_c_c.py Mstar Mcdm 128 --model VDM --synthetic --max-steps 1000

One process per GPU, each with its own micro-batch of ``batch_size`` samples, gradients all-reduced in one flat
bucket over NCCL/NVLink.  ``--synthetic`` stands in for the CAMELS LH set (see scripts/_common.py).
Checkpoints: ``{"state_dict": ...}`` files, loadable by ``vdm4cdm_b200.utils.get_model`` (src/utils.py:467).
"""
import argparse
import os
import time

from _common import init_distributed, synthetic_batch

import torch
import torch.distributed as dist

from mltools.models import sfm_model, vdm_model
from mltools.networks import networks
from vdm4cdm_b200.trainer import Trainer

# Hyper-parameters of the reference's scripts (tests/golden/train_presets.json, read from their syntax trees by
# oracle/make_golden_train_presets.py): the grid sizes with a dedicated script train on the re-gridded boxes CMD_{n};
# every other cropsize is the base script (trainVDM3D_c_c_... / trainSFM3D_c_c_...: crops of the 256^3 "CMD" boxes).
#   (model, cropsize) -> (chs, batch_size per GPU, dataset_name)
PRESETS = {("VDM", 128): ([32, 64, 128, 256], 2, "CMD_128"), ("VDM", 160): ([32, 64, 128, 256], 2, "CMD_160"),
           ("VDM", 192): ([32, 64, 128, 256], 2, "CMD_192"), ("VDM", 224): ([16, 32, 64, 128], 2, "CMD_224"),
           ("SFM", 128): ([32, 64, 128, 256], 4, "CMD_128"), ("SFM", 160): ([32, 64, 128, 256], 4, "CMD_160"),
           ("SFM", 192): ([32, 64, 128, 256], 2, "CMD_192")}
BASE_PRESET = ([16, 32, 64, 128], 2, "CMD")
NORM_GROUPS, DROPOUT_PROB, GAMMA_MAX, LEARNING_RATE, GRADIENT_CLIP_VAL = 8, 0.1, 13.3, 3.0e-4, 0.5


def preset(model: str, cropsize: int):
    return PRESETS.get((model, cropsize), BASE_PRESET)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("field_in")
    ap.add_argument("field_out")
    ap.add_argument("cropsize", type=int)
    ap.add_argument("--model", choices=["VDM", "SFM"], default="VDM")
    ap.add_argument("--synthetic", action="store_true")
    ap.add_argument("--data-root", default=os.environ.get("VDM4CDM_DATA_ROOT"),
                    help="directory with the CAMELS grids (Grids_{field}_{suite}_LH_{res}_z=0.0.npy) and params_LH_{suite}.txt")
    ap.add_argument("--suite-name", default="Astrid")
    ap.add_argument("--dataset-name", default=None, help="override the grid set (default: the reference script's, e.g. CMD_128)")
    ap.add_argument("--host-boxes", action="store_true", help="keep the boxes memory-mapped on the host (mmap=True)")
    ap.add_argument("--max-epochs", type=int, default=1000)
    ap.add_argument("--synthetic-boxes", action="store_true",
                    help="with --synthetic: synthetic RAW simulation boxes through the device-resident dataset pipeline")
    ap.add_argument("--max-steps", type=int, default=1_000_000)
    ap.add_argument("--batch-size", type=int, default=None, help="samples per GPU (default: the reference script's)")
    ap.add_argument("--chs", type=int, nargs="+", default=None, help="override the channel widths per level")
    ap.add_argument("--ckpt-dir", default="./checkpoints")
    ap.add_argument("--ckpt-every", type=int, default=10_000)
    ap.add_argument("--log-every", type=int, default=10)
    args = ap.parse_args()
    rank, world, device = init_distributed()
    torch.manual_seed(42)                                     # seed_everything(42)
    chs, batch_size, dataset_name = preset(args.model, args.cropsize)
    batch_size = args.batch_size or batch_size
    dataset_name = args.dataset_name or dataset_name
    chs = args.chs or chs
    n = args.cropsize
    net = networks.CUNet(shape=(1, n, n, n), chs=chs, s_conditioning_channels=1, v_conditioning_dims=[6],
                         t_conditioning=True, norm_groups=NORM_GROUPS, mid_attn=False, dropout_prob=DROPOUT_PROB,
                         conv_padding_mode="circular" if n == 256 else "zeros", n_attention_heads=4)
    if args.model == "VDM":
        model = vdm_model.LightVDM(score_model=net, draw_figure=None, gamma_max=GAMMA_MAX, learning_rate=LEARNING_RATE)
    else:
        model = sfm_model.LightSFM(velocity_model=net, draw_figure=None, learning_rate=LEARNING_RATE)
    model = model.to(device)
    trainer = Trainer(model, gradient_clip_val=GRADIENT_CLIP_VAL)
    stream = None
    if not args.synthetic:
        # trainVDM3D128_...:75-89: LH set of the re-gridded boxes, stage "fit", mmap=False (boxes resident in HBM)
        from vdm4cdm_b200.dataset import get_dataset
        key_c, key_x = ("conditioning", "x") if args.model == "VDM" else ("x0", "x1")
        dm = get_dataset(dataset_name=dataset_name, suite_name=args.suite_name,
                         return_func=lambda fields, params: {key_c: fields[0], key_x: fields[1], "conditioning_values": [params]},
                         set_name="LH", z_name="z_0.0", channel_names=[args.field_in, args.field_out], stage="fit",
                         batch_size=batch_size, cropsize=n, mmap=args.host_boxes, data_root=args.data_root, device=device,
                         seed=42, rank=rank, world=world)

        def epochs():
            for _ in range(args.max_epochs):
                yield from dm.train_dataloader()
        stream = epochs()
    if args.synthetic_boxes:
        # the full data path: raw boxes resident in HBM -> vdm_augment_crop (periodic crop, log-normalise, flip, permute)
        from vdm4cdm_b200.dataset import DeviceAstroDataset
        S, n_sims = max(2 * n, 256) if n < 256 else n, 4
        g = torch.Generator().manual_seed(7)
        base = torch.randn((n_sims, S, S, S), generator=g)
        raw_x = (10.0 ** (base * 0.552 + 10.019)).to(device)
        raw_c = (10.0 ** ((0.7 * base + 0.3 * torch.randn((n_sims, S, S, S), generator=g)) * 0.6 + 8.0)).to(device)
        params = torch.rand((n_sims, 6), generator=g)
        key_c, key_x = ("conditioning", "x") if args.model == "VDM" else ("x0", "x1")
        ds = DeviceAstroDataset([raw_c, raw_x], params, lambda fields, params: {key_c: fields[0], key_x: fields[1],
                                                                               "conditioning_values": [params]},
                                alphas=[1.0, 1.0], means=[8.0, 10.019], stds=[0.6, 0.552], crop=n, seed=42 + rank)
        stream = ds.batches(batch_size, rank=rank, world=world)
    t0 = time.perf_counter()
    for step in range(args.max_steps):
        if stream is not None:
            batch = next(stream, None)
            if batch is None:
                break
            loss = trainer.training_step(batch)
            if rank == 0 and (step + 1) % args.log_every == 0:
                dt = time.perf_counter() - t0
                print(f"step {step + 1}: loss {loss.item():.5f}  {(step + 1) * batch_size * world / dt:.1f} samples/s")
            continue
        raw = synthetic_batch(batch_size, n, 42 + step * world + rank, device=device)
        if args.model == "VDM":
            batch = raw
        else:                                                  # trainSFM3D160_...:71-72
            batch = {"x0": raw["conditioning"], "x1": raw["x"], "conditioning_values": raw["conditioning_values"]}
        loss = trainer.training_step(batch)
        if rank == 0 and (step + 1) % args.log_every == 0:
            dt = time.perf_counter() - t0
            print(f"step {step + 1}: loss {loss.item():.5f}  {(step + 1) * batch_size * world / dt:.1f} samples/s")
        if rank == 0 and (step + 1) % args.ckpt_every == 0:
            os.makedirs(args.ckpt_dir, exist_ok=True)
            name = f"{args.model}_{args.field_in}_{args.field_out}_c_c_{n}_step={step + 1}.ckpt"
            torch.save({"state_dict": model.state_dict()}, os.path.join(args.ckpt_dir, name))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
