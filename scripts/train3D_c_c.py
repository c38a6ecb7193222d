#!/usr/bin/env python
"""Training entry point for the 3-D conditional models: the reference's
``trainVDM3D{,128,160,192,224}_c_c_from_field_name_thick_lowbatch.py`` and ``trainSFM3D*_c_c_..._lowbatch.py``
(positional CLI ``field_in field_out cropsize`` as there; hyper-parameters from those files: chs, batch_size,
dropout 0.1, norm_groups 8, gamma_max 13.3, lr 3e-4, gradient_clip_val 0.5, seed 42) with the Lightning /
Comet orchestration replaced by ``vdm4cdm_b200.trainer.Trainer``:

    torchrun --nproc-per-node 8 scripts/train3D_c_c.py Mstar Mcdm 128 --model VDM --synthetic --max-steps 1000
    python scripts/trainVDM3D128_c_c_from_field_name_thick_lowbatch.py Mstar Mcdm 128      # the reference's file names

One process per GPU, each with its own micro-batch of ``batch_size`` samples, gradients all-reduced in one flat
bucket over NCCL/NVLink.  What ``lightning.Trainer(max_steps=1_000_000, val_check_interval=..., gradient_clip_val=0.5,
callbacks=[ModelCheckpoint(save_top_k=-1, every_n_train_steps=10_000)]).fit`` does around the step
(trainVDM3D128_...:38-49) is done here, on every data path:

  * checkpoints every ``--ckpt-every`` steps and at the end: ``{"state_dict", "global_step", "optimizer", ...}``
    (``Trainer.state_dict``), loadable by ``vdm4cdm_b200.utils.get_model`` (src/utils.py:467) and by ``--resume``
    (the manual ``vdm.load_state_dict(ckpt['state_dict'])`` of trainVDM3D_c_c_...:133-135, plus the optimizer state);
  * validation every ``--val-check-interval`` steps: the loss over the validation loader and, on rank 0, what the
    reference's ``draw_figure`` computes from ``draw_samples`` (trainVDM3D128_...:97-112, src/utils.py:130-195) -- the
    P(k) of the truth / sample / conditioning and r(k) of sample vs truth -- as numbers in ``--log-jsonl``, no figure.

``--synthetic`` stands in for the CAMELS LH set (see scripts/_common.py).
"""
import argparse
import json
import os
import time

from _common import init_distributed, synthetic_batch, unnorm_mcdm

import torch
import torch.distributed as dist

from mltools.models import sfm_model, vdm_model
from mltools.networks import networks
from vdm4cdm_b200 import utils
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
# Trainer(val_check_interval=...) of each reference script: (model, grid size in the script's NAME or None) -> steps
VAL_CHECK_INTERVAL = {("VDM", 128): 1000, ("VDM", 160): 5000, ("VDM", 192): 5000, ("VDM", 224): 5000, ("VDM", None): 5000,
                      ("SFM", 128): 1000, ("SFM", 160): 1000, ("SFM", 192): 1000, ("SFM", None): 1000}


def preset(model: str, cropsize: int):
    return PRESETS.get((model, cropsize), BASE_PRESET)


def script_preset(model: str, script_grid):
    """(chs, batch_size, dataset_name, val_check_interval) of the reference script ``train{model}3D{script_grid}_c_c_...``
    (``script_grid=None``: the base script).  A reference script takes its network / dataset from its own literals, not
    from the ``cropsize`` argument."""
    chs, batch, name = PRESETS[(model, script_grid)] if script_grid is not None else BASE_PRESET
    return chs, batch, name, VAL_CHECK_INTERVAL[(model, script_grid)]


def validate(model, kind, val_batches, device, world, n_sampling_steps, sample_batch, seed):
    """Validation numbers of one ``val_check_interval`` tick: mean loss over ``val_batches`` (all-reduced over ranks) and,
    for the VDM, sample-vs-truth statistics of ``draw_samples`` on the first validation batch (what draw_figure plots:
    ``pk_func`` = utils.pk(unnorm(f) / sum), ``cc_func`` = utils.get_ccs of the two, trainVDM3D128_...:97-103)."""
    model.eval()
    out = {}
    losses = []
    first = None
    with torch.no_grad():
        for batch in val_batches:
            if first is None:
                first = batch
            losses.append(model.validation_step(batch).float())
    tot = torch.stack([torch.stack(losses).sum() if losses else torch.zeros((), device=device),
                       torch.tensor(float(len(losses)), device=device)])
    if world > 1:
        dist.all_reduce(tot)
    out["val_loss"] = (tot[0] / tot[1].clamp_min(1.0)).item()
    out["val_batches"] = int(tot[1].item())
    rank = dist.get_rank() if world > 1 else 0
    if kind == "VDM" and first is not None and n_sampling_steps > 0 and rank == 0:
        b = min(sample_batch, first["x"].shape[0])
        truth = first["x"][:b]
        cond = first["conditioning"][:b].contiguous()
        vals = [v[:b].contiguous() for v in first["conditioning_values"]]
        samples = model.draw_samples(batch_size=b, n_sampling_steps=n_sampling_steps, s_conditioning=cond,
                                     v_conditionings=vals, seed=seed)
        t_un, s_un = unnorm_mcdm(truth), unnorm_mcdm(samples)
        t_un = (t_un / t_un.sum((2, 3, 4), keepdim=True)).contiguous()
        s_un = (s_un / s_un.sum((2, 3, 4), keepdim=True)).contiguous()
        ks, pk_t, _ = utils.pk(t_un)
        _, pk_s, _ = utils.pk(s_un)
        _, cc = utils.get_ccs(t_un, s_un, full=False)
        out.update(k=ks[0].tolist(), pk_truth=pk_t[0].tolist(), pk_sample=pk_s[0].tolist(),
                   pk_ratio=(pk_s[0] / pk_t[0]).tolist(), cc=cc[0].tolist(),
                   sample_mean=samples.mean().item(), sample_std=samples.std().item(), n_sampling_steps=n_sampling_steps)
    if world > 1:
        dist.barrier()
    return out


def main(model_kind=None, script_grid="cli"):
    """``model_kind`` / ``script_grid`` are fixed by the thin launchers that carry the reference's file names
    (``script_grid`` = the grid size in the name, None for the base scripts); ``--model`` otherwise."""
    ap = argparse.ArgumentParser()
    ap.add_argument("field_in")
    ap.add_argument("field_out")
    ap.add_argument("cropsize", type=int)
    if model_kind is None:
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
    ap.add_argument("--resume", default=None, help="checkpoint to continue from (weights, optimizer state, step, RNG)")
    ap.add_argument("--val-check-interval", type=int, default=None, help="default: the reference script's (1000 or 5000)")
    ap.add_argument("--val-batches", type=int, default=0, help="validation batches per check (0: the whole validation loader)")
    ap.add_argument("--val-sampling-steps", type=int, default=250, help="draw_samples steps at validation (0: loss only)")
    ap.add_argument("--log-jsonl", default=None, help="metrics file (default: <ckpt-dir>/metrics.jsonl)")
    ap.add_argument("--log-every", type=int, default=10)
    args = ap.parse_args()
    kind = model_kind or args.model
    rank, world, device = init_distributed()
    torch.manual_seed(42)                                     # seed_everything(42)
    n = args.cropsize
    if script_grid == "cli":
        chs, batch_size, dataset_name = preset(kind, n)
        val_every = VAL_CHECK_INTERVAL.get((kind, n if (kind, n) in PRESETS else None))
    else:
        chs, batch_size, dataset_name, val_every = script_preset(kind, script_grid)
    batch_size = args.batch_size or batch_size
    dataset_name = args.dataset_name or dataset_name
    chs = args.chs or chs
    val_every = args.val_check_interval or val_every
    net = networks.CUNet(shape=(1, n, n, n), chs=chs, s_conditioning_channels=1, v_conditioning_dims=[6],
                         t_conditioning=True, norm_groups=NORM_GROUPS, mid_attn=False, dropout_prob=DROPOUT_PROB,
                         conv_padding_mode="circular" if n == 256 else "zeros", n_attention_heads=4)
    if kind == "VDM":
        model = vdm_model.LightVDM(score_model=net, draw_figure=None, gamma_max=GAMMA_MAX, learning_rate=LEARNING_RATE)
    else:
        model = sfm_model.LightSFM(velocity_model=net, draw_figure=None, learning_rate=LEARNING_RATE)
    model = model.to(device)
    trainer = Trainer(model, gradient_clip_val=GRADIENT_CLIP_VAL)
    start_step, resumed = 0, None
    if args.resume:
        resumed = trainer.load_checkpoint(args.resume)
        start_step = int(resumed["global_step"])
        if rank == 0:
            print(f"resumed from {args.resume} at step {start_step}")
    key_c, key_x = ("conditioning", "x") if kind == "VDM" else ("x0", "x1")

    def as_model_batch(raw):
        if kind == "VDM":
            return raw
        return {"x0": raw["conditioning"], "x1": raw["x"], "conditioning_values": raw["conditioning_values"]}   # trainSFM3D160_...:71-72

    stream, val_loader, loader = None, None, None
    if not args.synthetic:
        # trainVDM3D128_...:75-89: LH set of the re-gridded boxes, stage "fit", mmap=False (boxes resident in HBM)
        from vdm4cdm_b200.dataset import get_dataset
        dm = get_dataset(dataset_name=dataset_name, suite_name=args.suite_name,
                         return_func=lambda fields, params: {key_c: fields[0], key_x: fields[1], "conditioning_values": [params]},
                         set_name="LH", z_name="z_0.0", channel_names=[args.field_in, args.field_out], stage="fit",
                         batch_size=batch_size, cropsize=n, mmap=args.host_boxes, data_root=args.data_root, device=device,
                         seed=42, rank=rank, world=world)
        loader = dm.train_dataloader()
        if resumed is not None and "data" in resumed:
            loader.load_state(resumed["data"])                   # continue the sample stream where the run stopped

        def epochs():
            for _ in range(args.max_epochs):
                yield from loader
        stream = epochs()
        val_loader = dm.val_dataloader()
    elif args.synthetic_boxes:
        # the full data path: raw boxes resident in HBM -> vdm_augment_crop (periodic crop, log-normalise, flip, permute)
        from vdm4cdm_b200.dataset import DeviceAstroDataset
        S, n_sims = max(2 * n, 256) if n < 256 else n, 4
        g = torch.Generator().manual_seed(7)
        base = torch.randn((n_sims, S, S, S), generator=g)
        raw_x = (10.0 ** (base * 0.552 + 10.019)).to(device)
        raw_c = (10.0 ** ((0.7 * base + 0.3 * torch.randn((n_sims, S, S, S), generator=g)) * 0.6 + 8.0)).to(device)
        params = torch.rand((n_sims, 6), generator=g)
        ds = DeviceAstroDataset([raw_c, raw_x], params, lambda fields, params: {key_c: fields[0], key_x: fields[1],
                                                                               "conditioning_values": [params]},
                                alphas=[1.0, 1.0], means=[8.0, 10.019], stds=[0.6, 0.552], crop=n, seed=42 + rank)
        stream = ds.batches(batch_size, rank=rank, world=world)

    def next_batch(step):
        if stream is not None:
            return next(stream, None)
        return as_model_batch(synthetic_batch(batch_size, n, 42 + step * world + rank, device=device))

    def val_batches():
        if val_loader is not None:
            for i, b in enumerate(val_loader):
                if args.val_batches and i >= args.val_batches:
                    break
                yield b
        else:          # synthetic stand-in: fixed validation fields, different on every rank
            for i in range(args.val_batches or 2):
                yield as_model_batch(synthetic_batch(batch_size, n, 9_000_000 + i * world + rank, device=device))

    log_path = args.log_jsonl or os.path.join(args.ckpt_dir, "metrics.jsonl")
    if rank == 0:
        os.makedirs(args.ckpt_dir, exist_ok=True)
        os.makedirs(os.path.dirname(os.path.abspath(log_path)), exist_ok=True)

    def log(record):
        if rank == 0:
            with open(log_path, "a") as f:
                f.write(json.dumps(record) + "\n")

    def save(step):
        if rank == 0:
            name = f"{kind}_{args.field_in}_{args.field_out}_c_c_{n}_step={step}.ckpt"
            trainer.save_checkpoint(os.path.join(args.ckpt_dir, name),
                                    extra=None if loader is None else {"data": loader.state()})
            print(f"step {step}: checkpoint {name}")

    t0 = time.perf_counter()
    step, last_saved = start_step, start_step
    while step < args.max_steps:
        batch = next_batch(step)
        if batch is None:                                   # --max-epochs exhausted
            break
        loss = trainer.training_step(batch)
        step += 1
        if step % args.log_every == 0:
            lv = loss.item()
            if rank == 0:
                dt = time.perf_counter() - t0
                rate = (step - start_step) * batch_size * world / dt
                print(f"step {step}: loss {lv:.5f}  {rate:.1f} samples/s")
                log({"step": step, "train_loss": lv, "samples_per_s": rate})
        if val_every and step % val_every == 0:
            trainer.sync_inference_weights()
            rec = validate(model, kind, val_batches(), device, world, args.val_sampling_steps, batch_size, seed=42 + step)
            rec["step"] = step
            log(rec)
            if rank == 0:
                tail = f"  r(k) first/last bin {rec['cc'][0]:.3f}/{rec['cc'][-1]:.3f}" if "cc" in rec else ""
                print(f"step {step}: val_loss {rec['val_loss']:.5f} over {rec['val_batches']} batches{tail}")
        if step % args.ckpt_every == 0:
            save(step)
            last_saved = step
    if step != last_saved:
        save(step)                                          # final weights, whatever ended the loop
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
