#!/usr/bin/env python
"""Summary statistics of generated ensembles: the P(k) / r(k) part of the reference's ``calc_SS.py``
(get_pk_3d / get_pk_2d: calc_SS.py:67-75; posterior mean/std: :150-152) on the cuFFT + binning kernels.

    python scripts/calc_SS.py SAVE_PATH [--truth truth.npy]

reads every ``gen_*.npy`` (rep, 1, N, N, N) written by scripts/generate_3D.py, un-normalises the fields
(``10**(x*std+mean)-1``), and writes ``SAVE_PATH/<file>_summary.npz`` next to every ensemble file with: the 3-D P(k) of every
realisation, the projected 2-D P(k) of the half / quarter slabs, per-realisation mean/std, posterior mean/std of the
3-D field over the ensemble (``post_means`` / ``post_stds``, calc_SS.py:150-152) and of the half slab and,
when a truth field is given, the cross-correlation coefficient r(k) of every realisation with it.
Also the log-PDF histograms (get_logpdf_3d / get_logpdf_2d: calc_SS.py:51-65) on the histogram kernel.
Files are independent units: file i goes to rank i mod world_size.  (The wavelet-scattering statistics
(``mltools.archive.LWT``) are out of scope: SURVEY.md section 8f.)
"""
import argparse
import glob
import os

import numpy as np

from _common import init_distributed, unnorm_mcdm

import torch
import torch.distributed as dist

from vdm4cdm_b200 import utils
from vdm4cdm_b200.trainer import shard_indices


def get_pk_3d(fields):
    fields_u = fields / fields.sum((2, 3, 4), keepdims=True)
    ks, pk, _ = utils.pk(fields_u)
    return ks.cpu().numpy(), pk.cpu().numpy()


def get_pk_2d(fields):
    fields_u = fields / fields.sum((2, 3), keepdims=True)
    ks, pk, _ = utils.pk(fields_u)
    return ks.cpu().numpy(), pk.cpu().numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("save_path")
    ap.add_argument("--truth", default=None, help=".npy with the true normalised field (1, 1, N, N, N)")
    ap.add_argument("--chunk", type=int, default=16, help="realisations transformed per batched FFT")
    args = ap.parse_args()
    rank, world, device = init_distributed()
    # gen_{count}.npy (CV run types) and {name}_{rep}.npy (1P run types, generate_3D_1P.py:70)
    files = sorted(f for f in glob.glob(os.path.join(args.save_path, "*.npy")) if not f.endswith("_old.npy"))
    assert files, f"no ensemble .npy files under {args.save_path}"
    truth = None if args.truth is None else unnorm_mcdm(torch.from_numpy(np.load(args.truth)).float().to(device))
    for i in shard_indices(len(files), rank, world):
        data = np.load(files[i])
        n = data.shape[-1]
        half, quarter = n // 2, n // 4
        out = {"pk3d": [], "pk2d_half": [], "pk2d_quarter": [], "cc": [], "logpdf3d": [], "logpdf2d_half": [],
               "logpdf2d_quarter": [], "mean3d": [], "std3d": [], "mean2d_half": [], "std2d_half": [],
               "mean2d_quarter": [], "std2d_quarter": []}
        mean_acc = torch.zeros((1, 1, n, n), dtype=torch.float64, device=device)
        sq_acc = torch.zeros_like(mean_acc)
        mean3_acc = torch.zeros((1, 1, n, n, n), dtype=torch.float64, device=device)     # calc_SS.py:150-152
        sq3_acc = torch.zeros_like(mean3_acc)
        for j0 in range(0, data.shape[0], args.chunk):
            x = unnorm_mcdm(torch.from_numpy(data[j0:j0 + args.chunk]).float().to(device))
            k3, p3 = get_pk_3d(x)
            slab_h, slab_q = x[:, :, :half].sum(2), x[:, :, :quarter].sum(2)
            k2, p2h = get_pk_2d(slab_h)
            _, p2q = get_pk_2d(slab_q)
            out["pk3d"].append(p3); out["pk2d_half"].append(p2h); out["pk2d_quarter"].append(p2q)
            out["logpdf3d"].append(utils.get_logpdf_3d(x).cpu().numpy())               # calc_SS.py:51-57
            out["logpdf2d_half"].append(utils.get_logpdf_2d(slab_h).cpu().numpy())     # calc_SS.py:59-65
            out["logpdf2d_quarter"].append(utils.get_logpdf_2d(slab_q).cpu().numpy())
            for name, f in (("3d", x), ("2d_half", slab_h), ("2d_quarter", slab_q)):       # calc_SS.py:80-81,85-86,92-93
                flat = f.reshape(f.shape[0], -1)
                out[f"mean{name}"].append(flat.mean(1).cpu().numpy())
                out[f"std{name}"].append(flat.std(1).cpu().numpy())
            mean3_acc += x.double().sum(0, keepdim=True)
            sq3_acc += (x.double() ** 2).sum(0, keepdim=True)
            mean_acc += slab_h.double().sum(0, keepdim=True)
            sq_acc += (slab_h.double() ** 2).sum(0, keepdim=True)
            if truth is not None:
                t = truth.expand(x.shape[0], -1, -1, -1, -1).contiguous()
                _, cc = utils.get_ccs(x / x.sum((2, 3, 4), keepdims=True), t / t.sum((2, 3, 4), keepdims=True))
                out["cc"].append(cc.cpu().numpy())
        rep = data.shape[0]
        mean = mean_acc / rep
        std = torch.sqrt(torch.clamp(sq_acc / rep - mean ** 2, min=0.0) * rep / max(rep - 1, 1))
        mean3 = mean3_acc / rep
        std3 = torch.sqrt(torch.clamp(sq3_acc / rep - mean3 ** 2, min=0.0) * rep / max(rep - 1, 1))
        res = {"k3d": k3[0], "k2d": k2[0], "post_means": mean3.float().cpu().numpy(), "post_stds": std3.float().cpu().numpy(),
               "post_mean_half": mean.float().cpu().numpy(), "post_std_half": std.float().cpu().numpy()}
        res.update({k: np.concatenate(v, axis=0) for k, v in out.items() if v})
        np.savez(files[i].replace(".npy", "_summary.npz"), **res)
        print(f"[rank {rank}] {os.path.basename(files[i])}: {rep} realisations, P(k) in {len(res['k3d'])} bins")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
