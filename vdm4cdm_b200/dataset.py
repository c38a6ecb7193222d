"""Device-resident training data: the reference's ``AstroDataset`` / ``AstroDataModule`` batch pipeline
(src/dataset/CAMELS_3D_dataset.py:19-74, 76-199) with the simulation boxes held in HBM and every batch element
produced by ONE gather kernel (``vdm_augment_crop``: periodic crop + log-normalisation + flip + axis permutation;
src/dataset/augmentation.py:8-127).

Same batch schema: ``return_func(fields=[...], params=...)`` per sample, collated like ``AstroDataModule.collate_fn``
(:158-171).  The random choices (crop index, anchor shift, flip axes, permutation) are drawn on the host from a
seeded ``torch.Generator`` exactly where the reference draws them; they are a few integers per sample.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np
import torch

from . import ops


def norm_func(field, alpha, mean, std):
    """AstroDataModule.norm_func (CAMELS_3D_dataset.py:152-156)."""
    return (torch.log10(field + alpha) - mean) / std


def unnorm_func(field, alpha, mean, std):
    """AstroDataModule.unnorm_func (CAMELS_3D_dataset.py:146-150)."""
    return 10 ** (field * std + mean) - alpha


class DeviceAstroDataset:
    """``fields``: list of CUDA fp32 tensors (n_sims, S, S, S) of RAW (un-normalised) boxes, one per channel name;
    ``params``: (n_sims, P).  ``get_batch(indices)`` returns the collated batch dict."""

    def __init__(self, fields: Sequence[torch.Tensor], params: torch.Tensor, return_func: Callable, alphas, means, stds,
                 do_crop: bool = True, crop: int = 128, aug_shift: bool = True, augment: bool = True, seed: int = 42):
        assert len(fields) >= 1 and all(f.is_cuda and f.dtype == torch.float32 and f.dim() == 4 for f in fields), \
            "fields must be CUDA fp32 (n_sims, S, S, S) tensors (this package has no CPU path)"
        self.fields = [f.contiguous() for f in fields]
        self.n_sims, self.fullsize = fields[0].shape[0], fields[0].shape[-1]
        assert all(f.shape == fields[0].shape for f in fields) and len(params) == self.n_sims
        self.params = params.to(fields[0].device, torch.float32)
        self.return_func = return_func
        self.alphas, self.means, self.stds = list(alphas), list(means), list(stds)
        self.do_crop, self.crop, self.aug_shift, self.augment = do_crop, (crop if do_crop else self.fullsize), aug_shift, augment
        # anchors of the crop grid (augmentation.py:97-106)
        r = range(0, self.fullsize, self.crop)
        self.anchors = np.array([(a, b, c) for a in r for b in r for c in r], dtype=np.int64) if do_crop else np.zeros((1, 3), np.int64)
        self.ncrops = len(self.anchors)
        self.gen = torch.Generator().manual_seed(seed)

    def __len__(self):
        return self.n_sims * self.ncrops

    def draw(self, idx: int):
        """(simulation index, anchor, flip mask, permutation) of sample ``idx``; consumes the host generator like the
        reference's Crop / Flip / Permutate do."""
        bidx, icrop = divmod(int(idx), self.ncrops)
        anchor = self.anchors[icrop].copy()
        if self.do_crop and self.aug_shift:
            anchor += torch.randint(self.crop, (3,), generator=self.gen).numpy()
        if self.augment:
            flip = torch.randint(2, (3,), generator=self.gen).numpy()
            perm = torch.randperm(3, generator=self.gen).numpy()
        else:
            flip, perm = np.zeros(3, np.int64), np.arange(3)
        return bidx, anchor, flip, perm

    def get_batch(self, indices: Sequence[int]):
        n, b = self.crop, len(indices)
        outs = [torch.empty((b, 1, n, n, n), dtype=torch.float32, device=self.fields[0].device) for _ in self.fields]
        samples = []
        for j, idx in enumerate(indices):
            bidx, anchor, flip, perm = self.draw(idx)
            for k, f in enumerate(self.fields):
                ops.augment_crop(f[bidx], (n, n, n), anchor, flip, perm, alpha=self.alphas[k], mean=self.means[k],
                                 std=self.stds[k], do_log=True, out=outs[k][j, 0])
            samples.append(self.return_func(fields=[o[j] for o in outs], params=self.params[bidx]))
        return collate(samples)

    def batches(self, batch_size: int, rank: int = 0, world: int = 1, shuffle: bool = True):
        """Endless stream of batches; sample ids are sharded ``i -> rank i mod world`` over a shuffled epoch order."""
        while True:
            order = torch.randperm(len(self), generator=self.gen).tolist() if shuffle else list(range(len(self)))
            mine = order[rank::world]
            for i in range(0, len(mine) - batch_size + 1, batch_size):
                yield self.get_batch(mine[i:i + batch_size])


def collate(batch: List[dict]) -> dict:
    """AstroDataModule.collate_fn (CAMELS_3D_dataset.py:158-171)."""
    out = {}
    b0 = batch[0]
    for key in b0.keys():
        if b0[key] is None:
            out[key] = None
        elif isinstance(b0[key], torch.Tensor):
            out[key] = torch.stack([b[key] for b in batch], dim=0)
        elif isinstance(b0[key], list):
            out[key] = [torch.stack([b[key][i] for b in batch], dim=0) for i in range(len(b0[key]))]
        else:
            raise ValueError(f"Type of {key} not recognized")
    return out
