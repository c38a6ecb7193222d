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
                 do_crop: bool = True, crop: int = 128, aug_shift: bool = True, augment: bool = True, seed: int = 42,
                 device=None):
        # CUDA tensors stay resident in HBM; numpy arrays (also memory-mapped ones) stay on the host and the one
        # box a sample needs is staged to ``device`` when it is drawn (67 MB for a 256^3 box)
        assert len(fields) >= 1 and all(f.ndim == 4 for f in fields), "fields must be (n_sims, S, S, S)"
        assert all((f.is_cuda and f.dtype == torch.float32) if torch.is_tensor(f) else hasattr(f, "__getitem__")
                   for f in fields), "fields must be CUDA fp32 tensors or host numpy arrays (there is no CPU compute path)"
        self.fields = [f.contiguous() if torch.is_tensor(f) else f for f in fields]
        self.n_sims, self.fullsize = fields[0].shape[0], fields[0].shape[-1]
        assert all(tuple(f.shape) == tuple(fields[0].shape) for f in fields) and len(params) == self.n_sims
        cuda_fields = [f for f in fields if torch.is_tensor(f)]
        self.device = torch.device(device) if device is not None else (cuda_fields[0].device if cuda_fields else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0))
        self.params = torch.as_tensor(np.asarray(params) if not torch.is_tensor(params) else params).to(torch.float32)
        self.return_func = return_func
        self.alphas, self.means, self.stds = list(alphas), list(means), list(stds)
        self.do_crop, self.crop, self.aug_shift, self.augment = do_crop, (crop if do_crop else self.fullsize), aug_shift, augment
        # anchors of the crop grid (augmentation.py:97-106)
        r = range(0, self.fullsize, self.crop)
        self.anchors = np.array([(a, b, c) for a in r for b in r for c in r], dtype=np.int64) if do_crop else np.zeros((1, 3), np.int64)
        self.ncrops = len(self.anchors)
        self.seed = int(seed)
        self.gen = torch.Generator().manual_seed(self.seed)

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
        outs = [torch.empty((b, 1, n, n, n), dtype=torch.float32, device=self.device) for _ in self.fields]
        params = self.params.to(self.device)
        samples = []
        for j, idx in enumerate(indices):
            bidx, anchor, flip, perm = self.draw(idx)
            for k, f in enumerate(self.fields):
                box = f[bidx] if torch.is_tensor(f) else \
                    torch.from_numpy(np.array(f[bidx], dtype=np.float32)).to(self.device, non_blocking=True)
                ops.augment_crop(box, (n, n, n), anchor, flip, perm, alpha=self.alphas[k], mean=self.means[k],
                                 std=self.stds[k], do_log=True, out=outs[k][j, 0])
            samples.append(self.return_func(fields=[o[j] for o in outs], params=params[bidx]))
        return collate(samples)

    def batches(self, batch_size: int, rank: int = 0, world: int = 1, shuffle: bool = True, seed: int = 42):
        """Endless stream of full batches; the epoch order is a function of (seed, epoch) shared by all ranks, padded to a
        multiple of ``world * batch_size`` and sharded ``i -> rank i mod world`` (see ``_Loader``)."""
        loader = _Loader(self, range(len(self)), batch_size, shuffle, rank=rank, world=world, drop_last=True, seed=seed)
        while True:
            yield from loader


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


# ------------------------------------------------------------------------------------------------------------------
# AstroDataModule: the reference's LightningDataModule (CAMELS_3D_dataset.py:76-199) without Lightning.
#
# The reference reads three JSON tables with absolute paths on its author's cluster (CAMELS_3D_dataset.py:10-17:
# per-channel log offsets, per-channel mean/std of the log field, and the .npy path of every
# dataset/suite/set/redshift/channel).  Here the two small tables are constants and the grid / parameter files are
# looked up under ONE directory (``data_root`` or $VDM4CDM_DATA_ROOT) by CAMELS' own file names.
ALPHAS_3D = {"Mcdm": 1.0, "Mstar": 1.0, "B": 1.0, "HI": 1.0, "Mgas": 1.0, "MgFe": 1.0, "ne": 1.0, "P": 1.0, "T": 1.0,
             "Z": 1.0, "Go7": 2.0, "Go8": 2.0, "Go9": 2.0}
NORMALIZATIONS_3D = {"Mcdm": (10.019186475678042, 0.5520203178284999), "Mstar": (0.010429391444558287, 0.3219291117577123),
                     "Go7": (0.0, 1.0), "Go8": (0.0, 1.0), "Go9": (0.0, 1.0)}
CV_EXCLUDED = (2, 8, 17)            # CAMELS_3D_dataset.py:114-119, 126-131: three CV boxes the reference drops
REFERENCE_FULLSIZE = 256            # get_dataset: do_crop = (cropsize != 256)   (CAMELS_3D_dataset.py:225)


def grid_resolution(dataset_name: str) -> int:
    """"CMD" holds the 256^3 grids, "CMD_128" / "CMD_160" / "CMD_192" / "CMD_224" the re-gridded boxes
    (data_source_3d.json of the reference)."""
    return int(dataset_name.split("_")[1]) if "_" in dataset_name else REFERENCE_FULLSIZE


def grid_file_name(channel_name: str, selection: dict) -> str:
    """File name of a CAMELS 3-D grid stack, as in the reference's data_source_3d.json."""
    z = selection.get("z_name", "z_0.0").replace("z_", "z=")
    return (f"Grids_{channel_name}_{selection['suite_name']}_{selection['set_name']}_"
            f"{grid_resolution(selection['dataset_name'])}_{z}.npy")


def params_file_name(selection: dict) -> str:
    return f"params_{selection['set_name']}_{selection['suite_name']}.txt"     # CAMELS_3D_dataset.py:125


def epoch_permutation(n: int, seed: int, epoch: int) -> List[int]:
    """Shuffled order of ``n`` samples for one epoch: a function of (seed, epoch) ONLY, so every rank computes the same
    order without talking to the others and without touching the generator the augmentation draws come from."""
    g = torch.Generator().manual_seed((int(seed) * 1000003 + int(epoch)) % (2 ** 63 - 1))
    return torch.randperm(n, generator=g).tolist()


class _Loader:
    """Iterable of collated device batches over a fixed list of sample ids (what a ``DataLoader`` over a
    ``Subset`` is for the reference): ``shuffle`` reshuffles every epoch (``epoch_permutation``), a single rank keeps
    the last short batch like the reference's ``DataLoader``.

    With ``world > 1`` the (shuffled) epoch order is first padded, by wrapping around, to a multiple of
    ``world * batch_size`` and then sharded ``i -> rank i mod world`` (``torch.utils.data.DistributedSampler``'s rule,
    which Lightning applies to the reference under DDP, with the padding rounded up to whole batches): every rank runs
    the SAME number of full batches per epoch, so the per-step gradient all-reduce can never be left waiting for a rank
    that ran out of data, and the static-shape CUDA graph of the training step is never invalidated by a short batch."""

    def __init__(self, data: DeviceAstroDataset, ids: Sequence[int], batch_size: int, shuffle: bool, rank: int = 0,
                 world: int = 1, drop_last: bool = False, seed: int = 42):
        self.data, self.ids, self.batch_size, self.shuffle = data, list(ids), int(batch_size), shuffle
        self.rank, self.world, self.drop_last, self.seed = rank, world, drop_last, seed
        self.epoch, self.cursor, self._skip = 0, 0, 0

    def _per_rank(self) -> int:
        if self.world == 1:
            return len(self.ids)
        unit = self.world * self.batch_size
        return -(-len(self.ids) // unit) * unit // self.world

    def __len__(self):
        n = self._per_rank()
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def epoch_order(self, epoch: int) -> List[int]:
        """Sample ids of this rank for ``epoch``, in the order they are drawn."""
        order = self.ids
        if self.shuffle:
            order = [self.ids[i] for i in epoch_permutation(len(self.ids), self.seed, epoch)]
        if self.world > 1 and len(order) > 0:
            total = self._per_rank() * self.world
            order = (order * (-(-total // len(order))))[:total]
        return order[self.rank::self.world]

    def state(self) -> dict:
        """Position in the sample stream, for checkpoints: (epoch, batches already drawn in it)."""
        return {"epoch": self.epoch, "cursor": self.cursor}

    def load_state(self, state: dict) -> None:
        """Continue after ``state``: the next ``iter()`` replays the host-side augmentation draws of the batches that were
        already consumed (no kernels) and yields the following ones -- the same stream as an uninterrupted run, on every
        rank, because order and augmentation are functions of (seed, rank, epoch) only."""
        self.epoch, self._skip = int(state["epoch"]), int(state["cursor"])

    def __iter__(self):
        mine = self.epoch_order(self.epoch)
        # augmentation draws of epoch e come from a generator seeded by (dataset seed, e): epoch 0 continues the
        # constructor's seeding, so a fresh dataset replays it (tests/test_gpu_dataset.py)
        if self.epoch > 0 or self._skip:
            self.data.gen.manual_seed(self.data.seed + 104729 * self.epoch)
        skip, self._skip = self._skip, 0
        self.cursor = 0
        for bi, i in enumerate(range(0, len(mine), self.batch_size)):
            chunk = mine[i:i + self.batch_size]
            if len(chunk) < self.batch_size and self.drop_last:
                break
            if bi < skip:
                for idx in chunk:
                    self.data.draw(idx)
                self.cursor = bi + 1
                continue
            batch = self.data.get_batch(chunk)
            self.cursor = bi + 1
            yield batch
        self.epoch += 1
        self.cursor = 0


class AstroDataModule:
    """``AstroDataModule(selection, channel_names, return_func, stage, batch_size, do_crop, cropsize, ndim,
    num_workers, mmap)`` (CAMELS_3D_dataset.py:76-144), 3-D only.

    ``stage="fit"``: random periodic crops with anchor shift + Flip + Permutate, 95 % / 5 % train/validation split;
    ``stage="test"``: the deterministic crop grid, no augmentation.  ``num_workers`` is accepted and ignored -- there
    are no worker processes: a batch element is one gather kernel on boxes that already sit in HBM (``mmap=False``)
    or one 67 MB host-to-device copy plus that kernel (``mmap=True``: boxes stay in the page cache)."""

    def __init__(self, selection: dict, channel_names: Sequence[str], return_func: Callable, stage: str = "fit",
                 batch_size: int = 1, do_crop: bool = False, cropsize: int = 256, ndim: int = 3, num_workers: int = 1,
                 mmap: bool = True, data_root: Optional[str] = None, device=None, seed: int = 42, rank: int = 0,
                 world: int = 1):
        import os
        if ndim != 3:
            raise NotImplementedError("vdm4cdm_b200.dataset.AstroDataModule covers the 3-D data module only")
        assert stage in ["fit", "test"], f"stage {stage} not recognized"
        self.selection, self.channel_names, self.stage, self.batch_size = selection, list(channel_names), stage, batch_size
        self.do_crop, self.cropsize, self.ndim, self.num_workers, self.mmap = do_crop, cropsize, ndim, num_workers, mmap
        self.rank, self.world, self.seed, self._train_loader = rank, world, seed, None
        self.alphas = [ALPHAS_3D[c] for c in self.channel_names]
        self.means = [NORMALIZATIONS_3D[c][0] for c in self.channel_names]
        self.stds = [NORMALIZATIONS_3D[c][1] for c in self.channel_names]
        root = data_root or os.environ.get("VDM4CDM_DATA_ROOT")
        if root is None:
            raise FileNotFoundError("AstroDataModule: pass data_root= or set VDM4CDM_DATA_ROOT to the directory that holds "
                                    f"{grid_file_name(self.channel_names[0], selection)} and {params_file_name(selection)}")
        is_cv = selection["set_name"] == "CV"
        fields = []
        for c in self.channel_names:
            path = os.path.join(root, grid_file_name(c, selection))
            if not os.path.exists(path):
                raise FileNotFoundError(path)
            f = np.load(path, mmap_mode="r" if mmap else None)
            if is_cv:
                keep = np.ones(len(f), dtype=bool)
                keep[[i for i in CV_EXCLUDED if i < len(f)]] = False
                f = f[np.nonzero(keep)[0]] if not mmap else _RowSubset(f, np.nonzero(keep)[0])
            fields.append(f)
        params = np.loadtxt(os.path.join(root, params_file_name(selection)), ndmin=2)
        if is_cv:
            keep = np.ones(len(params), dtype=bool)
            keep[[i for i in CV_EXCLUDED if i < len(params)]] = False
            params = params[keep]
        if not mmap:
            if device is None and not torch.cuda.is_available():
                raise RuntimeError("AstroDataModule(mmap=False) keeps the boxes in HBM and needs a CUDA device")
            dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
            fields = [torch.from_numpy(np.ascontiguousarray(f, dtype=np.float32)).to(dev) for f in fields]
        self.data = DeviceAstroDataset(fields, torch.from_numpy(params), return_func, self.alphas, self.means, self.stds,
                                       do_crop=do_crop, crop=cropsize, aug_shift=(stage == "fit"), augment=(stage == "fit"),
                                       seed=seed + 7919 * rank, device=device)     # augmentation draws differ per rank
        if stage == "fit":
            n_train = int(len(self.data) * 0.95)                           # CAMELS_3D_dataset.py:137-139
            order = torch.randperm(len(self.data), generator=torch.Generator().manual_seed(seed)).tolist()
            self.train_ids, self.valid_ids = order[:n_train], order[n_train:]
        else:
            self.test_ids = list(range(len(self.data)))

    def unnorm_func(self, field, i_channel):
        return unnorm_func(field, self.alphas[i_channel], self.means[i_channel], self.stds[i_channel])

    def norm_func(self, field, i_channel):
        return norm_func(field, self.alphas[i_channel], self.means[i_channel], self.stds[i_channel])

    collate_fn = staticmethod(collate)

    def train_dataloader(self):
        if self._train_loader is None:           # one loader object: its epoch counter advances from epoch to epoch
            self._train_loader = _Loader(self.data, self.train_ids, self.batch_size, shuffle=True, rank=self.rank,
                                         world=self.world, seed=self.seed)
        return self._train_loader

    def val_dataloader(self):
        return _Loader(self.data, self.valid_ids, self.batch_size, shuffle=False, rank=self.rank, world=self.world)

    def test_dataloader(self):
        return _Loader(self.data, self.test_ids, self.batch_size, shuffle=False)


class _RowSubset:
    """Rows ``rows`` of a memory-mapped (n, S, S, S) stack without copying it."""

    def __init__(self, base, rows):
        self.base, self.rows = base, np.asarray(rows)
        self.shape, self.ndim, self.dtype = (len(self.rows),) + tuple(base.shape[1:]), base.ndim, base.dtype

    def __len__(self):
        return len(self.rows)

    def __getitem__(self, i):
        return self.base[self.rows[i]]


def get_dataset(dataset_name="CMD", suite_name="Astrid", set_name="LH", z_name="z_0.0", channel_names=("Mcdm",),
                return_func=None, stage="fit", batch_size=1, cropsize=256, ndim=3, num_workers=8, mmap=True, **kw):
    """CAMELS_3D_dataset.get_dataset (CAMELS_3D_dataset.py:200-232); ``kw``: data_root, device, seed, rank, world."""
    selection = {"dataset_name": dataset_name, "suite_name": suite_name, "set_name": set_name, "z_name": z_name}
    if return_func is None:
        def return_func(fields, params):
            return {"x": torch.cat(fields, dim=0), "conditioning": None, "conditioning_values": params}
    return AstroDataModule(selection=selection, channel_names=list(channel_names), return_func=return_func, stage=stage,
                           batch_size=batch_size, do_crop=cropsize != REFERENCE_FULLSIZE, cropsize=cropsize, ndim=ndim,
                           num_workers=num_workers, mmap=mmap, **kw)
