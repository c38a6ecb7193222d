"""Generates tests/golden/augment_golden.npz from the reference's own augmentation classes
(/root/reference/src/dataset/augmentation.py).  Run in the build container only (the reference checkout does not
travel to the GPU box): python oracle/make_golden_augment.py"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_augmentation", "/root/reference/src/dataset/augmentation.py")
aug = importlib.util.module_from_spec(spec)
spec.loader.exec_module(aug)

out = {}
rng = np.random.default_rng(11)
cases = [(12, 8, 0.5, 1.2, 0.7), (16, 8, 1.0, 10.019, 0.552), (10, 10, 1.0, 3.0, 2.0)]
for ci, (S, c, alpha, mean, std) in enumerate(cases):
    raw = [np.abs(rng.standard_normal((1, S, S, S))).astype(np.float32) * 10 for _ in range(2)]
    torch.manual_seed(100 + ci)
    crop = aug.Crop(3, c, 0, fullsize=S, do_augshift=True)
    icrop = int(rng.integers(crop.ncrops))
    fields = crop([f.copy() for f in raw], icrop)
    anchor = crop.anchors[icrop].copy()              # includes the random shift Crop added in place
    fields = [torch.from_numpy(f).to(torch.float32) for f in fields]
    fields = aug.LogTransform([alpha, alpha])(fields)
    fields = aug.Normalize(means=[mean, mean], stds=[std, std])(fields)
    fl = aug.Flip(ndim=3)
    fields = fl(fields)
    pm = aug.Permutate(ndim=3)
    fields = pm(fields)
    flip_mask = np.zeros(3, dtype=np.int32)
    flip_mask[fl.axes.numpy()] = 1
    out[f"c{ci}_raw0"], out[f"c{ci}_raw1"] = raw
    out[f"c{ci}_out0"], out[f"c{ci}_out1"] = [f.contiguous().numpy() for f in fields]
    out[f"c{ci}_anchor"] = np.asarray(anchor, dtype=np.int32)
    out[f"c{ci}_flip"] = flip_mask
    out[f"c{ci}_perm"] = pm.axes.numpy().astype(np.int32)
    out[f"c{ci}_meta"] = np.array([S, c, alpha, mean, std], dtype=np.float64)
# order of the crop grid (AstroDataset.__getitem__: bidx, icrop = divmod(idx, ncrops); CAMELS_3D_dataset.py:53-56),
# also for a box the crop size does not divide
for S, c in ((16, 8), (24, 8), (24, 10)):
    out[f"anchors_{S}_{c}"] = np.asarray(aug.Crop(3, c, 0, fullsize=S, do_augshift=False).anchors, dtype=np.int32)
np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "augment_golden.npz"), **out)
print("wrote", len(out), "arrays")
