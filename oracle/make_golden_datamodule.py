"""Generate tests/golden/datamodule_golden.npz by running the REFERENCE's own ``AstroDataModule``
(src/dataset/CAMELS_3D_dataset.py:76-199) on a miniature CAMELS-like directory.

Run in the build container only (needs /root/reference):   python oracle/make_golden_datamodule.py

The reference module needs three things that do not exist here, all supplied without touching its code:
  * ``lightning.pytorch.LightningDataModule``: a stand-in base class in ``sys.modules``;
  * three JSON tables opened by absolute path on the author's cluster at import time: ``open`` is redirected to the
    copies under /root/reference/src/dataset while the module is imported;
  * the grid / parameter files: a ``CMD_16`` entry pointing at the miniature files is added to its ``data_source``
    table, and ``np.loadtxt`` of the hard-coded parameter path is redirected to the miniature parameter file.
Only numerical outputs on seeded inputs are stored.
"""
from __future__ import annotations

import builtins
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "datamodule_golden.npz")
REF = "/root/reference"
N_SIMS, SIZE, CROP, CHANNELS = 20, 16, 8, ("Mstar", "Mcdm")


def write_camels_like(path: str, set_name: str = "CV", seed: int = 5) -> None:
    """The seeded miniature directory shared by the generator and the test."""
    rng = np.random.default_rng(seed)
    for c in CHANNELS:
        np.save(os.path.join(path, f"Grids_{c}_Astrid_{set_name}_{SIZE}_z=0.0.npy"),
                (10.0 ** (rng.standard_normal((N_SIMS, SIZE, SIZE, SIZE)) * 0.5 + 10.0)).astype(np.float32))
    np.savetxt(os.path.join(path, f"params_{set_name}_Astrid.txt"), rng.random((N_SIMS, 6)))


def return_func(fields, params):
    return {"conditioning": fields[0], "x": fields[1], "conditioning_values": [params]}


def load_reference_datamodule():
    lightning, lp = types.ModuleType("lightning"), types.ModuleType("lightning.pytorch")
    lp.LightningDataModule = type("LightningDataModule", (), {"__init__": lambda self: None})
    lightning.pytorch = lp
    sys.modules.update({"lightning": lightning, "lightning.pytorch": lp})
    real_open = builtins.open

    def redirected(file, *a, **kw):
        if isinstance(file, str) and file.startswith("/n/home12/cfpark00/Diffusion/vdm4cdm/src/dataset/"):
            file = os.path.join(REF, "src", "dataset", os.path.basename(file))
        return real_open(file, *a, **kw)

    sys.path.insert(0, REF)
    builtins.open = redirected
    try:
        import src.dataset.CAMELS_3D_dataset as mod
    finally:
        builtins.open = real_open
        sys.path.remove(REF)
    return mod


def main():
    mod = load_reference_datamodule()
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        write_camels_like(tmp, "CV")
        mod.data_source["CMD_16"] = {"Astrid": {"CV": {"z_0.0": {
            c: os.path.join(tmp, f"Grids_{c}_Astrid_CV_{SIZE}_z=0.0.npy") for c in CHANNELS}}}}
        real_np = mod.np
        proxy = types.SimpleNamespace(**{k: getattr(real_np, k) for k in dir(real_np) if not k.startswith("__")})
        proxy.loadtxt = lambda path, *a, **kw: real_np.loadtxt(os.path.join(tmp, os.path.basename(path)), *a, **kw)
        mod.np = proxy
        sel = {"dataset_name": "CMD_16", "suite_name": "Astrid", "set_name": "CV", "z_name": "z_0.0"}
        dm = mod.AstroDataModule(selection=sel, channel_names=list(CHANNELS), return_func=return_func, stage="test",
                                 batch_size=3, do_crop=True, cropsize=CROP, ndim=3, num_workers=0, mmap=False)
        out["test_len"] = np.array(len(dm.test_data))
        out["alphas"], out["means"], out["stds"] = np.array(dm.alphas, float), np.array(dm.means), np.array(dm.stds)
        batches = []
        for i, batch in enumerate(dm.test_dataloader()):
            batches.append(batch)
            if i == 3:
                break
        out["x"] = torch.cat([b["x"] for b in batches]).numpy()
        out["conditioning"] = torch.cat([b["conditioning"] for b in batches]).numpy()
        out["conditioning_values"] = torch.cat([b["conditioning_values"][0] for b in batches]).numpy()
        # one sample deep inside the set (simulation 11 of the 17 kept, crop 5)
        deep = dm.test_data[11 * 8 + 5]
        out["deep_x"], out["deep_params"] = deep["x"].numpy(), deep["conditioning_values"][0].numpy()
        torch.manual_seed(0)
        fit = mod.AstroDataModule(selection=sel, channel_names=list(CHANNELS), return_func=return_func, stage="fit",
                                  batch_size=3, do_crop=True, cropsize=CROP, ndim=3, num_workers=0, mmap=False)
        out["fit_train_len"], out["fit_valid_len"] = np.array(len(fit.train_data)), np.array(len(fit.valid_data))
        field = torch.from_numpy(out["x"][:1])
        out["unnorm_x0"] = dm.unnorm_func(field, 1).numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
