"""Host logic of vdm4cdm_b200.dataset.AstroDataModule against the REFERENCE's own AstroDataModule
(src/dataset/CAMELS_3D_dataset.py:76-199): tests/golden/datamodule_golden.npz holds the batches the reference class
produces on a seeded miniature CAMELS directory (generator: oracle/make_golden_datamodule.py).  Here the same
directory is rebuilt from the seed, OUR data module decides which simulation / crop / constants every sample uses
(boxes memory-mapped on the host, no device involved), and the numpy oracle (oracle/augment_ref.py) turns that into
values -- the CUDA gather kernel is tied to the same oracle by tests/test_gpu_dataset.py."""
import os

import numpy as np

from oracle import augment_ref
from oracle.make_golden_datamodule import CHANNELS, CROP, N_SIMS, return_func, write_camels_like
from vdm4cdm_b200 import dataset

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "datamodule_golden.npz"))


def _sample(dm, idx):
    bidx, anchor, flip, perm = dm.data.draw(idx)
    fields = [augment_ref.prepare(np.asarray(f[bidx])[None], anchor, (CROP,) * 3, flip, perm, dm.alphas[k], dm.means[k], dm.stds[k])
              for k, f in enumerate(dm.data.fields)]
    return fields, dm.data.params[bidx].numpy()


def test_test_stage_batches_match_the_reference_datamodule(tmp_path):
    write_camels_like(str(tmp_path), "CV")
    sel = {"dataset_name": "CMD_16", "suite_name": "Astrid", "set_name": "CV", "z_name": "z_0.0"}
    dm = dataset.AstroDataModule(sel, list(CHANNELS), return_func, stage="test", batch_size=3, do_crop=True, cropsize=CROP,
                                 data_root=str(tmp_path), mmap=True)
    assert len(dm.test_ids) == int(GOLD["test_len"]) == (N_SIMS - 3) * 8        # three CV boxes dropped, 8 crops each
    assert np.allclose(dm.alphas, GOLD["alphas"]) and np.allclose(dm.means, GOLD["means"], rtol=0, atol=0)
    assert np.allclose(dm.stds, GOLD["stds"], rtol=0, atol=0)
    assert len(dm.test_dataloader()) == -(-len(dm.test_ids) // 3)
    for idx in range(12):                                                       # the first four batches of three
        (cond, x), params = _sample(dm, dm.test_ids[idx])
        assert np.allclose(x, GOLD["x"][idx], rtol=2e-6, atol=2e-6), idx
        assert np.allclose(cond, GOLD["conditioning"][idx], rtol=2e-6, atol=2e-6), idx
        assert np.allclose(params, GOLD["conditioning_values"][idx], atol=1e-7), idx
    (_, x), params = _sample(dm, 11 * 8 + 5)                                    # deep inside: past all dropped boxes
    assert np.allclose(x, GOLD["deep_x"], rtol=2e-6, atol=2e-6) and np.allclose(params, GOLD["deep_params"], atol=1e-7)
    import torch
    un = dm.unnorm_func(torch.from_numpy(GOLD["x"][:1]), 1).numpy()
    assert np.allclose(un, GOLD["unnorm_x0"], rtol=1e-6)


def test_fit_stage_split_sizes_match_the_reference_datamodule(tmp_path):
    write_camels_like(str(tmp_path), "CV")
    sel = {"dataset_name": "CMD_16", "suite_name": "Astrid", "set_name": "CV", "z_name": "z_0.0"}
    dm = dataset.AstroDataModule(sel, list(CHANNELS), return_func, stage="fit", batch_size=3, do_crop=True, cropsize=CROP,
                                 data_root=str(tmp_path), mmap=True)
    assert len(dm.train_ids) == int(GOLD["fit_train_len"]) and len(dm.valid_ids) == int(GOLD["fit_valid_len"])
    assert dm.data.augment and dm.data.aug_shift                                # fit: random shift, flips, permutations
