"""GPU parity of vdm_augment_crop / DeviceAstroDataset against the reference's own augmentation classes (golden
fixtures) and the numpy oracle: the data movement is bit-exact, the log-normalisation agrees to fp32 rounding."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "augment_golden.npz"))


def test_augment_crop_matches_reference_golden():
    from vdm4cdm_b200 import ops
    for ci in range(3):
        S, c, alpha, mean, std = GOLD[f"c{ci}_meta"]
        c = int(c)
        for f in (0, 1):
            raw = torch.from_numpy(GOLD[f"c{ci}_raw{f}"][0]).cuda()
            got = ops.augment_crop(raw, (c, c, c), GOLD[f"c{ci}_anchor"], GOLD[f"c{ci}_flip"], GOLD[f"c{ci}_perm"],
                                   alpha=float(alpha), mean=float(mean), std=float(std), do_log=True).cpu().numpy()
            want = GOLD[f"c{ci}_out{f}"][0]
            assert np.allclose(got, want, rtol=2e-6, atol=2e-6), np.abs(got - want).max()
            # pure data movement (no log): bit-exact against the oracle
            from oracle import augment_ref
            moved = ops.augment_crop(raw, (c, c, c), GOLD[f"c{ci}_anchor"], GOLD[f"c{ci}_flip"], GOLD[f"c{ci}_perm"]).cpu().numpy()
            ref = augment_ref.permutate(augment_ref.flip(augment_ref.crop_periodic(GOLD[f"c{ci}_raw{f}"], GOLD[f"c{ci}_anchor"],
                                                                                  (c, c, c)), GOLD[f"c{ci}_flip"]), GOLD[f"c{ci}_perm"])
            assert np.array_equal(moved, ref[0])


def test_anisotropic_crop_negative_anchor_and_bad_arguments():
    from oracle import augment_ref
    from vdm4cdm_b200 import ops
    rng = np.random.default_rng(3)
    raw = rng.standard_normal((1, 7, 9, 11)).astype(np.float32)
    got = ops.augment_crop(torch.from_numpy(raw[0]).cuda(), (4, 6, 5), (-3, 8, 20), (0, 1, 1), (1, 2, 0)).cpu().numpy()
    ref = augment_ref.permutate(augment_ref.flip(augment_ref.crop_periodic(raw, (-3, 8, 20), (4, 6, 5)), (0, 1, 1)), (1, 2, 0))
    assert got.shape == (6, 5, 4) and np.array_equal(got, ref[0])
    with pytest.raises(RuntimeError, match="permutation"):
        ops.augment_crop(torch.from_numpy(raw[0]).cuda(), (4, 4, 4), (0, 0, 0), (0, 0, 0), (0, 0, 1))


def test_device_dataset_batch_schema_and_statistics():
    from oracle import augment_ref
    from vdm4cdm_b200.dataset import DeviceAstroDataset, norm_func, unnorm_func
    g = torch.Generator().manual_seed(0)
    S, n = 32, 16
    raw_c = torch.rand((3, S, S, S), generator=g) * 1e10 + 1.0
    raw_x = torch.rand((3, S, S, S), generator=g) * 1e12 + 1.0
    params = torch.rand((3, 6), generator=g)

    def return_func(fields, params):                  # trainVDM3D128_c_c_from_field_name_thick_lowbatch.py:75-76
        return {"conditioning": fields[0], "x": fields[1], "conditioning_values": [params]}

    ds = DeviceAstroDataset([raw_c.cuda(), raw_x.cuda()], params, return_func, alphas=[1.0, 1.0], means=[9.0, 11.0],
                            stds=[0.5, 0.6], crop=n, seed=5)
    assert len(ds) == 3 * 8
    batch = ds.get_batch([0, 9, 23])
    assert batch["x"].shape == (3, 1, n, n, n) and batch["conditioning"].shape == (3, 1, n, n, n)
    assert isinstance(batch["conditioning_values"], list) and batch["conditioning_values"][0].shape == (3, 6)
    # replay the same host draws and compare with the oracle pipeline sample by sample
    ds2 = DeviceAstroDataset([raw_c.cuda(), raw_x.cuda()], params, return_func, alphas=[1.0, 1.0], means=[9.0, 11.0],
                             stds=[0.5, 0.6], crop=n, seed=5)
    for j, idx in enumerate([0, 9, 23]):
        bidx, anchor, flip, perm = ds2.draw(idx)
        want = augment_ref.prepare(raw_x[bidx][None].numpy(), anchor, (n, n, n), flip, perm, 1.0, 11.0, 0.6)
        assert np.allclose(batch["x"][j].cpu().numpy(), want, rtol=2e-6, atol=2e-6)
        assert torch.equal(batch["conditioning_values"][0][j].cpu(), params[bidx])
    x = batch["x"]
    assert torch.allclose(norm_func(unnorm_func(x, 1.0, 11.0, 0.6), 1.0, 11.0, 0.6), x, atol=1e-4)
