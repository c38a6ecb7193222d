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


def test_datamodule_batches_from_files_match_oracle(tmp_path):
    """utils.get_datamodule (src/utils.py:401-432) over a miniature CAMELS-like directory: test-stage batches are the
    log-normalised boxes themselves (HBM-resident and host-staged boxes give the same batch); fit-stage batches
    replay through the oracle's crop/flip/permutate."""
    from oracle import augment_ref
    from vdm4cdm_b200 import dataset, utils
    rng = np.random.default_rng(1)
    n_sims, S = 6, 16
    raws = {}
    for c in ("Mstar", "Mcdm"):
        raws[c] = (rng.random((n_sims, S, S, S)) * 1e10).astype(np.float32)
        np.save(tmp_path / f"Grids_{c}_Astrid_1P_16_z=0.0.npy", raws[c])
    par = rng.random((n_sims, 6))
    np.savetxt(tmp_path / "params_1P_Astrid.txt", par)
    config = {"cropsize": 16, "in_field_name": "Mstar", "out_field_name": "Mcdm",
              "data_params": {"dataset_name": "CMD_16", "set_name": "1P", "batch_size": 2}}
    dm = utils.get_datamodule(config, data_root=str(tmp_path))
    batches = list(dm.test_dataloader())
    assert len(batches) == 3 and batches[0]["x"].is_cuda and batches[0]["x"].shape == (2, 1, S, S, S)
    m, s = dataset.NORMALIZATIONS_3D["Mcdm"]
    want = (np.log10(raws["Mcdm"].astype(np.float64) + 1.0) - m) / s
    got = torch.cat([b["x"] for b in batches])[:, 0].cpu().numpy()
    assert np.allclose(got, want, rtol=2e-6, atol=2e-6)
    assert np.allclose(torch.cat([b["conditioning_values"][0] for b in batches]).cpu().numpy(), par, atol=1e-6)
    assert torch.allclose(dm.unnorm_func(batches[0]["x"], 1)[:, 0].cpu(), torch.from_numpy(raws["Mcdm"][:2]), rtol=2e-4)
    # boxes left on the host (mmap=True) give bit-identical batches
    rf = lambda fields, params: {"conditioning": fields[0], "x": fields[1], "conditioning_values": [params]}
    dmh = dataset.get_dataset(dataset_name="CMD_16", set_name="1P", channel_names=["Mstar", "Mcdm"], return_func=rf,
                              stage="test", batch_size=2, cropsize=16, data_root=str(tmp_path), mmap=True)
    assert torch.equal(next(iter(dmh.test_dataloader()))["conditioning"], batches[0]["conditioning"])
    # fit stage: random crops + flips + permutations, replayed through the oracle from the same seed
    kw = dict(dataset_name="CMD_16", set_name="1P", channel_names=["Mstar", "Mcdm"], return_func=rf, stage="fit",
              batch_size=4, cropsize=8, data_root=str(tmp_path), mmap=False, seed=9)
    dmf, dmr = dataset.get_dataset(**kw), dataset.get_dataset(**kw)
    loader = dmf.train_dataloader()
    assert len(dmf.train_ids) == int(n_sims * 8 * 0.95)
    batch = next(iter(loader))
    # epoch order: a function of (seed, epoch) shared by all ranks, separate from the augmentation generator
    order = [dmr.train_ids[i] for i in dataset.epoch_permutation(len(dmr.train_ids), 9, 0)][:4]
    assert order == loader.epoch_order(0)[:4]
    for j, idx in enumerate(order):
        bidx, anchor, flip, perm = dmr.data.draw(idx)
        want = augment_ref.prepare(raws["Mcdm"][bidx][None], anchor, (8, 8, 8), flip, perm, 1.0, m, s)
        assert np.allclose(batch["x"][j].cpu().numpy(), want, rtol=2e-6, atol=2e-6)
