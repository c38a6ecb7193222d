"""The entry-point scripts run end to end on a miniature CAMELS-like directory (files named as the reference's
data_source_3d.json names them): generate_3D.py / generate_3D_1P.py (generate_3D.py:43-97, generate_3D_1P.py:43-70),
train3D_c_c.py (trainVDM3D128_...:75-89,134-160) and calc_SS.py on the generated ensembles."""
import os
import subprocess
import sys

import numpy as np
import pytest
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args):
    env = dict(os.environ, PYTHONPATH=ROOT)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script), *map(str, args)], env=env, cwd=ROOT,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    return res.stdout


def _camels_dir(path, sets, n=16):
    rng = np.random.default_rng(0)
    for set_name, n_sims in sets.items():
        for c in ("Mstar", "Mcdm"):
            np.save(path / f"Grids_{c}_Astrid_{set_name}_{n}_z=0.0.npy",
                    (10.0 ** (rng.standard_normal((n_sims, n, n, n)) * 0.5 + 10.0)).astype(np.float32))
        np.savetxt(path / f"params_{set_name}_Astrid.txt", rng.random((n_sims, 6)))


def test_generate_scripts_on_a_camels_like_directory(tmp_path):
    _camels_dir(tmp_path, {"CV": 8, "1P": 29})
    cfg = {"VDM_Mstar_Mcdm_c_c_16": {"type": "VDM", "in_field_name": "Mstar", "out_field_name": "Mcdm", "cropsize": 16,
                                     "chs": [16, 32], "data_params": {"dataset_name": "CMD_16"}}}
    # a checkpoint in the reference's format ({"state_dict": ...}, src/utils.py:467) so that every run has the same weights
    import torch
    from vdm4cdm_b200 import utils
    torch.manual_seed(0)
    torch.save({"state_dict": utils.get_model(cfg["VDM_Mstar_Mcdm_c_c_16"]).state_dict()}, tmp_path / "model.ckpt")
    cfg["VDM_Mstar_Mcdm_c_c_16"]["ckpt_path"] = str(tmp_path / "model.ckpt")
    (tmp_path / "configs.yaml").write_text(yaml.safe_dump(cfg))
    common = ["--configs", tmp_path / "configs.yaml", "--data-root", tmp_path, "--n-sampling-steps", 3, "--rep", 3, "--batch", 2]
    out = tmp_path / "out"
    _run("generate_3D.py", "VDM_Mstar_Mcdm_c_c_16", out, "CV_1_128", *common)
    gen = np.load(out / "gen_0.npy")
    assert gen.shape == (3, 1, 16, 16, 16) and np.isfinite(gen).all() and gen.std() > 0
    assert not np.array_equal(gen[0], gen[1])                     # realisations differ (Philox key = (seed, r))
    _run("generate_3D_1P.py", "VDM_Mstar_Mcdm_c_c_16", out, "1P_24", *common, "--fields", 2)
    assert sorted(os.listdir(out)) == ["Om_m2_3.npy", "fid_3.npy", "gen_0.npy"]
    # calc_SS.py over ensembles of both naming schemes; P(k) and posterior statistics against the oracle / numpy
    # (an untrained network's samples are far outside the data range, so the ensemble here is synthetic)
    ss = tmp_path / "ss"
    os.makedirs(ss)
    np.save(ss / "fid_3.npy", np.random.default_rng(3).standard_normal((3, 1, 16, 16, 16)).astype(np.float32))
    np.save(ss / "gen_0.npy", np.random.default_rng(4).standard_normal((2, 1, 16, 16, 16)).astype(np.float32))
    _run("calc_SS.py", ss, "--chunk", 2)
    assert os.path.exists(ss / "gen_0_summary.npz")
    summ = np.load(ss / "fid_3_summary.npz")
    fid = 10.0 ** (np.load(ss / "fid_3.npy").astype(np.float64) * 0.552 + 10.019) - 1.0
    assert summ["pk3d"].shape == (3, 8) and summ["pk2d_half"].shape == (3, 8) and summ["logpdf3d"].shape == (3, 99)
    assert np.allclose(summ["post_means"], fid.mean(0, keepdims=True), rtol=1e-3)
    assert np.allclose(summ["post_stds"], fid.std(0, ddof=1, keepdims=True), rtol=1e-3, atol=1e-3 * fid.std())
    assert np.allclose(summ["mean3d"], fid.reshape(3, -1).mean(1), rtol=1e-3)
    from oracle import power_ref
    field = np.load(ss / "fid_3.npy")[:1].astype(np.float32)
    un = (10.0 ** (field * np.float32(0.552) + np.float32(10.019)) - 1.0).astype(np.float32)
    k_ref, p_ref, _ = power_ref.power(un / un.sum())
    assert np.allclose(summ["pk3d"][0], p_ref, rtol=2e-3), (summ["pk3d"][0], p_ref)
    # same ensemble whatever the batch size
    out1 = tmp_path / "out1"
    _run("generate_3D.py", "VDM_Mstar_Mcdm_c_c_16", out1, "CV_1_128", *common[:-1], 1)
    gen1 = np.load(out1 / "gen_0.npy")
    err = np.linalg.norm(gen1 - gen) / np.linalg.norm(gen)
    assert err < 1e-2, err                                        # bf16 tolerance (see test_philox_sampling_is_batch_independent)


def test_train_script_on_a_camels_like_directory(tmp_path):
    _camels_dir(tmp_path, {"LH": 6})
    log = _run("train3D_c_c.py", "Mstar", "Mcdm", 16, "--model", "VDM", "--data-root", tmp_path, "--dataset-name", "CMD_16",
               "--chs", 16, 32,
               "--max-steps", 6, "--log-every", 2, "--ckpt-dir", tmp_path / "ckpt")
    losses = [float(line.split("loss")[1].split()[0]) for line in log.splitlines() if line.startswith("step")]
    assert len(losses) == 3 and all(np.isfinite(losses))
