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
    # same ensemble whatever the batch size
    out1 = tmp_path / "out1"
    _run("generate_3D.py", "VDM_Mstar_Mcdm_c_c_16", out1, "CV_1_128", *common[:-1], 1)
    gen1 = np.load(out1 / "gen_0.npy")
    err = np.linalg.norm(gen1 - gen) / np.linalg.norm(gen)
    assert err < 1e-2, err                                        # bf16 tolerance (see test_philox_sampling_is_batch_independent)


def test_train_script_on_a_camels_like_directory(tmp_path):
    _camels_dir(tmp_path, {"LH": 6})
    log = _run("train3D_c_c.py", "Mstar", "Mcdm", 16, "--model", "VDM", "--data-root", tmp_path, "--chs", 16, 32,
               "--max-steps", 6, "--log-every", 2, "--ckpt-dir", tmp_path / "ckpt")
    losses = [float(line.split("loss")[1].split()[0]) for line in log.splitlines() if line.startswith("step")]
    assert len(losses) == 3 and all(np.isfinite(losses))
