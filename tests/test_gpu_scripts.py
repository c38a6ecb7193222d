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
    from vdm4cdm_b200.dataset import NORMALIZATIONS_3D
    mu, sd = NORMALIZATIONS_3D["Mcdm"]            # calc_SS.py:146: dm.unnorm_func(samples, i_channel=1)
    fid = 10.0 ** (np.load(ss / "fid_3.npy").astype(np.float64) * sd + mu) - 1.0
    assert summ["pk3d"].shape == (3, 8) and summ["pk2d_half"].shape == (3, 8) and summ["logpdf3d"].shape == (3, 99)
    assert np.allclose(summ["post_means"], fid.mean(0, keepdims=True), rtol=1e-3)
    assert np.allclose(summ["post_stds"], fid.std(0, ddof=1, keepdims=True), rtol=1e-3, atol=1e-3 * fid.std())
    assert np.allclose(summ["mean3d"], fid.reshape(3, -1).mean(1), rtol=1e-3)
    from oracle import power_ref
    field = np.load(ss / "fid_3.npy")[:1].astype(np.float32)
    un = (10.0 ** (field * np.float32(sd) + np.float32(mu)) - 1.0).astype(np.float32)
    k_ref, p_ref, _ = power_ref.power(un / un.sum())
    assert np.allclose(summ["pk3d"][0], p_ref, rtol=2e-3), (summ["pk3d"][0], p_ref)
    # same ensemble whatever the batch size
    out1 = tmp_path / "out1"
    _run("generate_3D.py", "VDM_Mstar_Mcdm_c_c_16", out1, "CV_1_128", *common[:-1], 1)
    gen1 = np.load(out1 / "gen_0.npy")
    err = np.linalg.norm(gen1 - gen) / np.linalg.norm(gen)
    assert err < 1e-2, err                                        # bf16 tolerance (see test_philox_sampling_is_batch_independent)


def _losses(log):
    return {int(line.split(":")[0].split()[1]): float(line.split("loss")[1].split()[0])
            for line in log.splitlines() if line.startswith("step") and " loss " in line}


def test_train_script_on_a_camels_like_directory(tmp_path):
    """The real-data path (AstroDataModule over files) through the launcher that carries the reference's file name:
    checkpoints are written (trainVDM3D128_...:47 ModelCheckpoint), validation numbers land in the JSONL log
    (:42 val_check_interval, :97-103 P(k) / r(k) of sample vs truth), and ``--resume`` continues the run
    (trainVDM3D_c_c_...:133-135)."""
    import json
    import torch
    _camels_dir(tmp_path, {"LH": 6})
    common = ["Mstar", "Mcdm", 16, "--data-root", tmp_path, "--dataset-name", "CMD_16", "--chs", 16, 32, "--log-every", 1,
              "--val-check-interval", 4, "--val-sampling-steps", 5]
    ck = tmp_path / "ckpt"
    log = _run("trainVDM3D128_c_c_from_field_name_thick_lowbatch.py", *common, "--max-steps", 8, "--ckpt-every", 4,
               "--ckpt-dir", ck)
    losses = _losses(log)
    assert sorted(losses) == list(range(1, 9)) and all(np.isfinite(list(losses.values())))
    names = sorted(os.listdir(ck))
    assert "VDM_Mstar_Mcdm_c_c_16_step=4.ckpt" in names and "VDM_Mstar_Mcdm_c_c_16_step=8.ckpt" in names, names
    state = torch.load(ck / "VDM_Mstar_Mcdm_c_c_16_step=4.ckpt", map_location="cpu", weights_only=False)
    assert state["global_step"] == 4 and state["optimizer"]["step"] == 4 and "model.score_model.conv_in.weight" in state["state_dict"]
    assert state["optimizer"]["exp_avg"].abs().sum() > 0
    # the checkpoint is what utils.get_model loads (src/utils.py:467-469)
    from vdm4cdm_b200 import utils
    model = utils.get_model({"type": "VDM", "cropsize": 16, "chs": [16, 32], "ckpt_path": str(ck / "VDM_Mstar_Mcdm_c_c_16_step=8.ckpt")})
    final = torch.load(ck / "VDM_Mstar_Mcdm_c_c_16_step=8.ckpt", map_location="cpu", weights_only=False)["state_dict"]
    assert torch.equal(model.state_dict()["model.score_model.conv_in.weight"], final["model.score_model.conv_in.weight"])
    assert not torch.equal(final["model.score_model.conv_in.weight"], state["state_dict"]["model.score_model.conv_in.weight"])
    # validation records: loss over the validation loader + P(k) / r(k) of a 5-step sample vs the truth
    recs = [json.loads(l) for l in open(ck / "metrics.jsonl")]
    vals = [r for r in recs if "val_loss" in r]
    assert [r["step"] for r in vals] == [4, 8] and all(np.isfinite(r["val_loss"]) and r["val_batches"] >= 1 for r in vals)
    # (an UNTRAINED network's 5-step samples are amplified by alpha_0/alpha_1 ~ 770 and overflow 10**x: r(k) may be NaN
    #  here exactly as it would be for the reference; the truth spectrum is always finite)
    assert len(vals[0]["pk_truth"]) == 8 and len(vals[0]["cc"]) == 8 and np.isfinite(vals[0]["pk_truth"]).all()
    assert all(np.isnan(c) or abs(c) <= 1.0 + 1e-4 for c in vals[0]["cc"]) and "sample_std" in vals[0]
    assert [r["step"] for r in recs if "train_loss" in r] == list(range(1, 9))
    # resume from step 4: steps 5..8 reproduce the uninterrupted run (same data order, RNG, optimizer state; the
    # tolerance covers the unordered fp32 atomics of the weight-gradient and statistics reductions)
    ck2 = tmp_path / "ckpt2"
    log2 = _run("trainVDM3D128_c_c_from_field_name_thick_lowbatch.py", *common, "--max-steps", 8, "--ckpt-every", 100,
                "--ckpt-dir", ck2, "--resume", ck / "VDM_Mstar_Mcdm_c_c_16_step=4.ckpt", "--val-check-interval", 1000)
    again = _losses(log2)
    assert sorted(again) == [5, 6, 7, 8], log2
    assert os.path.exists(ck2 / "VDM_Mstar_Mcdm_c_c_16_step=8.ckpt")        # final checkpoint at loop exit
    for k in (5, 6):
        assert abs(again[k] - losses[k]) < 2e-2 * abs(losses[k]), (k, again[k], losses[k])


def test_sfm_launcher_synthetic_boxes_writes_checkpoints(tmp_path):
    """--synthetic-boxes (the data-stream path that skipped the checkpoint block in round 1) with the SFM launcher."""
    ck = tmp_path / "ck"
    log = _run("trainSFM3D128_c_c_from_field_name_thick_lowbatch.py", "Mstar", "Mcdm", 16, "--synthetic", "--synthetic-boxes",
               "--chs", 16, 32, "--batch-size", 2, "--max-steps", 5, "--ckpt-every", 2, "--ckpt-dir", ck, "--log-every", 1,
               "--val-check-interval", 3)
    assert len(_losses(log)) == 5
    assert sorted(os.listdir(ck)) == ["SFM_Mstar_Mcdm_c_c_16_step=2.ckpt", "SFM_Mstar_Mcdm_c_c_16_step=4.ckpt",
                                      "SFM_Mstar_Mcdm_c_c_16_step=5.ckpt", "metrics.jsonl"]
