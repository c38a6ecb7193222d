"""CPU tests of the host-side logic: the ``mltools`` compat namespace, the model factory and registry, state_dict
compatibility with the oracle, the sampler's schedule algebra, and the data-parallel plumbing (flat buckets,
gradient all-reduce, realisation sharding) over a world_size-2 ``gloo`` group."""
import math
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mltools_namespace_resolves_to_this_package():
    import mltools.ml_utils as ml_utils
    import mltools.models.sfm_model as sfm_model
    import mltools.models.vdm_model as vdm_model
    import mltools.networks.networks as networks
    from mltools.utils import cuda_tools
    import vdm4cdm_b200.networks
    import vdm4cdm_b200.sfm_model
    import vdm4cdm_b200.vdm_model
    assert networks.CUNet is vdm4cdm_b200.networks.CUNet
    assert vdm_model.LightVDM is vdm4cdm_b200.vdm_model.LightVDM
    assert sfm_model.LightSFM is vdm4cdm_b200.sfm_model.LightSFM
    assert ml_utils.to_np(torch.arange(3.0, requires_grad=True)).tolist() == [0.0, 1.0, 2.0]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            cuda_tools.get_freer_device()


def test_registry_and_model_factory():
    from vdm4cdm_b200 import utils
    from vdm4cdm_b200.vdm_model import LightVDM
    configs = yaml.safe_load(open(os.path.join(ROOT, "configs.yaml")))
    ref_names = ["VDM_Go7_Mcdm_c_c_128", "VDM_Mstar_Mcdm_c_c_128", "VDM_Mstar_Mcdm_c_c_160", "VDM_Mstar_Mcdm_c_c_192",
                 "VDM_Mstar_Mcdm_c_c_224", "VDM_Mstar_Mcdm_c_c_256", "SFM_Mstar_Mcdm_c_c_128"]
    for n in ref_names:
        assert n in configs, n
    # the registry's checkpoint files are not shipped: a configured-but-missing ckpt_path raises like the reference's
    # torch.load (src/utils.py:468-469) unless the caller explicitly asks for random weights
    with pytest.raises(FileNotFoundError):
        utils.get_model(configs["VDM_Mstar_Mcdm_c_c_224"])
    with pytest.warns(UserWarning, match="RANDOM"):
        m = utils.get_model(configs["VDM_Mstar_Mcdm_c_c_224"], allow_random_init=True)
    assert isinstance(m, LightVDM) and m.model.score_model.shape == (1, 224, 224, 224)
    assert m.model.score_model.chs == [16, 32, 64, 128] and m.learning_rate == 3.0e-4 and m.model.gamma_max == 13.3
    assert utils.get_model(configs["SFM_Mstar_Mcdm_c_c_128"]) is None          # src/utils.py:472-473 does `pass`
    with pytest.raises(ValueError):
        utils.get_model({"type": "GAN"})
    m256 = utils.get_model({k: v for k, v in configs["VDM_Mstar_Mcdm_c_c_256"].items() if k != "ckpt_path"})                 # cropsize 256 -> circular padding (src/utils.py:460)
    assert m256.model.score_model.circular and m256.model.score_model.conv_in.padding_mode == "circular"
    assert not m.model.score_model.circular


def test_state_dict_is_interchangeable_with_the_oracle():
    from oracle.unet_ref import CUNet as RefNet
    from oracle.vdm_ref import LightVDM as RefLight
    from vdm4cdm_b200.networks import CUNet
    from vdm4cdm_b200.vdm_model import LightVDM
    kw = dict(shape=(1, 16, 16, 16), chs=(16, 32, 64), s_conditioning_channels=1, v_conditioning_dims=[6],
              t_conditioning=True)
    ref, mine = RefLight(RefNet(**kw)), LightVDM(CUNet(**kw))
    assert list(ref.state_dict().keys()) == list(mine.state_dict().keys())
    mine.load_state_dict(ref.state_dict(), strict=True)
    ref.load_state_dict(mine.state_dict(), strict=True)
    for (ka, a), (kb, b) in zip(ref.state_dict().items(), mine.state_dict().items()):
        assert ka == kb and torch.equal(a, b)


def test_step_coefficients_follow_the_reference_algebra():
    """z_s = alpha_s/alpha_t (z_t - c sigma_t eps) + sigma_s sqrt(c) noise (vdm_model.py:370-378) and its DDNM
    decomposition (src/utils.py:296-299)."""
    from vdm4cdm_b200.networks import CUNet
    from vdm4cdm_b200.vdm_model import VDM
    vdm = VDM(CUNet(shape=(1, 8, 8, 8), chs=(16, 32)))
    steps = torch.linspace(1.0, 0.0, 11)
    coef = vdm.step_coefficients(steps[:-1], steps[1:], final_rescale=True).double()
    g = vdm.gamma(steps).double()
    for i in range(10):
        gt, gs = g[i], g[i + 1]
        c = -math.expm1((gs - gt).item())
        a_t, a_s = math.sqrt(torch.sigmoid(-gt).item()), math.sqrt(torch.sigmoid(-gs).item())
        s_t, s_s = math.sqrt(torch.sigmoid(gt).item()), math.sqrt(torch.sigmoid(gs).item())
        assert coef[i, 0].item() == pytest.approx(a_s / a_t, rel=1e-6)
        assert coef[i, 1].item() == pytest.approx(-a_s / a_t * c * s_t, rel=1e-6)
        assert coef[i, 2].item() == pytest.approx(s_s * math.sqrt(c), rel=1e-6)
        assert coef[i, 3].item() == pytest.approx(1.0 / a_s if i == 9 else 1.0, rel=1e-6)
        # DDNM form: w_z z + w_x x0_hat + scale eps with x0_hat = (z - sigma_t eps)/alpha_t
        w_z, w_x = a_s * (1 - c) / a_t, a_s * c
        assert w_z + w_x / a_t == pytest.approx(a_s / a_t, rel=1e-9)
        assert -w_x * s_t / a_t == pytest.approx(-a_s / a_t * c * s_t, rel=1e-9)


def test_realisation_sharding_and_dgrad_chunks():
    from vdm4cdm_b200.autograd import _chunks
    from vdm4cdm_b200.trainer import shard_indices
    for world in (1, 2, 3, 8):
        got = sorted(i for r in range(world) for i in shard_indices(64, r, world))
        assert got == list(range(64))
    assert _chunks(32) == [(0, 32)] and _chunks(256) == [(0, 256)]
    assert _chunks(384) == [(0, 192), (192, 192)]
    for c in (16, 48, 96, 192, 320, 384, 512):
        parts = _chunks(c)
        assert sum(n for _, n in parts) == c and all(n <= 256 and n % 8 == 0 for _, n in parts)


def test_trainer_needs_cuda():
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU failure mode")
    from vdm4cdm_b200.trainer import Trainer
    m = torch.nn.Linear(4, 4)
    m.learning_rate = 1e-3
    with pytest.raises(RuntimeError, match="CUDA"):
        Trainer(m)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from vdm4cdm_b200.trainer import FlatBuckets, allreduce_gradients, broadcast_parameters, shard_indices
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                               # replicas start DIFFERENT on purpose
    model = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.SiLU(), torch.nn.Linear(5, 3))
    buckets = FlatBuckets(model)
    assert all(p.data_ptr() >= buckets.flat_param.data_ptr() for p in model.parameters())
    broadcast_parameters(buckets, src=0)
    # rank-dependent micro-batch -> rank-dependent gradients, written by autograd INTO the flat bucket
    buckets.zero_grad()
    x = torch.full((4, 7), float(rank + 1))
    model(x).sum().backward()
    local = buckets.flat_grad.clone()
    assert local.abs().sum() > 0, "autograd did not accumulate into the flat gradient views"
    allreduce_gradients(buckets)
    summed = buckets.flat_grad.clone()
    # the captured training step's form: small parameters' gradients stored by autograd, then moved into the bucket
    buckets.zero_grad(detach_small=True)
    model(x).sum().backward()
    buckets.collect_small_grads()
    assert torch.equal(buckets.flat_grad, local), "collect_small_grads does not reproduce the in-place accumulation"
    allreduce_gradients(buckets)
    assert torch.equal(buckets.flat_grad, summed)
    torch.save({"param": buckets.flat_param.clone(), "local": local, "summed": buckets.flat_grad.clone(),
                "mine": list(shard_indices(10, rank, world))}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_over_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_ddp_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"rank{i}.pt")) for i in range(world)]
    assert torch.equal(r[0]["param"], r[1]["param"]), "broadcast did not equalise the replicas"
    assert torch.allclose(r[0]["summed"], r[0]["local"] + r[1]["local"])
    assert torch.equal(r[0]["summed"], r[1]["summed"])
    assert sorted(r[0]["mine"] + r[1]["mine"]) == list(range(10))


def test_flat_bucket_small_gradient_collection_equals_in_place_accumulation():
    """FlatBuckets.zero_grad(detach_small=True) + collect_small_grads() (the form the captured training step uses: conv filters
    accumulate into their bucket views, everything of at most two dimensions is stored by autograd and copied in afterwards)
    leaves the same flat gradient as plain accumulation into the views, also for parameters that received no gradient."""
    from vdm4cdm_b200.trainer import FlatBuckets
    torch.manual_seed(3)
    model = torch.nn.Sequential(torch.nn.Conv3d(2, 4, 3, padding=1), torch.nn.GroupNorm(2, 4), torch.nn.SiLU(),
                                torch.nn.Conv3d(4, 1, 1), torch.nn.Flatten(), torch.nn.Linear(64, 3))
    unused = torch.nn.Parameter(torch.ones(5))               # registered, never reached by the loss
    model.register_parameter("unused", unused)
    buckets = FlatBuckets(model)
    x = torch.randn(2, 2, 4, 4, 4)
    buckets.zero_grad()
    model(x).square().sum().backward()
    ref = buckets.flat_grad.clone()
    assert ref.abs().sum() > 0
    for _ in range(2):                                       # twice: the views must be re-attached correctly for the next step
        buckets.zero_grad(detach_small=True)
        assert all((p.grad is None) == (p.dim() <= 2) for p in model.parameters())
        model(x).square().sum().backward()
        buckets.collect_small_grads()
        assert torch.equal(buckets.flat_grad, ref)
        for p, o in zip(buckets.params, buckets.offsets):
            assert p.grad.data_ptr() == buckets.flat_grad.data_ptr() + 4 * o
    o_unused = buckets.offsets[[id(p) for p in buckets.params].index(id(model.unused))]
    assert torch.equal(buckets.flat_grad[o_unused:o_unused + 5], torch.zeros(5))


def test_loader_state_resumes_the_same_sample_stream():
    """_Loader.state() / load_state(): a run resumed at (epoch, cursor) draws the same samples with the same augmentation
    draws as the uninterrupted run (checkpoint resume of scripts/train3D_c_c.py), on every rank."""
    from vdm4cdm_b200 import dataset

    class FakeData:
        def __init__(self, seed):
            self.seed, self.gen = seed, torch.Generator().manual_seed(seed)

        def draw(self, idx):
            return (idx, torch.randint(1000, (3,), generator=self.gen).tolist())

        def get_batch(self, chunk):
            return [self.draw(i) for i in chunk]

    def run(n_batches, state=None, rank=1):
        ld = dataset._Loader(FakeData(11 + 7919 * rank), list(range(23)), 2, shuffle=True, rank=rank, world=2, seed=11)
        if state is not None:
            ld.load_state(state)
        out = []
        while len(out) < n_batches:
            for b in ld:
                out.append(b)
                if len(out) == n_batches:
                    return out, ld.state()
        return out, ld.state()

    full, _ = run(15)
    for cut in (1, 5, 6, 7, 12):                       # 6 batches per epoch and rank: mid-epoch, at and across the boundary
        head, st = run(cut)
        tail, _ = run(15 - cut, state=st)
        assert head + tail == full, cut


def _write_camels_like(tmp_path, n_sims=20, size=16, set_name="CV", channels=("Mstar", "Mcdm"), res=16):
    import numpy as np
    rng = np.random.default_rng(0)
    for c in channels:
        np.save(tmp_path / f"Grids_{c}_Astrid_{set_name}_{res}_z=0.0.npy",
                (rng.random((n_sims, size, size, size)) * 1e10).astype(np.float32))
    np.savetxt(tmp_path / f"params_{set_name}_Astrid.txt", rng.random((n_sims, 6)))


def test_datamodule_file_lookup_cv_exclusion_and_split(tmp_path):
    """Host logic of AstroDataModule (CAMELS_3D_dataset.py:76-144) on a miniature CAMELS-like directory; boxes stay
    memory-mapped on the host (mmap=True), so no device is touched."""
    import numpy as np
    from vdm4cdm_b200 import dataset
    _write_camels_like(tmp_path)
    rf = lambda fields, params: {"conditioning": fields[0], "x": fields[1], "conditioning_values": [params]}
    sel = {"dataset_name": "CMD_16", "suite_name": "Astrid", "set_name": "CV", "z_name": "z_0.0"}
    assert dataset.grid_file_name("Mcdm", sel) == "Grids_Mcdm_Astrid_CV_16_z=0.0.npy"
    assert dataset.grid_file_name("Mcdm", dict(sel, dataset_name="CMD")) == "Grids_Mcdm_Astrid_CV_256_z=0.0.npy"
    dm = dataset.AstroDataModule(sel, ["Mstar", "Mcdm"], rf, stage="fit", batch_size=2, do_crop=True, cropsize=8,
                                 data_root=str(tmp_path), mmap=True)
    assert dm.data.n_sims == 17 and dm.data.ncrops == 8            # boxes 2, 8, 17 of the CV set are dropped
    raw = np.load(tmp_path / "Grids_Mcdm_Astrid_CV_16_z=0.0.npy")
    assert np.array_equal(dm.data.fields[1][2], raw[3]) and np.array_equal(dm.data.fields[1][15], raw[18])
    par = np.loadtxt(tmp_path / "params_CV_Astrid.txt")
    assert np.allclose(dm.data.params[7].numpy(), par[9], atol=1e-7)
    assert len(dm.train_ids) == int(17 * 8 * 0.95) and len(dm.train_ids) + len(dm.valid_ids) == 17 * 8
    assert not set(dm.train_ids) & set(dm.valid_ids)
    assert len(dm.train_dataloader()) == -(-len(dm.train_ids) // 2)
    assert dm.means[1] == pytest.approx(10.019186475678042) and dm.alphas == [1.0, 1.0]
    # two ranks see disjoint halves of an epoch
    a = dataset._Loader(dm.data, dm.train_ids, 2, shuffle=False, rank=0, world=2)
    b = dataset._Loader(dm.data, dm.train_ids, 2, shuffle=False, rank=1, world=2)
    assert not set(a.ids[0::2]) & set(b.ids[1::2])
    # ... and ALWAYS run the same number of full batches per epoch, whatever len(ids) % (world * batch) is (a rank
    # that runs out of data early leaves the others hanging in the gradient all-reduce): 950 ids, 4 ranks, batch 2
    for n_ids, world, bs in ((950, 4, 2), (129, 2, 2), (7, 8, 2), (16, 8, 2)):
        loaders = [dataset._Loader(dm.data, list(range(n_ids)), bs, shuffle=True, rank=r, world=world, seed=5)
                   for r in range(world)]
        assert len({len(l) for l in loaders}) == 1
        for epoch in (0, 1):
            orders = [l.epoch_order(epoch) for l in loaders]
            assert all(len(o) == len(loaders[0]) * bs for o in orders)
            seen = [i for o in orders for i in o]
            assert set(seen) == set(range(n_ids))                        # every sample is visited
            assert len(seen) - n_ids < world * bs                        # at most one padding unit of repeats
        assert loaders[0].epoch_order(0) != loaders[0].epoch_order(1)     # reshuffled per epoch, same on every rank
        assert dataset.epoch_permutation(n_ids, 5, 1) == dataset.epoch_permutation(n_ids, 5, 1)
    # test stage: no exclusion for non-CV sets, deterministic ids, missing files are reported by name
    _write_camels_like(tmp_path, n_sims=5, set_name="1P")
    dmt = dataset.get_dataset(dataset_name="CMD_16", set_name="1P", channel_names=["Mstar", "Mcdm"], return_func=rf,
                              stage="test", cropsize=16, data_root=str(tmp_path), mmap=True)
    assert dmt.data.n_sims == 5 and dmt.do_crop and dmt.data.ncrops == 1 and dmt.test_ids == list(range(5))
    assert dmt.data.draw(3)[1].tolist() == [0, 0, 0] and not dmt.data.augment
    with pytest.raises(FileNotFoundError, match="Grids_Go7_Astrid_1P_16"):
        dataset.get_dataset(dataset_name="CMD_16", set_name="1P", channel_names=["Go7"], stage="test",
                            data_root=str(tmp_path))
    with pytest.raises(FileNotFoundError, match="VDM4CDM_DATA_ROOT"):
        os.environ.pop("VDM4CDM_DATA_ROOT", None)
        dataset.get_dataset(dataset_name="CMD_16", set_name="1P", stage="test")


def test_ctypes_structs_match_the_c_header(tmp_path):
    """The ctypes mirrors in vdm4cdm_b200/_C.py must have the size and field offsets the C compiler gives the structs
    of include/vdm4cdm_b200.h (the header is plain C: compiled here with gcc, no CUDA needed)."""
    import ctypes
    import subprocess
    from vdm4cdm_b200 import _C
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = {"VdmConvDesc": _C.ConvDesc, "VdmConvEpilogue": _C.ConvEpilogue, "VdmWgradDesc": _C.WgradDesc,
             "VdmTensor": _C.Tensor, "VdmPackJob": _C.PackJob}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vdm4cdm_b200.h"', "int main(void) {"]
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, what, value = line.split()
        cls = pairs[cname]
        want = ctypes.sizeof(cls) if what == "size" else getattr(cls, what).offset
        assert int(value) == want, f"{cname}.{what}: C says {value}, ctypes says {want}"
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())


def test_get_model_builds_what_the_reference_get_model_builds():
    """tests/golden/get_model_kwargs.json holds the constructor arguments the REFERENCE's get_model (src/utils.py:434-471)
    passes for every entry of its configs.yaml (oracle/make_golden_get_model.py); our get_model on our copy of the
    registry must build the same networks."""
    import json
    from vdm4cdm_b200 import utils
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gold = json.load(open(os.path.join(root, "tests", "golden", "get_model_kwargs.json")))
    configs = yaml.safe_load(open(os.path.join(root, "configs.yaml")))
    assert set(gold) <= set(configs), set(gold) - set(configs)
    checked = 0
    for name, want in gold.items():
        cfg = {k: v for k, v in configs[name].items() if k != "ckpt_path"}
        model = utils.get_model(cfg)
        if want is None:
            assert model is None, name                       # SFM entries: get_model returns None (src/utils.py:472-473)
            continue
        net, w = model.model.score_model, want["CUNet"]
        assert list(net.shape) == w["shape"] and net.chs == w["chs"], name
        assert net.s_conditioning_channels == w["s_conditioning_channels"], name
        assert net.v_conditioning_dims == w["v_conditioning_dims"] and net.t_conditioning == w["t_conditioning"], name
        assert net.norm_groups == w["norm_groups"] and net.dropout_prob == w["dropout_prob"], name
        assert net.circular == (w["conv_padding_mode"] == "circular") and net.n_attention_heads == w["n_attention_heads"], name
        assert w["mid_attn"] is False
        assert model.model.gamma_max == want["LightVDM"]["gamma_max"], name
        assert model.learning_rate == want["LightVDM"]["learning_rate"] and model.draw_figure is None, name
        checked += 1
    assert checked >= 8


def test_get_datamodule_passes_what_the_reference_passes(monkeypatch):
    """tests/golden/get_datamodule_kwargs.json: the arguments the REFERENCE's get_datamodule (src/utils.py:401-432) hands
    to get_dataset for every registry entry and three data_params variants, and its return_func's batch schema
    (oracle/make_golden_get_datamodule.py).  Ours must hand the same to vdm4cdm_b200.dataset.get_dataset."""
    import json
    from vdm4cdm_b200 import dataset, utils
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gold = json.load(open(os.path.join(root, "tests", "golden", "get_datamodule_kwargs.json")))
    configs = yaml.safe_load(open(os.path.join(root, "configs.yaml")))
    calls = []
    monkeypatch.setattr(dataset, "get_dataset", lambda **kw: calls.append(kw) or "dm")
    variants = {"default": {}, "cv_fit": {"set_name": "CV", "stage": "fit", "batch_size": 4},
                "one_p": {"set_name": "1P", "stage": "test", "batch_size": 1}}
    for key, want in gold.items():
        name, variant = key.split("/")
        cfg = dict(configs[name], data_params=dict(configs[name]["data_params"], **variants[variant]))
        calls.clear()
        assert utils.get_datamodule(cfg, data_root="/somewhere") == "dm"
        kw = dict(calls[0])
        rf = kw.pop("return_func")
        assert kw.pop("data_root") == "/somewhere"            # the one argument the reference does not have
        assert kw == want["kwargs"], (key, kw, want["kwargs"])
        assert rf(fields=["F0", "F1"], params="P") == want["return_func"], key
    with pytest.raises(AssertionError, match="data_params"):
        utils.get_datamodule({"cropsize": 128})


def test_train_script_presets_match_the_reference_scripts():
    """tests/golden/train_presets.json: hyper-parameters read from the syntax trees of the reference's
    train{VDM,SFM}3D*_c_c_..._lowbatch.py (oracle/make_golden_train_presets.py); scripts/train3D_c_c.py must carry them."""
    import importlib.util
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "scripts"))
    spec = importlib.util.spec_from_file_location("train3D_c_c", os.path.join(root, "scripts", "train3D_c_c.py"))
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.path.remove(os.path.join(root, "scripts"))
    gold = json.load(open(os.path.join(root, "tests", "golden", "train_presets.json")))
    assert len(gold) == 9
    for name, want in gold.items():
        # scripts without a grid size in their name take it from the command line: any size without its own script
        cropsize = want["cropsize_in_name"] or 64
        chs, batch, dataset_name = mod.preset(want["model"], cropsize)
        assert chs == want["chs"] and batch == want["batch_size"] and dataset_name == want["dataset_name"], name
        assert mod.NORM_GROUPS == want["norm_groups"] and mod.DROPOUT_PROB == want["dropout_prob"], name
        assert mod.LEARNING_RATE == want["learning_rate"] and mod.GRADIENT_CLIP_VAL == want["gradient_clip_val"], name
        assert want["conditioning_values"] == 6 and want["conditioning_channels"] == 1, name      # CUNet(... [6], 1 ...)
        assert want["set_name"] == "LH" and want["stage"] == "fit" and want["mmap"] is False, name
        if want["model"] == "VDM":
            assert mod.GAMMA_MAX == want["gamma_max"], name
        # Trainer(max_steps=1_000_000, ModelCheckpoint(every_n_train_steps=10_000)): the script's defaults
        assert want["max_steps"] == 1_000_000 and want["every_n_train_steps"] == 10_000, name
        # the thin launcher that carries the reference's file name pins the script's own literals (a reference script
        # takes chs / batch_size / dataset_name / val_check_interval from itself, not from the cropsize argument)
        assert mod.script_preset(want["model"], want["cropsize_in_name"]) == \
            (want["chs"], want["batch_size"], want["dataset_name"], want["val_check_interval"]), name
        launcher = open(os.path.join(root, "scripts", name)).read()
        assert f'model_kind="{want["model"]}", script_grid={want["cropsize_in_name"]!r}' in launcher, name
    src = open(os.path.join(root, "scripts", "train3D_c_c.py")).read()
    assert '"--max-steps", type=int, default=1_000_000' in src and '"--ckpt-every", type=int, default=10_000' in src
    assert mod.preset("VDM", 256) == mod.BASE_PRESET and mod.preset("SFM", 224) == mod.BASE_PRESET


def _run_our_generate(monkeypatch, tmp_path, script_mode, model_name, runtype, conditioning_values):
    """scripts/generate_3D.py:main(mode) on the CPU with the same recording stand-ins oracle/make_golden_generate.py puts
    under the reference's scripts."""
    import importlib
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "scripts"))
    import _common
    from vdm4cdm_b200 import utils
    gen = importlib.import_module("generate_3D")
    draws, dms = [], []

    class FakeModel:
        def eval(self):
            return self

        def draw_samples(self, **kw):
            field = int(kw["s_conditioning"].flatten()[0])
            draws.append({"batch_size": kw["batch_size"], "field": field, "n_v_conditionings": len(kw["v_conditionings"]),
                          "ids": list(kw["realisation_ids"])})
            return torch.full((kw["batch_size"], 1, 2, 2, 2), float(field * 1000 + len(draws)))

    class FakeDM:
        def test_dataloader(self):
            for i in range(40):
                yield {"x": torch.full((1, 1, 2, 2, 2), float(i)), "conditioning": torch.full((1, 1, 2, 2, 2), float(i)),
                       "conditioning_values": [torch.full((1, 6), float(i))]}

    def fake_get_datamodule(config, **kw):
        dms.append({k: config["data_params"].get(k) for k in ("set_name", "stage", "batch_size")})
        return FakeDM()

    monkeypatch.setattr(gen, "init_distributed", lambda: (0, 1, torch.device("cpu")))
    monkeypatch.setattr(utils, "get_model", lambda config, device=None, allow_random_init=False: FakeModel())
    monkeypatch.setattr(utils, "get_datamodule", fake_get_datamodule)
    cfg = {model_name: {"type": "VDM", "cropsize": 2, "conditioning_values": conditioning_values,
                        "in_field_name": "Mstar", "out_field_name": "Mcdm", "data_params": {"dataset_name": "CMD"}}}
    (tmp_path / "configs.yaml").write_text(yaml.safe_dump(cfg))
    out = tmp_path / "out"
    monkeypatch.setattr(sys, "argv", ["generate", model_name, str(out), runtype, "--configs", str(tmp_path / "configs.yaml"),
                                      "--data-root", "unused", "--batch", "5"])
    gen.main(script_mode)
    files = {}
    for f in sorted(os.listdir(out)):
        a = np.load(out / f)
        first = a.reshape(a.shape[0], -1)[:, 0].astype(int) // 1000
        files[f] = {"shape": list(a.shape), "field": int(first[0]), "all_from_one_field": bool((first == first[0]).all())}
    return files, dms, draws


@pytest.mark.parametrize("key", ["generate_3D.py:VDM_Mstar_Mcdm_c_c_128:CV_12_12", "generate_3D.py:VDM_Mstar_Mcdm_c_c_128:CV_1_128",
                                 "generate_3D.py:VDM_Mstar_Mcdm_c_uc_256:CV_1_128", "generate_3D_1P.py:VDM_Mstar_Mcdm_c_c_128:1P_24",
                                 "generate_3D_1P.py:VDM_Mstar_Mcdm_c_c_256:1P_128"])
def test_generate_scripts_behave_like_the_reference_scripts(monkeypatch, tmp_path, key):
    """tests/golden/generate_scripts.json records what the reference's generate_3D.py / generate_3D_1P.py DO when run
    (unmodified, under recording stand-ins: oracle/make_golden_generate.py): which test batches they sample, how many
    realisations, the data_params they set, the files they write.  Ours must do the same (in batches of 5 here)."""
    import json
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    want = json.load(open(os.path.join(root, "tests", "golden", "generate_scripts.json")))[key]
    script, model_name, runtype = key.split(":")
    ref_calls = [json.loads(c) for c in want["distinct_draw_calls"]]
    n_v = ref_calls[0]["n_v_conditionings"]
    files, dms, draws = _run_our_generate(monkeypatch, tmp_path, "1P" if "1P" in script else "CV", model_name, runtype,
                                          conditioning_values=6 if n_v else 0)
    assert files == want["files"]
    assert dms == want["data_params"]
    assert sum(d["batch_size"] for d in draws) == want["n_draw_calls"]          # reference: batch_size=1 per call
    assert {d["n_v_conditionings"] for d in draws} == {n_v}
    assert {d["field"] for d in draws} == {c["field"] for c in ref_calls}
    rep = next(iter(want["files"].values()))["shape"][0]
    by_field = {}
    for d in draws:
        by_field.setdefault(d["field"], []).extend(d["ids"])
    assert all(sorted(ids) == list(range(rep)) for ids in by_field.values())     # every realisation exactly once
