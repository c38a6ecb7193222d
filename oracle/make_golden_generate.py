"""Generate tests/golden/generate_scripts.json by RUNNING the reference's ``generate_3D.py`` and ``generate_3D_1P.py``
(unmodified, with runpy) against recording stand-ins for everything outside the scripts: ``src.utils.get_model`` returns
a fake model whose ``draw_samples`` stamps each call, ``src.utils.get_datamodule`` a fake data module whose test loader
yields numbered batches, ``mltools.utils.cuda_tools.get_freer_device`` returns "cpu".  What is recorded is the scripts'
own behaviour: which loader batches are sampled, how many realisations each, which keyword arguments reach
``draw_samples``, which ``data_params`` they set, and the names / shapes of the files they write.

Run in the build container only (needs /root/reference):   python oracle/make_golden_generate.py
"""
from __future__ import annotations

import json
import os
import runpy
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "generate_scripts.json")
REF = "/root/reference"
N_BATCHES = 40


class FakeModel:
    def __init__(self, log):
        self.log = log

    def to(self, device):
        return self

    def eval(self):
        return self

    def draw_samples(self, **kw):
        rec = {"batch_size": kw.get("batch_size"), "field": int(kw["s_conditioning"].flatten()[0]),
               "n_v_conditionings": len(kw["v_conditionings"]),
               "extra_kwargs": sorted(k for k in kw if k not in ("batch_size", "s_conditioning", "v_conditionings"))}
        self.log.append(rec)
        return torch.full((kw["batch_size"], 1, 2, 2, 2), float(rec["field"] * 1000 + len(self.log)))


class FakeDM:
    def __init__(self, log, config):
        log.append({k: config["data_params"].get(k) for k in ("set_name", "stage", "batch_size")})

    def test_dataloader(self):
        for i in range(N_BATCHES):
            yield {"x": torch.full((1, 1, 2, 2, 2), float(i)), "conditioning": torch.full((1, 1, 2, 2, 2), float(i)),
                   "conditioning_values": [torch.full((1, 6), float(i))]}


def run(script, model_name, runtype):
    draws, dms = [], []
    utils = types.ModuleType("src.utils")
    utils.get_model = lambda config: FakeModel(draws)
    utils.get_datamodule = lambda config: FakeDM(dms, config)
    src = types.ModuleType("src")
    src.utils = utils
    cuda_tools = types.ModuleType("mltools.utils.cuda_tools")
    cuda_tools.get_freer_device = lambda: "cpu"
    mu, ml = types.ModuleType("mltools.utils"), types.ModuleType("mltools")
    mu.cuda_tools, ml.utils = cuda_tools, mu
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    stubs = {"src": src, "src.utils": utils, "mltools": ml, "mltools.utils": mu, "mltools.utils.cuda_tools": cuda_tools,
             "matplotlib": mpl, "matplotlib.pyplot": plt}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    cwd, argv = os.getcwd(), sys.argv
    with tempfile.TemporaryDirectory() as tmp:
        try:
            os.chdir(REF)                                  # the scripts read ./configs.yaml
            sys.argv = [script, model_name, tmp, runtype]
            runpy.run_path(os.path.join(REF, script), run_name="__main__")
        finally:
            os.chdir(cwd)
            sys.argv = argv
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v
        files = {}
        for f in sorted(os.listdir(tmp)):
            a = np.load(os.path.join(tmp, f))
            files[f] = {"shape": list(a.shape), "field": int(a.flat[0]) // 1000,
                        "all_from_one_field": bool((a.reshape(a.shape[0], -1)[:, 0].astype(int) // 1000 == int(a.flat[0]) // 1000).all())}
    per_call = {json.dumps(d, sort_keys=True) for d in draws}
    return {"files": files, "data_params": dms, "n_draw_calls": len(draws), "distinct_draw_calls": sorted(per_call)}


def main():
    out = {}
    for script, model, runtype in (("generate_3D.py", "VDM_Mstar_Mcdm_c_c_128", "CV_12_12"),
                                   ("generate_3D.py", "VDM_Mstar_Mcdm_c_c_128", "CV_1_128"),
                                   ("generate_3D.py", "VDM_Mstar_Mcdm_c_uc_256", "CV_1_128"),
                                   ("generate_3D_1P.py", "VDM_Mstar_Mcdm_c_c_128", "1P_24"),
                                   ("generate_3D_1P.py", "VDM_Mstar_Mcdm_c_c_256", "1P_128")):
        out[f"{script}:{model}:{runtype}"] = run(script, model, runtype)
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, {n: (d["shape"], d["field"]) for n, d in v["files"].items()}, v["data_params"], v["n_draw_calls"])


if __name__ == "__main__":
    main()
