"""Generate tests/golden/get_model_kwargs.json: the constructor arguments the REFERENCE's ``get_model``
(src/utils.py:434-471) passes to ``mltools.networks.networks.CUNet`` and ``mltools.models.vdm_model.LightVDM`` for every
entry of the reference's ``configs.yaml``.  ``mltools`` is absent, so recording stand-ins are injected into
``sys.modules``; what is pinned is the registry -> constructor mapping (defaults, padding mode, conditioning
dimensions), which is all ``get_model`` does besides loading a checkpoint.

Run in the build container only (needs /root/reference):   python oracle/make_golden_get_model.py
"""
from __future__ import annotations

import json
import os
import sys
import types

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.make_golden import REF, load_reference_utils  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "get_model_kwargs.json")


class _Recorder:
    def __init__(self, **kw):
        self.kw = kw


def main():
    ref = load_reference_utils()
    networks, vdm_model = types.ModuleType("mltools.networks.networks"), types.ModuleType("mltools.models.vdm_model")
    networks.CUNet, vdm_model.LightVDM = _Recorder, _Recorder
    pkg_n, pkg_m = types.ModuleType("mltools.networks"), types.ModuleType("mltools.models")
    pkg_n.networks, pkg_m.vdm_model = networks, vdm_model
    sys.modules.update({"mltools.networks": pkg_n, "mltools.networks.networks": networks, "mltools.models": pkg_m,
                        "mltools.models.vdm_model": vdm_model})
    configs = yaml.safe_load(open(os.path.join(REF, "configs.yaml")))
    out = {}
    for name, cfg in configs.items():
        cfg = {k: v for k, v in cfg.items() if k != "ckpt_path"}          # checkpoints live on the author's cluster
        model = ref.get_model(cfg)
        if model is None:
            out[name] = None                                              # SFM entries: the reference returns None
            continue
        light = dict(model.kw)
        net = light.pop("score_model").kw
        net["shape"] = list(net["shape"])
        out[name] = {"CUNet": net, "LightVDM": light}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT, len(out), "entries")


if __name__ == "__main__":
    main()
