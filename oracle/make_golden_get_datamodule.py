"""Generate tests/golden/get_datamodule_kwargs.json: the arguments the REFERENCE's ``get_datamodule``
(src/utils.py:401-432) passes to ``CAMELS_3D_dataset.get_dataset`` for every registry entry that has ``data_params``,
and the batch schema of its ``return_func``.  ``src.dataset.CAMELS_3D_dataset`` needs Lightning and the author's cluster
paths at import time, so a recording stand-in is injected into ``sys.modules``.

Run in the build container only (needs /root/reference):   python oracle/make_golden_get_datamodule.py
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

OUT = os.path.join(HERE, "..", "tests", "golden", "get_datamodule_kwargs.json")


def main():
    ref = load_reference_utils()
    calls = []
    stub = types.ModuleType("src.dataset.CAMELS_3D_dataset")
    stub.get_dataset = lambda **kw: calls.append(kw) or "dm"
    pkg, sub = types.ModuleType("src"), types.ModuleType("src.dataset")
    pkg.dataset, sub.CAMELS_3D_dataset = sub, stub
    sys.modules.update({"src": pkg, "src.dataset": sub, "src.dataset.CAMELS_3D_dataset": stub})
    configs = yaml.safe_load(open(os.path.join(REF, "configs.yaml")))
    out = {}
    for name, cfg in configs.items():
        if "data_params" not in cfg:
            continue
        for variant, extra in (("default", {}), ("cv_fit", {"set_name": "CV", "stage": "fit", "batch_size": 4}),
                               ("one_p", {"set_name": "1P", "stage": "test", "batch_size": 1})):
            c = dict(cfg, data_params=dict(cfg["data_params"], **extra))
            calls.clear()
            assert ref.get_datamodule(c) == "dm"
            kw = dict(calls[0])
            rf = kw.pop("return_func")
            sample = rf(fields=["F0", "F1"], params="P")
            out[f"{name}/{variant}"] = {"kwargs": kw, "return_func": sample}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT, len(out), "entries")


if __name__ == "__main__":
    main()
