"""Generate tests/golden/train_presets.json: the hyper-parameters of the reference's 3-D conditional training scripts
(``train{VDM,SFM}3D{,128,160,192,224}_c_c_from_field_name_thick_lowbatch.py``), read from their syntax trees: the
literal assignments ``chs``, ``batch_size``, ``norm_groups``, ``dropout_prob``, ``gamma_max``, the ``dataset_name`` /
``learning_rate`` / ``gradient_clip_val`` keyword literals.  ``scripts/train3D_c_c.py`` carries these as its PRESETS
table and constants; tests/test_host_logic.py compares.

Run in the build container only (needs /root/reference):   python oracle/make_golden_train_presets.py
"""
from __future__ import annotations

import ast
import glob
import json
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "train_presets.json")
NAMES = ("chs", "batch_size", "norm_groups", "dropout_prob", "gamma_max", "conditioning_values", "conditioning_channels")
KEYWORDS = ("dataset_name", "learning_rate", "gradient_clip_val", "set_name", "stage", "mmap", "max_steps",
            "val_check_interval", "every_n_train_steps", "devices")


def main():
    out = {}
    for path in sorted(glob.glob("/root/reference/train*3D*_c_c_from_field_name_thick_lowbatch.py")):
        name = os.path.basename(path)
        m = re.match(r"train(VDM|SFM)3D(\d*)_c_c", name)
        tree = ast.parse(open(path).read(), filename=path)
        rec = {"model": m.group(1), "cropsize_in_name": int(m.group(2)) if m.group(2) else None}
        for node in ast.walk(tree):
            if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) \
                    and node.targets[0].id in NAMES:
                try:
                    rec[node.targets[0].id] = ast.literal_eval(node.value)
                except ValueError:
                    pass
            if isinstance(node, ast.keyword) and node.arg in KEYWORDS:
                try:
                    rec[node.arg] = ast.literal_eval(node.value)
                except ValueError:
                    pass
        out[name] = rec
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, v)


if __name__ == "__main__":
    main()
