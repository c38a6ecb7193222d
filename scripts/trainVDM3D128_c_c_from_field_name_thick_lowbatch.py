#!/usr/bin/env python
"""``python trainVDM3D128_c_c_from_field_name_thick_lowbatch.py field_in field_out cropsize``: the reference's entry point of this name (trainVDM3D128_c_c_from_field_name_thick_lowbatch.py:51-133 there:
its own chs / batch_size / dataset_name / val_check_interval literals, pinned in tests/golden/train_presets.json) on
``scripts/train3D_c_c.py`` (checkpoints, validation, ``--resume``, ``--synthetic``, torchrun data parallelism)."""
import train3D_c_c

if __name__ == "__main__":
    train3D_c_c.main(model_kind="VDM", script_grid=128)
