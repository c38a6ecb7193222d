"""The oracle's DDNM loop (oracle/vdm_ref.py:get_ddnm_result) against the REFERENCE's own ``get_ddnm_result``
(src/utils.py:277-304): tests/golden/ddnm_golden.npz holds the output of the reference function driving the oracle's
VDM on a seeded toy problem (generator: oracle/make_golden_ddnm.py, run where /root/reference exists).  Both draw
their noise from torch's global generator in the same order, so the comparison is to fp32 rounding."""
import os

import numpy as np
import torch

from oracle.make_golden_ddnm import STEPS, TRAVEL, build
from oracle.vdm_ref import get_ddnm_result

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ddnm_golden.npz")


def test_ddnm_restatement_matches_the_reference_loop():
    gold = np.load(GOLD)["out"]
    vdm, cond, params, y, A, AT = build()
    torch.manual_seed(1234)
    out = get_ddnm_result(vdm, y, A, AT, n_sampling_steps=STEPS, l=list(TRAVEL), return_all=True,
                          s_conditioning=cond, v_conditionings=[params]).numpy()
    assert out.shape == gold.shape == (STEPS, 2, 1, 8, 8, 8)
    err = np.abs(out - gold).max() / np.abs(gold).max()
    assert err < 1e-5, err
    # a different time-travel schedule is a different result (the fixture pins the schedule handling too)
    torch.manual_seed(1234)
    other = get_ddnm_result(vdm, y, A, AT, n_sampling_steps=STEPS, l=0, return_all=True,
                            s_conditioning=cond, v_conditionings=[params]).numpy()
    assert np.abs(other - gold).max() / np.abs(gold).max() > 1e-3
    # the measured voxels of every iterate equal the measurement (x0_r = A^T y + (I - A^T A) x0)
    assert np.allclose(A(torch.from_numpy(out[-1])).numpy(), y.numpy(), atol=1e-4 * np.abs(gold).max())
