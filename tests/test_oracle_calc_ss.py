"""The oracles of the calc_SS statistics against the REFERENCE's own functions (calc_SS.py:51-75), executed from the
reference tree by oracle/make_golden_calc_ss.py -> tests/golden/calc_ss_golden.npz.  (The CUDA path is tied to these
oracles by tests/test_gpu_pk.py.)"""
import os

import numpy as np
import torch

from oracle import logpdf_ref, power_ref
from oracle.make_golden_calc_ss import mass_field

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "calc_ss_golden.npz"))


def test_logpdf_oracle_reproduces_the_reference_histograms():
    fields = torch.from_numpy(mass_field(21, (3, 1, 32, 32, 32)))
    half = fields[:, :, :16].sum(2)
    assert np.array_equal(logpdf_ref.get_logpdf_3d(fields), GOLD["logpdf3d"])
    assert np.array_equal(logpdf_ref.get_logpdf_2d(half), GOLD["logpdf2d"])
    assert GOLD["logpdf3d"].sum() > 0.99 * fields.numel()          # the fixture exercises the populated range


def test_pk_oracle_reproduces_get_pk_3d_and_get_pk_2d():
    fields = mass_field(21, (3, 1, 32, 32, 32))
    fu = fields / fields.sum((2, 3, 4), keepdims=True)              # calc_SS.py:68
    _, p3, _ = power_ref.pk(fu)
    assert np.allclose(p3, GOLD["pk3d"], rtol=1e-5)
    half = fields[:, :, :16].sum(2)
    hu = half / half.sum((2, 3), keepdims=True)                     # calc_SS.py:73
    _, p2, _ = power_ref.pk(hu)
    assert np.allclose(p2, GOLD["pk2d"], rtol=1e-5)
