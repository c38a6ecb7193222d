"""Generate tests/golden/ddnm_golden.npz by running the REFERENCE's own ``get_ddnm_result`` (src/utils.py:277-304).

Run in the build container only (needs /root/reference):   python oracle/make_golden_ddnm.py

The reference function is model agnostic: it drives ``vdm.model.sample_zt_given_zs`` / ``sample_zs_given_zt(...,
return_ddnm=True)`` of whatever VDM it is given and draws its noise from torch's global generator.  Here it drives the
ORACLE's VDM (oracle/vdm_ref.py) on a tiny seeded network, so the fixture pins the restated DDNM loop
(oracle/vdm_ref.py:get_ddnm_result) -- call order, time-travel indices, the (w_z, w_x, x0, scale) update, the
consumption order of the noise draws -- against the reference's own code.  Nothing from the reference is copied:
only its numerical output on seeded inputs is stored.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.make_golden import load_reference_utils  # noqa: E402
from oracle.unet_ref import CUNet  # noqa: E402
from oracle.vdm_ref import LightVDM  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "ddnm_golden.npz")
SHAPE, CHS, STEPS, TRAVEL = (1, 8, 8, 8), (16, 32), 5, [0, 1, 2, 1, 3]


def build(seed: int = 11):
    """The seeded model / measurement shared by the generator and the test."""
    torch.manual_seed(seed)
    net = CUNet(shape=SHAPE, chs=CHS, s_conditioning_channels=1, v_conditioning_dims=[6], t_conditioning=True,
                norm_groups=8, dropout_prob=0.1).eval()
    vdm = LightVDM(score_model=net).eval()
    g = torch.Generator().manual_seed(seed + 1)
    cond = torch.randn((2,) + SHAPE, generator=g)
    params = torch.rand((2, 6), generator=g)
    truth = torch.randn((2,) + SHAPE, generator=g)
    mask = (torch.rand(SHAPE, generator=g) > 0.5).float()

    def A(x):                      # inpainting measurement: keep the unmasked voxels
        return x * mask

    return vdm, cond, params, A(truth), A, A       # A is a projection: A^T = A


class _RefFacing(torch.nn.Module):
    """What the reference's loop touches: ``vdm.device`` and ``vdm.model.{score_model, sample_*}``; it passes
    ``conditioning=None`` (src/utils.py:296), a keyword the mltools VDM accepts and this oracle's UNet does not."""

    def __init__(self, light):
        super().__init__()
        self.light = light
        self.device = torch.device("cpu")
        outer = light.model

        class _Model:
            score_model = outer.score_model

            @staticmethod
            def sample_zt_given_zs(**kw):
                return outer.sample_zt_given_zs(**kw)

            @staticmethod
            def sample_zs_given_zt(**kw):
                kw.pop("conditioning", None)
                return outer.sample_zs_given_zt(**kw)

        self.model = _Model()


def main():
    ref = load_reference_utils()
    vdm, cond, params, y, A, AT = build()
    torch.manual_seed(1234)
    with torch.no_grad():
        out = ref.get_ddnm_result(_RefFacing(vdm), y, A, AT, n_sampling_steps=STEPS, l=list(TRAVEL), return_all=True,
                                  s_conditioning=cond, v_conditionings=[params])
    np.savez_compressed(OUT, out=out.numpy(), torch_version=np.array(torch.__version__))
    print("wrote", OUT, tuple(out.shape), float(out.abs().mean()))


if __name__ == "__main__":
    main()
