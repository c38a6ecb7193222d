"""Host mirror of the reference's ``src/utils.py`` for the hot path: ``power`` / ``pk`` / ``get_ccs``
(src/utils.py:16-128) and the model factory ``get_model`` (src/utils.py:434-475), same names, argument
meaning and return types, computed by the CUDA library (cuFFT R2C + the k-shell binning kernel).

Differences a user can observe, all deliberate:
  * inputs must live on a CUDA device (there is no CPU fallback; a CPU tensor raises),
  * bins are accumulated in fp64 and deterministically per block (the reference uses fp32
    ``torch.bincount``, whose CUDA path is an unordered atomicAdd), then returned in the reference's dtypes
    (fp32 k, fp32 P, int32 N),
  * ``pk`` transforms all samples with one batched plan and ``get_ccs`` transforms every field once
    (the reference runs four FFTs for two fields).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import ops


def _as_fields(x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("vdm4cdm_b200.utils: P(k) runs on the GPU only (tensor is on %s)" % x.device)
    return x.contiguous().float()


def power(x: torch.Tensor, x2: Optional[torch.Tensor] = None):
    """(k, P, N) of a (B, C, *spatial) field: batch mean of the channel sum of X conj(X2), binned by
    ceil(|k|) with Hermitian weights, bins 1..min(n)//2 (src/utils.py:16-83)."""
    f1 = _as_fields(x)[None]
    f2 = None if x2 is None else _as_fields(x2)[None]
    k, p, n = ops.pk_fields(f1, f2)
    return k[0].float(), p[0].float(), n[0].to(torch.int32)


def pk(fields: torch.Tensor, fields2: Optional[torch.Tensor] = None):
    """Per-sample spectra of (B, C, *spatial) fields, stacked on axis 0 (src/utils.py:85-102)."""
    f1 = _as_fields(fields)[:, None]
    f2 = None if fields2 is None else _as_fields(fields2)[:, None]
    k, p, n = ops.pk_fields(f1, f2)
    return k.float(), p.float(), n.to(torch.int32)


def pk_conversion(dim=2, boxsize=25):
    """Unit conversion used by the reference's plots (src/utils.py:104-108)."""
    assert dim == 2, "check code before 3d!"
    import numpy as np
    return 2 * np.pi / boxsize, boxsize ** 2


def get_ccs(fields1: torch.Tensor, fields2: torch.Tensor, full: bool = False):
    """Cross-correlation coefficient r(k) = P12 / sqrt(P11 P22) per sample, or for all pairs with
    ``full=True`` (src/utils.py:110-128)."""
    a, b = _as_fields(fields1), _as_fields(fields2)
    if not full:
        assert len(a) == len(b)
        k, p11, p22, p12, _ = ops.pk_cross3(a[:, None], b[:, None])
        return k.float(), (p12 / torch.sqrt(p11 * p22)).float()
    k, p11, _ = ops.pk_fields(a[:, None])
    p22 = ops.pk_fields(b[:, None])[1]
    rows = []
    for i in range(len(a)):
        rep = a[i][None].expand(len(b), *a.shape[1:]).contiguous()
        rows.append(ops.pk_fields(rep[:, None], b[:, None])[1])
    p12 = torch.stack(rows, dim=0)
    return k.float(), (p12 / torch.sqrt(p11[:, None] * p22[None, :])).float()


def get_model(config: dict, device=None):
    """``CUNet`` + ``LightVDM`` from a ``configs.yaml`` entry (src/utils.py:434-471): ``chs`` default
    [32, 64, 128, 256], ``norm_groups=8``, ``dropout_prob=0.1``, ``gamma_max=13.3``, circular padding iff
    ``cropsize == 256``; SFM entries return ``None`` exactly like the reference (src/utils.py:472-473)."""
    from .networks import CUNet
    from .vdm_model import LightVDM
    if config["type"] == "SFM":
        return None
    if config["type"] != "VDM":
        raise ValueError(f"Unknown model type {config['type']}")
    cropsize = int(config.get("cropsize", 128))
    chs = list(config.get("chs", [32, 64, 128, 256]))
    n_cond_values = int(config.get("conditioning_values", 6))
    s_channels = int(config.get("conditioning_channels", 1))
    net = CUNet(shape=(1, cropsize, cropsize, cropsize), chs=chs, s_conditioning_channels=s_channels,
                v_conditioning_dims=[n_cond_values] if n_cond_values else [], t_conditioning=True, norm_groups=8,
                mid_attn=False, dropout_prob=0.1, conv_padding_mode="circular" if cropsize == 256 else "zeros",
                n_attention_heads=4)
    model = LightVDM(score_model=net, draw_figure=None, gamma_max=13.3, learning_rate=3.0e-4)
    ckpt_path = config.get("ckpt_path")
    if ckpt_path and os.path.exists(ckpt_path):   # the registry's paths live on the author's cluster
        state = torch.load(ckpt_path, map_location="cpu")
        model.load_state_dict(state["state_dict"])
    if device is not None:
        model = model.to(device)
    return model
