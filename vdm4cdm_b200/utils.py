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


def get_logpdf(fields: torch.Tensor, lo: float, hi: float, nbins: int = 99) -> torch.Tensor:
    """Histogram of log10(field + 1) per sample, ``np.histogram(..., bins=np.linspace(lo, hi, nbins + 1))`` as in
    calc_SS.py:51-65 (3-D: lo, hi = 8.5, 15; 2-D slabs: 10.5, 15.5; 99 bins).  fields: CUDA fp32 (B, C, ...) -> int64 (B, nbins)."""
    f = _as_fields(fields)
    return ops.log_histogram(f.reshape(f.shape[0], -1), float(lo), float(hi), int(nbins), 1.0)


def get_logpdf_3d(fields: torch.Tensor) -> torch.Tensor:
    return get_logpdf(fields, 8.5, 15.0, 99)


def get_logpdf_2d(fields: torch.Tensor) -> torch.Tensor:
    return get_logpdf(fields, 10.5, 15.5, 99)


def get_ddnm_result(vdm, y, A, AT, n_sampling_steps=250, l=10, return_all=False, verbose=0, noise_fn=None, seed=0,
                    **kwargs):
    """DDNM inpainting / linear-inverse sampler with time travel (src/utils.py:277-304), same signature and loop:

        z ~ N(0, I);  for i in range(n):  L = min(l[i], i)
            z = sample_zt_given_zs(z, t=steps[i-L], s=steps[i])                      # re-noise L steps back
            for j in L..0:  (w_z, w_x, x0_hat, scale) = sample_zs_given_zt(z, t=steps[i-j], s=steps[i+1-j], return_ddnm=True)
                            x0_r = A^T y + x0_hat - A^T A x0_hat                      # range-null space correction
                            z = w_z z + w_x x0_r + scale N(0, I)

    The denoiser call and the three-term update run on the CUDA kernels (``vdm_conv3d`` trunk, ``vdm_sampler_step``
    with coefficients (w_z, w_x, scale)); ``A`` / ``AT`` are the caller's torch callables.  Noise comes from the
    counter-based Philox stream (draw index = position in the loop) or from ``noise_fn(draw, shape)``.
    """
    import numpy as np
    if not isinstance(l, np.ndarray):
        if isinstance(l, int):
            l = np.full(n_sampling_steps, l)
        elif isinstance(l, list):
            l = np.array(l)
    assert np.all(l >= 0), "l must be non-negative"
    assert np.issubdtype(l.dtype, np.integer), "l must be integer"
    assert isinstance(l, np.ndarray) and l.ndim == 1 and len(l) == n_sampling_steps, \
        "l must be 1d array of length n_sampling_steps or a single integer>0 or a list of integers>0"
    model = vdm.model
    dev = vdm.device
    steps = torch.linspace(1.0, 0.0, n_sampling_steps + 1, device=dev)
    shape = (y.shape[0], *model.score_model.shape)
    draw = [0]

    def noise(shape_):
        d = draw[0]
        draw[0] += 1
        if noise_fn is not None:
            return noise_fn(d, shape_).to(dev).float().contiguous()
        return ops.philox_normal(shape_, seed, d, None, dev)

    z = noise(shape)
    ATy = AT(y)
    xs = []
    it = range(n_sampling_steps)
    if verbose >= 1:
        from tqdm import trange
        it = trange(n_sampling_steps, desc="sampling")
    x_0t_r = None
    with torch.no_grad():
        for i in it:
            L = int(min(l[i], i))
            z = model.sample_zt_given_zs(zs=z, t=steps[i - L], s=steps[i], noise=noise(shape))
            for j in range(L, -1, -1):
                w_z, w_x_0t, x_0t, scale = model.sample_zs_given_zt(zt=z, t=steps[i - j], s=steps[i + 1 - j],
                                                                    return_ddnm=True, **kwargs)
                x_0t_r = ATy + x_0t - AT(A(x_0t))
                coef = torch.stack([w_z.reshape(()), w_x_0t.reshape(()), scale.reshape(()),
                                    torch.ones((), device=dev)]).float().reshape(1, 4).contiguous()
                z = ops.sampler_step(z.contiguous().float(), x_0t_r.contiguous().float(), coef, noise=noise(shape))
            if return_all:
                xs.append(x_0t_r)
    if return_all:
        return torch.stack(xs, dim=0)
    return x_0t_r


def get_datamodule(config: dict, **kw):
    """``AstroDataModule`` of a ``configs.yaml`` entry (src/utils.py:401-432): channels ``[in_field_name,
    out_field_name]``, batch schema {"conditioning", "x", "conditioning_values": [params]}, defaults suite Astrid /
    set CV / z_0.0 / stage test / batch 1, boxes resident in HBM (the reference passes ``mmap=False``).
    ``kw`` (data_root, device, seed, rank, world) goes to ``vdm4cdm_b200.dataset.get_dataset``."""
    from . import dataset
    if "data_params" not in config:
        assert False, "data_params not in config"
    data_params = config["data_params"]

    def return_func(fields, params):
        return {"conditioning": fields[0], "x": fields[1], "conditioning_values": [params]}

    return dataset.get_dataset(dataset_name=data_params["dataset_name"], suite_name=data_params.get("suite_name", "Astrid"),
                               return_func=return_func, set_name=data_params.get("set_name", "CV"),
                               z_name=data_params.get("z_name", "z_0.0"),
                               channel_names=[config["in_field_name"], config["out_field_name"]],
                               stage=data_params.get("stage", "test"), batch_size=data_params.get("batch_size", 1),
                               cropsize=config["cropsize"], num_workers=8, mmap=False, **kw)


def get_model(config: dict, device=None, allow_random_init: bool = False):
    """``CUNet`` + ``LightVDM`` from a ``configs.yaml`` entry (src/utils.py:434-471): ``chs`` default
    [32, 64, 128, 256], ``norm_groups=8``, ``dropout_prob=0.1``, ``gamma_max=13.3``, circular padding iff
    ``cropsize == 256``; SFM entries return ``None`` exactly like the reference (src/utils.py:472-473).

    A configured ``ckpt_path`` is loaded, and a missing file raises like the reference's ``torch.load`` does
    (src/utils.py:468-469).  ``allow_random_init=True`` (the scripts' ``--allow-random-init``, smoke runs only) keeps
    the random initialisation instead and says so."""
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
    if ckpt_path:
        if os.path.exists(ckpt_path):
            # trusted checkpoint ({"state_dict": ...}, Lightning's or scripts/train3D_c_c.py's): full unpickling
            state = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            model.load_state_dict(state["state_dict"])
        elif allow_random_init:
            import warnings
            warnings.warn(f"vdm4cdm_b200.utils.get_model: checkpoint {ckpt_path!r} does not exist; the model keeps its "
                          "RANDOM initialisation (allow_random_init=True)")
        else:
            raise FileNotFoundError(f"get_model: ckpt_path {ckpt_path!r} is configured but the file does not exist "
                                    "(pass allow_random_init=True / --allow-random-init for a smoke run on random weights)")
    if device is not None:
        model = model.to(device)
    return model
