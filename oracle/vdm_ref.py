"""Oracle VDM / LightVDM: plain PyTorch fp32 restatement (TEST INFRASTRUCTURE, parity unpinned).

The reference imports these from the absent package ``mltools.models.vdm_model``
(trainVDM3D128_c_c_from_field_name_thick_lowbatch.py:128-132, src/utils.py:463-467).
Lines recoverable from the traceback in model_test.ipynb:678-682 are followed exactly:

  vdm_model.py:318-324  get_pred_noise: score_model(zt, t=(gamma_t-gamma_min)/(gamma_max-gamma_min), **kw)
                        when ``w_cfg is None or self.training``; otherwise a guidance branch that
                        needs ``v_conditionings``
  vdm_model.py:370-378  sample_zs_given_zt: sigma_t, sigma_s, pred_noise, and
                        ``mean = alpha_s / alpha_t * (zt - c * sigma_t * pred_noise)``
  vdm_model.py:429-442  sample: loop ``z = sample_zs_given_zt(zt=z, t=steps[i], s=steps[i+1], **kw)``
  vdm_model.py:531-557  LightVDM.draw_samples -> self.model.sample(..., device=self.device, ...)
  src/utils.py:286-299  steps = linspace(1, 0, n+1); return_ddnm=True gives (w_z, w_x_0t, x_0t, scale)
                        with z_s = w_z z + w_x_0t x_0t + scale eps

Everything else follows Kingma, Salimans, Poole, Ho (2021) "Variational Diffusion Models"
 and is recorded in oracle/DECISIONS.md.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


class LearnedLinearSchedule(nn.Module):
    """gamma(t) = b + |w| t, initialised to [gamma_min, gamma_max]."""

    def __init__(self, gamma_min: float, gamma_max: float):
        super().__init__()
        self.b = nn.Parameter(torch.tensor(float(gamma_min)))
        self.w = nn.Parameter(torch.tensor(float(gamma_max - gamma_min)))

    def forward(self, t):
        return self.b + self.w.abs() * t

    def slope(self):
        return self.w.abs()


class FixedLinearSchedule(nn.Module):
    def __init__(self, gamma_min: float, gamma_max: float):
        super().__init__()
        self.register_buffer("b", torch.tensor(float(gamma_min)))
        self.register_buffer("w", torch.tensor(float(gamma_max - gamma_min)))

    def forward(self, t):
        return self.b + self.w * t

    def slope(self):
        return self.w


class VDM(nn.Module):
    def __init__(self, score_model, noise_schedule: str = "learned_linear", gamma_min: float = -13.3,
                 gamma_max: float = 13.3, antithetic_time_sampling: bool = True,
                 data_noise: float = 1.0e-3, w_cfg=None):
        super().__init__()
        self.score_model = score_model
        self.gamma_min = gamma_min
        self.gamma_max = gamma_max
        self.antithetic_time_sampling = antithetic_time_sampling
        self.data_noise = data_noise
        self.w_cfg = w_cfg
        if noise_schedule == "learned_linear":
            self.gamma = LearnedLinearSchedule(gamma_min, gamma_max)
        elif noise_schedule == "fixed_linear":
            self.gamma = FixedLinearSchedule(gamma_min, gamma_max)
        else:
            raise ValueError(f"Unknown noise schedule {noise_schedule}")

    # ---- schedule helpers -------------------------------------------------
    @staticmethod
    def sigma(gamma):
        return torch.sqrt(torch.sigmoid(gamma))

    @staticmethod
    def alpha(gamma):
        return torch.sqrt(torch.sigmoid(-gamma))

    def _gamma5(self, t, ref):
        g = self.gamma(torch.as_tensor(t, dtype=torch.float32, device=ref.device))
        return g.reshape(-1, *([1] * (ref.dim() - 1)))

    # ---- network call -----------------------------------------------------
    def get_pred_noise(self, zt, gamma_t, **kwargs):
        t_net = ((gamma_t - self.gamma_min) / (self.gamma_max - self.gamma_min)).reshape(-1)
        if self.w_cfg is None or self.training:
            return self.score_model(zt, t=t_net, **kwargs)
        assert "v_conditionings" in kwargs, "Need v_conditionings to mask out"
        cond = self.score_model(zt, t=t_net, **kwargs)
        masked = dict(kwargs)
        masked["v_conditionings"] = [torch.zeros_like(v) for v in kwargs["v_conditionings"]]
        uncond = self.score_model(zt, t=t_net, **masked)
        return (1.0 + self.w_cfg) * cond - self.w_cfg * uncond

    # ---- forward / reverse transitions --------------------------------------
    def sample_zt_given_x(self, x, t, noise):
        gamma_t = self._gamma5(t, x)
        return self.alpha(gamma_t) * x + self.sigma(gamma_t) * noise, gamma_t

    def sample_zs_given_zt(self, zt, t, s, return_ddnm=False, noise=None, **kwargs):
        gamma_t = self._gamma5(t, zt)
        gamma_s = self._gamma5(s, zt)
        c = -torch.expm1(gamma_s - gamma_t)
        alpha_t = self.alpha(gamma_t)
        alpha_s = self.alpha(gamma_s)
        sigma_t = self.sigma(gamma_t)
        sigma_s = self.sigma(gamma_s)
        pred_noise = self.get_pred_noise(zt=zt, gamma_t=gamma_t, **kwargs)
        scale = sigma_s * torch.sqrt(c)
        if not return_ddnm:
            mean = alpha_s / alpha_t * (zt - c * sigma_t * pred_noise)
            if noise is None:
                noise = torch.randn_like(zt)
            return mean + scale * noise
        x_0t = (zt - sigma_t * pred_noise) / alpha_t
        return alpha_s * (1.0 - c) / alpha_t, alpha_s * c, x_0t, scale

    def sample_zt_given_zs(self, zs, t, s, noise=None):
        gamma_t = self._gamma5(t, zs)
        gamma_s = self._gamma5(s, zs)
        c = -torch.expm1(gamma_s - gamma_t)
        if noise is None:
            noise = torch.randn_like(zs)
        return self.alpha(gamma_t) / self.alpha(gamma_s) * zs + self.sigma(gamma_t) * torch.sqrt(c) * noise

    @torch.no_grad()
    def sample(self, batch_size, n_sampling_steps, device, z=None, return_all=False, verbose=False,
               noise_fn=None, **kwargs):
        """Ancestral sampling; ``noise_fn(draw, shape)`` injects the noise (draw 0 = initial latent)."""
        shape = (batch_size, *self.score_model.shape)
        if z is None:
            z = noise_fn(0, shape) if noise_fn is not None else torch.randn(shape, device=device)
        steps = torch.linspace(1.0, 0.0, n_sampling_steps + 1, device=device)
        zs = []
        for i in range(n_sampling_steps):
            noise = noise_fn(i + 1, shape) if noise_fn is not None else None
            z = self.sample_zs_given_zt(zt=z, t=steps[i], s=steps[i + 1], noise=noise, **kwargs)
            if return_all:
                zs.append(z)
        gamma_0 = self._gamma5(0.0, z)
        x = z / self.alpha(gamma_0)
        if return_all:
            return torch.stack(zs + [x], dim=0)
        return x

    # ---- training loss ------------------------------------------------------
    def sample_times(self, batch_size, device, t0=None):
        if self.antithetic_time_sampling:
            if t0 is None:
                t0 = torch.rand((), device=device)
            return torch.remainder(t0 + torch.arange(batch_size, device=device) / batch_size, 1.0)
        return torch.rand(batch_size, device=device)

    def get_loss(self, x, noise=None, noise0=None, times=None, **kwargs):
        """Continuous-time VDM loss in bits per dimension, plus its three terms (batch means)."""
        bsz = x.shape[0]
        red = tuple(range(1, x.dim()))
        if times is None:
            times = self.sample_times(bsz, x.device)
        if noise is None:
            noise = torch.randn_like(x)
        if noise0 is None:
            noise0 = torch.randn_like(x)
        zt, gamma_t = self.sample_zt_given_x(x, times, noise)
        pred = self.get_pred_noise(zt, gamma_t, **kwargs)
        diffusion = 0.5 * self.gamma.slope() * ((noise - pred) ** 2).sum(dim=red)

        gamma_1 = self._gamma5(1.0, x)
        var_1 = torch.sigmoid(gamma_1)
        latent = 0.5 * (var_1 + torch.sigmoid(-gamma_1) * x * x - torch.log(var_1) - 1.0).sum(dim=red)

        gamma_0 = self._gamma5(0.0, x)
        z0_rescaled = x + torch.exp(0.5 * gamma_0) * noise0          # (alpha_0 x + sigma_0 eps)/alpha_0
        dn = self.data_noise
        recons = (0.5 * ((x - z0_rescaled) / dn) ** 2 + math.log(dn) + 0.5 * math.log(2.0 * math.pi)).sum(dim=red)

        bpd = 1.0 / (x[0].numel() * math.log(2.0))
        loss = (diffusion + latent + recons).mean() * bpd
        return loss, {"diffusion_loss": diffusion.mean() * bpd, "latent_loss": latent.mean() * bpd,
                      "reconstruction_loss": recons.mean() * bpd}


class LightVDM(nn.Module):
    """Lightning-free stand-in for ``LightVDM`` (ctor: trainVDM3D128_...:128-132)."""

    def __init__(self, score_model, draw_figure=None, gamma_min=-13.3, gamma_max=13.3,
                 noise_schedule="learned_linear", learning_rate=3.0e-4, **vdm_kwargs):
        super().__init__()
        self.model = VDM(score_model, noise_schedule=noise_schedule, gamma_min=gamma_min,
                         gamma_max=gamma_max, **vdm_kwargs)
        self.draw_figure = draw_figure
        self.learning_rate = learning_rate

    @property
    def device(self):
        return next(self.parameters()).device

    def get_loss(self, batch, **kw):
        return self.model.get_loss(batch["x"], s_conditioning=batch.get("conditioning"),
                                   v_conditionings=batch.get("conditioning_values"), **kw)

    def training_step(self, batch, batch_idx=0):
        return self.get_loss(batch)[0]

    def configure_optimizers(self):
        return torch.optim.AdamW(self.parameters(), lr=self.learning_rate)

    def draw_samples(self, batch_size, n_sampling_steps=250, verbose=False, return_all=False, **kwargs):
        return self.model.sample(batch_size=batch_size, n_sampling_steps=n_sampling_steps,
                                 device=self.device, verbose=verbose, return_all=return_all, **kwargs)


def get_ddnm_result(vdm, y, A, AT, n_sampling_steps=250, l=10, return_all=False, noise_fn=None, **kwargs):
    """Oracle restatement of src/utils.py:277-304 (DDNM with time travel) with injectable noise:
    ``noise_fn(draw, shape)`` is called once for the initial latent and once per re-noise / reverse update, in loop
    order.  ``vdm`` is an oracle ``LightVDM``."""
    import numpy as np
    if isinstance(l, int):
        l = np.full(n_sampling_steps, l)
    l = np.asarray(l)
    model = vdm.model
    steps = torch.linspace(1.0, 0.0, n_sampling_steps + 1)
    shape = (y.shape[0], *model.score_model.shape)
    draw = [0]

    def noise():
        d = draw[0]
        draw[0] += 1
        return noise_fn(d, shape) if noise_fn is not None else torch.randn(shape)

    z = noise()
    ATy = AT(y)
    xs = []
    with torch.no_grad():
        for i in range(n_sampling_steps):
            L = int(min(l[i], i))
            z = model.sample_zt_given_zs(zs=z, t=steps[i - L], s=steps[i], noise=noise())
            for j in range(L, -1, -1):
                w_z, w_x_0t, x_0t, scale = model.sample_zs_given_zt(zt=z, t=steps[i - j], s=steps[i + 1 - j],
                                                                    return_ddnm=True, **kwargs)
                x_0t_r = ATy + x_0t - AT(A(x_0t))
                z = w_z * z + w_x_0t * x_0t_r + scale * noise()
            if return_all:
                xs.append(x_0t_r)
    return torch.stack(xs, dim=0) if return_all else x_0t_r
