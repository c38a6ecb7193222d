"""``LightSFM``: stochastic flow matching around the CUNet velocity network, B200-native.

Mirrors what the reference imports as ``mltools.models.sfm_model.LightSFM`` (constructor call:
trainSFM3D160_c_c_from_field_name_thick_lowbatch.py:124-127 ``LightSFM(velocity_model=..., draw_figure=...,
learning_rate=3.0e-4)``; batch schema ``{"x0", "x1", "conditioning_values"}`` :71-72; the velocity network is a
CUNet with one spatial conditioning channel and a time input :112-123).  The module is absent from the
reference checkout; the objective is the one fixed in oracle/DECISIONS.md (``oracle/sfm_ref.py``):
``x_t = (1-t) x0 + t x1 (+ sigma sqrt(t(1-t)) eps)``, target ``x1 - x0``, mean squared error of
``v(x_t, t; s_conditioning=x0, v_conditionings=params)``.

The network forward/backward run on the sm_100a kernels (``vdm4cdm_b200.autograd``); the interpolation and
the MSE are a handful of elementwise torch ops on (B, 1, N^3) fp32 tensors.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class LightSFM(nn.Module):
    def __init__(self, velocity_model, draw_figure=None, learning_rate=3.0e-4, sigma: float = 0.0):
        super().__init__()
        self.velocity_model = velocity_model
        self.draw_figure = draw_figure
        self.learning_rate = learning_rate
        self.sigma = sigma

    @property
    def device(self):
        return next(self.parameters()).device

    def get_loss(self, batch, times=None, noise=None):
        x0, x1 = batch["x0"], batch["x1"]
        bsz = x0.shape[0]
        if times is None:
            times = torch.rand(bsz, device=x0.device)
        tb = times.reshape(-1, *([1] * (x0.dim() - 1)))
        xt = (1.0 - tb) * x0 + tb * x1
        if self.sigma > 0.0:
            if noise is None:
                noise = torch.randn_like(x0)
            xt = xt + self.sigma * torch.sqrt(tb * (1.0 - tb)) * noise
        v = self.velocity_model(xt, t=times, s_conditioning=x0, v_conditionings=batch.get("conditioning_values"))
        return ((v - (x1 - x0)) ** 2).mean()

    def training_step(self, batch, batch_idx=0):
        return self.get_loss(batch)

    def validation_step(self, batch, batch_idx=0):
        with torch.no_grad():
            return self.get_loss(batch)

    def configure_optimizers(self):
        return torch.optim.AdamW(self.parameters(), lr=self.learning_rate)

    @torch.no_grad()
    def draw_samples(self, x0, n_sampling_steps=100, v_conditionings=None):
        """Euler integration of dx/dt = v(x, t | x0) from t = 0 (x0) to t = 1.  (The reference never samples
        its SFM models: generate_3D.py:16-17 raises NotImplementedError for them.)"""
        x = x0.clone()
        dt = 1.0 / n_sampling_steps
        for i in range(n_sampling_steps):
            t = torch.full((x0.shape[0],), i * dt, device=x0.device)
            x = x + dt * self.velocity_model(x, t=t, s_conditioning=x0, v_conditionings=v_conditionings)
        return x
