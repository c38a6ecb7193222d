"""Oracle LightSFM: plain PyTorch fp32 restatement (TEST INFRASTRUCTURE, parity unpinned).

The reference imports ``mltools.models.sfm_model.LightSFM`` (absent); only its
constructor call is visible (trainSFM3D160_c_c_from_field_name_thick_lowbatch.py:124-127:
``LightSFM(velocity_model=..., draw_figure=..., learning_rate=3.0e-4)``) together with the
batch schema ``{"x0", "x1", "conditioning_values"}`` (same file :71-72) and the fact that the
velocity network is a CUNet with one spatial conditioning channel and a time input
(:112-123).  The objective below is the standard (stochastic-interpolant) flow matching loss;
every choice is recorded in oracle/DECISIONS.md.  The reference never samples from SFM models
(generate_3D.py:16-17 raises NotImplementedError), so ``draw_samples`` is an addition.
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
        v = self.velocity_model(xt, t=times, s_conditioning=x0,
                                v_conditionings=batch.get("conditioning_values"))
        return ((v - (x1 - x0)) ** 2).mean()

    def training_step(self, batch, batch_idx=0):
        return self.get_loss(batch)

    def configure_optimizers(self):
        return torch.optim.AdamW(self.parameters(), lr=self.learning_rate)

    @torch.no_grad()
    def draw_samples(self, x0, n_sampling_steps=100, v_conditionings=None):
        """Euler integration of dx/dt = v(x, t | x0) from t=0 (x0) to t=1."""
        x = x0.clone()
        dt = 1.0 / n_sampling_steps
        for i in range(n_sampling_steps):
            t = torch.full((x0.shape[0],), i * dt, device=x0.device)
            x = x + dt * self.velocity_model(x, t=t, s_conditioning=x0, v_conditionings=v_conditionings)
        return x
