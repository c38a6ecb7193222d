"""Oracle CUNet: plain PyTorch fp32 restatement (TEST INFRASTRUCTURE, parity unpinned).

The reference imports this network from the absent third-party package
``mltools.networks.networks`` (call sites: trainVDM3D128_c_c_from_field_name_thick_lowbatch.py:116-127,
src/utils.py:451-462).  What the reference pins, and what this file follows:

  * constructor keywords and ``.shape``               -- call sites above, src/utils.py:287
  * ``forward(x, t, s_conditioning, v_conditionings)`` with the ``downs`` loop,
    ``no_down`` on the last level and ``if h_skip is not None`` skip collection
                                                       -- model_test.ipynb:684 (networks.py:259-265)
  * ``ResNetDown.forward(x, conditionings, no_down)`` looping ``resnet_blocks``
    (+ optional ``attention_blocks``)                  -- model_test.ipynb:686 (blocks.py:166-170)
  * ``ResNetBlock.forward(x, conditionings)``: ``h = self.net1(x)`` with ``net1`` a
    Sequential that starts with GroupNorm; one projection per entry of
    ``conditioning_dims``                              -- model_test.ipynb:688-692 (blocks.py:129-132)

Everything else is this oracle's documented choice (oracle/DECISIONS.md), and the
oracle is then the parity authority for the CUDA path.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


def timestep_embedding(t: torch.Tensor, dim: int, scale: float = 1000.0) -> torch.Tensor:
    """Sinusoidal features of a (B,) time in [0,1]: [sin(s t f_j), cos(s t f_j)], f_j = 1e4^(-j/half)."""
    half = dim // 2
    j = torch.arange(half, dtype=torch.float32, device=t.device)
    freqs = torch.exp(-math.log(10000.0) * j / half)
    args = (t.to(torch.float32) * scale)[:, None] * freqs[None, :]
    return torch.cat([torch.sin(args), torch.cos(args)], dim=1)


class ResNetBlock(nn.Module):
    """Pre-activation residual block (blocks.py:129-132 in the traceback)."""

    def __init__(self, ch_in: int, ch_out: int, conditioning_dims: Sequence[int],
                 dropout_prob: float, norm_groups: int, padding_mode: str = "zeros"):
        super().__init__()
        self.ch_in, self.ch_out = ch_in, ch_out
        self.conditioning_dims = list(conditioning_dims)
        self.net1 = nn.Sequential(
            nn.GroupNorm(norm_groups, ch_in),
            nn.SiLU(),
            nn.Conv3d(ch_in, ch_out, 3, padding=1, padding_mode=padding_mode),
        )
        self.cond_projs = nn.ModuleList([nn.Linear(d, ch_out) for d in self.conditioning_dims])
        self.net2 = nn.Sequential(
            nn.GroupNorm(norm_groups, ch_out),
            nn.SiLU(),
            nn.Dropout(dropout_prob),
            nn.Conv3d(ch_out, ch_out, 3, padding=1, padding_mode=padding_mode),
        )
        self.skip_conv = nn.Conv3d(ch_in, ch_out, 1) if ch_in != ch_out else None

    def forward(self, x, conditionings=None):
        h = self.net1(x)
        if conditionings is not None:
            assert len(conditionings) == len(self.conditioning_dims)
            for c, proj in zip(conditionings, self.cond_projs):
                h = h + proj(c)[:, :, None, None, None]
        h = self.net2(h)
        return h + (x if self.skip_conv is None else self.skip_conv(x))


class ResNetDown(nn.Module):
    """Residual blocks, then keep a skip and halve the grid unless ``no_down`` (blocks.py:166-170)."""

    def __init__(self, resnet_blocks: List[ResNetBlock]):
        super().__init__()
        self.resnet_blocks = nn.ModuleList(resnet_blocks)
        self.attention_blocks = None

    def forward(self, x, conditionings, no_down=False):
        for i, resnet_block in enumerate(self.resnet_blocks):
            x = resnet_block(x, conditionings)
            if self.attention_blocks is not None:
                x = self.attention_blocks[i](x)
        if no_down:
            return x, None
        return F.avg_pool3d(x, 2), x


class ResNetUp(nn.Module):
    """Nearest x2 up-sampling, channel concat with the skip, residual blocks."""

    def __init__(self, resnet_blocks: List[ResNetBlock]):
        super().__init__()
        self.resnet_blocks = nn.ModuleList(resnet_blocks)

    def forward(self, x, x_skip, conditionings):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        x = torch.cat([x, x_skip], dim=1)
        for resnet_block in self.resnet_blocks:
            x = resnet_block(x, conditionings)
        return x


class CUNet(nn.Module):
    def __init__(self, shape=(1, 128, 128, 128), chs=(32, 64, 128, 256),
                 s_conditioning_channels: int = 0, v_conditioning_dims: Sequence[int] = (),
                 t_conditioning: bool = False, norm_groups: int = 8, mid_attn: bool = False,
                 dropout_prob: float = 0.1, conv_padding_mode: str = "zeros",
                 n_attention_heads: int = 4, t_embedding_dim: int = 64, v_embedding_dim: int = 64,
                 out_channels: Optional[int] = None):
        super().__init__()
        if len(shape) != 4:
            raise NotImplementedError("oracle covers the 3-D networks only (SURVEY.md section 2.1 rows 1-2)")
        if mid_attn:
            raise NotImplementedError("mid_attn=True is used by the 2-D scripts only (out of scope)")
        self.shape = tuple(shape)
        self.chs = list(chs)
        self.s_conditioning_channels = s_conditioning_channels
        self.v_conditioning_dims = list(v_conditioning_dims)
        self.t_conditioning = t_conditioning
        self.t_embedding_dim = t_embedding_dim
        self.v_embedding_dim = v_embedding_dim
        self.n_attention_heads = n_attention_heads
        in_ch = shape[0] + s_conditioning_channels
        out_ch = shape[0] if out_channels is None else out_channels
        pm = conv_padding_mode

        cond_dims = []
        if t_conditioning:
            self.t_embed = nn.Sequential(nn.Linear(t_embedding_dim, 4 * t_embedding_dim), nn.SiLU(),
                                         nn.Linear(4 * t_embedding_dim, t_embedding_dim), nn.SiLU())
            cond_dims.append(t_embedding_dim)
        self.v_embeds = nn.ModuleList()
        for d in self.v_conditioning_dims:
            self.v_embeds.append(nn.Sequential(nn.Linear(d, 4 * v_embedding_dim), nn.SiLU(),
                                               nn.Linear(4 * v_embedding_dim, v_embedding_dim), nn.SiLU()))
            cond_dims.append(v_embedding_dim)
        self.conditioning_dims = cond_dims

        def block(ci, co):
            return ResNetBlock(ci, co, cond_dims, dropout_prob, norm_groups, pm)

        c = self.chs
        self.conv_in = nn.Conv3d(in_ch, c[0], 3, padding=1, padding_mode=pm)
        self.downs = nn.ModuleList(
            [ResNetDown([block(c[max(i - 1, 0)], c[i])]) for i in range(len(c))])
        self.mid1 = block(c[-1], c[-1])
        self.mid2 = block(c[-1], c[-1])
        self.ups = nn.ModuleList(
            [ResNetUp([block(c[i + 1] + c[i], c[i])]) for i in reversed(range(len(c) - 1))])
        self.conv_out = nn.Sequential(nn.GroupNorm(norm_groups, c[0]), nn.SiLU(),
                                      nn.Conv3d(c[0], out_ch, 3, padding=1, padding_mode=pm))

    def conditionings(self, batch: int, t, v_conditionings, device):
        out = []
        if self.t_conditioning:
            tt = torch.as_tensor(t, dtype=torch.float32, device=device).reshape(-1)
            if tt.numel() == 1:
                tt = tt.expand(batch)
            out.append(self.t_embed(timestep_embedding(tt, self.t_embedding_dim)))
        v_conditionings = [] if v_conditionings is None else v_conditionings
        assert len(v_conditionings) == len(self.v_embeds)
        for v, emb in zip(v_conditionings, self.v_embeds):
            out.append(emb(v.to(torch.float32)))
        return out if len(out) else None

    def forward(self, x, t=None, s_conditioning=None, v_conditionings=None):
        conditionings = self.conditionings(x.shape[0], t, v_conditionings, x.device)
        h = x if s_conditioning is None else torch.cat([x, s_conditioning], dim=1)
        h = self.conv_in(h)
        skips = []
        for i, down in enumerate(self.downs):
            h, h_skip = down(h, conditionings=conditionings, no_down=(i == (len(self.downs) - 1)))
            if h_skip is not None:
                skips.append(h_skip)
        h = self.mid1(h, conditionings)
        h = self.mid2(h, conditionings)
        for up in self.ups:
            h = up(h, skips.pop(), conditionings)
        return self.conv_out(h)


def conv_flops(model: CUNet, x, **kw) -> int:
    """Algorithmic conv FLOPs of one forward: sum over Conv3d modules of 2 k^3 Cin Cout D H W B
    (collect with forward hooks, do not hand-count)."""
    total = [0]
    hooks = []

    def hook(mod, inp, out):
        k = mod.kernel_size[0] * mod.kernel_size[1] * mod.kernel_size[2]
        total[0] += 2 * k * mod.in_channels * mod.out_channels * out.shape[0] * out.shape[2] * out.shape[3] * out.shape[4]

    for m in model.modules():
        if isinstance(m, nn.Conv3d):
            hooks.append(m.register_forward_hook(hook))
    with torch.no_grad():
        model(x, **kw)
    for h in hooks:
        h.remove()
    return total[0]
