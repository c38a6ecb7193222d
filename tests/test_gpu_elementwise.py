"""GPU parity of the fused elementwise kernels against plain PyTorch fp32 ops on the same inputs."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from vdm4cdm_b200 import ops
    return ops


@pytest.mark.parametrize("shape", [(2, 32, 6, 10, 14), (1, 96, 4, 8, 8), (3, 16, 2, 6, 4)])
def test_channel_stats(shape):
    ops = _ops()
    dev = torch.device("cuda:0")
    x = torch.randn(shape, generator=torch.Generator().manual_seed(1)).to(dev)
    xp = ops.to_planar(x)
    xr = ops.from_planar(xp).double()
    st = ops.channel_stats(xp, shape[1])
    torch.cuda.synchronize()
    assert torch.allclose(st[..., 0], xr.sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4)
    assert torch.allclose(st[..., 1], (xr ** 2).sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("shape,groups", [((2, 32, 6, 10, 14), 8), ((1, 96, 4, 8, 8), 8), ((2, 16, 4, 6, 6), 8),
                                          ((1, 384, 2, 4, 4), 8)])
def test_gn_silu_matches_torch(shape, groups):
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(shape, generator=g) * 1.7 + 0.3).to(dev)
    gamma = (torch.rand(shape[1], generator=g) + 0.5).to(dev)
    beta = (torch.randn(shape[1], generator=g) * 0.2).to(dev)
    xp = ops.to_planar(x)
    xr = ops.from_planar(xp)
    st = ops.channel_stats(xp, shape[1])
    y = ops.gn_silu(xp, shape[1], groups, st, gamma, beta, 1e-5)
    torch.cuda.synchronize()
    ref = F.silu(F.group_norm(xr, groups, gamma, beta, 1e-5))
    got = ops.from_planar(y)
    # output is bf16: half an ulp = 2^-9 relative
    assert torch.allclose(got, ref, rtol=2 ** -8, atol=2e-3), (got - ref).abs().max()


def test_gn_silu_dropout_is_a_scaled_mask():
    ops = _ops()
    dev = torch.device("cuda:0")
    shape, p = (2, 32, 8, 16, 16), 0.1
    x = torch.randn(shape, generator=torch.Generator().manual_seed(3)).to(dev)
    gamma, beta = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    xp = ops.to_planar(x)
    st = ops.channel_stats(xp, 32)
    y0 = ops.from_planar(ops.gn_silu(xp, 32, 8, st, gamma, beta))
    y1 = ops.from_planar(ops.gn_silu(xp, 32, 8, st, gamma, beta, dropout_p=p, seed=42, layer_tag=3))
    y2 = ops.from_planar(ops.gn_silu(xp, 32, 8, st, gamma, beta, dropout_p=p, seed=42, layer_tag=3))
    y3 = ops.from_planar(ops.gn_silu(xp, 32, 8, st, gamma, beta, dropout_p=p, seed=42, layer_tag=4))
    torch.cuda.synchronize()
    assert torch.equal(y1, y2), "dropout must be a pure function of (seed, layer_tag, element)"
    assert not torch.equal(y1, y3)
    dropped = (y1 == 0) & (y0 != 0)
    frac = dropped.float().mean().item()
    assert abs(frac - p) < 0.01, frac
    kept = ~dropped
    expect = (y0 / (1 - p)).to(torch.bfloat16).float()
    assert torch.allclose(y1[kept], expect[kept], rtol=2 ** -7, atol=1e-3)


def test_avgpool_and_upsample_with_stats_and_windows():
    ops = _ops()
    dev = torch.device("cuda:0")
    b, c, d, h, w = 2, 32, 4, 8, 12
    x = torch.randn((b, c, d, h, w), generator=torch.Generator().manual_seed(4)).to(dev)
    xp = ops.to_planar(x)
    xr = ops.from_planar(xp)
    st = torch.zeros((b, c, 2), dtype=torch.float64, device=dev)
    yp = ops.avgpool2(xp, c, stats=st)
    torch.cuda.synchronize()
    ref = F.avg_pool3d(xr, 2).to(torch.bfloat16).float()
    got = ops.from_planar(yp)
    assert torch.allclose(got, ref, rtol=2 ** -7, atol=1e-6)
    assert torch.allclose(st[..., 0], got.double().sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4)
    assert torch.allclose(st[..., 1], (got.double() ** 2).sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4)
    # up-sample the pooled tensor into planes 0..3 of a 6-plane concat buffer
    cat = torch.zeros((b, 6, d, h, w, 8), dtype=torch.bfloat16, device=dev)
    st2 = torch.zeros((b, 48, 2), dtype=torch.float64, device=dev)
    ops.upsample2(yp, c, cat, stats=st2)
    torch.cuda.synchronize()
    up = F.interpolate(got, scale_factor=2, mode="nearest")
    assert torch.equal(ops.from_planar(cat[:, :4].contiguous()), up)
    assert cat[:, 4:].float().abs().sum().item() == 0
    assert torch.allclose(st2[:, :32, 0], up.double().sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4)
    assert torch.allclose(st2[:, :32, 1], (up.double() ** 2).sum(dim=(2, 3, 4)), rtol=1e-5, atol=1e-4)


def test_pack_input_layout():
    ops = _ops()
    dev = torch.device("cuda:0")
    b, d, h, w = 2, 4, 6, 8
    g = torch.Generator().manual_seed(5)
    z = torch.randn((b, 1, d, h, w), generator=g).to(dev)
    cond = torch.randn((b, 1, d, h, w), generator=g).to(dev)
    packed = ops.pack_input(z, cond, 16)
    torch.cuda.synchronize()
    dense = ops.from_planar(packed)
    assert torch.equal(dense[:, 0:1], z.to(torch.bfloat16).float())
    assert torch.equal(dense[:, 1:2], cond.to(torch.bfloat16).float())
    assert dense[:, 2:].abs().sum().item() == 0
