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


def test_gn_silu_view_channel_window_and_fused_upsampling():
    """vdm_gn_silu_view: GroupNorm+SiLU over cat([interpolate(coarse), skip]) without the up-sampled tensor; groups
    straddle the boundary between the two parts (96 channels, 8 groups of 12, boundary at 64)."""
    from vdm4cdm_b200 import ops
    g = torch.Generator().manual_seed(4)
    b, c_up, c_skip, grid = 2, 64, 32, (8, 12, 16)
    cg = tuple(n // 2 for n in grid)
    coarse = (torch.randn((b, c_up) + cg, generator=g) * 1.5 + 0.3).to(torch.bfloat16).float()
    skip = (torch.randn((b, c_skip) + grid, generator=g) * 0.7 - 0.2).to(torch.bfloat16).float()
    cat = torch.cat([F.interpolate(coarse, scale_factor=2, mode="nearest"), skip], dim=1)
    gamma, beta = torch.randn(96, generator=g), torch.randn(96, generator=g)
    want = F.silu(F.group_norm(cat, 8, gamma, beta, 1e-5))
    # statistics of the fine-resolution concat: 8x the coarse sums for the up-sampled channels
    stats = torch.zeros((b, 96, 2), dtype=torch.float64)
    stats[:, :c_up, 0] = 8 * coarse.double().sum(dim=(2, 3, 4))
    stats[:, :c_up, 1] = 8 * (coarse.double() ** 2).sum(dim=(2, 3, 4))
    stats[:, c_up:, 0] = skip.double().sum(dim=(2, 3, 4))
    stats[:, c_up:, 1] = (skip.double() ** 2).sum(dim=(2, 3, 4))
    stats = stats.cuda()
    # coarse tensor in a window of a wider buffer; skip in the concat buffer at its usual planes
    cbuf = torch.zeros((b, c_up // 8 + 2) + cg + (8,), dtype=torch.bfloat16, device="cuda")
    cbuf[:, 1:1 + c_up // 8] = ops.to_planar(coarse.cuda())
    catbuf = torch.full((b, 12) + grid + (8,), 7.0, dtype=torch.bfloat16, device="cuda")
    catbuf[:, c_up // 8:] = ops.to_planar(skip.cuda())
    out = torch.zeros((b, 12) + grid + (8,), dtype=torch.bfloat16, device="cuda")
    ops.gn_silu_view(cbuf, c_up, 0, 96, 8, stats, gamma.cuda(), beta.cuda(), 1e-5, out, x_plane0=1, out_plane0=0, upsample=True)
    ops.gn_silu_view(catbuf, c_skip, c_up, 96, 8, stats, gamma.cuda(), beta.cuda(), 1e-5, out, x_plane0=c_up // 8,
                     out_plane0=c_up // 8)
    got = ops.from_planar(out, 96).cpu()
    assert torch.allclose(got, want, rtol=1e-2, atol=1e-2), (got - want).abs().max()
    assert ((got - want).norm() / want.norm()).item() < 3e-3
    with pytest.raises(RuntimeError, match="outside the 96-channel norm"):
        ops.gn_silu_view(catbuf, 64, 64, 96, 8, stats, gamma.cuda(), beta.cuda(), 1e-5, out)
