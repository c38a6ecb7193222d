"""GPU parity of vdm_conv3d (tcgen05 implicit GEMM on halo tiles) against torch's conv3d.

Integer-valued inputs make every product and partial sum exact in fp32, so the comparison with the
fp32 reference is BIT-EXACT after the same bf16 rounding; random-normal inputs are checked to the
bf16 tolerance named by BASELINE.json's north_star (1e-2 relative).
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from vdm4cdm_b200 import ops
    return ops


def _int_tensor(shape, lo, hi, gen, device):
    return torch.randint(lo, hi + 1, shape, generator=gen, device="cpu").float().to(device)


def _reference(x, w, chan_add=None, residual=None, exact_integers=True):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    k = w.shape[-1]
    y = F.conv3d(x.double(), w.double(), padding=k // 2)
    if exact_integers:
        y = y.round()      # torch's fp64 conv is not exact to the last bit; the true sums are integers
    y = y.float()
    if chan_add is not None:
        y = y + chan_add[:, :, None, None, None]
    if residual is not None:
        y = y + residual
    return y


CASES = [
    # B, Cin, Cout, D, H, W, k
    (1, 16, 16, 4, 16, 8, 3),
    (1, 32, 32, 8, 16, 16, 3),
    (2, 32, 32, 6, 20, 12, 3),      # ragged: H, W not multiples of the 16x8 tile, D not of MT
    (1, 64, 64, 8, 16, 16, 3),
    (1, 96, 32, 4, 16, 16, 3),      # concat-shaped input (KC=32, 3 chunks)
    (1, 32, 64, 5, 9, 7, 3),        # tiny odd grid
    (1, 256, 256, 4, 16, 8, 3),     # coarse level: output channels split over CTAs
    (1, 384, 128, 2, 16, 8, 3),
    (1, 96, 32, 4, 16, 16, 1),      # 1x1x1 skip conv
    (3, 16, 32, 3, 8, 8, 3),        # H < tile height
    (1, 48, 16, 5, 16, 8, 3),       # kd-folded schedule over 3 channel chunks (the 224^3 network's level-0 up block)
    (2, 64, 32, 7, 20, 12, 3),      # folded, 2 or 4 chunks, ragged grid
    (1, 128, 32, 4, 16, 8, 3),      # folded only if the resident weights fit; else the generic path
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv3d_exact_integers(case):
    ops = _ops()
    b, ci, co, d, h, w, k = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1234 + ci + co)
    x = _int_tensor((b, ci, d, h, w), -2, 2, g, dev)
    wt = _int_tensor((co, ci, k, k, k), -1, 1, g, dev)
    ref = _reference(x, wt).to(torch.bfloat16).float()
    taps = ops.TAPS_3X3X3 if k == 3 else ops.TAPS_1X1X1
    y = ops.conv3d(ops.to_planar(x), ops.pack_conv_weight(wt), co, taps=taps)
    torch.cuda.synchronize()
    got = ops.from_planar(y, co)
    bad = (got != ref).sum().item()
    assert bad == 0, f"{bad} of {ref.numel()} outputs differ; max |diff| = {(got - ref).abs().max().item()}"


def test_conv3d_epilogue_bias_residual_stats_and_windows():
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(7)
    b, ci, co, d, h, w = 2, 32, 32, 6, 16, 16
    x = _int_tensor((b, ci, d, h, w), -2, 2, g, dev)
    wt = _int_tensor((co, ci, 3, 3, 3), -1, 1, g, dev)
    cadd = _int_tensor((b, co), -3, 3, g, dev)
    res = _int_tensor((b, co, d, h, w), -4, 4, g, dev)
    ref = _reference(x, wt, cadd, res).to(torch.bfloat16).float()
    # x lives in planes 2..5 of a 7-plane buffer, the output goes to planes 1..4 of a 6-plane buffer,
    # the residual sits at plane 3 of an 8-plane buffer, stats at channel 8 of a 48-channel table
    xbuf = torch.zeros((b, 7, d, h, w, 8), dtype=torch.bfloat16, device=dev)
    xbuf[:, 2:6] = ops.to_planar(x)
    rbuf = torch.zeros((b, 8, d, h, w, 8), dtype=torch.bfloat16, device=dev)
    rbuf[:, 3:7] = ops.to_planar(res)
    ybuf = torch.full((b, 6, d, h, w, 8), 77.0, dtype=torch.bfloat16, device=dev)
    stats = torch.zeros((b, 48, 2), dtype=torch.float64, device=dev)
    ops.conv3d(xbuf, ops.pack_conv_weight(wt), co, x_plane0=2, c_in=ci, out=ybuf, out_plane0=1, chan_add=cadd,
               residual=rbuf, residual_plane0=3, stats=stats, stats_c0=8)
    torch.cuda.synchronize()
    got = ops.from_planar(ybuf[:, 1:5].contiguous(), co)
    assert (got != ref).sum().item() == 0, f"max |diff| = {(got - ref).abs().max().item()}"
    assert (ybuf[:, 0].float() == 77.0).all() and (ybuf[:, 5].float() == 77.0).all(), "wrote outside the plane window"
    s1 = ref.double().sum(dim=(2, 3, 4))
    s2 = (ref.double() ** 2).sum(dim=(2, 3, 4))
    assert torch.allclose(stats[:, 8:40, 0], s1, rtol=1e-6, atol=1e-3), (stats[:, 8:40, 0] - s1).abs().max()
    assert torch.allclose(stats[:, 8:40, 1], s2, rtol=1e-6, atol=1e-3), (stats[:, 8:40, 1] - s2).abs().max()
    assert stats[:, :8].abs().sum().item() == 0 and stats[:, 40:].abs().sum().item() == 0


@pytest.mark.parametrize("b,ci,co,d,h,w", [
    (1, 32, 32, 37, 16, 8),       # one long column: ring wraps twice, segments of unequal length
    (40, 16, 16, 3, 16, 8),       # more samples than the shared-memory bias table holds (rows read from global memory)
    (2, 32, 16, 19, 20, 12),      # ragged faces, depth not a multiple of the segment length
    (3, 16, 32, 5, 9, 7),         # tiny odd grid: a segment shorter than the three-slice span
    (1, 32, 32, 130, 16, 8),      # the 128^3 regime: 130 slices in one column
    (2, 32, 32, 20, 32, 16),      # even grid: also with the coarse (nearest x2) residual
])
def test_conv3d_marching_schedule_edges_exact_integers(b, ci, co, d, h, w):
    """The d-marching schedule (conv3d_march.cuh: narrow 3x3x3 layers) on the cases specific to it: TMEM ring wrap, unit /
    segment boundaries, segments shorter than the kd span, ragged faces, the bias-row fallback, with bias + residual +
    statistics and a coarse (up-sampled) residual -- exact on integers."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(500 + d + b)
    x = _int_tensor((b, ci, d, h, w), -2, 2, g, dev)
    wt = _int_tensor((co, ci, 3, 3, 3), -1, 1, g, dev)
    cadd = _int_tensor((b, co), -3, 3, g, dev)
    res = _int_tensor((b, co, d, h, w), -4, 4, g, dev)
    ref = _reference(x, wt, cadd, res).to(torch.bfloat16).float()
    stats = torch.zeros((b, co, 2), dtype=torch.float64, device=dev)
    y = ops.conv3d(ops.to_planar(x), ops.pack_conv_weight(wt), co, chan_add=cadd, residual=ops.to_planar(res), stats=stats)
    torch.cuda.synchronize()
    got = ops.from_planar(y, co)
    assert (got != ref).sum().item() == 0, f"max |diff| = {(got - ref).abs().max().item()}"
    assert torch.allclose(stats[..., 0], ref.double().sum(dim=(2, 3, 4)), rtol=1e-6, atol=1e-3)
    assert torch.allclose(stats[..., 1], (ref.double() ** 2).sum(dim=(2, 3, 4)), rtol=1e-6, atol=1e-3)
    if d % 2 == 0 and h % 2 == 0 and w % 2 == 0:
        rc = _int_tensor((b, co, d // 2, h // 2, w // 2), -4, 4, g, dev)
        ref2 = _reference(x, wt, cadd, F.interpolate(rc, scale_factor=2, mode="nearest")).to(torch.bfloat16).float()
        y2 = ops.conv3d(ops.to_planar(x), ops.pack_conv_weight(wt), co, chan_add=cadd, residual=ops.to_planar(rc),
                        residual_upsample=True)
        torch.cuda.synchronize()
        assert (ops.from_planar(y2, co) != ref2).sum().item() == 0


def test_conv3d_fp32_single_channel_output():
    """conv_out of the UNet: C -> 1, fp32 NCDHW result."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    b, ci, d, h, w = 2, 32, 5, 16, 24
    x = _int_tensor((b, ci, d, h, w), -2, 2, g, dev)
    wt = _int_tensor((1, ci, 3, 3, 3), -1, 1, g, dev)
    cadd = _int_tensor((b, 1), -3, 3, g, dev)
    ref = _reference(x, wt, cadd)
    y = ops.conv3d(ops.to_planar(x), ops.pack_conv_weight(wt), 1, out_fp32=True, chan_add=cadd)
    torch.cuda.synchronize()
    assert (y != ref).sum().item() == 0, f"max |diff| = {(y - ref).abs().max().item()}"


@pytest.mark.parametrize("case", [(1, 32, 32, 16, 32, 32), (1, 128, 128, 8, 16, 16), (2, 64, 32, 8, 24, 16)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv3d_random_normal_bf16_tolerance(case):
    ops = _ops()
    b, ci, co, d, h, w = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(99)
    x = torch.randn((b, ci, d, h, w), generator=g).to(dev)
    wt = (torch.randn((co, ci, 3, 3, 3), generator=g) / (27 * ci) ** 0.5).to(dev)
    xb, wb = x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float()
    ref = _reference(xb, wb, exact_integers=False)
    got = ops.from_planar(ops.conv3d(ops.to_planar(x), ops.pack_conv_weight(wt), co), co)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-2, f"relative max error {err}"   # bf16 output rounding: 2^-8 relative


def test_conv3d_dgrad_weight_packing():
    """dgrad = the same kernel run with the flipped, transposed filter."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    b, ci, co, d, h, w = 1, 32, 64, 4, 16, 8
    wt = _int_tensor((co, ci, 3, 3, 3), -1, 1, g, dev)
    dy = _int_tensor((b, co, d, h, w), -2, 2, g, dev)
    x = torch.zeros((b, ci, d, h, w), device=dev, requires_grad=True)
    F.conv3d(x, wt, padding=1).backward(dy)
    ref = x.grad.to(torch.bfloat16).float()
    got = ops.from_planar(ops.conv3d(ops.to_planar(dy), ops.pack_conv_weight(wt, transpose_flip=True), ci), ci)
    torch.cuda.synchronize()
    assert (got != ref).sum().item() == 0, f"max |diff| = {(got - ref).abs().max().item()}"


def test_conv3d_rejects_bad_arguments():
    ops = _ops()
    dev = torch.device("cuda:0")
    x = torch.zeros((1, 1, 4, 16, 8, 8), dtype=torch.bfloat16, device=dev)   # 8 channels: not a multiple of 16
    w = torch.zeros((27, 1, 16, 8), dtype=torch.bfloat16, device=dev)
    with pytest.raises(RuntimeError, match="multiple of 16"):
        ops.conv3d(x, w, 16)


@pytest.mark.parametrize("case", [(2, 32, 32, 8, 16, 16), (1, 96, 32, 5, 20, 12), (1, 64, 64, 6, 16, 8), (1, 16, 128, 4, 16, 8)],
                         ids=lambda c: "x".join(map(str, c)))
def test_conv3d_circular_padding_exact_integers(case):
    """Conv3d(padding_mode="circular") of the cropsize == 256 models: vdm_pad_circular + vdm_conv3d(circular=1)."""
    ops = _ops()
    b, ci, co, d, h, w = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(77 + ci + co)
    x = _int_tensor((b, ci, d, h, w), -2, 2, g, dev)
    wt = _int_tensor((co, ci, 3, 3, 3), -1, 1, g, dev)
    xp = F.pad(x.double(), (1, 1, 1, 1, 1, 1), mode="circular")
    ref = F.conv3d(xp, wt.double()).round().float().to(torch.bfloat16).float()
    xpad = ops.pad_circular(ops.to_planar(x), ci)
    assert torch.equal(ops.from_planar(xpad), xp.float())
    y = ops.conv3d(xpad, ops.pack_conv_weight(wt), co, circular=True)
    torch.cuda.synchronize()
    assert torch.equal(ops.from_planar(y, co), ref)
    # dgrad and wgrad of the circular conv
    xg = x.double().clone().requires_grad_(True)
    wg = wt.double().clone().requires_grad_(True)
    dy = _int_tensor((b, co, d, h, w), -1, 1, g, dev)
    F.conv3d(F.pad(xg, (1, 1, 1, 1, 1, 1), mode="circular"), wg).backward(dy.double())
    dyp = ops.pad_circular(ops.to_planar(dy), co)
    dx = ops.conv3d(dyp, ops.pack_conv_weight(wt, transpose_flip=True), ci, circular=True)
    assert torch.equal(ops.from_planar(dx, ci), xg.grad.round().float().to(torch.bfloat16).float())
    dw = ops.conv3d_wgrad(xpad, ops.pad_circular(ops.to_planar(dy, 16), -(-co // 16) * 16), ci, co, 3, a_padded=True, g_padded=True)
    assert torch.equal(ops.wgrad_to_torch(dw, 3), wg.grad.round().float())


def test_conv3d_residual_read_through_upsampling():
    """residual_upsample: y = conv1x1(x) + bias + interpolate(r_coarse) -- the split skip conv of the up blocks."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(21)
    for (b, ci, co, d, h, w, k) in [(2, 32, 32, 6, 20, 12, 1), (1, 64, 64, 4, 16, 8, 3), (1, 16, 128, 2, 6, 10, 1)]:
        x = _int_tensor((b, ci, d, h, w), -2, 2, g, dev)
        wt = _int_tensor((co, ci, k, k, k), -1, 1, g, dev)
        cadd = _int_tensor((b, co), -3, 3, g, dev)
        rc = _int_tensor((b, co, d // 2, h // 2, w // 2), -4, 4, g, dev)
        ref = _reference(x, wt, cadd, F.interpolate(rc, scale_factor=2, mode="nearest")).to(torch.bfloat16).float()
        rbuf = torch.zeros((b, co // 8 + 3, d // 2, h // 2, w // 2, 8), dtype=torch.bfloat16, device=dev)
        rbuf[:, 2:2 + co // 8] = ops.to_planar(rc)
        taps = ops.TAPS_3X3X3 if k == 3 else ops.TAPS_1X1X1
        y = ops.conv3d(ops.to_planar(x), ops.pack_conv_weight(wt), co, taps=taps, chan_add=cadd, residual=rbuf,
                       residual_plane0=2, residual_upsample=True)
        torch.cuda.synchronize()
        assert torch.equal(ops.from_planar(y, co), ref), (b, ci, co, d, h, w, k)
    with pytest.raises(RuntimeError, match="even grid"):
        ops.conv3d(ops.to_planar(torch.zeros((1, 16, 3, 8, 8), device=dev)), ops.pack_conv_weight(torch.zeros((16, 16, 1, 1, 1), device=dev)),
                   16, taps=ops.TAPS_1X1X1, residual=torch.zeros((1, 2, 1, 4, 4, 8), dtype=torch.bfloat16, device=dev),
                   residual_upsample=True)


@pytest.mark.parametrize("ci,co", [(16, 16), (32, 32), (96, 32), (64, 64)])
def test_conv3d_circular_is_shift_equivariant_bit_for_bit(ci, co):
    """Periodic padding: conv(roll(x)) == roll(conv(x)) BIT FOR BIT on random bf16 data for shifts that are not multiples of
    the tile (every output voxel sums the same products in the same order wherever it sits -- tile and marching schedules
    alike), and the GroupNorm statistics of the two runs agree to fp32 summation order."""
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(5)
    b, d, h, w = 2, 16, 32, 16
    x = torch.randn((b, ci // 8, d, h, w, 8), device=dev, generator=g).to(torch.bfloat16)
    wt = ops.pack_conv_weight(torch.randn((co, ci, 3, 3, 3), device=dev, generator=g) / (27 * ci) ** 0.5)
    cadd = torch.randn((b, co), device=dev, generator=g)
    outs = []
    for sh in ((0, 0, 0), (3, 5, 4), (4, 8, 4)):
        st = torch.zeros((b, co, 2), dtype=torch.float64, device=dev)
        y = ops.conv3d(ops.pad_circular(torch.roll(x, sh, (2, 3, 4)), ci), wt, co, chan_add=cadd, stats=st, circular=True)
        torch.cuda.synchronize()
        outs.append((sh, y, st))
    for sh, y, st in outs[1:]:
        assert torch.equal(y, torch.roll(outs[0][1], sh, (2, 3, 4))), sh
        scale = outs[0][2][..., 1:].sqrt() * (d * h * w) ** 0.5          # ~ sum |y|: the sums' natural scale
        assert float(((st[..., 0] - outs[0][2][..., 0]).abs() / scale[..., 0]).max()) < 1e-6
        assert float(((st[..., 1] - outs[0][2][..., 1]).abs() / outs[0][2][..., 1]).max()) < 1e-6


@pytest.mark.parametrize("case", [(2, 32, 32, 32, 6, 20, 12), (1, 16, 16, 16, 5, 16, 8), (1, 32, 32, 64, 4, 16, 16),
                                  (1, 64, 32, 32, 8, 16, 8), (1, 16, 32, 48, 3, 9, 7)], ids=lambda c: "x".join(map(str, c)))
def test_conv3d_fused_skip_conv_exact_integers(case):
    """skip_x / skip_w: y = conv3x3x3(x) + conv1x1x1(x_skip) + bias + interpolate(r_coarse) in ONE launch (the up
    blocks' net2 conv + skip conv), with the GroupNorm statistics of the sum."""
    ops = _ops()
    b, ci, co, cs, d, h, w = case
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(300 + ci + cs)
    x = _int_tensor((b, ci, d, h, w), -2, 2, g, dev)
    xs = _int_tensor((b, cs, d, h, w), -2, 2, g, dev)
    wt = _int_tensor((co, ci, 3, 3, 3), -1, 1, g, dev)
    ws = _int_tensor((co, cs, 1, 1, 1), -1, 1, g, dev)
    cadd = _int_tensor((b, co), -3, 3, g, dev)
    even = d % 2 == 0 and h % 2 == 0 and w % 2 == 0
    ref = _reference(x, wt, cadd) + F.conv3d(xs.double(), ws.double()).round().float()
    rbuf = None
    if even:
        rc = _int_tensor((b, co, d // 2, h // 2, w // 2), -4, 4, g, dev)
        ref = ref + F.interpolate(rc, scale_factor=2, mode="nearest")
        rbuf = ops.to_planar(rc)
    ref = ref.to(torch.bfloat16).float()
    # the skip tensor sits at planes 3.. of a wider buffer (the concat buffer of the up block)
    sbuf = torch.zeros((b, cs // 8 + 4, d, h, w, 8), dtype=torch.bfloat16, device=dev)
    sbuf[:, 3:3 + cs // 8] = ops.to_planar(xs)
    stats = torch.zeros((b, co, 2), dtype=torch.float64, device=dev)
    y = ops.conv3d(ops.to_planar(x), ops.pack_conv_weight(wt), co, chan_add=cadd, residual=rbuf, residual_upsample=even,
                   stats=stats, skip_x=sbuf, skip_w=ops.pack_conv_weight(ws), skip_plane0=3)
    torch.cuda.synchronize()
    got = ops.from_planar(y, co)
    assert torch.equal(got, ref), f"{(got != ref).sum().item()} differ, max |diff| {(got - ref).abs().max().item()}"
    assert torch.allclose(stats[..., 0], ref.double().sum(dim=(2, 3, 4)), rtol=1e-6, atol=1e-3)


def test_conv3d_fused_skip_conv_declined_for_wide_layers():
    ops = _ops()
    dev = torch.device("cuda:0")
    x = torch.zeros((1, 8, 4, 16, 8, 8), dtype=torch.bfloat16, device=dev)
    w = ops.pack_conv_weight(torch.zeros((64, 64, 3, 3, 3), device=dev))
    ws = ops.pack_conv_weight(torch.zeros((64, 32, 1, 1, 1), device=dev))
    with pytest.raises(ops.UnsupportedFusion):
        ops.conv3d(x, w, 64, skip_x=torch.zeros((1, 4, 4, 16, 8, 8), dtype=torch.bfloat16, device=dev), skip_w=ws)


# ---- fused input transform: GroupNorm + SiLU applied by the conv kernel to its input tile (VdmConvEpilogue.in_norm) ----
def _gn_silu_ref(x, gamma, beta, groups, eps=1e-5):
    import torch.nn.functional as F
    return F.silu(F.group_norm(x.double(), groups, gamma.double(), beta.double(), eps))


@pytest.mark.parametrize("ci,co,grid,batch,taps", [
    (32, 32, (8, 32, 16), 2, 27),       # kd-folded, resident weights; two samples (coefficient table switch)
    (32, 32, (5, 20, 12), 1, 27),       # ragged tiles: every tile touches the boundary
    (64, 64, (8, 16, 16), 1, 27),       # kd-folded, streamed weights, two channel chunks
    (16, 32, (6, 16, 8), 2, 27),        # one 16-channel chunk: two planes, voxels split between warps
    (128, 128, (4, 16, 8), 1, 27),      # generic schedule, 64-channel chunks (8 planes: two per warp)
    (64, 32, (4, 16, 16), 2, 1),        # 1x1x1: box without halo
    (32, 1, (8, 16, 16), 2, 27),        # fp32 single-channel output (conv_out)
])
def test_conv3d_fused_groupnorm_silu_input(ci, co, grid, batch, taps):
    """conv(silu(groupnorm(x))) with the normalisation applied by the conv kernel to every halo tile in shared memory
    (ops.conv3d in_norm + ops.gn_coef) against (a) fp64 torch and (b) the two-pass path (ops.gn_silu, then ops.conv3d).
    Zero padding must stay zero AFTER the non-linearity: a transformed halo would show up as an O(1) error on every
    boundary voxel, so the maximum error is checked next to the relative L2."""
    import torch.nn.functional as F
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    d, h, w = grid
    k = 3 if taps == 27 else 1
    x = (torch.randn((batch, ci, d, h, w), generator=g) * 1.5 + 0.3).to(dev)
    wt = (torch.randn((co, ci, k, k, k), generator=g) / (taps * ci) ** 0.5).to(dev)
    gamma = (1.0 + 0.2 * torch.randn(ci, generator=g)).to(dev)
    beta = (0.2 * torch.randn(ci, generator=g)).to(dev)
    cadd = torch.randn((batch, co), generator=g).to(dev)
    xp = ops.to_planar(x)
    xr = ops.from_planar(xp, ci)                                   # the bf16 tensor both paths normalise
    stats = ops.channel_stats(xp, ci)
    coef = ops.gn_coef(stats, gamma, beta, 8, d * h * w)
    wp = ops.pack_conv_weight(wt)
    kw = dict(taps=ops.TAPS_3X3X3 if k == 3 else ops.TAPS_1X1X1, chan_add=cadd)
    if co == 1:
        kw["out_fp32"] = True
    fused = ops.conv3d(xp, wp, co, in_norm=coef, **kw)
    a = ops.gn_silu(xp, ci, 8, stats, gamma, beta)
    two_pass = ops.conv3d(a, wp, co, **kw)
    torch.cuda.synchronize()
    if co > 1:
        fused, two_pass = ops.from_planar(fused, co), ops.from_planar(two_pass, co)
    ref = (F.conv3d(_gn_silu_ref(xr, gamma, beta, 8), wt.double(), padding=k // 2) + cadd.double()[:, :, None, None, None]).float()
    scale = ref.abs().max().item()
    for name, got in (("fused", fused), ("two-pass", two_pass)):
        rel = ((got - ref).norm() / ref.norm()).item()
        mx = (got - ref).abs().max().item() / scale
        print(f"{name}: relative L2 {rel:.3e}, max abs / max |ref| {mx:.3e}")
        assert rel < 6e-3 and mx < 3e-2, (name, rel, mx)
    assert ((fused - two_pass).norm() / two_pass.norm()).item() < 4e-3


def test_conv3d_fused_input_with_skip_chunk_residual_and_stats():
    """The up block's net2 conv as the inference trunk launches it: raw h normalised in the kernel, the 1x1x1 skip conv
    fused as an UN-transformed extra channel chunk, coarse residual, statistics."""
    import torch.nn.functional as F
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(12)
    b, c, cs, (d, h, w) = 2, 32, 32, (8, 32, 16)
    x = torch.randn((b, c, d, h, w), generator=g).to(dev)
    xs = torch.randn((b, cs, d, h, w), generator=g).to(dev)
    rc = torch.randn((b, c, d // 2, h // 2, w // 2), generator=g).to(dev)
    wt = (torch.randn((c, c, 3, 3, 3), generator=g) / (27 * c) ** 0.5).to(dev)
    ws = (torch.randn((c, cs, 1, 1, 1), generator=g) / cs ** 0.5).to(dev)
    gamma, beta = (1.0 + 0.2 * torch.randn(c, generator=g)).to(dev), (0.2 * torch.randn(c, generator=g)).to(dev)
    cadd = torch.randn((b, c), generator=g).to(dev)
    xp, sp, rp = ops.to_planar(x), ops.to_planar(xs), ops.to_planar(rc)
    stats = ops.channel_stats(xp, c)
    coef = ops.gn_coef(stats, gamma, beta, 8, d * h * w)
    out_stats = torch.zeros((b, c, 2), dtype=torch.float64, device=dev)
    y = ops.conv3d(xp, ops.pack_conv_weight(wt), c, chan_add=cadd, residual=rp, residual_upsample=True, stats=out_stats,
                   skip_x=sp, skip_w=ops.pack_conv_weight(ws), in_norm=coef)
    torch.cuda.synchronize()
    got = ops.from_planar(y, c)
    ref = F.conv3d(_gn_silu_ref(ops.from_planar(xp, c), gamma, beta, 8), wt.double(), padding=1) + \
        F.conv3d(ops.from_planar(sp, cs).double(), ws.double()) + cadd.double()[:, :, None, None, None] + \
        F.interpolate(ops.from_planar(rp, c).double(), scale_factor=2, mode="nearest")
    rel = ((got - ref.float()).norm() / ref.float().norm()).item()
    mx = (got - ref.float()).abs().max().item() / ref.abs().max().item()
    print(f"fused input + skip chunk: relative L2 {rel:.3e}, max {mx:.3e}")
    assert rel < 6e-3 and mx < 3e-2, (rel, mx)
    assert torch.allclose(out_stats[..., 0], got.double().sum((2, 3, 4)), rtol=1e-5, atol=1e-2)


def test_conv3d_fused_input_is_declined_for_circular_padding():
    ops = _ops()
    x = torch.zeros((1, 2, 6, 18, 10, 8), dtype=torch.bfloat16, device="cuda")
    w = torch.zeros((27, 2, 16, 8), dtype=torch.bfloat16, device="cuda")
    coef = torch.zeros((1, 16, 2), dtype=torch.float32, device="cuda")
    with pytest.raises(ops.UnsupportedFusion):
        ops.conv3d(x, w, 16, circular=True, in_norm=coef)


# ---- polyphase up-conv: conv3x3x3(interpolate(a)) as eight 2x2x2-tap convolutions of the coarse tensor ------------------
@pytest.mark.parametrize("cc,co,cgrid,batch,n_par", [(32, 32, (4, 16, 8), 2, 1), (64, 32, (5, 10, 6), 1, 1), (128, 64, (4, 8, 8), 1, 1),
                                                      (64, 32, (5, 10, 6), 2, 4), (128, 64, (4, 16, 8), 1, 4),
                                                      (256, 128, (3, 8, 8), 1, 2), (32, 32, (4, 16, 8), 1, 8)])
def test_polyphase_upconv_is_exact_on_integers(cc, co, cgrid, batch, n_par):
    """The up blocks' conv over the up-sampled half of the concat (ResNetUp: interpolate -> cat -> ResNetBlock.net1) in the
    form the inference trunk runs it: per output parity a 2x2x2-tap conv of the COARSE tensor with summed filter taps
    (ops.polyphase_weight / polyphase_taps) into a parity-planar buffer, read back by the consumer conv through the
    depth-to-space residual (residual_upsample="d2s").  Integer inputs: every product and sum is exact, so the result
    must equal F.conv3d(F.interpolate(a)) bit for bit -- including the zero padding at the fine-grid boundary."""
    import torch.nn.functional as F
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(21)
    dc, hc, wc = cgrid
    a = _int_tensor((batch, cc, dc, hc, wc), -2, 2, g, dev)
    # sparse +-1 filter: every partial and final sum stays an integer bf16 holds exactly (|.| < 256)
    wt = _int_tensor((co, cc, 3, 3, 3), -1, 1, g, dev) * (_int_tensor((co, cc, 3, 3, 3), 0, 2 if cc < 256 else 11, g, dev) == 0).float()
    ap = ops.to_planar(a)
    part = torch.full((batch, co, dc, hc, wc, 8), 55.0, dtype=torch.bfloat16, device=dev)     # 8 x co channels
    if n_par == 1:                       # one launch per parity
        for pi in range(8):
            parity = (pi >> 2, (pi >> 1) & 1, pi & 1)
            we = ops.polyphase_weight(wt, parity)
            assert we.shape == (co, cc, 2, 2, 2)
            ops.conv3d(ap, ops.pack_conv_weight(we), co, taps=ops.polyphase_taps(parity), out=part, out_plane0=pi * (co // 8))
    else:                                # n_par parities per launch, stacked along N (what the inference trunk launches)
        for grp in range(8 // n_par):
            we, taps = ops.polyphase_group(wt, grp, n_par)
            assert we.shape[0] == n_par * co and len(taps) == we.shape[2] * we.shape[3] * we.shape[4]
            ops.conv3d(ap, ops.pack_conv_weight(we), n_par * co, taps=taps, out=part, out_plane0=grp * n_par * (co // 8))
    # consumer: a conv with zero weights, so its output is bias + the depth-to-space residual
    xs = torch.zeros((batch, 2, 2 * dc, 2 * hc, 2 * wc, 8), dtype=torch.bfloat16, device=dev)
    wz = torch.zeros((27, 2, co, 8), dtype=torch.bfloat16, device=dev)
    cadd = _int_tensor((batch, co), -3, 3, g, dev)
    y = ops.conv3d(xs, wz, co, chan_add=cadd, residual=part, residual_upsample="d2s")
    torch.cuda.synchronize()
    # (fp64 reference on the CPU: cuDNN picks an FFT algorithm for some fp64 shapes and returns 113.00000000000001)
    ref = (F.conv3d(F.interpolate(a.double().cpu(), scale_factor=2, mode="nearest"), wt.double().cpu(), padding=1) +
           cadd.double().cpu()[:, :, None, None, None]).to(dev)
    assert ref.abs().max() < 256                                     # integers bf16 holds exactly
    got = ops.from_planar(y, co).double()
    assert torch.equal(got, ref), f"max |diff| = {(got - ref).abs().max().item()}"
