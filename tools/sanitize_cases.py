#!/usr/bin/env python
"""Small invocations of every hand-written kernel family, for compute-sanitizer (tools/sanitize.sh): the conv kernel on
each schedule (kd-folded resident / streamed / generic, fused skip chunk, residual ring, statistics), the wgrad kernel
(narrow and generic), GroupNorm+SiLU forward/backward, pooling, the sampler update and the P(k) binning.  Results are
checked against torch so that a sanitizer-clean run is also a correct run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from vdm4cdm_b200 import ops, utils  # noqa: E402


def ints(shape, lo, hi, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(lo, hi + 1, shape, generator=g, device="cuda").float()


def conv_case(ci, co, grid, batch=1, residual=False, stats=False, taps=27):
    d, h, w = grid
    x = ints((batch, ci, d, h, w), -1, 1, 1)
    k = 3 if taps == 27 else 1
    wt = ints((co, ci, k, k, k), -1, 1, 2) * (ints((co, ci, k, k, k), 0, 3, 3) == 0).float()
    res = ints((batch, co, d, h, w), -2, 2, 4) if residual else None
    st = torch.zeros((batch, co, 2), dtype=torch.float64, device="cuda") if stats else None
    y = ops.conv3d(ops.to_planar(x), ops.pack_conv_weight(wt), co, taps=ops.TAPS_3X3X3 if taps == 27 else ops.TAPS_1X1X1,
                   residual=None if res is None else ops.to_planar(res), stats=st)
    want = F.conv3d(x.double(), wt.double(), padding=k // 2)
    if res is not None:
        want = want + res.double()
    got = ops.from_planar(y, co).double()
    assert torch.equal(got, want), f"conv {ci}->{co} {grid} taps={taps}: max diff {(got - want).abs().max().item()}"
    if st is not None:
        assert torch.allclose(st[..., 0], want.sum((2, 3, 4)), atol=1e-6)
    print(f"conv {ci}->{co} grid={grid} B={batch} taps={taps} residual={residual} stats={stats}: exact", flush=True)


def wgrad_case(ci, co, grid, batch=1):
    d, h, w = grid
    a, g = ints((batch, ci, d, h, w), -1, 1, 5), ints((batch, co, d, h, w), -1, 1, 6)
    dw = ops.conv3d_wgrad(ops.to_planar(a), ops.to_planar(g, 16), ci, co, 3)
    ap = F.pad(a, (1, 1, 1, 1, 1, 1)).double()
    want = torch.stack([torch.einsum("bcdhw,bodhw->co", ap[:, :, kd:kd + d, kh:kh + h, kw:kw + w], g.double())
                        for kd in range(3) for kh in range(3) for kw in range(3)]).float()
    assert torch.equal(dw, want), f"wgrad {ci}->{co}: max diff {(dw - want).abs().max().item()}"
    print(f"wgrad {ci}->{co} grid={grid} B={batch}: exact", flush=True)


def main():
    conv_case(32, 32, (8, 16, 16), residual=True, stats=True)        # kd-folded, resident weights, residual ring
    conv_case(16, 32, (6, 16, 8), stats=True)
    conv_case(64, 64, (8, 16, 16), stats=True)                       # kd-folded, streamed weights (N = 192)
    conv_case(128, 128, (4, 16, 8), batch=2, residual=True)          # generic schedule, n_split
    conv_case(64, 32, (4, 16, 8), taps=1, stats=True)                # 1x1x1
    conv_case(32, 16, (5, 20, 12))                                   # ragged tiles
    wgrad_case(32, 32, (8, 16, 16))                                  # narrow (d-slices folded into M)
    wgrad_case(128, 64, (4, 16, 8), batch=2)                         # generic
    # elementwise + sampler + P(k): one tiny UNet forward/backward and a spectrum
    from vdm4cdm_b200.networks import CUNet
    from vdm4cdm_b200.vdm_model import LightVDM
    torch.manual_seed(0)
    net = CUNet(shape=(1, 16, 16, 16), chs=(16, 32), s_conditioning_channels=1, v_conditioning_dims=[6], t_conditioning=True)
    model = LightVDM(net).cuda().train()
    x = torch.randn((2, 1, 16, 16, 16), device="cuda")
    batch = {"x": x, "conditioning": 0.7 * x, "conditioning_values": [torch.rand((2, 6), device="cuda")]}
    loss = model.training_step(batch)
    loss.backward()
    assert torch.isfinite(loss)
    model.eval()
    model.model.use_cuda_graph = False            # the sanitizer instruments eager launches
    xs = model.draw_samples(batch_size=2, n_sampling_steps=2, s_conditioning=batch["conditioning"],
                            v_conditionings=batch["conditioning_values"], seed=1)
    assert torch.isfinite(xs).all()
    k, p, n = utils.pk(torch.randn((2, 1, 16, 16, 16), device="cuda"))
    assert torch.isfinite(p).all()
    torch.cuda.synchronize()
    print("training step, 2-step sample, P(k): finite", flush=True)


if __name__ == "__main__":
    main()
