"""GPU parity of the fused sampler update and its counter-based noise against the CPU oracle
(oracle/philox_ref.py, oracle/vdm_ref.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from vdm4cdm_b200 import ops
    return ops


@pytest.mark.parametrize("n", [4, 37, 4096, 32 ** 3 + 3])
def test_philox_normal_matches_oracle(n):
    from oracle import philox_ref
    ops = _ops()
    seed, draw = 0x1234_5678_9ABC_DEF0, 7
    rid = torch.tensor([0, 5, 1 << 20], dtype=torch.int32, device="cuda:0")
    got = ops.philox_normal((3, n), seed, draw, rid).cpu().numpy()
    for i, r in enumerate([0, 5, 1 << 20]):
        want = philox_ref.normal_field(seed, r, draw, n)
        # same Philox words bit for bit; logf/sincosf differ from numpy's by a few ulp
        np.testing.assert_allclose(got[i], want, rtol=0, atol=2e-6)
    # moments of a large draw
    if n > 30000:
        assert abs(got.mean()) < 0.01 and abs(got.std() - 1.0) < 0.01


def test_sampler_step_injected_noise_matches_oracle_formula():
    """z_s = alpha_s/alpha_t (z_t - c sigma_t eps_hat) + sigma_s sqrt(c) eps  (vdm_model.py:370-378)."""
    from oracle.vdm_ref import VDM
    ops = _ops()
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    shape = (2, 1, 6, 10, 14)
    z = torch.randn(shape, generator=g)
    eps = torch.randn(shape, generator=g)
    noise = torch.randn(shape, generator=g)

    class Net(torch.nn.Module):
        shape = (1, 6, 10, 14)

        def forward(self, zt, t=None, **kw):
            return eps

    vdm = VDM(Net())
    t, s = 0.62, 0.58
    want = vdm.sample_zs_given_zt(z, t, s, noise=noise)
    gt, gs = torch.tensor(-13.3 + 26.6 * t, dtype=torch.float64), torch.tensor(-13.3 + 26.6 * s, dtype=torch.float64)
    c = -torch.expm1(gs - gt)
    a_t, a_s = torch.sigmoid(-gt).sqrt(), torch.sigmoid(-gs).sqrt()
    s_t, s_s = torch.sigmoid(gt).sqrt(), torch.sigmoid(gs).sqrt()
    coef = torch.tensor([[a_s / a_t, -a_s / a_t * c * s_t, s_s * c.sqrt(), 1.0]], dtype=torch.float32, device=dev)
    got = ops.sampler_step(z.to(dev), eps.to(dev), coef, noise=noise.to(dev)).cpu()
    # fp32 tolerance of the north_star (1e-5): the oracle evaluates alpha/sigma in fp32, the coefficients here are fp64
    assert torch.allclose(got, want.detach(), rtol=1e-5, atol=1e-5), (got - want).abs().max()


def test_sampler_step_philox_noise_steps_and_packed_output():
    from oracle import philox_ref
    ops = _ops()
    dev = torch.device("cuda:0")
    b, d, h, w = 2, 4, 6, 10
    n = d * h * w
    g = torch.Generator().manual_seed(1)
    z = torch.randn((b, 1, d, h, w), generator=g).to(dev)
    eps = torch.randn((b, 1, d, h, w), generator=g).to(dev)
    cond = torch.randn((b, 1, d, h, w), generator=g).to(dev)
    coef = torch.tensor([[1.0, 0.0, 0.0, 1.0], [0.5, -0.25, 2.0, 1.5], [0.9, 0.1, 0.0, 1.0]], dtype=torch.float32,
                        device=dev)
    step = torch.tensor([1], dtype=torch.int32, device=dev)
    rid = torch.tensor([3, 11], dtype=torch.int32, device=dev)
    packed = ops.pack_input(z, cond, 16)
    seed = 2024
    out = ops.sampler_step(z, eps, coef, step_ptr=step, seed=seed, realisation_id=rid, draw_base=1, cond=cond,
                           packed_out=packed)
    torch.cuda.synchronize()
    for i, r in enumerate([3, 11]):
        nz = torch.from_numpy(philox_ref.normal_field(seed, r, 1 + 1, n)).reshape(1, d, h, w)
        want = 1.5 * (0.5 * z[i].cpu() - 0.25 * eps[i].cpu() + 2.0 * nz)
        assert torch.allclose(out[i].cpu(), want, rtol=1e-5, atol=1e-5), (out[i].cpu() - want).abs().max()
    dense = ops.from_planar(packed)
    assert torch.equal(dense[:, 0:1], out.to(torch.bfloat16).float()), "plane 0 must carry the new latent"
    assert torch.equal(dense[:, 1:2], cond.to(torch.bfloat16).float())
    assert dense[:, 2:].abs().sum().item() == 0
    # noise_scale == 0 (last step) must not touch the RNG path and stays deterministic
    step.fill_(2)
    o2 = ops.sampler_step(z, eps, coef, step_ptr=step, seed=seed)
    assert torch.allclose(o2, 0.9 * z + 0.1 * eps, rtol=1e-6, atol=1e-6)
    ops.increment(step)
    torch.cuda.synchronize()
    assert step.item() == 3


def test_ddnm_inpainting_matches_oracle():
    """get_ddnm_result (src/utils.py:277-304): masked-observation inpainting with time travel, injected noise."""
    from oracle.unet_ref import CUNet as RefNet
    from oracle.vdm_ref import LightVDM as RefLight, get_ddnm_result as ref_ddnm
    from vdm4cdm_b200 import utils
    from vdm4cdm_b200.networks import CUNet
    from vdm4cdm_b200.vdm_model import LightVDM
    torch.manual_seed(0)
    shape, chs, batch, n_steps = (1, 16, 16, 16), (16, 32), 2, 4
    kw = dict(shape=shape, chs=chs, s_conditioning_channels=1, v_conditioning_dims=[6], t_conditioning=True)
    ref_net = RefNet(**kw).eval()
    net = CUNet(**kw)
    net.load_state_dict(ref_net.state_dict())
    ref, mod = RefLight(ref_net).eval(), LightVDM(net).cuda().eval()
    g = torch.Generator().manual_seed(7)
    truth = torch.randn((batch,) + shape, generator=g)
    cond = torch.randn((batch,) + shape, generator=g)
    vals = [torch.rand(batch, 6, generator=g)]
    mask = (torch.rand((1,) + shape, generator=g) > 0.5).float()
    noises = [torch.randn((batch,) + shape, generator=g) for _ in range(64)]
    y = mask * truth
    want = ref_ddnm(ref, y, lambda x: mask * x, lambda x: mask * x, n_sampling_steps=n_steps, l=[0, 1, 2, 1],
                    return_all=True, noise_fn=lambda d, s: noises[d], s_conditioning=cond, v_conditionings=vals)
    mc = mask.cuda()
    got = utils.get_ddnm_result(mod, y.cuda(), lambda x: mc * x, lambda x: mc * x, n_sampling_steps=n_steps, l=[0, 1, 2, 1],
                                return_all=True, noise_fn=lambda d, s: noises[d].cuda(), s_conditioning=cond.cuda(),
                                v_conditionings=[vals[0].cuda()]).cpu()
    assert got.shape == want.shape == (n_steps, batch) + shape
    # the observed voxels are reproduced at every step (range-space projection), up to the fp32 rounding of
    # (A^T y + x0) - A^T A x0 (|x0| reaches 1e3 at t = 1 where alpha ~ 1e-3)
    for i in range(n_steps):
        tol = 4e-7 * max(1.0, got[i].abs().max().item()) * 4
        assert torch.allclose(got[i] * mask, y, atol=tol), (i, (got[i] * mask - y).abs().max().item(), tol)
    # ... and the inpainted part follows the oracle (errors compound over the 9 network calls of this schedule)
    err = ((got - want).norm() / want.norm()).item()
    print(f"ddnm: relative L2 vs oracle {err:.3e}")
    assert err < 5e-2, err
    with pytest.raises(AssertionError):
        utils.get_ddnm_result(mod, y.cuda(), lambda x: x, lambda x: x, n_sampling_steps=4, l=[1, 2])


def test_classifier_free_guidance_branch_matches_oracle():
    """VDM.get_pred_noise with w_cfg (vdm_model.py:318-327): (1 + w) eps(cond) - w eps(parameters masked out), and a
    guided chain samples through the generic (non-graph) loop."""
    import sys
    import os
    sys.path.insert(0, os.path.dirname(__file__))
    from test_gpu_unet import _models, _rel_l2
    from oracle.vdm_ref import VDM as RefVDM
    from vdm4cdm_b200.vdm_model import VDM
    shape, chs, batch = (1, 16, 16, 16), (16, 32), 2
    ref_net, net = _models(shape, chs)
    ref_vdm, vdm = RefVDM(ref_net, w_cfg=1.5).eval(), VDM(net, w_cfg=1.5).cuda().eval()
    g = torch.Generator().manual_seed(12)
    zt = torch.randn((batch,) + shape, generator=g)
    cond = torch.randn((batch,) + shape, generator=g)
    v = [torch.rand(batch, 6, generator=g)]
    gamma_t = torch.full((batch,), 2.0)
    with torch.no_grad():
        want = ref_vdm.get_pred_noise(zt, gamma_t, s_conditioning=cond, v_conditionings=v)
        plain = RefVDM(ref_net).eval().get_pred_noise(zt, gamma_t, s_conditioning=cond, v_conditionings=v)
        got = vdm.get_pred_noise(zt.cuda(), gamma_t.cuda(), s_conditioning=cond.cuda(), v_conditionings=[v[0].cuda()]).cpu()
    # guidance extrapolates the difference of two network outputs, so their bf16 noise is amplified by (1 + 2w)
    assert _rel_l2(got, want) < 4e-2, _rel_l2(got, want)
    assert _rel_l2(want, plain) > 1e-3                  # the guided output is a different function
    with pytest.raises(AssertionError, match="v_conditionings"):
        vdm.get_pred_noise(zt.cuda(), gamma_t.cuda(), s_conditioning=cond.cuda())
    xs = vdm.sample(batch, 3, "cuda:0", seed=3, s_conditioning=cond.cuda(), v_conditionings=[v[0].cuda()])
    assert xs.shape == (batch,) + shape and torch.isfinite(xs).all()
