import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.test_gpu_unet import _models, _rel_l2
from oracle.vdm_ref import VDM as RefVDM
from vdm4cdm_b200.vdm_model import VDM
shape, chs, batch, n_steps = (1, 16, 16, 16), (16, 32), 2, 6
ref_net, net = _models(shape, chs)
ref_vdm, vdm = RefVDM(ref_net).eval(), VDM(net).cuda().eval()
g = torch.Generator().manual_seed(3)
noises = [torch.randn((batch,) + shape, generator=g) for _ in range(n_steps + 1)]
cond = torch.randn((batch,) + shape, generator=g)
v = [torch.rand(batch, 6, generator=g)]
kw_ref = dict(s_conditioning=cond, v_conditionings=v)
kw = dict(s_conditioning=cond.cuda(), v_conditionings=[v[0].cuda()])
steps = torch.linspace(1.0, 0.0, n_steps + 1)
for mode in (False, True):
    vdm.use_cuda_graph = mode
    traj = vdm.sample(batch, n_steps, "cuda:0", return_all=True, noise_fn=lambda d, s: noises[d].cuda(), **kw).cpu()
    z = noises[0]
    with torch.no_grad():
        for i in range(n_steps):
            want = ref_vdm.sample_zs_given_zt(zt=z, t=steps[i], s=steps[i + 1], noise=noises[i + 1], **kw_ref)
            eps_ref = ref_vdm.get_pred_noise(z, ref_vdm._gamma5(steps[i], z).expand(batch, 1, 1, 1, 1), **kw_ref)
            print(f"graph={mode} step {i}: rel err {_rel_l2(traj[i], want):.3e}  |z|={z.norm():.3f} |want|={want.norm():.3f} |eps|={eps_ref.norm():.3f}")
            z = traj[i]
