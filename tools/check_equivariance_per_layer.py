"""Per conv call: is the marching network equivariant under a w-shift?  Compares every conv output / statistics of the
shifted run with the rolled output of the unshifted run (bring-up build)."""
import os, sys
os.environ["VDM4CDM_BRINGUP"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from vdm4cdm_b200 import _C, ops
from test_gpu_unet import _models
lib = _C.lib()
real = ops.conv3d
log = []
def rec(x, w_packed, c_out, **kw):
    y = real(x, w_packed, c_out, **kw)
    torch.cuda.synchronize()
    st = kw.get("stats")
    log.append((tuple(x.shape), tuple(w_packed.shape), c_out, y.float().clone(), None if st is None else st.clone(), kw.get("skip_x") is not None))
    return y
ops.conv3d = rec
shape, chs, batch = (1, 16, 32, 16), (16, 32, 64), 2
ref, net = _models(shape, chs, padding="circular")
g = torch.Generator().manual_seed(2)
x = torch.randn((batch,) + shape, generator=g)
cond = 0.7 * x + 0.3 * torch.randn((batch,) + shape, generator=g)
t, v = torch.rand(batch, generator=g), [torch.rand(batch, 6, generator=g)]
SH = 4
for nm in (1, 0):
    lib.vdm_debug_set(7, nm)
    runs = []
    for s in (0, SH):
        log.clear()
        with torch.no_grad():
            net(torch.roll(x, s, 4).cuda(), t=t.cuda(), s_conditioning=torch.roll(cond, s, 4).cuda(), v_conditionings=[v[0].cuda()])
        runs.append(list(log))
    print("=== no_march", nm)
    for k, (a, b) in enumerate(zip(*runs)):
        ya, yb = a[3], b[3]
        lvl = shape[3] // ya.shape[-2] if ya.dim() == 6 else 1
        wdim = 4 if ya.dim() == 6 else 4
        sh = SH // lvl
        ra = torch.roll(ya, sh, wdim)
        diff = (ra - yb).abs()
        nbad = int((diff > 0).sum())
        ds = 0.0 if a[4] is None else float(((a[4] - b[4]).abs() / a[4].abs().clamp_min(1e-20)).max())
        print(f"call {k:2d} x={a[0]} w={a[1]} c_out={a[2]} skip={a[5]}: max|diff| {float(diff.max()):.3g} differing {nbad}/{diff.numel()}  stats rel diff {ds:.3g}", flush=True)
