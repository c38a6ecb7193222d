"""Shared pieces of the entry-point scripts: repository root on sys.path (so that ``mltools`` and ``vdm4cdm_b200``
resolve), torch.distributed bring-up, and the synthetic stand-in for the CAMELS data module.

The reference's ``AstroDataModule`` (src/dataset/CAMELS_3D_dataset.py) reads files that live on the author's
cluster; it is the "next" row of SURVEY.md section 8f.  Until it is built, every script accepts ``--synthetic``:
Gaussian random fields of the configured grid, a correlated conditioning field and CAMELS-range parameters
(SURVEY.md section 8d), in the reference's batch schema {"x", "conditioning", "conditioning_values": [params]}.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

PARAM_LO = [0.1, 0.6, 0.25, 0.25, 0.5, 0.5]
PARAM_HI = [0.5, 1.0, 4.0, 4.0, 2.0, 2.0]
from vdm4cdm_b200.dataset import ALPHAS_3D, NORMALIZATIONS_3D, unnorm_func  # noqa: E402


def init_distributed():
    """(rank, world, device).  One process per GPU under torchrun; a single process otherwise."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("vdm4cdm_b200 needs a CUDA device (B200); none is visible")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return rank, world, dev


def synthetic_batch(batch, grid, seed, n_params=6, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((batch, 1, grid, grid, grid), generator=g)
    cond = 0.7 * x + 0.3 * torch.randn((batch, 1, grid, grid, grid), generator=g)
    out = {"x": x.to(device), "conditioning": cond.to(device), "conditioning_values": []}
    if n_params:
        lo, hi = torch.tensor(PARAM_LO[:n_params]), torch.tensor(PARAM_HI[:n_params])
        out["conditioning_values"] = [(lo + (hi - lo) * torch.rand((batch, n_params), generator=g)).to(device)]
    return out


def unnorm_mcdm(x):
    """Normalised log field -> mass field: the consumer-side ``dm.unnorm_func(samples, i_channel=1)`` of calc_SS.py:146
    with the SAME constants the data were normalised with (``vdm4cdm_b200.dataset.NORMALIZATIONS_3D["Mcdm"]``, the
    reference's normalizations_3d.json) -- the one source of truth, not a rounded copy."""
    mean, std = NORMALIZATIONS_3D["Mcdm"]
    return unnorm_func(x, ALPHAS_3D["Mcdm"], mean, std)
