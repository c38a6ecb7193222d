"""``mltools.networks.networks`` as the reference imports it (trainVDM3D128_...:11, src/utils.py:448)."""
from vdm4cdm_b200.networks import CUNet, ResNetBlock, ResNetDown, ResNetUp, timestep_embedding  # noqa: F401
