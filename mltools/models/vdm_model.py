"""``mltools.models.vdm_model`` as the reference imports it (trainVDM3D128_...:10, src/utils.py:449)."""
from vdm4cdm_b200.vdm_model import VDM, FixedLinearSchedule, LearnedLinearSchedule, LightVDM, SamplerSession  # noqa: F401
