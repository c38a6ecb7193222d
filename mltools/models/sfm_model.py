"""``mltools.models.sfm_model`` as the reference imports it (trainSFM3D160_c_c_from_field_name_thick_lowbatch.py:10)."""
from vdm4cdm_b200.sfm_model import LightSFM  # noqa: F401
