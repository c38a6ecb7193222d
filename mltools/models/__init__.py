from . import sfm_model, vdm_model  # noqa: F401
