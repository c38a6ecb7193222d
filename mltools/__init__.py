"""Compatibility namespace: the reference imports its model code by module path from the third-party package
``mltools`` (trainVDM3D128_c_c_from_field_name_thick_lowbatch.py:10-12, src/utils.py:448-449, generate_3D.py:31).
Putting this repository's root on ``sys.path`` makes those imports resolve to the B200-native implementations
in ``vdm4cdm_b200`` without touching the call sites:

    mltools.networks.networks.CUNet          -> vdm4cdm_b200.networks.CUNet
    mltools.models.vdm_model.LightVDM / VDM  -> vdm4cdm_b200.vdm_model
    mltools.models.sfm_model.LightSFM        -> vdm4cdm_b200.sfm_model
    mltools.ml_utils.to_np                   -> tensor -> numpy helper
    mltools.utils.cuda_tools.get_freer_device
"""
