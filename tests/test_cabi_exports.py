"""CPU: the C-ABI library loads without a GPU and exports every symbol include/vdm4cdm_b200.h declares;
compute entry points fail loudly (no CPU fallback) when there is no device."""
import ctypes

import pytest
import torch

from vdm4cdm_b200 import _C


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_C.LIB_PATH)
    names = _C.declared_symbols()
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(_C._SIGNATURES) == names, "ctypes signature table and header disagree"


def test_version_and_error_string():
    lib = _C.lib()
    assert lib.vdm_version() >= 100
    assert isinstance(lib.vdm_last_error_string(), bytes)
    # the bring-up knobs ("results are wrong by construction") are not reachable from the release library
    assert not hasattr(lib, "vdm_debug_set")


def test_bad_arguments_are_rejected_without_touching_a_device():
    lib = _C.lib()
    rc = lib.vdm_conv3d(None, None, None, None, None, None)
    assert rc == -1 and b"NULL" in lib.vdm_last_error_string()
    assert lib.vdm_pk_work_bytes(1, 1, 1, 128, 128, 128, 0) >= 128 * 128 * 65 * 8
    assert lib.vdm_pk_work_bytes(0, 1, 1, 128, 128, 128, 0) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from vdm4cdm_b200 import ops
    x = torch.zeros((1, 2, 4, 16, 8, 8), dtype=torch.bfloat16)
    w = torch.zeros((27, 2, 16, 8), dtype=torch.bfloat16)
    with pytest.raises((ValueError, RuntimeError)):
        ops.conv3d(x, w, 16)
    assert _C.lib().vdm_device_supported(0) != 0
