"""vdm4cdm_b200: B200-native (sm_100a) implementation of the vdm4cdm hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); the arithmetic runs in
hand-written CUDA behind the C ABI of ``include/vdm4cdm_b200.h`` (``libvdm4cdm_b200.so``).  There is
no CPU or PyTorch fallback: importing works anywhere, computing needs the library and a B200.
"""
__version__ = "0.1.0"
