"""``mltools.ml_utils``: the one helper the hot path's callers use (``to_np``: trainVDM3D128_...:103-113, calc_SS.py)."""
import numpy as np
import torch


def to_np(x):
    """torch tensor (any device, any grad state) -> numpy array; numpy arrays pass through."""
    if isinstance(x, np.ndarray):
        return x
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)
