"""``mltools.utils.cuda_tools.get_freer_device`` (generate_3D.py:31-32, calc_SS.py)."""
import torch


def get_freer_device(verbose: bool = False):
    """The CUDA device with the most free memory.  Under torchrun (LOCAL_RANK set) a rank keeps its own GPU.
    There is no CPU fallback in this package: without a GPU this raises."""
    import os
    if not torch.cuda.is_available():
        raise RuntimeError("vdm4cdm_b200 needs a CUDA device (B200); none is visible")
    if "LOCAL_RANK" in os.environ:
        return torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    free = [torch.cuda.mem_get_info(i)[0] for i in range(torch.cuda.device_count())]
    best = max(range(len(free)), key=lambda i: free[i])
    if verbose:
        print(f"Selected cuda:{best} ({free[best] / 2 ** 30:.1f} GiB free)")
    return torch.device("cuda", best)
