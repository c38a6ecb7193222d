from . import cuda_tools  # noqa: F401
