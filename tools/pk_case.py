#!/usr/bin/env python
"""A few P(k) / r(k) / log-PDF calls on 16 fields of 128^3 (for an ncu launch list: which kernels make up vdm_pk)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from vdm4cdm_b200 import utils  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x = torch.rand((16, 1, n, n, n), device="cuda") + 0.5
y = torch.rand((16, 1, n, n, n), device="cuda") + 0.5
for _ in range(3):
    utils.pk(x)
    utils.get_ccs(x, y)
    utils.get_logpdf_3d(x * 1e10)
torch.cuda.synchronize()
print("done")
