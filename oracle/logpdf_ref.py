"""Oracle for the log-PDF statistics (TEST INFRASTRUCTURE): calc_SS.py:51-65 restated line for line with numpy /
torch on the CPU (the reference's own arithmetic: fp32 ``torch.log10(fields + 1)``, ``np.histogram`` against
``np.linspace`` edges).  The two functions differ only in their bin range."""
import numpy as np
import torch


def get_logpdf(fields, lo, hi, n_edges=100):
    bins = np.linspace(lo, hi, n_edges)
    logfields = torch.log10(torch.as_tensor(fields) + 1).detach().cpu().numpy()
    pdfs = []
    for i in range(logfields.shape[0]):
        pdfs.append(np.histogram(logfields[i].flatten(), bins=bins)[0])
    return np.array(pdfs)


def get_logpdf_3d(fields):
    return get_logpdf(fields, 8.5, 15)


def get_logpdf_2d(fields):
    return get_logpdf(fields, 10.5, 15.5)
