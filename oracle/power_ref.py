"""Oracle restatement of the reference's P(k) / r(k) code (TEST INFRASTRUCTURE).

Follows /root/reference/src/utils.py:
  * ``power``   -- src/utils.py:16-83
  * ``pk``      -- src/utils.py:85-102
  * ``get_ccs`` -- src/utils.py:110-128

The restatement is written with numpy (complex64 FFT semantics emulated by
casting) so that it is independent of both the reference code and of the CUDA
product.  It is pinned against the reference's own functions by
``oracle/make_golden.py`` -> ``tests/golden/power_*.npz`` and against the
analytic known answers of SURVEY.md section 4 in ``tests/test_oracle_power.py``.

Conventions restated (each one matters for bit-level agreement of the bins):
  * transform: real-to-complex FFT over all spatial axes, unnormalised.
  * spectrum : X * conj(X2); mean over batch axis 0, then sum over channel axis.
  * wave-number of a mode: integer frequency per axis, with the wrap
    ``j -> j - n`` applied when ``j > n // 2`` on every full axis (so the Nyquist
    index n/2 stays positive); the last (half) axis is 0..n/2.  |k| is the fp32
    Euclidean norm.
  * Hermitian weight: 2 for every mode, except 1 on the last-axis plane 0 and,
    when the last axis is even, on its Nyquist plane.
  * bin index: ceil(|k|) as int32; bins 1..kmax are returned, kmax = min(n)//2.
  * outputs: (sum w k / sum w, sum w P / sum w, sum w) -- fp32, fp32, int32.
"""
from __future__ import annotations

import numpy as np


def _axis_freqs(n_modes: int, full_axis: bool) -> np.ndarray:
    j = np.arange(n_modes, dtype=np.float32)
    if full_axis:
        j = j - n_modes * (j > (n_modes // 2))
    return j.astype(np.float32)


def power(x: np.ndarray, x2: np.ndarray | None = None):
    """(k_mean, P_mean, N_modes) of a (B, C, *spatial) field; src/utils.py:16-83."""
    x = np.asarray(x, dtype=np.float32)
    nd = x.ndim - 2
    spatial = x.shape[2:]
    axes = tuple(range(2, 2 + nd))
    kmax = min(spatial) // 2
    last_even = spatial[-1] % 2 == 0

    fx = np.fft.rfftn(x.astype(np.float64), s=spatial, axes=axes).astype(np.complex64)
    if x2 is None:
        fx2 = fx
    else:
        x2 = np.asarray(x2, dtype=np.float32)
        fx2 = np.fft.rfftn(x2.astype(np.float64), s=spatial, axes=axes).astype(np.complex64)
    spec = (fx * np.conj(fx2)).astype(np.complex64)
    spec = spec.mean(axis=0).sum(axis=0)          # batch mean, channel sum
    spec = spec.real.astype(np.float32)

    mode_shape = spec.shape
    freqs = [_axis_freqs(n, full_axis=(a < nd - 1)) for a, n in enumerate(mode_shape)]
    grids = np.meshgrid(*freqs, indexing="ij")
    k2 = np.zeros(mode_shape, dtype=np.float32)
    for g in grids:
        k2 = k2 + (g * g).astype(np.float32)
    kmag = np.sqrt(k2).astype(np.float32)

    w = np.full(mode_shape, 2, dtype=np.int32)
    w[..., 0] = 1
    if last_even:
        w[..., -1] = 1

    kflat = kmag.reshape(-1)
    pflat = spec.reshape(-1)
    wflat = w.reshape(-1)
    kbin = np.ceil(kflat).astype(np.int32)
    nb = int(kbin.max()) + 1
    ksum = np.bincount(kbin, weights=(kflat * wflat).astype(np.float64), minlength=nb)
    psum = np.bincount(kbin, weights=(pflat.astype(np.float64) * wflat), minlength=nb)
    nsum = np.bincount(kbin, weights=wflat.astype(np.float64), minlength=nb)

    sl = slice(1, 1 + kmax)
    n_out = np.rint(nsum[sl]).astype(np.int32)
    k_out = (ksum[sl] / n_out).astype(np.float32)
    p_out = (psum[sl] / n_out).astype(np.float32)
    return k_out, p_out, n_out


def pk(fields: np.ndarray, fields2: np.ndarray | None = None):
    """Per-sample spectra, stacked on axis 0; src/utils.py:85-102."""
    ks, ps, ns = [], [], []
    for i in range(len(fields)):
        f2 = None if fields2 is None else fields2[i][None]
        k, p, n = power(fields[i][None], f2)
        ks.append(k)
        ps.append(p)
        ns.append(n)
    return np.stack(ks), np.stack(ps), np.stack(ns)


def get_ccs(fields1: np.ndarray, fields2: np.ndarray, full: bool = False):
    """Cross-correlation coefficient r(k) = P12 / sqrt(P11 P22); src/utils.py:110-128."""
    ks, p11, _ = pk(fields1)
    p22 = pk(fields2)[1]
    if full:
        n2 = len(fields2)
        rows = []
        for f1 in fields1:
            rep = np.repeat(f1[None], n2, axis=0)
            rows.append(pk(rep, fields2)[1])
        p12 = np.stack(rows)
        return ks, p12 / np.sqrt(p11[:, None] * p22[None, :])
    assert len(fields1) == len(fields2)
    p12 = pk(fields1, fields2)[1]
    return ks, p12 / np.sqrt(p11 * p22)
