"""Oracle for the sampler's counter-based noise (TEST INFRASTRUCTURE).

Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11; the generator behind curand's Philox and torch's CUDA ``randn``) restated
in numpy, plus the Box-Muller mapping this repository defines for its fused
sampler kernel.  The reference itself sets no seed in generate_3D.py and calls
``torch.randn``, so there is no reference stream to
match: parity means the CUDA kernel and this restatement agree on the *same
definition*, which is:

  element e of realisation r, draw d (d = 0: initial latent z_1; d = i + 1: the
  noise consumed by reverse step i), seed s (64 bit):
      counter = (e // 4, d, r, 0x56444D34)        # 'VDM4'
      key     = (s & 0xffffffff, s >> 32)
      (w0, w1, w2, w3) = philox4x32_10(counter, key)
      u_j = ((w_j >> 8) + 0.5) * 2^-24            # exact in fp32, inside (0, 1)
      rad_a = sqrt(-2 ln u0), rad_b = sqrt(-2 ln u2)
      n = (rad_a sin(2 pi u1), rad_a cos(2 pi u1), rad_b sin(2 pi u3), rad_b cos(2 pi u3))
      noise[e] = n[e % 4]

The Philox core is pinned by the Random123 known-answer vectors in
``tests/test_oracle_philox.py``.
"""
from __future__ import annotations

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = np.uint32(0x9E3779B9)
W1 = np.uint32(0xBB67AE85)
STREAM_TAG = 0x56444D34


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. Inputs: uint32 arrays (broadcastable). Returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint32)
    c1 = np.asarray(c1, dtype=np.uint32)
    c2 = np.asarray(c2, dtype=np.uint32)
    c3 = np.asarray(c3, dtype=np.uint32)
    k0 = np.asarray(k0, dtype=np.uint32)
    k1 = np.asarray(k1, dtype=np.uint32)
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & mask).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & mask).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = (k0 + W0).astype(np.uint32)
            k1 = (k1 + W1).astype(np.uint32)
    return c0, c1, c2, c3


def _unit(w):
    return ((w >> np.uint32(8)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -24)


def normal_field(seed: int, realisation: int, draw: int, n_elements: int) -> np.ndarray:
    """fp32 N(0,1) noise for elements 0..n_elements-1 of one realisation (see module docstring)."""
    n_groups = (n_elements + 3) // 4
    g = np.arange(n_groups, dtype=np.uint64).astype(np.uint32)
    w0, w1, w2, w3 = philox4x32_10(g, np.uint32(draw), np.uint32(realisation), np.uint32(STREAM_TAG),
                                   np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF))
    u0, u1, u2, u3 = _unit(w0), _unit(w1), _unit(w2), _unit(w3)
    two_pi = np.float32(2.0 * np.pi)
    ra = np.sqrt(np.float32(-2.0) * np.log(u0)).astype(np.float32)
    rb = np.sqrt(np.float32(-2.0) * np.log(u2)).astype(np.float32)
    out = np.empty((n_groups, 4), dtype=np.float32)
    out[:, 0] = ra * np.sin(two_pi * u1)
    out[:, 1] = ra * np.cos(two_pi * u1)
    out[:, 2] = rb * np.sin(two_pi * u3)
    out[:, 3] = rb * np.cos(two_pi * u3)
    return out.reshape(-1)[:n_elements]
