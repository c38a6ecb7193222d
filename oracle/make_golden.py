"""Generate tests/golden/power_*.npz by running the REFERENCE's own P(k) code.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):   python oracle/make_golden.py

The reference module ``src/utils.py`` imports ``matplotlib.pyplot`` and
``mltools.ml_utils`` at top level; neither is installed here and neither is used
by ``power`` / ``pk`` / ``get_ccs`` (src/utils.py:16-128), so they are stubbed in
``sys.modules`` before the file is loaded with importlib.  Nothing from the
reference is copied into this repository: only its numerical outputs on seeded
inputs are stored.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def load_reference_utils():
    for name in ("matplotlib", "matplotlib.pyplot", "mltools", "mltools.ml_utils"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["mltools"].ml_utils = sys.modules["mltools.ml_utils"]
    spec = importlib.util.spec_from_file_location("ref_utils", os.path.join(REF, "src", "utils.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def seeded_field(seed: int, shape) -> np.ndarray:
    """The input definition shared by the generator and the tests."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal(size=shape, dtype=np.float32)


def mass_field(seed: int, shape) -> np.ndarray:
    """calc_SS.py-style input: unnormalised Mcdm-like field divided by its sum
    (calc_SS.py:67-70; constants from src/dataset/normalizations_3d.json, Mcdm)."""
    g = seeded_field(seed, shape).astype(np.float64)
    m = 10.0 ** (g * 0.552 + 10.019) - 1.0
    m = m / m.sum(axis=tuple(range(2, m.ndim)), keepdims=True)
    return m.astype(np.float32)


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    ref = load_reference_utils()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))

    cases = {
        # name: (kind, seed, shape)
        "g8": ("gauss", 1, (1, 1, 8, 8, 8)),
        "g9_odd": ("gauss", 2, (1, 1, 9, 9, 9)),
        "g16": ("gauss", 3, (1, 1, 16, 16, 16)),
        "g16_bc": ("gauss", 4, (3, 2, 8, 8, 8)),
        "g32": ("gauss", 5, (2, 1, 32, 32, 32)),
        "m64": ("mass", 6, (2, 1, 64, 64, 64)),
        "m128": ("mass", 7, (1, 1, 128, 128, 128)),
        "g2d_64": ("gauss", 8, (2, 1, 64, 64)),
        "m2d_128": ("mass", 9, (2, 1, 128, 128)),
        "g_aniso": ("gauss", 10, (1, 1, 16, 12, 20)),
    }
    out = {}
    for name, (kind, seed, shape) in cases.items():
        x = seeded_field(seed, shape) if kind == "gauss" else mass_field(seed, shape)
        y = seeded_field(seed + 100, shape) if kind == "gauss" else mass_field(seed + 100, shape)
        xt, yt = torch.from_numpy(x), torch.from_numpy(y)
        k, p, n = ref.power(xt)
        kc, pc, nc = ref.power(xt, yt)
        kb, pb, nb = ref.pk(xt)
        kcc, cc = ref.get_ccs(xt, yt, full=False)
        out[f"{name}.kind"] = np.array(kind)
        out[f"{name}.seed"] = np.array(seed)
        out[f"{name}.shape"] = np.array(shape)
        out[f"{name}.xdigest"] = np.array(digest(x))
        out[f"{name}.ydigest"] = np.array(digest(y))
        out[f"{name}.power_k"] = k.numpy()
        out[f"{name}.power_p"] = p.numpy()
        out[f"{name}.power_n"] = n.numpy()
        out[f"{name}.cross_p"] = pc.numpy()
        out[f"{name}.pk_k"] = kb.numpy()
        out[f"{name}.pk_p"] = pb.numpy()
        out[f"{name}.pk_n"] = nb.numpy()
        out[f"{name}.ccs_k"] = kcc.numpy()
        out[f"{name}.ccs"] = cc.numpy()
        print(name, shape, "bins", len(k), "sumN", int(n.sum()))
    # full=True pair matrix.  The reference's full branch only runs for 2-D fields
    # (src/utils.py:120 repeats with four factors, which raises for (C,D,H,W) inputs).
    a = torch.from_numpy(seeded_field(21, (3, 1, 16, 16)))
    b = torch.from_numpy(seeded_field(22, (3, 1, 16, 16)))
    kf, ccf = ref.get_ccs(a, b, full=True)
    out["full16.ccs"] = ccf.numpy()
    out["full16.k"] = kf.numpy()

    # analytic known-answer inputs, evaluated by the reference
    d8 = torch.zeros(1, 1, 8, 8, 8); d8[0, 0, 0, 0, 0] = 1.0
    k, p, n = ref.power(d8)
    out["delta8.k"], out["delta8.p"], out["delta8.n"] = k.numpy(), p.numpy(), n.numpy()
    d16 = torch.zeros(1, 1, 16, 16, 16); d16[0, 0, 0, 0, 0] = 1.0
    k, p, n = ref.power(d16)
    out["delta16.k"], out["delta16.p"], out["delta16.n"] = k.numpy(), p.numpy(), n.numpy()
    i = torch.arange(16, dtype=torch.float32)
    c = torch.cos(2 * np.pi * 3 * i / 16)
    for ax in range(3):
        shp = [1, 1, 1, 1, 1]
        shp[2 + ax] = 16
        f = c.reshape(shp).expand(1, 1, 16, 16, 16).contiguous()
        k, p, n = ref.power(f)
        out[f"cos16_ax{ax}.p"] = p.numpy()
        out[f"cos16_ax{ax}.n"] = n.numpy()

    np.savez_compressed(os.path.join(OUT, "power_golden.npz"), **out)
    print("wrote", os.path.join(OUT, "power_golden.npz"))


if __name__ == "__main__":
    main()
