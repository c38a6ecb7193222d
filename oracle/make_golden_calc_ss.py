"""Generate tests/golden/calc_ss_golden.npz by executing the REFERENCE's own summary-statistic functions of
``calc_SS.py`` (get_logpdf_3d / get_logpdf_2d :51-65, get_pk_3d / get_pk_2d :67-75).

``calc_SS.py`` is a script: importing it parses the command line and asserts that the author's data folder exists.  The
four function definitions are therefore located in its syntax tree and executed on their own, in a namespace that
provides what they use (``np``, ``torch`` and the reference's ``utils``).  Nothing is copied into this repository: the
functions run from /root/reference/calc_SS.py and only their outputs on seeded inputs are stored.

Run in the build container only:   python oracle/make_golden_calc_ss.py
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from oracle.make_golden import REF, load_reference_utils, seeded_field  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden", "calc_ss_golden.npz")
WANTED = ("get_logpdf_3d", "get_logpdf_2d", "get_pk_3d", "get_pk_2d")


def mass_field(seed, shape, mean=10.019, std=0.552):
    """Un-normalised Mcdm-like field (calc_SS.py:146 applies unnorm_func before the statistics)."""
    g = seeded_field(seed, shape).astype(np.float64)
    return (10.0 ** (g * std + mean) - 1.0).astype(np.float32)


def load_reference_functions():
    path = os.path.join(REF, "calc_SS.py")
    tree = ast.parse(open(path).read(), filename=path)
    defs = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in WANTED]
    assert sorted(d.name for d in defs) == sorted(WANTED)
    ns = {"np": np, "torch": torch, "utils": load_reference_utils()}
    exec(compile(ast.Module(body=defs, type_ignores=[]), path, "exec"), ns)
    return ns


def main():
    ref = load_reference_functions()
    fields = torch.from_numpy(mass_field(21, (3, 1, 32, 32, 32)))
    half = fields[:, :, :16].sum(2)                       # calc_SS.py:84: projected slab of half the box
    out = {"logpdf3d": ref["get_logpdf_3d"](fields), "logpdf2d": ref["get_logpdf_2d"](half),
           "pk3d": ref["get_pk_3d"](fields), "pk2d": ref["get_pk_2d"](half)}
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: (v.shape, v.dtype) for k, v in out.items()}, int(out["logpdf3d"].sum()), int(out["logpdf2d"].sum()))


if __name__ == "__main__":
    main()
