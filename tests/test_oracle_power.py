"""CPU: the oracle restatement of P(k)/r(k) (oracle/power_ref.py) is pinned against outputs of the
REFERENCE's own src/utils.py:16-128 (tests/golden/power_golden.npz, made by oracle/make_golden.py) and
against the analytic known answers of SURVEY.md section 4."""
import os

import numpy as np
import pytest

from oracle import power_ref
from oracle.make_golden import digest, mass_field, seeded_field

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "power_golden.npz")
Z = np.load(GOLDEN)
CASES = sorted({k.split(".")[0] for k in Z.files if k.endswith(".kind")})


def _inputs(name):
    kind, seed = str(Z[f"{name}.kind"]), int(Z[f"{name}.seed"])
    shape = tuple(int(v) for v in Z[f"{name}.shape"])
    f = seeded_field if kind == "gauss" else mass_field
    x, y = f(seed, shape), f(seed + 100, shape)
    assert digest(x) == str(Z[f"{name}.xdigest"]) and digest(y) == str(Z[f"{name}.ydigest"]), "input definition drifted"
    return x, y


@pytest.mark.parametrize("name", CASES)
def test_power_matches_reference(name):
    x, y = _inputs(name)
    k, p, n = power_ref.power(x)
    np.testing.assert_array_equal(n, Z[f"{name}.power_n"])
    # k-mean: the reference sums up to ~5e4 fp32 terms per bin in fp32 (torch.bincount), which costs it up to
    # 3e-5 relative at 128^3; the oracle (and the CUDA path) accumulate in fp64 and are the exact values.
    np.testing.assert_allclose(k, Z[f"{name}.power_k"], rtol=5e-5)
    np.testing.assert_allclose(p, Z[f"{name}.power_p"], rtol=1e-5)
    _, pc, _ = power_ref.power(x, y)
    scale = np.sqrt(np.abs(p * power_ref.power(y)[1]))
    np.testing.assert_allclose(pc / scale, Z[f"{name}.cross_p"] / scale, atol=2e-5)


@pytest.mark.parametrize("name", [c for c in CASES if c != "m128"])
def test_pk_and_ccs_match_reference(name):
    x, y = _inputs(name)
    kb, pb, nb = power_ref.pk(x[:, None]) if False else power_ref.pk(x)
    np.testing.assert_allclose(pb, Z[f"{name}.pk_p"], rtol=1e-5)
    np.testing.assert_array_equal(nb, Z[f"{name}.pk_n"])
    _, cc = power_ref.get_ccs(x, y)
    np.testing.assert_allclose(cc, Z[f"{name}.ccs"], atol=2e-5)


def test_ccs_full_matrix_matches_reference():
    a, b = seeded_field(21, (3, 1, 16, 16)), seeded_field(22, (3, 1, 16, 16))
    _, cc = power_ref.get_ccs(a, b, full=True)
    np.testing.assert_allclose(cc, Z["full16.ccs"], atol=2e-5)


def test_known_answers():
    d8 = np.zeros((1, 1, 8, 8, 8), np.float32); d8[0, 0, 0, 0, 0] = 1
    k, p, n = power_ref.power(d8)
    np.testing.assert_array_equal(n, [6, 26, 90, 131])
    np.testing.assert_allclose(p, 1.0, rtol=1e-6)
    np.testing.assert_allclose(k, Z["delta8.k"], rtol=1e-6)
    d16 = np.zeros((1, 1, 16, 16, 16), np.float32); d16[0, 0, 0, 0, 0] = 1
    np.testing.assert_array_equal(power_ref.power(d16)[2], [6, 26, 90, 134, 258, 410, 494, 687])
    i = np.arange(16, dtype=np.float32)
    c = np.cos(2 * np.pi * 3 * i / 16).astype(np.float32)
    for ax in range(3):
        shp = [1, 1, 1, 1, 1]
        shp[2 + ax] = 16
        f = np.broadcast_to(c.reshape(shp), (1, 1, 16, 16, 16)).copy()
        k, p, n = power_ref.power(f)
        np.testing.assert_allclose(p[2] * n[2], 8388608.0, rtol=1e-5)
    a = seeded_field(3, (2, 1, 16, 16, 16))
    np.testing.assert_allclose(power_ref.get_ccs(a, a)[1], 1.0, rtol=1e-6)
    np.testing.assert_allclose(power_ref.get_ccs(a, -a)[1], -1.0, rtol=1e-6)
    k9, p9, n9 = power_ref.power(seeded_field(2, (1, 1, 9, 9, 9)))
    np.testing.assert_array_equal(n9, [6, 26, 90, 134])
