"""CPU: the oracle's Philox4x32-10 is pinned by the Random123 known-answer vectors (kat_vectors,
Salmon et al. SC'11), and the Box-Muller mapping by its moments."""
import numpy as np

from oracle import philox_ref

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        got = philox_ref.philox4x32_10(*[np.uint32(c) for c in ctr], *[np.uint32(k) for k in key])
        assert tuple(int(g) for g in got) == want


def test_normal_field_is_reproducible_and_standard_normal():
    a = philox_ref.normal_field(42, 3, 1, 200003)
    b = philox_ref.normal_field(42, 3, 1, 200003)
    c = philox_ref.normal_field(42, 4, 1, 200003)
    d = philox_ref.normal_field(42, 3, 2, 200003)
    assert a.dtype == np.float32 and np.array_equal(a, b)
    assert not np.array_equal(a, c) and not np.array_equal(a, d)
    assert np.all(np.isfinite(a))
    assert abs(a.mean()) < 0.01 and abs(a.std() - 1) < 0.01
    assert abs((a ** 3).mean()) < 0.03 and abs((a ** 4).mean() - 3) < 0.1
    # a prefix of a longer field is the shorter field: element e depends on (e, draw, realisation, seed) only
    assert np.array_equal(philox_ref.normal_field(42, 3, 1, 1001), a[:1001])
