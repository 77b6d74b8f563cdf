"""oracle/rng_oracle.py: Philox4x32-10 against the known-answer vectors of Random123 (kat_vectors, philox4x32 10 rounds),
and the uniform / discrete-Gaussian maps built on it (no GPU)."""
import numpy as np

from oracle import rng_oracle as R


def test_philox_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = R.philox4x32_10(*[np.array([c], dtype=np.uint64) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def test_uniform_is_in_range_reproducible_and_flat():
    moduli = [268440577, 8392193, 97]
    a = R.uniform(moduli, 3, 4096, seed=0x1234567890abcdef, stream=5)
    b = R.uniform(moduli, 3, 4096, seed=0x1234567890abcdef, stream=5)
    c = R.uniform(moduli, 3, 4096, seed=0x1234567890abcdef, stream=6)
    assert (a == b).all() and (a != c).any()
    for l, q in enumerate(moduli):
        assert a[:, l].min() >= 0 and a[:, l].max() < q
    small = a[:, 2].ravel()
    counts = np.bincount(small, minlength=97)
    expect = small.size / 97
    assert ((counts - expect) ** 2 / expect).sum() < 97 + 6 * (2 * 97) ** 0.5      # chi-square, 6 sigma


def test_gaussian_table_and_moments():
    sigma = 3.2
    cdt = R.gaussian_cdt(sigma)
    assert all(x <= y for x, y in zip(cdt, cdt[1:])) and cdt[-1] == (1 << 64) - 1 and len(cdt) <= 128
    x = R.gaussian(cdt, 8, 1 << 15, seed=99, stream=1)
    assert abs(x.mean()) < 0.05 and abs(x.var() - sigma * sigma) < 0.3
    assert (R.gaussian(cdt, 8, 1 << 15, seed=99, stream=1) == x).all()
