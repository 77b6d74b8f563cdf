"""Tier-0: the CPU oracle against the mathematics (no GPU).

The reference has no golden vectors for this path, so the oracle is pinned
to convention-independent facts: CRT rows are evaluations at the primitive m-th roots,
crt is a ring homomorphism from Z_q[X]/Phi_m(X), crtInv inverts it, G and L satisfy their
defining relations, twace/embed commute with CRT.
"""
import numpy as np
import pytest

from oracle import oracle as O

CASES = [(8, 17), (16, 97), (9, 19), (27, 109), (7, 29), (15, 31), (45, 181), (32, 193),
         (21, 43), (63, 127), (25, 101), (36, 37), (12, 13), (100, 101), (105, 211), (40, 41)]


def cyclotomic(m, cache={}):
    """Phi_m(X) over Z, coefficients low to high"""
    if m in cache:
        return cache[m]
    num = np.zeros(m + 1, dtype=object)
    num[0], num[m] = -1, 1
    for d in range(1, m):
        if m % d == 0:
            den = cyclotomic(d)
            quo = np.zeros(len(num) - len(den) + 1, dtype=object)
            rem = num.copy()
            for k in range(len(quo) - 1, -1, -1):
                quo[k] = rem[k + len(den) - 1] // den[-1]
                rem[k:k + len(den)] -= quo[k] * den
            assert not rem.any()
            num = quo
    cache[m] = num
    return num


def poly_of_pow(m, q, x):
    """pow-basis coefficient vector -> polynomial in X=zeta_m of degree < m"""
    e = O.pow_exponents(m)
    a = np.zeros(m, dtype=object)
    for j, k in enumerate(e):
        a[k] = (a[k] + int(x[j])) % q
    return a


def reduce_mod_phi(m, q, a):
    phi = cyclotomic(m)
    d = len(phi) - 1
    a = np.concatenate([np.array(a, dtype=object), np.zeros(max(0, d - len(a)), dtype=object)]) % q
    for k in range(len(a) - 1, d - 1, -1):
        c = a[k]
        if c:
            a[k - d:k + 1] = (a[k - d:k + 1] - c * phi) % q
    return a[:d] % q


@pytest.mark.parametrize("m,q", CASES)
def test_structured_equals_definitional(m, q):
    rng = np.random.default_rng(m * 1000 + q)
    n = O.totient(O.factor(m))
    x = rng.integers(0, q, size=n)
    assert (O.crt(m, q, x) == O.crt_def(m, q, x)).all()
    assert (O.crt_inv(m, q, O.crt(m, q, x)) == x).all()
    assert (O.crt(m, q, O.crt_inv(m, q, x)) == x).all()


@pytest.mark.parametrize("m,q", CASES)
def test_rows_are_evaluations_at_primitive_roots(m, q):
    n = O.totient(O.factor(m))
    w = O.root_of_unity(q, m)
    units, exps = O.crt_units(m), O.pow_exponents(m)
    assert sorted(units.tolist()) == [u for u in range(m) if np.gcd(u, m) == 1]
    assert len(set(exps.tolist())) == n
    rng = np.random.default_rng(7)
    x = rng.integers(0, q, size=n)
    y = O.crt(m, q, x)
    for i in range(n):
        acc = sum(int(x[j]) * pow(w, int(units[i]) * int(exps[j]), q) for j in range(n)) % q
        assert acc == y[i]


@pytest.mark.parametrize("m,q", CASES)
def test_crt_is_ring_homomorphism_vs_schoolbook(m, q):
    rng = np.random.default_rng(m + q)
    n = O.totient(O.factor(m))
    a, b = rng.integers(0, q, size=n), rng.integers(0, q, size=n)
    c = O.crt_inv(m, q, (O.crt(m, q, a).astype(object) * O.crt(m, q, b).astype(object) % q).astype(np.int64))
    pa, pb, pc = poly_of_pow(m, q, a), poly_of_pow(m, q, b), poly_of_pow(m, q, c)
    prod = np.zeros(2 * m, dtype=object)
    for i, ai in enumerate(pa):
        if ai:
            prod[i:i + m] = (prod[i:i + m] + ai * pb) % q
    assert (reduce_mod_phi(m, q, prod) == reduce_mod_phi(m, q, pc)).all()


@pytest.mark.parametrize("m,q", CASES)
def test_g_and_l(m, q):
    rng = np.random.default_rng(3 * m + q)
    n = O.totient(O.factor(m))
    x = rng.integers(0, q, size=n)
    # mulG in the CRT basis is pointwise multiplication by g's evaluations
    assert (O.crt(m, q, O.mul_g_pow(m, q, x)) == O.mul_g_crt(m, q, O.crt(m, q, x))).all()
    # g = prod_{odd p | m} (1 - zeta_p) as a polynomial
    one = np.zeros(n, dtype=np.int64)
    one[0] = 1
    g = np.array([1], dtype=object)
    for p, _ in O.factor(m):
        if p != 2:
            f = np.zeros(m // p + 1, dtype=object)
            f[0], f[m // p] = 1, -1
            g = np.convolve(g, f)
    assert (reduce_mod_phi(m, q, poly_of_pow(m, q, O.mul_g_pow(m, q, one))) == reduce_mod_phi(m, q, g)).all()
    for mul, div in ((O.mul_g_pow, O.div_g_pow), (O.mul_g_dec, O.div_g_dec), (O.mul_g_crt, O.div_g_crt)):
        assert (div(m, q, mul(m, q, x)) == x).all()
        assert (mul(m, q, div(m, q, x)) == x).all()
    assert (O.l_inv(m, q, O.l(m, q, x)) == x).all()
    # G in the decoding basis is L^-1 G_pow L
    assert (O.mul_g_dec(m, q, x) == O.l_inv(m, q, O.mul_g_pow(m, q, O.l(m, q, x)))).all()


EXT = [(4, 16, 97), (3, 9, 19), (3, 21, 43), (7, 21, 43), (5, 45, 181), (9, 45, 181), (15, 45, 181),
       (4, 36, 37), (12, 36, 37), (1, 15, 31), (8, 40, 41), (21, 105, 211), (35, 105, 211)]


@pytest.mark.parametrize("m,mp,q", EXT)
def test_embed_twace_commute_with_crt(m, mp, q):
    rng = np.random.default_rng(m * mp)
    n, npr = O.totient(O.factor(m)), O.totient(O.factor(mp))
    ext, base = O.ext_indices_pow(m, mp), O.base_indices_crt(m, mp)
    x = rng.integers(0, q, size=n)
    # embedPow: scatter; the embedded element is the same polynomial (zeta_m = zeta_mp^(mp/m))
    xe = np.zeros(npr, dtype=np.int64)
    xe[ext] = x
    lo, hi = poly_of_pow(m, q, x), poly_of_pow(mp, q, xe)
    lifted = np.zeros(mp, dtype=object)
    lifted[::mp // m] = lo
    assert (reduce_mod_phi(mp, q, lifted) == reduce_mod_phi(mp, q, hi)).all()
    # embedCRT = replicate through base indices
    assert (O.crt(mp, q, xe) == O.crt(m, q, x)[base]).all()
    # twace: Tw(y) = (mhat/mhat') Tr(y g'/g); in CRT: sum over the rows above, weighted by g'/g
    y = rng.integers(0, q, size=npr)
    mhat = lambda k: k // 2 if k % 2 == 0 else k
    scale = mhat(m) * pow(mhat(mp), q - 2, q) % q
    gy = O.crt(mp, q, y).astype(object) * O.g_crt(mp, q).astype(object) % q
    tw = np.zeros(n, dtype=object)
    np.add.at(tw, base, gy)
    ginv = np.array([pow(int(v), q - 2, q) for v in O.g_crt(m, q)], dtype=object)
    tw_crt = (tw * ginv * scale % q).astype(np.int64)
    # ... and it equals the index gather in the powerful (and decoding) basis
    assert (O.crt(m, q, y[ext]) == tw_crt).all()
    ydec = O.l_inv(mp, q, y)
    assert (O.l(m, q, ydec[ext]) == O.crt_inv(m, q, tw_crt)).all()
    # twace . embed = id
    assert (xe[ext] == x).all()


@pytest.mark.parametrize("m,mp,q", EXT)
def test_coeffs_factor_the_powerful_basis(m, mp, q):
    n, npr = O.totient(O.factor(m)), O.totient(O.factor(mp))
    tab = O.coeff_indices(m, mp)
    assert sorted(tab.reshape(-1).tolist()) == list(range(npr))
    assert (tab[0] == O.ext_indices_pow(m, mp)).all()
    # pow'(tab[k, j]) = pow'(tab[k, 0]) * embed(pow(j))
    ehi, elo = O.pow_exponents(mp), O.pow_exponents(m)
    for k in range(tab.shape[0]):
        for j in range(n):
            assert ehi[tab[k, j]] == (ehi[tab[k, 0]] + elo[j] * (mp // m)) % mp


def test_power_of_two_is_the_negacyclic_ntt():
    m, q = 64, 193
    n = 32
    psi = O.root_of_unity(q, m)
    rng = np.random.default_rng(5)
    x = rng.integers(0, q, size=n)
    y = O.crt(m, q, x)
    rev = lambda a: int(format(a, "05b")[::-1], 2)
    for i in range(n):
        assert y[i] == sum(int(x[a]) * pow(psi, rev(a) * (2 * i + 1), q) for a in range(n)) % q
    pl = O.Plan(m, q)
    assert (pl.crt_generic(x.copy()) == y).all()


def test_example_rings_have_roots_and_roundtrip():
    rings = {512: [268440577, 8392193], 11648: [537264001], 20475: [1543651201, 1567238401],
             54600: [720720001], 40320: [1073479681, 949616641], 65536: [1073479681, 1052835841, 1032192001, 949616641]}
    rng = np.random.default_rng(11)
    for m, qs in rings.items():
        n = O.totient(O.factor(m))
        for q in qs:
            x = rng.integers(0, q, size=(2, n))
            pl = O.Plan(m, q)
            y = pl.crt_many(x.copy(), threads=2)
            assert (pl.crt_many(y.copy(), inverse=True) == x).all()
            # linearity
            z = pl.crt_many(((x[0] + x[1]) % q)[None].copy())
            assert (z[0] == (y[0] + y[1]) % q).all()
