"""Tier-2: the SymmSHE-level oracle (oracle/she_oracle.py) against convention-independent facts (no GPU).

The reference's own criterion is semantic -- decrypt(eval ct) == eval pt on random inputs
(examples/Arithmetic.hs:75, examples/HomomRLWR.hs:71) -- so each ciphertext operation is checked through
the *phase* c(s) = c0 + c1 s (+ c2 s^2): multiplication is exact, key switching and tunnelling preserve the
phase (resp. map it through the E-linear function) up to a small, explicitly bounded noise term, and modulus
switching scales it by 1/q_drop up to rounding.
"""
from math import prod

import numpy as np
import pytest

from oracle import oracle as O
from oracle import she_oracle as S


def primes_1_mod(m, lo, count):
    out, q = [], lo - lo % m + 1
    while len(out) < count:
        if q > lo and all(q % d for d in range(2, int(q ** 0.5) + 1)):
            out.append(q)
        q += m
    return out


def crt_lift(res, moduli):
    """residues [L, ...] -> centred integers mod prod(moduli) (python ints)"""
    Q = prod(moduli)
    acc = np.zeros(res.shape[1:], dtype=object)
    for l, q in enumerate(moduli):
        Ql = Q // q
        acc = (acc + res[l].astype(object) * (Ql * pow(Ql, -1, q))) % Q
    return np.where(acc * 2 >= Q, acc - Q, acc), Q


def small(rng, shape, bound=1):
    return rng.integers(-bound, bound + 1, size=shape)


def to_rns(x, moduli):
    return np.stack([np.mod(x, q).astype(np.int64) for q in moduli], axis=-2)


def ring_mul(m, moduli, a, b):
    """product of powerful-basis elements a, b [..., L, n] -> powerful basis"""
    qv = np.array(moduli, dtype=np.int64)[:, None]
    return S.crt_inv(m, moduli, S.crt(m, moduli, a) * S.crt(m, moduli, b) % qv)


@pytest.mark.parametrize("base", [0, 2, 3, 16])
def test_decompose_against_gadget(base):
    moduli = [97, 193, 257]
    rng = np.random.default_rng(base)
    x = np.stack([rng.integers(0, q, size=(4, 8)) for q in moduli], axis=1)
    x[0, :, 0] = [q - 1 for q in moduli]
    x[0, :, 1] = [q // 2 for q in moduli]
    x[0, :, 2] = [q // 2 + 1 for q in moduli]
    d = S.decompose_signed(x, moduli, base)
    gad = S.gadget_vector(moduli, base)
    assert d.shape[-2] == len(gad) == S.gadget_len(moduli, base)
    for l, q in enumerate(moduli):
        acc = sum(d[:, k] * v for k, (gl, v) in enumerate(gad) if gl == l)
        assert (np.mod(acc, q) == x[:, l]).all()
    if base:
        for k, (gl, v) in enumerate(gad):
            last = k + 1 == len(gad) or gad[k + 1][0] != gl
            if not last:
                assert (d[:, k] >= -(base // 2) - (base % 2)).all() and (2 * d[:, k] < base).all()
            else:
                assert (np.abs(d[:, k]) <= base).all()
    red = S.decompose(x, moduli, moduli[:2], base)
    assert red.shape == (4, len(gad), 2, 8) and (red >= 0).all() and (red[:, :, 1] < moduli[1]).all()


def test_rescale_is_exact_division_by_the_dropped_modulus():
    moduli = [1073479681, 1052835841, 1032192001]
    rng = np.random.default_rng(2)
    x = np.stack([rng.integers(0, q, size=(3, 16)) for q in moduli], axis=1)
    X, Q = crt_lift(x.transpose(1, 0, 2), moduli)
    for drop in (2, 0, 1):
        kept = [q for i, q in enumerate(moduli) if i != drop]
        r = S.rescale_drop(x, moduli, drop)
        c = S.center(x[:, drop], moduli[drop]).astype(object)
        want = (X - c) // moduli[drop]
        assert ((X - c) % moduli[drop] == 0).all()
        for k, q in enumerate(kept):
            assert (r[:, k] == np.mod(want, q).astype(np.int64)).all()
        # |x/q_drop - rescale(x)| <= 1/2
        assert (abs(want * moduli[drop] - X) * 2 <= moduli[drop]).all()


def test_ct_mul_is_the_product_of_phases_times_g():
    m, moduli = 36, primes_1_mod(36, 1 << 20, 2)
    n = O.totient(O.factor(m))
    rng = np.random.default_rng(3)
    qv = np.array(moduli, dtype=np.int64)[None, :, None]
    a = np.stack([rng.integers(0, q, size=(2, 2, n)) for q in moduli], axis=2)
    b = np.stack([rng.integers(0, q, size=(2, 2, n)) for q in moduli], axis=2)
    s = S.crt(m, moduli, to_rns(small(rng, (n,)), moduli))[None]
    c = S.ct_mul(m, moduli, a, b)
    g = np.stack([O.g_crt(m, q) for q in moduli])[None]
    pa = (a[:, 0] + a[:, 1] * s % qv) % qv
    pb = (b[:, 0] + b[:, 1] * s % qv) % qv
    pc = (c[:, 0] + c[:, 1] * s % qv + c[:, 2] * (s * s % qv) % qv) % qv
    assert (pc == pa * pb % qv * g % qv).all()


@pytest.mark.parametrize("base", [0, 2, 8])
def test_key_switch_preserves_the_phase_up_to_bounded_noise(base):
    m = 32
    n = 16
    moduli = primes_1_mod(m, 1 << 26, 2)
    rng = np.random.default_rng(4 + base)
    qv = np.array(moduli, dtype=np.int64)[None, :, None]
    s_pow = to_rns(small(rng, (n,)), moduli)
    s = S.crt(m, moduli, s_pow)
    s2 = s * s % qv[0]
    gad = S.gadget_vector(moduli, base)
    D = len(gad)
    # hint_d = (b_d, a_d) with b_d + a_d s = gadget_d s^2 + e_d, e_d small
    a = np.stack([rng.integers(0, q, size=(D, n)) for q in moduli], axis=1)
    e = small(rng, (D, n), 2)
    e_crt = S.crt(m, moduli, to_rns(e, moduli))
    gs2 = np.zeros((D, len(moduli), n), dtype=np.int64)
    for d, (l, v) in enumerate(gad):
        gs2[d, l] = s2[l] * v % moduli[l]
    b = (gs2 + e_crt - a * s[None] % qv) % qv
    hint = np.stack([b, a], axis=1)
    c3 = np.stack([rng.integers(0, q, size=(3, 3, n)) for q in moduli], axis=2)
    out = S.key_switch_quad(m, moduli, c3, hint, base)
    ph_in = (c3[:, 0] + c3[:, 1] * s % qv + c3[:, 2] * s2[None] % qv) % qv
    ph_out = (out[:, 0] + out[:, 1] * s % qv) % qv
    diff = S.crt_inv(m, moduli, (ph_out - ph_in) % qv)
    err, _ = crt_lift(diff.transpose(1, 0, 2), moduli)
    digit_bound = max(moduli) // 2 + 1 if base == 0 else base
    assert (abs(err) <= D * n * digit_bound * 2).all()
    assert abs(err).max() > 0            # there is noise: the hint errors were used


def test_mod_switch_scales_the_phase():
    m, n = 32, 16
    moduli = primes_1_mod(m, 1 << 26, 3)
    rng = np.random.default_rng(6)
    s_int = small(rng, (n,))
    for frm, to in ((3, 2), (3, 1), (2, 1)):
        q = moduli[:frm]
        qv = np.array(q, dtype=np.int64)[None, :, None]
        s = S.crt(m, q, to_rns(s_int, q))
        c = np.stack([rng.integers(0, qq, size=(2, 2, n)) for qq in q], axis=2)
        out = S.mod_switch(m, moduli, c, to)
        assert out.shape == (2, 2, to, n)
        ph_in = S.crt_inv(m, q, (c[:, 0] + c[:, 1] * s % qv) % qv)
        qt = moduli[:to]
        qvt = qv[:, :to]
        ph_out = S.crt_inv(m, qt, (out[:, 0] + out[:, 1] * s[:to] % qvt) % qvt)
        X, Q = crt_lift(ph_in.transpose(1, 0, 2), q)
        Y, Qt = crt_lift(ph_out.transpose(1, 0, 2), qt)
        dropped = Q // Qt
        # Y = X / dropped + rounding, rounding <= (1 + |s|_1) / 2 per dropped limb (accumulated, with slack)
        d = (Y * dropped - X) % Q
        d = np.where(d * 2 >= Q, d - Q, d)
        assert (abs(d) <= dropped * (1 + np.abs(s_int).sum()) * (frm - to)).all()
    # switching up multiplies by the new moduli exactly and appends zero limbs
    c = np.stack([rng.integers(0, qq, size=(1, 2, n)) for qq in moduli[:1]], axis=2)
    up = S.mod_switch(m, moduli, c, 3)
    assert (up[:, :, 1:] == 0).all()
    assert (up[:, :, 0] == c[:, :, 0] * (moduli[1] * moduli[2] % moduli[0]) % moduli[0]).all()
    assert (S.mod_switch(m, moduli, c, 1) == c).all()


def test_mod_switch_rounds_each_polynomial_in_the_basis_its_mask_bit_names():
    """dec_mask of S.mod_switch (Lol rescaleLinearMSD: c0 through rescaleDec, c1 through rescalePow): in the chosen basis every
    coefficient of the result is the input coefficient divided by the dropped moduli, off by at most half a unit per dropped limb;
    on a ring with odd primes the two bases give different ciphertexts."""
    m, n = 45, 24
    moduli = primes_1_mod(m, 1 << 24, 3)
    rng = np.random.default_rng(9)
    c = np.stack([rng.integers(0, qq, size=(3, 2, n)) for qq in moduli], axis=2)
    outs = {}
    for dec_mask in (0, 1, 2, 3):
        for to in (2, 1):
            out = S.mod_switch(m, moduli, c, to, dec_mask)
            outs[(dec_mask, to)] = out
            qt, dropped = moduli[:to], prod(moduli[to:])
            x_in, x_out = S.crt_inv(m, moduli, c), S.crt_inv(m, qt, out)
            for i in range(2):
                a, b = x_in[:, i], x_out[:, i]
                if (dec_mask >> i) & 1:
                    a, b = S.per_limb(O.l_inv, m, moduli, a), S.per_limb(O.l_inv, m, qt, b)
                X, Q = crt_lift(a.transpose(1, 0, 2), moduli)
                Y, Qt = crt_lift(b.transpose(1, 0, 2), qt)
                d = (Y * dropped - X) % Q
                d = np.where(d * 2 >= Q, d - Q, d)
                assert (2 * abs(d) <= dropped * (3 - to)).all(), (dec_mask, to, i)
    assert (outs[(1, 1)][:, 1] == outs[(0, 1)][:, 1]).all() and (outs[(1, 1)][:, 0] != outs[(0, 1)][:, 0]).any()
    assert (outs[(2, 1)][:, 0] == outs[(0, 1)][:, 0]).all() and (outs[(3, 1)][:, 0] == outs[(1, 1)][:, 0]).all()


TUNNELS = [(12, 4, 20), (8, 4, 24), (9, 3, 21), (28, 4, 12), (45, 15, 15), (16, 16, 48)]


@pytest.mark.parametrize("r,e,s_idx", TUNNELS)
@pytest.mark.parametrize("base", [0, 4])
def test_tunnel_maps_the_phase_through_an_e_linear_function(r, e, s_idx, base):
    from math import lcm
    moduli = primes_1_mod(lcm(r, s_idx), 1 << 24, 2)
    L = len(moduli)
    nr, ne, ns = (O.totient(O.factor(x)) for x in (r, e, s_idx))
    K = nr // ne
    rng = np.random.default_rng(r * 100 + s_idx + base)
    qv = np.array(moduli, dtype=np.int64)[None, :, None]
    tab = O.coeff_indices(e, r)

    # an arbitrary E-linear f: O_r -> O_s is fixed by its values y_k on the relative decoding basis
    yk = S.crt(s_idx, moduli, np.stack([rng.integers(0, q, size=(K, ns)) for q in moduli], axis=1))

    def f_lin(x_pow):
        """x in the powerful basis of O_r [.., L, nr] -> f(x) in the CRT basis of O_s"""
        dec = S.per_limb(O.l_inv, r, moduli, x_pow)
        acc = 0
        for k in range(K):
            part = S.per_limb(O.l, e, moduli, dec[..., tab[k]])
            acc = (acc + S.crt(s_idx, moduli, S.embed_pow(e, s_idx, part)) * yk[k] % qv[0]) % qv[0]
        return acc

    s_in = to_rns(small(rng, (nr,)), moduli)
    s_out = S.crt(s_idx, moduli, to_rns(small(rng, (ns,)), moduli))
    gad = S.gadget_vector(moduli, base)
    D = len(gad)
    # hint_{k,d} = (b, a): b + a s_out = gadget_d f(relpow_k s_in) + small
    hint = np.zeros((K, D, 2, L, ns), dtype=np.int64)
    for k in range(K):
        relpow = np.zeros((L, nr), dtype=np.int64)
        relpow[:, tab[k][0]] = 1
        target = f_lin(ring_mul(r, moduli, relpow, s_in))
        for d, (l, v) in enumerate(gad):
            a = np.stack([rng.integers(0, q, size=ns) for q in moduli])
            err = S.crt(s_idx, moduli, to_rns(small(rng, (ns,), 2), moduli))
            gt = np.zeros((L, ns), dtype=np.int64)
            gt[l] = target[l] * v % moduli[l]
            hint[k, d, 0] = (gt + err - a * s_out % qv[0]) % qv[0]
            hint[k, d, 1] = a
    c_pow = np.stack([rng.integers(0, q, size=(2, 2, nr)) for q in moduli], axis=2)
    c = S.crt(r, moduli, c_pow)
    out = S.tunnel(r, e, s_idx, moduli, c, yk, hint, base)
    assert out.shape == (2, 2, L, ns)
    phase_in = (c_pow[:, 0] + ring_mul(r, moduli, c_pow[:, 1], s_in[None])) % qv
    want = f_lin(phase_in)
    got = (out[:, 0] + out[:, 1] * s_out[None] % qv) % qv
    diff = S.crt_inv(s_idx, moduli, (got - want) % qv)
    err, _ = crt_lift(diff.transpose(1, 0, 2), moduli)
    digit_bound = max(moduli) // 2 + 1 if base == 0 else base
    assert (abs(err) <= K * D * ns * 8 * digit_bound * 2).all()
    assert prod(moduli) > 1000 * K * D * ns * 8 * digit_bound * 2      # the bound is far below the modulus
