"""The SymmSHE boundary oracle (oracle/symm_oracle.py) against the reference's own criterion (no GPU):
decrypt (eval ct) == eval pt on random inputs -- examples/Arithmetic.hs:75, examples/HomomRLWR.hs:71 -- for
encrypt/decrypt, (+), (*), addPublic, mulPublic, modSwitchPT and the PT2CT `mul_` pipeline
modSwitch . keySwitchQuad . modSwitch . (*) (Crypto/Alchemy/Interpreter/PT2CT.hs:172-177)."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import symm_oracle as Y
from oracle import she_oracle as S

ARITH_Q = [268440577, 8392193, 1073750017]


def primes_1_mod(m, lo, count):
    out, q = [], lo - lo % m + 1
    while len(out) < count:
        if q > lo and all(q % d for d in range(2, int(q ** 0.5) + 1)):
            out.append(q)
        q += m
    return out


def plain_mul(m, p, a, b):
    """product in Z_p[X]/Phi_m, powerful basis: through the CRT over one big prime, exact over the integers"""
    q0 = primes_1_mod(m, 1 << 29, 1)[0]          # products of residues stay below 2^62
    lift = lambda x: np.mod(np.where(2 * np.mod(x, p) >= p, np.mod(x, p) - p, np.mod(x, p)), q0)
    c = O.crt_inv(m, q0, O.crt(m, q0, lift(a)) * O.crt(m, q0, lift(b)) % q0)
    return np.mod(np.where(2 * c > q0, c - q0, c), p)


def setup(m, mp, moduli, p, B, seed):
    rng = np.random.default_rng(seed)
    sch = Y.SymmSHE(m, mp, moduli, p)
    sk = sch.gen_sk(rng.integers(-1, 2, size=sch.np_))
    draw = lambda L: dict(mu=rng.integers(0, p, size=(B, sch.n)), noise=rng.integers(-2, 3, size=(B, sch.np_)),
                          c1=np.stack([rng.integers(0, q, size=(B, sch.np_)) for q in moduli[:L]], axis=1))
    return rng, sch, sk, draw


RINGS = [(4, 16, primes_1_mod(16, 1 << 20, 3), 7), (3, 45, primes_1_mod(45, 1 << 22, 3), 4), (9, 63, primes_1_mod(63, 1 << 22, 2), 8),
         (4, 512, ARITH_Q, 7)]


@pytest.mark.parametrize("m,mp,moduli,p", RINGS)
def test_decrypt_inverts_encrypt_and_respects_ring_operations(m, mp, moduli, p):
    rng, sch, sk, draw = setup(m, mp, moduli, p, 3, m + mp)
    L = len(moduli)
    x, y = draw(L), draw(L)
    cx, cy = sch.encrypt(sk, x["mu"], x["noise"], x["c1"]), sch.encrypt(sk, y["mu"], y["noise"], y["c1"])
    assert (sch.decrypt(sk, cx) == x["mu"]).all()
    assert (sch.decrypt(sk, sch.to_msd(cx)) == x["mu"]).all()
    assert (sch.decrypt(sk, sch.add(cx, cy)) == (x["mu"] + y["mu"]) % p).all()
    prod_ct = sch.mul(cx, cy)
    want = np.stack([plain_mul(m, p, a[None], b[None])[0] for a, b in zip(x["mu"], y["mu"])])
    assert prod_ct.k == 1 and (sch.decrypt(sk, prod_ct) == want).all()
    # (x * y) + x: aligns k (mulG) on the fresh ciphertext
    assert (sch.decrypt(sk, sch.add(prod_ct, cx)) == (want + x["mu"]) % p).all()
    pub = rng.integers(0, p, size=sch.n)
    assert (sch.decrypt(sk, sch.add_public(prod_ct, pub)) == (want + pub) % p).all()
    want_pub = np.stack([plain_mul(m, p, a[None], pub[None])[0] for a in x["mu"]])
    assert (sch.decrypt(sk, sch.mul_public(cx, pub)) == want_pub).all()
    assert (sch.error_rate(sk, cx) < 1e-3).all()


@pytest.mark.parametrize("m,mp,moduli,p,lin,lh,lout,base", [(4, 16, primes_1_mod(16, 1 << 22, 3), 7, 2, 3, 1, 0),
                                                            (3, 45, primes_1_mod(45, 1 << 24, 3), 4, 2, 2, 1, 2),
                                                            (4, 512, ARITH_Q, 7, 2, 2, 1, 0)])      # examples/Arithmetic.hs, SURVEY B.3
def test_add_mul_circuit_of_arithmetic_hs(m, mp, moduli, p, lin, lh, lout, base):
    """addMul = \\x y -> (x + y) * y, compiled as PT2CT does: (*) then modSwitch . keySwitchQuad . modSwitch"""
    rng, sch, sk, draw = setup(m, mp, moduli, p, 2, 7 * m + mp)
    x, y = draw(lin), draw(lin)
    cx, cy = sch.encrypt(sk, x["mu"], x["noise"], x["c1"]), sch.encrypt(sk, y["mu"], y["noise"], y["c1"])
    D = S.gadget_len(moduli[:lh], base)
    mask = np.stack([rng.integers(0, q, size=(D, sch.np_)) for q in moduli[:lh]], axis=1)
    hint = sch.ks_hint(sk, mask, rng.integers(-1, 2, size=(D, sch.np_)), lh, base)
    ct = sch.mul(sch.add(cx, cy), cy)
    ct = sch.mod_switch(sch.key_switch_quad(sch.mod_switch(ct, lh), hint, base), lout)
    assert ct.c.shape[1:3] == (2, lout)
    want = np.stack([plain_mul(m, p, ((a + b) % p)[None], b[None])[0] for a, b in zip(x["mu"], y["mu"])])
    assert (sch.decrypt(sk, ct) == want).all()
    assert (sch.error_rate(sk, ct) < 0.25).all()


def test_mod_switch_pt_divides_the_plaintext():
    """div2_ (Eval.hs:95): a plaintext that is a multiple of p / p' is divided by it and lands in Z_p'"""
    m, mp, p, p2 = 3, 45, 8, 4
    moduli = primes_1_mod(45, 1 << 24, 2)
    rng, sch, sk, draw = setup(m, mp, moduli, p, 2, 5)
    x = draw(2)
    x["mu"] = 2 * rng.integers(0, p2, size=x["mu"].shape)
    ct = sch.mod_switch_pt(sch.encrypt(sk, x["mu"], x["noise"], x["c1"]), p2)
    got = Y.SymmSHE(m, mp, moduli, p2).decrypt(sk, ct)
    assert (got == x["mu"] // 2).all()


def test_tunnel_with_generated_hint_maps_the_phase():
    """tunnel_hint + S.tunnel: the phase of the output is f(phase of the input) up to bounded noise (no GPU)"""
    from math import lcm, prod
    r, e, s_idx, base = 12, 4, 20, 4
    moduli = primes_1_mod(lcm(r, s_idx), 1 << 24, 2)
    L = len(moduli)
    nr, ne, ns = (O.totient(O.factor(x)) for x in (r, e, s_idx))
    K, D = nr // ne, S.gadget_len(moduli, base)
    rng = np.random.default_rng(77)
    qv = np.array(moduli, dtype=np.int64)[:, None]
    rnd = lambda lead, n: np.stack([rng.integers(0, q, size=tuple(lead) + (n,)) for q in moduli], axis=-2)
    flin = rnd((K,), ns)
    s_in = S.crt(r, moduli, Y.lift_plain(rng.integers(-1, 2, size=nr), moduli, 1 << 40))
    s_out = S.crt(s_idx, moduli, Y.lift_plain(rng.integers(-1, 2, size=ns), moduli, 1 << 40))
    hint = Y.tunnel_hint(r, e, s_idx, moduli, s_in, s_out, flin, rnd((K, D), ns), rng.integers(-2, 3, size=(K, D, ns)), base)
    c = rnd((2, 2), nr)
    out = S.tunnel(r, e, s_idx, moduli, c, flin, hint, base)
    phase_in = S.crt_inv(r, moduli, (c[:, 0] + c[:, 1] * s_in % qv) % qv)
    want = Y.eval_lin(r, e, s_idx, moduli, phase_in, flin)
    got = (out[:, 0] + out[:, 1] * s_out % qv) % qv
    diff = S.crt_inv(s_idx, moduli, (got - want) % qv)
    Q = prod(moduli)
    acc = np.zeros((2, ns), dtype=object)
    for l, q in enumerate(moduli):
        Ql = Q // q
        acc = (acc + diff[:, l].astype(object) * (Ql * pow(Ql, -1, q))) % Q
    err = np.where(acc * 2 > Q, acc - Q, acc)
    bound = K * D * ns * 8 * base * 2 * ns
    assert (abs(err) <= bound).all() and Q > 1000 * bound
