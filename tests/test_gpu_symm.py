"""GPU parity for the SymmSHE boundary (alchemy_b200/symm.py over the C ABI) against oracle/symm_oracle.py, stage by
stage and bit for bit, on the circuit of examples/Arithmetic.hs -- addMul = \\x y -> (x + y) * y compiled as
PT2CT.hs:172-177 prescribes -- plus addPublic / mulPublic / modSwitchPT / errorRate, ending in the reference's own
criterion: decrypt (eval ct) == eval pt (examples/Arithmetic.hs:75)."""
import numpy as np
import pytest

from oracle import she_oracle as S
from oracle import symm_oracle as Y
from test_symm_oracle import ARITH_Q, plain_mul, primes_1_mod

pytestmark = pytest.mark.gpu

RLWR_Q = [1543651201, 689270401, 718099201, 720720001, 1556755201, 1567238401]


def dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.int64)).cuda()


def same(got, want):
    return (got.cpu().numpy() == want).all()


CASES = [  # m, m', moduli, p, L_in, L_hint, L_out, gadget base
    (4, 512, ARITH_Q, 7, 2, 2, 1, 0),                       # examples/Arithmetic.hs with its limb schedule
    (4, 16, primes_1_mod(16, 1 << 22, 3), 7, 2, 3, 1, 0),
    (3, 45, primes_1_mod(45, 1 << 24, 3), 4, 2, 2, 1, 2),   # odd primes in m': g != 1, mulG / divG on the decrypt side
    (9, 63, primes_1_mod(63, 1 << 24, 3), 8, 2, 3, 1, 4),
    (4095, 20475, RLWR_Q, 2, 2, 3, 1, 0),                   # HomomRLWR's last ring pair H5 -> H5' (Common.hs:52-54), Z_2 plaintexts
]


@pytest.mark.parametrize("m,mp,moduli,p,lin,lh,lout,base", CASES)
def test_arithmetic_circuit_stage_by_stage(m, mp, moduli, p, lin, lh, lout, base):
    from alchemy_b200 import symm
    rng = np.random.default_rng(m * 131 + mp)
    ora, gpu = Y.SymmSHE(m, mp, moduli, p), symm.SymmSHE(m, mp, moduli, p)
    B, n, np_ = 2, ora.n, ora.np_
    s_dec = rng.integers(-1, 2, size=np_)
    sk_o, sk_g = ora.gen_sk(s_dec), gpu.genSK(dev(s_dec))
    assert same(sk_g, sk_o)

    def enc(L):
        mu, noise = rng.integers(0, p, size=(B, n)), rng.integers(-2, 3, size=(B, np_))
        c1 = np.stack([rng.integers(0, q, size=(B, np_)) for q in moduli[:L]], axis=1)
        co, cg = ora.encrypt(sk_o, mu, noise, c1), gpu.encrypt(sk_g, dev(mu), dev(noise), dev(c1))
        assert same(cg.c, co.c)
        return mu, co, cg

    x, xo, xg = enc(lin)
    y, yo, yg = enc(lin)
    assert same(gpu.decrypt(sk_g, xg), x) and (ora.decrypt(sk_o, xo) == x).all()
    D = S.gadget_len(moduli[:lh], base)
    mask = np.stack([rng.integers(0, q, size=(D, np_)) for q in moduli[:lh]], axis=1)
    hnoise = rng.integers(-1, 2, size=(D, np_))
    hint_o, hint_g = ora.ks_hint(sk_o, mask, hnoise, lh, base), gpu.ksQuadCircHint(sk_g, dev(mask), dev(hnoise), lh, base)
    assert same(hint_g, hint_o)

    so, sg = ora.add(xo, yo), gpu.add(xg, yg)
    assert same(sg.c, so.c)
    po, pg = ora.mul(so, yo), gpu.mul(sg, yg)
    assert same(pg.c, po.c) and (pg.k, pg.l, pg.enc) == (po.k, po.l, po.enc)
    want = np.stack([plain_mul(m, p, ((a + b) % p)[None], b[None])[0] for a, b in zip(x, y)])
    assert same(gpu.decrypt(sk_g, pg), want)                 # degree-2 ciphertext, k = 1: divG on the way out
    for step_o, step_g in ((lambda c: ora.mod_switch(c, lh), lambda c: gpu.modSwitch(c, lh)),
                           (lambda c: ora.key_switch_quad(c, hint_o, base), lambda c: gpu.keySwitchQuad(c, hint_g, base)),
                           (lambda c: ora.mod_switch(c, lout), lambda c: gpu.modSwitch(c, lout))):
        po, pg = step_o(po), step_g(pg)
        assert same(pg.c, po.c) and (pg.k, pg.l, pg.enc) == (po.k, po.l, po.enc)
    assert tuple(pg.c.shape[1:3]) == (2, lout)
    # the one-call pipeline (ab_mul_key_switch_dev with the toMSD scalars folded into the product) is Lol's sequence, bit for bit
    fg = gpu.mulKeySwitch(sg, yg, hint_g, lh, lout, base)
    assert same(fg.c, po.c) and (fg.k, fg.l, fg.enc) == (po.k, po.l, po.enc)
    got = gpu.decrypt(sk_g, pg)
    assert same(got, ora.decrypt(sk_o, po)) and same(got, want)        # the reference's PASS criterion
    rate_g, rate_o = gpu.errorRate(sk_g, pg).cpu().numpy(), ora.error_rate(sk_o, po)
    assert np.allclose(rate_g, rate_o, rtol=1e-9, atol=0) and (rate_g < 0.25).all()

    # (x + y) * y + x: aligning a fresh ciphertext (k = 0, LSD) with the evaluated one; addPublic / mulPublic
    xo1, xg1 = ora.mod_switch(xo, lout), gpu.modSwitch(xg, lout)
    ao, ag = ora.add(po, xo1), gpu.add(pg, xg1)
    assert same(ag.c, ao.c) and same(gpu.decrypt(sk_g, ag), (want + x) % p)
    pub = rng.integers(0, p, size=n)
    bo, bg = ora.add_public(po, pub), gpu.addPublic(pg, dev(pub))
    assert same(bg.c, bo.c) and same(gpu.decrypt(sk_g, bg), (want + pub) % p)
    mo, mg = ora.mul_public(xo, pub), gpu.mulPublic(xg, dev(pub))
    want_pub = np.stack([plain_mul(m, p, a[None], pub[None])[0] for a in x])
    assert same(mg.c, mo.c) and same(gpu.decrypt(sk_g, mg), want_pub)


def test_mod_switch_pt_and_rns_lift_edges():
    """div2_ (Eval.hs:95) through the device path; centred lift at the extremes of the RNS range"""
    import ctypes as C

    import torch

    import alchemy_b200 as ab
    from alchemy_b200 import symm
    m, mp, p, p2 = 3, 45, 8, 4
    moduli = primes_1_mod(45, 1 << 24, 3)
    rng = np.random.default_rng(11)
    ora, gpu = Y.SymmSHE(m, mp, moduli, p), symm.SymmSHE(m, mp, moduli, p)
    s_dec = rng.integers(-1, 2, size=ora.np_)
    sk_o, sk_g = ora.gen_sk(s_dec), gpu.genSK(dev(s_dec))
    mu = 2 * rng.integers(0, p2, size=(2, ora.n))
    noise = rng.integers(-2, 3, size=(2, ora.np_))
    c1 = np.stack([rng.integers(0, q, size=(2, ora.np_)) for q in moduli[:2]], axis=1)
    cg = gpu.modSwitchPT(gpu.encrypt(sk_g, dev(mu), dev(noise), dev(c1)), p2)
    co = ora.mod_switch_pt(ora.encrypt(sk_o, mu, noise, c1), p2)
    assert same(cg.c, co.c) and cg.l == co.l
    assert same(symm.SymmSHE(m, mp, moduli, p2).decrypt(sk_g, cg), mu // 2)
    # centred lift: 0, +-1, +-(Q-1)/2 and random residues, for 1..3 limbs, against python integers
    plan = ab.Plan(mp, moduli)
    for L in (1, 2, 3):
        q = moduli[:L]
        Q = int(np.prod([int(v) for v in q], dtype=object))
        vals = [0, 1, -1, (Q - 1) // 2, -((Q - 1) // 2), (Q - 1) // 2 - 1, 12345, -98765] + [int(rng.integers(0, 1 << 62)) % Q - Q // 2 for _ in range(plan.n - 8)]
        x = np.array([[v % ql for v in vals] for ql in q], dtype=np.int64)[None]
        for pmod in (2, 7, 32, (1 << 31) - 1):
            out = torch.empty((1, plan.n), dtype=torch.int64, device="cuda")
            rate = torch.zeros(1, dtype=torch.float64, device="cuda")
            ab.check(ab.lib().ab_rns_lift_reduce_dev(plan.h, C.c_void_p(out.data_ptr()), C.c_void_p(rate.data_ptr()), C.c_void_p(dev(x).data_ptr()), 1, L, pmod, None))
            assert out.cpu().numpy()[0].tolist() == [v % pmod for v in vals]
            assert abs(rate.item() - max(abs(v) for v in vals) / Q) < 1e-12


def test_error_rate_of_small_negative_terms_at_six_limbs_and_large_batches():
    """errorRate_ (Eval.hs:155-160) at ALCHEMY's real moduli: Q = 2^180, error terms of a few units, so rates near 1e-54.  A negative
    term must report |x| / Q to double precision (not 1 - x/Q, which bottoms out at 1e-16); more than 65535 ring elements in one call."""
    import ctypes as C

    import torch

    import alchemy_b200 as ab
    plan = ab.Plan(20475, RLWR_Q[:6])
    n, rng = plan.n, np.random.default_rng(5)
    for L in (4, 6):
        q = RLWR_Q[:L]
        Q = int(np.prod([int(v) for v in q], dtype=object))
        rows = [[-3] + [0] * (n - 1), [2, -7] + [1] * (n - 2), [-(1 << 70)] + [5] * (n - 1), [int(v) for v in rng.integers(-50, 51, size=n)]]
        x = np.array([[[v % ql for v in row] for ql in q] for row in rows], dtype=np.int64)
        out = torch.empty((len(rows), n), dtype=torch.int64, device="cuda")
        rate = torch.zeros(len(rows), dtype=torch.float64, device="cuda")
        ab.check(ab.lib().ab_rns_lift_reduce_dev(plan.h, C.c_void_p(out.data_ptr()), C.c_void_p(rate.data_ptr()), C.c_void_p(dev(x).data_ptr()), len(rows), L, 32, None))
        assert out.cpu().numpy().tolist() == [[v % 32 for v in row] for row in rows]
        want = np.array([max(abs(v) for v in row) / Q for row in rows])
        assert np.allclose(rate.cpu().numpy(), want, rtol=1e-12, atol=0), (rate.cpu().numpy(), want)
    small = ab.Plan(16, primes_1_mod(16, 1 << 22, 2))
    B, q = 70000, primes_1_mod(16, 1 << 22, 2)
    vals = rng.integers(-1000, 1001, size=(B, small.n))
    x = np.stack([np.mod(vals, ql) for ql in q], axis=1)
    out = torch.empty((B, small.n), dtype=torch.int64, device="cuda")
    rate = torch.zeros(B, dtype=torch.float64, device="cuda")
    ab.check(ab.lib().ab_rns_lift_reduce_dev(small.h, C.c_void_p(out.data_ptr()), C.c_void_p(rate.data_ptr()), C.c_void_p(dev(x).data_ptr()), B, 2, 7, None))
    assert (out.cpu().numpy() == np.mod(vals, 7)).all()
    assert np.allclose(rate.cpu().numpy(), np.abs(vals).max(axis=1) / float(q[0] * q[1]), rtol=1e-12, atol=0)


def test_device_randomness_matches_the_cpu_restatement_and_drives_the_circuit():
    """keys, hint and encryptions drawn on the device (Philox4x32-10), bit-identical to oracle/rng_oracle.py, and the
    addMul circuit of examples/Arithmetic.hs still decrypts to the plaintext result"""
    import alchemy_b200 as ab
    from alchemy_b200 import symm
    from oracle import rng_oracle as R
    m, mp, moduli, p, lin, lh, lout = 4, 512, ARITH_Q, 7, 2, 2, 1
    gpu, ora = symm.SymmSHE(m, mp, moduli, p), Y.SymmSHE(m, mp, moduli, p)
    cdt = R.gaussian_cdt(3.0 ** 0.5)                     # evalKeysHints (3.0 :: Double), examples/Arithmetic.hs:50
    seed = 0x5eed5eed12345678
    for stream_id, batch in ((2, 3), (4, 2)):
        assert same(ab.GT.uniform(gpu.hi, batch, lin, seed, stream_id), R.uniform(moduli[:lin], batch, ora.np_, seed, stream_id))
    assert same(ab.GT.uniform(gpu.hi, 1, 3, 7, 9), R.uniform(moduli, 1, ora.np_, 7, 9))
    for stream_id, batch in ((0, 1), (1, 5), (3, 2)):
        assert same(ab.GT.gaussianDec(gpu.hi, batch, cdt, seed, stream_id), R.gaussian(cdt, batch, ora.np_, seed, stream_id))
    sk = gpu.sampleSK(cdt, seed)
    assert same(sk, ora.gen_sk(R.gaussian(cdt, 1, ora.np_, seed, 0)[0]))
    hint = gpu.sampleKSQuadCircHint(sk, lh, cdt, seed)
    D = S.gadget_len(moduli[:lh], 0)
    assert same(hint, ora.ks_hint(sk.cpu().numpy(), R.uniform(moduli[:lh], D, ora.np_, seed, 4), R.gaussian(cdt, D, ora.np_, seed, 3), lh, 0))
    rng = np.random.default_rng(3)
    x, y = rng.integers(0, p, size=(4, ora.n)), rng.integers(0, p, size=(4, ora.n))
    cx, cy = gpu.encryptRandom(sk, dev(x), lin, cdt, seed), gpu.encryptRandom(sk, dev(y), lin, cdt, seed + 1)
    ct = gpu.mul(gpu.add(cx, cy), cy)
    ct = gpu.modSwitch(gpu.keySwitchQuad(gpu.modSwitch(ct, lh), hint), lout)
    want = np.stack([plain_mul(m, p, ((a + b) % p)[None], b[None])[0] for a, b in zip(x, y)])
    assert same(gpu.decrypt(sk, ct), want)
    assert (gpu.errorRate(sk, ct).cpu().numpy() < 0.25).all()


@pytest.mark.parametrize("r,e,s_idx,moduli,L,base", [(12, 4, 20, primes_1_mod(60, 1 << 24, 2), 2, 4), (9, 3, 21, primes_1_mod(63, 1 << 24, 2), 2, 0),
                                                     (11648, 5824, 29120, [537264001, 539884801], 1, 2)])      # Tunnel.hs hop 1, BaseBGad 2
def test_tunnel_hint_generated_on_the_device(r, e, s_idx, moduli, L, base):
    """Lol tunnelHint (KeysHints.hs:113) from device kernels: bit-identical to the CPU restatement, and tunnelling with it
    maps the phase c0 + c1 s_in through the E-linear function up to bounded noise"""
    from math import prod

    import alchemy_b200 as ab
    from alchemy_b200 import she, symm
    from oracle import oracle as O
    q = moduli[:L]
    rng = np.random.default_rng(r + s_idx + base)
    nr, ne, ns = (O.totient(O.factor(x)) for x in (r, e, s_idx))
    K, D = nr // ne, S.gadget_len(q, base)
    qv = np.array(q, dtype=np.int64)[:, None]
    rnd = lambda lead, n: np.stack([rng.integers(0, ql, size=tuple(lead) + (n,)) for ql in q], axis=-2)
    flin, mask, noise = rnd((K,), ns), rnd((K, D), ns), rng.integers(-2, 3, size=(K, D, ns))
    s_in = S.crt(r, q, Y.lift_plain(rng.integers(-1, 2, size=nr), q, 1 << 40))
    s_out = S.crt(s_idx, q, Y.lift_plain(rng.integers(-1, 2, size=ns), q, 1 << 40))
    want_hint = Y.tunnel_hint(r, e, s_idx, q, s_in, s_out, flin, mask, noise, base)
    pr, ps, pe = ab.Plan(r, q), ab.Plan(s_idx, q), ab.Plan(e, q, crt=False)
    tn = she.Tunnel(pr, pe, ps)
    hint = symm.tunnelHint(tn, ab.Ext(pe, pr), q, dev(s_in), dev(s_out), dev(flin), dev(mask), dev(noise), base)
    assert same(hint, want_hint)
    c = rnd((2, 2), nr)
    out = tn(dev(c), dev(flin), hint, base).cpu().numpy()
    assert (out == S.tunnel(r, e, s_idx, q, c, flin, want_hint, base)).all()
    phase_in = S.crt_inv(r, q, (c[:, 0] + c[:, 1] * s_in % qv) % qv)
    want = Y.eval_lin(r, e, s_idx, q, phase_in, flin)
    got = (out[:, 0] + out[:, 1] * s_out % qv) % qv
    diff = S.crt_inv(s_idx, q, (got - want) % qv)
    Q = prod(q)
    acc = np.zeros((2, ns), dtype=object)
    for l, ql in enumerate(q):
        Ql = Q // ql
        acc = (acc + diff[:, l].astype(object) * (Ql * pow(Ql, -1, ql))) % Q
    err = np.where(acc * 2 > Q, acc - Q, acc)
    digit = max(q) // 2 + 1 if base == 0 else base
    bound = K * D * ns * 8 * digit * 2 * 13
    assert (abs(err) <= bound).all()
