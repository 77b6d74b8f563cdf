"""Tier-1: every Tensor method of the CUDA backend, through the C ABI, bit-exact against the CPU
oracle on the same seeded inputs (run under gpurun: `pytest -m gpu`).

Rings: BASELINE configs 1-5 plus small odd cases that exercise every stage kind
of the mixed-radix program (p^e with e > 1 for p = 2, 3, 5; primes up to 13; m = 1, 2).
"""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu

CFG2_Q = [1073479681, 1052835841, 1032192001, 949616641]
ARITH_Q = [268440577, 8392193]
RLWR_Q = [1543651201, 689270401, 718099201, 720720001, 1556755201, 1567238401]
TUNNEL_Q = [537264001, 539884801, 555609601, 560851201, 566092801]

# (m, moduli)
POW2 = [(64, [193, 257]), (128, ARITH_Q[:1]), (256, ARITH_Q), (512, ARITH_Q), (1024, ARITH_Q[:1]),
        (2048, CFG2_Q), (4096, CFG2_Q[:2]), (8192, CFG2_Q[:3]), (16384, CFG2_Q), (32768, CFG2_Q[:2]), (65536, CFG2_Q)]
MIXED = [(9, [19, 37]), (27, [109]), (25, [101]), (45, [181, 541]), (36, [37, 73]), (12, [13]), (105, [211, 421]),
         (63, [127]), (40, [41]), (7, [29]), (13, [53]), (72, [73]), (200, [401]), (1, [17]), (2, [17]), (4, [17]),
         (40320, CFG2_Q[:2]), (11648, TUNNEL_Q[:2]), (29120, TUNNEL_Q[:1]), (43680, RLWR_Q[:2]), (54600, RLWR_Q[3:5]),
         (27300, RLWR_Q[:1]), (20475, RLWR_Q[:3]),
         # the 10-point pair step (p = 11), alone, beside 13 (stride 12: two columns per item) and over a modulus above 2^30 (signed sums)
         (11, [23, 67]), (55, [331]), (143, [859]), (1001, [1500018521, 2003])]
# 2^30 < q < 2^31: the exact (non-lazy) butterfly path; 2147483647 = 2^31 - 1
BIGQ = [(512, 2147483137), (2048, 2013265921), (63, 2147483647), (20475, 2147418001),
        (32768, 2013265921), (65536, 2013265921)]       # the exact butterflies in the large kernels (n = 16384; n = 32768: cluster forward, split inverse)


def torch_mod():
    import torch
    return torch


def backend():
    import alchemy_b200 as ab
    return ab


def rand_batch(rng, batch, moduli, n):
    return np.stack([rng.integers(0, q, size=(batch, n), dtype=np.int64) for q in moduli], axis=1)


def oracle_limbs(fn, m, moduli, x):
    """apply a one-limb oracle function to x[batch, L, n]"""
    return np.stack([fn(m, q, x[:, l]) for l, q in enumerate(moduli)], axis=1)


def dev(x):
    return torch_mod().from_numpy(np.ascontiguousarray(x)).cuda()


@pytest.mark.parametrize("m,moduli", POW2 + MIXED)
def test_crt_and_inverse_match_oracle(m, moduli):
    ab = backend()
    rng = np.random.default_rng(m)
    plan = ab.Plan(m, moduli)
    n = plan.n
    assert n == O.totient(O.factor(m))
    for l, q in enumerate(moduli):
        assert plan.root(l) == O.root_of_unity(q, m)
    batch = 3 if n > 4096 else 5
    x = rand_batch(rng, batch, moduli, n)
    want = oracle_limbs(O.crt, m, moduli, x)
    y = dev(x)
    ab.GT.crt(plan, y)
    assert (y.cpu().numpy() == want).all()
    # crtInv of the oracle's CRT returns the input; crtInv of fresh data matches the oracle
    ab.GT.crtInv(plan, y)
    assert (y.cpu().numpy() == x).all()
    z = dev(x)
    ab.GT.crtInv(plan, z)
    assert (z.cpu().numpy() == oracle_limbs(O.crt_inv, m, moduli, x)).all()
    # interleaved tuples (Lol's Storable layout) give the same numbers
    if len(moduli) > 1:
        yi = dev(x.transpose(0, 2, 1))
        ab.GT.crt(plan, yi, layout=ab.INTERLEAVED)
        assert (yi.cpu().numpy().transpose(0, 2, 1) == want).all()
        ab.GT.crtInv(plan, yi, layout=ab.INTERLEAVED)
        assert (yi.cpu().numpy().transpose(0, 2, 1) == x).all()


def test_crt_large_moduli_take_the_exact_butterflies():
    ab = backend()
    for m, q in BIGQ:
        plan = ab.Plan(m, [q])
        rng = np.random.default_rng(q % 1000)
        x = rand_batch(rng, 4, [q], plan.n)
        x[0, 0, :] = q - 1          # extreme residues
        y = dev(x)
        ab.GT.crt(plan, y)
        assert (y.cpu().numpy() == oracle_limbs(O.crt, m, [q], x)).all()
        ab.GT.crtInv(plan, y)
        assert (y.cpu().numpy() == x).all()


@pytest.mark.parametrize("m,moduli", [(20475, RLWR_Q[:1] + RLWR_Q[4:6]), (43680, RLWR_Q[:2]), (40320, CFG2_Q[:1]), (1001, [1500018521, 2003])])
def test_mixed_kernels_at_the_extremes_of_the_residue_range(m, moduli):
    """worst cases of the 64-bit sums of the dense and pair steps: all residues q - 1 (narrow: largest unsigned products), all at
    +-(q-1)/2 and alternating around the centring threshold (wide moduli: largest signed products of either sign, sign handling of the
    reduced values), against the oracle, forward and inverse"""
    ab = backend()
    plan = ab.Plan(m, moduli)
    n = plan.n
    rows = []
    for q in moduli:
        hq = (q - 1) // 2
        alt = np.where(np.arange(n) % 2 == 0, hq, hq + 1)
        rows.append(np.stack([np.full(n, q - 1), np.full(n, hq), np.full(n, hq + 1), alt, (alt[::-1]).copy(), np.where(np.arange(n) % 3 == 0, q - 1, 1)]))
    x = np.stack(rows, axis=1).astype(np.int64)          # [6, L, n]
    y = dev(x)
    ab.GT.crt(plan, y)
    assert (y.cpu().numpy() == oracle_limbs(O.crt, m, moduli, x)).all()
    z = dev(x)
    ab.GT.crtInv(plan, z)
    assert (z.cpu().numpy() == oracle_limbs(O.crt_inv, m, moduli, x)).all()


def test_limb_subrange_and_empty_batch():
    ab = backend()
    torch = torch_mod()
    m, moduli = 2048, CFG2_Q
    plan = ab.Plan(m, moduli)
    rng = np.random.default_rng(1)
    x = rand_batch(rng, 2, moduli[1:3], plan.n)
    y = dev(x)
    ab.GT.crt(plan, y, limb0=1, nlimbs=2)
    assert (y.cpu().numpy() == oracle_limbs(O.crt, m, moduli[1:3], x)).all()
    empty = torch.empty((0, 4, plan.n), dtype=torch.int64, device="cuda")
    ab.GT.crt(plan, empty)
    ab.GT.mul(plan, empty, empty)


def test_errors_are_status_codes_not_aborts():
    ab = backend()
    with pytest.raises(ab.AlchemyError) as e:
        ab.Plan(64, [101])          # 64 does not divide 100
    assert e.value.code == 2
    with pytest.raises(ab.AlchemyError) as e:
        ab.Plan(64, [15])                    # not prime
    assert e.value.code == 1
    plan = ab.Plan(64, [193], crt=False)
    y = dev(np.zeros((1, 1, 32), dtype=np.int64))
    with pytest.raises(ab.AlchemyError) as e:
        ab.GT.crt(plan, y)
    assert e.value.code == 2
    ab.GT.l(plan, y)                         # non-CRT methods still work without a root


LG = [(9, [19, 37]), (45, [181]), (105, [211, 421]), (64, [193]), (20475, RLWR_Q[:2]), (43680, RLWR_Q[:1]), (40320, CFG2_Q[:1])]


@pytest.mark.parametrize("m,moduli", LG)
def test_l_and_g_match_oracle(m, moduli):
    ab = backend()
    rng = np.random.default_rng(m + 1)
    plan = ab.Plan(m, moduli)
    x = rand_batch(rng, 3, moduli, plan.n)
    for name, fn in [("l", O.l), ("lInv", O.l_inv), ("mulGPow", O.mul_g_pow), ("mulGDec", O.mul_g_dec),
                     ("divGPow", O.div_g_pow), ("divGDec", O.div_g_dec), ("mulGCRT", O.mul_g_crt), ("divGCRT", O.div_g_crt)]:
        y = dev(x)
        getattr(ab.GT, name)(plan, y)
        assert (y.cpu().numpy() == oracle_limbs(fn, m, moduli, x)).all(), name
        if len(moduli) > 1:
            yi = dev(x.transpose(0, 2, 1))
            getattr(ab.GT, name)(plan, yi, layout=ab.INTERLEAVED)
            assert (yi.cpu().numpy().transpose(0, 2, 1) == oracle_limbs(fn, m, moduli, x)).all(), name


@pytest.mark.parametrize("m,moduli", [(512, ARITH_Q), (20475, RLWR_Q[:3])])
def test_pointwise_match_numpy(m, moduli):
    ab = backend()
    rng = np.random.default_rng(m + 2)
    plan = ab.Plan(m, moduli)
    qv = np.array(moduli, dtype=np.int64)[None, :, None]
    x, w = rand_batch(rng, 4, moduli, plan.n), rand_batch(rng, 4, moduli, plan.n)
    x[0, :, :3] = 0
    x[1, :, :] = qv - 1
    w[1, :, :] = qv - 1
    for name, want in [("mul", x * w % qv), ("add", (x + w) % qv), ("sub", (x - w) % qv)]:
        y = dev(x)
        getattr(ab.GT, name)(plan, y, dev(w))
        assert (y.cpu().numpy() == want).all(), name
        # broadcast of a single public element over the batch
        y = dev(x)
        getattr(ab.GT, name)(plan, y, dev(w[:1]))
        want_b = {"mul": x * w[:1] % qv, "add": (x + w[:1]) % qv, "sub": (x - w[:1]) % qv}[name]
        assert (y.cpu().numpy() == want_b).all(), name
    y = dev(x)
    ab.GT.neg(plan, y)
    assert (y.cpu().numpy() == (-x) % qv).all()
    sc = [int(q) - 2 for q in moduli]
    y = dev(x)
    ab.GT.scalarMul(plan, y, sc)
    assert (y.cpu().numpy() == x * np.array(sc)[None, :, None] % qv).all()


EXT = [(3, 9, [19]), (3, 21, [43]), (7, 21, [43, 127]), (5, 45, [181]), (9, 45, [181]), (4, 36, [37]), (12, 36, [37]),
       (1, 15, [31]), (8, 40, [41]), (21, 105, [211]), (4, 512, ARITH_Q), (128, 11648, TUNNEL_Q[:2]),
       (5824, 11648, TUNNEL_Q[:1]), (5824, 29120, TUNNEL_Q[:1]), (4095, 20475, RLWR_Q[:1]), (6825, 27300, RLWR_Q[:2])]


@pytest.mark.parametrize("m,mp,moduli", EXT)
def test_twace_embed_coeffs_match_oracle(m, mp, moduli):
    ab = backend()
    rng = np.random.default_rng(m * 7 + mp)
    lo, hi = ab.Plan(m, moduli), ab.Plan(mp, moduli)
    ext = ab.Ext(lo, hi)
    n, npr = lo.n, hi.n
    idx, base, tab = O.ext_indices_pow(m, mp), O.base_indices_crt(m, mp), O.coeff_indices(m, mp)
    assert ext.rel == npr // n
    x, y = rand_batch(rng, 3, moduli, n), rand_batch(rng, 3, moduli, npr)
    # embedPow: scatter with zero fill; twacePowDec: gather
    want = np.zeros_like(y)
    want[:, :, idx] = x
    assert (ab.GT.embedPow(ext, dev(x)).cpu().numpy() == want).all()
    assert (ab.GT.twacePowDec(ext, dev(y)).cpu().numpy() == y[:, :, idx]).all()
    # embedDec = lInv' . embedPow . l
    want = oracle_limbs(lambda mm, q, v: O.l_inv(mp, q, _scatter(O.l(m, q, v), idx, npr)), m, moduli, x)
    assert (ab.GT.embedDec(ext, dev(x)).cpu().numpy() == want).all()
    # embedCRT = replicate; twaceCRT = crt . twacePow . crtInv'
    assert (ab.GT.embedCRT(ext, dev(x)).cpu().numpy() == x[:, :, base]).all()
    want = oracle_limbs(lambda mm, q, v: O.crt(m, q, O.crt_inv(mp, q, v)[..., idx]), m, moduli, y)
    assert (ab.GT.twaceCRT(ext, dev(y)).cpu().numpy() == want).all()
    # coeffs / uncoeffs
    c = ab.GT.coeffs(ext, dev(y))
    assert (c.cpu().numpy() == y[:, :, tab].transpose(0, 2, 1, 3)).all()
    assert (ab.GT.uncoeffs(ext, c).cpu().numpy() == y).all()


def _scatter(v, idx, npr):
    out = np.zeros(v.shape[:-1] + (npr,), dtype=np.int64)
    out[..., idx] = v
    return out


def test_host_entry_interleaved_tuples():
    ab = backend()
    m, moduli = 512, ARITH_Q
    plan = ab.Plan(m, moduli)
    rng = np.random.default_rng(9)
    x = rand_batch(rng, 6, moduli, plan.n)
    host = np.ascontiguousarray(x.transpose(0, 2, 1))
    ab.GT.hostOp(plan, "crt", host)
    assert (host.transpose(0, 2, 1) == oracle_limbs(O.crt, m, moduli, x)).all()
    ab.GT.hostOp(plan, "crt_inv", host)
    assert (host.transpose(0, 2, 1) == x).all()


@pytest.mark.parametrize("switch", ["AB_POW2_SPLIT=0", "AB_POW2_PF_WAVES=0", "AB_POW2_PF_WAVES=1.5"])
def test_power_of_two_kernel_variants_compute_the_same_bits(switch):
    """the large power-of-two inverses run as two independent half-transform CTAs and every large kernel prefetches a later CTA's input
    into L2; the one-CTA inverse kernels and other prefetch distances must give the same bits (switches are read once per process)"""
    import os
    import subprocess
    import sys
    env = dict(os.environ)
    k, v = switch.split("=")
    env[k] = v
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                          "test_crt_and_inverse_match_oracle and (8192 or 16384 or 32768 or 65536)", "-p", "no:cacheprovider"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
