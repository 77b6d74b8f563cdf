"""Tier-1 for the ciphertext-level entries (ab_ct_mul_dev, ab_mod_switch_dev, ab_key_switch_quad_dev,
ab_tunnel_dev, ab_decompose_dev, ab_rescale_drop_dev): bit-exact against oracle/she_oracle.py, which
computes the same operations definitionally through the powerful basis (run under gpurun: `pytest -m gpu`).
Shapes follow the limb schedule of the reference's examples."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import she_oracle as S

pytestmark = pytest.mark.gpu

ARITH_Q = [268440577, 8392193, 1073750017]
RLWR_Q = [1543651201, 689270401, 718099201, 720720001, 1556755201, 1567238401]
TUNNEL_Q = [537264001, 539884801, 555609601, 560851201, 566092801]


def dev(x):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def rand(rng, lead, moduli, n):
    return np.stack([rng.integers(0, q, size=tuple(lead) + (n,), dtype=np.int64) for q in moduli], axis=-2)


@pytest.mark.parametrize("m,moduli,base", [(32, [97, 193, 257], 0), (32, [97, 193, 257], 2), (45, [181, 541], 3),
                                            (512, ARITH_Q[:2], 0), (20475, RLWR_Q[:3], 0), (11648, TUNNEL_Q[:1], 2)])
def test_decompose_and_rescale(m, moduli, base):
    import alchemy_b200 as ab
    rng = np.random.default_rng(m + base)
    plan = ab.Plan(m, moduli, crt=False)
    x = rand(rng, (3,), moduli, plan.n)
    x[0, :, 0] = [q - 1 for q in moduli]
    x[0, :, 1] = [q // 2 for q in moduli]
    x[0, :, 2] = [q // 2 + 1 for q in moduli]
    x[0, :, 3] = 0
    got = ab.GT.decompose(plan, dev(x), base=base).cpu().numpy()
    assert (got == S.decompose(x, moduli, moduli, base)).all()
    assert ab.GT.gadgetLen(plan, len(moduli), base) == S.gadget_len(moduli, base)
    if len(moduli) > 1:
        got = ab.GT.decompose(plan, dev(x), base=base, out_limbs=1).cpu().numpy()
        assert (got == S.decompose(x, moduli, moduli[:1], base)).all()
        for drop in range(len(moduli)):
            got = ab.GT.rescaleDrop(plan, dev(x), drop=drop).cpu().numpy()
            assert (got == S.rescale_drop(x, moduli, drop)).all()


MULS = [  # (m, moduli, L_in, L_hint, L_out, base): the limb schedule of SURVEY.md B.3
    (512, ARITH_Q, 2, 2, 1, 0),             # examples/Arithmetic.hs
    (20475, RLWR_Q, 4, 5, 3, 0),            # HomomRLWR level 1
    (20475, RLWR_Q, 3, 4, 2, 0),            # level 2
    (20475, RLWR_Q, 2, 3, 1, 0),            # level 3
    (20475, RLWR_Q, 1, 2, 1, 0),            # level 4
    (36, [37, 73, 109], 2, 3, 1, 0),
    (64, [193, 257, 449], 2, 2, 1, 4),      # BaseBGad on a product ring
    (45, [181, 541, 631], 2, 3, 1, 3),      # odd base, mixed-radix ring (fused digit load)
    (36, [37, 73, 109], 2, 2, 1, 2),        # binary gadget, mixed-radix ring
    (11648, TUNNEL_Q, 2, 2, 1, 2),          # Tunnel.hs ring with its BaseBGad 2
    (20475, RLWR_Q, 2, 2, 1, 16),
    # power-of-two rings, TrivGad: the single-kernel pipeline (ab_mulks_pow2.cu) incl. scale-up and multi-limb drops
    (64, [193, 257, 449, 577, 641], 2, 4, 1, 0),
    (64, [193, 257, 449, 577, 641], 3, 5, 2, 0),
    (64, [193, 257, 449], 1, 2, 1, 0),
    (64, [193, 257, 449], 3, 3, 3, 0),
    (4, [1073479681, 1052835841], 1, 2, 1, 0),          # n = 2
    (1024, [1073479681, 1052835841, 1032192001], 2, 3, 1, 0),
    (2048, [1073479681, 1052835841, 1032192001], 2, 3, 2, 0),
    (8192, [1073479681, 1052835841, 1032192001], 2, 3, 1, 0),   # too large for shared memory: the launch sequence behind the same entry
    (45, [181, 541, 631], 3, 2, 1, 0), (64, [193, 257, 449], 3, 2, 1, 0),   # hint modulus SMALLER than the input's: rescale down before the key switch
]


@pytest.mark.parametrize("m,moduli,lin,lh,lout,base", MULS)
def test_mul_keyswitch_pipeline(m, moduli, lin, lh, lout, base):
    """PT2CT.hs:172-177: modSwitch . keySwitchQuad hint . modSwitch . (*) -- every stage against the oracle"""
    import alchemy_b200 as ab
    from alchemy_b200 import she
    rng = np.random.default_rng(m + lin)
    plan = ab.Plan(m, moduli)
    n, B = plan.n, 3
    a, b = rand(rng, (B, 2), moduli[:lin], n), rand(rng, (B, 2), moduli[:lin], n)
    D = S.gadget_len(moduli[:lh], base)
    hint = rand(rng, (D, 2), moduli[:lh], n)
    c3 = she.ctMul(plan, dev(a), dev(b))
    w3 = S.ct_mul(m, moduli[:lin], a, b)
    assert (c3.cpu().numpy() == w3).all()
    # dec_mask 0: every polynomial rounded in the powerful basis, inputs taken as they are; dec_mask 1 with the toMSD scalars of a
    # plaintext modulus p = 7: Lol's modSwitch . keySwitchQuadCirc . modSwitch . (*) (c0 through the decoding basis)
    for dec_mask, pre in ((0, None), (1, [pow(7, -1, q) for q in moduli[:lin]])):
        w3m, c3m = w3, c3
        if pre is not None:
            w3m = w3 * np.array(pre, dtype=np.int64)[None, None, :, None] % np.array(moduli[:lin], dtype=np.int64)[None, None, :, None]
            c3m = c3.clone()
            ab.GT.scalarMul(plan, c3m.reshape(B * 3, lin, n), pre, nlimbs=lin)
        c3h = she.modSwitch(plan, c3m, lh, dec_mask)
        w3h = S.mod_switch(m, moduli, w3m, lh, dec_mask)
        assert (c3h.cpu().numpy() == w3h).all()
        c2 = she.keySwitchQuadCirc(plan, c3h, dev(hint), base)
        w2 = S.key_switch_quad(m, moduli, w3h, hint, base)
        assert (c2.cpu().numpy() == w2).all()
        out = she.modSwitch(plan, c2, lout, dec_mask)
        want = S.mod_switch(m, moduli, w2, lout, dec_mask)
        assert (out.cpu().numpy() == want).all()
        # and the one-call entry gives the same thing
        assert (she.mulKeySwitch(plan, dev(a), dev(b), dev(hint), lh, lout, base, prescale=pre, dec_mask=dec_mask).cpu().numpy() == want).all()


@pytest.mark.parametrize("m,moduli,frm,to", [(512, ARITH_Q, 3, 1), (512, ARITH_Q, 1, 3), (20475, RLWR_Q, 6, 2), (20475, RLWR_Q, 5, 5),
                                              (45, [181, 541, 631], 3, 2), (11648, TUNNEL_Q, 2, 1),
                                              # several limbs dropped in one rescale (Garner digits in the load phase of the transform)
                                              (20475, RLWR_Q, 5, 3), (20475, RLWR_Q, 6, 1), (45, [181, 541, 631], 3, 1), (43680, RLWR_Q, 6, 3),
                                              (11648, TUNNEL_Q, 5, 2), (36, [37, 73, 109], 3, 1),
                                              # a prime above 13: outside the shared-memory kernel, launch sequence with one pass per polynomial
                                              (51, [103, 307, 409], 3, 1), (51, [103, 307, 409], 2, 1)])
def test_mod_switch(m, moduli, frm, to):
    import alchemy_b200 as ab
    from alchemy_b200 import she
    rng = np.random.default_rng(m + frm * 10 + to)
    plan = ab.Plan(m, moduli)
    c = rand(rng, (4, 2), moduli[:frm], plan.n)
    c[0, :, :, 0] = np.array(moduli[:frm]) - 1
    got = she.modSwitch(plan, dev(c), to).cpu().numpy()
    assert (got == S.mod_switch(m, moduli, c, to)).all()
    # rounding bases per polynomial: c0 in the decoding basis (Lol rescaleLinearMSD), c1, both; a degree-2 ciphertext with c0 and c2
    for dec_mask in (1, 2, 3):
        assert (she.modSwitch(plan, dev(c), to, dec_mask).cpu().numpy() == S.mod_switch(m, moduli, c, to, dec_mask)).all(), dec_mask
    c3 = rand(rng, (3, 3), moduli[:frm], plan.n)
    for dec_mask in (1, 5):
        assert (she.modSwitch(plan, dev(c3), to, dec_mask).cpu().numpy() == S.mod_switch(m, moduli, c3, to, dec_mask)).all(), dec_mask
    if frm == to:       # the host mirror returns its input for equal moduli; the ABI entry itself copies
        import ctypes as C

        import torch
        cd, out = dev(c), torch.zeros((4, 2, to, plan.n), dtype=torch.int64, device="cuda")
        ab.check(ab.lib().ab_mod_switch_dev(plan.h, C.c_void_p(out.data_ptr()), C.c_void_p(cd.data_ptr()), 4, 2, frm, to, 1, None, None))
        assert torch.equal(out, cd)


TUNNELS = [  # (r', e', s', moduli, L, base)
    (12, 4, 20, [61, 181, 241], 2, 0), (9, 3, 21, [127, 379], 2, 3), (45, 15, 15, [181, 541], 1, 2), (16, 16, 48, [97, 193], 2, 0),
    (11648, 5824, 29120, TUNNEL_Q, 1, 2),        # examples/Tunnel.hs hop 1 -> 2 (BaseBGad 2, one limb)
    (29120, 14560, 43680, RLWR_Q, 2, 0),         # HomomRLWR hop 2 (TrivGad), 2 of its limbs
    (43680, 10920, 54600, RLWR_Q, 1, 0),         # hop 3: dim 4
    (54600, 27300, 27300, RLWR_Q, 2, 0),         # hop 4: s' = e'
    (27300, 6825, 20475, RLWR_Q, 1, 0),          # hop 5
]


@pytest.mark.parametrize("r,e,s,moduli,L,base", TUNNELS)
def test_tunnel(r, e, s, moduli, L, base):
    import alchemy_b200 as ab
    from alchemy_b200 import she
    rng = np.random.default_rng(r + s)
    q = moduli[:L]
    pr, ps = ab.Plan(r, q), ab.Plan(s, q)
    pe = ab.Plan(e, q, crt=False)
    tn = she.Tunnel(pr, pe, ps)
    K = pr.n // pe.n
    D = S.gadget_len(q, base)
    B = 2
    c = rand(rng, (B, 2), q, pr.n)
    flin = rand(rng, (K,), q, ps.n)
    hint = rand(rng, (K, D, 2), q, ps.n)
    got = tn(dev(c), dev(flin), dev(hint), base).cpu().numpy()
    want = S.tunnel(r, e, s, q, c, flin, hint, base)
    assert (got == want).all()


def test_full_size_roundtrip_properties():
    """size-independent checks at a BASELINE-size batch: crtInv.crt = id, linearity of key switching in (c0,c1)"""
    import torch

    import alchemy_b200 as ab
    from alchemy_b200 import she
    m, moduli = 20475, RLWR_Q[:4]
    plan = ab.Plan(m, moduli)
    n, B = plan.n, 512
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.stack([torch.randint(0, q, (B, n), dtype=torch.int64, device="cuda", generator=g) for q in moduli], dim=1)
    y = x.clone()
    ab.GT.crt(plan, y)
    assert not torch.equal(x, y)
    ab.GT.crtInv(plan, y)
    assert torch.equal(x, y)
    # key switch: out(c0 + d0, c1, c2) - out(c0, c1, c2) == d0 exactly
    D = S.gadget_len(moduli, 0)
    hint = torch.stack([torch.randint(0, q, (D, 2, n), dtype=torch.int64, device="cuda", generator=g) for q in moduli], dim=2)
    c3 = torch.stack([torch.randint(0, q, (64, 3, n), dtype=torch.int64, device="cuda", generator=g) for q in moduli], dim=2)
    d0 = torch.stack([torch.randint(0, q, (64, n), dtype=torch.int64, device="cuda", generator=g) for q in moduli], dim=1)
    qv = torch.tensor(moduli, device="cuda")[None, :, None]
    o1 = she.keySwitchQuadCirc(plan, c3, hint)
    c3b = c3.clone()
    c3b[:, 0] = (c3b[:, 0] + d0) % qv
    o2 = she.keySwitchQuadCirc(plan, c3b, hint)
    assert torch.equal((o2[:, 0] - o1[:, 0]) % qv, d0) and torch.equal(o2[:, 1], o1[:, 1])


def test_empty_batches_and_bad_arguments_of_the_fused_entries():
    """batch = 0 is a no-op with status 0; null pointers / limb counts outside the plan are AB_ERR_ARG, never a crash"""
    import ctypes as C

    import torch

    import alchemy_b200 as ab
    lib = ab.lib()
    plan = ab.Plan(45, [181, 541, 631])
    z = torch.zeros(8, dtype=torch.int64, device="cuda")
    p = C.c_void_p(z.data_ptr())
    assert lib.ab_mul_key_switch_dev(plan.h, p, p, p, p, 0, 2, 3, 1, 0, None, 0, None, None) == 0
    assert lib.ab_ct_phase_dev(plan.h, p, p, p, 0, 2, 2, None) == 0
    assert lib.ab_rns_lift_reduce_dev(plan.h, p, None, p, 0, 2, 7, None) == 0
    assert lib.ab_lift_plain_dev(plan.h, p, p, 0, 2, 7, None) == 0
    assert lib.ab_ct_add_public_dev(plan.h, p, p, 0, 2, 2, None) == 0 and lib.ab_ct_mul_public_dev(plan.h, p, p, 0, 2, 2, None) == 0
    assert lib.ab_uniform_dev(plan.h, p, 0, 2, 1, 0, None) == 0
    assert lib.ab_mul_key_switch_dev(plan.h, None, p, p, p, 1, 2, 3, 1, 0, None, 0, p, None) == ab._lib.AB_ERR_ARG
    assert lib.ab_mul_key_switch_dev(plan.h, p, p, p, p, 1, 2, 4, 1, 0, None, 0, p, None) == ab._lib.AB_ERR_ARG          # 4 limbs > plan
    assert lib.ab_mul_key_switch_dev(plan.h, p, p, p, p, 1, 2, 3, 1, 0, None, 0, None, None) == ab._lib.AB_ERR_ARG       # scratch missing
    assert lib.ab_rns_lift_reduce_dev(plan.h, p, None, p, 1, 2, 1, None) == ab._lib.AB_ERR_ARG                  # modulus 1
    assert lib.ab_gaussian_dev(plan.h, p, 1, None, 4, 1, 0, None) == ab._lib.AB_ERR_ARG
    assert lib.ab_mul_key_switch_work_bytes(plan.h, 4, 2, 3, 1, 0) > 0 and lib.ab_mul_key_switch_work_bytes(plan.h, 4, 2, 9, 1, 0) == -1
    # composite plaintext moduli: L / mulG work, divG refuses when an odd prime of m is not invertible
    p9 = ab.Plan(45, [9], crt=False)
    y = torch.arange(p9.n, dtype=torch.int64, device="cuda").reshape(1, 1, -1) % 9
    keep = y.clone()
    ab.GT.lInv(p9, ab.GT.l(p9, y))
    assert torch.equal(y, keep)
    with pytest.raises(ab.AlchemyError):
        ab.GT.divGPow(p9, y)


@pytest.mark.parametrize("switch", ["AB_NO_FUSE=1", "AB_NO_MAC=1", "AB_NO_BITS=1", "AB_NO_MULTIDROP=1", "AB_NO_MULKS=1", "AB_NO_MULKS256=1", "AB_MX_MAC_PARTS=5", "AB_MX_T=128", "AB_MX_MAC_T=256",
                                    "AB_NO_SUBRING=1", "AB_MX_MAC_GROUP=1", "AB_MX_MAC_GROUP=3"])
def test_every_kernel_variant_computes_the_same_bits(switch):
    """the A/B switches of DESIGN.md section 5 select alternative kernels / launch sequences; each must pass the same parity tests
    (the switches are read once per process, hence the subprocess)"""
    import subprocess
    import sys
    env = dict(os.environ)
    k, v = switch.split("=")
    env[k] = v
    here = os.path.dirname(os.path.abspath(__file__))
    sel = "mul_keyswitch_pipeline or test_mod_switch or test_tunnel"
    out = subprocess.run([sys.executable, "-m", "pytest", os.path.join(here, "test_gpu_she.py"), "-x", "-q", "-m", "gpu", "-k", sel, "-p", "no:cacheprovider"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_arithmetic_mul_one_warp_kernel_on_a_ragged_batch():
    """examples/Arithmetic.hs `mul_` (n = 256, limbs 2 -> 2 -> 1) runs as one warp per op (ab_mulks_n256.cu): a batch that does not
    fill the last CTA, against the stage-by-stage calls (themselves checked against the CPU restatement above)"""
    import torch

    import alchemy_b200 as ab
    from alchemy_b200 import she
    rng = np.random.default_rng(256)
    plan = ab.Plan(512, ARITH_Q)
    n, B = plan.n, 4099
    a, b, hint = dev(rand(rng, (B, 2), ARITH_Q[:2], n)), dev(rand(rng, (B, 2), ARITH_Q[:2], n)), dev(rand(rng, (2, 2), ARITH_Q[:2], n))
    for pre in (None, [pow(7, -1, q) for q in ARITH_Q[:2]]):
        c3 = she.ctMul(plan, a, b)
        if pre is not None:
            ab.GT.scalarMul(plan, c3.reshape(B * 3, 2, n), pre, nlimbs=2)
        want = she.modSwitch(plan, she.keySwitchQuadCirc(plan, c3, hint, 0), 1, 1)
        got = she.mulKeySwitch(plan, a, b, hint, 2, 1, 0, prescale=pre, dec_mask=1)
        assert torch.equal(got, want)
    # extreme residues: 0 and q - 1 everywhere
    top = torch.tensor([q - 1 for q in ARITH_Q[:2]], dtype=torch.int64, device="cuda").reshape(1, 1, 2, 1).expand(5, 2, 2, n).contiguous()
    for x, y in ((top, top), (torch.zeros_like(top), top)):
        want = she.modSwitch(plan, she.keySwitchQuadCirc(plan, she.ctMul(plan, x, y), top[:2].contiguous(), 0), 1, 0)
        assert torch.equal(she.mulKeySwitch(plan, x, y, top[:2].contiguous(), 2, 1, 0), want)


def test_chains_replay_from_a_cuda_graph():
    """every *_dev entry is asynchronous on the caller's stream with caller-owned scratch, so a chain of them can be captured once
    into a CUDA graph and replayed (small batches are launch-bound: examples/Tunnel.hs on one ciphertext is 20 launches): a tunnelling
    hop followed by a rescale, and the one-call `mul_` pipeline, replayed on new inputs, give the bits of the eager calls"""
    import torch

    import alchemy_b200 as ab
    from alchemy_b200 import she
    rng = np.random.default_rng(77)
    r, e, s_idx, moduli, L, base = 11648, 5824, 29120, TUNNEL_Q, 2, 2
    pr, pe, ps = ab.Plan(r, moduli), ab.Plan(e, moduli, crt=False), ab.Plan(s_idx, moduli)
    tn = she.Tunnel(pr, pe, ps)
    K, D, B = pr.n // pe.n, S.gadget_len(moduli[:L], base), 2
    flin, hint = dev(rand(rng, (K,), moduli[:L], ps.n)), dev(rand(rng, (K, D, 2), moduli[:L], ps.n))
    p5 = ab.Plan(20475, RLWR_Q)
    kh = dev(rand(rng, (3, 2), RLWR_Q[:3], p5.n))
    pre = [pow(2, -1, q) for q in RLWR_Q[:2]]

    def chain(c, a, b):
        return she.modSwitch(ps, tn(c, flin, hint, base), 1, 1), she.mulKeySwitch(p5, a, b, kh, 3, 1, 0, prescale=pre, dec_mask=1)

    def inputs():
        return dev(rand(rng, (B, 2), moduli[:L], pr.n)), dev(rand(rng, (B, 2), RLWR_Q[:2], p5.n)), dev(rand(rng, (B, 2), RLWR_Q[:2], p5.n))

    c, a, b = inputs()
    chain(c, a, b)                                   # warm-up outside the capture (lazy kernel attributes, allocator)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = chain(c, a, b)
    for _ in range(2):
        c2, a2, b2 = inputs()
        c.copy_(c2); a.copy_(a2); b.copy_(b2)
        g.replay()
        torch.cuda.synchronize()
        want = chain(c2, a2, b2)
        assert torch.equal(out[0], want[0]) and torch.equal(out[1], want[1])
