"""Host-side mirror of Lol's SymmSHE as ALCHEMY drives it: ciphertext bookkeeping (encoding flag, g-power k,
scalar l) around the device kernels.  Reference call sites: encrypt / decrypt Crypto/Alchemy/Interpreter/
PT2CT.hs:78-99; add_/mul_ Eval.hs:58-67; modSwitchPT_/modSwitch_/addPublic_/mulPublic_/keySwitchQuad_
Eval.hs:129-133; errorRate_ Eval.hs:155-160; hint generation Crypto/Alchemy/Interpreter/KeysHints.hs:96,113.
SymmSHE itself lives in the absent `lol-apps` package ([RECALL]); the conventions this
file fixes (S1-S7) are listed in DESIGN.md section 8; the test suite restates the same operations on the CPU.

Every ring operation below is a CUDA kernel behind the C ABI; Python only carries the three scalars
(enc, k, l) per ciphertext *type* -- in ALCHEMY they are compile-time data of the typed DSL.
Randomness (uniform c1 / hint masks, small error terms, the secret) is passed in by the caller.
"""
import ctypes as C

import torch

from . import she
from ._lib import check, lib
from .tensor import GT, Ext, Plan, _ptr, _stream

LIFT_ALL = 1 << 40          # "plaintext modulus" that leaves small signed integers as they are


class CT:
    """Lol `CT m zp (Cyc t m' zq)`: encoding (LSD/MSD), power k of g, scalar l (mod p), polynomial in the secret.
    c: int64 CUDA tensor [B, degree+1, L, n'] in the CRT basis over the first L moduli."""

    def __init__(self, c, k=0, l=1, enc="LSD"):
        self.c, self.k, self.l, self.enc = c, int(k), int(l), enc

    @property
    def L(self):
        return self.c.shape[2]


class SymmSHE:
    # which polynomials modSwitch rounds in the decoding basis (bit i = c_i): Lol's rescaleLinearMSD takes c0 through rescaleDec and
    # c1 through rescalePow [RECALL; generalised here to "c0 Dec, every further coefficient Pow" for the degree-2 ciphertext that
    # PT2CT.hs:177 rescales before the key switch].  0 = every polynomial through the powerful basis.
    rescale_dec_mask = 1

    def __init__(self, m, mp, moduli, p):
        assert mp % m == 0
        self.m, self.mp, self.moduli, self.p = m, mp, [int(q) for q in moduli], int(p)
        self.hi = Plan(mp, moduli)                        # O_m' over Z_q, CRT tables
        self.lo = Plan(m, moduli, crt=False)              # O_m over the same moduli: source of embed
        self.ext = Ext(self.lo, self.hi)
        self.hi_p = Plan(mp, [p], crt=False)              # O_m' over Z_p: where decrypt divides by g and takes the twace
        self.lo_p = Plan(m, [p], crt=False)
        self.ext_p = Ext(self.lo_p, self.hi_p)

    # ---- helpers ---------------------------------------------------------------------------
    def _lift_plain(self, plan, x, nlimbs, pmod):
        """[B, n] integers -> [B, L, n]: lift_pmod(x) mod q_l"""
        B, n = x.shape
        assert n == plan.n
        out = torch.empty((B, nlimbs, n), dtype=torch.int64, device=x.device)
        check(lib().ab_lift_plain_dev(plan.h, _ptr(out), _ptr(x.contiguous()), B, nlimbs, pmod, _stream()))
        return out

    def _scale(self, c, scalar_of_q):
        """multiply every polynomial of c[B,P,L,n] by a per-limb scalar"""
        B, P, L, n = c.shape
        y = c.reshape(B * P, L, n)
        GT.scalarMul(self.hi, y, [scalar_of_q(q) % q for q in self.moduli[:L]], nlimbs=L)
        return c

    def _public_crt(self, b, k, l, L):
        """public plaintext b[n] (mod p, powerful basis of O_m) in the form of an LSD ciphertext with (k, l):
        g_{m'}^k * embed(reduce(lift(b / l))), CRT basis of O_m' -> [1, L, n']"""
        b = (b.to(torch.int64) * pow(l, -1, self.p)) % self.p
        x = self._lift_plain(self.lo, b.reshape(1, -1), L, self.p)
        x = GT.embedPow(self.ext, x, nlimbs=L)
        for _ in range(k):                                  # the ciphertext carries g_{m'}^k: multiply in O_m'
            GT.mulGPow(self.hi, x, nlimbs=L)
        return GT.crt(self.hi, x, nlimbs=L)

    # ---- keys --------------------------------------------------------------------------------
    def genSK(self, s_dec):
        """secret key from small integer coefficients in the decoding basis of O_m' -> CRT residues [L, n']"""
        x = self._lift_plain(self.hi, s_dec.reshape(1, -1), self.hi.L, LIFT_ALL)
        return GT.crt(self.hi, GT.l(self.hi, x))[0]

    def ksQuadCircHint(self, sk, mask, noise_dec, L, base=0):
        """hint for s^2 -> s over the first L moduli: hint_d = (gamma_d s^2 - a_d s + p e_d, a_d)
        mask a[D, L, n'] uniform (CRT basis), noise_dec[D, n'] small integers (decoding basis) -> [D, 2, L, n']"""
        D = GT.gadgetLen(self.hi, L, base)
        assert tuple(mask.shape) == (D, L, self.hi.n) and tuple(noise_dec.shape) == (D, self.hi.n)
        s = sk[:L].unsqueeze(0).contiguous()
        e = GT.l(self.hi, self._lift_plain(self.hi, noise_dec, L, LIFT_ALL), nlimbs=L)
        GT.scalarMul(self.hi, e, [self.p % q for q in self.moduli[:L]], nlimbs=L)
        h0 = GT.crt(self.hi, e, nlimbs=L)                              # p e_d
        t = mask.clone()
        GT.mul(self.hi, t, s, nlimbs=L)
        GT.sub(self.hi, h0, t, nlimbs=L)                               # - a_d s
        s2 = s.clone()
        GT.mul(self.hi, s2, s, nlimbs=L)
        d, n = 0, self.hi.n
        for l, q in enumerate(self.moduli[:L]):                        # + gamma_d s^2: gamma_d lives in limb l only
            for i in range(_digits(q, base)):
                g = s2[:, l:l + 1].clone()
                GT.scalarMul(self.hi, g, [1 if base == 0 else pow(base, i, q)], limb0=l, nlimbs=1)
                GT.add(self.hi, h0[d, l].view(1, 1, n), g, limb0=l, nlimbs=1)
                d += 1
        return torch.stack([h0, mask], dim=1).contiguous()

    # ---- the same with device randomness (KeysHints.hs:96 getKey / genSK, :113 getQuadCircHint) -------------------
    # stream ids: 0 secret key, 1 encryption noise, 2 encryption mask, 3 hint noise, 4 hint mask; + 8 * nonce.
    # TEST AND BENCHMARK RANDOMNESS ONLY: Philox4x32-10 is a statistical generator with a 64-bit key, not a CSPRNG, and every draw
    # is a pure function of (seed, stream id, position) -- two calls with the same (seed, nonce) produce the SAME masks and noise.
    # Give every call of one key's lifetime its own nonce (or seed); production keys need a cryptographic generator behind the
    # same entry points (the kernels take masks and noise as inputs: encrypt / ksQuadCircHint / tunnelHint above and below).
    def sampleSK(self, cdt, seed, nonce=0):
        return self.genSK(GT.gaussianDec(self.hi, 1, cdt, seed, 0 + 8 * nonce)[0])

    def sampleKSQuadCircHint(self, sk, L, cdt, seed, base=0, nonce=0):
        D = GT.gadgetLen(self.hi, L, base)
        return self.ksQuadCircHint(sk, GT.uniform(self.hi, D, L, seed, 4 + 8 * nonce), GT.gaussianDec(self.hi, D, cdt, seed, 3 + 8 * nonce), L, base)

    def encryptRandom(self, sk, mu, L, cdt, seed, nonce=0):
        B = mu.shape[0]
        return self.encrypt(sk, mu, GT.gaussianDec(self.hi, B, cdt, seed, 1 + 8 * nonce), GT.uniform(self.hi, B, L, seed, 2 + 8 * nonce))

    # ---- encrypt / decrypt ---------------------------------------------------------------------
    def encrypt(self, sk, mu, noise_dec, c1):
        """mu[B, n] plaintexts mod p (powerful basis of O_m); noise_dec[B, n'] small integers (decoding basis of
        O_m'); c1[B, L, n'] uniform residues (CRT basis).  e = embed(lift mu) + p noise; c0 = e - c1 s."""
        L = c1.shape[1]
        x = GT.embedPow(self.ext, self._lift_plain(self.lo, mu, L, self.p), nlimbs=L)
        e = GT.l(self.hi, self._lift_plain(self.hi, noise_dec, L, LIFT_ALL), nlimbs=L)
        GT.scalarMul(self.hi, e, [self.p % q for q in self.moduli[:L]], nlimbs=L)
        GT.add(self.hi, e, x, nlimbs=L)
        c0 = GT.crt(self.hi, e, nlimbs=L)
        t = c1.clone()
        GT.mul(self.hi, t, sk[:L].unsqueeze(0).contiguous(), nlimbs=L)
        GT.sub(self.hi, c0, t, nlimbs=L)
        return CT(torch.stack([c0, c1], dim=1).contiguous(), 0, 1, "LSD")

    def errorTerm(self, sk, ct, pmod, want_rate=False):
        """lift (evaluate c s) in the decoding basis, reduced mod pmod: [B, n'] (and max |e| / Q per ciphertext)"""
        ct = self.toLSD(ct)
        B, P, L, n = ct.c.shape
        ph = torch.empty((B, L, n), dtype=torch.int64, device=ct.c.device)
        check(lib().ab_ct_phase_dev(self.hi.h, _ptr(ph), _ptr(ct.c), _ptr(sk[:L].contiguous()), B, P, L, _stream()))
        GT.lInv(self.hi, GT.crtInv(self.hi, ph, nlimbs=L), nlimbs=L)
        out = torch.empty((B, n), dtype=torch.int64, device=ct.c.device)
        rate = torch.zeros(B, dtype=torch.float64, device=ct.c.device) if want_rate else None
        check(lib().ab_rns_lift_reduce_dev(self.hi.h, _ptr(out), C.c_void_p(rate.data_ptr()) if want_rate else None, _ptr(ph), B, L, pmod, _stream()))
        return (out, rate) if want_rate else out

    def errorRate(self, sk, ct):
        return self.errorTerm(sk, ct, self.p, want_rate=True)[1]

    def decrypt(self, sk, ct):
        """l * twace(divG^k (reduce (errorTerm ct))) -> [B, n] mod p, powerful basis of O_m"""
        ct = self.toLSD(ct)
        e = self.errorTerm(sk, ct, self.p).unsqueeze(1)                # [B, 1, n'] over Z_p, decoding basis
        for _ in range(ct.k):
            GT.divGDec(self.hi_p, e)
        x = GT.twacePowDec(self.ext_p, e)
        GT.scalarMul(self.lo_p, x, [ct.l % self.p])
        return GT.l(self.lo_p, x)[:, 0]

    # ---- encodings -----------------------------------------------------------------------------
    def _Q(self, L):
        Q = 1
        for q in self.moduli[:L]:
            Q *= q
        return Q

    def toMSD(self, ct):
        if ct.enc == "MSD":
            return ct
        p = self.p
        c = self._scale(ct.c.clone(), lambda q: pow(p, -1, q))
        return CT(c, ct.k, ct.l * (-self._Q(ct.L)) % p, "MSD")

    def toLSD(self, ct):
        if ct.enc == "LSD":
            return ct
        p = self.p
        c = self._scale(ct.c.clone(), lambda q: p)
        return CT(c, ct.k, ct.l * pow(-self._Q(ct.L), -1, p) % p, "LSD")

    # ---- ring operations on ciphertexts -------------------------------------------------------------
    def _align(self, a, b):
        """same encoding, g-power and scalar (Lol's Additive instance for CT)"""
        if a.enc != b.enc:
            a, b = self.toLSD(a), self.toLSD(b)
        if a.k < b.k:
            a, b = b, a
        if b.k < a.k:
            c = b.c.clone()
            B, P, L, n = c.shape
            for _ in range(a.k - b.k):
                GT.mulGCRT(self.hi, c.reshape(B * P, L, n), nlimbs=L)
            b = CT(c, a.k, b.l, b.enc)
        if a.l != b.l:
            # b carries l_b^-1 mu_b: multiply by lift(l_b / l_a) so that it carries l_a^-1 mu_b
            f = b.l * pow(a.l, -1, self.p) % self.p
            f = f - self.p if 2 * f >= self.p else f
            b = CT(self._scale(b.c.clone(), lambda q: f), b.k, a.l, b.enc)
        return a, b

    def add(self, a, b):
        a, b = self._align(a, b)
        Pa, Pb = a.c.shape[1], b.c.shape[1]
        if Pa < Pb:
            a, b, Pa, Pb = b, a, Pb, Pa
        c = a.c.clone()
        B, _, L, n = c.shape
        lo = c[:, :Pb].reshape(B * Pb, L, n) if Pa == Pb else None
        if lo is not None:
            GT.add(self.hi, lo, b.c.reshape(B * Pb, L, n), nlimbs=L)
        else:
            for i in range(Pb):
                t = c[:, i].contiguous()
                GT.add(self.hi, t, b.c[:, i].contiguous(), nlimbs=L)
                c[:, i] = t
        return CT(c, a.k, a.l, a.enc)

    def mul(self, a, b):
        a, b = self.toLSD(a), self.toLSD(b)
        return CT(she.ctMul(self.hi, a.c, b.c), a.k + b.k + 1, a.l * b.l % self.p, "LSD")

    def modSwitch(self, ct, to_limbs):
        ct = self.toMSD(ct)
        return CT(she.modSwitch(self.hi, ct.c, to_limbs, self.rescale_dec_mask), ct.k, ct.l, "MSD")

    def mulKeySwitch(self, a, b, hint, l_hint, l_out, base=0):
        """PT2CT `mul_` (PT2CT.hs:172-177) = modSwitch . keySwitchQuad hint . modSwitch . (*) as ONE ABI call.  (*) works on LSD
        forms and yields an LSD ciphertext; the first modSwitch begins with toMSD (polynomials times p^-1 mod q_l, l times -Q_in
        mod p): the scalars ride along in the product kernel.  Bit-identical to self.modSwitch(self.keySwitchQuad(self.modSwitch(
        self.mul(a, b), l_hint), hint, base), l_out)."""
        a, b = self.toLSD(a), self.toLSD(b)
        assert a.L == b.L
        pre = [pow(self.p, -1, q) for q in self.moduli[:a.L]]
        c = she.mulKeySwitch(self.hi, a.c, b.c, hint, l_hint, l_out, base, prescale=pre, dec_mask=self.rescale_dec_mask)
        return CT(c, a.k + b.k + 1, a.l * b.l * (-self._Q(a.L)) % self.p, "MSD")

    def keySwitchQuad(self, ct, hint, base=0):
        ct = self.toMSD(ct)
        return CT(she.keySwitchQuadCirc(self.hi, ct.c, hint, base), ct.k, ct.l, "MSD")

    def modSwitchPT(self, ct, p_new):
        """Lol modSwitchPT (div2_ when p / p_new = 2): the polynomial is untouched in MSD form; the scalar is re-read mod p'.
        Returns the ciphertext and the scheme object for the smaller plaintext modulus."""
        assert self.p % p_new == 0
        ct = self.toMSD(ct)
        return CT(ct.c, ct.k, ct.l % p_new, "MSD")

    def addPublic(self, ct, b):
        ct = self.toLSD(ct)
        pub = self._public_crt(b, ct.k, ct.l, ct.L)
        c = ct.c.clone()
        B, P, L, n = c.shape
        check(lib().ab_ct_add_public_dev(self.hi.h, _ptr(c), _ptr(pub), B, P, L, _stream()))
        return CT(c, ct.k, ct.l, "LSD")

    def mulPublic(self, ct, a):
        pub = self._public_crt(a, 0, 1, ct.L)
        c = ct.c.clone()
        B, P, L, n = c.shape
        check(lib().ab_ct_mul_public_dev(self.hi.h, _ptr(c), _ptr(pub), B, P, L, _stream()))
        return CT(c, ct.k, ct.l, ct.enc)


def _digits(q, base):
    if base == 0:
        return 1
    k, v = 0, 1
    while v < q:
        v *= base
        k += 1
    return k


def tunnelHint(tn, ext_er, moduli, s_in_crt, s_out_crt, flin, mask, noise_dec, base=0):
    """Lol `tunnelHint` (KeysHints.hs:113): hint[k, d] = (b, a) with b + a s_out = gamma_d * f(relpow_k * s_in) + e_{k,d}.
    tn: she.Tunnel for (r', e', s'); ext_er: Ext(e', r') over the same moduli; s_in_crt [L, n_r], s_out_crt [L, n_s] (CRT);
    flin [K, L, n_s]; mask [K, D, L, n_s] uniform; noise_dec [K, D, n_s] small integers (decoding basis of O_s').
    f(relpow_k * s_in) is obtained from the tunnel kernel itself: tunnelling the degree-0 ciphertext (x, 0) yields (f(x), 0)."""
    pr, ps = tn.r, tn.s
    K, D, L, ns = mask.shape
    nr = pr.n
    dev = mask.device
    # relpow_k = uncoeffs of the k-th unit vector over O_e' (powerful basis), times s_in, as degree-0 ciphertexts
    onehot = torch.zeros((K, K, L, ext_er.lo.n), dtype=torch.int64, device=dev)
    for k in range(K):
        onehot[k, k, :, 0] = 1
    rel = GT.uncoeffs(ext_er, onehot, nlimbs=L)                       # [K, L, n_r]
    GT.crt(pr, rel, nlimbs=L)
    GT.mul(pr, rel, s_in_crt[:L].unsqueeze(0).contiguous(), nlimbs=L)
    c = torch.zeros((K, 2, L, nr), dtype=torch.int64, device=dev)
    c[:, 0] = rel
    zero_hint = torch.zeros((K, D, 2, L, ns), dtype=torch.int64, device=dev)
    target = tn(c, flin, zero_hint, base)[:, 0].contiguous()           # [K, L, n_s] = f(relpow_k s_in), CRT basis of O_s'
    # b = e - a s_out (+ gamma_d * target in the limb the gadget element lives in)
    e = torch.empty((K * D, L, ns), dtype=torch.int64, device=dev)
    check(lib().ab_lift_plain_dev(ps.h, _ptr(e), _ptr(noise_dec.reshape(K * D, ns).contiguous()), K * D, L, LIFT_ALL, _stream()))
    GT.crt(ps, GT.l(ps, e, nlimbs=L), nlimbs=L)
    t = mask.reshape(K * D, L, ns).clone()
    GT.mul(ps, t, s_out_crt[:L].unsqueeze(0).contiguous(), nlimbs=L)
    GT.sub(ps, e, t, nlimbs=L)
    b = e.reshape(K, D, L, ns)
    d = 0
    for l, q in enumerate(moduli[:L]):
        for i in range(_digits(q, base)):
            g = target[:, l:l + 1].clone()                             # [K, 1, n_s]
            GT.scalarMul(ps, g, [1 if base == 0 else pow(base, i, q)], limb0=l, nlimbs=1)
            row = b[:, d, l].contiguous().view(K, 1, ns)
            GT.add(ps, row, g, limb0=l, nlimbs=1)
            b[:, d, l] = row[:, 0]
            d += 1
    return torch.stack([b, mask], dim=2).contiguous()
