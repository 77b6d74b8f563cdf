"""CPU restatement of the SymmSHE boundary ALCHEMY drives: encrypt / decrypt (Crypto/Alchemy/Interpreter/
PT2CT.hs:78-99), add_/mul_ (Eval.hs:58-67), modSwitchPT_/modSwitch_/addPublic_/mulPublic_/keySwitchQuad_
(Eval.hs:129-133), errorRate_ (Eval.hs:155-160) and the key-switch hint of KeysHints.hs:96.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: SymmSHE is in the absent `lol-apps` package; this file restates the
scheme of Crockett-Peikert ("Lambda o lambda", section 6) with python big integers for every cross-limb step
(the CUDA path uses Garner mixed-radix arithmetic instead) and is pinned by the reference's own criterion:
decrypt (eval ct) == eval pt (examples/Arithmetic.hs:75, examples/HomomRLWR.hs:71).

Conventions fixed here ([RECALL]; each would have to be checked against lol-apps):
  S1  LSD ciphertext (k, l): c(s) = e with e = l^-1 g^k mu (mod p), e short.  encrypt: k = 0, l = 1,
      e = embed(lift mu) + p * noise, noise given in the decoding basis; c0 = e - c1 s.
  S2  MSD <-> LSD: LSD -> MSD multiplies c by p^-1 mod q and l by (-q) mod p; MSD -> LSD multiplies c by p and
      l by (-q)^-1 mod p.
  S3  errorTerm lifts c(s) in the DECODING basis, centred modulo the full RNS product.
  S4  decrypt = l * twace(divG^k(reduce(errorTerm))) computed in the decoding basis over Z_p.
  S5  (*) multiplies LSD forms: k = k1 + k2 + 1, l = l1 l2, every coefficient times g.
  S6  (+) aligns encodings (to LSD), g-powers (mulG on the smaller k) and scalars (multiply by lift(l_b / l_a)).
  S7  modSwitch and keySwitchQuadCirc act on MSD forms; the key-switch hint satisfies
      h0_d + h1_d s = gamma_d s^2 + p e_d.
  S8  modSwitch rounds c0 in the decoding basis and every further coefficient in the powerful basis (Lol rescaleLinearMSD:
      rescaleDec c0, rescalePow c1; extended to the degree-2 ciphertext PT2CT.hs:177 rescales before the key switch).
      gadget decomposition of c2 (key switch) and of the E'-coefficients of c1 (tunnel) is taken in the powerful basis.
"""
from math import prod

import numpy as np

from . import oracle as O
from . import she_oracle as S


class CT:
    def __init__(self, c, k=0, l=1, enc="LSD"):
        self.c, self.k, self.l, self.enc = np.asarray(c, dtype=np.int64), int(k), int(l), enc

    @property
    def L(self):
        return self.c.shape[2]


def _mods(moduli):
    return np.array(moduli, dtype=np.int64)[:, None]


def lift_plain(x, moduli, pmod):
    """[..., n] integers -> [..., L, n]: centred representative mod pmod (in [-pmod/2, pmod/2)), reduced mod q_l"""
    x = np.mod(np.asarray(x, dtype=np.int64), pmod)
    x = np.where(2 * x >= pmod, x - pmod, x)
    return np.stack([np.mod(x, q) for q in moduli], axis=-2)


class SymmSHE:
    rescale_dec_mask = 1      # S8: modSwitch rounds c0 in the decoding basis, every further coefficient in the powerful basis

    def __init__(self, m, mp, moduli, p):
        self.m, self.mp, self.moduli, self.p = m, mp, [int(q) for q in moduli], int(p)
        self.n, self.np_ = O.totient(O.factor(m)), O.totient(O.factor(mp))

    def _q(self, L):
        return self.moduli[:L]

    def _scale(self, c, scalar_of_q):
        L = c.shape[-2]
        f = np.array([scalar_of_q(q) % q for q in self._q(L)], dtype=np.int64)[:, None]
        return c * f % _mods(self._q(L))

    def gen_sk(self, s_dec):
        q = self.moduli
        x = np.stack([np.mod(np.asarray(s_dec, dtype=np.int64), ql) for ql in q])
        return S.crt(self.mp, q, S.per_limb(O.l, self.mp, q, x))

    def ks_hint(self, sk, mask, noise_dec, L, base=0):
        q = self._q(L)
        qv = _mods(q)
        s = sk[:L]
        gad = S.gadget_vector(q, base)
        e = S.crt(self.mp, q, S.per_limb(O.l, self.mp, q, lift_plain(noise_dec, q, 1 << 40)) * self.p % qv)
        h0 = (e - mask * s % qv) % qv
        s2 = s * s % qv
        for d, (l, v) in enumerate(gad):
            h0[d, l] = (h0[d, l] + s2[l] * v) % q[l]
        return np.stack([h0, mask], axis=1)

    def encrypt(self, sk, mu, noise_dec, c1):
        L = c1.shape[1]
        q = self._q(L)
        qv = _mods(q)
        x = S.embed_pow(self.m, self.mp, lift_plain(mu, q, self.p))
        e = (S.per_limb(O.l, self.mp, q, lift_plain(noise_dec, q, 1 << 40)) * self.p + x) % qv
        c0 = (S.crt(self.mp, q, e) - c1 * sk[:L] % qv) % qv
        return CT(np.stack([c0, c1], axis=1), 0, 1, "LSD")

    def error_term(self, sk, ct):
        """python-int array [B, n']: centred lift of c(s) in the decoding basis, and Q"""
        ct = self.to_lsd(ct)
        L = ct.L
        q = self._q(L)
        qv = _mods(q)
        s = sk[:L]
        ph = np.zeros_like(ct.c[:, 0])
        for i in range(ct.c.shape[1] - 1, -1, -1):
            ph = (ph * s % qv + ct.c[:, i]) % qv
        dec = S.per_limb(O.l_inv, self.mp, q, S.crt_inv(self.mp, q, ph))
        Q = prod(q)
        acc = np.zeros((dec.shape[0], dec.shape[2]), dtype=object)
        for l, ql in enumerate(q):
            Ql = Q // ql
            acc = (acc + dec[:, l].astype(object) * (Ql * pow(Ql, -1, ql))) % Q
        return np.where(acc * 2 > Q, acc - Q, acc), Q

    def error_rate(self, sk, ct):
        e, Q = self.error_term(sk, ct)
        return np.array([float(max(abs(int(v)) for v in row)) / float(Q) for row in e])

    def decrypt(self, sk, ct):
        ct = self.to_lsd(ct)
        e, _ = self.error_term(sk, ct)
        x = np.mod(e, self.p).astype(np.int64)
        x = self._mod_p_ring(x, ct.k)
        return x * ct.l % self.p

    def _mod_p_ring(self, e_dec, k):
        """divG^k in the decoding basis of O_m' over Z_p, then twace to O_m, then to the powerful basis"""
        p = self.p
        x = e_dec
        for _ in range(k):
            x = div_g_dec_any(self.mp, p, x)
        idx = O.ext_indices_pow(self.m, self.mp)          # twacePowDec: the same index gather in both bases
        x = x[..., idx]
        return l_any(self.m, p, x)

    def to_msd(self, ct):
        if ct.enc == "MSD":
            return ct
        p, Q = self.p, prod(self._q(ct.L))
        return CT(self._scale(ct.c, lambda q: pow(p, -1, q)), ct.k, ct.l * (-Q) % p, "MSD")

    def to_lsd(self, ct):
        if ct.enc == "LSD":
            return ct
        p, Q = self.p, prod(self._q(ct.L))
        return CT(self._scale(ct.c, lambda q: p), ct.k, ct.l * pow(-Q, -1, p) % p, "LSD")

    def _align(self, a, b):
        if a.enc != b.enc:
            a, b = self.to_lsd(a), self.to_lsd(b)
        if a.k < b.k:
            a, b = b, a
        if b.k < a.k:
            q = self._q(b.L)
            g = np.stack([O.g_crt(self.mp, ql) for ql in q])
            c = b.c
            for _ in range(a.k - b.k):
                c = c * g % _mods(q)
            b = CT(c, a.k, b.l, b.enc)
        if a.l != b.l:
            f = b.l * pow(a.l, -1, self.p) % self.p
            f = f - self.p if 2 * f >= self.p else f
            b = CT(self._scale(b.c, lambda q: f), b.k, a.l, b.enc)
        return a, b

    def add(self, a, b):
        a, b = self._align(a, b)
        if a.c.shape[1] < b.c.shape[1]:
            a, b = b, a
        c = a.c.copy()
        P = b.c.shape[1]
        c[:, :P] = (c[:, :P] + b.c) % _mods(self._q(a.L))
        return CT(c, a.k, a.l, a.enc)

    def mul(self, a, b):
        a, b = self.to_lsd(a), self.to_lsd(b)
        return CT(S.ct_mul(self.mp, self._q(a.L), a.c, b.c), a.k + b.k + 1, a.l * b.l % self.p, "LSD")

    def mod_switch(self, ct, to):
        ct = self.to_msd(ct)
        return CT(S.mod_switch(self.mp, self.moduli, ct.c, to, self.rescale_dec_mask), ct.k, ct.l, "MSD")

    def key_switch_quad(self, ct, hint, base=0):
        ct = self.to_msd(ct)
        return CT(S.key_switch_quad(self.mp, self.moduli, ct.c, hint, base), ct.k, ct.l, "MSD")

    def mod_switch_pt(self, ct, p_new):
        ct = self.to_msd(ct)
        return CT(ct.c, ct.k, ct.l % p_new, "MSD")

    def _public_crt(self, b, k, l, L):
        q = self._q(L)
        x = lift_plain(np.mod(np.asarray(b, dtype=np.int64) * pow(l, -1, self.p), self.p), q, self.p)
        x = S.embed_pow(self.m, self.mp, x)
        for _ in range(k):                                  # the ciphertext carries g_{m'}^k: multiply in O_m'
            x = S.per_limb(O.mul_g_pow, self.mp, q, x)
        return S.crt(self.mp, q, x)

    def add_public(self, ct, b):
        ct = self.to_lsd(ct)
        c = ct.c.copy()
        c[:, 0] = (c[:, 0] + self._public_crt(b, ct.k, ct.l, ct.L)[None]) % _mods(self._q(ct.L))
        return CT(c, ct.k, ct.l, "LSD")

    def mul_public(self, ct, a):
        pub = self._public_crt(a, 0, 1, ct.L)
        return CT(ct.c * pub[None, None] % _mods(self._q(ct.L)), ct.k, ct.l, ct.enc)


# ---- L and divG over an arbitrary (possibly composite) modulus, per odd prime axis: definitional -----------------
def _axes(m):
    """[(p, phi(p^e), rts)] over the prime powers of m, first prime slowest (C3/C4 of alchemy_oracle.c)"""
    pps = O.factor(m)
    dims = [(p - 1) * p ** (e - 1) for p, e in pps]
    out, rts = [], 1
    for (p, e), d in reversed(list(zip(pps, dims))):
        out.append((p, d, rts))
        rts *= d
    return list(reversed(out))


def _along_j0(m, x, fn):
    """apply fn to every length-(p-1) vector along the j0 axis of each odd prime power (index j = j1 (p-1) + j0)"""
    x = np.array(x, dtype=np.int64)
    n = x.shape[-1]
    for p, d, rts in _axes(m):
        if p == 2:
            continue
        lts = n // (d * rts)
        v = x.reshape(x.shape[:-1] + (lts, d // (p - 1), p - 1, rts))
        v = fn(p, v)
        x = v.reshape(x.shape[:-1] + (n,))
    return x


def l_any(m, mod, x):
    """Dec -> Pow over Z_mod: running sum along j0"""
    return _along_j0(m, x, lambda p, v: np.mod(np.cumsum(v, axis=-2), mod))


def div_g_dec_any(m, mod, x):
    """the inverse of mulGDec over Z_mod, by Gaussian-free solve: uses the oracle's structured operator on a prime
    modulus when mod is prime, otherwise the closed form below (needs every odd prime of m invertible mod `mod`)"""
    def solve(p, v):
        # mulGDec on one axis is the (p-1) x (p-1) matrix G with G p G^-1 integral: find G from the oracle's
        # definition (decoding basis: g = 1 - zeta_p acts as d_i' = d_i - d_(i-1) with wrap through sum), invert mod `mod`
        G = g_dec_matrix(p)
        Ginv = mat_inv_mod(G, mod)
        return np.mod(np.einsum("ij,...jr->...ir", Ginv, v.astype(object)).astype(object), mod).astype(np.int64)
    return _along_j0(m, x, solve)


_GDEC = {}


def g_dec_matrix(p):
    """matrix of multiplication by (1 - zeta_p) in the decoding basis of O_p, read off the C oracle on a prime modulus"""
    if p not in _GDEC:
        q = 1
        while True:
            q += p
            if q > 1000 and all(q % d for d in range(2, int(q ** 0.5) + 1)):
                break
        G = np.zeros((p - 1, p - 1), dtype=np.int64)
        for j in range(p - 1):
            e = np.zeros((1, p - 1), dtype=np.int64)
            e[0, j] = 1
            col = O.mul_g_dec(p, q, e)[0]
            G[:, j] = np.where(col > q // 2, col - q, col)          # entries are in {-1, 0, 1, 2}
        _GDEC[p] = G
    return _GDEC[p]


def mat_inv_mod(G, mod):
    n = G.shape[0]
    a = [[int(v) % mod for v in row] for row in G]
    inv = [[int(i == j) for j in range(n)] for i in range(n)]
    for col in range(n):
        piv = next(r for r in range(col, n) if np.gcd(a[r][col], mod) == 1)
        a[col], a[piv] = a[piv], a[col]
        inv[col], inv[piv] = inv[piv], inv[col]
        d = pow(a[col][col], -1, mod)
        a[col] = [v * d % mod for v in a[col]]
        inv[col] = [v * d % mod for v in inv[col]]
        for r in range(n):
            if r != col and a[r][col]:
                f = a[r][col]
                a[r] = [(x - f * y) % mod for x, y in zip(a[r], a[col])]
                inv[r] = [(x - f * y) % mod for x, y in zip(inv[r], inv[col])]
    return np.array(inv, dtype=object)


# ---- tunnel hint (KeysHints.hs:113 `getTunnelHint`; Lol `tunnelHint` [RECALL]) --------------------------------------------
def eval_lin(r, e, s_idx, moduli, x_pow, flin):
    """the E-linear map fixed by flin[k] = f(relative decoding basis element k), applied to x (powerful basis of O_r,
    [..., L, n_r]) -> CRT basis of O_s: sum_k embed(coeffsDec(x)_k) * flin_k"""
    qv = _mods(moduli)
    tab = O.coeff_indices(e, r)
    dec = S.per_limb(O.l_inv, r, moduli, x_pow)
    acc = 0
    for k in range(tab.shape[0]):
        part = S.per_limb(O.l, e, moduli, dec[..., tab[k]])
        acc = (acc + S.crt(s_idx, moduli, S.embed_pow(e, s_idx, part)) * flin[k] % qv) % qv
    return acc


def tunnel_hint(r, e, s_idx, moduli, s_in_crt, s_out_crt, flin, mask, noise_dec, base):
    """hint[k, d] = (b, a) with b + a s_out = gamma_d * f(relpow_k * s_in) + e_{k,d}: key-switches the K = [O_r : O_e]
    E-coefficients of c1 s_in into O_s.  s_in_crt [L, n_r], s_out_crt [L, n_s] (CRT bases), mask [K, D, L, n_s] uniform,
    noise_dec [K, D, n_s] small integers in the decoding basis of O_s -> [K, D, 2, L, n_s]"""
    L = len(moduli)
    qv = _mods(moduli)
    tab = O.coeff_indices(e, r)
    K, nr, ns = tab.shape[0], s_in_crt.shape[-1], s_out_crt.shape[-1]
    gad = S.gadget_vector(moduli, base)
    hint = np.zeros((K, len(gad), 2, L, ns), dtype=np.int64)
    for k in range(K):
        relpow = np.zeros((L, nr), dtype=np.int64)
        relpow[:, tab[k][0]] = 1
        prod_pow = S.crt_inv(r, moduli, S.crt(r, moduli, relpow) * s_in_crt % qv)
        target = eval_lin(r, e, s_idx, moduli, prod_pow, flin)
        for d, (l, v) in enumerate(gad):
            err = S.crt(s_idx, moduli, S.per_limb(O.l, s_idx, moduli, lift_plain(noise_dec[k, d], moduli, 1 << 40)))
            b = (err - mask[k, d] * s_out_crt % qv) % qv
            b[l] = (b[l] + target[l] * v) % moduli[l]
            hint[k, d, 0], hint[k, d, 1] = b, mask[k, d]
    return hint
