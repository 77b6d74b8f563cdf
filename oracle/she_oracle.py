"""CPU restatement of the cross-limb (RNS) and ciphertext-level operations of Lol's SymmSHE that
ALCHEMY's evaluator calls (Crypto/Alchemy/Interpreter/Eval.hs:65-67 `(*)`, :130 modSwitch, :133
keySwitchQuadCirc, :134 tunnel; chained as PT2CT.hs:172-177 and :224-229 prescribe).

TEST INFRASTRUCTURE ONLY (see oracle/alchemy_oracle.c).  PARITY UNPINNED: SymmSHE lives in the absent
`lol-apps` package (stack.yaml:55-60); what follows restates the published algorithms (Crockett-Peikert,
"Lambda o lambda", sections 4-6; Lyubashevsky-Peikert-Regev toolkit) definitionally -- every operation goes
through the powerful basis with whole-vector crtInv/crt from the C oracle, NOT through the fused CRT-domain
shortcuts the CUDA path takes -- so agreement is a real cross-check.

Conventions ([RECALL]):
  C5  arrays here are limb-major x[..., L, n]; Lol's interleaved tuples are a transpose of that.
  C7  gadget on a product ring concatenates per-limb gadgets: decompose (a,b) = decompose a ++ decompose b.
      TrivGad: one digit, the centred lift.  BaseBGad b: k = logCeil b q digits by repeated divModCent
      (remainder in [-b/2, b/2)), the last digit being the remaining quotient.
  C8  rescale (a,b) -> a drops limb b:  (a - lift_centred(b)) * q_b^-1 mod q_a; lift: r if 2r < q else r - q.
"""
import numpy as np

from . import oracle as O


def center(x, q):
    """Lol `lift` on ZqBasic: representative in [-q/2, q/2)"""
    x = np.asarray(x, dtype=np.int64)
    return np.where(2 * x < q, x, x - q)


def invmod(a, q):
    return pow(int(a) % q, q - 2, q)


def per_limb(fn, m, moduli, x):
    """x[..., L, n] -> apply fn(m, q_l, x[..., l, :]) on every limb"""
    x = np.asarray(x, dtype=np.int64)
    return np.stack([fn(m, q, np.ascontiguousarray(x[..., l, :])) for l, q in enumerate(moduli)], axis=-2)


def crt(m, moduli, x):
    return per_limb(O.crt, m, moduli, x)


def crt_inv(m, moduli, x):
    return per_limb(O.crt_inv, m, moduli, x)


def gadget_digits(q, base):
    if base == 0:
        return 1
    k, v = 0, 1
    while v < q:
        v *= base
        k += 1
    return k


def gadget_len(moduli, base):
    return sum(gadget_digits(q, base) for q in moduli)


def gadget_vector(moduli, base):
    """[(limb, value)]: gadget element d is `value` in limb `limb` and 0 elsewhere"""
    out = []
    for l, q in enumerate(moduli):
        for d in range(gadget_digits(q, base)):
            out.append((l, 1 if base == 0 else pow(base, d, q)))
    return out


def decompose_signed(x, moduli, base):
    """x[..., L, n] (powerful basis) -> signed digits [..., D, n] (integers, before any reduction)"""
    x = np.asarray(x, dtype=np.int64)
    digs = []
    for l, q in enumerate(moduli):
        v = center(x[..., l, :], q)
        k = gadget_digits(q, base)
        for d in range(k):
            if base == 0 or d == k - 1:
                digs.append(v)
            else:
                qq, r = np.floor_divide(v, base), np.mod(v, base)      # floor division, r in [0, b)
                hi = 2 * r >= base
                r, qq = np.where(hi, r - base, r), np.where(hi, qq + 1, qq)
                digs.append(r)
                v = qq
    return np.stack(digs, axis=-2)


def decompose(x, moduli, out_moduli, base):
    """digits reduced into each of out_moduli: [..., D, L_out, n]"""
    d = decompose_signed(x, moduli, base)
    return np.stack([np.mod(d, q) for q in out_moduli], axis=-2)


def rescale_drop(x, moduli, drop):
    """x[..., L, n] in the Pow/Dec basis -> [..., L-1, n]"""
    x = np.asarray(x, dtype=np.int64)
    c = center(x[..., drop, :], moduli[drop])
    out = []
    for l, q in enumerate(moduli):
        if l == drop:
            continue
        out.append(np.mod(np.mod(x[..., l, :] - c, q) * invmod(moduli[drop], q), q))
    return np.stack(out, axis=-2)


def ct_mul(m, moduli, a, b):
    """a, b: [B, 2, L, n] CRT basis -> [B, 3, L, n]: the product polynomial in the secret, times g"""
    qv = np.array(moduli, dtype=np.int64)[None, :, None]
    g = np.stack([O.g_crt(m, q) for q in moduli])[None]
    a0, a1, b0, b1 = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    c0 = a0 * b0 % qv
    c1 = (a0 * b1 % qv + a1 * b0 % qv) % qv
    c2 = a1 * b1 % qv
    return np.stack([c0 * g % qv, c1 * g % qv, c2 * g % qv], axis=1)


def mod_switch(m, moduli, c, to, dec_mask=0):
    """c[B, P, from, n] in the CRT basis -> [B, P, to, n] in the CRT basis: rescaleCyc on every coefficient, polynomial i rounded in
    the decoding basis when bit i of dec_mask is set (Lol rescaleLinearMSD [RECALL]: c0 rescaleDec, c1 rescalePow -> dec_mask 1),
    else in the powerful basis."""
    frm = c.shape[-2]
    if to == frm:
        return c.copy()
    if to > frm:
        out = np.zeros(c.shape[:-2] + (to, c.shape[-1]), dtype=np.int64)
        for k in range(frm):
            s = 1
            for l in range(frm, to):
                s = s * (moduli[l] % moduli[k]) % moduli[k]
            out[..., k, :] = c[..., k, :] * s % moduli[k]
        return out
    x = crt_inv(m, moduli[:frm], c)
    dec = [i for i in range(c.shape[1]) if (dec_mask >> i) & 1]
    for i in dec:
        x[:, i] = per_limb(O.l_inv, m, moduli[:frm], x[:, i])
    for L in range(frm, to, -1):
        x = rescale_drop(x, moduli[:L], L - 1)
    for i in dec:
        x[:, i] = per_limb(O.l, m, moduli[:to], x[:, i])
    return crt(m, moduli[:to], x)


def key_switch_quad(m, moduli, c3, hint, base):
    """c3[B, 3, L, n], hint[D, 2, L, n], all CRT basis -> [B, 2, L, n]"""
    L = c3.shape[-2]
    q = moduli[:L]
    qv = np.array(q, dtype=np.int64)[None, :, None]
    c2 = crt_inv(m, q, c3[:, 2])
    dig = crt(m, q, decompose(c2, q, q, base))          # [B, D, L, n]
    out = [c3[:, 0].copy(), c3[:, 1].copy()]
    for d in range(dig.shape[1]):
        for i in range(2):
            out[i] = (out[i] + dig[:, d] * hint[d, i][None] % qv) % qv
    return np.stack(out, axis=1)


def embed_pow(m, mp, x):
    idx = O.ext_indices_pow(m, mp)
    out = np.zeros(x.shape[:-1] + (O.totient(O.factor(mp)),), dtype=np.int64)
    out[..., idx] = x
    return out


def tunnel(r, e, s, moduli, c2, flin, hint, base):
    """ring switch O_r -> O_s through O_e (e | r, e | s).  c2[B, 2, L, n_r] CRT; flin[K, L, n_s] CRT of O_s;
    hint[K, D, 2, L, n_s] CRT of O_s -> [B, 2, L, n_s] CRT of O_s."""
    L = c2.shape[-2]
    q = moduli[:L]
    qv = np.array(q, dtype=np.int64)[None, :, None]
    tab = O.coeff_indices(e, r)                           # [K, n_e]
    K = tab.shape[0]
    pw = crt_inv(r, q, c2)                                # powerful basis of O_r
    c0_dec = per_limb(O.l_inv, r, q, pw[:, 0])            # decoding basis
    ns = O.totient(O.factor(s))
    out0 = np.zeros((c2.shape[0], L, ns), dtype=np.int64)
    out1 = np.zeros_like(out0)
    for k in range(K):
        # E-coefficient k of c0 w.r.t. the relative decoding basis, as an element of O_e (decoding basis),
        # embedded into O_s: Dec -> Pow in O_e, embedPow, CRT_s
        part = per_limb(O.l, e, q, c0_dec[..., tab[k]])
        out0 = (out0 + crt(s, q, embed_pow(e, s, part)) * flin[k][None] % qv) % qv
        # E-coefficient k of c1 w.r.t. the relative powerful basis, embedded, decomposed, key-switched
        dig = crt(s, q, decompose(embed_pow(e, s, pw[:, 1][..., tab[k]]), q, q, base))
        for d in range(dig.shape[1]):
            out0 = (out0 + dig[:, d] * hint[k, d, 0][None] % qv) % qv
            out1 = (out1 + dig[:, d] * hint[k, d, 1][None] % qv) % qv
    return np.stack([out0, out1], axis=1)
