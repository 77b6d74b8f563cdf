"""Host-side mirror of the SymmSHE ciphertext operations ALCHEMY's evaluator calls
(Crypto/Alchemy/Interpreter/Eval.hs:65-67 `mul_`, :130 `modSwitch_`, :133 `keySwitchQuad_`, :134 `tunnel_`,
:131-132 addPublic/mulPublic; chained as Crypto/Alchemy/Interpreter/PT2CT.hs:172-177 and :224-229 say).

Ciphertext polynomials are CUDA int64 tensors c[batch, degree+1, L, n] in the CRT basis (planar limbs).
The (encoding, k, l) bookkeeping fields of Lol's `CT` stay on the host side of the boundary (they are
scalars per ciphertext type, fixed at compile time in ALCHEMY's typed DSL); this module moves residues only.
"""
import ctypes as C

import torch

from ._lib import check, lib
from .tensor import Plan, _ptr, _stream


def _work(nbytes, device):
    return torch.empty(max(8, int(nbytes)), dtype=torch.uint8, device=device)


def ctMul(plan: Plan, a, b):
    """SymmSHE (*): [B,2,L,n] x [B,2,L,n] -> [B,3,L,n] (includes the mulG of every coefficient)"""
    B, two, L, n = a.shape
    assert two == 2 and b.shape == a.shape and n == plan.n
    out = torch.empty((B, 3, L, n), dtype=torch.int64, device=a.device)
    check(lib().ab_ct_mul_dev(plan.h, _ptr(out), _ptr(a), _ptr(b), B, L, _stream()))
    return out


def modSwitch(plan: Plan, c, to_limbs, dec_mask=0):
    """rescaleCyc on every coefficient of an MSD ciphertext: [B,P,from,n] -> [B,P,to,n] over the leading moduli of the plan.
    dec_mask bit i: polynomial i is rounded in the decoding basis (1 = Lol's rescaleLinearMSD: c0 Dec, the rest Pow; 0 = all Pow)"""
    B, P, frm, n = c.shape
    assert n == plan.n
    if to_limbs == frm:
        return c          # Lol's RescaleCyc instance for equal moduli is the identity; callers never mutate ciphertexts
    out = torch.empty((B, P, to_limbs, n), dtype=torch.int64, device=c.device)
    nbytes = lib().ab_mod_switch_work_bytes(plan.h, B, P, frm, to_limbs)
    assert nbytes >= 0
    work = _work(nbytes, c.device)
    check(lib().ab_mod_switch_dev(plan.h, _ptr(out), _ptr(c), B, P, frm, to_limbs, dec_mask, C.c_void_p(work.data_ptr()), _stream()))
    return out


def keySwitchQuadCirc(plan: Plan, c3, hint, base=0):
    """[B,3,L,n] -> [B,2,L,n]; hint[D,2,L,n] with D = gadget length (base 0 = TrivGad, b = BaseBGad b)"""
    B, three, L, n = c3.shape
    D = lib().ab_gadget_len(plan.h, L, base)
    assert three == 3 and n == plan.n and tuple(hint.shape) == (D, 2, L, n), (tuple(c3.shape), tuple(hint.shape), D)
    out = torch.empty((B, 2, L, n), dtype=torch.int64, device=c3.device)
    work = _work(lib().ab_key_switch_work_bytes(plan.h, B, L, base), c3.device)
    check(lib().ab_key_switch_quad_dev(plan.h, _ptr(out), _ptr(c3), _ptr(hint), B, L, base, C.c_void_p(work.data_ptr()), _stream()))
    return out


def mulKeySwitch(plan: Plan, a, b, hint, l_hint, l_out, base=0, prescale=None, dec_mask=0):
    """PT2CT `mul_` (PT2CT.hs:172-177): modSwitch . keySwitchQuad hint . modSwitch . (*)
    a, b at L_in limbs (LSD forms); the hint lives at l_hint limbs; the result (MSD form) at l_out limbs.
    prescale: per-limb scalars of Lol's toMSD (p^-1 mod q_l for the L_in moduli) folded into the product; None = 1
    (symm.SymmSHE.mulKeySwitch passes them and keeps the (k, l) bookkeeping).  dec_mask: as in modSwitch."""
    B, two, lin, n = a.shape
    assert two == 2 and b.shape == a.shape and n == plan.n and hint.shape[2] == l_hint
    out = torch.empty((B, 2, l_out, n), dtype=torch.int64, device=a.device)
    work = _work(lib().ab_mul_key_switch_work_bytes(plan.h, B, lin, l_hint, l_out, base), a.device)
    ps = None
    if prescale is not None:
        assert len(prescale) == lin
        ps = (C.c_int64 * lin)(*[int(v) for v in prescale])
    check(lib().ab_mul_key_switch_dev(plan.h, _ptr(out), _ptr(a), _ptr(b), _ptr(hint), B, lin, l_hint, l_out, base, ps, dec_mask,
                                      C.c_void_p(work.data_ptr()), _stream()))
    return out


class Tunnel:
    """ring switch O_r' -> O_s' through E' = O_e' (ab_tunnel): PT2CT `linearCyc_` core (PT2CT.hs:224-229)"""

    def __init__(self, plan_r: Plan, plan_e: Plan, plan_s: Plan):
        self.r, self.e, self.s = plan_r, plan_e, plan_s
        h = C.c_void_p()
        check(lib().ab_tunnel_create(C.byref(h), plan_r.h, plan_e.h, plan_s.h))
        self.h = h
        self.K = plan_r.n // plan_e.n

    def __call__(self, c2, flin, hint, base=0):
        B, two, L, nr = c2.shape
        D = lib().ab_gadget_len(self.s.h, L, base)
        ns = self.s.n
        assert two == 2 and nr == self.r.n
        assert tuple(flin.shape) == (self.K, L, ns) and tuple(hint.shape) == (self.K, D, 2, L, ns)
        out = torch.empty((B, 2, L, ns), dtype=torch.int64, device=c2.device)
        work = _work(lib().ab_tunnel_work_bytes(self.h, B, L, base), c2.device)
        check(lib().ab_tunnel_dev(self.h, _ptr(out), _ptr(c2), _ptr(flin), _ptr(hint), B, L, base, C.c_void_p(work.data_ptr()), _stream()))
        return out

    def close(self):
        if getattr(self, "h", None):
            lib().ab_tunnel_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
