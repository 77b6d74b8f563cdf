"""Host-side mirror of Lol's `class Tensor t` for the B200 backend (tag `GT`, standing where the
reference uses the lol-cpp tag `CT`: examples/Arithmetic.hs:19,23; HomomRLWR.hs:24,47; Tunnel.hs:20,41).

Method names follow Lol's class ([RECALL]): crt, crtInv, l, lInv, mulGPow,
mulGDec, divGPow, divGDec, mulGCRT, divGCRT, twacePowDec, embedPow, embedDec, twaceCRT, embedCRT,
coeffs, zipWithT-style mul/add/sub.  Every method takes a batch of ring elements as a CUDA int64
tensor y[batch, L, n] (planar) or y[batch, n, L] (Lol's interleaved tuples) and works in place,
like the C functions it replaces.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import INTERLEAVED, PLANAR, check, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    assert t.is_cuda and t.dtype == torch.int64 and t.is_contiguous(), "need a contiguous CUDA int64 tensor"
    return C.c_void_p(t.data_ptr())


class Plan:
    """device tables for one cyclotomic index m and a list of moduli (ab_plan)"""

    def __init__(self, m, moduli, roots=None, crt=True):
        self.m, self.moduli = int(m), [int(q) for q in moduli]
        pe, npe = _lib.prime_exps(self.m)
        qs = (C.c_int64 * len(self.moduli))(*self.moduli)
        h = C.c_void_p()
        if crt:
            rs = (C.c_int64 * len(self.moduli))(*roots) if roots is not None else None
            check(lib().ab_plan_create(C.byref(h), pe, npe, qs, len(self.moduli), rs))
        else:
            check(lib().ab_plan_create_nocrt(C.byref(h), pe, npe, qs, len(self.moduli)))
        self.h = h
        self.n = int(lib().ab_plan_n(h))
        self.L = len(self.moduli)
        self.has_crt = crt

    def root(self, limb):
        return int(lib().ab_plan_root(self.h, limb))

    def close(self):
        if getattr(self, "h", None):
            lib().ab_plan_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Ext:
    """index tables for an extension O_m in O_m' (ab_ext): twace/embed/coeffs"""

    def __init__(self, lo: Plan, hi: Plan):
        self.lo, self.hi = lo, hi
        h = C.c_void_p()
        check(lib().ab_ext_create(C.byref(h), lo.h, hi.h))
        self.h = h
        self.rel = int(lib().ab_ext_rel_dim(h))

    def close(self):
        if getattr(self, "h", None):
            lib().ab_ext_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GT:
    """The tensor backend.  Stateless; all methods are static and in place."""

    @staticmethod
    def _shape(plan, y, layout, limb0, nlimbs):
        nl = plan.L - limb0 if nlimbs is None else nlimbs
        want = (nl, plan.n) if layout == PLANAR else (plan.n, nl)
        assert y.dim() == 3 and tuple(y.shape[1:]) == want, f"expected [batch,{want[0]},{want[1]}], got {tuple(y.shape)}"
        return y.shape[0], nl

    @staticmethod
    def _op(name, plan, y, layout=PLANAR, limb0=0, nlimbs=None):
        b, nl = GT._shape(plan, y, layout, limb0, nlimbs)
        check(getattr(lib(), name)(plan.h, _ptr(y), b, limb0, nl, layout, _stream()))
        return y

    # ---- Tensor class methods -------------------------------------------------------------
    @staticmethod
    def crt(plan, y, **kw):
        return GT._op("ab_crt_dev", plan, y, **kw)

    @staticmethod
    def crtInv(plan, y, **kw):
        return GT._op("ab_crt_inv_dev", plan, y, **kw)

    @staticmethod
    def l(plan, y, **kw):  # noqa: E743
        return GT._op("ab_l_dev", plan, y, **kw)

    @staticmethod
    def lInv(plan, y, **kw):
        return GT._op("ab_l_inv_dev", plan, y, **kw)

    @staticmethod
    def mulGPow(plan, y, **kw):
        return GT._op("ab_mul_g_pow_dev", plan, y, **kw)

    @staticmethod
    def mulGDec(plan, y, **kw):
        return GT._op("ab_mul_g_dec_dev", plan, y, **kw)

    @staticmethod
    def divGPow(plan, y, **kw):
        return GT._op("ab_div_g_pow_dev", plan, y, **kw)

    @staticmethod
    def divGDec(plan, y, **kw):
        return GT._op("ab_div_g_dec_dev", plan, y, **kw)

    @staticmethod
    def mulGCRT(plan, y, **kw):
        return GT._op("ab_mul_g_crt_dev", plan, y, **kw)

    @staticmethod
    def divGCRT(plan, y, **kw):
        return GT._op("ab_div_g_crt_dev", plan, y, **kw)

    @staticmethod
    def neg(plan, y, **kw):
        return GT._op("ab_neg_dev", plan, y, **kw)

    @staticmethod
    def _zip(name, plan, y, x, layout=PLANAR, limb0=0, nlimbs=None):
        b, nl = GT._shape(plan, y, layout, limb0, nlimbs)
        xb, _ = GT._shape(plan, x, layout, limb0, nlimbs)
        check(getattr(lib(), name)(plan.h, _ptr(y), _ptr(x), b, xb, limb0, nl, layout, _stream()))
        return y

    @staticmethod
    def mul(plan, y, x, **kw):
        return GT._zip("ab_mul_dev", plan, y, x, **kw)

    @staticmethod
    def add(plan, y, x, **kw):
        return GT._zip("ab_add_dev", plan, y, x, **kw)

    @staticmethod
    def sub(plan, y, x, **kw):
        return GT._zip("ab_sub_dev", plan, y, x, **kw)

    @staticmethod
    def scalarMul(plan, y, scalars, layout=PLANAR, limb0=0, nlimbs=None):
        b, nl = GT._shape(plan, y, layout, limb0, nlimbs)
        sc = (C.c_int64 * nl)(*[int(s) for s in scalars])
        check(lib().ab_scalar_mul_dev(plan.h, _ptr(y), sc, b, limb0, nl, layout, _stream()))
        return y

    # ---- extension rings ------------------------------------------------------------------
    @staticmethod
    def _ext(name, ext, x, n_in, n_out, nlimbs, rel_in=1, rel_out=1):
        nl = ext.lo.L if nlimbs is None else nlimbs
        b = x.shape[0]
        assert tuple(x.shape[1:]) == ((rel_in, nl, n_in) if rel_in > 1 else (nl, n_in)), tuple(x.shape)
        shape = (b, rel_out, nl, n_out) if rel_out > 1 else (b, nl, n_out)
        out = torch.empty(shape, dtype=torch.int64, device=x.device)
        check(getattr(lib(), name)(ext.h, _ptr(out), _ptr(x), b, nl, _stream()))
        return out

    @staticmethod
    def twacePowDec(ext, x, nlimbs=None):
        return GT._ext("ab_twace_pow_dec_dev", ext, x, ext.hi.n, ext.lo.n, nlimbs)

    @staticmethod
    def embedPow(ext, x, nlimbs=None):
        return GT._ext("ab_embed_pow_dev", ext, x, ext.lo.n, ext.hi.n, nlimbs)

    @staticmethod
    def embedDec(ext, x, nlimbs=None):
        return GT._ext("ab_embed_dec_dev", ext, x, ext.lo.n, ext.hi.n, nlimbs)

    @staticmethod
    def twaceCRT(ext, x, nlimbs=None):
        return GT._ext("ab_twace_crt_dev", ext, x, ext.hi.n, ext.lo.n, nlimbs)

    @staticmethod
    def embedCRT(ext, x, nlimbs=None):
        return GT._ext("ab_embed_crt_dev", ext, x, ext.lo.n, ext.hi.n, nlimbs)

    @staticmethod
    def coeffs(ext, x, nlimbs=None):
        """[b, L, n'] -> [b, rel, L, n]: coefficients w.r.t. the relative powerful/decoding basis"""
        out = GT._ext("ab_coeffs_dev", ext, x, ext.hi.n, ext.lo.n, nlimbs, rel_out=ext.rel)
        return out if ext.rel > 1 else out.unsqueeze(1)

    @staticmethod
    def uncoeffs(ext, x, nlimbs=None):
        if ext.rel == 1:
            x = x.squeeze(1)
        return GT._ext("ab_uncoeffs_dev", ext, x, ext.lo.n, ext.hi.n, nlimbs, rel_in=ext.rel)

    # ---- cross-limb -----------------------------------------------------------------------
    @staticmethod
    def gadgetLen(plan, nlimbs, base):
        return int(lib().ab_gadget_len(plan.h, nlimbs, base))

    @staticmethod
    def decompose(plan, x, base=0, out_limbs=None):
        """[b, L, n] (Pow basis) -> [b, digits, L_out, n]: gadget digits reduced into each modulus"""
        b, nl, n = x.shape
        ol = nl if out_limbs is None else out_limbs
        nd = GT.gadgetLen(plan, nl, base)
        out = torch.empty((b, nd, ol, n), dtype=torch.int64, device=x.device)
        check(lib().ab_decompose_dev(plan.h, _ptr(out), _ptr(x), b, nl, ol, base, _stream()))
        return out

    @staticmethod
    def rescaleDrop(plan, x, drop=None):
        """Lol RescaleCyc (a,b) -> b in the Pow/Dec basis: drop one limb"""
        b, nl, n = x.shape
        d = nl - 1 if drop is None else drop
        out = torch.empty((b, nl - 1, n), dtype=torch.int64, device=x.device)
        check(lib().ab_rescale_drop_dev(plan.h, _ptr(out), _ptr(x), b, nl, d, _stream()))
        return out

    # ---- randomness (Lol `tGaussianDec`, MonadRandom draws of KeysHints.hs:96,113) ------------------
    @staticmethod
    def uniform(plan, batch, nlimbs, seed, stream_id, device="cuda"):
        """[batch, L, n] residues uniform in [0, q_l): Philox4x32-10 at (seed, stream_id, position)"""
        out = torch.empty((batch, nlimbs, plan.n), dtype=torch.int64, device=device)
        check(lib().ab_uniform_dev(plan.h, _ptr(out), batch, nlimbs, seed, stream_id, _stream()))
        return out

    @staticmethod
    def gaussianDec(plan, batch, cdt, seed, stream_id, device="cuda"):
        """[batch, n] signed integers from the discrete Gaussian whose folded cumulative table is `cdt`
        (floor(2^64 P(|X| <= k)), k = 0, 1, ...), i.i.d. per decoding-basis coefficient"""
        out = torch.empty((batch, plan.n), dtype=torch.int64, device=device)
        tab = (C.c_uint64 * len(cdt))(*[int(t) for t in cdt])
        check(lib().ab_gaussian_dev(plan.h, _ptr(out), batch, tab, len(cdt), seed, stream_id, _stream()))
        return out

    # ---- host-buffer entry (what a Haskell foreign import would call) ------------------------
    @staticmethod
    def hostOp(plan, op, y_host, layout=INTERLEAVED):
        """y_host: pinned or pageable CPU int64 tensor / numpy array, shape [batch, ...]; in place"""
        import numpy as np
        if isinstance(y_host, np.ndarray):
            assert y_host.dtype == np.int64 and y_host.flags.c_contiguous
            ptr, batch = C.c_void_p(y_host.ctypes.data), y_host.shape[0]
        else:
            assert (not y_host.is_cuda) and y_host.dtype == torch.int64 and y_host.is_contiguous()
            ptr, batch = C.c_void_p(y_host.data_ptr()), y_host.shape[0]
        check(lib().ab_tensor_op_host(plan.h, _lib.OPS[op], ptr, batch, layout))
        return y_host
