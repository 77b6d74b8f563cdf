"""ctypes binding of the CPU oracle (oracle/alchemy_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under alchemy_b200/ may import this module.
PARITY UNPINNED against lol-cpp (see the header of alchemy_oracle.c).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libalchemy_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "alchemy_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"])
    return _SO


class PP(C.Structure):
    _fields_ = [("p", C.c_int32), ("e", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        i64, p64, ppp = C.c_int64, C.POINTER(C.c_int64), C.POINTER(PP)
        for name in ("ao_crt", "ao_crt_inv", "ao_l", "ao_l_inv", "ao_mul_g_pow", "ao_div_g_pow",
                     "ao_mul_g_dec", "ao_div_g_dec", "ao_mul_g_crt", "ao_div_g_crt"):
            getattr(_lib, name).argtypes = [C.c_int, ppp, i64, p64]
            getattr(_lib, name).restype = None
        _lib.ao_crt_def.argtypes = [C.c_int, ppp, i64, p64, p64]
        _lib.ao_g_crt.argtypes = [C.c_int, ppp, i64, p64]
        for name in ("ao_pow_exponents", "ao_crt_units"):
            getattr(_lib, name).argtypes = [C.c_int, ppp, p64]
        for name in ("ao_ext_indices_pow", "ao_base_indices_crt", "ao_coeff_indices"):
            getattr(_lib, name).argtypes = [C.c_int, ppp, ppp, p64]
        _lib.ao_factor.argtypes = [i64, ppp]
        _lib.ao_factor.restype = C.c_int
        for name in ("ao_smallest_generator",):
            getattr(_lib, name).argtypes = [i64]
            getattr(_lib, name).restype = i64
        _lib.ao_root_of_unity.argtypes = [i64, i64]
        _lib.ao_root_of_unity.restype = i64
        _lib.ao_powmod.argtypes = [i64, i64, i64]
        _lib.ao_powmod.restype = i64
        _lib.ao_plan_create.argtypes = [C.c_int, ppp, i64]
        _lib.ao_plan_create.restype = C.c_void_p
        _lib.ao_plan_destroy.argtypes = [C.c_void_p]
        for name in ("ao_plan_crt", "ao_plan_crt_inv", "ao_plan_crt_generic"):
            getattr(_lib, name).argtypes = [C.c_void_p, p64]
        _lib.ao_plan_crt_many.argtypes = [C.c_void_p, p64, i64, C.c_int, C.c_int]
    return _lib


def _p64(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def factor(m):
    """prime powers of m, primes ascending: [(p, e), ...]"""
    buf = (PP * 32)()
    k = lib().ao_factor(int(m), buf)
    return [(buf[i].p, buf[i].e) for i in range(k)]


def _pps(pps):
    arr = (PP * max(1, len(pps)))()
    for i, (p, e) in enumerate(pps):
        arr[i].p, arr[i].e = p, e
    return arr


def totient(pps):
    n = 1
    for p, e in pps:
        n *= p ** (e - 1) * (p - 1) if e > 0 else 1
    return n


def value(pps):
    m = 1
    for p, e in pps:
        m *= p ** e
    return m


def smallest_generator(q):
    return int(lib().ao_smallest_generator(int(q)))


def root_of_unity(q, m):
    return int(lib().ao_root_of_unity(int(q), int(m)))


def _inplace(name, m, q, x):
    pps = factor(m)
    y = np.ascontiguousarray(x, dtype=np.int64).copy()
    flat = y.reshape(-1, totient(pps))
    fn = getattr(lib(), name)
    for row in flat:
        fn(len(pps), _pps(pps), int(q), _p64(row))
    return y


def crt(m, q, x):
    """Tensor.crt for one modulus; x[..., phi(m)] residues in [0,q)."""
    return _inplace("ao_crt", m, q, x)


def crt_inv(m, q, x):
    return _inplace("ao_crt_inv", m, q, x)


def l(m, q, x):
    return _inplace("ao_l", m, q, x)


def l_inv(m, q, x):
    return _inplace("ao_l_inv", m, q, x)


def mul_g_pow(m, q, x):
    return _inplace("ao_mul_g_pow", m, q, x)


def div_g_pow(m, q, x):
    return _inplace("ao_div_g_pow", m, q, x)


def mul_g_dec(m, q, x):
    return _inplace("ao_mul_g_dec", m, q, x)


def div_g_dec(m, q, x):
    return _inplace("ao_div_g_dec", m, q, x)


def mul_g_crt(m, q, x):
    return _inplace("ao_mul_g_crt", m, q, x)


def div_g_crt(m, q, x):
    return _inplace("ao_div_g_crt", m, q, x)


def crt_def(m, q, x):
    """definitional O(n^2) CRT of a single vector"""
    pps = factor(m)
    x = np.ascontiguousarray(x, dtype=np.int64)
    y = np.empty_like(x)
    lib().ao_crt_def(len(pps), _pps(pps), int(q), _p64(x), _p64(y))
    return y


def g_crt(m, q):
    pps = factor(m)
    out = np.empty(totient(pps), dtype=np.int64)
    lib().ao_g_crt(len(pps), _pps(pps), int(q), _p64(out))
    return out


def pow_exponents(m):
    pps = factor(m)
    out = np.empty(totient(pps), dtype=np.int64)
    lib().ao_pow_exponents(len(pps), _pps(pps), _p64(out))
    return out


def crt_units(m):
    pps = factor(m)
    out = np.empty(totient(pps), dtype=np.int64)
    lib().ao_crt_units(len(pps), _pps(pps), _p64(out))
    return out


def _hi_lo(m, mp):
    assert mp % m == 0
    hi = factor(mp)
    lo_map = dict(factor(m))
    lo = [(p, lo_map.get(p, 0)) for p, _ in hi]
    return hi, lo


def ext_indices_pow(m, mp):
    """for m | mp: index in [phi(mp)] of each pow/dec basis element of O_m"""
    hi, lo = _hi_lo(m, mp)
    out = np.empty(totient(lo), dtype=np.int64)
    lib().ao_ext_indices_pow(len(hi), _pps(hi), _pps(lo), _p64(out))
    return out


def base_indices_crt(m, mp):
    """for m | mp: CRT index in [phi(m)] lying below each CRT index of O_mp"""
    hi, lo = _hi_lo(m, mp)
    out = np.empty(totient(hi), dtype=np.int64)
    lib().ao_base_indices_crt(len(hi), _pps(hi), _pps(lo), _p64(out))
    return out


def coeff_indices(m, mp):
    """[phi(mp)/phi(m), phi(m)] index table of Lol `coeffs`"""
    hi, lo = _hi_lo(m, mp)
    n, npr = totient(lo), totient(hi)
    out = np.empty(npr, dtype=np.int64)
    lib().ao_coeff_indices(len(hi), _pps(hi), _pps(lo), _p64(out))
    return out.reshape(npr // n, n)


class Plan:
    """precomputed root tables for one (m, q); used for batches and CPU timing"""

    def __init__(self, m, q):
        self.pps = factor(m)
        self.n = totient(self.pps)
        self.q = int(q)
        self.h = lib().ao_plan_create(len(self.pps), _pps(self.pps), int(q))
        if not self.h:
            raise ValueError(f"no {m}-th root of unity mod {q}")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ao_plan_destroy(self.h)
            self.h = None

    def crt_many(self, y, inverse=False, threads=1):
        """in place on y[count, n]"""
        assert y.dtype == np.int64 and y.flags.c_contiguous and y.shape[-1] == self.n
        lib().ao_plan_crt_many(self.h, _p64(y), y.size // self.n, int(inverse), int(threads))
        return y

    def crt_generic(self, y):
        assert y.dtype == np.int64 and y.flags.c_contiguous and y.size == self.n
        lib().ao_plan_crt_generic(self.h, _p64(y))
        return y
