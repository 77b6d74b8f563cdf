"""ctypes binding of libalchemy_b200.so (the C ABI declared in include/alchemy_b200.h).

There is no CPU fallback: if the CUDA library is missing this module raises on import of the
symbols, and every compute entry returns AB_ERR_CUDA without a device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALCHEMY_B200_LIB") or os.path.join(_HERE, "lib", "libalchemy_b200.so")   # env override: kernel-variant experiments

AB_OK, AB_ERR_ARG, AB_ERR_NO_ROOT, AB_ERR_CUDA, AB_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
PLANAR, INTERLEAVED = 0, 1
RESCALE_POW, RESCALE_LOL_LINEAR = 0, 1      # dec_mask of modSwitch: all polynomials through Pow / c0 through Dec (Lol rescaleLinearMSD)
OPS = {"crt": 0, "crt_inv": 1, "l": 2, "l_inv": 3, "mul_g_pow": 4, "mul_g_dec": 5, "div_g_pow": 6,
       "div_g_dec": 7, "mul_g_crt": 8, "div_g_crt": 9}


class PrimeExp(C.Structure):
    _fields_ = [("prime", C.c_int32), ("exponent", C.c_int32)]


class WireHeader(C.Structure):
    _fields_ = [("magic", C.c_uint32), ("version", C.c_uint16), ("kind", C.c_uint16), ("m", C.c_int64), ("n", C.c_int64), ("batch", C.c_int64),
                ("npoly", C.c_int32), ("nlimbs", C.c_int32), ("enc", C.c_int32), ("k", C.c_int32), ("l", C.c_int64), ("pmod", C.c_int64),
                ("moduli", C.c_int64 * 8), ("checksum", C.c_uint64)]


class AlchemyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"alchemy_b200 error {code}: {msg}")
        self.code = code


_i64, _i32, _vp = C.c_int64, C.c_int32, C.c_void_p
_tensor_sig = [_vp, _vp, _i64, _i32, _i32, C.c_int, _vp]
_zip_sig = [_vp, _vp, _vp, _i64, _i64, _i32, _i32, C.c_int, _vp]
_ext_sig = [_vp, _vp, _vp, _i64, _i32, _vp]

# every symbol include/alchemy_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "ab_last_error": (C.c_char_p, []),
    "ab_device_count": (C.c_int, []),
    "ab_version": (C.c_char_p, []),
    "ab_plan_create": (C.c_int, [C.POINTER(_vp), C.POINTER(PrimeExp), _i32, C.POINTER(_i64), _i32, C.POINTER(_i64)]),
    "ab_plan_create_nocrt": (C.c_int, [C.POINTER(_vp), C.POINTER(PrimeExp), _i32, C.POINTER(_i64), _i32]),
    "ab_plan_destroy": (None, [_vp]),
    "ab_plan_m": (_i64, [_vp]),
    "ab_plan_n": (_i64, [_vp]),
    "ab_plan_limbs": (_i32, [_vp]),
    "ab_plan_root": (_i64, [_vp, _i32]),
    "ab_crt_dev": (C.c_int, _tensor_sig),
    "ab_crt_inv_dev": (C.c_int, _tensor_sig),
    "ab_l_dev": (C.c_int, _tensor_sig),
    "ab_l_inv_dev": (C.c_int, _tensor_sig),
    "ab_mul_g_pow_dev": (C.c_int, _tensor_sig),
    "ab_mul_g_dec_dev": (C.c_int, _tensor_sig),
    "ab_div_g_pow_dev": (C.c_int, _tensor_sig),
    "ab_div_g_dec_dev": (C.c_int, _tensor_sig),
    "ab_mul_g_crt_dev": (C.c_int, _tensor_sig),
    "ab_div_g_crt_dev": (C.c_int, _tensor_sig),
    "ab_mul_dev": (C.c_int, _zip_sig),
    "ab_add_dev": (C.c_int, _zip_sig),
    "ab_sub_dev": (C.c_int, _zip_sig),
    "ab_neg_dev": (C.c_int, _tensor_sig),
    "ab_scalar_mul_dev": (C.c_int, [_vp, _vp, C.POINTER(_i64), _i64, _i32, _i32, C.c_int, _vp]),
    "ab_ext_create": (C.c_int, [C.POINTER(_vp), _vp, _vp]),
    "ab_ext_destroy": (None, [_vp]),
    "ab_twace_pow_dec_dev": (C.c_int, _ext_sig),
    "ab_embed_pow_dev": (C.c_int, _ext_sig),
    "ab_embed_dec_dev": (C.c_int, _ext_sig),
    "ab_twace_crt_dev": (C.c_int, _ext_sig),
    "ab_embed_crt_dev": (C.c_int, _ext_sig),
    "ab_coeffs_dev": (C.c_int, _ext_sig),
    "ab_uncoeffs_dev": (C.c_int, _ext_sig),
    "ab_ext_rel_dim": (_i64, [_vp]),
    "ab_gadget_len": (_i32, [_vp, _i32, _i32]),
    "ab_decompose_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "ab_rescale_drop_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "ab_ct_mul_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _vp]),
    "ab_mod_switch_work_bytes": (_i64, [_vp, _i64, _i32, _i32, _i32]),
    "ab_mod_switch_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, C.c_uint32, _vp, _vp]),
    "ab_key_switch_work_bytes": (_i64, [_vp, _i64, _i32, _i32]),
    "ab_key_switch_quad_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "ab_mul_key_switch_work_bytes": (_i64, [_vp, _i64, _i32, _i32, _i32, _i32]),
    "ab_mul_key_switch_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, C.POINTER(_i64), C.c_uint32, _vp, _vp]),
    "ab_tunnel_create": (C.c_int, [C.POINTER(_vp), _vp, _vp, _vp]),
    "ab_tunnel_destroy": (None, [_vp]),
    "ab_tunnel_work_bytes": (_i64, [_vp, _i64, _i32, _i32]),
    "ab_tunnel_dev": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "ab_ct_add_public_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "ab_ct_mul_public_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "ab_ct_phase_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _vp]),
    "ab_rns_lift_reduce_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i64, _vp]),
    "ab_lift_plain_dev": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i64, _vp]),
    "ab_uniform_dev": (C.c_int, [_vp, _vp, _i64, _i32, C.c_uint64, C.c_uint32, _vp]),
    "ab_gaussian_dev": (C.c_int, [_vp, _vp, _i64, C.POINTER(C.c_uint64), _i32, C.c_uint64, C.c_uint32, _vp]),
    "ab_wire_bytes": (_i64, [C.POINTER(WireHeader)]),
    "ab_wire_pack_dev": (C.c_int, [_vp, C.POINTER(WireHeader), _vp, _vp]),
    "ab_wire_read_header": (C.c_int, [C.POINTER(WireHeader), _vp, _i64]),
    "ab_wire_unpack_dev": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ab_tensor_op_host": (C.c_int, [_vp, C.c_int, _vp, _i64, C.c_int]),
    "ab_mul_host": (C.c_int, [_vp, _vp, _vp, _i64, C.c_int]),
    "tensorCRTRq": (C.c_int, [C.c_int16, _vp, _i64, C.POINTER(PrimeExp), C.c_int16, C.POINTER(_vp), C.POINTER(_i64)]),
    "tensorCRTInvRq": (C.c_int, [C.c_int16, _vp, _i64, C.POINTER(PrimeExp), C.c_int16, C.POINTER(_vp), C.POINTER(_i64), C.POINTER(_i64)]),
    "tensorLRq": (C.c_int, [C.c_int16, _vp, _i64, C.POINTER(PrimeExp), C.c_int16, C.POINTER(_i64)]),
    "tensorLInvRq": (C.c_int, [C.c_int16, _vp, _i64, C.POINTER(PrimeExp), C.c_int16, C.POINTER(_i64)]),
    "tensorGPowRq": (C.c_int, [C.c_int16, _vp, _i64, C.POINTER(PrimeExp), C.c_int16, C.POINTER(_i64)]),
    "tensorGDecRq": (C.c_int, [C.c_int16, _vp, _i64, C.POINTER(PrimeExp), C.c_int16, C.POINTER(_i64)]),
    "tensorGInvPowRq": (C.c_int, [C.c_int16, _vp, _i64, C.POINTER(PrimeExp), C.c_int16, C.POINTER(_i64)]),
    "tensorGInvDecRq": (C.c_int, [C.c_int16, _vp, _i64, C.POINTER(PrimeExp), C.c_int16, C.POINTER(_i64)]),
    "mulRq": (C.c_int, [C.c_int16, _vp, _vp, _i64, C.POINTER(_i64)]),
    "ab_kernel_launches": (C.c_uint64, []),
}

_lib = None


def lib():
    """load the CUDA library; fails loudly when it has not been built"""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  alchemy_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the header and the library diverge
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != AB_OK:
        raise AlchemyError(rc, lib().ab_last_error().decode())


def factor(m):
    """prime-power factorisation, primes ascending (Lol `ppsFact`)"""
    out, p = [], 2
    while p * p <= m:
        if m % p == 0:
            e = 0
            while m % p == 0:
                m //= p
                e += 1
            out.append((p, e))
        p += 1
    if m > 1:
        out.append((m, 1))
    return out


def prime_exps(m):
    pps = factor(m)
    arr = (PrimeExp * max(1, len(pps)))()
    for i, (p, e) in enumerate(pps):
        arr[i].prime, arr[i].exponent = p, e
    return arr, len(pps)
