"""Wire / on-disk format for ring elements, ciphertexts and hints (ab_wire.cu; include/alchemy_b200.h `ab_wire_header`).
The reference has no serialisation (keys and hints are in-memory Dynamics, Crypto/Alchemy/Interpreter/KeysHints.hs:36-39);
this is what moves ciphertext shards between processes or boxes at half the bytes of the int64 tensors."""
import ctypes as C

import torch

from ._lib import WireHeader, check, lib
from .tensor import _ptr, _stream

KIND = {"elements": 1, "ciphertexts": 2, "ks_hint": 3, "tunnel_hint": 4}


def _header(kind, m, x, moduli, enc="LSD", k=0, l=1, pmod=0):
    """x: [..., npoly, L, n]; leading dimensions are flattened into `batch`"""
    assert x.dtype == torch.int64 and x.is_cuda and x.dim() >= 3
    n, L, npoly = x.shape[-1], x.shape[-2], x.shape[-3]
    batch = x.numel() // (n * L * npoly)
    h = WireHeader()
    h.kind, h.m, h.n, h.batch, h.npoly, h.nlimbs = KIND[kind], int(m), n, batch, npoly, L
    h.enc, h.k, h.l, h.pmod = (1 if enc == "MSD" else 0), int(k), int(l), int(pmod)
    for i, q in enumerate(moduli[:L]):
        h.moduli[i] = int(q)
    return h


def pack(kind, m, x, moduli, **meta):
    """device tensor -> device uint8 tensor holding header + uint32 payload"""
    h = _header(kind, m, x, moduli, **meta)
    nbytes = lib().ab_wire_bytes(C.byref(h))
    out = torch.empty((nbytes + 7) // 8 * 8, dtype=torch.uint8, device=x.device)
    check(lib().ab_wire_pack_dev(C.c_void_p(out.data_ptr()), C.byref(h), _ptr(x.contiguous()), _stream()))
    return out[:nbytes]


def unpack(wire):
    """device uint8 tensor -> (header fields, int64 tensor [batch, npoly, L, n]); raises on truncation / corruption"""
    assert wire.dtype == torch.uint8 and wire.is_cuda
    buf = wire if wire.data_ptr() % 8 == 0 else wire.clone()
    h = WireHeader()
    check(lib().ab_wire_read_header(C.byref(h), C.c_void_p(buf.data_ptr()), buf.numel()))
    out = torch.empty((h.batch, h.npoly, h.nlimbs, h.n), dtype=torch.int64, device=wire.device)
    check(lib().ab_wire_unpack_dev(_ptr(out), C.c_void_p(buf.data_ptr()), buf.numel(), _stream()))
    meta = {"kind": {v: k for k, v in KIND.items()}[h.kind], "m": h.m, "enc": "MSD" if h.enc else "LSD", "k": h.k, "l": h.l, "pmod": h.pmod,
            "moduli": [h.moduli[i] for i in range(h.nlimbs)]}
    return meta, out


def save(path, kind, m, x, moduli, **meta):
    with open(path, "wb") as f:
        f.write(pack(kind, m, x, moduli, **meta).cpu().numpy().tobytes())


def load(path, device="cuda"):
    import numpy as np
    with open(path, "rb") as f:
        raw = np.frombuffer(f.read(), dtype=np.uint8)
    return unpack(torch.from_numpy(raw.copy()).to(device))
