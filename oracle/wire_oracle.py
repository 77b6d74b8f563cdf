"""Independent reader / writer of the wire format (include/alchemy_b200.h `ab_wire_header`), numpy + struct only.
TEST INFRASTRUCTURE ONLY: checks that what the CUDA path writes is the documented byte layout, not merely something it can read back."""
import struct

import numpy as np

FMT = "<IHHqqqiiiiqq8qQ"          # 136 bytes, little endian
MAGIC = 0x57324241
VERSION = 2
K = 0x9E3779B97F4A7C15
M64 = (1 << 64) - 1


def checksum(head128, payload):
    """sum_j (j+1) K w_j + sum_i (i+1) v_i mod 2^64 over the 16 header words before the checksum field and the uint32 payload"""
    w = struct.unpack("<16Q", head128)
    s = sum((j + 1) * K * v for j, v in enumerate(w)) & M64
    v = np.asarray(payload, dtype=np.uint64).reshape(-1)
    with np.errstate(over="ignore"):
        s = (s + int((v * np.arange(1, v.size + 1, dtype=np.uint64)).sum(dtype=np.uint64))) & M64
    return s


def parse(raw):
    raw = bytes(raw)
    f = struct.unpack(FMT, raw[:struct.calcsize(FMT)])
    h = dict(magic=f[0], version=f[1], kind=f[2], m=f[3], n=f[4], batch=f[5], npoly=f[6], nlimbs=f[7], enc=f[8], k=f[9], l=f[10], pmod=f[11],
             moduli=list(f[12:20]), checksum=f[20])
    count = h["batch"] * h["npoly"] * h["nlimbs"] * h["n"]
    payload = np.frombuffer(raw[struct.calcsize(FMT):], dtype="<u4")
    assert h["magic"] == MAGIC and h["version"] == VERSION and payload.size == count
    assert checksum(raw[:128], payload) == h["checksum"]
    return h, payload.astype(np.int64).reshape(h["batch"], h["npoly"], h["nlimbs"], h["n"])


def build(kind, m, x, moduli, enc=0, k=0, l=1, pmod=0):
    x = np.asarray(x, dtype=np.int64)
    batch, npoly, L, n = x.shape
    mods = list(moduli[:L]) + [0] * (8 - L)
    payload = x.astype("<u4")
    head = struct.pack(FMT, MAGIC, VERSION, kind, m, n, batch, npoly, L, enc, k, l, pmod, *mods, 0)
    return head[:128] + struct.pack("<Q", checksum(head[:128], payload)) + payload.tobytes()
