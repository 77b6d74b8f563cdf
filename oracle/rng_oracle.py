"""CPU restatement of the device randomness (alchemy_b200/csrc/ab_rng.cu): Philox4x32-10 keyed by a 64-bit seed,
counter = (position lo, position hi, stream id, draw index).  TEST INFRASTRUCTURE ONLY.
Anchors: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3" (SC'11), whose known-answer
vectors for Philox4x32-10 are checked in tests/test_rng_oracle.py; the reference draws the corresponding values with
Haskell's MonadRandom (Crypto/Alchemy/Interpreter/KeysHints.hs:96,113), so streams cannot match upstream."""
import math

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """vectorised over numpy uint64 arrays holding 32-bit words"""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & np.uint64(MASK) for c in (c0, c1, c2, c3))
    k0, k1 = int(k0) & MASK, int(k1) & MASK
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c0, np.uint64(M1) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & np.uint64(MASK), p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def _draw(count, seed, stream, draw):
    i = np.arange(count, dtype=np.uint64)
    z = np.zeros(count, dtype=np.uint64)
    return philox4x32_10(i & np.uint64(MASK), i >> np.uint64(32), z + np.uint64(stream), z + np.uint64(draw), seed & MASK, (seed >> 32) & MASK)


def uniform(moduli, batch, n, seed, stream):
    """[batch, L, n] residues: floor(r * q / 2^64) with r = (word0 << 32 | word1)"""
    L = len(moduli)
    r0, r1, _, _ = _draw(batch * L * n, seed, stream, 0)
    q = np.tile(np.repeat(np.array(moduli, dtype=np.uint64), n), batch)
    # (r * q) >> 64 with r = r0 2^32 + r1:  (r0 q + ((r1 q) >> 32)) >> 32, every term below 2^64
    v = (r0 * q + ((r1 * q) >> np.uint64(32))) >> np.uint64(32)
    return v.astype(np.int64).reshape(batch, L, n)


def gaussian_cdt(sigma, tail=12.0):
    """T[i] = floor(2^64 P(|X| <= i)) for the discrete Gaussian over Z with parameter sigma (standard deviation),
    cut where the tail mass drops below the table's resolution; double-precision weights, exact integer table"""
    kmax = max(1, int(math.ceil(tail * sigma)))
    w = [math.exp(-k * k / (2.0 * sigma * sigma)) for k in range(kmax + 1)]
    total = w[0] + 2.0 * sum(w[1:])
    out, acc = [], 0.0
    for k in range(kmax + 1):
        acc += (w[k] if k == 0 else 2.0 * w[k]) / total
        out.append(min((1 << 64) - 1, int(acc * 18446744073709551616.0)))
    while len(out) > 1 and out[-2] == (1 << 64) - 1:
        out.pop()
    out[-1] = (1 << 64) - 1        # the last magnitude absorbs the rounding slack; u = 2^64 - 1 still maps past it at most once
    return out


def gaussian(cdt, batch, n, seed, stream):
    """[batch, n] signed samples: magnitude = #{k : T[k] <= u}, sign from bit 0 of word 2"""
    r0, r1, r2, _ = _draw(batch * n, seed, stream, 1)
    u = (r0 << np.uint64(32)) | r1
    mag = np.zeros(batch * n, dtype=np.int64)
    for t in cdt:
        mag += (np.uint64(t) <= u).astype(np.int64)
    return np.where((r2 & np.uint64(1)) == 1, -mag, mag).reshape(batch, n)
