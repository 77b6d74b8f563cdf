#!/usr/bin/env python
"""Regenerate tests/golden/*.npz.

What these fixtures are -- and are not.  The reference ships no golden vectors for this path (alchemy.cabal has no
test-suite), and lol-cpp cannot be built here, so nothing below was produced by the reference.
They are produced by the DEFINITIONAL side of the CPU oracle (O(n^2) evaluation at the primitive roots, schoolbook
digit / rescale arithmetic on python integers), on inputs from a fixed linear-congruential sequence, and frozen so that
  * the structured oracle, and the CUDA path through the C ABI, are checked against numbers that do not move when
    either of them is refactored (tests/test_golden.py, CPU and GPU legs);
  * a maintainer with lol-cpp at hand can run the same inputs through it and diff: the conventions C1-C9 / S1-S7 that
    bit-equality depends on are exactly what these files encode.
Run from the repository root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from oracle import she_oracle as S  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def lcg(count, q, seed):
    out, x = [], seed
    for _ in range(count):
        x = (x * 6364136223846793005 + 1442695040888963407) % (1 << 64)
        out.append((x >> 33) % q)
    return np.array(out, dtype=np.int64)


def primes_1_mod(m, lo, count):
    out, q = [], lo - lo % m + 1
    while len(out) < count:
        if q > lo and all(q % d for d in range(2, int(q ** 0.5) + 1)):
            out.append(q)
        q += m
    return out


def tensor_fixture(m, moduli):
    n = O.totient(O.factor(m))
    x = np.stack([lcg(2 * n, q, 17 * m + l).reshape(2, n) for l, q in enumerate(moduli)], axis=1)        # [2, L, n]
    x[0, :, 0] = [q - 1 for q in moduli]
    out = {"m": m, "moduli": np.array(moduli, dtype=np.int64), "x": x}
    out["crt"] = np.stack([np.stack([O.crt_def(m, q, x[b, l]) for b in range(x.shape[0])]) for l, q in enumerate(moduli)], axis=1)   # definitional O(n^2), one vector at a time
    for name in ("l", "l_inv", "mul_g_pow", "mul_g_dec", "mul_g_crt"):
        out[name] = S.per_limb(getattr(O, name), m, moduli, x)
    for base in (0, 2, 3):
        out[f"decompose_b{base}"] = S.decompose(x, moduli, moduli, base)
    out["rescale_drop_last"] = S.rescale_drop(x, moduli, len(moduli) - 1)
    return out


def main():
    for m, lo in ((16, 1 << 20), (45, 1 << 24), (63, 1 << 29), (36, 1 << 30)):
        moduli = primes_1_mod(m, lo, 3)
        np.savez_compressed(os.path.join(HERE, f"tensor_m{m}.npz"), **tensor_fixture(m, moduli))
        print("wrote", f"tensor_m{m}.npz", moduli)


if __name__ == "__main__":
    main()
