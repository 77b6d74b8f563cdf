#!/usr/bin/env python
"""Time batched CRT / crtInv over the BASELINE config-2/3 sweep (device-resident, CUDA events, buffers > L2).
Writes one JSON object per case to stdout; used under gpurun to fill profiles/*.json."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import alchemy_b200 as ab  # noqa: E402

CFG2 = [1073479681, 1052835841, 1032192001, 949616641]
TUNNEL_Q = [537264001, 539884801, 555609601, 560851201, 566092801]
RLWR_Q = [1543651201, 689270401, 718099201, 720720001, 1556755201, 1567238401]


def time_case(m, moduli, target_bytes=1 << 30, iters=10):
    plan = ab.Plan(m, moduli)
    n, L = plan.n, len(moduli)
    batch = max(1, target_bytes // (8 * n * L))
    y = torch.empty((batch, L, n), dtype=torch.int64, device="cuda")
    for l, q in enumerate(moduli):
        y[:, l] = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda")
    head = y[:1].clone()
    st = torch.cuda.current_stream()
    sp, yp = C.c_void_p(st.cuda_stream), C.c_void_p(y.data_ptr())
    lib = ab.lib()
    out = {"m": m, "n": n, "L": L, "batch": batch}
    for name, fn in (("crt", lib.ab_crt_dev), ("crt_inv", lib.ab_crt_inv_dev)):
        for _ in range(3):
            ab.check(fn(plan.h, yp, batch, 0, L, ab.PLANAR, sp))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(iters):
            ab.check(fn(plan.h, yp, batch, 0, L, ab.PLANAR, sp))
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[name + "_ms"] = ms
        out[name + "_gbs"] = 16 * n * batch * L / (ms * 1e-3) / 1e9
        out[name + "_Mtransforms_s"] = batch * L / (ms * 1e-3) / 1e6
    # 13 forward + 13 inverse in total: bring the buffer back to the input and check
    for _ in range(13):
        pass
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="pow2,mixed")
    ap.add_argument("--peak", type=float, default=6461.5)
    ap.add_argument("--only", type=int, default=0, help="a single index m with the four 30-bit synthetic moduli")
    ap.add_argument("--moduli", default="", help="with --only: comma-separated moduli instead of the synthetic ones")
    ap.add_argument("--bytes", type=int, default=1 << 30)
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    cases = []
    if "pow2" in args.cases:
        for logn in range(10, 16):
            for L in (2, 4):
                cases.append((2 ** (logn + 1), CFG2[:L]))
        cases.append((512, [268440577, 8392193]))
    if "mixed" in args.cases:
        cases += [(40320, CFG2[:4]), (11648, TUNNEL_Q[:2]), (29120, TUNNEL_Q[:1]), (43680, RLWR_Q[:5]), (54600, RLWR_Q[:5]),
                  (27300, RLWR_Q[:5]), (20475, RLWR_Q[:4])]
    if args.only:
        cases = [(args.only, [int(v) for v in args.moduli.split(",")] if args.moduli else CFG2[:4])]
    for m, q in cases:
        r = time_case(m, q, target_bytes=args.bytes, iters=args.iters)
        r["crt_frac"] = r["crt_gbs"] / args.peak
        r["crt_inv_frac"] = r["crt_inv_gbs"] / args.peak
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
