#!/usr/bin/env python
"""Secondary metrics on synthetic ciphertexts (device-resident, CUDA events): the PT2CT `mul_` pipeline
(PT2CT.hs:172-177) at every level of HomomRLWR and of Arithmetic.hs, and the ring-tunnelling chain of
examples/Tunnel.hs (intended `switch3`): 3 x (modSwitch, tunnel, modSwitch) over H0' -> H1' -> H2' -> H3'.
Prints one JSON object per case.  Under torchrun every rank runs its shard of the ciphertexts (weak scaling is the
driver's business; here B is the TOTAL and each rank takes shard_range)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import alchemy_b200 as ab  # noqa: E402
from alchemy_b200 import she, shard  # noqa: E402

RLWR_Q = [1543651201, 689270401, 718099201, 720720001, 1556755201, 1567238401]
ARITH_Q = [268440577, 8392193, 1073750017]
TUNNEL_Q = [537264001, 539884801, 555609601, 560851201, 566092801]
PEAK = 6461.5


def rnd(moduli, lead, n):
    return torch.stack([torch.randint(0, q, tuple(lead) + (n,), dtype=torch.int64, device="cuda") for q in moduli], dim=-2)


def timeit(fn, iters):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def mul_cases(iters):
    out = []
    for name, m, moduli, lin, lh, lout, B in (
            ("HomomRLWR mul level 1", 20475, RLWR_Q, 4, 5, 3, 2048), ("HomomRLWR mul level 2", 20475, RLWR_Q, 3, 4, 2, 2048),
            ("HomomRLWR mul level 3", 20475, RLWR_Q, 2, 3, 1, 2048), ("HomomRLWR mul level 4", 20475, RLWR_Q, 1, 2, 1, 2048),
            ("Arithmetic mul", 512, ARITH_Q, 2, 2, 1, 32768)):
        plan = ab.Plan(m, moduli)
        n = plan.n
        a, b, hint = rnd(moduli[:lin], (B, 2), n), rnd(moduli[:lin], (B, 2), n), rnd(moduli[:lh], (lh, 2), n)
        l0 = ab.lib().ab_kernel_launches()
        kw = {"prescale": [pow(7 if m == 512 else 32, -1, q) for q in moduli[:lin]], "dec_mask": 1}     # Lol: toMSD scalars, c0 rounded in Dec
        ms = timeit(lambda: she.mulKeySwitch(plan, a, b, hint, lh, lout, **kw), iters)
        nbytes = (32 * n * lin + 16 * n * lout) * B
        out.append({"case": name, "m": m, "n": n, "L_in": lin, "L_hint": lh, "L_out": lout, "B": B, "ms": ms, "ops_per_s": B / (ms * 1e-3),
                    "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / PEAK,
                    "launches_per_call": int((ab.lib().ab_kernel_launches() - l0) // (iters + 2))})
    return out


def tunnel_chain(total, iters, rank, world):
    """examples/Tunnel.hs: input at 2 limbs; each hop = modSwitch . tunnel hint . modSwitch (PT2CT.hs:224-229);
    BaseBGad 2, one-limb hints: hop 1: 2 -> 1 -> 1, hops 2-3: 1 -> 1 -> 1"""
    lo, hi = shard.shard_range(total, world, rank)
    B = hi - lo
    hops = [(11648, 5824, 29120), (29120, 14560, 43680), (43680, 10920, 54600)]
    q = TUNNEL_Q[:2]
    plans = {}
    for r, e, s in hops:
        for m in (r, s):
            if m not in plans:
                plans[m] = ab.Plan(m, q)
    tn, flin, hint = [], [], []
    for r, e, s in hops:
        pe = ab.Plan(e, q, crt=False)
        t = she.Tunnel(plans[r], pe, plans[s])
        D = ab.GT.gadgetLen(plans[s], 1, 2)
        tn.append(t)
        flin.append(rnd(q[:1], (t.K,), plans[s].n))
        hint.append(rnd(q[:1], (t.K, D, 2), plans[s].n))
    x = rnd(q, (B, 2), plans[11648].n)

    def run():
        c = she.modSwitch(plans[11648], x, 1, 1)
        for i, (r, e, s) in enumerate(hops):
            c = tn[i](c, flin[i], hint[i], base=2)
            c = she.modSwitch(plans[s], c, 1, 1)       # equal moduli: the identity rescale of Lol
        return c
    l0 = ab.lib().ab_kernel_launches()
    ms = timeit(run, iters)
    n = [plans[m].n for m in (11648, 29120, 43680, 54600)]
    nbytes = B * (16 * n[0] * 2 + 16 * n[1] + 16 * n[1] + 16 * n[2] + 16 * n[2] + 16 * n[3])
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    return {"case": "Tunnel.hs chain (switch3), BaseBGad 2", "ciphertexts_total": total, "n_gpus": world, "per_gpu": B, "ms": ms,
            "ciphertexts_per_s": total / (ms * 1e-3), "algorithmic_gbs_per_gpu": nbytes / (ms * 1e-3) / 1e9,
            "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / PEAK, "launches_per_call": int((ab.lib().ab_kernel_launches() - l0) // (iters + 2))}


def rlwr_chain(B, iters, graph=False):
    """examples/HomomRLWR.hs as a chain of device calls on synthetic ciphertexts and hints: mulPublic,
    five ring-tunnelling hops H0' -> ... -> H5' (TrivGad; modSwitch to the hint modulus, tunnel, modSwitch back: 5 -> 6 -> 5 limbs,
    last hop 5 -> 5 -> 4), then the rescaling tree at H5': 8 `mul_` pipelines (1 + 4 + 2 + 1 over the four levels), 15 div2
    (bookkeeping only in MSD form) and 9 addPublic.  One unit = one evaluation of the whole circuit on one ciphertext."""
    import ctypes as C
    Q = RLWR_Q
    idx = [11648, 29120, 43680, 54600, 27300, 20475]
    es = [5824, 14560, 10920, 27300, 6825]
    plans = {m: ab.Plan(m, Q) for m in idx}
    hops = []
    for i in range(5):
        r, s, e = idx[i], idx[i + 1], es[i]
        lh, lout = (6, 5) if i < 4 else (5, 4)
        t = she.Tunnel(plans[r], ab.Plan(e, Q, crt=False), plans[s])
        D = ab.GT.gadgetLen(plans[s], lh, 0)
        hops.append((t, s, lh, lout, rnd(Q[:lh], (t.K,), plans[s].n), rnd(Q[:lh], (t.K, D, 2), plans[s].n)))
    p5 = plans[20475]
    n5 = p5.n
    levels = [(4, 5, 3, 1), (3, 4, 2, 4), (2, 3, 1, 2), (1, 2, 1, 1)]          # (L_in, L_hint, L_out, how many)
    hints = [rnd(Q[:lh], (lh, 2), n5) for _, lh, _, _ in levels]
    x0 = rnd(Q[:5], (B, 2), plans[11648].n)
    pub = rnd(Q[:5], (1,), plans[11648].n)[0].contiguous()
    pub5 = [rnd(Q[:lin], (1,), n5)[0].contiguous() for lin, _, _, _ in levels]
    lib = ab.lib()
    st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr())

    def run():
        c = x0.clone()
        ab.check(lib.ab_ct_mul_public_dev(plans[11648].h, ptr(c), ptr(pub), B, 2, 5, st()))
        for t, s, lh, lout, flin, hint in hops:
            c = she.modSwitch(t.r, c, lh, 1)
            c = t(c, flin, hint, base=0)
            c = she.modSwitch(plans[s], c, lout, 1)
        # rescaling tree: level k consumes pairs of level k-1 results; synthetic operands of the right shape per level
        cur = c                                                      # [B, 2, 4, n5]
        for (lin, lh, lout, count), hint, pb in zip(levels, hints, pub5):
            a = cur[:, :, :lin].contiguous()
            outs = []
            for _ in range(count):
                b = a.clone()
                ab.check(lib.ab_ct_add_public_dev(p5.h, ptr(b), ptr(pb), B, 2, lin, st()))      # the (z_i + y) leaves / node operands
                outs.append(she.mulKeySwitch(p5, a, b, hint, lh, lout, prescale=[pow(32, -1, q) for q in Q[:lin]], dec_mask=1))
            cur = outs[0]
        return cur
    l0 = lib.ab_kernel_launches()
    ms = timeit(run, iters)
    out = {"case": "HomomRLWR.hs evaluation chain (mulPublic, 5 tunnel hops, 8 mul_ pipelines, TrivGad)", "B": B, "ms": ms,
           "evaluations_per_s": B / (ms * 1e-3), "launches_per_call": int((lib.ab_kernel_launches() - l0) // (iters + 2))}
    if graph:      # the same chain captured once into a CUDA graph and replayed (the entries are asynchronous on the caller's stream)
        want = run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            got = run()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(got, want)
        out["graph_ms"] = timeit(g.replay, iters)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--ciphertexts", type=int, default=4096)
    ap.add_argument("--cases", default="mul,rlwr,tunnel")
    ap.add_argument("--rlwr-batch", type=int, default=1024)
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    if "mul" in args.cases and rank == 0:
        for r in mul_cases(args.iters):
            print(json.dumps(r), flush=True)
    if "rlwr" in args.cases and rank == 0:
        print(json.dumps(rlwr_chain(args.rlwr_batch, args.iters)), flush=True)
    if "tunnel" in args.cases:
        r = tunnel_chain(args.ciphertexts, args.iters, rank, world)
        if rank == 0:
            print(json.dumps(r), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
