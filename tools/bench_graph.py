#!/usr/bin/env python
"""Latency of the examples/Tunnel.hs chain (switch3, BaseBGad 2) on SMALL batches: eager launches against one CUDA-graph replay.
The *_dev entries are asynchronous on the caller's stream with caller-owned scratch, so the whole chain (20 launches) captures
into one graph.  Prints one JSON line per batch size."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402

import alchemy_b200 as ab  # noqa: E402
from alchemy_b200 import she  # noqa: E402
from bench_she import TUNNEL_Q, rnd  # noqa: E402


def main():
    hops = [(11648, 5824, 29120), (29120, 14560, 43680), (43680, 10920, 54600)]
    q = TUNNEL_Q[:2]
    plans = {m: ab.Plan(m, q) for m in {h[0] for h in hops} | {h[2] for h in hops}}
    tn, flin, hint = [], [], []
    for r, e, s in hops:
        t = she.Tunnel(plans[r], ab.Plan(e, q, crt=False), plans[s])
        D = ab.GT.gadgetLen(plans[s], 1, 2)
        tn.append(t)
        flin.append(rnd(q[:1], (t.K,), plans[s].n))
        hint.append(rnd(q[:1], (t.K, D, 2), plans[s].n))
    for B in (1, 8, 64, 512):
        x = rnd(q, (B, 2), plans[11648].n)

        def run():
            c = she.modSwitch(plans[11648], x, 1, 1)
            for i, (r, e, s) in enumerate(hops):
                c = tn[i](c, flin[i], hint[i], base=2)
            return c
        l0 = ab.lib().ab_kernel_launches()
        want = run()
        launches = ab.lib().ab_kernel_launches() - l0
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = run()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, want)

        def timeit(fn, iters=20):
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        eager, graph = timeit(run), timeit(g.replay)
        print(json.dumps({"case": "Tunnel.hs chain, small batch", "B": B, "launches": int(launches), "eager_ms": eager, "graph_ms": graph,
                          "ciphertexts_per_s_graph": B / (graph * 1e-3)}), flush=True)


def rlwr_small():
    """one HomomRLWR.hs evaluation at a time (the reference's own timed call, examples/HomomRLWR.hs:68) and small batches"""
    from bench_she import rlwr_chain
    for B in (1, 8, 64):
        r = rlwr_chain(B, 5, graph=True)
        print(json.dumps({"case": "HomomRLWR.hs evaluation chain, small batch", "B": B, "eager_ms": r["ms"], "graph_ms": r["graph_ms"],
                          "evaluations_per_s_graph": B / (r["graph_ms"] * 1e-3), "launches": r["launches_per_call"]}), flush=True)


if __name__ == "__main__":
    main()
    rlwr_small()
