#!/usr/bin/env python
"""bench.py -- batched CRT/crtInv throughput of the B200 tensor backend (BASELINE.json metric M1).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--logn 15] [--limbs 4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE config 2): power-of-two index m = 2^(logn+1), n = 2^logn, `limbs` of the
synthetic 30-bit moduli, B ring elements per GPU sized so the in-place buffer is ~1 GiB (8x the 126 MB L2).
One step = one forward CRT launch over the whole batch followed by one crtInv launch: 2*B*L transforms, after
which the buffer holds the input again (checked once after the timed region: crtInv . crt = id, bit for bit).

value  : transforms/s, all ranks, data resident in HBM, CUDA events on the launching stream, max over ranks.
e2e    : transforms/s through the host-buffer C-ABI entry a `foreign import ccall` binds (ab_tensor_op_host):
         pinned host buffers, H2D + kernel + D2H inside the timed region.
roofline: 16*n algorithmic bytes per transform / mean launch time vs the measured HBM copy peak.
cpu_baseline / --impl reference: the CPU oracle (oracle/alchemy_oracle.c, "port": lol-cpp cannot be built here,
         SURVEY.md F3) on all host cores over a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2_MODULI = [1073479681, 1052835841, 1032192001, 949616641]
METRIC, UNIT = "crt_transforms_per_s", "transforms/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--logn", type=int, default=15)
    ap.add_argument("--limbs", type=int, default=4)
    ap.add_argument("--m", type=int, default=0, help="non-power-of-two index instead of 2^(logn+1)")
    ap.add_argument("--batch", type=int, default=0, help="ring elements per GPU (default: ~1 GiB of residues)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-m2", action="store_true", help="skip the secondary mul+key-switch measurement")
    return ap.parse_args()


def workload(args):
    from oracle import oracle as O      # only for phi(m); the bench's checker / CPU leg
    m = args.m if args.m else 2 ** (args.logn + 1)
    n = O.totient(O.factor(m))
    moduli = CFG2_MODULI[:args.limbs]
    batch = args.batch if args.batch else max(1, (1 << 30) // (8 * n * len(moduli)))
    name = f"batched CRT+crtInv, m={m} (n={n}), L={len(moduli)} moduli {moduli}, B={batch} per GPU"
    return m, n, moduli, batch, name


def measured_traffic(m, L, batch):
    """DRAM bytes per launch of the two CRT kernels from the committed ncu capture (profiles/), if it is for this workload"""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
        p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        t = json.load(f)
    if t.get("workload_key") != f"m={m},L={L},B={batch}":
        return None
    return (t["fwd"] + t["inv"]) / 2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], 0.0, set()
        for ts, line in self.rows:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mhz = float(f[0]); smax = max(smax, float(f[1]))
            except ValueError:
                continue
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(mhz)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            sm = [float(r[1].split(",")[0]) for r in self.rows[-3:] if r[1].split(",")[0].strip().replace(".", "").isdigit()] or [0.0]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU leg
def cpu_leg(m, n, moduli, target_s=10.0, steps=1, warmup=0):
    """the CPU oracle (all host cores) on a bounded sample: `count` transforms per modulus, forward + inverse"""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    plans = [O.Plan(m, q) for q in moduli]
    rng = np.random.default_rng(1)
    probe = rng.integers(0, moduli[0], size=(cores, n), dtype=np.int64)
    t = time.perf_counter()
    plans[0].crt_many(probe, threads=cores)
    per_wave = max(1e-4, time.perf_counter() - t)           # `cores` transforms in parallel
    waves = max(1, int(target_s / (per_wave * 2 * len(moduli) * max(1, steps + warmup))))
    count = cores * waves
    data = [rng.integers(0, q, size=(count, n), dtype=np.int64) for q in moduli]
    times = []
    for it in range(warmup + steps):
        t = time.perf_counter()
        for pl, d in zip(plans, data):
            pl.crt_many(d, threads=cores)
            pl.crt_many(d, inverse=True, threads=cores)
        times.append(time.perf_counter() - t)
    times = times[warmup:]
    per_step = 2 * count * len(moduli)
    # the same restatement on ONE core ([RECALL] lol-cpp itself is single-threaded), a handful of transforms
    one = data[0][:2]
    t = time.perf_counter()
    plans[0].crt_many(one, threads=1)
    plans[0].crt_many(one, inverse=True, threads=1)
    one_core = 4 / max(1e-9, time.perf_counter() - t)
    return {"value": per_step * len(times) / sum(times), "unit": UNIT, "cores": cores, "kind": "port", "value_1core": one_core,
            "sample": f"{count} ring elements x {len(moduli)} moduli, CRT+crtInv each, {len(times)} pass(es); "
                      "C restatement (oracle/alchemy_oracle.c), not lol-cpp"}, sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    m, n, moduli, batch, name = workload(args)
    base, ms = cpu_leg(m, n, moduli, target_s=60.0, steps=args.steps, warmup=args.warmup)
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
            "data": "synthetic", "impl": "reference", "config": {"workload": name, "note": "CPU oracle port on host cores, bounded sample"},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ M2
def mul_keyswitch_leg(ab, torch, iters=5):
    """PT2CT `mul_` pipeline modSwitch . keySwitchQuad . modSwitch . (*) (PT2CT.hs:172-177) on synthetic ciphertexts:
    HomomRLWR's deepest level (m' = 20475, n = 8640, L 4 -> hint 5 -> 3, TrivGad) and Arithmetic.hs (m' = 512, 2 -> 2 -> 1)."""
    from alchemy_b200 import she
    out = []
    for name, m, moduli, lin, lh, lout, B, pt in (
            ("HomomRLWR level 1", 20475, [1543651201, 689270401, 718099201, 720720001, 1556755201, 1567238401], 4, 5, 3, 1024, 32),
            ("Arithmetic", 512, [268440577, 8392193, 1073750017], 2, 2, 1, 65536, 7)):
        plan = ab.Plan(m, moduli)
        n = plan.n
        rnd = lambda lead, L: torch.stack([torch.randint(0, q, tuple(lead) + (n,), dtype=torch.int64, device="cuda") for q in moduli[:L]], dim=-2)
        a, b = rnd((B, 2), lin), rnd((B, 2), lin)
        hint = rnd((lh, 2), lh)
        # Lol's sequence: (*) on LSD forms, toMSD scalars p^-1 mod q_l folded into the product, c0 rounded in the decoding basis
        kw = {"prescale": [pow(pt, -1, q) for q in moduli[:lin]], "dec_mask": 1}
        lib = ab.lib()
        for _ in range(2):
            she.mulKeySwitch(plan, a, b, hint, lh, lout, **kw)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.ab_kernel_launches()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            she.mulKeySwitch(plan, a, b, hint, lh, lout, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        nbytes = (32 * n * lin + 16 * n * lout) * B
        out.append({"workload": f"{name}: m'={m}, n={n}, L {lin}->{lh}->{lout}, B={B}", "ops_per_s": B / (ms * 1e-3), "ms": ms,
                    "algorithmic_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peaks()[0],
                    "bytes_per_op": "32 n L_in + 16 n L_out (SURVEY 8d); " + (
                        "the pipeline is 30 n-point transforms per op at level 1: instruction-bound, not HBM-bound" if m != 512 else
                        "8 transforms of 256 points per op in one warp (k_mulks_n256): bound by the integer-multiply pipe, not by HBM"),
                    "launches_per_call": int((lib.ab_kernel_launches() - l0) // iters)})
    return out


# ------------------------------------------------------------------------------------------ config 3
def mixed_leg(ab, torch, peak, iters=10):
    """BASELINE config 3: batched CRT / crtInv on the non-power-of-two index m = 40320 = 2^7 3^2 5 7 (n = 9216, mixed-radix stages),
    same four moduli, ~1 GiB in place (larger than L2), CUDA events per direction on the launching stream."""
    m, moduli = 40320, CFG2_MODULI
    plan = ab.Plan(m, moduli)
    n, L = plan.n, len(moduli)
    batch = (1 << 30) // (8 * n * L)
    y = torch.empty((batch, L, n), dtype=torch.int64, device="cuda")
    for l, q in enumerate(moduli):
        y[:, l] = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda")
    head = y[:2].clone()
    st = torch.cuda.current_stream()
    sp, yp, lib = C.c_void_p(st.cuda_stream), C.c_void_p(y.data_ptr()), ab.lib()
    out = {"workload": f"batched CRT / crtInv, m={m} (n={n}), L={L}, B={batch}", "kernel": "k_crt_mixed<0> / k_crt_mixed<1>", "bytes_per_launch": 16 * n * batch * L}
    for name, fn in (("fwd", lib.ab_crt_dev), ("inv", lib.ab_crt_inv_dev)):
        for _ in range(3):
            ab.check(lib.ab_crt_dev(plan.h, yp, batch, 0, L, ab.PLANAR, sp))
            ab.check(lib.ab_crt_inv_dev(plan.h, yp, batch, 0, L, ab.PLANAR, sp))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(iters):
            ab.check(fn(plan.h, yp, batch, 0, L, ab.PLANAR, sp))
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out[name + "_ms"] = ms
        out[name + "_gbs"] = out["bytes_per_launch"] / (ms * 1e-3) / 1e9
        out[name + "_frac"] = out[name + "_gbs"] / peak
        # undo the `iters` one-directional applications so that the round-trip check below sees the input
        for _ in range(iters):
            ab.check((lib.ab_crt_inv_dev if name == "fwd" else lib.ab_crt_dev)(plan.h, yp, batch, 0, L, ab.PLANAR, sp))
    torch.cuda.synchronize()
    assert torch.equal(head, y[:2]), "mixed-radix crtInv . crt != id"
    out["transforms_per_s"] = 2 * batch * L / ((out["fwd_ms"] + out["inv_ms"]) * 1e-3)
    out["frac"] = (out["fwd_frac"] + out["inv_frac"]) / 2
    return out


# ------------------------------------------------------------------------------------------ GPU legs
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist

    import alchemy_b200 as ab

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus or world == 1, f"launched {world} ranks for --gpus {args.gpus}"
    n_gpus = world

    m, n, moduli, batch, name = workload(args)
    L = len(moduli)
    plan = ab.Plan(m, moduli)
    lib = ab.lib()
    gen = torch.Generator(device="cuda")
    gen.manual_seed(0xA1C4E5 + rank)
    y = torch.empty((batch, L, n), dtype=torch.int64, device="cuda")
    for l, q in enumerate(moduli):
        y[:, l] = torch.randint(0, q, (batch, n), dtype=torch.int64, device="cuda", generator=gen)
    ref_sum = (y.sum(dim=(0, 2)) % (1 << 62)).clone()
    ref_head = y[:2].clone()
    stream = torch.cuda.current_stream()
    sp = C.c_void_p(stream.cuda_stream)
    yp = C.c_void_p(y.data_ptr())

    def step(ev=None):
        if ev:
            ev[0].record(stream)
        ab.check(lib.ab_crt_dev(plan.h, yp, batch, 0, L, ab.PLANAR, sp))
        if ev:
            ev[1].record(stream)
        ab.check(lib.ab_crt_inv_dev(plan.h, yp, batch, 0, L, ab.PLANAR, sp))
        if ev:
            ev[2].record(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    launches0 = lib.ab_kernel_launches()
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    t0 = time.time()
    for k in range(args.steps):
        step(evs[k])
    barrier()
    t1 = time.time()
    launches = lib.ab_kernel_launches() - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    total_ms = evs[0][0].elapsed_time(evs[-1][2])
    fwd_ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    inv_ms = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    tt = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    ms_per_step = total_ms / args.steps
    per_step = 2 * batch * L
    value = n_gpus * per_step / (ms_per_step * 1e-3)

    # size-independent check at full size: the buffer must hold the input again, bit for bit
    assert torch.equal(ref_head, y[:2]) and torch.equal(ref_sum, y.sum(dim=(0, 2)) % (1 << 62)), "crtInv . crt != id"

    # ---- end to end through the host-buffer ABI entry (pinned host memory, copies inside the timed region)
    e2e = None
    if not args.no_e2e:
        eb = max(1, batch // 4)
        host = torch.empty((eb, L, n), dtype=torch.int64).pin_memory()
        host.copy_(y[:eb])
        hp = C.c_void_p(host.data_ptr())
        keep = host[:1].clone()

        def e2e_step():
            ab.check(lib.ab_tensor_op_host(plan.h, 0, hp, eb, ab.PLANAR))
            ab.check(lib.ab_tensor_op_host(plan.h, 1, hp, eb, ab.PLANAR))

        for _ in range(2):
            e2e_step()
        ksteps = max(3, min(args.steps, 10))
        barrier()
        t = time.perf_counter()
        for _ in range(ksteps):
            e2e_step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        assert torch.equal(keep, host[:1])
        nbytes = eb * L * n * 8
        e2e = {"value": n_gpus * 2 * eb * L * ksteps / float(dt.item()), "unit": UNIT, "h2d_bytes_per_step": 2 * nbytes,
               "d2h_bytes_per_step": 2 * nbytes, "batch": eb, "steps": ksteps,
               "api": "ab_tensor_op_host(CRT) + ab_tensor_op_host(CRT_INV), pinned host buffers"}

    # ---- secondary metric M2 (rank 0, N = 1 only): ciphertext mul + key-switch pipelines per second
    m2 = m3 = mx = None
    if rank == 0 and world == 1 and not args.no_m2:
        mx = mixed_leg(ab, torch, peaks()[0])
        m2 = mul_keyswitch_leg(ab, torch)
        # M3: the ring-tunnelling chain of examples/Tunnel.hs on a slice of the 4096-ciphertext workload (tools/bench_she.py)
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        from bench_she import tunnel_chain
        m3 = tunnel_chain(1024, 2, 0, 1)

    # ---- gather a checksum of every rank's result over NCCL (the only collective of the path)
    if world > 1:
        sums = [torch.zeros_like(ref_sum) for _ in range(world)]
        dist.all_gather(sums, ref_sum)

    if rank == 0:
        peak, peak_src = peaks()
        bytes_per_launch = 16 * n * batch * L
        mean_launch_ms = (fwd_ms + inv_ms) / 2
        achieved = bytes_per_launch / (mean_launch_ms * 1e-3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 (int64 residues in HBM)",
                "data": "synthetic", "config": {"workload": name, "m": m, "n": n, "limbs": L, "batch_per_gpu": batch,
                                                 "l2": "in-place buffer of %.0f MB per GPU, larger than the 126 MB L2" % (batch * L * n * 8 / 1e6),
                                                 "step": "1 CRT launch + 1 crtInv launch over the whole batch"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": measured_traffic(m, L, batch),
                             "peak_source": peak_src, "kernel": ("k_crt_pow2_fwd_c2 (2-CTA cluster) / k_crt_pow2_inv_h2 (two independent half-transform CTAs)" if n == 32768 else "k_crt_pow2_fwd / k_crt_pow2_inv") if not args.m else "k_crt_mixed",
                             "bytes_per_launch": bytes_per_launch, "fwd_ms": fwd_ms, "inv_ms": inv_ms,
                             "fwd_gbs": bytes_per_launch / (fwd_ms * 1e-3) / 1e9, "inv_gbs": bytes_per_launch / (inv_ms * 1e-3) / 1e9},
                "gpu_launches": int(launches), "clocks": clocks}
        if e2e:
            line["e2e"] = e2e
        if mx:
            line["crt_mixed"] = mx
        if m2:
            line["mul_keyswitch"] = m2
        if m3:
            line["tunnel_chain"] = m3
        if not args.no_cpu:
            line["cpu_baseline"], _ = cpu_leg(m, n, moduli, target_s=10.0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
