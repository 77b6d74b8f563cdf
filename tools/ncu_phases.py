"""Condense the SASS source page of an ncu report (`ncu -i X.ncu-rep --page source --csv`) into phases: the instruction stream of
each kernel is cut at BAR.SYNC / WARPSYNC instructions and every phase reports its share of stall samples, of executed warp
instructions, the dominant stall reasons and the opcode mix.  Usage: python tools/ncu_phases.py file.csv [kernel-substring]"""
import csv
import sys
from collections import Counter, defaultdict


def main(path, want=""):
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "ins": []}
            kernels.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
            cur["ins"].append(r)
    for k in kernels:
        if want not in k["name"]:
            continue
        h = {n: i for i, n in enumerate(k["hdr"])}
        stall_cols = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
        tot_s = sum(int(r[h["# Samples"]] or 0) for r in k["ins"])
        tot_i = sum(int(r[h["Instructions Executed"]] or 0) for r in k["ins"])
        print(f"== {k['name']}: {len(k['ins'])} SASS instructions, {tot_s} samples, {tot_i} warp instructions")
        phase, phases = {"s": 0, "i": 0, "st": Counter(), "op": Counter(), "n": 0, "wf": 0, "wfi": 0}, []
        for r in k["ins"]:
            sass = r[h["Source"]].strip()
            op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
            s, i = int(r[h["# Samples"]] or 0), int(r[h["Instructions Executed"]] or 0)
            phase["s"] += s; phase["i"] += i; phase["n"] += 1
            phase["op"][op.split(".")[0]] += i
            phase["wf"] += int(r[h["L1 Wavefronts Shared"]] or 0); phase["wfi"] += int(r[h["L1 Wavefronts Shared Ideal"]] or 0)
            for c in stall_cols:
                phase["st"][c] += int(r[h[c]] or 0)
            if op.startswith("BAR") or op.startswith("EXIT"):
                phases.append(phase)
                phase = {"s": 0, "i": 0, "st": Counter(), "op": Counter(), "n": 0, "wf": 0, "wfi": 0}
        if phase["n"]:
            phases.append(phase)
        for n, p in enumerate(phases):
            if p["i"] == 0 and p["s"] == 0:
                continue
            st = ", ".join(f"{c[6:]} {v * 100 // max(1, p['s'])}%" for c, v in p["st"].most_common(4))
            ops = ", ".join(f"{o} {v * 100 // max(1, p['i'])}%" for o, v in p["op"].most_common(6))
            print(f" phase {n:2d}: {p['n']:4d} sass, samples {p['s'] * 100.0 / max(1, tot_s):5.1f}%, warp-instr {p['i'] * 100.0 / max(1, tot_i):5.1f}% ({p['i']}), smem wavefronts {p['wf']} (ideal {p['wfi']}) | {st} | {ops}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
