"""Opcode histogram of the first kernel of an ncu source page (`ncu -i X.ncu-rep --page source --csv`), in executed warp instructions
and, given the residues processed per launch, per residue.  Usage: python tools/ncu_opmix.py file.csv [residues-per-launch] [kernel#]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
res = float(sys.argv[2]) if len(sys.argv) > 2 else 0
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
k, hdr, n, name = [], None, -1, ""
for r in rows:
    if r and r[0] == "Kernel Name":
        n += 1
        if n == which:
            name = r[1]
    elif r and r[0] == "Address":
        hdr = r
    elif hdr and len(r) > 10 and n == which:
        k.append(r)
h = {x: i for i, x in enumerate(hdr)}
ops, tot = Counter(), 0
for r in k:
    s = r[h["Source"]].strip().split()
    op = s[1] if s[0].startswith("@") else s[0]
    i = int(r[h["Instructions Executed"]] or 0)
    ops[op] += i
    tot += i
print(name, tot, "warp instructions", f"({tot * 32 / res:.1f} per residue)" if res else "")
for o, v in ops.most_common(40):
    print(f"{o:28s} {v:12d} {100 * v / tot:5.1f}%" + (f"  {v * 32 / res:6.2f}/residue" if res else ""))
