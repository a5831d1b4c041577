#!/usr/bin/env python
"""Opcode histogram (executed warp instructions, stall samples) from `ncu --page source --csv` output."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ex, st = Counter(), Counter()
tot_e = tot_s = 0
top = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr) - 5 or r[ix["Source"]] == "Source":
        continue
    op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith("@")]
    if not op:
        continue
    o = op[0].split(".")[0]
    if o in ("LDS", "STS", "LDG", "STG", "LDL", "STL", "LD", "ST"):
        o = op[0]
    e = int(r[ix["Instructions Executed"]] or 0)
    s = int(r[ix["# Samples"]] or 0)
    ex[o] += e
    st[o] += s
    tot_e += e
    tot_s += s
    top.append((s, n, r[ix["Source"]].strip()[:70]))
print(f"total executed {tot_e}  samples {tot_s}  static {len(rows)-2}")
for o, e in ex.most_common(28):
    print(f"{o:14s} exec {e:10d} {100*e/tot_e:5.1f}%   samples {st[o]:7d} {100*st[o]/max(tot_s,1):5.1f}%")
if len(sys.argv) > 2:
    print("top stall sites:")
    for s, n, t in sorted(top, reverse=True)[:int(sys.argv[2])]:
        print(f"  {s:6d}  #{n:5d}  {t}")
