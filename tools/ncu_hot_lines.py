#!/usr/bin/env python
"""ncu -i <rep> --page source --csv --kernel-name regex:<k> --print-source cuda,sass > file.csv; this script lists the
source lines / SASS instructions with the most warp-stall samples and their dominant stall reasons."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
secs, cur = [], None
for r in rows:
    if r and r[0] == "Line No":
        cur = {"hdr": r, "rows": []}
        secs.append(cur)
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
num = lambda x: int(x) if x.strip().lstrip("-").isdigit() else 0
s = secs[which]
h = s["hdr"]
i_s = h.index("# Samples")
stall_cols = [i for i, n in enumerate(h) if n.startswith("stall_") and not n.endswith("_not_issued")]
tot = sum(num(r[i_s]) for r in s["rows"])
print("sections", [len(x["rows"]) for x in secs], "total samples", tot)
for r in sorted(s["rows"], key=lambda r: -num(r[i_s]))[:top_n]:
    st = sorted(((num(r[i]), h[i][6:]) for i in stall_cols), reverse=True)[:3]
    st = " ".join(f"{n}:{c}" for c, n in st if c)
    print(f"{num(r[i_s]):6d} {100.0 * num(r[i_s]) / max(tot, 1):5.1f}%  L{r[0]:>5s} {r[3][:70]:70s} | {st}")
