#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: one line per kernel (registers, stack, spills)."""
import re
import sys

text = open(sys.argv[1]).read()
pat = re.compile(r"Function properties for (\S+)\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n"
                 r"ptxas info\s+: Used (\d+) registers")
flt = sys.argv[2] if len(sys.argv) > 2 else ""


def short(n):
    m = re.match(r"_ZN5serac4b200\d+(\w+?_kernel)I", n)
    if not m:
        return n[:120]
    k = m.group(1)
    if k == "evaluation_kernel":
        g = re.search(r"GeometryE(\d)ELi(\d)ELi(\d)E", n)
        mat = re.search(r"solid_mechanics\d+([A-Za-z0-9]+?)E", n) or re.search(r"heat_transfer\d+([A-Za-z0-9]+?)E", n)
        wrt = re.search(r"EEELi(n?\d)ENS", n)
        tail = re.search(r"ELi(\d)ELi(\d)ELb(\d)EEEv", n)
        return (f"{k} geom={g.group(1)} q={g.group(2)} epb={g.group(3)} wrt={wrt.group(1) if wrt else '?'} "
                f"mat={mat.group(1) if mat else '?'} rows/minb/fast={tail.groups() if tail else '?'}")
    args = re.findall(r"Li(\d+)E", n)
    return f"{k}<{','.join(args)}>"


for m in pat.finditer(text):
    s = short(m.group(1))
    if flt in s:
        print(f"regs={m.group(5):>3} stack={m.group(2):>5} spill_st={m.group(3):>5} spill_ld={m.group(4):>5}  {s}")
