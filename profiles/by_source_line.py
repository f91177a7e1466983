#!/usr/bin/env python
"""Attributes the stall samples of an ncu capture to CUDA source lines: joins the SASS rows of `ncu --page source`
with the line table of the matching function in `nvdisasm -gi` output.
usage: python profiles/by_source_line.py <source.csv> <nvdisasm_all.txt> <mangled function name> [top]"""
import collections
import csv
import re
import sys

src_csv, dis, fn = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows, hdr = [], None
for r in csv.reader(open(src_csv)):
    if r and r[0] == "Address":
        hdr = r
    elif len(r) > 6 and r[0].startswith("0x"):
        rows.append(r)
iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[0][0], 16)
by_off = {int(r[0], 16) - base: (int(r[isamp]), int(r[iex])) for r in rows}
# line table: the LAST "//## File" comment before an instruction is the outermost (non-inlined) location
inside, loc_chain, line_of = False, [], {}
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + fn
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        loc_chain.append(m)
        continue
    m2 = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m2:
        off = int(m2.group(1), 16)
        if loc_chain:
            inner, outer = loc_chain[0], loc_chain[-1]
            line_of[off] = ((inner.group(1).split("/")[-1], int(inner.group(2))), (outer.group(1).split("/")[-1], int(outer.group(2))))
            last = line_of[off]
        else:
            line_of[off] = last
        loc_chain = []
tot = sum(v[0] for v in by_off.values())
outer, inner = collections.Counter(), collections.Counter()
oinst = collections.Counter()
for off, (s, e) in by_off.items():
    if off in line_of:
        outer[line_of[off][1]] += s
        inner[line_of[off][0]] += s
        oinst[line_of[off][1]] += e
print(f"total samples {tot}")
print("-- by outermost source line (share of samples, warp-instructions executed)")
for k, v in outer.most_common(top):
    print(f"  {100 * v / tot:5.1f}%  {oinst[k]:>10d}  {k[0]}:{k[1]}")
print("-- by innermost (inlined) source line")
for k, v in inner.most_common(top):
    print(f"  {100 * v / tot:5.1f}%  {k[0]}:{k[1]}")
