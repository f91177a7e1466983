#!/usr/bin/env python
"""Stall samples and executed instructions of an ncu capture attributed to the source lines of ONE function body (the outermost
location inside [lo, hi] of `file` on each instruction's inline chain): which phase of a fused kernel the time goes to.
usage: python profiles/by_phase.py <source.csv> <nvdisasm -gi of the library> <mangled kernel name> <lo> <hi> <file> [units]"""
import collections, csv, re, sys
src_csv, dis, fn = sys.argv[1], sys.argv[2], sys.argv[3]
lo, hi, fname = int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
rows, hdr = [], None
for r in csv.reader(open(src_csv)):
    if r and r[0] == "Address": hdr = r
    elif len(r) > 6 and r[0].startswith("0x"): rows.append(r)
iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(rows[0][0], 16)
by_off = {int(r[0], 16) - base: (int(r[isamp]), int(r[iex])) for r in rows}
inside, chain, line_of, last = False, [], {}, None
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + fn; continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: chain.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m2 = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m2:
        off = int(m2.group(1), 16)
        if chain:
            pick = None
            for f, l in chain:  # inner -> outer: keep the outermost inside [lo, hi] of fname
                if f == fname and lo <= l <= hi: pick = l
            last = pick
        line_of[off] = last
        chain = []
samp, inst = collections.Counter(), collections.Counter()
for off, (s, e) in by_off.items():
    k = line_of.get(off)
    samp[k] += s; inst[k] += e
tot = sum(samp.values()); ti = sum(inst.values())
units = float(sys.argv[7]) if len(sys.argv) > 7 else 1
for k in sorted(samp, key=lambda x: (x is None, x)):
    if samp[k] * 200 > tot or inst[k] * 200 > ti:
        print(f"  line {k}: samples {100*samp[k]/tot:5.1f}%  instr {inst[k]/units:9.1f}/unit ({100*inst[k]/ti:4.1f}%)")
