#!/usr/bin/env python
"""Summarise an `ncu -i X.ncu-rep --page source --csv --print-source sass` export by address REGION: stall samples by reason,
instructions executed and instructions per issue slot for each range of SASS offsets.  Usage:
    python tools/ncu_src_regions.py src.csv 0x1670:0x2710=hold_cascade 0x2d90:0x3e20=hold_parallel ...
Offsets are relative to the first instruction of the kernel (the addresses cuobjdump prints)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {n: i for i, n in enumerate(hdr)}
data = rows[2:]
base = int(data[0][col["Address"]], 16)
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
regions = []
for a in sys.argv[2:]:
    rng, name = a.split("=")
    lo, hi = (int(x, 16) for x in rng.split(":"))
    regions.append((lo, hi, name))
if not regions:
    regions = [(0, 1 << 30, "all")]
tot_samples = sum(int(r[col["# Samples"]]) for r in data)
for lo, hi, name in regions:
    sel = [r for r in data if lo <= int(r[col["Address"]], 16) - base <= hi]
    samples = sum(int(r[col["# Samples"]]) for r in sel)
    inst = sum(int(r[col["Instructions Executed"]]) for r in sel)
    st = {s: sum(int(r[col[s]]) for r in sel) for s in stalls}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:9]
    print("%-18s %4d instr  samples %6.2f %%  warp-inst %.3e  | %s" % (
        name, len(sel), 100.0 * samples / tot_samples, inst,
        "  ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(samples, 1)) for k, v in top)))
