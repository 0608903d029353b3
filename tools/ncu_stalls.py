#!/usr/bin/env python3
"""Warp-stall samples of one kernel by reason and by SASS opcode, from an .ncu-rep taken with --import-source on.
usage: ncu_stalls.py report.ncu-rep > profiles/NAME.txt"""
import collections, csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
print("kernel:", rows[0][1])
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, byop, nsamp = collections.Counter(), collections.defaultdict(collections.Counter), collections.Counter()
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    p = r[col["Source"]].split()
    op = (p[1] if p[0].startswith("@") else p[0]).rstrip(";")
    nsamp[op] += int(r[col["# Samples"]] or 0)
    for s in stalls:
        v = int(r[col[s]] or 0)
        tot[s] += v
        byop[op][s] += v
T = sum(tot.values())
print(f"warp-stall samples: {T}")
for s, v in tot.most_common(8):
    print(f"  {s:24s} {v:9d} {100 * v / T:5.1f} %")
print("by opcode: samples, share, top reasons")
for op, n in nsamp.most_common(16):
    c = byop[op]
    tt = max(sum(c.values()), 1)
    print(f"  {op:22s} {n:8d} {100 * n / T:5.1f} %  " + ", ".join(f"{k[6:]} {100 * v / tt:.0f} %" for k, v in c.most_common(4)))
