#!/usr/bin/env python
"""ncu_opmix.py — executed warp-instructions per 32 node-stages by opcode, and where the stall samples fall, from the
source page of an ncu report:   ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv ; python tools/ncu_opmix.py src.csv NODES"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nodes = float(sys.argv[2]) / 32 if len(sys.argv) > 2 else 512 ** 3 / 32
hdr = None
k = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        cls, samp = collections.Counter(), collections.Counter()
        k += 1
        data = []
        continue
    if r and r[0] == "Kernel Name":
        if hdr and data:
            break          # first kernel only
        print(r[1][:150])
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
tot = 0
dp = 0
for r in data:
    src = r[ix["Source"]].strip().split()
    op = src[1] if src[0].startswith("@") else src[0]
    op = op.split(".")[0]
    n = int(r[ix["Instructions Executed"]])
    cls[op] += n
    samp[op] += int(r[ix["# Samples"]])
    tot += n
    if op in ("DFMA", "DMUL", "DADD", "DSETP"):
        dp += n
print(f"total {tot / nodes:.1f} warp-instructions per 32 node-stages, FP64-pipe {dp / nodes:.1f}, other {(tot - dp) / nodes:.1f}; "
      f"issue-model slots 2*DP+other = {(tot + dp) / nodes:.1f}")
for op, n in cls.most_common(45):
    print("%-10s %8.2f   samples %5.1f%%" % (op, n / nodes, 100 * samp[op] / max(tot_samp, 1)))
