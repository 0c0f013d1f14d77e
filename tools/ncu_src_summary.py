"""Summarise `ncu --page source --csv` output: opcode mix and hot regions.
usage: python tools/ncu_src_summary.py src.csv [chunk]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 250
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
data = [r for r in rows if len(r) == len(hdr) and r is not hdr and r[0].startswith("0x")]
ia, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
tot = sum(int(r[isamp]) for r in data)
totex = sum(int(r[iex]) for r in data)
print("total samples", tot, "warp-instr executed", totex, "SASS lines", len(data))
ops, ops_s = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    op = m.group(2) if m else "?"
    op = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "BAR")) else op.split(".")[0]
    ops[op] += int(r[iex])
    ops_s[op] += int(r[isamp])
for op, c in ops.most_common(22):
    print(f"{op:12s} exec {c / totex * 100:5.1f}%  samples {ops_s[op] / tot * 100:5.1f}%")
print("--- regions")
for i in range(0, len(data), chunk):
    blk = data[i:i + chunk]
    s = sum(int(r[isamp]) for r in blk)
    e = sum(int(r[iex]) for r in blk)
    nf = sum(1 for r in blk if "FFMA" in r[ia])
    if s / max(tot, 1) > 0.01:
        print(f"{i:6d} samples {s / tot * 100:5.1f}% exec {e / totex * 100:5.1f}% ffma_static {nf}")
