"""Per-source-line warp-sample shares from `ncu -i x.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K`
(the report carries its own sources: no cubin needed).  usage: python tools/ncu_cuda_lines.py export.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
agg, src, cur, hdr = collections.Counter(), {}, None, None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr, isamp = r, r.index("# Samples")
    elif hdr and len(r) == len(hdr) and r[0].strip().isdigit():
        try:
            s = int(r[isamp])
        except ValueError:
            s = 0
        agg[(cur, int(r[0]))] += s
        src[(cur, int(r[0]))] = r[1]
tot = sum(agg.values())
byfile = collections.Counter()
for (fn, _), s in agg.items():
    byfile[fn] += s
print(f"total warp samples {tot}; by file: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in byfile.most_common(5)))
for (fn, ln), s in agg.most_common(top):
    print(f"{100 * s / tot:5.1f}%  {fn}:{ln}  {src[(fn, ln)][:110]}")
