"""Aggregate ncu per-SASS samples by CUDA source line using nvdisasm -g line info.
usage: python tools/ncu_by_line.py <src.csv from `ncu --page source --csv`> <cubin> <mangled kernel substring> [top]"""
import collections
import csv
import re
import subprocess
import sys

src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of = {}
cur, infunc = None, False
for ln in dis:
    if ln.startswith(".text."):
        infunc = kname in ln
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
data = [r for r in rows if len(r) == len(hdr) and r[0].startswith("0x")]
base = int(data[0][0], 16)
seen = set()
isamp, iex, ia = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
agg = collections.defaultdict(lambda: [0, 0, 0])
tot_s = tot_e = 0
for r in data:
    off = int(r[0], 16) - base
    if off in seen:
        continue
    seen.add(off)
    key = line_of.get(off, ("?", 0))
    a = agg[key]
    a[0] += int(r[isamp]); a[1] += int(r[iex]); a[2] += 1
    tot_s += int(r[isamp]); tot_e += int(r[iex])
print(f"total samples {tot_s}, warp-instr {tot_e}, sass {len(seen)}")
for key, (s, e, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{key[0]}:{key[1]:<5d} samples {s / tot_s * 100:5.1f}%  exec {e / tot_e * 100:5.1f}%  sass {n}")
