"""Per-source-line aggregation of an `ncu --page source --csv` export with the dominant stall reasons.
usage: python tools/ncu_by_line2.py <src.csv> <cubin> <kernel substring> [top] [exclude-line-regex]"""
import collections, csv, re, subprocess, sys
src_csv, cubin, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, infunc = {}, None, False
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
isamp, iex = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
tot_s = tot_e = 0
seen = set()
for r in data:
    off = int(r[0], 16) - base
    if off in seen:
        continue
    seen.add(off)
    key = line_of.get(off, ("?", 0))
    a = agg[key]
    a[0] += int(r[isamp]); a[1] += int(r[iex])
    for i, n in stall_cols:
        if r[i] not in ("", "0"):
            a[2][n] += int(r[i])
    tot_s += int(r[isamp]); tot_e += int(r[iex])
print(f"total samples {tot_s}, warp-instr {tot_e}")
for key, (s_, e, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    tops = ", ".join(f"{n}:{v}" for n, v in st.most_common(3))
    print(f"{key[0]}:{key[1]:<5d} samples {s_ / tot_s * 100:5.1f}%  exec {e / tot_e * 100:5.1f}%  [{tops}]")
