"""Turn the ncu captures of a gpurun call into the tracked summaries under profiles/.
usage (here, no GPU): python tools/make_profiles.py [tag]
expects gpurun_out/<tag>_bench_tc.ncu-rep (--set full of both tensor kernels at the bench size) and
gpurun_out/<tag>_launches.csv (--metrics gpu__time_duration.sum launch list of bench.py)."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rep = os.path.join(GO, f"{tag}_bench_tc.ncu-rep")
so = os.path.join(ROOT, "forecasting-influenza-using-universal-differential-equations_b200", "csrc", "libude_b200.so")


def sh(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw).stdout


# ---- full-metric summary + traffic
raw = list(csv.reader(sh(["ncu", "-i", rep, "--page", "raw", "--csv"]).splitlines()))
hdr, units, rows = raw[0], raw[1], raw[2:]
keep = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio"]
keep += [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
with open(os.path.join(PR, f"{tag}_tcgen05_bench_ncu_full_metrics.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows))])
    for h in keep:
        if h in hdr:
            i = hdr.index(h)
            w.writerow([h, units[i]] + [r[i] for r in rows])
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
ki, ri, wi, ti = (hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum"))
tr = {"batch": 1 << 20, "grid_steps": 8,
      "source": f"ncu --set full --clock-control none, python bench.py --steps 1 --warmup 3 --no-cpu-baseline (profiles/{tag}_tcgen05_bench_ncu_full_metrics.csv)"}
for r in rows:
    k = "bwd" if "bwd" in r[ki] else "fwd"
    rd, wr = float(r[ri]) * scale[units[ri]], float(r[wi]) * scale[units[wi]]
    tr[k + "_dram_bytes_per_launch"] = int(rd + wr)
    tr[k + "_dram_read_bytes"], tr[k + "_dram_write_bytes"], tr[k + "_ncu_ms"] = int(rd), int(wr), float(r[ti])
json.dump(tr, open(os.path.join(PR, f"{tag}_traffic.json"), "w"), indent=1)

# ---- launch list
src = os.path.join(GO, f"{tag}_launches.csv")
lrows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in lrows:
    name = r[4].split("(")[0].replace("void ", "")[-70:]
    agg[name][0] += 1
    agg[name][1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
with open(os.path.join(PR, f"{tag}_bench_launch_summary.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n"
            "# (cold-cache, serialised launches: compare SHARES; first 400 launches = warm-up + timed steps + start of the e2e leg)\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{v[1] / 1e6:10.3f} ms {v[1] / tot * 100:6.2f}%  x{v[0]:3d}  {k}\n")
with open(os.path.join(PR, f"{tag}_bench_launches.csv"), "w") as f:
    f.write(open(src).read())

# ---- per-line / per-opcode samples
subprocess.run(["cuobjdump", "-xelf", "ude_tc_host", so], cwd=GO, capture_output=True)
cubin = os.path.join(GO, "ude_tc_host.sm_100a.cubin")
for k in ("bwd", "fwd"):
    page = sh(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:rk4_{k}"])
    path = os.path.join(GO, f"src_{tag}_{k}.csv")
    open(path, "w").write(page)
    out = sh([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line2.py"), path, cubin, f"rk4_{k}_tc_kernelILi1ELi8", "40"])
    open(os.path.join(PR, f"{tag}_tcgen05_{k}_by_line.txt"), "w").write(
        f"# per-source-line warp-state samples, rk4_{k}_tc_kernel<1,8>, bench workload (ncu --set full --import-source on); tools/ncu_by_line2.py\n" + out)
    data = list(csv.reader(page.splitlines()))
    h2 = next(r for r in data if "Source" in r and "# Samples" in r)
    body = [r for r in data if len(r) == len(h2) and r[0].startswith("0x")]
    ia, iex, isamp = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
    sc = [(i, x[6:]) for i, x in enumerate(h2) if x.startswith("stall_") and "Not Issued" not in x]
    ops = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    seen = set()
    for r in body:
        if r[0] in seen:
            continue
        seen.add(r[0])
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
        op = ".".join((m.group(2) if m else "?").split(".")[:2])
        o = ops[op]
        o[0] += int(r[isamp]); o[1] += int(r[iex])
        for i, n in sc:
            if r[i] not in ("", "0"):
                o[2][n] += int(r[i])
    ts, te = sum(o[0] for o in ops.values()), sum(o[1] for o in ops.values())
    with open(os.path.join(PR, f"{tag}_tcgen05_{k}_opcode_mix.txt"), "w") as f:
        f.write(f"# opcode mix + warp-state samples, rk4_{k}_tc_kernel<1,8>, bench workload (B=2^20, 8 steps); total samples {ts}, warp instructions {te}\n")
        for op, (s_, e, st) in sorted(ops.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"{op:24s} exec {e / te * 100:5.1f}%  samples {s_ / ts * 100:5.1f}%  top stalls {dict(st.most_common(3))}\n")
print(json.dumps(tr))
