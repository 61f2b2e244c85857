#!/usr/bin/env python
"""Turns the artefacts of profiles/capture.sh (gpurun_out/launches.csv, gpurun_out/prof_fused.ncu-rep)
into the committed summaries under profiles/.  Run here (ncu CLI, no GPU needed):
    python profiles/summarize.py r01                       (artefacts in gpurun_out/)
    python profiles/summarize.py r02 gpurun_out/r02        (artefacts of profiles/capture_r02.sh)
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
GP = os.path.join(ROOT, sys.argv[2]) if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")

# ---- launch list: per-kernel share of the step
rows = [r for r in csv.reader(l for l in open(os.path.join(GP, "launches.csv")) if not l.startswith("=="))]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0].strip()
    t = float(r[iv].replace(",", ""))
    c = agg.setdefault(name, [0, 0.0])
    c[0] += 1
    c[1] += t
tot = sum(v[1] for v in agg.values())
with open(os.path.join(OUT, f"{tag}_launches_summary.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 5 --warmup 3 "
            "--no-cpu --no-e2e (per-launch times are cold-cache and serialised: read the SHARES)\n")
    f.write("kernel,launches,total_us,avg_us,share\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{n},{t / 1e3:.1f},{t / 1e3 / n:.1f},{t / tot:.4f}\n")

# ---- full capture of the fused kernel
rep = os.path.join(GP, "prof_fused.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units, data = rr[0], rr[1], rr[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__warps_eligible.avg.per_cycle_active"]
want += [x for x in h if "issue_stalled" in x and x.endswith("per_issue_active.ratio") and "not_issued" not in x]
lines = []
vals = {}
for w in want:
    if w in h:
        i = h.index(w)
        v = [d[i] for d in data]
        vals[w] = v
        lines.append(f"{w:88s} {units[i]:14s} {'  '.join(v)}")
def gb(name):
    i = h.index(name)
    u = units[i].lower()
    f = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}[u]
    return [float(d[i]) * f for d in data]
rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
traffic = sum(rd) / len(rd) + sum(wr) / len(wr)
json.dump({"kernel": data[0][h.index("Kernel Name")], "dram_bytes_per_launch": traffic,
           "dram_read_bytes": sum(rd) / len(rd), "dram_write_bytes": sum(wr) / len(wr),
           "algorithmic_bytes_per_launch": 1184 * 1024 * 1640,
           "source": f"profiles/{tag}_fused_ncu.txt (ncu --set full --clock-control none, {len(data)} launches)"},
          open(os.path.join(OUT, "fused_traffic.json"), "w"), indent=1)
with open(os.path.join(OUT, f"{tag}_fused_ncu.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c {len(data)}\n"
            "# command: python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e ; kernel: "
            f"{data[0][h.index('Kernel Name')]}\n# one column per captured launch\n")
    f.write("\n".join(lines) + "\n")
    f.write(f"\ndram traffic per launch = {traffic / 1e9:.4f} GB (algorithmic {1184 * 1024 * 1640 / 1e9:.4f} GB)\n")

# ---- executed-instruction mix of the fused kernel (source page)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
import re
rows = list(csv.reader(io.StringIO(src)))
hd, body = None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        if hd is not None:
            break
        continue
    if r and r[0] == "Address":
        hd = r
        continue
    if hd:
        body.append(r)
ia, isrc, ist = hd.index("Instructions Executed"), hd.index("Source"), hd.index("Warp Stall Sampling (All Samples)")
ops, st = collections.Counter(), collections.Counter()
for r in body:
    m = re.match(r"\s*(@!?U?P\d\s+)?([A-Z0-9_.]+)", r[isrc])
    if m:
        ops[m.group(2)] += int(r[ia])
        st[m.group(2)] += int(r[ist])
nbf = 1024 * 1640
with open(os.path.join(OUT, f"{tag}_fused_instruction_mix.txt"), "w") as f:
    f.write("# executed warp-instructions of k_fused per bridge-frame (640 channel-samples), from the ncu source page\n")
    f.write(f"total {sum(ops.values()) / nbf:.1f} warp-instr / bridge-frame\n")
    for k, v in ops.most_common(40):
        f.write(f"{k:34s} {v / nbf:8.2f}   stall samples {st[k]}\n")
print(open(os.path.join(OUT, f"{tag}_launches_summary.csv")).read())
print(open(os.path.join(OUT, f"{tag}_fused_ncu.txt")).read())
