"""per-source-line executed warp-instructions of one kernel: ncu report x nvdisasm -g line table (run here, no GPU).
usage: linemix.py report.ncu-rep libigate_dsp.so kernel_substr [nbf] [launch index in the report, default 0]"""
import csv, io, re, subprocess, sys, collections, os, tempfile
rep, so, ksub = sys.argv[1:4]
nbf = float(sys.argv[4]) if len(sys.argv) > 4 else 1024 * 1640
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
dis = []                          # every embedded cubin (one per translation unit)
for cub in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
    dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.splitlines()
# instruction list (offset, line, text) of the kernel
ins, cur, on = [], None, False
for l in dis:
    if l.startswith(".text."):
        on = ksub in l
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), cur, m.group(2).strip()))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hd, body, seen = None, [], -1
for r in rows:                       # one captured launch only
    if r and r[0] == "Kernel Name":
        continue
    if r and r[0] == "Address":
        seen += 1
        if seen > which:
            break
        hd, body = r, []
        continue
    if hd and seen == which:
        body.append(r)
ia = hd.index("Instructions Executed"); isrc = hd.index("Source")
assert len(body) == len(ins), (len(body), len(ins))
per = collections.Counter(); ops = collections.defaultdict(collections.Counter)
for (off, ln, txt), r in zip(ins, body):
    n = int(r[ia])
    per[ln] += n
    op = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", txt).group(2)
    ops[ln][op] += n
tot = sum(per.values())
print(f"total {tot / nbf:.1f} warp-instr per bridge-frame")
lines = {}
for f in set(k[0] for k in per if k):
    for d in ("/root/repo/igate4xsoftphonedsp_b200/csrc/",):
        if os.path.exists(d + f):
            lines[f] = open(d + f).read().splitlines()
for ln, n in sorted(per.items(), key=lambda kv: (kv[0] or ("", 0))):
    if n / nbf < 0.05:
        continue
    text = lines.get(ln[0], [""] * 99999)[ln[1] - 1].strip()[:70] if ln else ""
    top = " ".join(f"{k}:{v / nbf:.1f}" for k, v in ops[ln].most_common(4))
    print(f"{ln[0] if ln else '?':16s}{ln[1] if ln else 0:5d} {n / nbf:7.2f}  {text:70s} | {top}")
