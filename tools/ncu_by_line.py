#!/usr/bin/env python
"""Aggregate an ncu SASS source page (csv) by CUDA source line using nvdisasm -g line info.
usage: ncu_by_line.py <report.ncu-rep> <kernel-regex> <mangled-function-substring> [launch-skip]"""
import csv, re, subprocess, sys, collections, os, tempfile
rep, kre, fn = sys.argv[1], sys.argv[2], sys.argv[3]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
so = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "practical-path-guiding_b200", "csrc", "libppg_b200.so")
dis = []
build = os.path.join(os.path.dirname(os.path.abspath(so)), "build")
for obj in sorted(os.listdir(build)):          # one object (one cubin) per translation unit: take the one that holds the function
    if not obj.endswith(".o"): continue
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(build, obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
    for cubin in os.listdir(tmp):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        if any(l.startswith("//--------------------- .text.") and fn in l for l in txt.splitlines()):
            dis = txt.splitlines()
    if dis: break
# map instruction ordinal within function -> (file, line) with inline chain's outermost user line
infn = False; cur = None; amap = []
for l in dis:
    if l.startswith("//--------------------- .text."):
        infn = fn in l; continue
    if not infn: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        amap.append((int(m.group(1), 16), cur, m.group(2)))
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
hdr = next(r for r in rows if r and r[0] == "Address")
H = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if len(r) == len(hdr) and r[0] != "Address"]
def num(x):
    try: return float(x.replace(",", ""))
    except Exception: return 0.0
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
if len(data) == 2 * len(amap):      # the CLI prints the instruction table twice
    data = data[:len(amap)]
if len(data) != len(amap):
    print("warning: instruction count mismatch", len(data), len(amap), file=sys.stderr)
for i, r in enumerate(data):
    key = amap[i][1] if i < len(amap) else None
    a = agg[key]; a[0] += num(r[H["# Samples"]]); a[1] += num(r[H["Instructions Executed"]]); a[2] += num(r[H["Thread Instructions Executed"]])
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
src = {}
for k in agg:
    if k and k[0] not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(so)), k[0])
        src[k[0]] = open(p).read().splitlines() if os.path.exists(p) else []
print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:int(os.environ.get("TOP", "40"))]:
    text = src[k[0]][k[1] - 1].strip()[:110] if k and src.get(k[0]) and k[1] <= len(src[k[0]]) else ""
    print(f"{a[0]/ts*100:5.1f}% smp {a[1]/ti*100:5.1f}% inst thr/inst {a[2]/max(a[1],1):5.1f} | {k[0] if k else '?'}:{k[1] if k else 0:4d} {text}")
