#!/usr/bin/env python
"""Bucket the per-line ncu data of bounce_kernel by pipeline stage (uses ncu_by_line's machinery with TOP=100000)."""
import os, re, subprocess, sys
rep, fn = sys.argv[1], sys.argv[2]
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_by_line.py"), rep, "bounce" if "bounce" in fn else "commit", fn],
                     capture_output=True, text=True, env=dict(os.environ, TOP="100000")).stdout
dev = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "practical-path-guiding_b200", "csrc", "ppg_device.cuh")).read().splitlines()
# function ranges in ppg_device.cuh by scanning for __device__ definitions
marks = []
for i, l in enumerate(dev, 1):
    m = re.search(r"__device__ __forceinline__ [\w:<> ]+?[ \*&](\w+)\(", l)
    if m: marks.append((i, m.group(1)))
def fn_of(line):
    name = "?"
    for i, n in marks:
        if i <= line: name = n
    return name
group = {"tri_intersect": "intersect", "bvh_intersect": "intersect", "fill_its": "fill_its", "stree_lookup": "stree", "spread3": "stree", "voxel_size": "stree",
         "dtree_sample": "dtree_sample", "dtree_pdf": "dtree_pdf", "quad_child_index": "dtree_pdf", "sum4": "dtree(sum4/child16)", "child16": "dtree(sum4/child16)",
         "dir_to_canonical": "dir<->canonical", "canonical_to_dir": "dir<->canonical", "square_to_cosine_hemisphere": "bsdf", "bsdf_eval": "bsdf", "bsdf_pdf": "bsdf",
         "bsdf_sample": "bsdf", "load_bsdf": "bsdf", "nextU32": "rng", "next1D": "rng", "seed": "rng", "splitmix64": "rng"}
agg = {}
for l in out.splitlines()[1:]:
    m = re.match(r"\s*([\d.]+)% smp\s+([\d.]+)% inst thr/inst\s+([\d.]+) \| (\S+):\s*(\d+)", l)
    if not m: continue
    smp, inst, lanes, f, line = float(m.group(1)), float(m.group(2)), float(m.group(3)), m.group(4), int(m.group(5))
    if f == "ppg_device.cuh":
        fnn = fn_of(line); key = group.get(fnn, "vecmath/" + fnn if fnn in ("operator*", "operator+", "operator-", "dot", "cross", "normalize", "f3") else fnn)
        if key.startswith("vecmath"): key = "vecmath"
    elif f == "ppg_kernels.cuh": key = "kernel body"
    else: key = f
    a = agg.setdefault(key, [0, 0, 0]); a[0] += smp; a[1] += inst; a[2] += inst * lanes
print(out.splitlines()[0])
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{k:28s} samples {a[0]:5.1f}%  instructions {a[1]:5.1f}%  lanes/inst {a[2] / max(a[1], 1e-9):5.1f}")
