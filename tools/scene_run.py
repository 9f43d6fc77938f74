#!/usr/bin/env python
"""GPU box: one render of a fixture scene (for ncu captures / timing).  usage: scene_run.py <fixture> W H budget [name=value ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import load_fixture_scene
from ppg_b200.integrator import GuidedPathTracer
name, W, H, budget = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
sc = load_fixture_scene(name)
if W: sc = sc.with_film(W, H)
props = dict(sc.integrator, budget=budget)
for kv in sys.argv[5:]:
    k, v = kv.split("="); props[k] = v
g = GuidedPathTracer(props); g.set_scene(sc)
t = time.time(); ptr, st = g.render_device(); dt = time.time() - t
print("Msamples/s %.1f" % (st["total_vertices"] / st["render_device_ms"] / 1e3), "device ms %.1f" % st["render_device_ms"], "wall %.2f" % dt, "verts/path %.2f" % (st["total_vertices"] / st["total_paths"]),
      {k: round(v, 1) for k, v in st["kernel_ms"].items()}, "launches", st["kernel_launches"], "invalid rays", st["invalid_rays"], "truncated", st["truncated_paths"], "sub-batches", st["sub_batches"])
for it in st["iterations"]:
    print(it["iteration"], it["passes"], "sec %.3f" % it["seconds"], "verts", it["vertices"], "Mverts/s %.1f" % (it["vertices"] / max(it["seconds"], 1e-9) / 1e6), "leaves", it["s_tree_leaves"], "d_S %.1f" % it["s_tree_depth_avg"], "d_D %.1f" % it["depth_avg"])
