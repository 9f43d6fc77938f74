#!/usr/bin/env python
"""GPU box: one render of the spaceship fixture (for ncu captures / timing).  usage: spaceship_run.py [W H budget]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import load_fixture_scene
from ppg_b200.integrator import GuidedPathTracer
W, H, budget = (int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]) if len(sys.argv) > 3 else (640, 360, "15")
sc = load_fixture_scene("spaceship-improved").with_film(W, H)
g = GuidedPathTracer(dict(sc.integrator, budget=budget)); g.set_scene(sc)
ptr, st = g.render_device()
print("Msamples/s", st["total_vertices"] / st["render_device_ms"] / 1e3, "ms", st["render_device_ms"], st["kernel_ms"])
