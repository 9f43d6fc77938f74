#!/usr/bin/env python
"""GPU box: CUDA render vs oracle on the BVH-path test scene (CBOX + tessellated sphere)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from common import cbox_with_sphere, relmse
from ppg_b200.integrator import GuidedPathTracer
for subdiv, smooth in ((2, True), (2, False), (3, False)):
    sc = cbox_with_sphere(128, subdiv=subdiv, smooth=smooth)
    for budget in ("4", "60"):
        props = dict(sc.integrator, budget=budget)
        o = O.Oracle(O.params_from_xml(props), sc, kind="port"); oi, ost = o.render()
        g = GuidedPathTracer(props); g.set_scene(sc); gi, gst = g.render()
        bad = ~np.isclose(gi, oi, rtol=1e-3, atol=1e-4).all(axis=2)
        print(subdiv, smooth, budget, "relMSE %.3e" % relmse(gi, oi), "verts o/g", ost["total_vertices"], gst["total_vertices"], "bad pixels", int(bad.sum()),
              "means", float(oi.mean()), float(gi.mean()), "first bad", np.argwhere(bad)[:5].tolist(), flush=True)
        o.close(); g.close()
