#!/usr/bin/env python
"""Calibration script (GPU box): equal-spp relMSE of the CUDA render vs the CPU oracle for the default and the
'improved' CBOX configurations, against a converged reference assembled from many oracle renders."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from common import load_cbox, relmse
from ppg_b200.integrator import GuidedPathTracer

size = int(sys.argv[1]) if len(sys.argv) > 1 else 128
budget = sys.argv[2] if len(sys.argv) > 2 else "127"
nref = int(sys.argv[3]) if len(sys.argv) > 3 else 16
sc = load_cbox(size)
refs = []
for seed in range(nref):
    o = O.Oracle(O.params_from_xml(dict(sc.integrator, budget="508"), seed=500 + seed), sc, kind="port")
    refs.append(o.render()[0]); o.close()
ref = np.mean(refs, axis=0)
out = {}
for name, improved in (("default", False), ("improved", True)):
    s2 = load_cbox(size, improved=improved)
    props = dict(s2.integrator, budget=budget)
    ro, rg, fr = [], [], []
    for seed in range(4):
        o = O.Oracle(O.params_from_xml(props, seed=seed), s2, kind="port")
        oi, ost = o.render()
        e = o.export(0)
        g = GuidedPathTracer(dict(props, seed=str(seed))); g.set_scene(s2)
        gi, gst = g.render()
        ro.append(relmse(oi, ref)); rg.append(relmse(gi, ref))
        th = e["adam"][e["s_is_leaf"] == 1][:, 3]
        fr.append(float(np.mean(1 / (1 + np.exp(-th)))))
        o.close(); g.close()
    out[name] = {"relmse_oracle": ro, "relmse_gpu": rg, "mean_oracle": float(np.mean(ro)), "mean_gpu": float(np.mean(rg)), "oracle_mean_fraction": fr,
                 "gpu_var": gst["final_variance"], "oracle_var": ost["final_variance"]}
    print(name, json.dumps(out[name]), flush=True)
