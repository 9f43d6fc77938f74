#!/usr/bin/env python
"""GPU box: per-feature parity of the CUDA render against the CPU oracle (same seeds), one improvement at a time."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from common import load_cbox, relmse
from ppg_b200.integrator import GuidedPathTracer
size = 128
sc = load_cbox(size)
variants = {"default": {}, "dbox": {"directionalFilter": "box"}, "stochastic": {"spatialFilter": "stochastic"}, "sbox": {"spatialFilter": "box"},
            "kl": {"bsdfSamplingFractionLoss": "kl"}, "var": {"bsdfSamplingFractionLoss": "var"}, "inversevar": {"sampleCombination": "inversevar"}, "spp1": {"sppPerPass": "1"},
            "thr4000": {"sTreeThreshold": "4000"}, "kickstart": {"nee": "kickstart"}, "always": {"nee": "always"}, "combo": {"nee": "kickstart", "spatialFilter": "stochastic", "directionalFilter": "box", "budget": "300"},
            "kick_stoch": {"nee": "kickstart", "spatialFilter": "stochastic"}, "kick_dbox": {"nee": "kickstart", "directionalFilter": "box"}, "kick300": {"nee": "kickstart", "budget": "300"}}
sel = sys.argv[1:] or list(variants)
for name in sel:
    props = dict(dict(sc.integrator, budget="60"), **variants[name])
    o = O.Oracle(O.params_from_xml(props), sc, kind="port"); oi, ost = o.render()
    g = GuidedPathTracer(props); g.set_scene(sc); gi, gst = g.render()
    ow = [round(i["weight_avg"] * i["s_tree_leaves"]) for i in ost["iterations"]]; gw = [round(i["weight_avg"] * i["s_tree_leaves"]) for i in gst["iterations"]]
    print(name, "relMSE(gpu,oracle)=%.3e" % relmse(gi, oi), "var o/g", [round(i["variance"], 5) for i in ost["iterations"]], [round(i["variance"], 5) for i in gst["iterations"]],
          "weights o/g", ow, gw, "leaves", [i["s_tree_leaves"] for i in ost["iterations"]], [i["s_tree_leaves"] for i in gst["iterations"]], flush=True)
    o.close(); g.close()
