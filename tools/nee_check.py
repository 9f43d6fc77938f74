import sys, os
ROOT='/root/repo'
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from common import load_cbox, relmse
from ppg_b200.integrator import GuidedPathTracer
from ppg_b200.builtin_scenes import cbox_thin_glass, cbox_blinds
for name, mk in (("thin", cbox_thin_glass), ("blinds", cbox_blinds)):
    sc = mk(load_cbox(128))
    for seed in ("1234", "77", "20260924"):
        for budget in ("4", "12", "60"):
            props = dict(sc.integrator, budget=budget, nee="always", seed=seed)
            g = GuidedPathTracer(props); g.set_scene(sc); img, st = g.render()
            o = O.Oracle(O.params_from_xml(props), sc, kind="port"); ref, ost = o.render()
            o2 = O.Oracle(O.params_from_xml(props), sc, kind="port"); ref2, ost2 = o2.render()
            bad = ~np.isclose(img, ref, rtol=1e-3, atol=1e-5).all(axis=2)
            print(name, seed, budget, "relMSE g/o %.2e o/o %.2e" % (relmse(img, ref), relmse(ref2, ref)), "bad px", int(bad.sum()), "verts", st["total_vertices"], ost["total_vertices"],
                  "W", [round(i["weight_avg"] * i["s_tree_leaves"]) for i in st["iterations"]], [round(i["weight_avg"] * i["s_tree_leaves"]) for i in ost["iterations"]], flush=True)
