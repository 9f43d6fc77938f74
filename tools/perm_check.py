#!/usr/bin/env python
"""GPU box: variance of the early iterations of SPACESHIP 640x360 (authors' log: 0.0976 / 0.0399 / 0.0180) for several seeds and PPG_PERM_RUN values."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import load_fixture_scene
from ppg_b200.integrator import GuidedPathTracer
sc = load_fixture_scene("spaceship-improved")
for run in sys.argv[1:]:
    os.environ["PPG_PERM_RUN"] = run
    for seed in (1, 2, 3, 4):
        g = GuidedPathTracer(dict(sc.integrator, budget="31", seed=str(seed))); g.set_scene(sc)
        _, st = g.render(); g.close()
        it = st["iterations"]
        print("run", run, "seed", seed, "var", ["%.4f" % it[k]["variance"] for k in (1, 2, 3)], "weight", ["%.0f" % it[k]["weight_avg"] for k in (1, 2, 3)], "leaves", [it[k]["s_tree_leaves"] for k in (2, 3)], "sub", st["sub_batches"], flush=True)
