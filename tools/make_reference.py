#!/usr/bin/env python
"""GPU box: converged image of a bench workload for the equal-spp relMSE of bench.py:
    python tools/make_reference.py cbox [spp]      -> gpurun_out/cbox_1024x1024_ref.npy (float16; copy to scenes/ref/)
The product itself at a very high sample count (default 32768 spp, default parameters of the scene: the final iteration holds about half of
them).  Its remaining noise (relMSE ~ 1e-4 of a 252-spp render's) is far below the differences bench.py reports."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200"))
import numpy as np
import bench
from ppg_b200.integrator import GuidedPathTracer
name = sys.argv[1] if len(sys.argv) > 1 else "cbox"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
_, W, H, _, _ = bench.SCENES[name]
sc = bench.load_scene(name)
g = GuidedPathTracer(dict(sc.integrator, budgetType="spp", budget=str(spp), seed="987654321")); g.set_scene(sc)
t = time.time(); img, st = g.render(); print("rendered", spp, "spp in %.1f s" % (time.time() - t), "mean", img.mean(axis=(0, 1)), "final variance", st["final_variance"])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"{name}_{W}x{H}_ref.npy"), img.astype(np.float16))
