#!/usr/bin/env python
"""GPU box: CUDA render vs oracle on the textured / bump-mapped / environment-lit CBOX and on KITCHEN (reduced resolution)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from common import load_fixture_scene, relmse
from ppg_b200.integrator import GuidedPathTracer

what = sys.argv[1] if len(sys.argv) > 1 else "all"
if what in ("all", "tex"):
    for name, size, budget in (("cbox-textured-flat", 96, "4"), ("cbox-textured-flat", 96, "60"), ("cbox-textured", 96, "4"), ("cbox-textured", 96, "60")):
        sc = load_fixture_scene(name, size)
        props = dict(sc.integrator, budget=budget)
        o = O.Oracle(O.params_from_xml(props), sc, kind="port"); oi, ost = o.render()
        g = GuidedPathTracer(props); g.set_scene(sc); gi, gst = g.render()
        bad = ~np.isclose(gi, oi, rtol=1e-3, atol=1e-4).all(axis=2)
        print(name, size, budget, "relMSE %.3e" % relmse(gi, oi), "verts o/g", ost["total_vertices"], gst["total_vertices"], "bad pixels", int(bad.sum()),
              "means", oi.mean(axis=(0, 1)), gi.mean(axis=(0, 1)), flush=True)
        o.close(); g.close()
if what in ("all", "kitchen"):
    W, H, spp = (int(a) for a in (sys.argv[2:5] if len(sys.argv) > 4 else (175, 100, 31)))
    t = time.time(); sc = load_fixture_scene("kitchen-improved").with_film(W, H); print("load", time.time() - t, flush=True)
    props = dict(sc.integrator, budget=str(spp))
    t = time.time(); g = GuidedPathTracer(props); g.set_scene(sc); print("set_scene", time.time() - t, flush=True)
    t = time.time(); gi, gst = g.render(); print("gpu render", time.time() - t, gst["total_vertices"], {k: round(v, 1) for k, v in gst["kernel_ms"].items()}, flush=True)
    t = time.time(); o = O.Oracle(O.params_from_xml(props), sc, kind="port"); oi, ost = o.render(); print("oracle render", time.time() - t, ost["total_vertices"], flush=True)
    for a, b in zip(gst["iterations"], ost["iterations"]):
        print(a["iteration"], a["passes"], "var g/o %.4f %.4f" % (a["variance"], b["variance"]), "w_avg %.1f %.1f" % (a["weight_avg"], b["weight_avg"]),
              "rec %d %d" % (a["recorded_vertices"], b["recorded_vertices"]), "leaves %d %d" % (a["s_tree_leaves"], b["s_tree_leaves"]),
              "mean_rad %.4f %.4f" % (a["mean_radiance_avg"], b["mean_radiance_avg"]), flush=True)
    print("means g/o", gi.mean(axis=(0, 1)), oi.mean(axis=(0, 1)), "finite", np.isfinite(gi).all(), "relMSE g vs o %.4f" % relmse(gi, oi))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.save(os.path.join(ROOT, "gpurun_out", f"kitchen_gpu_{W}x{H}_{spp}.npy"), gi)
if what == "kitchen_gold":
    W, H, spp = (int(a) for a in sys.argv[2:5])
    sc = load_fixture_scene("kitchen-improved").with_film(W, H)
    g = GuidedPathTracer(dict(sc.integrator, budget=str(spp))); g.set_scene(sc)
    t = time.time(); gi, gst = g.render(); dt = time.time() - t
    f = W // 175
    small = gi.astype(np.float64).reshape(100, f, 175, f, 3).mean(axis=(1, 3))
    for tag in ("reference", "improved"):
        gold = np.load(os.path.join(ROOT, "tests", "golden", f"kitchen_{tag}_175x100.npy")).astype(np.float64)
        print(tag, "means gpu/gold", small.mean(axis=(0, 1)), gold.mean(axis=(0, 1)), "relMSE %.4f" % relmse(small, gold), flush=True)
    print("render s", dt, "Msamples/s", gst["total_vertices"] / dt / 1e6, "vertices/path", gst["total_vertices"] / gst["total_paths"], {k: round(v, 1) for k, v in gst["kernel_ms"].items()})
    for a in gst["iterations"]:
        print(a["iteration"], a["passes"], "var %.4f" % a["variance"], "w_avg %.1f" % a["weight_avg"], "leaves", a["s_tree_leaves"], "sec %.2f" % a["seconds"])
