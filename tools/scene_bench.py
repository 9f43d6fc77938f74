#!/usr/bin/env python
"""GPU box: throughput of the CUDA path on the non-diffuse CBOX variants (and any .npz scene), one line per (library, scene).
usage: scene_bench.py [--size 512] [--budget 124] [--libs default,build_variants/x.so] [--scenes plastic,metal,glass,diffuse]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512); ap.add_argument("--budget", default="124")
ap.add_argument("--libs", default="default"); ap.add_argument("--scenes", default="plastic,metal,glass")
ap.add_argument("--extra", default="", help="k=v,k=v integrator properties")
a = ap.parse_args()
from common import load_cbox, load_fixture_scene


def SceneDesc_load(name):
    from ppg_b200.scene import SceneDesc
    return SceneDesc.load(name)


def scene(name):
    from ppg_b200 import builtin_scenes as B
    if name == "diffuse": return load_cbox(a.size or None)
    if name == "plastic": return load_fixture_scene("cbox-plastic", a.size or 512)
    if name == "metal": return B.cbox_rough_metal(load_cbox(a.size or 512))
    if name == "glass": return B.cbox_rough_glass(load_cbox(a.size or 512))
    if name == "mirror": return B.cbox_glass_mirror(load_cbox(a.size or 512))
    if name == "blinds": return B.cbox_blinds(load_cbox(a.size or 512))
    if name.endswith(".npz") and a.size: return SceneDesc_load(name).with_film(a.size, a.size)
    from ppg_b200.scene import SceneDesc
    return SceneDesc.load(name)


for lib in a.libs.split(","):
    if lib != "default": os.environ["PPG_B200_LIB"] = os.path.join(ROOT, lib)
    else: os.environ.pop("PPG_B200_LIB", None)
    import importlib
    from ppg_b200 import capi
    importlib.reload(capi)
    from ppg_b200 import integrator as I
    importlib.reload(I)
    for sn in a.scenes.split(","):
        sc = scene(sn)
        props = dict(sc.integrator, budget=a.budget)
        for kv in filter(None, a.extra.split(",")):
            k, v = kv.split("="); props[k] = v
        best = None
        for rep in range(2):
            g = I.GuidedPathTracer(props); g.set_scene(sc); ptr, st = g.render_device(); g.close()
            ms = st["render_device_ms"]; v = st["total_vertices"] / ms / 1e3
            best = max(best or 0, v)
        km = st.get("kernel_ms", [])
        kb = km.get("bounce", 0.0) if isinstance(km, dict) else 0.0
        print(f"{lib:40s} {sn:10s} {best:9.1f} Msamples/s  ms {ms:8.1f}  bounce_ms {kb:8.1f} ({st['total_vertices'] / max(kb, 1e-9) / 1e3:7.1f} Msamples/s of bounce-kernel time) "
              f"commit {km.get('commit', 0):.1f} adam {km.get('adam', 0):.1f}", flush=True)
