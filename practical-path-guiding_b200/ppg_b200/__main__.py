"""python -m ppg_b200 scene.xml [-o out.exr|.pfm|.npy] [-D name=value ...] [--size W H] [--device N] [--sdt prefix] [--check]

The reference's `mitsuba [-D name=value] [-o out] scene.xml` for the guided_path integrator: reads the Mitsuba 0.5 scene XML (or a .npz fixture) unchanged,
applies -D overrides to the <integrator> block (XML names and value strings: budget, budgetType, nee, sampleCombination, spatialFilter, ...), renders on the
B200 through libppg_b200.so and writes the film (OpenEXR through OpenCV, PFM, or .npy).  The per-iteration block of the reference's log (GP:1176-1186,
1323-1326) goes to stderr.  --sdt PREFIX sets dumpSDTree=true and the destination prefix (PREFIX-00.sdt, ... readable by the reference's visualizer);
--check stops after loading the scene and validating the parameters (no CUDA device needed).  There is no CPU fallback."""
import argparse
import os
import sys

import numpy as np


def write_image(path, img):
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npy":
        np.save(path, img)
    elif ext == ".pfm":
        with open(path, "wb") as f:
            f.write(b"PF\n%d %d\n-1.0\n" % (img.shape[1], img.shape[0])); f.write(np.ascontiguousarray(img[::-1], "<f4").tobytes())
    elif ext == ".exr":
        os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
        import cv2
        if not cv2.imwrite(path, np.ascontiguousarray(img[..., ::-1], np.float32)):
            raise OSError(f"cannot write {path}")
    else:
        raise ValueError(f"output format '{ext}' (use .exr, .pfm or .npy)")


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m ppg_b200", description=__doc__.split("\n\n")[1], formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("scene"); ap.add_argument("-o", "--output"); ap.add_argument("-D", action="append", default=[], metavar="name=value")
    ap.add_argument("--size", nargs=2, type=int, metavar=("W", "H")); ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--sdt", metavar="PREFIX"); ap.add_argument("--check", action="store_true")
    a = ap.parse_args(argv)
    from . import capi
    from .scene import SceneDesc, load_mitsuba_xml
    sc = SceneDesc.load(a.scene) if a.scene.endswith(".npz") else load_mitsuba_xml(a.scene)
    if a.size:
        sc = sc.with_film(*a.size)
    props = dict(sc.integrator)
    for kv in a.D:
        if "=" not in kv:
            ap.error(f"-D expects name=value, got '{kv}'")
        k, v = kv.split("=", 1); props[k] = v
    if a.sdt:
        props["dumpSDTree"] = "true"
    lib = capi.load_library()
    prm = capi.PpgParams(); lib.ppg_params_default(prm)
    for k, v in props.items():                                   # the plugin constructor's validation and messages (GP:1014-1085)
        if lib.ppg_params_set(prm, k.encode(), str(v).encode()) != 0:
            print(f"ppg_b200: {lib.ppg_last_error().decode()}", file=sys.stderr); return 1
    if lib.ppg_params_validate(prm) != 0:
        print(f"ppg_b200: {lib.ppg_last_error().decode()}", file=sys.stderr); return 1
    print(f"scene: {len(sc.indices)} triangles, {len(sc.shapes)} shapes, {len(sc.bsdfs)} materials, {len(sc.area_radiance)} area emitters"
          f"{' + environment map' if sc.envmap else ''}, film {sc.film_width} x {sc.film_height}", file=sys.stderr)
    if a.check:
        print("check ok", file=sys.stderr); return 0
    out = a.output or os.path.splitext(a.scene)[0] + ".exr"
    from .integrator import GuidedPathTracer, PpgError
    try:
        g = GuidedPathTracer(props, device=a.device)
    except PpgError as e:
        print(f"ppg_b200: {e}", file=sys.stderr); return 2
    g.set_scene(sc)
    if a.sdt:
        g.set_destination(a.sdt)
    img, st = g.render()
    for it in st["iterations"]:
        print(f"ITERATION {it['iteration']}{' (FINAL)' if it['is_final'] else ''}, {it['passes']} passes, {it['seconds']:.2f} s, Var: {it['variance']:g} | D-tree depth "
              f"{it['depth_min']}..{it['depth_max']} (avg {it['depth_avg']:.2f}), nodes {it['nodes_min']}..{it['nodes_max']} (avg {it['nodes_avg']:.1f}), "
              f"stat. weight avg {it['weight_avg']:.1f}, {it['s_tree_leaves']} S-tree leaves", file=sys.stderr)
    ms = st["render_device_ms"]
    print(f"{st['total_paths']} paths, {st['total_vertices']} vertices in {st['render_seconds']:.3f} s ({st['total_vertices'] / ms * 1e-3 if ms > 0 else 0:.1f} Msamples/s on the device)", file=sys.stderr)
    write_image(out, img)
    return 0 if g.last_status == 0 else 5


if __name__ == "__main__":
    sys.exit(main())
