#!/usr/bin/env python
"""Regenerate the data fixtures that must travel to the GPU box (where
/root/reference does not exist).  Run in the build container:

    python tools/make_fixtures.py

Writes
  practical-path-guiding_b200/ppg_b200/data/cie1931.npz  CIE 1931 observer + D65 tables (standard colorimetric data)
  scenes/cbox.npz            flat-array form of /root/reference/scenes/cbox/cbox.xml (our loader's output)
  scenes/cbox-improved.npz   same for cbox-improved.xml
  scenes/cbox-plastic.npz    CBOX with rough-plastic boxes; carries the per-material rough-transmittance tables reduced from
                             /root/reference/mitsuba/data/microfacet/{beckmann,ggx}.dat (ppg_b200/rtrans.py)
  scenes/spaceship-improved.npz   flat-array form of /root/reference/scenes/spaceship/spaceship-improved.xml (457 560 triangles + 1 sphere)
  tests/golden/spaceship_log_stats.json   the same for spaceship-improved.exr (first six iterations)
  scenes/kitchen-improved.npz     flat-array form of /root/reference/scenes/kitchen/kitchen-improved.xml (1 414 391 triangles, 13 bitmap textures as
                                  half-precision level-0 texels, the sunsky emitter baked to a 512x256 environment map by ppg_b200/sunsky.py)
  scenes/cbox-textured.npz        procedural CBOX variant with bitmap textures, bump maps and an environment map (builtin_scenes.cbox_textured)
  scenes/cbox-textured-flat.npz   the same without the bump maps
  tests/golden/kitchen_log_stats.json     known answers of the authors' kitchen-improved.exr / kitchen.exr logs (first seven iterations)
  tests/golden/kitchen_improved_175x100.npy, kitchen_reference_175x100.npy   the golden images box-downsampled 4x4 (float16)
  tests/golden/cbox_log_stats.json   known-answer statistics parsed from the logs embedded
                                      in the reference's golden EXRs (hdrfilm attachLog)
"""
import json, os, re, struct, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200"))
from ppg_b200 import scene as S  # noqa: E402

REF = "/root/reference"


def exr_attr(fn, want):
    b = open(fn, "rb").read()
    assert b[:4] == b"\x76\x2f\x31\x01"
    p = 8
    while b[p] != 0:
        e = b.index(b"\0", p); name = b[p:e].decode(); p = e + 1
        e = b.index(b"\0", p); p = e + 1
        sz = struct.unpack("<i", b[p:p + 4])[0]; p += 4
        if name == want:
            return b[p:p + sz]
        p += sz
    return None


def parse_log(log):
    its = []
    cur = None
    triple = lambda s: [float(x) for x in re.findall(r"[-+0-9.einf]+", s.split("=")[1])]
    for l in log.splitlines():
        m = re.search(r"ITERATION (\d+), (\d+) passes", l)
        if m:
            cur = {"iteration": int(m.group(1)), "passes": int(m.group(2))}; its.append(cur); continue
        m = re.search(r"([0-9.]+) seconds, Total passes: (\d+), Var: ([-0-9.einf]+)", l)
        if m and cur is not None:
            cur.update(seconds=float(m.group(1)), total_passes=int(m.group(2)), var=float(m.group(3)))
        if cur is None:
            continue
        if "Depth  " in l: cur["depth"] = triple(l)
        if "Mean radiance" in l: cur["mean_radiance"] = triple(l)
        if "Node count" in l: cur["node_count"] = triple(l)
        if "Stat. weight" in l: cur["stat_weight"] = triple(l)
        m = re.search(r"FINAL (\d+) passes", l)
        if m: cur["final_passes"] = int(m.group(1))
    m = re.search(r"Render time: ([0-9.]+)s", log)
    m2 = re.search(r"\((\d+)x(\d+), (\d+) cores", log)
    return {"iterations": its, "render_time_s": float(m.group(1)) if m else None,
            "width": int(m2.group(1)), "height": int(m2.group(2)), "cores": int(m2.group(3))}


def plastic():
    from ppg_b200.builtin_scenes import cbox_rough_plastic
    sc = cbox_rough_plastic(S.SceneDesc.load(os.path.join(ROOT, "scenes", "cbox.npz")))
    sc.save(os.path.join(ROOT, "scenes", "cbox-plastic.npz"))
    print("cbox-plastic", sc.bsdf_names[-2:], "fdr", sc.bsdfs[-2:, 19], "ssw", sc.bsdfs[-2:, 20], "T(1)", sc.bsdf_tables[:, -1], "T(0)", sc.bsdf_tables[:, 0])


def spaceship():
    sc = S.load_mitsuba_xml(f"{REF}/scenes/spaceship/spaceship-improved.xml")
    sc.save(os.path.join(ROOT, "scenes", "spaceship-improved.npz"))
    print("spaceship", "tris", len(sc.indices), "spheres", sc.spheres, "bsdfs", sc.bsdf_names, sc.integrator)


def spaceship_log():
    """Known answers of the authors' own render of spaceship-improved.xml (640x360, log embedded in the golden EXR)."""
    log = exr_attr(f"{REF}/scenes/spaceship/spaceship-improved.exr", "log").decode(errors="replace")
    st = parse_log(log)
    st["iterations"] = st["iterations"][:6]
    with open(os.path.join(ROOT, "tests", "golden", "spaceship_log_stats.json"), "w") as f:
        json.dump({"spaceship-improved": st}, f, indent=1)
    print("spaceship log:", [(i["passes"], i.get("var"), i["stat_weight"][1]) for i in st["iterations"]])
    # the golden image itself, box-downsampled 4x4 to 160x90 RGB (float16, ~86 KB): image-level known answer for the GPU test
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    im = cv2.imread(f"{REF}/scenes/spaceship/spaceship-improved.exr", cv2.IMREAD_UNCHANGED)[..., :3][..., ::-1].astype(np.float64)
    small = im.reshape(90, 4, 160, 4, 3).mean(axis=(1, 3))
    np.save(os.path.join(ROOT, "tests", "golden", "spaceship_improved_160x90.npy"), small.astype(np.float16))
    print("spaceship golden image mean rgb", im.mean(axis=(0, 1)))


def kitchen():
    sc = S.load_mitsuba_xml(f"{REF}/scenes/kitchen/kitchen-improved.xml")
    sc.save(os.path.join(ROOT, "scenes", "kitchen-improved.npz"))
    print("kitchen", "tris", len(sc.indices), "verts", len(sc.positions), "bsdfs", len(sc.bsdfs), "textures", len(sc.textures), "texels", sc.texels.size,
          "envmap", sc.envmap["texels"].shape, sc.integrator)


def kitchen_log():
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    out = {}
    for nm in ("kitchen-improved", "kitchen"):
        st = parse_log(exr_attr(f"{REF}/scenes/kitchen/{nm}.exr", "log").decode(errors="replace"))
        st["iterations"] = st["iterations"][:7]
        out[nm] = st
    with open(os.path.join(ROOT, "tests", "golden", "kitchen_log_stats.json"), "w") as f:
        json.dump(out, f, indent=1)
    for nm, tag in (("kitchen-improved", "improved"), ("kitchen-reference", "reference")):
        im = cv2.imread(f"{REF}/scenes/kitchen/{nm}.exr", cv2.IMREAD_UNCHANGED)[..., :3][..., ::-1].astype(np.float64)
        small = im.reshape(100, 4, 175, 4, 3).mean(axis=(1, 3))
        np.save(os.path.join(ROOT, "tests", "golden", f"kitchen_{tag}_175x100.npy"), small.astype(np.float16))
        print(nm, "mean rgb", im.mean(axis=(0, 1)))


def textured():
    from ppg_b200.builtin_scenes import cbox_textured
    sc = cbox_textured(S.SceneDesc.load(os.path.join(ROOT, "scenes", "cbox.npz")))
    sc.save(os.path.join(ROOT, "scenes", "cbox-textured.npz"))
    cbox_textured(S.SceneDesc.load(os.path.join(ROOT, "scenes", "cbox.npz")), bump=False).save(os.path.join(ROOT, "scenes", "cbox-textured-flat.npz"))
    print("cbox-textured", sc.bsdf_names[-4:], "textures", len(sc.textures), "envmap", sc.envmap["texels"].shape)


def main():
    if sys.argv[1:] == ["kitchen"]:
        kitchen(); return kitchen_log()
    if sys.argv[1:] == ["textured"]:
        return textured()
    if sys.argv[1:] == ["spaceship_log"]:
        return spaceship_log()
    if sys.argv[1:] == ["plastic"]:
        return plastic()
    if sys.argv[1:] == ["spaceship"]:
        return spaceship()
    os.makedirs(os.path.join(ROOT, "scenes"), exist_ok=True)
    cie = S.extract_cie_tables()
    assert len(cie["x"]) == 471
    np.savez_compressed(S._CIE_NPZ, **cie)
    S._cie_cache = None
    for name in ("cbox", "cbox-improved"):
        sc = S.load_mitsuba_xml(f"{REF}/scenes/cbox/{name}.xml")
        sc.save(os.path.join(ROOT, "scenes", f"{name}.npz"))
        print(name, "tris", len(sc.indices), "verts", len(sc.positions), "bsdfs", sc.bsdf_names, sc.integrator)
        print("  aabb", sc.aabb_min, sc.aabb_max, "xfov", sc.x_fov_deg)
        print("  refl", sc.bsdfs[:, 2:5], "radiance", sc.area_radiance)
    stats = {}
    for name in ("cbox", "cbox-improved"):
        log = exr_attr(f"{REF}/scenes/cbox/{name}.exr", "log").decode(errors="replace")
        stats[name] = parse_log(log)
    with open(os.path.join(ROOT, "tests", "golden", "cbox_log_stats.json"), "w") as f:
        json.dump(stats, f, indent=1)
    spaceship_log()
    print(json.dumps(stats["cbox"]["iterations"][:2], indent=1))
    plastic()
    spaceship()
    textured()
    kitchen(); kitchen_log()


if __name__ == "__main__":
    main()
