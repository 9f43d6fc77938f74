"""Procedural scenes (no assets needed).

``torus_scene`` is a STAND-IN for the reference's TORUS benchmark scene, which is not bundled with the reference
("available for download on the Mitsuba website", /root/reference/README.md:63-65) and cannot be fetched here: a diffuse
torus inside a glass cube on a diffuse floor, lit by a small area light -- i.e. every light path to the torus is
specular-diffuse-specular (SDS), the caustic case path guiding was designed for.  It is NOT the original asset.
"""
from __future__ import annotations

import numpy as np

from .scene import BSDF_CONDUCTOR, BSDF_DIELECTRIC, BSDF_DIFFUSE, SceneDesc, _look_at, _make_bsdf


def _quad(p0, p1, p2, p3):
    return np.array([p0, p1, p2, p3], np.float32), np.array([[0, 1, 2], [0, 2, 3]], np.uint32)


def _box(lo, hi):
    x0, y0, z0 = lo; x1, y1, z1 = hi
    faces = [((x0, y0, z0), (x0, y1, z0), (x1, y1, z0), (x1, y0, z0)), ((x0, y0, z1), (x1, y0, z1), (x1, y1, z1), (x0, y1, z1)),
             ((x0, y0, z0), (x0, y0, z1), (x0, y1, z1), (x0, y1, z0)), ((x1, y0, z0), (x1, y1, z0), (x1, y1, z1), (x1, y0, z1)),
             ((x0, y0, z0), (x1, y0, z0), (x1, y0, z1), (x0, y0, z1)), ((x0, y1, z0), (x0, y1, z1), (x1, y1, z1), (x1, y1, z0))]
    P, I = [], []
    for f in faces:
        p, i = _quad(*f)
        I.append(i + 4 * len(P)); P.append(p)
    return np.concatenate(P), np.concatenate(I)


def _torus(R, r, nu, nv, center):
    u = np.linspace(0, 2 * np.pi, nu, endpoint=False); v = np.linspace(0, 2 * np.pi, nv, endpoint=False)
    U, V = np.meshgrid(u, v, indexing="ij")
    P = np.stack([(R + r * np.cos(V)) * np.cos(U), r * np.sin(V), (R + r * np.cos(V)) * np.sin(U)], -1).reshape(-1, 3)
    N = np.stack([np.cos(V) * np.cos(U), np.sin(V), np.cos(V) * np.sin(U)], -1).reshape(-1, 3)
    idx = lambda i, j: (i % nu) * nv + (j % nv)
    I = []
    for i in range(nu):
        for j in range(nv):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            I += [(a, c, b), (a, d, c)]          # outward-facing winding
    return (P + np.array(center)).astype(np.float32), N.astype(np.float32), np.array(I, np.uint32)


def torus_scene(size=256, nu=48, nv=24) -> SceneDesc:
    meshes = []   # (P, N or None, I, bsdf, emitter)
    bsdfs = [_make_bsdf(BSDF_DIFFUSE, 0, (0.7, 0.7, 0.7)), _make_bsdf(BSDF_DIFFUSE, 0, (0.8, 0.25, 0.15)),
             _make_bsdf(BSDF_DIELECTRIC, 0, (1, 1, 1), (1, 1, 1), (1.5046 / 1.000277,) * 3), _make_bsdf(BSDF_DIFFUSE, 0, (0, 0, 0))]
    names = ["floor", "torus", "glass", "__black"]
    P, I = _quad((-6, -1.001, -6), (-6, -1.001, 6), (6, -1.001, 6), (6, -1.001, -6)); meshes.append((P, None, I, 0, -1))
    P, N, I = _torus(0.6, 0.22, nu, nv, (0, -0.3, 0)); meshes.append((P, N, I, 1, -1))
    P, I = _box((-1, -1, -1), (1, 1, 1)); meshes.append((P, None, I, 2, -1))
    P, I = _quad((-0.15, 3.0, -0.15), (0.15, 3.0, -0.15), (0.15, 3.0, 0.15), (-0.15, 3.0, 0.15)); meshes.append((P, None, I, 3, 0))   # faces down (-y)
    Ps, Ns, Is, TS, shapes = [], [], [], [], []
    voff = toff = 0
    for k, (P, N, I, b, e) in enumerate(meshes):
        shapes.append([toff, len(I), b, e, 0 if N is None else 1, 0, 0, 0])
        Ps.append(P); Ns.append(N if N is not None else np.zeros_like(P)); Is.append(I + voff); TS.append(np.full(len(I), k, np.uint32))
        voff += len(P); toff += len(I)
    P = np.concatenate(Ps).astype(np.float32)
    cam = _look_at((2.6, 1.9, 3.4), (0, -0.2, 0), (0, 1, 0))
    mn = np.minimum(P.min(0), cam[:3, 3]); mx = np.maximum(P.max(0), cam[:3, 3])
    integ = {"strictNormals": "true", "maxDepth": "12", "rrDepth": "12", "budgetType": "spp", "budget": "127", "sTreeThreshold": "4000", "sppPerPass": "1"}
    return SceneDesc(positions=P, normals=np.concatenate(Ns).astype(np.float32), uvs=np.zeros((len(P), 2), np.float32),
                     indices=np.concatenate(Is).astype(np.uint32), triangle_shape=np.concatenate(TS).astype(np.uint32),
                     shapes=np.asarray(shapes, np.int32), bsdfs=np.asarray(bsdfs, np.float32), area_radiance=np.array([[400.0, 400.0, 380.0]], np.float32),
                     cam_to_world=cam.astype(np.float32), x_fov_deg=35.0, near_clip=0.1, far_clip=100.0, film_width=size, film_height=size,
                     aabb_min=mn.astype(np.float32), aabb_max=mx.astype(np.float32), integrator=integ, bsdf_names=names)


def cbox_glass_mirror(cbox: SceneDesc) -> SceneDesc:
    """CBOX with a glass small box (dielectric, bk7 in air) and a mirror large box (conductor, material=none)."""
    import copy
    sc = copy.copy(cbox)
    nb = len(sc.bsdfs)
    sc.bsdfs = np.concatenate([_pad_bsdfs(sc.bsdfs), np.stack([_make_bsdf(BSDF_DIELECTRIC, 0, (1, 1, 1), (1, 1, 1), (1.5046 / 1.000277,) * 3),
                                                   _make_bsdf(BSDF_CONDUCTOR, 0, (0.95, 0.95, 0.95), (0, 0, 0), (0, 0, 0), (1, 1, 1))])]).astype(np.float32)
    sc.bsdf_names = list(sc.bsdf_names) + ["glass", "mirror"]
    shapes = sc.shapes.copy()
    shapes[6, 2] = nb          # cbox_smallbox
    shapes[7, 2] = nb + 1      # cbox_largebox
    sc.shapes = shapes
    return sc


def cbox_rough_glass(cbox: SceneDesc) -> SceneDesc:
    """CBOX whose small box is frosted glass (roughdielectric GGX alpha 0.15, eta 1.5) and whose large box is a lightly rough
    glass (Beckmann alpha 0.05, tinted transmittance): glossy transmission is ESmooth, i.e. guided, recorded and light-sampled."""
    import copy
    from .scene import BSDF_ROUGHDIELECTRIC
    sc = copy.copy(cbox)
    nb = len(sc.bsdfs)
    sc.bsdfs = np.concatenate([_pad_bsdfs(sc.bsdfs), np.stack([
        _make_bsdf(BSDF_ROUGHDIELECTRIC, 0, (1, 1, 1), (1, 1, 1), (1.5,) * 3, (0, 0, 0), 0.15, 1),
        _make_bsdf(BSDF_ROUGHDIELECTRIC, 0, (1, 1, 1), (0.8, 0.95, 0.9), (1.33,) * 3, (0, 0, 0), 0.05, 0)])]).astype(np.float32)
    sc.bsdf_names = list(sc.bsdf_names) + ["frosted_glass_ggx", "rough_water_beckmann"]
    shapes = sc.shapes.copy(); shapes[6, 2] = nb; shapes[7, 2] = nb + 1
    sc.shapes = shapes
    return sc


def cbox_with_analytic_spheres(cbox: SceneDesc) -> SceneDesc:
    """CBOX plus three analytic spheres (src/shapes/sphere.cpp): a white diffuse ball on the floor, a small emitting ball (light
    sampling from outside: uniform cone) and a huge inward-facing emitting shell around everything like spaceship.xml's
    (light sampling from inside: uniform sphere).  The scene box grows to the shell (Sphere::getAABB)."""
    import copy
    from .scene import make_sphere, BSDF_DIFFUSE
    sc = copy.copy(cbox)
    ns, ne = len(sc.shapes), len(sc.area_radiance)
    base = _pad_bsdfs(sc.bsdfs); nb = len(base)
    sc.bsdfs = np.concatenate([base, _make_bsdf(BSDF_DIFFUSE, 0, (0, 0, 0))[None]]).astype(np.float32)        # emitters without a BSDF are black (shape.cpp)
    sc.bsdf_names = list(sc.bsdf_names) + ["__black"]
    nt = len(sc.indices)
    sc.shapes = np.concatenate([sc.shapes, np.array([[nt, 0, 1, -1, 0, 0, 0, 0], [nt, 0, nb, ne, 0, 0, 0, 0], [nt, 0, nb, ne + 1, 0, 0, 0, 0]], np.int32)])
    sc.area_radiance = np.concatenate([sc.area_radiance, np.array([[12.0, 10.0, 6.0], [0.08, 0.1, 0.14]], np.float32)])
    sc.spheres = np.stack([make_sphere((415.0, 60.0, 135.0), 60.0, ns), make_sphere((150.0, 400.0, 250.0), 25.0, ns + 1),
                           make_sphere((278.0, 273.0, -100.0), 1500.0, ns + 2, True)])
    sc.aabb_min = np.minimum(sc.aabb_min, np.float32([278, 273, -100]) - np.float32(1500)).astype(np.float32)
    sc.aabb_max = np.maximum(sc.aabb_max, np.float32([278, 273, -100]) + np.float32(1500)).astype(np.float32)
    return sc


def cbox_smooth_plastic(cbox: SceneDesc) -> SceneDesc:
    """CBOX whose boxes are smooth plastics (plastic.cpp): a delta coat reflection mixed with a diffuse base, so a guided vertex can
    return a delta sample (GP:1670-1676).  Small box nonlinear red, large box linear blue; the floor becomes a glossy dark plastic."""
    import copy
    from .scene import make_plastic
    sc = copy.copy(cbox)
    base = _pad_bsdfs(sc.bsdfs); nb = len(base)
    sc.bsdfs = np.concatenate([base, np.stack([make_plastic(0, (0.6, 0.1, 0.08), (1, 1, 1), 1.49 / 1.000277, True),
                                               make_plastic(0, (0.1, 0.25, 0.6), (0.9, 0.9, 0.9), 1.9, False),
                                               make_plastic(0, (0.2, 0.2, 0.2), (1, 1, 1), 1.5, False)])]).astype(np.float32)
    sc.bsdf_names = list(sc.bsdf_names) + ["red_plastic_nonlinear", "blue_plastic", "dark_floor_plastic"]
    shapes = sc.shapes.copy(); shapes[6, 2] = nb; shapes[7, 2] = nb + 1; shapes[1, 2] = nb + 2
    sc.shapes = shapes
    return sc


def cbox_thin_glass(cbox: SceneDesc) -> SceneDesc:
    """CBOX with two thin-dielectric panes (thindielectric.cpp): a horizontal one under the ceiling light, so that ALL direct light
    reaches the room through an index-matched (ENull) surface -- light sampling multiplies the pane's transmittance
    (Scene::evalTransmittance) and the emitter lookup after BSDF sampling looks through it (GP:2184-2245) -- and a tilted, tinted
    one in front of the boxes that the camera looks through (null transition on the primary ray: `scattered` stays false)."""
    import copy
    from .scene import BSDF_THINDIELECTRIC
    sc = copy.copy(cbox)
    base = _pad_bsdfs(sc.bsdfs); nb = len(base)
    sc.bsdfs = np.concatenate([base, np.stack([_make_bsdf(BSDF_THINDIELECTRIC, 0, (1, 1, 1), (1, 1, 1), (1.5046 / 1.000277,) * 3),
                                               _make_bsdf(BSDF_THINDIELECTRIC, 0, (1, 1, 1), (0.7, 0.9, 0.95), (1.33,) * 3)])]).astype(np.float32)
    sc.bsdf_names = list(sc.bsdf_names) + ["thin_glass", "thin_tinted"]
    quads = [np.float32([[0.5, 500, 0.5], [555.5, 500, 0.5], [555.5, 500, 558.7], [0.5, 500, 558.7]]),
             np.float32([[60, 20, 40], [500, 20, 60], [500, 430, 150], [60, 430, 130]])]
    nv, nt, ns = len(sc.positions), len(sc.indices), len(sc.shapes)
    P = np.concatenate(quads)
    sc.positions = np.concatenate([sc.positions, P]); sc.normals = np.concatenate([sc.normals, np.zeros_like(P)])
    sc.uvs = np.concatenate([sc.uvs, np.zeros((len(P), 2), np.float32)])
    tris = np.uint32([[0, 1, 2], [0, 2, 3], [4, 5, 6], [4, 6, 7]]) + nv
    sc.indices = np.concatenate([sc.indices, tris]).astype(np.uint32)
    sc.triangle_shape = np.concatenate([sc.triangle_shape, np.uint32([ns, ns, ns + 1, ns + 1])])
    sc.shapes = np.concatenate([sc.shapes, np.array([[nt, 2, nb, -1, 0, 0, 0, 0], [nt + 2, 2, nb + 1, -1, 0, 0, 0, 0]], np.int32)])
    return sc


def cbox_blinds(cbox: SceneDesc) -> SceneDesc:
    """CBOX with two `mask` panes like kitchen.xml's "Blinds" (mask.cpp: constant opacity around a twosided diffuse BSDF), in the places
    of cbox_thin_glass's panes.  A mask is a smooth/null hybrid: guided, light-sampled, looked through by the emitter lookup and
    shadow rays, and with a sampling-fraction loss its null transitions are recorded as delta vertices (GP:2049-2066)."""
    from .scene import BSDF_FLAG_MASK, BSDF_FLAG_TWOSIDED, BSDF_DIFFUSE
    sc = cbox_thin_glass(cbox)
    b = sc.bsdfs.copy()
    blinds = _make_bsdf(BSDF_DIFFUSE, BSDF_FLAG_MASK | BSDF_FLAG_TWOSIDED, (0.612066, 0.499505, 0.378676)); blinds[22:25] = (0.612066,) * 3
    veil = _make_bsdf(BSDF_DIFFUSE, BSDF_FLAG_MASK | BSDF_FLAG_TWOSIDED, (0.2, 0.5, 0.7)); veil[22:25] = (0.3, 0.5, 0.8)
    b[-2] = blinds; b[-1] = veil
    sc.bsdfs = b
    sc.bsdf_names = list(sc.bsdf_names[:-2]) + ["blinds", "veil"]
    return sc


def cbox_rough_metal(cbox: SceneDesc) -> SceneDesc:
    """CBOX whose boxes are rough conductors: the small box GGX alpha 0.1 with the eta/k of spaceship.xml's "RoughAluminium",
    the large box Beckmann alpha 0.3 -- glossy BSDFs are guided (ESmooth) and take part in light sampling."""
    import copy
    from .scene import BSDF_ROUGHCONDUCTOR
    sc = copy.copy(cbox)
    nb = len(sc.bsdfs)
    sc.bsdfs = np.concatenate([_pad_bsdfs(sc.bsdfs), np.stack([
        _make_bsdf(BSDF_ROUGHCONDUCTOR, 0, (0.578596,) * 3, (0, 0, 0), (1.65746, 0.880369, 0.521229), (9.22387, 6.26952, 4.837), 0.1, 1),
        _make_bsdf(BSDF_ROUGHCONDUCTOR, 0, (1, 1, 1), (0, 0, 0), (0.2, 0.92, 1.1), (3.9, 2.45, 2.14), 0.3, 0)])]).astype(np.float32)
    sc.bsdf_names = list(sc.bsdf_names) + ["rough_aluminium_ggx", "rough_gold_beckmann"]
    shapes = sc.shapes.copy(); shapes[6, 2] = nb; shapes[7, 2] = nb + 1
    sc.shapes = shapes
    return sc


def _pad_bsdfs(b):
    b = np.asarray(b, np.float32)
    return b if b.shape[1] >= 28 else np.concatenate([b, np.zeros((len(b), 28 - b.shape[1]), np.float32)], axis=1)


def cbox_rough_plastic(cbox: SceneDesc) -> SceneDesc:
    """CBOX whose boxes are rough plastics (needs the reference's data/microfacet tables at build time):
    the small box like spaceship.xml's "PinkLeather" (Beckmann 0.4, nonlinear), the large box GGX alpha 0.2, linear."""
    import copy
    from .scene import make_roughplastic
    sc = copy.copy(cbox)
    tables = []
    extra = [make_roughplastic(0, (0.256, 0.013, 0.08), (1, 1, 1), 1.5 / 1.000277, 0.4, 0, True, tables),
             make_roughplastic(0, (0.1, 0.3, 0.6), (1, 1, 1), 1.5 / 1.000277, 0.2, 1, False, tables)]
    base = _pad_bsdfs(sc.bsdfs)
    nb = len(base)
    sc.bsdfs = np.concatenate([base, np.stack(extra)]).astype(np.float32)
    sc.bsdf_tables = np.asarray(tables, np.float32)
    sc.bsdf_names = list(sc.bsdf_names) + ["pink_leather_beckmann", "blue_plastic_ggx"]
    shapes = sc.shapes.copy(); shapes[6, 2] = nb; shapes[7, 2] = nb + 1
    sc.shapes = shapes
    return sc


def _half_bits(img):
    with np.errstate(over="ignore"):
        return np.ascontiguousarray(img, np.float32).astype(np.float16).view(np.uint16).reshape(-1)


def cbox_textured(cbox: SceneDesc, bump=True, env=True, size=64) -> SceneDesc:
    """CBOX with procedural bitmap textures (no assets): a checker/gradient RGB texture on the floor, ceiling and back wall
    (meshes without texture coordinates: uv = barycentrics, skdtree.h:404) and on the tall box (given explicit, scaled and
    offset UVs with `repeat` wrapping), a one-channel bump map on the two boxes (tall box: twosided diffuse with a textured
    reflectance; short box: rough plastic with a textured diffuseReflectance), and a small lat-long environment map that
    lights the scene through the open front.  Exercises every texture / bumpmap / envmap code path of the hot loop."""
    import copy
    from .scene import (BSDF_FLAG_BUMPMAP, BSDF_FLAG_TWOSIDED, TEXTURE_DTYPE, make_roughplastic)
    sc = copy.copy(cbox)
    rng = np.random.default_rng(7)
    # texture 0: RGB checker with a smooth gradient and a little noise (size x size)
    y, x = np.mgrid[0:size, 0:size].astype(np.float64) / size
    chk = ((np.floor(x * 8) + np.floor(y * 8)) % 2)
    rgb = np.stack([0.15 + 0.7 * chk * x, 0.2 + 0.6 * (1 - chk) * y, 0.25 + 0.5 * (0.5 + 0.5 * np.sin(12 * x) * np.cos(9 * y))], -1)
    rgb = np.clip(rgb + rng.uniform(-0.03, 0.03, rgb.shape), 0.0, 0.95)
    # texture 1: luminance bump (smooth bumps), texture 2: RGB stripes with non-square size and clamp / mirror wrapping
    hgt = 0.5 + 0.5 * np.sin(2 * np.pi * 6 * x) * np.sin(2 * np.pi * 5 * y)
    w2, h2 = 48, 20
    yy, xx = np.mgrid[0:h2, 0:w2].astype(np.float64)
    stripes = np.stack([0.2 + 0.6 * ((xx // 4) % 2), 0.3 + 0.5 * (yy / h2), 0.6 - 0.4 * (xx / w2)], -1)
    texs = [(rgb, 3, 0, 0, (1.0, 1.0), (0.0, 0.0)), (hgt[..., None] * 4.0, 1, 0, 0, (3.0, 2.0), (0.25, 0.5)), (stripes, 3, 1, 2, (2.5, 1.5), (-0.2, 0.1))]
    meta = np.zeros(len(texs), TEXTURE_DTYPE); texels = []; off = 0
    for i, (img, ch, wu, wv, scl, ofs) in enumerate(texs):
        meta[i] = (img.shape[1], img.shape[0], ch, wu, wv, scl, ofs, 0, off)
        hb = _half_bits(img); texels.append(hb); off += len(hb)
    sc.textures = meta; sc.texels = np.concatenate(texels)
    # materials
    nb = len(sc.bsdfs); B = _pad_bsdfs(sc.bsdfs)
    def u32(v): return np.array([v], np.uint32).view(np.float32)
    floor = _make_bsdf(BSDF_DIFFUSE, 0, rgb.reshape(-1, 3).mean(0)); floor[25:26] = u32(1)
    wall = _make_bsdf(BSDF_DIFFUSE, BSDF_FLAG_TWOSIDED, stripes.reshape(-1, 3).mean(0)); wall[25:26] = u32(3)
    tall = _make_bsdf(BSDF_DIFFUSE, BSDF_FLAG_TWOSIDED | (BSDF_FLAG_BUMPMAP if bump else 0), rgb.reshape(-1, 3).mean(0)); tall[25:26] = u32(1); tall[26:27] = u32(2 if bump else 0)
    tables = [t for t in (sc.bsdf_tables if sc.bsdf_tables is not None else [])]
    short = make_roughplastic(BSDF_FLAG_TWOSIDED | (BSDF_FLAG_BUMPMAP if bump else 0), stripes.reshape(-1, 3).mean(0), np.ones(3, np.float32), 1.49 / 1.000277, 0.15, 1, False, tables)
    short[25:26] = u32(3); short[26:27] = u32(2 if bump else 0)
    sc.bsdf_tables = np.asarray(tables, np.float32).reshape(-1, 100)
    sc.bsdfs = np.concatenate([B, np.stack([floor, wall, tall, short])]).astype(np.float32)
    sc.bsdf_names = list(sc.bsdf_names) + ["tex_floor", "tex_wall", "tex_bump_tall", "tex_bump_short"]
    shapes = sc.shapes.copy()
    shapes[1, 2] = nb          # floor: textured diffuse, uv = barycentrics
    shapes[3, 2] = nb + 1      # back wall: stripes (clamp / mirror)
    shapes[7, 2] = nb + 2      # tall box: explicit UVs below
    shapes[6, 2] = nb + 3      # short box: rough plastic, barycentric uv
    # explicit texture coordinates for the tall box: planar projection of its vertices, beyond [0,1] to exercise wrapping
    first, n = int(shapes[7, 0]), int(shapes[7, 1])
    vid = np.unique(sc.indices[first:first + n].reshape(-1))
    uv = sc.uvs.copy()
    P = sc.positions[vid]
    uv[vid, 0] = (P[:, 0] + 0.37 * P[:, 2]) / 150.0
    uv[vid, 1] = (P[:, 1] + 0.21 * P[:, 2]) / 120.0 - 0.3
    sc.uvs = uv.astype(np.float32)
    shapes[7, 5] = 1
    sc.shapes = shapes
    if env:
        eh, ew = 16, 32
        t, p = np.mgrid[0:eh, 0:ew].astype(np.float64)
        sky = np.stack([0.3 + 0.2 * np.cos(p / ew * 2 * np.pi), 0.4 + 0.0 * t, 0.6 - 0.3 * t / eh], -1) * (t[..., None] < eh / 2 + 2)
        sky[3, 9] += np.array([60.0, 50.0, 40.0])                 # a small bright "sun" texel
        ang = np.radians(35.0)
        rot = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
        sc.envmap = {"texels": _half_bits(sky).reshape(eh, ew, 3), "scale": 1.5, "world_to_env": np.linalg.inv(rot).astype(np.float32)}
    return sc


def env_lit_scene(size=64, torus=False) -> SceneDesc:
    """A diffuse ground plate with a diffuse box and a mirror-like rough-conductor slab on it, lit ONLY by a lat-long environment
    map (dim sky + one bright "sun" texel, rotated about y): no area emitter at all, so every light sample of `nee = always | kickstart`
    is an environment sample (EnvironmentMap::sampleDirect) and every emitter hit of a BSDF / guided sample is a ray that leaves the scene.
    `torus` adds a smooth-shaded diffuse torus of 2304 triangles: the scene then no longer fits the shared-memory staging and is intersected
    through the BVH (and, from 32 768 paths per wavefront, by the separate nearest-hit pass)."""
    from .scene import BSDF_ROUGHCONDUCTOR
    meshes = []
    bsdfs = [_make_bsdf(BSDF_DIFFUSE, 0, (0.6, 0.6, 0.55)), _make_bsdf(BSDF_DIFFUSE, 0, (0.7, 0.3, 0.2)),
             _make_bsdf(BSDF_ROUGHCONDUCTOR, 0, (0.9, 0.9, 0.9), (0, 0, 0), (0.2, 0.9, 1.1), (3.9, 2.4, 2.2), 0.2, 1)]
    names = ["ground", "box", "slab"]
    P, I = _quad((-3, 0, -3), (-3, 0, 3), (3, 0, 3), (3, 0, -3)); meshes.append((P, I, 0))
    P, I = _box((-0.6, 0.0, -0.5), (0.3, 0.9, 0.4)); meshes.append((P, I, 1))
    P, I = _quad((0.7, 0.0, -1.0), (0.7, 1.2, -1.0), (1.5, 1.2, 0.2), (1.5, 0.0, 0.2)); meshes.append((P, I, 2))
    normals = [np.zeros((len(m[0]), 3), np.float32) for m in meshes]
    if torus:
        P, N, I = _torus(0.55, 0.2, 48, 24, (-1.4, 0.2, 1.2)); meshes.append((P, I, 1)); normals.append(N)
    Ps, Is, TS, shapes = [], [], [], []
    voff = toff = 0
    for k, (P, I, b) in enumerate(meshes):
        shapes.append([toff, len(I), b, -1, 1 if np.any(normals[k]) else 0, 0, 0, 0])
        Ps.append(P); Is.append(I + voff); TS.append(np.full(len(I), k, np.uint32))
        voff += len(P); toff += len(I)
    P = np.concatenate(Ps).astype(np.float32)
    cam = _look_at((2.8, 2.4, 3.6), (0.2, 0.4, 0), (0, 1, 0))
    mn = np.minimum(P.min(0), cam[:3, 3]); mx = np.maximum(P.max(0), cam[:3, 3])
    eh, ew = 16, 32
    t, p = np.mgrid[0:eh, 0:ew].astype(np.float64)
    sky = np.stack([0.25 + 0.15 * np.cos(p / ew * 2 * np.pi), 0.35 + 0.0 * t, 0.55 - 0.3 * t / eh], -1) * (t[..., None] < eh / 2 + 1)
    sky[2, 20] += np.array([90.0, 80.0, 60.0])                     # the "sun"
    sky[5, 3] += np.array([4.0, 6.0, 9.0])                         # a second, dimmer lobe
    ang = np.radians(-20.0)
    rot = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    integ = {"strictNormals": "true", "maxDepth": "6", "rrDepth": "5", "budgetType": "spp", "budget": "28", "sppPerPass": "4"}
    return SceneDesc(positions=P, normals=np.concatenate(normals).astype(np.float32), uvs=np.zeros((len(P), 2), np.float32),
                     indices=np.concatenate(Is).astype(np.uint32), triangle_shape=np.concatenate(TS).astype(np.uint32),
                     shapes=np.asarray(shapes, np.int32), bsdfs=np.asarray(bsdfs, np.float32), area_radiance=np.zeros((0, 3), np.float32),
                     cam_to_world=cam.astype(np.float32), x_fov_deg=40.0, near_clip=0.1, far_clip=100.0, film_width=size, film_height=size,
                     aabb_min=mn.astype(np.float32), aabb_max=mx.astype(np.float32), integrator=integ, bsdf_names=names,
                     envmap={"texels": _half_bits(sky).reshape(eh, ew, 3), "scale": 1.2, "world_to_env": np.linalg.inv(rot).astype(np.float32)})
