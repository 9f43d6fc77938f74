"""Minimal Mitsuba-0.5 scene loader -> flat arrays for ``ppg_set_scene``.

Restates only what the bundled scenes of the reference need (SURVEY.md §2 #11,
§7 step 2): it is a *host loader for the hot path*, not a scene graph.

Reference behaviour followed (paths under /root/reference/mitsuba):
  * <spectrum value="l:v,..."> -> zero-extended piecewise-linear spectrum,
    convolved with the CIE 1931 matching functions, -> XYZ -> linear Rec.709
    (src/librender/scenehandler.cpp:597-613, src/libcore/spectrum.cpp:172-191,
    222-227, 630-686).  Single-valued emitter spectra are multiplied by D65
    (scenehandler.cpp:575-593); <rgb> is taken verbatim.
  * transforms compose left-to-right as ``T_new * T_so_far``
    (scenehandler.cpp:348-441); lookAt builds (left, up, dir, origin) columns
    (src/libcore/transform.cpp:191-214).
  * OBJ: n-gons are fanned (src/shapes/obj.cpp:313-324), vertices merged per
    (p, n, uv) key, missing normals replaced by angle-weighted smooth normals
    (src/librender/trimesh.cpp:608-690) unless ``faceNormals``.
  * other mesh sources: ``ply`` (src/shapes/ply.cpp: ascii / binary, triangles and quads), ``serialized``
    (TriMesh::loadCompressed, src/librender/trimesh.cpp:176-335: zlib stream, versions 3 / 4, ``shapeIndex``),
    ``cube`` (src/shapes/cube.cpp) and ``rectangle`` as two triangles; ``sphere`` stays analytic.
  * perspective sensor: fovAxis resolution (src/librender/sensor.cpp:239-300).
  * scene AABB = geometry AABB + sensor position + emitter AABBs
    (src/librender/scene.cpp:387-413).
  * shapes with an emitter and no BSDF get a black diffuse BSDF, shapes with
    neither get 0.5 grey diffuse (src/librender/shape.cpp:48-72).
  * bitmap textures: 8-bit files are converted from sRGB (or the explicit
    ``gamma``) to linear floats and stored in half precision, level 0 only
    (src/textures/bitmap.cpp:178-183, 301-330; include/mitsuba/render/mipmap.h:226-231);
    ``bumpmap`` wraps its nested BSDF with the displacement texture
    (src/bsdfs/bumpmap.cpp:119-137); nested BSDFs that carry an ``id`` are
    registered on their own like any named object (scenehandler.cpp), so a
    ``<ref>`` to the inner id of kitchen.xml's bump-mapped cushions gets the
    un-bumped material, exactly as in the reference.
  * ``sunsky`` (and its halves ``sky`` / ``sun``) is baked to a lat-long environment map (ppg_b200/sunsky.py);
    environment emitters do not change the scene AABB (envmap.cpp:558-564).
"""
from __future__ import annotations

import math
import os
import re
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
_CIE_NPZ = os.path.join(_DATA_DIR, "cie1931.npz")
_REF_SPECTRUM_CPP = "/root/reference/mitsuba/src/libcore/spectrum.cpp"

BSDF_DIFFUSE = 0
BSDF_NULL_BLACK = 1
BSDF_DIELECTRIC = 2
BSDF_CONDUCTOR = 3
BSDF_ROUGHCONDUCTOR = 4
BSDF_ROUGHPLASTIC = 5
BSDF_ROUGHDIELECTRIC = 6
BSDF_PLASTIC = 7
BSDF_THINDIELECTRIC = 8
BSDF_FLAG_NONLINEAR = 2
BSDF_FLAG_MASK = 4
BSDF_FLAG_TWOSIDED = 1
BSDF_FLAG_BUMPMAP = 8
WRAP_MODES = {"repeat": 0, "clamp": 1, "mirror": 2}
# numpy mirror of ppg_texture (include/ppg.h), 48 bytes
TEXTURE_DTYPE = np.dtype([("width", "<u4"), ("height", "<u4"), ("channels", "<u4"), ("wrap_u", "<u4"), ("wrap_v", "<u4"),
                          ("uv_scale", "<f4", 2), ("uv_offset", "<f4", 2), ("reserved", "<u4"), ("first_texel", "<u8")])
assert TEXTURE_DTYPE.itemsize == 48

# a few entries of Mitsuba's named indices of refraction (src/bsdfs/ior.h); defaults: intIOR "bk7", extIOR "air"
_IOR = {"vacuum": 1.0, "air": 1.000277, "water": 1.3330, "bk7": 1.5046, "diamond": 2.419, "pyrex": 1.470, "acrylic glass": 1.49,
        "polypropylene": 1.49, "sodium chloride": 1.544, "amber": 1.55, "pet": 1.5750, "helium": 1.000036, "hydrogen": 1.000132, "water ice": 1.31}


def _lookup_ior(v, default):
    if v is None:
        v = default
    try:
        return float(v)
    except ValueError:
        return _IOR[v.lower()]


# --------------------------------------------------------------------------- CIE data

def _parse_c_array(text: str, name: str) -> np.ndarray:
    m = re.search(r"const\s+Float\s+" + name + r"\s*\[[^\]]*\]\s*=\s*\{(.*?)\};", text, re.S)
    if not m:
        raise RuntimeError(f"table {name} not found")
    body = re.sub(r"/\*.*?\*/", "", m.group(1), flags=re.S)
    vals = [float(t.rstrip("fF")) for t in re.split(r"[,\s]+", body.strip()) if t]
    return np.asarray(vals, dtype=np.float64)


def extract_cie_tables(ref_cpp: str = _REF_SPECTRUM_CPP) -> dict:
    """Read the CIE 1931 2-degree observer + D65 tables (standard colorimetric
    data, 471 samples 360..830 nm) out of the reference tree.  Only used by
    tools/make_fixtures.py to (re)generate ppg_b200/data/cie1931.npz."""
    text = open(ref_cpp, "r", errors="replace").read()
    out = {}
    for key, name in (("x", "CIE_X_entries"), ("y", "CIE_Y_entries"), ("z", "CIE_Z_entries"), ("d65", "CIE_D65_entries")):
        out[key] = _parse_c_array(text, name)
    n = len(out["x"])
    # CIE_wavelengths is 360..830 in 1 nm steps (spectrum.cpp: CIE_wavelengths[CIE_samples])
    out["lambda"] = np.arange(360.0, 360.0 + n, 1.0)
    # D65 is normalised so that it has unit luminance (spectrum.cpp staticInitialization)
    return out


_cie_cache = None


def _cie():
    global _cie_cache
    if _cie_cache is None:
        if os.path.exists(_CIE_NPZ):
            d = np.load(_CIE_NPZ)
            _cie_cache = {k: d[k] for k in d.files}
        elif os.path.exists(_REF_SPECTRUM_CPP):
            _cie_cache = extract_cie_tables()
        else:
            raise RuntimeError("CIE tables unavailable: run tools/make_fixtures.py where /root/reference exists")
    return _cie_cache


def _eval_reversed(l, v, x_mid, x):
    """InterpolatedSpectrum::eval (src/libcore/spectrum.cpp:688-714) strictly inside a
    knot interval.  NOTE the argument order of its lerp: ``math::lerp(t, fb, fa)`` --
    the reference interpolates *mirrored* inside every interval (value fb at the left
    knot side, fa at the right).  fromContinuousSpectrum integrates exactly this
    function (Gauss-Lobatto over ProductSpectrum::eval), so the RGB values of coarse
    spectra (the CBOX emitter: 100 nm knots) depend on the quirk; we reproduce it.
    ``x_mid`` selects the interval, ``x`` is where the interval's line is evaluated."""
    j = np.searchsorted(l, x_mid, side="right") - 1
    inside = (x_mid > l[0]) & (x_mid < l[-1])
    j = np.clip(j, 0, len(l) - 2)
    a, b = l[j], l[j + 1]
    t = (x - a) / (b - a)
    val = (1.0 - t) * v[j + 1] + t * v[j]
    return np.where(inside, val, 0.0)


def _integrate_product(l1, v1, l2, v2, lo, hi):
    """Integral over [lo,hi] of the product of two InterpolatedSpectrum::eval functions
    (piecewise linear per knot interval, zero outside): Simpson per merged knot interval
    is exact for the quadratic pieces."""
    knots = np.unique(np.concatenate([l1, l2, [lo, hi]]))
    knots = knots[(knots >= lo) & (knots <= hi)]
    a, b = knots[:-1], knots[1:]
    mid = 0.5 * (a + b)
    fa = _eval_reversed(l1, v1, mid, a) * _eval_reversed(l2, v2, mid, a)
    fm = _eval_reversed(l1, v1, mid, mid) * _eval_reversed(l2, v2, mid, mid)
    fb = _eval_reversed(l1, v1, mid, b) * _eval_reversed(l2, v2, mid, b)
    return float(np.sum((b - a) / 6.0 * (fa + 4.0 * fm + fb)))


def continuous_to_rgb(lam, val) -> np.ndarray:
    """Spectrum::fromContinuousSpectrum of an InterpolatedSpectrum in an RGB build (src/libcore/spectrum.cpp:172-191, 222-227);
    not clamped."""
    lam = np.asarray(lam, dtype=np.float64); val = np.asarray(val, dtype=np.float64)
    c = _cie()
    cl = c["lambda"]
    lo, hi = cl[0], cl[-1]
    X = _integrate_product(lam, val, cl, c["x"], lo, hi)
    Y = _integrate_product(lam, val, cl, c["y"], lo, hi)
    Z = _integrate_product(lam, val, cl, c["z"], lo, hi)
    ynorm = float(np.trapezoid(c["y"], cl))
    X, Y, Z = X / ynorm, Y / ynorm, Z / ynorm
    return np.array([
        3.240479 * X - 1.537150 * Y - 0.498535 * Z,
        -0.969256 * X + 1.875991 * Y + 0.041556 * Z,
        0.055648 * X - 0.204043 * Y + 1.057311 * Z,
    ])


def spectrum_to_rgb(wavelengths, values) -> np.ndarray:
    """InterpolatedSpectrum(...).zeroExtend() -> Spectrum::fromContinuousSpectrum
    (RGB build) -> clampNegative."""
    lam = np.asarray(wavelengths, dtype=np.float64)
    val = np.asarray(values, dtype=np.float64)
    if len(lam) < 2:
        raise ValueError("InterpolatedSpectrum::zeroExtend() -- at least 2 entries are needed!")
    spacing = float(np.mean(np.diff(lam)))
    if val[0] != 0:
        lam = np.concatenate([[lam[0] - spacing], lam]); val = np.concatenate([[0.0], val])
    if val[-1] != 0:
        lam = np.concatenate([lam, [lam[-1] + spacing]]); val = np.concatenate([val, [0.0]])
    return np.maximum(continuous_to_rgb(lam, val), 0.0).astype(np.float32)


def d65_rgb() -> np.ndarray:
    """Spectrum::getD65() in an RGB build is Spectrum(1.0f) (spectrum.cpp:163)."""
    return np.ones(3, dtype=np.float32)


# --------------------------------------------------------------------------- transforms

def _translate(x, y, z):
    m = np.eye(4); m[:3, 3] = (x, y, z); return m


def _scale(x, y, z):
    return np.diag([x, y, z, 1.0])


def _rotate(axis, angle_deg):
    """Transform::rotate(axis, angle) (src/libcore/transform.cpp) -- Rodrigues."""
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    s, c = math.sin(math.radians(angle_deg)), math.cos(math.radians(angle_deg))
    x, y, z = a
    m = np.eye(4)
    m[0, 0] = x * x + (1 - x * x) * c; m[0, 1] = x * y * (1 - c) - z * s; m[0, 2] = x * z * (1 - c) + y * s
    m[1, 0] = x * y * (1 - c) + z * s; m[1, 1] = y * y + (1 - y * y) * c; m[1, 2] = y * z * (1 - c) - x * s
    m[2, 0] = x * z * (1 - c) - y * s; m[2, 1] = y * z * (1 - c) + x * s; m[2, 2] = z * z + (1 - z * z) * c
    return m


def _look_at(origin, target, up):
    o = np.asarray(origin, dtype=np.float64); t = np.asarray(target, dtype=np.float64); u = np.asarray(up, dtype=np.float64)
    d = t - o; d /= np.linalg.norm(d)
    left = np.cross(u, d); left /= np.linalg.norm(left)
    new_up = np.cross(d, left)
    m = np.eye(4)
    m[:3, 0] = left; m[:3, 1] = new_up; m[:3, 2] = d; m[:3, 3] = o
    return m


def _floats(s):
    return [float(t) for t in re.split(r"[,\s]+", s.strip()) if t]


def _parse_transform(node) -> np.ndarray:
    m = np.eye(4)
    for ch in node:
        a = ch.attrib
        if ch.tag == "translate":
            t = _translate(float(a.get("x", 0)), float(a.get("y", 0)), float(a.get("z", 0)))
        elif ch.tag == "rotate":
            t = _rotate((float(a.get("x", 0)), float(a.get("y", 0)), float(a.get("z", 0))), float(a["angle"]))
        elif ch.tag == "scale":
            if "value" in a:
                v = float(a["value"]); t = _scale(v, v, v)
            else:
                t = _scale(float(a.get("x", 1)), float(a.get("y", 1)), float(a.get("z", 1)))
        elif ch.tag in ("lookAt", "lookat"):
            up = _floats(a["up"]) if a.get("up") else [0, 1, 0]
            t = _look_at(_floats(a["origin"]), _floats(a["target"]), up)
        elif ch.tag == "matrix":
            t = np.asarray(_floats(a["value"]), dtype=np.float64).reshape(4, 4)
        else:
            raise ValueError(f"unsupported transform element <{ch.tag}>")
        m = t @ m
    return m


# --------------------------------------------------------------------------- meshes

def _load_obj_fast(path, to_world, face_normals, flip_normals):
    """Vectorised path for the common export style: triangles only, every corner 'p/t/n' with positive indices.
    Returns None when the file does not fit (the generic loader takes over).  Vertices are merged per (p, t, n) index
    triple instead of per value -- same surface, possibly a few more vertices."""
    v, vn, vt, fl = [], [], [], []
    with open(path, "r", errors="replace") as f:
        for line in f:
            if line.startswith("v "): v.append(line[2:])
            elif line.startswith("vn "): vn.append(line[3:])
            elif line.startswith("vt "): vt.append(line[3:])
            elif line.startswith("f "): fl.append(line[2:])
    if not fl or not v or not vn or not vt:
        return None
    ftxt = " ".join(fl)
    if "//" in ftxt or "-" in ftxt:
        return None
    try:
        F = np.array(ftxt.replace("/", " ").split(), dtype=np.int64)
    except ValueError:
        return None
    if F.size != 9 * len(fl) or ftxt.count("/") != 6 * len(fl):
        return None
    verts = np.array(" ".join(v).split(), dtype=np.float64).reshape(len(v), -1)[:, :3]
    norms = np.array(" ".join(vn).split(), dtype=np.float64).reshape(len(vn), -1)[:, :3]
    uvs = np.array(" ".join(vt).split(), dtype=np.float64).reshape(len(vt), -1)[:, :2]
    corners = F.reshape(-1, 3)                                   # (3T, [p, t, n]) 1-based
    uniq, inv = np.unique(corners, axis=0, return_inverse=True)
    M = to_world
    P = (verts[uniq[:, 0] - 1] @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
    Nw = norms[uniq[:, 2] - 1] @ np.linalg.inv(M[:3, :3])     # (M^-T n)^T = n^T M^-1
    l = np.linalg.norm(Nw, axis=1, keepdims=True)
    N = np.where(l > 0, Nw / np.maximum(l, 1e-300), Nw).astype(np.float32)
    UV = uvs[uniq[:, 1] - 1].astype(np.float32); UV[:, 1] = 1 - UV[:, 1]        # flipTexCoords default true (obj.cpp:211, 306)
    I = inv.reshape(-1, 3).astype(np.uint32)
    if face_normals:
        if flip_normals: I = I[:, [1, 0, 2]]
        return P, None, UV, I
    if flip_normals: N = -N
    return P, N, UV, I


def _load_obj(path, to_world, face_normals=False, flip_normals=False):
    """WavefrontOBJ -> one merged triangle mesh (the bundled OBJs hold one object each)."""
    fast = _load_obj_fast(path, to_world, face_normals, flip_normals)
    if fast is not None:
        return fast
    verts, norms, uvs, tris = [], [], [], []
    with open(path, "r", errors="replace") as f:
        for line in f:
            tok = line.split()
            if not tok:
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "vn":
                norms.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "vt":
                uvs.append([float(tok[1]), float(tok[2])])
            elif tok[0] == "f":
                corners = []
                for s in tok[1:]:
                    parts = s.split("/")
                    p = int(parts[0]); uv = 0; n = 0
                    if len(parts) == 2:
                        uv = int(parts[1]) if parts[1] else 0
                    elif len(parts) == 3:
                        uv = int(parts[1]) if parts[1] else 0
                        n = int(parts[2]) if parts[2] else 0
                    corners.append((p, uv, n))
                for i in range(1, len(corners) - 1):  # fan, like obj.cpp:313-324
                    tris.append((corners[0], corners[i], corners[i + 1]))
    verts = np.asarray(verts, dtype=np.float64).reshape(-1, 3)
    norms = np.asarray(norms, dtype=np.float64).reshape(-1, 3)
    uvs = np.asarray(uvs, dtype=np.float64).reshape(-1, 2)
    M = to_world
    # normals transform with the inverse transpose
    Mn = np.linalg.inv(M[:3, :3]).T
    key_to_idx = {}
    P, N, UV, I = [], [], [], []
    has_n = has_uv = False
    for tri in tris:
        idx = []
        for (p, uv, n) in tri:
            if p < 0: p += len(verts) + 1
            if n < 0: n += len(norms) + 1
            if uv < 0: uv += len(uvs) + 1
            pw = (M[:3, :3] @ verts[p - 1] + M[:3, 3]).astype(np.float32)
            if n != 0:
                nw = Mn @ norms[n - 1]
                l = np.linalg.norm(nw)
                if l > 0: nw = nw / l
                nw = nw.astype(np.float32); has_n = True
            else:
                nw = np.zeros(3, np.float32)
            if uv != 0:
                t = uvs[uv - 1].astype(np.float32).copy(); t[1] = 1 - t[1]  # flipTexCoords default true
                has_uv = True
            else:
                t = np.zeros(2, np.float32)
            key = (pw.tobytes(), nw.tobytes(), t.tobytes())
            k = key_to_idx.get(key)
            if k is None:
                k = len(P); key_to_idx[key] = k
                P.append(pw); N.append(nw); UV.append(t)
            idx.append(k)
        I.append(idx)
    P = np.asarray(P, np.float32).reshape(-1, 3); N = np.asarray(N, np.float32).reshape(-1, 3)
    UV = np.asarray(UV, np.float32).reshape(-1, 2); I = np.asarray(I, np.uint32).reshape(-1, 3)
    if face_normals:
        if flip_normals:
            I = I[:, [1, 0, 2]]
        return P, None, (UV if has_uv else None), I
    if has_n:
        if flip_normals: N = -N
    else:
        N = _smooth_normals(P, I)
        if flip_normals: N = -N
    return P, N, (UV if has_uv else None), I


def _finish_mesh(P, N, UV, I, to_world, face_normals, flip_normals):
    """Object-space arrays of a mesh file -> world space, with the conventions of _load_obj: positions by toWorld, normals by its inverse transpose
    (normalised), `faceNormals` drops vertex normals (flipNormals then swaps the winding), meshes without normals get angle-weighted ones (TriMesh::configure)."""
    M = np.asarray(to_world, np.float64)
    P = (np.asarray(P, np.float64) @ M[:3, :3].T + M[:3, 3]).astype(np.float32)
    I = np.asarray(I, np.uint32).reshape(-1, 3)
    if len(I) and int(I.max()) >= len(P):
        raise ValueError("mesh file: vertex index out of range")
    UV = None if UV is None else np.asarray(UV, np.float32).reshape(-1, 2)
    if face_normals:
        return P, None, UV, (I[:, [1, 0, 2]] if flip_normals else I)
    if N is not None:
        Nw = np.asarray(N, np.float64) @ np.linalg.inv(M[:3, :3])            # rows x (M^-1) == (M^-T n) per row
        l = np.linalg.norm(Nw, axis=1, keepdims=True)
        N = np.where(l > 0, Nw / np.maximum(l, 1e-300), Nw).astype(np.float32)
    else:
        N = _smooth_normals(P, I)
    return P, (-N if flip_normals else N), UV, I


def _load_ply(path, to_world, face_normals=False, flip_normals=False):
    """Stanford PLY as src/shapes/ply.cpp reads it: ascii / binary_little_endian / binary_big_endian; vertex properties x y z [nx ny nz] [u v | s t |
    texture_u texture_v] (anything else, e.g. colours, is skipped), faces `vertex_indices` / `vertex_index` with 3 or 4 corners -- a quad (a, b, c, d)
    becomes (a, b, c), (d, a, c) (ply.cpp:299-312)."""
    types = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4",
             "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt = None; elements = []
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                elements.append([tok[1], int(tok[2]), []])
            elif tok[0] == "property":
                elements[-1][2].append(tuple(tok[1:]))
            elif tok[0] == "end_header":
                break
        if fmt not in ("ascii", "binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unknown PLY format {fmt}")
        end = ">" if fmt == "binary_big_endian" else "<"
        verts = {}; faces = []
        for name, count, props in elements:
            has_list = any(p[0] == "list" for p in props)
            if fmt == "ascii":
                rows = [f.readline().split() for _ in range(count)]
                if name == "vertex":
                    arr = np.asarray(rows, np.float64).reshape(count, len(props))
                    verts = {p[-1]: arr[:, k] for k, p in enumerate(props)}
                elif name == "face":
                    for r in rows:                                           # (only the index list; leading scalar properties are not used by ply.cpp either)
                        k = 0
                        for pr in props:
                            if pr[0] == "list":
                                n = int(r[k]); vals = [int(v) for v in r[k + 1:k + 1 + n]]; k += 1 + n
                                if pr[-1] in ("vertex_indices", "vertex_index"):
                                    faces.append(vals)
                            else:
                                k += 1
            elif not has_list:
                dt = np.dtype([(p[-1], end + types[p[0]]) for p in props])
                arr = np.frombuffer(f.read(dt.itemsize * count), dt, count)
                if name == "vertex":
                    verts = {n: arr[n].astype(np.float64) for n in arr.dtype.names}
            else:
                if name == "face" and len(props) == 1 and props[0][-1] in ("vertex_indices", "vertex_index"):
                    # the common case in one pass: every face has the same corner count (all triangles or all quads)
                    ct, it = np.dtype(end + types[props[0][1]]), np.dtype(end + types[props[0][2]])
                    pos = f.tell(); n0 = int(np.frombuffer(f.read(ct.itemsize), ct, 1)[0]); f.seek(pos)
                    rec = np.dtype([("n", ct), ("v", it, (n0,))])
                    raw = f.read(rec.itemsize * count)
                    arr = np.frombuffer(raw, rec, count) if len(raw) == rec.itemsize * count else None
                    if arr is not None and (arr["n"] == n0).all():
                        faces = arr["v"].astype(np.int64).tolist() if n0 not in (3, 4) else arr["v"].astype(np.int64)
                        continue
                    f.seek(pos)
                for _ in range(count):
                    for pr in props:
                        if pr[0] == "list":
                            ct, it = np.dtype(end + types[pr[1]]), np.dtype(end + types[pr[2]])
                            n = int(np.frombuffer(f.read(ct.itemsize), ct, 1)[0])
                            vals = np.frombuffer(f.read(it.itemsize * n), it, n)
                            if name == "face" and pr[-1] in ("vertex_indices", "vertex_index"):
                                faces.append([int(v) for v in vals])
                        else:
                            f.read(np.dtype(types[pr[0]]).itemsize)
    if not all(k in verts for k in "xyz"):
        raise ValueError(f"{path}: PLY without vertex positions")
    P = np.stack([verts["x"], verts["y"], verts["z"]], 1)
    N = np.stack([verts["nx"], verts["ny"], verts["nz"]], 1) if all(k in verts for k in ("nx", "ny", "nz")) else None
    UV = None
    for a, b in (("u", "v"), ("s", "t"), ("texture_u", "texture_v")):
        if a in verts and b in verts:
            UV = np.stack([verts[a], verts[b]], 1)
    tris = []
    if isinstance(faces, np.ndarray):
        tris = faces if faces.shape[1] == 3 else np.concatenate([faces[:, [0, 1, 2]][:, None], faces[:, [3, 0, 2]][:, None]], 1).reshape(-1, 3)
    else:
        for fc in faces:
            if len(fc) not in (3, 4):
                raise NotImplementedError("Only triangle and quad-based PLY meshes are supported for now.")       # ply.cpp:274-284
            tris.append(fc[:3])
            if len(fc) == 4:
                tris.append([fc[3], fc[0], fc[2]])
    return _finish_mesh(P, N, UV, np.asarray(tris, np.int64).reshape(-1, 3), to_world, face_normals, flip_normals)


def _load_serialized(path, shape_index, to_world, face_normals=False, flip_normals=False):
    """Mitsuba's .serialized triangle meshes (TriMesh::loadCompressed, src/librender/trimesh.cpp:176-250; shapes/serialized.cpp): little endian,
    [u16 0x041C][u16 version 3 | 4] then a zlib stream {u32 flags, (v4) name\0, u64 vertices, u64 triangles, positions, [normals], [texcoords],
    [colours], u32 indices}; several meshes per file are found through the offset table at the end (u64 offsets for v4, u32 for v3, then u32 count)."""
    import struct
    import zlib
    with open(path, "rb") as f:
        data = f.read()
    def header(off):
        fmt, ver = struct.unpack_from("<HH", data, off)
        if fmt != 0x041C:
            raise ValueError(f"{path}: Encountered an invalid file format!")
        if ver not in (3, 4):
            raise ValueError(f"{path}: Encountered an incompatible file version!")
        return ver
    ver = header(0); off = 0
    if shape_index != 0:
        count, = struct.unpack_from("<I", data, len(data) - 4)
        if shape_index < 0 or shape_index >= count:
            raise ValueError(f"Unable to unserialize mesh, shape index is out of range! (requested {shape_index} out of 0..{count - 1})")
        if ver == 4:
            off, = struct.unpack_from("<Q", data, len(data) - 8 * (count - shape_index) - 4)
        else:
            off, = struct.unpack_from("<I", data, len(data) - 4 * (count - shape_index + 1))
        header(off)
    raw = zlib.decompressobj().decompress(data[off + 4:])
    flags, = struct.unpack_from("<I", raw, 0); p = 4
    if ver == 4:
        p = raw.index(b"\0", p) + 1
    nv, nt = struct.unpack_from("<QQ", raw, p); p += 16
    ft = "<f8" if flags & 0x2000 else "<f4"; fs = 8 if flags & 0x2000 else 4
    def take(ncomp):
        nonlocal p
        a = np.frombuffer(raw, ft, nv * ncomp, p).reshape(nv, ncomp).astype(np.float64); p += fs * nv * ncomp
        return a
    P = take(3)
    N = take(3) if flags & 0x0001 else None
    UV = take(2) if flags & 0x0002 else None
    if flags & 0x0008:
        take(3)                                                             # vertex colours: not used by any model in scope
    I = np.frombuffer(raw, "<u4", nt * 3, p).reshape(nt, 3)
    return _finish_mesh(P, N, UV, I, to_world, face_normals or bool(flags & 0x0010), flip_normals)


def _cube(to_world, flip_normals=False):
    """shapes/cube.cpp:24-30, 81-108: 24 vertices (4 per face, own normals and [0,1]^2 texture coordinates), 12 triangles, [-1,1]^3 under toWorld."""
    P = [(1, -1, -1), (1, -1, 1), (-1, -1, 1), (-1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, 1, 1), (1, 1, 1), (1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1),
         (1, -1, 1), (1, 1, 1), (-1, 1, 1), (-1, -1, 1), (-1, -1, 1), (-1, 1, 1), (-1, 1, -1), (-1, -1, -1), (1, 1, -1), (1, -1, -1), (-1, -1, -1), (-1, 1, -1)]
    N = [(0, -1, 0)] * 4 + [(0, 1, 0)] * 4 + [(1, 0, 0)] * 4 + [(0, 0, 1)] * 4 + [(-1, 0, 0)] * 4 + [(0, 0, -1)] * 4
    UV = [(0, 1), (1, 1), (1, 0), (0, 0)] * 6
    I = [(0, 1, 2), (3, 0, 2), (4, 5, 6), (7, 4, 6), (8, 9, 10), (11, 8, 10), (12, 13, 14), (15, 12, 14), (16, 17, 18), (19, 16, 18), (20, 21, 22), (23, 20, 22)]
    return _finish_mesh(np.array(P, np.float64), np.array(N, np.float64), np.array(UV, np.float64), np.array(I), to_world, False, flip_normals)


def _smooth_normals(P, I):
    """Angle-weighted vertex normals (Thuermer & Wuethrich), trimesh.cpp:636-690."""
    N = np.zeros((len(P), 3), np.float64)
    Pd = P.astype(np.float64)
    for tri in I:
        for i in range(3):
            v0, v1, v2 = Pd[tri[i]], Pd[tri[(i + 1) % 3]], Pd[tri[(i + 2) % 3]]
            a, b = v1 - v0, v2 - v0
            if i == 0:
                n = np.cross(a, b); l = np.linalg.norm(n)
                if l == 0:
                    break
                n = n / l
            la, lb = np.linalg.norm(a), np.linalg.norm(b)
            if la == 0 or lb == 0:
                continue
            ang = math.acos(max(-1.0, min(1.0, float(np.dot(a, b) / (la * lb)))))
            N[tri[i]] += n * ang
    l = np.linalg.norm(N, axis=1, keepdims=True)
    N = np.where(l > 0, N / np.maximum(l, 1e-300), np.array([1.0, 0.0, 0.0]))
    return N.astype(np.float32)


def _rectangle(to_world):
    """shapes/rectangle.cpp:125-148: unit square [-1,1]^2 in the xy-plane, normal +z."""
    p = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], np.float64)
    P = (p @ to_world[:3, :3].T + to_world[:3, 3]).astype(np.float32)
    n = np.linalg.inv(to_world[:3, :3]).T @ np.array([0, 0, 1.0]); n /= np.linalg.norm(n)
    N = np.tile(n.astype(np.float32), (4, 1))
    UV = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    I = np.array([[0, 1, 2], [0, 2, 3]], np.uint32)
    return P, N, UV, I


# --------------------------------------------------------------------------- scene

@dataclass
class SceneDesc:
    positions: np.ndarray      # (V,3) f32
    normals: np.ndarray        # (V,3) f32
    uvs: np.ndarray            # (V,2) f32
    indices: np.ndarray        # (T,3) u32
    triangle_shape: np.ndarray  # (T,) u32
    shapes: np.ndarray         # (S,8) u32/i32: first_tri, n_tris, bsdf, emitter, has_normals, has_uvs, 0, 0
    bsdfs: np.ndarray          # (B,28) f32 view of ppg_bsdf (type/flags bit-cast)
    area_radiance: np.ndarray  # (E,3) f32
    cam_to_world: np.ndarray   # (4,4) f32
    x_fov_deg: float
    near_clip: float
    far_clip: float
    film_width: int
    film_height: int
    aabb_min: np.ndarray
    aabb_max: np.ndarray
    integrator: dict = field(default_factory=dict)  # XML name -> string value
    bsdf_names: list = field(default_factory=list)
    bsdf_tables: np.ndarray = None   # (T,100) f32 rough-transmittance tables referenced by roughplastic materials
    spheres: np.ndarray = None       # (K,6) f32 view of ppg_sphere: center[3], radius, shape (int bits), flip_normals (int bits)
    textures: np.ndarray = None      # (N,) TEXTURE_DTYPE == ppg_texture
    texels: np.ndarray = None        # uint16 (IEEE half bits) of all textures
    envmap: dict = None              # {"texels": (H,W,3) uint16 half bits, "scale": float, "world_to_env": (3,3) f32} or None

    def with_film(self, w: int, h: int) -> "SceneDesc":
        """Same scene, different film size (x fov re-resolved only when the aspect
        changes the 'smaller' axis -- callers keep aspect)."""
        import copy
        s = copy.copy(self); s.film_width = int(w); s.film_height = int(h); return s

    def save(self, path):
        np.savez_compressed(
            path, positions=self.positions, normals=self.normals, uvs=self.uvs, indices=self.indices,
            triangle_shape=self.triangle_shape, shapes=self.shapes, bsdfs=self.bsdfs,
            area_radiance=self.area_radiance, cam_to_world=self.cam_to_world,
            cam=np.array([self.x_fov_deg, self.near_clip, self.far_clip, self.film_width, self.film_height], np.float64),
            aabb=np.stack([self.aabb_min, self.aabb_max]).astype(np.float32),
            integrator=np.array([f"{k}={v}" for k, v in self.integrator.items()]),
            bsdf_names=np.array(self.bsdf_names), bsdf_tables=(self.bsdf_tables if self.bsdf_tables is not None else np.zeros((0, 100), np.float32)),
            spheres=(self.spheres if self.spheres is not None else np.zeros((0, 6), np.float32)),
            textures=(self.textures if self.textures is not None else np.zeros(0, TEXTURE_DTYPE)).view(np.uint8),
            texels=(self.texels if self.texels is not None else np.zeros(0, np.uint16)),
            env_texels=(self.envmap["texels"] if self.envmap else np.zeros((0, 0, 3), np.uint16)),
            env_meta=(np.concatenate([[self.envmap["scale"]], np.asarray(self.envmap["world_to_env"], np.float64).reshape(9)]) if self.envmap else np.zeros(0)))

    def save_flat(self, path):
        """Flat binary form for C / C++ hosts (read by ppg_scene_file_load, include/ppg.h): magic "PPGSCN02", then named arrays
        [u32 name length][name][u32 dtype: 0 f32, 1 u32, 2 i32, 3 u16, 4 u8, 5 f64][u32 ndim][u64 dims...][raw little-endian data]."""
        import struct
        codes = {"float32": 0, "uint32": 1, "int32": 2, "uint16": 3, "uint8": 4, "float64": 5}
        arrays = {
            "positions": np.asarray(self.positions, np.float32), "normals": np.asarray(self.normals, np.float32), "uvs": np.asarray(self.uvs, np.float32),
            "indices": np.asarray(self.indices, np.uint32), "triangle_shape": np.asarray(self.triangle_shape, np.uint32), "shapes": np.asarray(self.shapes, np.int32),
            "bsdfs": np.asarray(self.bsdfs, np.float32), "area_radiance": np.asarray(self.area_radiance, np.float32).reshape(-1, 3),
            "bsdf_tables": np.asarray(self.bsdf_tables if self.bsdf_tables is not None else np.zeros((0, 100)), np.float32),
            "spheres": np.asarray(self.spheres if self.spheres is not None else np.zeros((0, 6)), np.float32),
            "cam_to_world": np.asarray(self.cam_to_world, np.float32).reshape(16),
            "cam": np.array([self.x_fov_deg, self.near_clip, self.far_clip, self.film_width, self.film_height], np.float64),
            "aabb": np.stack([self.aabb_min, self.aabb_max]).astype(np.float32),
            "textures": (self.textures if self.textures is not None else np.zeros(0, TEXTURE_DTYPE)).view(np.uint8),
            "texels": np.asarray(self.texels if self.texels is not None else np.zeros(0), np.uint16),
            "env_texels": np.asarray(self.envmap["texels"] if self.envmap else np.zeros((0, 0, 3)), np.uint16),
            "env_meta": (np.concatenate([[self.envmap["scale"]], np.asarray(self.envmap["world_to_env"], np.float64).reshape(9)]) if self.envmap else np.zeros(0)).astype(np.float32),
            "integrator": np.frombuffer("\n".join(f"{k}={v}" for k, v in self.integrator.items()).encode(), np.uint8),
        }
        with open(path, "wb") as f:
            f.write(b"PPGSCN02")
            for name, a in arrays.items():
                a = np.ascontiguousarray(a)
                f.write(struct.pack("<I", len(name))); f.write(name.encode())
                f.write(struct.pack("<II", codes[a.dtype.name], a.ndim)); f.write(struct.pack(f"<{a.ndim}Q", *a.shape))
                f.write(a.tobytes())

    @staticmethod
    def load(path) -> "SceneDesc":
        d = np.load(path, allow_pickle=False)
        cam = d["cam"]
        integ = dict(s.split("=", 1) for s in d["integrator"].tolist())
        sc = SceneDesc(d["positions"], d["normals"], d["uvs"], d["indices"], d["triangle_shape"], d["shapes"],
                       d["bsdfs"], d["area_radiance"], d["cam_to_world"], float(cam[0]), float(cam[1]), float(cam[2]),
                       int(cam[3]), int(cam[4]), d["aabb"][0], d["aabb"][1], integ, d["bsdf_names"].tolist(),
                       d["bsdf_tables"] if "bsdf_tables" in d.files else None, d["spheres"] if "spheres" in d.files else None)
        if "textures" in d.files and d["textures"].size:
            sc.textures = d["textures"].view(TEXTURE_DTYPE).copy(); sc.texels = d["texels"]
        if "env_texels" in d.files and d["env_texels"].size:
            m = d["env_meta"]
            sc.envmap = {"texels": d["env_texels"], "scale": float(m[0]), "world_to_env": m[1:10].reshape(3, 3).astype(np.float32)}
        return sc


def make_sphere(center, radius, shape, flip_normals=False):
    """One ppg_sphere (include/ppg.h) as 6 floats."""
    r = np.zeros(6, np.float32)
    r[:3] = center; r[3] = radius
    r[4:6] = np.array([shape, 1 if flip_normals else 0], np.int32).view(np.float32)
    return r


def _make_bsdf(type_, flags, refl, trans=(0, 0, 0), eta=(0, 0, 0), k=(0, 0, 0), alpha=0.1, distribution=0):
    """One ppg_bsdf (include/ppg.h) as 28 floats: type, flags, reflectance[3], specular_transmittance[3], eta[3], k[3], alpha, distribution (int bits),
    specular_reflectance[3], fdr_int, specular_sampling_weight, table (int bits), opacity[3], reflectance_texture / bump_texture / reserved (uint bits);
    the trailing fields stay 0 here."""
    b = np.zeros(28, np.float32)
    b[:2] = np.array([type_, flags], np.uint32).view(np.float32)
    b[2:5] = refl; b[5:8] = trans; b[8:11] = eta; b[11:14] = k; b[14] = alpha
    b[15:16] = np.array([distribution], np.int32).view(np.float32)
    return b


def fresnel_diffuse_reflectance(eta: float) -> float:
    """fresnelDiffuseReflectance(eta, fast=false) (src/libcore/util.cpp:807-862): integral over xi in [0,1] of
    fresnelDielectricExt(sqrt(xi), eta).  The reference integrates adaptively (Gauss-Lobatto, relative error 1e-5); here a fixed
    composite Simpson rule in double precision after the substitution xi = c^2 (smooth integrand) -- same value to ~1e-7."""
    n = 1 << 14
    c = np.linspace(0.0, 1.0, n + 1)
    scale = np.where(c > 0, 1.0 / eta, eta)
    ct2 = 1.0 - (1.0 - c * c) * scale * scale
    tir = ct2 <= 0
    ct = np.sqrt(np.where(tir, 0.0, ct2))
    with np.errstate(invalid="ignore", divide="ignore"):
        rs = (c - eta * ct) / (c + eta * ct); rp = (eta * c - ct) / (eta * c + ct)
        F = np.where(tir, 1.0, 0.5 * (rs * rs + rp * rp))
    F = np.where(np.isfinite(F), F, 1.0)
    f = F * 2.0 * c                                   # d(xi) = 2 c dc
    w = np.ones(n + 1); w[1:-1:2] = 4; w[2:-1:2] = 2
    return float(np.sum(f * w) / (3.0 * n))


def make_plastic(flags, diffuse, specular, eta, nonlinear):
    """ppg_bsdf for the smooth plastic (src/bsdfs/plastic.cpp:185-212)."""
    lum = lambda c: 0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]
    d_avg, s_avg = lum(diffuse), lum(specular)
    b = _make_bsdf(BSDF_PLASTIC, flags | (BSDF_FLAG_NONLINEAR if nonlinear else 0), diffuse, (0, 0, 0), (eta, eta, eta))
    b[16:19] = specular; b[19] = fresnel_diffuse_reflectance(1.0 / eta); b[20] = s_avg / (d_avg + s_avg)
    return b


def make_roughplastic(flags, diffuse, specular, eta, alpha, distribution, nonlinear, tables):
    """ppg_bsdf for roughplastic (src/bsdfs/roughplastic.cpp:190-290); appends its rough-transmittance table to `tables`."""
    from . import rtrans
    lut, fdr = rtrans.reduce_for_material("ggx" if distribution == 1 else "beckmann", eta, alpha)
    lum = lambda c: 0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]
    d_avg, s_avg = lum(diffuse), lum(specular)
    b = _make_bsdf(BSDF_ROUGHPLASTIC, flags | (BSDF_FLAG_NONLINEAR if nonlinear else 0), diffuse, (0, 0, 0), (eta, eta, eta), (0, 0, 0), alpha, distribution)
    b[16:19] = specular; b[19] = fdr; b[20] = s_avg / (d_avg + s_avg)
    b[21:22] = np.array([len(tables)], np.int32).view(np.float32)
    tables.append(lut)
    return b


def _prop_children(node):
    out = {}
    for ch in node:
        if "name" in ch.attrib and ch.tag in ("string", "integer", "float", "boolean"):
            out[ch.attrib["name"]] = ch.attrib["value"]
    return out


def _parse_color(node, is_emitter=False):
    if node.tag == "rgb":
        return np.asarray(_floats(node.attrib["value"]), np.float32)
    if node.tag == "srgb":
        v = np.asarray(_floats(node.attrib["value"]), np.float64)
        lin = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)
        return lin.astype(np.float32)
    if node.tag == "spectrum":
        val = node.attrib["value"]
        if ":" in val:
            pairs = [t.split(":") for t in re.split(r"[,\s]+", val.strip()) if t]
            return spectrum_to_rgb([float(p[0]) for p in pairs], [float(p[1]) for p in pairs])
        toks = _floats(val)
        if len(toks) == 1:
            return (d65_rgb() * toks[0]) if is_emitter else np.full(3, toks[0], np.float32)
        if len(toks) == 3:
            return np.asarray(toks, np.float32)
    raise ValueError(f"unsupported colour element <{node.tag}>")


_TABLES = []      # rough-transmittance tables collected while parsing one scene
_TEXTURES = []    # (meta dict, half-bit texel array) collected while parsing one scene
_TEX_CACHE = {}


def load_bitmap(path, gamma=0.0):
    """Bitmap(EAuto, file) -> expand()->convert(ERGB | ELuminance, EFloat, gamma 1) like TMIPMap's constructor
    (include/mitsuba/render/mipmap.h:226-231; src/libcore/bitmap.cpp): 8-bit data are sRGB-decoded (file gamma -1) unless an
    explicit gamma is given.  Returns float32 (H, W, C) with C = 1 or 3."""
    import cv2
    im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if im is None:
        raise FileNotFoundError(f"Texture file \"{path}\" could not be found / decoded!")
    if im.ndim == 2:
        im = im[:, :, None]
    elif im.shape[2] >= 3:
        im = im[:, :, 2::-1]                     # BGR(A) -> RGB, alpha dropped (ERGBA -> ERGB)
    elif im.shape[2] == 2:
        im = im[:, :, :1]                        # luminance + alpha
    if im.dtype == np.uint8:
        v = im.astype(np.float64) / 255.0
    elif im.dtype == np.uint16:
        v = im.astype(np.float64) / 65535.0
    else:
        return np.ascontiguousarray(im, np.float32)      # float formats are linear already
    if gamma == 0:
        lin = np.where(v <= 0.04045, v / 12.92, ((v + 0.055) / 1.055) ** 2.4)
    else:
        lin = v ** gamma
    return np.ascontiguousarray(lin, np.float32)


def _parse_texture(node, base):
    """<texture type="bitmap"> -> index into the scene's texture list (src/textures/bitmap.cpp:178-330)."""
    if node.attrib.get("type") != "bitmap":
        raise NotImplementedError(f"texture '{node.attrib.get('type')}' (only 'bitmap' is in scope)")
    pr = _prop_children(node)
    if pr.get("channel"):
        raise NotImplementedError("bitmap texture: 'channel' extraction")
    path = os.path.join(base, pr["filename"])
    gamma = float(pr.get("gamma", 0))
    wrap = pr.get("wrapMode", "repeat")
    wu, wv = pr.get("wrapModeU", wrap), pr.get("wrapModeV", wrap)
    uvscale = float(pr.get("uvscale", 1.0))
    meta = dict(path=path, gamma=gamma, wrap_u=WRAP_MODES[wu], wrap_v=WRAP_MODES[wv],
                uv_scale=(float(pr.get("uscale", uvscale)), float(pr.get("vscale", uvscale))),
                uv_offset=(float(pr.get("uoffset", 0.0)), float(pr.get("voffset", 0.0))))
    key = tuple(sorted((k, v) for k, v in meta.items()))
    if key in _TEX_CACHE:
        return _TEX_CACHE[key]
    img = load_bitmap(path, gamma)
    meta["width"], meta["height"], meta["channels"] = img.shape[1], img.shape[0], img.shape[2]
    meta["average"] = img.reshape(-1, img.shape[2]).mean(axis=0, dtype=np.float64)      # m_average (mipmap.h:230): of the float data
    with np.errstate(over="ignore"):
        half = img.astype(np.float16)
    _TEXTURES.append((meta, half.view(np.uint16).reshape(-1)))
    _TEX_CACHE[key] = len(_TEXTURES) - 1
    return _TEX_CACHE[key]


def texture_average_rgb(idx):
    a = _TEXTURES[idx][0]["average"]
    return (np.full(3, a[0]) if len(a) == 1 else a).astype(np.float32)


_WRAPPERS = ("bumpmap", "mask", "twosided")


def _parse_bsdf(node, bsdf_table, names, by_id, base="."):
    """One <bsdf> element -> index into bsdf_table.  Wrappers must nest as bumpmap > mask > twosided > model (the order the
    device applies them); every level that carries an id is registered on its own."""
    chain = []
    cur = node
    while cur.attrib["type"] in _WRAPPERS:
        chain.append(cur)
        kids = [c for c in cur if c.tag == "bsdf"]
        if len(kids) != 1:
            raise ValueError(f"{cur.attrib['type']}: exactly one nested BSDF is required")
        cur = kids[0]
    order = [_WRAPPERS.index(c.attrib["type"]) for c in chain]
    if order != sorted(set(order)):
        raise NotImplementedError("BSDF wrappers must nest as bumpmap > mask > twosided")
    inner = cur
    typ = inner.attrib["type"]
    twosided = any(c.attrib["type"] == "twosided" for c in chain)
    colors = {c.attrib.get("name"): c for c in inner if c.tag in ("rgb", "srgb", "spectrum")}
    texs = {c.attrib.get("name"): c for c in inner if c.tag == "texture"}
    props = _prop_children(inner)

    def diffuse_input(nm_list, default):
        """constant colour or bitmap texture for a reflectance slot -> (rgb used for averages, 1-based texture index or 0)"""
        for nm in nm_list:
            if nm in texs:
                ti = _parse_texture(texs[nm], base)
                return texture_average_rgb(ti), ti + 1
            if nm in colors:
                return _parse_color(colors[nm]), 0
        return np.full(3, default, np.float32), 0

    for nm in texs:
        if nm not in ("reflectance", "diffuseReflectance"):
            raise NotImplementedError(f"textured BSDF parameter '{nm}'")
    flags = 0          # wrapper flags are added level by level below; the transmissive models refuse twosided
    tex_idx = 0
    if typ == "diffuse":
        refl, tex_idx = diffuse_input(("reflectance", "diffuseReflectance"), 0.5)   # diffuse.cpp:70-74
        entry = _make_bsdf(BSDF_DIFFUSE, flags, refl)
    elif typ == "dielectric":           # src/bsdfs/dielectric.cpp:157-190
        if twosided:
            raise ValueError("twosided cannot wrap a transmissive BSDF")
        eta = _lookup_ior(props.get("intIOR"), "bk7") / _lookup_ior(props.get("extIOR"), "air")
        sr = _parse_color(colors["specularReflectance"]) if "specularReflectance" in colors else np.ones(3, np.float32)
        st = _parse_color(colors["specularTransmittance"]) if "specularTransmittance" in colors else np.ones(3, np.float32)
        entry = _make_bsdf(BSDF_DIELECTRIC, flags, sr, st, (eta, eta, eta))
    elif typ == "thindielectric":       # src/bsdfs/thindielectric.cpp:88-104
        if twosided:
            raise ValueError("twosided cannot wrap a transmissive BSDF")
        eta = _lookup_ior(props.get("intIOR"), "bk7") / _lookup_ior(props.get("extIOR"), "air")
        sr = _parse_color(colors["specularReflectance"]) if "specularReflectance" in colors else np.ones(3, np.float32)
        st = _parse_color(colors["specularTransmittance"]) if "specularTransmittance" in colors else np.ones(3, np.float32)
        entry = _make_bsdf(BSDF_THINDIELECTRIC, flags, sr, st, (eta, eta, eta))
    elif typ == "roughdielectric":      # src/bsdfs/roughdielectric.cpp:183-210
        if twosided:
            raise ValueError("twosided cannot wrap a transmissive BSDF")
        eta = _lookup_ior(props.get("intIOR"), "bk7") / _lookup_ior(props.get("extIOR"), "air")
        distr = props.get("distribution", "beckmann").lower()
        if distr not in ("beckmann", "ggx"):
            raise NotImplementedError(f"microfacet distribution '{distr}'")
        sr = _parse_color(colors["specularReflectance"]) if "specularReflectance" in colors else np.ones(3, np.float32)
        st = _parse_color(colors["specularTransmittance"]) if "specularTransmittance" in colors else np.ones(3, np.float32)
        entry = _make_bsdf(BSDF_ROUGHDIELECTRIC, flags, sr, st, (eta, eta, eta), (0, 0, 0), float(props.get("alpha", 0.1)), 1 if distr == "ggx" else 0)
    elif typ == "plastic":              # src/bsdfs/plastic.cpp:143-163
        eta = _lookup_ior(props.get("intIOR"), "polypropylene") / _lookup_ior(props.get("extIOR"), "air")
        dr, tex_idx = diffuse_input(("diffuseReflectance",), 0.5)
        sr = _parse_color(colors["specularReflectance"]) if "specularReflectance" in colors else np.ones(3, np.float32)
        entry = make_plastic(flags, dr, sr, eta, props.get("nonlinear", "false") == "true")
    elif typ == "roughplastic":         # src/bsdfs/roughplastic.cpp:190-232
        eta = _lookup_ior(props.get("intIOR"), "polypropylene") / _lookup_ior(props.get("extIOR"), "air")
        distr = props.get("distribution", "beckmann").lower()
        if distr not in ("beckmann", "ggx"):
            raise NotImplementedError(f"microfacet distribution '{distr}'")
        dr, tex_idx = diffuse_input(("diffuseReflectance",), 0.5)
        sr = _parse_color(colors["specularReflectance"]) if "specularReflectance" in colors else np.ones(3, np.float32)
        entry = make_roughplastic(flags, dr, sr, eta, float(props.get("alpha", 0.1)), 1 if distr == "ggx" else 0, props.get("nonlinear", "false") == "true", _TABLES)
    elif typ == "roughconductor":       # src/bsdfs/roughconductor.cpp:190-222
        ext = _lookup_ior(props.get("extEta"), "air")
        if "eta" in colors and "k" in colors:
            eta, k = _parse_color(colors["eta"]), _parse_color(colors["k"])
        elif props.get("material", "Cu").lower() == "none":
            eta, k = np.zeros(3, np.float32), np.ones(3, np.float32)
        else:
            raise NotImplementedError("roughconductor: named materials need Mitsuba's data/ior/*.spd files; give eta/k or material=none")
        if "alphaU" in props or "alphaV" in props:
            raise NotImplementedError("anisotropic roughness")
        distr = props.get("distribution", "beckmann").lower()
        if distr not in ("beckmann", "ggx"):
            raise NotImplementedError(f"microfacet distribution '{distr}'")
        sr = _parse_color(colors["specularReflectance"]) if "specularReflectance" in colors else np.ones(3, np.float32)
        entry = _make_bsdf(BSDF_ROUGHCONDUCTOR, flags, sr, (0, 0, 0), eta / ext, k / ext, float(props.get("alpha", 0.1)), 1 if distr == "ggx" else 0)
    elif typ == "conductor":            # src/bsdfs/conductor.cpp:152-176
        ext = _lookup_ior(props.get("extEta"), "air")
        if "eta" in colors and "k" in colors:
            eta, k = _parse_color(colors["eta"]), _parse_color(colors["k"])
        elif props.get("material", "Cu").lower() == "none":
            eta, k = np.zeros(3, np.float32), np.ones(3, np.float32)
        else:
            raise NotImplementedError("conductor: named materials need Mitsuba's data/ior/*.spd files; give eta/k or material=none")
        sr = _parse_color(colors["specularReflectance"]) if "specularReflectance" in colors else np.ones(3, np.float32)
        entry = _make_bsdf(BSDF_CONDUCTOR, flags, sr, (0, 0, 0), eta / ext, k / ext)
    else:
        raise NotImplementedError(f"BSDF '{typ}' is outside the hot-path scope")
    entry[25:26] = np.array([tex_idx], np.uint32).view(np.float32)

    def push(e, nd):
        idx = len(bsdf_table)
        bsdf_table.append(e)
        names.append(nd.attrib.get("id", f"bsdf{idx}"))
        if "id" in nd.attrib:
            by_id[nd.attrib["id"]] = idx
        return idx

    def set_flag(e, f):
        e = e.copy(); e[1:2] = (e[1:2].view(np.uint32) | np.uint32(f)).view(np.float32); return e

    levels = [(inner, entry)]
    for w in reversed(chain):           # from the innermost wrapper outwards
        e = levels[-1][1]
        wt = w.attrib["type"]
        if wt == "twosided":
            e = set_flag(e, BSDF_FLAG_TWOSIDED)
        elif wt == "mask":               # src/bsdfs/mask.cpp:63-66: constant opacity around the nested BSDF
            if any(c.tag == "texture" for c in w):
                raise NotImplementedError("mask: textured opacity")
            if int(e[:1].view(np.uint32)[0]) == BSDF_THINDIELECTRIC:
                raise NotImplementedError("mask around another null-type BSDF")
            opacity = np.full(3, 0.5, np.float32)
            for c in w:
                if c.tag in ("rgb", "srgb", "spectrum") and c.attrib.get("name") == "opacity":
                    opacity = _parse_color(c)
            e = set_flag(e, BSDF_FLAG_MASK); e[22:25] = opacity
        else:                            # bumpmap.cpp:119-137: one displacement texture
            tx = [c for c in w if c.tag == "texture"]
            if len(tx) != 1:
                raise ValueError("bumpmap: A displacement texture must be specified")
            if tx[0].attrib.get("type") == "bitmap" and "gamma" not in _prop_children(tx[0]):
                raise ValueError("When using a bitmap texture as a bump map, please explicitly specify the 'gamma' parameter of the bitmap plugin.")
            e = set_flag(e, BSDF_FLAG_BUMPMAP); e[26:27] = np.array([_parse_texture(tx[0], base) + 1], np.uint32).view(np.float32)
        levels.append((w, e))
    result = None
    for k, (nd, e) in enumerate(levels):
        if k == len(levels) - 1 or "id" in nd.attrib:
            result = push(e, nd)
    return result


def _parse_environment(node, base):
    """Root-level <emitter>: sunsky / sky / sun (baked, ppg_b200/sunsky.py) or envmap (src/emitters/envmap.cpp:102-190) -> envmap dict."""
    typ = node.attrib.get("type")
    pr = _prop_children(node)
    to_world = np.eye(4)
    for t in node.findall("transform"):
        if t.attrib.get("name") == "toWorld":
            to_world = _parse_transform(t)
    if typ in ("sunsky", "sky", "sun"):
        from . import sunsky
        for c in node:
            if c.attrib.get("name") == "albedo" and c.tag in ("rgb", "srgb", "spectrum"):
                pr["albedo"] = _parse_color(c)
        if typ == "sky":                  # src/emitters/sky.cpp: the sky dome alone -- what sunsky.cpp:122-158 nests with scale = skyScale
            pr = dict(pr, skyScale=pr.get("scale", "1.0"), sunScale="0")
        elif typ == "sun":                # src/emitters/sun.cpp:142-225: the sun disc alone, rasterised by the same (0,2)-sequence splat (sunRadiusScale = 0, a directional emitter, is not in scope)
            pr = dict(pr, sunScale=pr.get("scale", "1.0"), skyScale="0")
        img, _ = sunsky.bake(pr)
        scale = 1.0                       # sunsky.cpp:209-219 passes no scale to the nested envmap
    elif typ == "envmap":
        os.environ.setdefault("OPENCV_IO_ENABLE_OPENEXR", "1")
        img = load_bitmap(os.path.join(base, pr["filename"]), float(pr.get("gamma", 0)))
        if img.shape[2] == 1:
            img = np.repeat(img, 3, axis=2)
        scale = float(pr.get("scale", 1.0))
    else:
        raise NotImplementedError(f"emitter '{typ}' (scene-level emitters in scope: sunsky, sky, sun, envmap)")
    with np.errstate(over="ignore"):
        half = np.ascontiguousarray(img, np.float32).astype(np.float16)
    if not np.isfinite(half.astype(np.float32)).all():
        raise ValueError("The environment map contains an invalid floating point value (nan/inf) -- giving up.")
    return {"texels": half.view(np.uint16), "scale": scale, "world_to_env": np.linalg.inv(to_world[:3, :3]).astype(np.float32)}


def load_mitsuba_xml(path: str, film_size=None) -> SceneDesc:
    root = ET.parse(path).getroot()
    base = os.path.dirname(os.path.abspath(path))
    del _TABLES[:]
    del _TEXTURES[:]
    _TEX_CACHE.clear()
    integrator = {}
    inode = root.find("integrator")
    if inode is not None:
        if inode.attrib.get("type") != "guided_path":
            raise ValueError("only <integrator type=\"guided_path\"> is handled")
        integrator = _prop_children(inode)

    # sensor
    snode = root.find("sensor")
    if snode is None or snode.attrib.get("type") != "perspective":
        raise NotImplementedError("only the perspective sensor is in scope")
    sp = _prop_children(snode)
    cam_to_world = np.eye(4)
    for t in snode.findall("transform"):
        if t.attrib.get("name") == "toWorld":
            cam_to_world = _parse_transform(t)
    fnode = snode.find("film")
    fp = _prop_children(fnode) if fnode is not None else {}
    W, H = int(fp.get("width", 768)), int(fp.get("height", 576))
    if film_size is not None:
        W, H = film_size
    if fnode is not None:
        rf = fnode.find("rfilter")
        if rf is not None and rf.attrib.get("type") != "box":
            raise NotImplementedError("only the box reconstruction filter is in scope (all bundled scenes use it)")
    aspect = W / H
    fov = float(sp.get("fov", 0) or 0)
    if "fov" not in sp:
        f = sp.get("focalLength", "50mm").rstrip("m")
        fov_diag = 2 * 180 / math.pi * math.atan(math.sqrt(36 * 36 + 24 * 24) / (2 * float(f)))
        diagonal = 2 * math.tan(0.5 * math.radians(fov_diag))
        width = diagonal / math.sqrt(1.0 + 1.0 / (aspect * aspect))
        xfov = math.degrees(2 * math.atan(width * 0.5))
    else:
        axis = sp.get("fovAxis", "x").lower()
        if axis == "smaller":
            axis = "y" if aspect > 1 else "x"
        elif axis == "larger":
            axis = "x" if aspect > 1 else "y"
        if axis == "x":
            xfov = fov
        elif axis == "y":
            xfov = math.degrees(2 * math.atan(math.tan(0.5 * math.radians(fov)) * aspect))
        elif axis == "diagonal":
            diagonal = 2 * math.tan(0.5 * math.radians(fov))
            width = diagonal / math.sqrt(1.0 + 1.0 / (aspect * aspect))
            xfov = math.degrees(2 * math.atan(width * 0.5))
        else:
            raise ValueError("The 'fovAxis' parameter must be set to one of 'smaller', 'larger', 'diagonal', 'x', or 'y'!")
    near = float(sp.get("nearClip", 1e-2)); far = float(sp.get("farClip", 1e4))

    bsdf_table, names, by_id = [], [], {}
    for b in root.findall("bsdf"):
        _parse_bsdf(b, bsdf_table, names, by_id, base)
    envmap = None
    for e in root.findall("emitter"):
        if envmap is not None:
            raise NotImplementedError("more than one scene-level emitter")
        envmap = _parse_environment(e, base)

    P_all, N_all, UV_all, I_all, TS_all, shapes, radiance = [], [], [], [], [], [], []
    voff = 0; toff = 0
    black = grey = None
    sphere_list = []
    for sh in root.findall("shape"):
        typ = sh.attrib["type"]
        props = _prop_children(sh)
        to_world = np.eye(4)
        for t in sh.findall("transform"):
            if t.attrib.get("name") == "toWorld":
                to_world = _parse_transform(t)
        if typ == "obj":
            P, N, UV, I = _load_obj(os.path.join(base, props["filename"]), to_world,
                                    props.get("faceNormals", "false") == "true", props.get("flipNormals", "false") == "true")
        elif typ in ("ply", "serialized"):
            if "maxSmoothAngle" in props:
                raise NotImplementedError(f"{typ}: maxSmoothAngle")
            fn, fl = props.get("faceNormals", "false") == "true", props.get("flipNormals", "false") == "true"
            mesh = os.path.join(base, props["filename"])
            P, N, UV, I = _load_ply(mesh, to_world, fn, fl) if typ == "ply" else _load_serialized(mesh, int(props.get("shapeIndex", 0)), to_world, fn, fl)
        elif typ == "cube":
            P, N, UV, I = _cube(to_world, props.get("flipNormals", "false") == "true")
        elif typ == "rectangle":
            P, N, UV, I = _rectangle(to_world)
        elif typ == "sphere":           # src/shapes/sphere.cpp:107-132: toWorld's scale folds into the radius
            ctr = np.zeros(3)
            for c in sh.findall("point"):
                if c.attrib.get("name") == "center":
                    ctr = np.array([float(c.attrib.get(k, 0)) for k in "xyz"])
            radius = float(props.get("radius", 1.0)) * np.linalg.norm(to_world[:3, 0])
            rot = to_world[:3, :3] / np.linalg.norm(to_world[:3, 0])
            if not np.allclose(rot, np.eye(3), atol=1e-6):
                raise NotImplementedError("sphere: toWorld must be translation + uniform scale")
            center = to_world[:3, :3] @ ctr + to_world[:3, 3]
            sphere_list.append([center, radius, len(shapes), props.get("flipNormals", "false") == "true"])
            P = np.zeros((0, 3), np.float32); N = None; UV = None; I = np.zeros((0, 3), np.uint32)
        else:
            raise NotImplementedError(f"shape '{typ}' (shapes in scope: obj, ply, serialized, rectangle, cube, sphere)")
        bsdf_idx = None
        ref = sh.find("ref")
        if ref is not None:
            bsdf_idx = by_id[ref.attrib["id"]]
        bn = sh.find("bsdf")
        if bn is not None:
            bsdf_idx = _parse_bsdf(bn, bsdf_table, names, by_id, base)
        em = sh.find("emitter")
        em_idx = -1
        if em is not None:
            if em.attrib.get("type") != "area":
                raise NotImplementedError("only area emitters on shapes")
            rad = np.ones(3, np.float32)
            for c in em:
                if c.attrib.get("name") == "radiance":
                    rad = _parse_color(c, is_emitter=True)
            em_idx = len(radiance); radiance.append(rad)
        if bsdf_idx is None:
            if em_idx >= 0:
                if black is None:
                    black = len(bsdf_table); bsdf_table.append(_make_bsdf(BSDF_DIFFUSE, 0, np.zeros(3, np.float32))); names.append("__black")
                bsdf_idx = black
            else:
                if grey is None:
                    grey = len(bsdf_table); bsdf_table.append(_make_bsdf(BSDF_DIFFUSE, 0, np.full(3, 0.5, np.float32))); names.append("__grey")
                bsdf_idx = grey
        shapes.append([toff, len(I), bsdf_idx, em_idx, 0 if N is None else 1, 0 if UV is None else 1, 0, 0])
        P_all.append(P)
        N_all.append(N if N is not None else np.zeros_like(P))
        UV_all.append(UV if UV is not None else np.zeros((len(P), 2), np.float32))
        I_all.append(I + voff)
        TS_all.append(np.full(len(I), len(shapes) - 1, np.uint32))
        voff += len(P); toff += len(I)

    textures = np.zeros(len(_TEXTURES), TEXTURE_DTYPE); off = 0
    for i, (m, tx) in enumerate(_TEXTURES):
        textures[i] = (m["width"], m["height"], m["channels"], m["wrap_u"], m["wrap_v"], m["uv_scale"], m["uv_offset"], 0, off)
        off += len(tx)
    texels = np.concatenate([tx for _, tx in _TEXTURES]) if _TEXTURES else np.zeros(0, np.uint16)
    P = np.concatenate(P_all).astype(np.float32)
    aabb_min = P.min(axis=0).astype(np.float64); aabb_max = P.max(axis=0).astype(np.float64)
    cam_pos = cam_to_world[:3, 3]
    aabb_min = np.minimum(aabb_min, cam_pos); aabb_max = np.maximum(aabb_max, cam_pos)
    for c, r, _, _ in sphere_list:         # Sphere::getAABB, sphere.cpp:152-157
        aabb_min = np.minimum(aabb_min, np.float32(c) - np.float32(r)); aabb_max = np.maximum(aabb_max, np.float32(c) + np.float32(r))
    return SceneDesc(
        positions=P, normals=np.concatenate(N_all).astype(np.float32), uvs=np.concatenate(UV_all).astype(np.float32),
        indices=np.concatenate(I_all).astype(np.uint32), triangle_shape=np.concatenate(TS_all).astype(np.uint32),
        shapes=np.asarray(shapes, np.int64).astype(np.int32), bsdfs=np.asarray(bsdf_table, np.float32).reshape(-1, 28),
        bsdf_tables=np.asarray(_TABLES, np.float32).reshape(-1, 100),
        spheres=np.asarray([make_sphere(c, r, si, fl) for c, r, si, fl in sphere_list], np.float32).reshape(-1, 6),
        area_radiance=np.asarray(radiance, np.float32).reshape(-1, 3), cam_to_world=cam_to_world.astype(np.float32),
        x_fov_deg=float(xfov), near_clip=near, far_clip=far, film_width=W, film_height=H,
        aabb_min=aabb_min.astype(np.float32), aabb_max=aabb_max.astype(np.float32), integrator=integrator, bsdf_names=names,
        textures=textures, texels=texels, envmap=envmap)
