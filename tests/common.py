import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED_OVERRIDE = None     # set by conftest for tests marked `seeds3`


def with_seed(props):
    """props with the seed of the current `seeds3` round (unless the test sets one itself)."""
    if SEED_OVERRIDE is not None and "seed" not in props:
        return dict(props, seed=str(SEED_OVERRIDE))
    return props


def load_cbox(size=None, improved=False):
    from ppg_b200.scene import SceneDesc
    sc = SceneDesc.load(os.path.join(ROOT, "scenes", "cbox-improved.npz" if improved else "cbox.npz"))
    if size is not None:
        sc = sc.with_film(size, size)
    return sc


def load_fixture_scene(name, size=None):  # size None or 0: the XML film size
    """scenes/<name>.npz written by tools/make_fixtures.py."""
    from ppg_b200.scene import SceneDesc
    sc = SceneDesc.load(os.path.join(ROOT, "scenes", name + ".npz"))
    return sc.with_film(size, size) if size else sc


def relmse(img, ref):
    img = np.asarray(img, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + 1e-3)))


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def cbox_with_sphere(size=128, subdiv=3, center=(185.0, 240.0, 169.0), radius=75.0, smooth=True):
    """CBOX plus a tessellated diffuse sphere (20 * 4^subdiv triangles, smooth vertex normals): more than 64 triangles,
    so the CUDA path intersects through the BVH walk instead of the lock-step tiny-scene test."""
    import copy
    sc = copy.copy(load_cbox(size))
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    v = [np.array(p, np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache = {}; nf = []
        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                m = v[a] + v[b]; v.append(m / np.linalg.norm(m)); cache[k] = len(v) - 1
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    V = np.array(v); F = np.array(f, np.uint32)
    P = (np.array(center) + radius * V).astype(np.float32); N = V.astype(np.float32)
    nv, nt, ns = len(sc.positions), len(sc.indices), len(sc.shapes)
    sc.positions = np.concatenate([sc.positions, P]); sc.normals = np.concatenate([sc.normals, N if smooth else np.zeros_like(N)])
    sc.uvs = np.concatenate([sc.uvs, np.zeros((len(P), 2), np.float32)])
    sc.indices = np.concatenate([sc.indices, F + nv]).astype(np.uint32)
    sc.triangle_shape = np.concatenate([sc.triangle_shape, np.full(len(F), ns, np.uint32)])
    sc.shapes = np.concatenate([sc.shapes, np.array([[nt, len(F), 1, -1, 1 if smooth else 0, 0, 0, 0]], np.int32)])   # bsdf 1 = "white"
    return sc
