import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cbox(size=None, improved=False):
    from ppg_b200.scene import SceneDesc
    sc = SceneDesc.load(os.path.join(ROOT, "scenes", "cbox-improved.npz" if improved else "cbox.npz"))
    if size is not None:
        sc = sc.with_film(size, size)
    return sc


def relmse(img, ref):
    img = np.asarray(img, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + 1e-3)))


def gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
