"""Rough-transmittance tables of the reference's roughplastic BSDF, reduced on the host to what the device needs.

Mitsuba ships data/microfacet/{beckmann,ggx}.dat (header "MTS_TRANSMITTANCE", 3 x u64 sample counts eta/alpha/theta,
4 x f32 ranges, then for each of 2*eta blocks x alpha: theta transmittance samples + 1 diffuse transmittance) and
RoughTransmittance (src/bsdfs/rtrans.h:75-151 loader, :292-346 setEta, :353-388 setAlpha, :183-283 eval/evalDiffuse)
reduces them for a material with constant eta and alpha to a 100-entry 1-D table over cos(theta)^(1/4) plus scalars.
This module restates that reduction (Catmull-Rom style cubic tensor interpolation, src/libcore/spline.cpp:22-60, 236-304)
in float32; the per-material result travels inside the scene description (ppg_scene_desc.bsdf_tables), so the .dat files
are only needed where scenes are converted."""
from __future__ import annotations

import os
import struct

import numpy as np

_DATA_DIR = "/root/reference/mitsuba/data/microfacet"
_cache = {}
f32 = np.float32


def load_table(name: str, data_dir: str = _DATA_DIR):
    if name in _cache:
        return _cache[name]
    b = open(os.path.join(data_dir, name + ".dat"), "rb").read()
    hdr = b"MTS_TRANSMITTANCE"
    assert b[:len(hdr)] == hdr, "invalid transmittance data file"
    p = len(hdr)
    n_eta, n_alpha, n_theta = struct.unpack("<QQQ", b[p:p + 24]); p += 24
    eta_min, eta_max, alpha_min, alpha_max = struct.unpack("<ffff", b[p:p + 16]); p += 16
    raw = np.frombuffer(b[p:], dtype="<f4").reshape(2 * n_eta, n_alpha, n_theta + 1)
    t = dict(trans=np.ascontiguousarray(raw[:, :, :n_theta]), diff=np.ascontiguousarray(raw[:, :, n_theta]), n_eta=n_eta, n_alpha=n_alpha,
             n_theta=n_theta, eta_min=f32(eta_min), eta_max=f32(eta_max), alpha_min=f32(alpha_min), alpha_max=f32(alpha_max))
    _cache[name] = t
    return t


def _weights(x, size):
    """Node weights of evalCubicInterp{2,3}D for one dimension (spline.cpp:242-287), x in [0,1]."""
    x = f32(x)
    if not (x >= 0 and x <= 1):
        return None, None
    t = f32(x * f32(size - 1))
    knot = min(int(t), size - 2)
    t = f32(t - f32(knot))
    t2 = f32(t * t); t3 = f32(t2 * t)
    w = [f32(0), f32(2 * t3 - 3 * t2 + 1), f32(-2 * t3 + 3 * t2), f32(0)]
    d0 = f32(t3 - 2 * t2 + t); d1 = f32(t3 - t2)
    if knot > 0:
        w[2] = f32(w[2] + f32(0.5) * d0); w[0] = f32(w[0] - f32(0.5) * d0)
    else:
        w[2] = f32(w[2] + d0); w[1] = f32(w[1] - d0)
    if knot + 2 < size:
        w[3] = f32(w[3] + f32(0.5) * d1); w[1] = f32(w[1] - f32(0.5) * d1)
    else:
        w[2] = f32(w[2] + d1); w[1] = f32(w[1] - d1)
    return knot, w


def _interp(values, coords):
    """Tensor-product cubic interpolation; `values` indexed [slowest..fastest], `coords` given fastest-first like Mitsuba's Point."""
    dims = values.shape[::-1]          # fastest first
    ks, ws = [], []
    for x, n in zip(coords, dims):
        k, w = _weights(x, n)
        if k is None:
            return f32(0)
        ks.append(k); ws.append(w)
    result = f32(0)
    nd = len(dims)
    if nd == 1:
        for x in range(-1, 3):
            w = ws[0][x + 1]
            if w == 0: continue
            result = f32(result + values[ks[0] + x] * w)
    elif nd == 2:
        for y in range(-1, 3):
            wy = ws[1][y + 1]
            for x in range(-1, 3):
                wxy = f32(ws[0][x + 1] * wy)
                if wxy == 0: continue
                result = f32(result + values[ks[1] + y, ks[0] + x] * wxy)
    else:
        for z in range(-1, 3):
            wz = ws[2][z + 1]
            for y in range(-1, 3):
                wyz = f32(ws[1][y + 1] * wz)
                for x in range(-1, 3):
                    wxyz = f32(ws[0][x + 1] * wyz)
                    if wxyz == 0: continue
                    result = f32(result + values[ks[2] + z, ks[1] + y, ks[0] + x] * wxyz)
    return result


def _set_eta(t, eta):
    """RoughTransmittance::setEta: 3-D (eta, alpha, theta) -> 2-D (alpha, theta) and the diffuse row (alpha)."""
    eta = f32(eta)
    trans, diff = t["trans"][:t["n_eta"]], t["diff"][:t["n_eta"]]
    if eta < 1:
        trans, diff = t["trans"][t["n_eta"]:], t["diff"][t["n_eta"]:]
        eta = f32(1) / eta
    if eta < t["eta_min"]:
        eta = t["eta_min"]
    warped_eta = f32(np.power(f32((eta - t["eta_min"]) / (t["eta_max"] - t["eta_min"])), f32(0.25)))
    na, nt = t["n_alpha"], t["n_theta"]
    d_alpha = f32(1.0) / f32(na - 1); d_theta = f32(1.0) / f32(nt - 1)
    new_trans = np.zeros((na, nt), f32); new_diff = np.zeros(na, f32)
    for i in range(na):
        for j in range(nt):
            new_trans[i, j] = _interp(trans, (f32(j * d_theta), f32(i * d_alpha), warped_eta))
        new_diff[i] = _interp(diff, (f32(i * d_alpha), warped_eta))
    return new_trans, new_diff


def _warped_alpha(t, alpha):
    return f32(np.power(f32((f32(alpha) - t["alpha_min"]) / (t["alpha_max"] - t["alpha_min"])), f32(0.25)))


def reduce_for_material(distribution: str, eta: float, alpha: float, data_dir: str = _DATA_DIR):
    """Returns (lut[100] float32 over cos(theta)^(1/4) of the EXTERNAL transmittance at (eta, alpha), fdr_int) where
    fdr_int = 1 - internal diffuse transmittance (eta -> 1/eta) evaluated at alpha (roughplastic.cpp:276-287, 364-366)."""
    t = load_table(distribution, data_dir)
    if not (t["alpha_min"] <= alpha <= t["alpha_max"]):
        raise ValueError(f"roughness alpha={alpha} outside the tabulated range [{t['alpha_min']}, {t['alpha_max']}]")
    wa = _warped_alpha(t, alpha)
    ext_trans, _ = _set_eta(t, eta)
    nt = t["n_theta"]
    d_theta = f32(1.0) / f32(nt - 1)
    lut = np.array([_interp(ext_trans, (f32(i * d_theta), wa)) for i in range(nt)], f32)      # setAlpha
    _, int_diff = _set_eta(t, 1.0 / eta)
    int_diffuse = min(f32(1), max(f32(0), _interp(int_diff, (wa,))))                             # evalDiffuse(alpha), eta fixed
    return lut, float(f32(1) - int_diffuse)
