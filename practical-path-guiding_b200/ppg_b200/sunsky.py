"""Host-side bake of Mitsuba's ``sunsky`` emitter into the lat-long environment map the hot path evaluates.

The reference never evaluates the sun/sky models while rendering: ``SunSkyEmitter`` (src/emitters/sunsky.cpp:100-225)
rasterises the Hosek-Wilkie sky (src/emitters/sky.cpp:219-260, 418-456; model code src/emitters/sunsky/skymodel.cpp,
coefficient tables src/emitters/sunsky/skymodeldata.h) into a ``resolution x resolution/2`` RGB bitmap, splats the
sun disc into it with a (0,2)-sequence (sunsky.cpp:160-207; sun position "Computing the Solar Vector" and Preetham's
sun radiance, src/emitters/sunsky/sunmodel.h:119-376) and hands the bitmap to the ``envmap`` plugin, which stores it in
half precision (src/emitters/envmap.cpp:102-103) and looks it up bilinearly on ray misses (:380-410).  This module
restates that scene-preparation step; what travels to the device is the baked half-precision map
(``ppg_scene_desc.envmap``).  The coefficient tables are published model data (Hosek & Wilkie 2012): they are parsed out
of the Mitsuba tree at conversion time and never stored in this repository.
"""
from __future__ import annotations

import math
import os
import re

import numpy as np

_SKY_DATA_H = os.environ.get("PPG_MITSUBA_SRC", "/root/reference/mitsuba/src") + "/emitters/sunsky/skymodeldata.h"
SUN_APP_RADIUS = 0.5358          # sunsky.cpp:34
f32 = np.float32
_tables = None


def _sky_tables(path=None):
    """datasetsRGB[3] (9 x 6 x 10 x 2 doubles each) and datasetsRGBRad[3] (6 x 10 x 2) of skymodeldata.h."""
    global _tables
    if _tables is None:
        p = path or _SKY_DATA_H
        if not os.path.exists(p):
            raise RuntimeError(f"sunsky: the Hosek-Wilkie coefficient tables ({p}) are needed to bake the emitter; "
                               "set PPG_MITSUBA_SRC to a Mitsuba 0.5 'src' directory")
        text = open(p, "r", errors="replace").read()
        out = {}
        for name in ("datasetRGB1", "datasetRGB2", "datasetRGB3", "datasetRGBRad1", "datasetRGBRad2", "datasetRGBRad3"):
            m = re.search(r"double\s+" + name + r"\s*\[\s*\]\s*=\s*\{(.*?)\};", text, re.S)
            body = re.sub(r"//[^\n]*", "", m.group(1))
            out[name] = np.array([float(t) for t in re.split(r"[,\s]+", body.strip()) if t], np.float64)
        _tables = out
    return _tables


def _bezier5(mat, stride, x):
    """quintic Bezier in the cube-rooted normalised solar elevation (skymodel.cpp:76-160)."""
    return ((1 - x) ** 5 * mat[0 * stride:0 * stride + stride] + 5 * (1 - x) ** 4 * x * mat[1 * stride:2 * stride]
            + 10 * (1 - x) ** 3 * x ** 2 * mat[2 * stride:3 * stride] + 10 * (1 - x) ** 2 * x ** 3 * mat[3 * stride:4 * stride]
            + 5 * (1 - x) * x ** 4 * mat[4 * stride:5 * stride] + x ** 5 * mat[5 * stride:6 * stride])


def cook_configuration(dataset, dataset_rad, turbidity, albedo, solar_elevation):
    """ArHosekSkyModel_CookConfiguration + CookRadianceConfiguration (skymodel.cpp:76-211): 9 coefficients + 1 radiance."""
    it = int(turbidity)
    assert 1 <= it <= 10
    rem = turbidity - it
    x = (solar_elevation / (math.pi / 2.0)) ** (1.0 / 3.0)
    cfg = np.zeros(9); rad = 0.0
    for alb_w, alb_off in ((1.0 - albedo, 0), (albedo, 1)):
        for t_w, t_idx in ((1.0 - rem, it - 1), (rem, it)):
            if t_idx >= 10:          # int_turbidity == 10: the upper turbidity terms are skipped
                continue
            cfg += alb_w * t_w * _bezier5(dataset[9 * 6 * 10 * alb_off + 9 * 6 * t_idx:], 9, x)
            rad += alb_w * t_w * float(_bezier5(dataset_rad[6 * 10 * alb_off + 6 * t_idx:], 1, x)[0])
    return cfg, rad


def sky_radiance_internal(cfg, theta, gamma):
    """ArHosekSkyModel_GetRadianceInternal (skymodel.cpp:213-226), vectorised over theta/gamma."""
    cg = np.cos(gamma); ct = np.cos(theta)
    expM = np.exp(cfg[4] * gamma)
    rayM = cg * cg
    mieM = (1.0 + cg * cg) / np.power(1.0 + cfg[8] * cfg[8] - 2.0 * cfg[8] * cg, 1.5)
    zenith = np.sqrt(ct)
    return (1.0 + cfg[0] * np.exp(cfg[1] / (ct + 0.01))) * (cfg[2] + cfg[3] * expM + cfg[5] * rayM + cfg[6] * mieM + cfg[7] * zenith)


def sun_coordinates(props):
    """computeSunCoordinates(props) -> (elevation [zenith angle], azimuth), sunmodel.h:119-231 (date/time/location branch)."""
    if "sunDirection" in props:
        raise NotImplementedError("sunsky: explicit sunDirection")
    lat = float(f32(props.get("latitude", 35.6894))); lon = float(f32(props.get("longitude", 139.6917))); tz = float(props.get("timezone", 9))
    year = int(props.get("year", 2010)); day = int(props.get("day", 10)); month = int(props.get("month", 7))
    hour = float(props.get("hour", 15.0)); minute = float(props.get("minute", 0.0)); second = float(props.get("second", 0.0))
    dec_hours = hour - tz + (minute + second / 60.0) / 60.0
    aux1 = int((month - 14) / 12)                       # C integer division truncates towards zero
    aux2 = (int((1461 * (year + 4800 + aux1)) / 4) + int((367 * (month - 2 - 12 * aux1)) / 12)
            - int((3 * int((year + 4900 + aux1) / 100)) / 4) + day - 32075)
    julian = float(aux2) - 0.5 + dec_hours / 24.0
    elapsed = julian - 2451545.0
    omega = 2.1429 - 0.0010394594 * elapsed
    mean_long = 4.8950630 + 0.017202791698 * elapsed
    anomaly = 6.2400600 + 0.0172019699 * elapsed
    ecl_long = mean_long + 0.03341607 * math.sin(anomaly) + 0.00034894 * math.sin(2 * anomaly) - 0.0001134 - 0.0000203 * math.sin(omega)
    ecl_obl = 0.4090928 - 6.2140e-9 * elapsed + 0.0000396 * math.cos(omega)
    s_el = math.sin(ecl_long)
    dY = math.cos(ecl_obl) * s_el; dX = math.cos(ecl_long)
    ra = math.atan2(dY, dX)
    if ra < 0:
        ra += 2 * math.pi
    decl = math.asin(math.sin(ecl_obl) * s_el)
    gmst = 6.6974243242 + 0.0657098283 * elapsed + dec_hours
    lmst = float(f32(f32(gmst * 15 + lon) * f32(math.pi / 180.0)))            # degToRad((Float) ...)
    lat_r = float(f32(f32(lat) * f32(math.pi / 180.0)))
    hour_angle = lmst - ra
    elevation = math.acos(math.cos(lat_r) * math.cos(hour_angle) * math.cos(decl) + math.sin(decl) * math.sin(lat_r))
    dY = -math.sin(hour_angle)
    dX = math.tan(decl) * math.cos(lat_r) - math.sin(lat_r) * math.cos(hour_angle)
    azimuth = math.atan2(dY, dX)
    if azimuth < 0:
        azimuth += 2 * math.pi
    elevation += (6371.01 / 149597890.0) * math.sin(elevation)
    return float(f32(elevation)), float(f32(azimuth))


def _interp_eval(l, v, lam):
    """InterpolatedSpectrum::eval (src/libcore/spectrum.cpp:688-714) incl. its mirrored lerp inside a knot interval."""
    if lam < l[0] or lam > l[-1]:
        return 0.0
    i1 = int(np.searchsorted(l, lam, side="left")); i2 = int(np.searchsorted(l, lam, side="right"))
    if i1 == i2:
        a, b, fa, fb = l[i1 - 1], l[i1], v[i1 - 1], v[i1]
        t = (lam - a) / (b - a)
        return (1 - t) * fb + t * fa
    return float(v[i1])


# sunmodel.h:237-300 (data "lifted from MI" by Preetham et al.; units cm^-1 / W m^-2 nm^-1 sr^-1)
_K_O_L = [300, 305, 310, 315, 320, 325, 330, 335, 340, 345, 350, 355, 445, 450, 455, 460, 465, 470, 475, 480, 485, 490, 495, 500, 505, 510, 515, 520, 525, 530,
          535, 540, 545, 550, 555, 560, 565, 570, 575, 580, 585, 590, 595, 600, 605, 610, 620, 630, 640, 650, 660, 670, 680, 690, 700, 710, 720, 730, 740, 750,
          760, 770, 780, 790]
_K_O_A = [10.0, 4.8, 2.7, 1.35, .8, .380, .160, .075, .04, .019, .007, .0, .003, .003, .004, .006, .008, .009, .012, .014, .017, .021, .025, .03, .035, .04, .045,
          .048, .057, .063, .07, .075, .08, .085, .095, .103, .110, .12, .122, .12, .118, .115, .12, .125, .130, .12, .105, .09, .079, .067, .057, .048, .036,
          .028, .023, .018, .014, .011, .010, .009, .007, .004, .0, .0]          # the reference passes the first 64 of its 65 entries
_K_G_L = [759, 760, 770, 771]; _K_G_A = [0, 3.0, 0.210, 0]
_K_WA_L = [689, 690, 700, 710, 720, 730, 740, 750, 760, 770, 780, 790, 800]
_K_WA_A = [0, 0.160e-1, 0.240e-1, 0.125e-1, 0.100e+1, 0.870, 0.610e-1, 0.100e-2, 0.100e-4, 0.100e-4, 0.600e-3, 0.175e-1, 0.360e-1]
_SOL_L = list(range(380, 751, 10))
_SOL_A = [16559.0, 16233.7, 21127.5, 25888.2, 25829.1, 24232.3, 26760.5, 29658.3, 30545.4, 30057.5, 30663.7, 28830.4, 28712.1, 27825.0, 27100.6, 27233.6, 26361.3,
          25503.8, 25060.2, 25311.6, 25355.9, 25134.2, 24631.5, 24173.2, 23685.3, 23212.1, 22827.7, 22339.8, 21970.2, 21526.7, 21097.9, 20728.3, 20240.4, 19870.8,
          19427.2, 19072.4, 18628.9, 18259.2]


def sun_radiance(theta, turbidity):
    """computeSunRadiance (sunmodel.h:302-376): Preetham's attenuated solar spectrum -> RGB (fromContinuousSpectrum, RGB build)."""
    from .scene import continuous_to_rgb
    ko = (np.array(_K_O_L, float), np.array(_K_O_A[:64], float)); kg = (np.array(_K_G_L, float), np.array(_K_G_A, float))
    kwa = (np.array(_K_WA_L, float), np.array(_K_WA_A, float)); sol = (np.array(_SOL_L, float), np.array(_SOL_A, float))
    beta = 0.04608365822050 * turbidity - 0.04586025928522
    m = 1.0 / (math.cos(theta) + 0.15 * math.pow(93.885 - theta / math.pi * 180.0, -1.253))
    lam = np.arange(350.0, 801.0, 5.0); data = np.zeros(91)
    for i, l in enumerate(lam):
        tauR = math.exp(-m * 0.008735 * math.pow(l / 1000.0, -4.08))
        tauA = math.exp(-m * beta * math.pow(l / 1000.0, -1.3))
        tauO = math.exp(-m * _interp_eval(*ko, l) * 0.35)
        g = _interp_eval(*kg, l)
        tauG = math.exp(-1.41 * g * m / math.pow(1 + 118.93 * g * m, 0.45))
        wa = _interp_eval(*kwa, l)
        tauWA = math.exp(-0.2385 * wa * 2.0 * m / math.pow(1 + 20.07 * wa * 2.0 * m, 0.45))
        data[i] = _interp_eval(*sol, l) * tauR * tauA * tauO * tauG * tauWA
    return np.maximum(continuous_to_rgb(lam, data), 0.0)


def _radical_inverse2(n):
    n = np.asarray(n, np.uint32)
    n = ((n & np.uint32(0xffff)) << np.uint32(16)) | (n >> np.uint32(16))
    n = ((n & np.uint32(0x00ff00ff)) << np.uint32(8)) | ((n & np.uint32(0xff00ff00)) >> np.uint32(8))
    n = ((n & np.uint32(0x0f0f0f0f)) << np.uint32(4)) | ((n & np.uint32(0xf0f0f0f0)) >> np.uint32(4))
    n = ((n & np.uint32(0x33333333)) << np.uint32(2)) | ((n & np.uint32(0xcccccccc)) >> np.uint32(2))
    n = ((n & np.uint32(0x55555555)) << np.uint32(1)) | ((n & np.uint32(0xaaaaaaaa)) >> np.uint32(1))
    return (n >> np.uint32(8)).astype(np.float32) / f32(1 << 24)            # qmc.h:43-59


def _sobol2(n):
    n = np.asarray(n, np.uint64).copy(); scr = np.zeros_like(n); v = np.uint64(1 << 31)
    while n.any():                                                           # qmc.h:82-87
        scr = np.where((n & np.uint64(1)) != 0, scr ^ v, scr)
        n >>= np.uint64(1); v ^= v >> np.uint64(1)
    return (scr.astype(np.float32) / f32(2.0 ** 32)).astype(np.float32)


def _coordinate_system(a):
    """coordinateSystem(a, b, c), src/libcore/util.cpp:592-601."""
    if abs(a[0]) > abs(a[1]):
        inv = 1.0 / math.sqrt(a[0] * a[0] + a[2] * a[2]); c = np.array([a[2] * inv, 0.0, -a[0] * inv])
    else:
        inv = 1.0 / math.sqrt(a[1] * a[1] + a[2] * a[2]); c = np.array([0.0, a[2] * inv, -a[1] * inv])
    return np.cross(c, a), c


def bake(props: dict, data_path=None):
    """SunSkyEmitter::SunSkyEmitter (sunsky.cpp:100-225): returns the (H, W, 3) float32 lat-long radiance map, BEFORE the
    envmap plugin's half-precision quantisation.  ``props``: the emitter's XML properties as strings."""
    T = _sky_tables(data_path)
    scale = float(props.get("scale", 1.0)); sun_scale = float(props.get("sunScale", scale)); sky_scale = float(props.get("skyScale", scale))
    sun_radius_scale = float(props.get("sunRadiusScale", 1.0))
    turbidity = float(props.get("turbidity", 3.0)); stretch = float(props.get("stretch", 1.0))
    if props.get("extend", "false") == "true":
        raise NotImplementedError("sunsky: extend=true")
    if not 1 <= turbidity <= 10 or not 1 <= stretch <= 2:
        raise ValueError("The turbidity parameter must be in [1,10] and stretch in [1,2]")
    albedo = np.full(3, 0.2) if "albedo" not in props else np.asarray(props["albedo"], float)
    res = int(props.get("resolution", 512)); W, H = res, res // 2
    sun_el, sun_az = sun_coordinates(props)
    sun_elevation = 0.5 * math.pi - sun_el
    if sun_elevation < 0:
        raise ValueError("The sun is below the horizon -- this is not supported by the sky model.")
    theta = ((np.arange(H) + 0.5) * (math.pi / H))[:, None] / stretch            # sky.cpp:420: theta = elevation / stretch
    phi = ((np.arange(W) + 0.5) * (2 * math.pi / W))[None, :]
    img = np.zeros((H, W, 3), np.float64)
    up = np.cos(theta) > 0
    th = np.where(up, theta, 0.0) + 0 * phi
    cos_gamma = np.cos(th) * math.cos(sun_el) + np.sin(th) * math.sin(sun_el) * np.cos(phi - sun_az)
    gamma = np.arccos(np.clip(cos_gamma, -1.0, 1.0))
    for c in range(3):
        cfg, rad = cook_configuration(T[f"datasetRGB{c + 1}"], T[f"datasetRGBRad{c + 1}"], turbidity, float(albedo[c]), sun_elevation)
        val = sky_radiance_internal(cfg, th, gamma) * rad / 106.856980
        img[:, :, c] = np.where(up, np.maximum(val, 0.0) * sky_scale, 0.0)
    img = img.astype(np.float32)
    # ---- the sun disc, splatted with a (0,2)-sequence (sunsky.cpp:160-207)
    sun_rad = sun_radiance(sun_el, turbidity).astype(np.float64) * sun_scale
    sun_el_s = sun_el * stretch
    n = np.array([math.sin(sun_az) * math.sin(sun_el_s), math.cos(sun_el_s), -math.cos(sun_az) * math.sin(sun_el_s)])   # toSphere, sunmodel.h:90-97
    s, t = _coordinate_system(n)
    theta_sun = float(f32(f32(SUN_APP_RADIUS * 0.5) * f32(math.pi / 180.0)))
    if sun_radius_scale == 0:
        raise NotImplementedError("sunsky: sunRadiusScale=0 (directional sun emitter)")
    pixel_count = res * res // 2
    cos_theta = f32(np.cos(f32(theta_sun * sun_radius_scale)))
    covered = f32(0.5) * (f32(1) - cos_theta)
    n_samples = int(max(f32(100), f32(pixel_count) * covered * f32(1000)))
    value = sun_rad * (2 * math.pi * (1 - math.cos(theta_sun))) * float(W * H) / (2 * math.pi * math.pi * n_samples)
    idx = np.arange(n_samples, dtype=np.uint32)
    sx = _radical_inverse2(idx).astype(np.float64); sy = _sobol2(idx).astype(np.float64)
    ct = (1 - sx) + sx * float(cos_theta); st = np.sqrt(np.maximum(0.0, 1 - ct * ct))       # warp::squareToUniformCone
    local = np.stack([np.cos(2 * math.pi * sy) * st, np.sin(2 * math.pi * sy) * st, ct], axis=1)
    d = local[:, :1] * s[None] + local[:, 1:2] * t[None] + local[:, 2:] * n[None]
    sin_theta = np.sqrt(np.maximum(0.0, 1 - d[:, 1] ** 2))
    az = np.arctan2(d[:, 0], -d[:, 2]); az = np.where(az < 0, az + 2 * math.pi, az)
    el = np.arccos(np.clip(d[:, 1], -1, 1))
    px = np.clip((az * (W / (2 * math.pi))).astype(np.int64), 0, W - 1); py = np.clip((el * (H / math.pi)).astype(np.int64), 0, H - 1)
    acc = img.astype(np.float64)
    np.add.at(acc, (py, px), value[None, :] / np.maximum(1e-3, sin_theta)[:, None])
    return acc.astype(np.float32), dict(sun_elevation=sun_el, sun_azimuth=sun_az, n_samples=n_samples, sun_radiance=sun_rad)
