"""The product's CUDA device functions (practical-path-guiding_b200/csrc/ppg_device.cuh) compiled for the HOST with g++ (tests/hostdev/: CUDA's own headers supply
float3 / float4, the intrinsics get host meanings) and compared, function by function and bit for bit, with the reference's own code compiled verbatim
(oracle/_ref/libmicrofacet_ref.so) -- or with the oracle's restatement where the reference tree was absent at build time.  The kernels are built with -fmad=false
-prec-div=true -prec-sqrt=true, this harness with -ffp-contract=off: +, -, *, / and sqrt round identically; only the libm calls are the host's here and the
device's on the GPU (their ulps are what the `-m gpu` parity tests allow for).  So: reference source == oracle == the source the kernels are compiled from."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from common import ROOT

CUDA_INC = next((p for p in (os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include"), "/usr/local/cuda/include") if os.path.exists(os.path.join(p, "cuda_runtime.h"))), None)
pytestmark = pytest.mark.skipif(CUDA_INC is None or shutil.which("g++") is None, reason="needs g++ and the CUDA headers")
f32p = C.POINTER(C.c_float)


@pytest.fixture(scope="module")
def dev(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hostdev") / "libdevice_on_host.so")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I" + CUDA_INC, "-D__device__=", "-D__host__=", "-D__global__=", "-D__shared__=",
           "-D__forceinline__=inline", "-D__noinline__=", "-Wno-unused-function", os.path.join(ROOT, "tests", "hostdev", "device_on_host.cpp"), "-o", so]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return C.CDLL(so)


def _truth():
    """The reference's own functions when they were compiled here, the oracle's restatement (bit-equal to them, tests/test_oracle_bsdf.py) otherwise."""
    return O.microfacet("ref" if os.path.exists(O.MFREF_SO) else "port")


def _dirs(rng, n, upper):
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    if upper:
        d[:, 2] = np.abs(d[:, 2])
    return np.ascontiguousarray(d, np.float32)


@pytest.mark.parametrize("type_", [0, 1])
@pytest.mark.parametrize("alpha", [0.01, 0.1, 0.6])
def test_device_microfacet_source_equals_the_reference(dev, type_, alpha):
    ref = _truth(); rng = np.random.default_rng(7 + 10 * type_ + int(100 * alpha)); n = 100000
    m, v, wi = _dirs(rng, n, False), _dirs(rng, n, False), _dirs(rng, n, True)
    wi[:500] = [0, 0, 1]; wi[500:1000, 2] = 1e-3 * rng.random(500); wi[500:1000] /= np.linalg.norm(wi[500:1000], axis=1, keepdims=True)
    smp = rng.random((n, 2), dtype=np.float32)
    out = np.zeros(n, np.float32)
    dev.dev_mf_eval.argtypes = [C.c_int, C.c_float, C.c_size_t, f32p, f32p]
    dev.dev_mf_eval(type_, alpha, n, m.ctypes.data_as(f32p), out.ctypes.data_as(f32p)); assert np.array_equal(out, ref.eval(type_, alpha, m))
    dev.dev_mf_smith_g1.argtypes = [C.c_int, C.c_float, C.c_size_t, f32p, f32p, f32p]
    dev.dev_mf_smith_g1(type_, alpha, n, v.ctypes.data_as(f32p), m.ctypes.data_as(f32p), out.ctypes.data_as(f32p)); assert np.array_equal(out, ref.smith_g1(type_, alpha, v, m))
    ma = np.ascontiguousarray(np.abs(m))
    dev.dev_mf_pdf.argtypes = [C.c_int, C.c_float, C.c_size_t, f32p, f32p, f32p]
    dev.dev_mf_pdf(type_, alpha, n, wi.ctypes.data_as(f32p), ma.ctypes.data_as(f32p), out.ctypes.data_as(f32p)); assert np.array_equal(out, ref.pdf(type_, alpha, wi, ma))
    mo = np.zeros((n, 3), np.float32); po = np.zeros(n, np.float32)
    dev.dev_mf_sample.argtypes = [C.c_int, C.c_float, C.c_size_t, f32p, f32p, f32p, f32p]
    dev.dev_mf_sample(type_, alpha, n, wi.ctypes.data_as(f32p), smp.ctypes.data_as(f32p), mo.ctypes.data_as(f32p), po.ctypes.data_as(f32p))
    mr, pr = ref.sample(type_, alpha, wi, smp)
    ok = np.isfinite(mr).all(axis=1)
    assert ok.mean() > 0.999 and np.array_equal(mo[ok], mr[ok]) and np.array_equal(po[ok], pr[ok])


def test_device_helper_sources_equal_the_reference(dev):
    ref = _truth(); rng = np.random.default_rng(3); n = 100000
    x = np.concatenate([rng.uniform(-0.999999, 0.999999, n), rng.uniform(-6, 6, 1000)]).astype(np.float32)
    a = np.zeros_like(x); b = np.zeros_like(x)
    dev.dev_erf.argtypes = [C.c_size_t, f32p, f32p, f32p]
    dev.dev_erf(len(x), x.ctypes.data_as(f32p), a.ctypes.data_as(f32p), b.ctypes.data_as(f32p))
    er, eir = ref.erf(x)
    assert np.array_equal(a, er) and np.array_equal(b[np.abs(x) < 1], eir[np.abs(x) < 1])
    c = np.concatenate([rng.uniform(-1, 1, n), [0.0, 1.0, -1.0, 1e-6]]).astype(np.float32)
    f = np.zeros_like(c); ct = np.zeros_like(c)
    dev.dev_fresnel_dielectric_ext.argtypes = [C.c_size_t, f32p, C.c_float, f32p, f32p]
    for eta in (1.0, 1.5046 / 1.000277, 1 / 1.5, 1.33):
        dev.dev_fresnel_dielectric_ext(len(c), c.ctypes.data_as(f32p), eta, f.ctypes.data_as(f32p), ct.ctypes.data_as(f32p))
        fr, tr = ref.fresnel_dielectric_ext(c, eta)
        assert np.array_equal(f, fr) and np.array_equal(ct, tr)
    ca = np.ascontiguousarray(np.abs(c)); out3 = np.zeros((len(c), 3), np.float32)
    dev.dev_fresnel_conductor_exact.argtypes = [C.c_size_t, f32p, f32p, f32p, f32p]
    for eta, k in (((0.2, 0.9, 1.1), (3.9, 2.4, 2.2)), ((1.65746, 0.880369, 0.521229), (9.22387, 6.26952, 4.837)), ((0, 0, 0), (1, 1, 1))):
        e = np.float32(eta); kk = np.float32(k)
        dev.dev_fresnel_conductor_exact(len(c), ca.ctypes.data_as(f32p), e.ctypes.data_as(f32p), kk.ctypes.data_as(f32p), out3.ctypes.data_as(f32p))
        assert np.array_equal(out3, ref.fresnel_conductor_exact(ca, eta, k))
    d = _dirs(rng, n, False); d[:3] = np.eye(3)
    bo = np.zeros_like(d); co = np.zeros_like(d)
    dev.dev_coordinate_system.argtypes = [C.c_size_t, f32p, f32p, f32p]
    dev.dev_coordinate_system(n, d.ctypes.data_as(f32p), bo.ctypes.data_as(f32p), co.ctypes.data_as(f32p))
    br, cr = ref.coordinate_system(d)
    assert np.array_equal(bo, br) and np.array_equal(co, cr)
    smp = rng.random((n, 2), dtype=np.float32); smp[:3] = [[0.5, 0.5], [0, 0], [0.99999994, 0.5]]
    vo = np.zeros((n, 3), np.float32)
    dev.dev_square_to_cosine_hemisphere.argtypes = [C.c_size_t, f32p, f32p]
    dev.dev_square_to_cosine_hemisphere(n, smp.ctypes.data_as(f32p), vo.ctypes.data_as(f32p))
    assert np.array_equal(vo, ref.square_to_cosine_hemisphere(smp))


def test_device_triangle_test_and_transmittance_sources_equal_the_reference(dev):
    """tri_intersect on the accel rows ppg_set_scene packs, fed with the reference's own TriAccel constants: same hit decision, same (t, u, v) as TriAccel::rayIntersect."""
    truth = C.CDLL(O.MFREF_SO) if os.path.exists(O.MFREF_SO) else O.load("port"); name = "mfref_triaccel" if os.path.exists(O.MFREF_SO) else "ppgo_triaccel"
    rng = np.random.default_rng(5); n = 200000
    A = rng.normal(size=(n, 3)).astype(np.float32) * 10; B = A + rng.normal(size=(n, 3)).astype(np.float32); Cc = A + rng.normal(size=(n, 3)).astype(np.float32)
    B[:100] = A[:100]
    w = rng.random((n, 2)); tgt = A + (B - A) * (w[:, :1] * 1.4 - 0.2) + (Cc - A) * (w[:, 1:] * 1.4 - 0.2)
    o = (tgt + rng.normal(size=(n, 3)) * 5).astype(np.float32); d = tgt - o; d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    mint = np.full(n, 1e-4, np.float32); maxt = np.where(rng.random(n) < 0.1, 3.0, np.inf).astype(np.float32)
    k = np.zeros(n, np.int32); c9 = np.zeros((n, 9), np.float32); hit = np.zeros(n, np.uint8); tuv = np.zeros((n, 3), np.float32)
    fn = getattr(truth, name); fn.argtypes = [C.c_size_t] + [f32p] * 7 + [C.POINTER(C.c_int), f32p, C.POINTER(C.c_ubyte), f32p]
    arrs = [np.ascontiguousarray(a, np.float32) for a in (A, B, Cc, o, d, mint, maxt)]
    fn(n, *[a.ctypes.data_as(f32p) for a in arrs], k.ctypes.data_as(C.POINTER(C.c_int)), c9.ctypes.data_as(f32p), hit.ctypes.data_as(C.POINTER(C.c_ubyte)), tuv.ctypes.data_as(f32p))
    hit2 = np.zeros(n, np.uint8); tuv2 = np.zeros((n, 3), np.float32)
    dev.dev_tri_intersect.argtypes = [C.c_size_t, C.POINTER(C.c_int), f32p, f32p, f32p, f32p, f32p, C.POINTER(C.c_ubyte), f32p]
    dev.dev_tri_intersect(n, k.ctypes.data_as(C.POINTER(C.c_int)), c9.ctypes.data_as(f32p), arrs[3].ctypes.data_as(f32p), arrs[4].ctypes.data_as(f32p), mint.ctypes.data_as(f32p),
                          maxt.ctypes.data_as(f32p), hit2.ctypes.data_as(C.POINTER(C.c_ubyte)), tuv2.ctypes.data_as(f32p))
    assert 0.1 < hit.mean() < 0.5 and np.array_equal(hit, hit2) and np.array_equal(tuv[hit == 1], tuv2[hit == 1])
    from ppg_b200 import rtrans
    lut, _ = rtrans.reduce_for_material("beckmann", 1.49, 0.1); lut = np.ascontiguousarray(lut, np.float32)
    cs = np.concatenate([rng.uniform(-0.2, 1, 100000), [0.0, 1.0]]).astype(np.float32)
    a = np.zeros_like(cs); b = np.zeros_like(cs)
    dev.dev_rough_transmittance.argtypes = [C.c_size_t, f32p, f32p, f32p]
    dev.dev_rough_transmittance(len(cs), cs.ctypes.data_as(f32p), lut.ctypes.data_as(f32p), a.ctypes.data_as(f32p))
    port = O.load("port"); port.ppgo_rough_transmittance.argtypes = [C.c_size_t, f32p, f32p, f32p]
    port.ppgo_rough_transmittance(len(cs), cs.ctypes.data_as(f32p), lut.ctypes.data_as(f32p), b.ctypes.data_as(f32p))
    assert np.array_equal(a, b)                                                     # (the oracle's lookup is itself checked against evalCubicInterp1D, tests/test_oracle_bsdf.py)
