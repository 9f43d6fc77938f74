"""GPU box: diagnostics of the environment light sampling (nee != never with an envmap) against the CPU oracle -- prints the deviations the tests in
tests/test_gpu_z_envnee.py assert on, so that one run shows how far inside / outside the tolerances the CUDA path is."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "practical-path-guiding_b200"), os.path.join(ROOT, "tests")]
import oracle_lib as O  # noqa: E402
from common import load_fixture_scene, relmse  # noqa: E402
from ppg_b200 import builtin_scenes as B  # noqa: E402
from ppg_b200.integrator import GuidedPathTracer  # noqa: E402


def ops(name, sc):
    props = dict(sc.integrator, nee="always")
    g = GuidedPathTracer(props); g.set_scene(sc); o = O.Oracle(O.params_from_xml(props), sc, kind="port")
    rng = np.random.default_rng(11); n = 100000
    lo, hi = np.asarray(sc.aabb_min, np.float64), np.asarray(sc.aabb_max, np.float64)
    ref = (lo + (hi - lo) * (0.1 + 0.8 * rng.random((n, 3)))).astype(np.float32)
    nrm = rng.normal(size=(n, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True); nrm[n // 2:] = 0; nrm = nrm.astype(np.float32)
    smp = rng.random((n, 2), dtype=np.float32)
    d, val, pdf, dist = g.op_emitter_sample_direct(ref, nrm, smp, 3)
    d0, val0, pdf0, dist0 = o.emitter_sample_direct(ref, nrm, smp, 3)
    ok = pdf0 > 0
    print(f"[{name}] sample_direct: lit {ok.mean():.3f}  pdf equal(1e-4) {np.isclose(pdf, pdf0, rtol=1e-4, atol=1e-7).mean():.5f}  "
          f"value equal {np.isclose(val, val0, rtol=1e-4, atol=1e-6).all(axis=1).mean():.5f}  dist equal {np.isclose(dist, dist0, rtol=1e-4, atol=1e-5).mean():.5f}  "
          f"max |d - d0| (lit) {np.abs(d - d0)[ok].max():.2e}  pdf bit-equal {(pdf == pdf0).mean():.4f}  sum value {val.sum():.6g} vs {val0.sum():.6g}", flush=True)
    if sc.envmap is not None:
        dd = rng.normal(size=(200000, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True); dd = dd.astype(np.float32)
        p, v = g.op_env_pdf(dd); p0, v0 = o.env_pdf(dd)
        print(f"[{name}] env_pdf: equal(2e-3) {np.isclose(p, p0, rtol=2e-3, atol=1e-5).mean():.5f}  value equal {np.isclose(v, v0, rtol=2e-3, atol=1e-4).all(axis=1).mean():.5f}  "
              f"sum {p.sum():.6g} vs {p0.sum():.6g}", flush=True)
    g.close()


def render(name, sc, **extra):
    props = dict(dict(sc.integrator, budget="60"), **extra)
    t = time.time()
    g = GuidedPathTracer(props); g.set_scene(sc); img, st = g.render(); g.close()
    tg = time.time() - t
    ref, ost = O.Oracle(O.params_from_xml(props), sc, kind="port").render()
    close = np.isclose(img, ref, rtol=1e-3, atol=1e-5).all(axis=2)
    print(f"[{name} {extra}] vertices {st['total_vertices']} vs {ost['total_vertices']} ({st['total_vertices'] / ost['total_vertices'] - 1:+.2e})  pixels equal {close.mean():.4f}  "
          f"mean {img.mean():.5f} vs {ref.mean():.5f}  relMSE {relmse(img, ref):.2e}  finite {np.isfinite(img).all()}  gpu {tg:.2f}s", flush=True)
    for a, b in zip(st["iterations"], ost["iterations"]):
        print(f"    it {a['iteration']}: leaves {a['s_tree_leaves']} / {b['s_tree_leaves']}  weight {a['weight_avg'] * a['s_tree_leaves']:.1f} / {b['weight_avg'] * b['s_tree_leaves']:.1f}  "
              f"var {a['variance']:.4g} / {b['variance']:.4g}", flush=True)


if __name__ == "__main__":
    env = B.env_lit_scene(96); both = load_fixture_scene("cbox-textured-flat", 96)
    ops("env-only", B.env_lit_scene(32)); ops("area+env", load_fixture_scene("cbox-textured-flat", 32))
    render("env-only", env, nee="always", budget="4")
    render("env-only", env, nee="always")
    render("env-only", env, nee="kickstart", spatialFilter="stochastic", directionalFilter="box")
    render("area+env", both, nee="always")
    render("env-torus (BVH, trace pass)", B.env_lit_scene(96, torus=True), nee="always")
    render("env-only", env, nee="never")
