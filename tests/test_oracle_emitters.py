"""Direct sampling of area emitters in the oracle (Scene::sampleAttenuatedEmitterDirect -> AreaLight / TriMesh / Sphere::sampleDirect), checked like the emitter part of
the reference's src/tests/test_chisquare.cpp:391-623 checks it: the sampling weight integrates to radiance x solid angle, the reported density is the analytic one."""
import numpy as np

import oracle_lib as O
from common import load_cbox
from ppg_b200 import builtin_scenes as B


def _solid_angle_of_triangles(P, I, ref, n=400):
    """Unsigned solid angle of a planar mesh seen from `ref` by midpoint quadrature over each triangle's barycentric grid."""
    total = 0.0
    u, v = np.meshgrid((np.arange(n) + 0.5) / n, (np.arange(n) + 0.5) / n, indexing="ij")
    keep = (u + v) < 1.0
    w = np.where(np.abs(u + v - 1.0 + 0.5 / n) < 0.6 / n, 0.5, 1.0)[keep]            # half cells on the diagonal
    u, v = u[keep], v[keep]
    for a, b, c in I:
        p0, e1, e2 = P[a].astype(np.float64), (P[b] - P[a]).astype(np.float64), (P[c] - P[a]).astype(np.float64)
        nrm = np.cross(e1, e2); area = 0.5 * np.linalg.norm(nrm); nrm /= 2 * area
        x = p0 + u[:, None] * e1 + v[:, None] * e2 - ref
        r2 = (x * x).sum(1)
        total += (np.abs(x @ nrm) / r2 ** 1.5 * w).sum() / w.sum() * area
    return total


def test_quad_light_weight_integrates_to_radiance_times_solid_angle():
    sc = load_cbox(32)
    o = O.Oracle(O.params_from_xml(dict(sc.integrator, nee="always")), sc, kind="port")
    e = int(np.nonzero(sc.shapes[:, 3] >= 0)[0][0]); first, cnt = int(sc.shapes[e, 0]), int(sc.shapes[e, 1])
    rng = np.random.default_rng(2)
    n = 400000
    for ref in ([278.0, 520.0, 270.0], [150.0, 540.0, 400.0]):       # between the lamp (it shines UP at the ceiling in this CBOX, y = 471.2) and the ceiling
        refs = np.tile(np.float32(ref), (n, 1))
        d, val, pdf, dist = o.emitter_sample_direct(refs, np.zeros_like(refs), rng.random((n, 2), dtype=np.float32))
        assert (pdf > 0).all() and np.isfinite(val).all()                              # nothing in between, two-sided reference point
        omega = _solid_angle_of_triangles(sc.positions, sc.indices[first:first + cnt], np.float64(ref))
        est = val.astype(np.float64).mean(0)
        np.testing.assert_allclose(est, sc.area_radiance[0] * omega, rtol=4e-3)
        # density = dist^2 / (area * |cos|): value x pdf is the radiance, sample by sample
        np.testing.assert_allclose(val * pdf[:, None], np.tile(sc.area_radiance[0], (n, 1)), rtol=1e-4)
    # a reference point BELOW the lamp (it is one-sided) gets nothing
    refs = np.tile(np.float32([278.0, 300.0, 100.0]), (1000, 1))
    assert (o.emitter_sample_direct(refs, np.zeros_like(refs), rng.random((1000, 2), dtype=np.float32))[2] == 0).all()
    # a front-side normal that faces away from the lamp rejects every sample (dot(d, refN) >= 0, area.cpp:164-168)
    refs = np.tile(np.float32([278.0, 520.0, 270.0]), (1000, 1)); up = np.tile(np.float32([0, 1, 0]), (1000, 1))
    assert (o.emitter_sample_direct(refs, up, rng.random((1000, 2), dtype=np.float32))[2] == 0).all()


def test_sphere_lights_cone_and_inside_sampling():
    """cbox_with_analytic_spheres: three lights chosen uniformly (lamp, small emitting ball seen from outside: uniform cone; huge inward-facing shell seen from
    inside: uniform sphere area).  Samples of the ball carry the analytic cone density and integrate to radiance x cone solid angle."""
    sc = B.cbox_with_analytic_spheres(load_cbox(32))
    o = O.Oracle(O.params_from_xml(dict(sc.integrator, nee="always")), sc, kind="port")
    rng = np.random.default_rng(3)
    n = 300000
    ref = np.float32([278.0, 300.0, 200.0]); refs = np.tile(ref, (n, 1))
    smp = rng.random((n, 2), dtype=np.float32)
    d, val, pdf, dist = o.emitter_sample_direct(refs, np.zeros_like(refs), smp)
    ball = (smp[:, 0] >= 1 / 3) & (smp[:, 0] < 2 / 3)                                   # DiscreteDistribution over three equal weights, emitter order of the scene
    c, r = np.float64([150.0, 400.0, 250.0]), 25.0
    dist_c = np.linalg.norm(c - ref); cos_a = np.sqrt(1 - (r / dist_c) ** 2); omega = 2 * np.pi * (1 - cos_a)
    lit = ball & (pdf > 0)
    assert lit.sum() > 0.99 * ball.sum()                                                # nothing between the point and the ball
    np.testing.assert_allclose(pdf[lit], 1.0 / omega / 3.0, rtol=1e-4)
    to_c = (c - ref) / dist_c
    assert ((d[lit].astype(np.float64) @ to_c) >= cos_a - 1e-5).all()                   # inside the cone
    est = (val[ball].astype(np.float64).sum(0) / n)                                     # (1/3 of the samples) x (weight 3 L / pdf_SA)
    np.testing.assert_allclose(est, sc.area_radiance[-2] * omega, rtol=5e-3)
    # the shell (radius 1500 around the box, normals flipped) is sampled over its area from inside; only the open front of the box lets it through
    shell = smp[:, 0] >= 2 / 3
    seen = shell & (val.sum(1) > 0)                                                     # an occluded sample keeps its density and carries zero
    assert 0.0 < seen.sum() < 0.5 * shell.sum()
    np.testing.assert_allclose(val[seen] * pdf[seen, None], np.tile(sc.area_radiance[-1], (int(seen.sum()), 1)), rtol=1e-4)
    assert (d[seen, 2] < 0).all()                                                       # the box opens towards -z (the camera side)
