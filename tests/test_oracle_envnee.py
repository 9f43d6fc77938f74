"""Light sampling of the environment emitter in the oracle (nee = always | kickstart with an envmap / baked sunsky): EnvironmentMap::sampleDirect /
pdfDirect (src/emitters/envmap.cpp:516-633) behind Scene::sampleAttenuatedEmitterDirect (scene.cpp:876-897), checked the way the reference checks its
samplers (src/tests/test_chisquare.cpp): the sampled directions are distributed like the density the pdf function reports, sampling weight x density
= radiance, the density integrates to one -- and, end to end, a render with light sampling has the same expectation as one without."""
import numpy as np
import pytest

import oracle_lib as O
from ppg_b200 import builtin_scenes as B


def _oracle(nee="always", **kw):
    sc = B.env_lit_scene(48)
    props = dict(sc.integrator, nee=nee, **{k: str(v) for k, v in kw.items()})
    return O.Oracle(O.params_from_xml(props), sc, kind="port", nthreads=4), sc


def _sphere_grid(n_theta=512, n_phi=1024):
    th = (np.arange(n_theta) + 0.5) * np.pi / n_theta; ph = (np.arange(n_phi) + 0.5) * 2 * np.pi / n_phi
    T, P = np.meshgrid(th, ph, indexing="ij")
    d = np.stack([np.sin(T) * np.cos(P), np.cos(T), np.sin(T) * np.sin(P)], -1).reshape(-1, 3).astype(np.float32)
    dw = (np.sin(T) * (np.pi / n_theta) * (2 * np.pi / n_phi)).reshape(-1)
    return d, dw


def test_environment_density_integrates_to_one_and_follows_the_radiance():
    o, _ = _oracle()
    d, dw = _sphere_grid()
    pdf, val = o.env_pdf(d)
    assert np.isfinite(pdf).all() and (pdf >= 0).all()
    assert abs(float((pdf.astype(np.float64) * dw).sum()) - 1.0) < 5e-3          # the only emitter: choice probability 1
    lum = val @ np.array([0.212671, 0.715160, 0.072169])
    # the density is the bilinearly interpolated luminance x sin(theta_texel) / sin(theta): proportional to luminance up to the
    # row-weight interpolation (16 rows: sin(theta_row) / sin(theta) varies by tens of percent near the poles), and zero exactly where the map is black
    assert np.corrcoef(pdf, lum)[0, 1] > 0.95
    assert (pdf[lum == 0] == 0).all()


def test_environment_samples_are_distributed_like_the_density():
    o, sc = _oracle()
    rng = np.random.default_rng(5)
    n = 400000
    ref = np.tile(np.array([[0.0, 2.0, 0.0]], np.float32), (n, 1))               # above everything: only downward directions can be blocked
    d, val, pdf, dist = o.emitter_sample_direct(ref, np.zeros_like(ref), rng.random((n, 2), dtype=np.float32))
    # every sample whose density is positive points to a lit texel; value x pdf is the radiance there unless the ground plate blocks it
    up = (pdf > 0) & (d[:, 1] > 0)
    assert up.sum() > 0.9 * n
    pdf2, rad = o.env_pdf(d[up])
    # pdfDirect(sampleDirect().d) == sampleDirect().pdf after the uv round trip through atan2 / acos -- away from the poles: the reference's tent offset
    # can carry a sample of the first / last row across the pole (theta < 0), where its own pdfDirect looks up the mirrored texels (envmap.cpp:575, 593-599)
    inner = np.abs(d[up, 1]) < np.cos(1.5 * np.pi / 16)
    assert inner.mean() > 0.95
    np.testing.assert_allclose(pdf2[inner], pdf[up][inner], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose((val[up] * pdf[up, None])[inner], rad[inner], rtol=5e-3, atol=1e-4)
    assert np.abs(np.linalg.norm(d[up], axis=1) - 1).max() < 1e-5
    assert (dist[up] > 0).all()
    # histogram over (theta, phi) cells against the integrated density (chi-square like test_chisquare.cpp, pooled cells)
    nt, nph = 16, 32
    th = np.arccos(np.clip(d[up, 1], -1, 1)); ph = np.mod(np.arctan2(d[up, 2], d[up, 0]), 2 * np.pi)
    hist = np.histogram2d(th, ph, bins=[nt, nph], range=[[0, np.pi], [0, 2 * np.pi]])[0]
    g, dw = _sphere_grid(nt * 16, nph * 16)
    gp, _ = o.env_pdf(g)
    expected = (gp.astype(np.float64) * dw).reshape(nt, 16, nph, 16).sum(axis=(1, 3)) * n
    sel = expected > 20
    sel[0] = False; sel[nt // 2:] = False      # upper hemisphere (the plate blocks the rest), without the polar row (see above)
    chi2 = (((hist - expected) ** 2)[sel] / expected[sel]).sum()
    dof = sel.sum() - 1
    assert chi2 < dof + 6 * np.sqrt(2 * dof), (chi2, dof)
    assert hist[0].sum() < 0.05 * n


def test_samples_towards_occluders_carry_nothing():
    o, _ = _oracle()
    rng = np.random.default_rng(6)
    n = 20000
    ref = np.tile(np.array([[-0.15, 0.3, 0.0]], np.float32), (n, 1))             # inside the closed diffuse box
    d, val, pdf, dist = o.emitter_sample_direct(ref, np.zeros_like(ref), rng.random((n, 2), dtype=np.float32))
    assert (val == 0).all()


@pytest.mark.parametrize("nee", ["always", "kickstart"])
def test_render_with_environment_light_sampling_has_the_same_expectation(nee):
    """Unguided estimator identity: path tracing with MIS-combined environment light sampling converges to the image plain BSDF sampling converges to.
    Compared on block means of a 48 x 48 render (iteration structure 4+8+16 spp; with `always` the trained iterations use the guided mixture too)."""
    imgs = {}
    for mode in ("never", nee):
        acc = None
        for seed in (1, 2, 3):
            o, sc = _oracle(mode, seed=seed, budget=60, sampleCombination="discard")
            img, st = o.render()
            assert np.isfinite(img).all()
            acc = img if acc is None else acc + img
        imgs[mode] = acc / 3
    a = imgs["never"].reshape(6, 8, 6, 8, 3).mean(axis=(1, 3)); b = imgs[nee].reshape(6, 8, 6, 8, 3).mean(axis=(1, 3))
    lit = a.mean(-1) > 0.02
    rel = np.abs(a - b)[lit].mean() / a[lit].mean()
    assert rel < 0.06, rel
    assert abs(a.mean() - b.mean()) < 0.03 * a.mean(), (a.mean(), b.mean())
    # and light sampling does what it is for: the sun texel is found by every path, so the variance drops
    assert imgs[nee].std() < imgs["never"].std() * 1.05
