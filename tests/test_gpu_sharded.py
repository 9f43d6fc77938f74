"""N > 1 path of the CUDA library on ONE GPU (two processes share cuda:0, the exchange goes through gloo): tile sharding
+ the per-iteration tree allreduce reproduce the unsharded render."""
import numpy as np
import pytest

from common import load_cbox
from test_sharded_cpu import _spawn

pytestmark = pytest.mark.gpu


def test_two_ranks_match_single_rank_on_gpu():
    from ppg_b200.integrator import GuidedPathTracer
    r0, r1 = _spawn("gpu", 128, "60")
    assert np.array_equal(r0["img"], r1["img"])                      # the film is all-reduced: identical on both ranks
    assert list(r0["leaves"]) == list(r1["leaves"]) and np.array_equal(r0["weights"], r1["weights"]) and np.array_equal(r0["variance"], r1["variance"])
    assert int(r0["paths"]) + int(r1["paths"]) == 128 * 128 * 60
    sc = load_cbox(128)
    g = GuidedPathTracer(dict(sc.integrator, budget="60")); g.set_scene(sc)
    img, st = g.render()
    w = [i["weight_avg"] * i["s_tree_leaves"] for i in st["iterations"]]
    assert r0["weights"][0] == w[0]                                   # unguided iteration: same paths, integer count
    assert list(r0["leaves"])[:2] == [i["s_tree_leaves"] for i in st["iterations"]][:2]
    assert np.allclose(r0["weights"], w, rtol=0.01)
    assert np.allclose(r0["variance"][0], st["iterations"][0]["variance"], rtol=1e-3)
    assert abs(r0["img"].mean() - img.mean()) <= 0.02 * img.mean()
