"""GPU tests of the integrator's control surface: the time-budget schedule (renderTime, GP:1434-1514) driven by an injected clock, asynchronous
cancel (Integrator::cancel, GP:1643-1648) and the progressive film."""
import threading
import time

import numpy as np
import pytest

from common import load_cbox

pytestmark = pytest.mark.gpu


def _gpu(props, scene):
    from ppg_b200.integrator import GuidedPathTracer
    g = GuidedPathTracer(props); g.set_scene(scene)
    return g


@pytest.mark.parametrize("combination", ["discard", "automatic", "inversevar"])
def test_seconds_budget_schedule_with_an_injected_clock(combination):
    """budgetType=seconds with an injected clock that advances 0.05 s per rendered pass (read off the progressive-film callback): the schedule is then a
    pure function of the algorithm.  renderTime renders 2^k passes in iteration k and stops when the clock passes the budget; with `automatic` the
    CURRENT iteration becomes the final one as soon as less time remains than it took (or the extrapolated variance grew) and keeps rendering batches
    of 2^k passes into the same film until the time is up (GP:1482-1501).  Budget 4 s: iterations of 1, 2, 4, 8, 16 passes take 1.55 s; the 32 passes
    of iteration 5 end at 3.15 s with 0.85 s < 1.6 s left -> final, one more batch of 32 -> 64 passes.  Without `automatic` no iteration is ever
    declared final and iteration 6 (64 passes) runs past the budget."""
    sc = load_cbox(48)
    g = _gpu(dict(sc.integrator, budgetType="seconds", budget="4", sampleCombination=combination), sc)
    state = {"passes": 0, "reads": 0}

    def film(ptr, w, h, passes):
        state["passes"] = passes

    def clock():
        state["reads"] += 1
        return 0.05 * state["passes"] + 1e-5 * state["reads"]
    g.set_clock(clock); g.set_film_callback(film)
    img, st = g.render()
    it = st["iterations"]
    assert np.isfinite(img).all() and img.mean() > 0.01
    if combination == "automatic":
        assert [i["passes"] for i in it] == [1, 2, 4, 8, 16, 64] and [i["is_final"] for i in it] == [0, 0, 0, 0, 0, 1]
    else:
        assert [i["passes"] for i in it] == [1, 2, 4, 8, 16, 32, 64] and not any(i["is_final"] for i in it)
    assert st["total_passes"] == sum(i["passes"] for i in it)
    # the same clock again gives the same schedule (no wall-clock dependence is left)
    state.update(passes=0, reads=0)
    _, st2 = g.render()
    assert [i["passes"] for i in st2["iterations"]] == [i["passes"] for i in it]


def test_cancel_from_a_second_thread():
    """Integrator::cancel() is asynchronous and thread-safe: a render with an hour of budget returns PPG_ERR_CANCELLED within a fraction of a second
    of the call (the flag is polled before every bounce launch), the film holds what was finished, and the integrator renders again afterwards."""
    sc = load_cbox(512)
    g = _gpu(dict(sc.integrator, budgetType="seconds", budget="3600"), sc)
    out = {}

    def run():
        out["img"], out["st"] = g.render()
        out["t"] = time.perf_counter()
    th = threading.Thread(target=run); th.start()
    time.sleep(1.0)
    t_cancel = time.perf_counter()
    g.cancel()
    th.join(timeout=30)
    assert not th.is_alive() and g.last_status == -5                     # PPG_ERR_CANCELLED
    assert out["t"] - t_cancel < 1.0, out["t"] - t_cancel
    assert np.isfinite(out["img"]).all() and out["img"].mean() > 0.01 and out["st"]["total_passes"] >= 1
    g2 = _gpu(dict(sc.integrator, budget="8"), sc.with_film(64, 64))
    img, st = g2.render()
    assert g2.last_status == 0 and st["total_passes"] == 2


def test_progressive_film_callback():
    """ppg_set_film_callback: the film is handed to the host after every performRenderPasses (the reference puts finished blocks into the film while
    rendering, renderproc.cpp:143-151): one call per iteration in spp mode, passes increasing, and the last film is the image render() returns."""
    import torch
    sc = load_cbox(64)
    g = _gpu(dict(sc.integrator, budget="28"), sc)
    seen = []

    class _Ptr:
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}

    def film(ptr, w, h, passes):
        seen.append((passes, torch.as_tensor(_Ptr(ptr, w * h * 3), device="cuda").cpu().numpy().reshape(h, w, 3).copy()))
    g.set_film_callback(film)
    img, st = g.render()
    assert [p for p, _ in seen] == [1, 3, 7] and st["total_passes"] == 7
    assert all(np.isfinite(f).all() for _, f in seen) and np.allclose(seen[-1][1], img, rtol=1e-6)
