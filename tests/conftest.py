import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "practical-path-guiding_b200"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running statistical test")
    config.addinivalue_line("markers", "seeds3: pixel-level parity of a trained render, stated as a majority over three fixed seeds (see pytest_pyfunc_call)")


SEEDS3 = (1234, 20260924, 77)


@pytest.hookimpl(tryfirst=True)
def pytest_pyfunc_call(pyfuncitem):
    """Tests marked `seeds3` compare TRAINED renders of two implementations pixel by pixel.  Both sides accumulate the SD-tree statistics with
    floating-point atomics (the reference's addToAtomicFloat, GP:59-62; red.global.add.f32 on the device), so sums differ in the last ulp from
    run to run; about one training run in thirty has a path whose random number falls between two such roundings of a quadtree partition, takes
    the other child, and the later iterations decorrelate to noise level.  There is no seed for which that cannot happen, so these tests state
    their claim as a majority over three FIXED seeds: the assertion must hold for two of them (evaluated lazily: two passes or two failures end
    the test).  A defect that is not a one-in-thirty rounding flip fails two seeds.  Every other test runs exactly once."""
    if pyfuncitem.get_closest_marker("seeds3") is None:
        return None
    import common
    fn = pyfuncitem.obj
    args = {a: pyfuncitem.funcargs[a] for a in pyfuncitem._fixtureinfo.argnames}
    passed, failures = 0, []
    try:
        for seed in SEEDS3:
            common.SEED_OVERRIDE = seed
            try:
                fn(**args); passed += 1
            except AssertionError as e:
                failures.append((seed, e))
            if passed == 2 or len(failures) == 2:
                break
    finally:
        common.SEED_OVERRIDE = None
    if passed < 2:
        raise AssertionError(f"failed for seeds {[s for s, _ in failures]} of {SEEDS3}: {failures[-1][1]}") from failures[-1][1]
    return True
