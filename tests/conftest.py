import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def statistical(fn):
    """For the few tests whose assertions compare SINGLE DRAWS of the engine's run-to-run noise (whole-network loss /
    Dice / gradient-cosine against an fp32 oracle, graph-vs-eager trajectories): the engine is not bitwise reproducible
    (fp32 atomics re-quantised by bf16 storage, DESIGN.md "Precision"), so a bound at k sigma fails one run in N by
    construction.  A failure of such a test has to REPRODUCE to count: the test body is run once more and the second
    verdict stands (the first failure is printed).  Bit-exact and per-layer teacher-forced tests are never wrapped."""
    import functools
    import gc

    @functools.wraps(fn)
    def run(*args, **kwargs):
        try:
            return fn(*args, **kwargs)
        except AssertionError as e:
            print("STATISTICAL-RETRY %s: first attempt failed with %s" % (fn.__name__, str(e).splitlines()[0][:300] if str(e) else "assert"))
        gc.collect()
        try:
            import torch
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        except ImportError:
            pass
        return fn(*args, **kwargs)
    return run


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
