"""CPU: the HD95 oracle (scipy restatement of medpy.metric.binary.hd95) against an independent brute force."""
import numpy as np
import pytest

from oracle import hd95 as O


blobs = O.random_blobs


@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_brute_force(seed):
    rng = np.random.RandomState(seed)
    a, b = blobs(rng, 40, 52, 3), blobs(rng, 40, 52, 2)
    if not a.any() or not b.any():
        pytest.skip("empty mask")
    for pct in (95, 100, 50):
        assert O.hd95(a, b, pct) == pytest.approx(O.brute_force(a, b, pct), abs=1e-12)


def test_known_answers():
    a = np.zeros((20, 20), bool)
    b = np.zeros((20, 20), bool)
    a[5, 5] = True
    b[5, 9] = True
    assert O.hd95(a, b) == 4.0                 # single pixels: both directed distances are 4
    a[:] = False
    a[2:10, 2:10] = True
    assert O.hd95(a, a) == 0.0
    with pytest.raises(RuntimeError):
        O.hd95(np.zeros((4, 4), bool), a[:4, :4])
