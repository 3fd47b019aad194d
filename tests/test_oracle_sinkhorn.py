"""CPU: self-consistency pins of the Sinkhorn oracle (geomloss is absent: parity unpinned)."""
import numpy as np

from aadg_b200.synth import feature_cloud
from oracle import sinkhorn as S


def clouds(n=8, d=128):
    return feature_cloud(n, d, 0), feature_cloud(n, d, 1), feature_cloud(n, d, 2)


def test_zero_on_identical_clouds_and_symmetry():
    x, y, z = clouds()
    assert abs(S.sinkhorn_divergence(x, x.copy())) < 1e-12
    assert abs(S.sinkhorn_divergence(x, y) - S.sinkhorn_divergence(y, x)) < 1e-12
    assert S.sinkhorn_divergence(x, y) > 0


def test_epsilon_schedule_shape():
    eps = S.epsilon_schedule(27.0)
    assert eps[0] == 27.0 ** 2 and abs(eps[1] - eps[0]) < 1e-9 and eps[-1] == 0.05 ** 2
    assert all(a >= b * (1 - 1e-12) for a, b in zip(eps, eps[1:]))
    ratios = [eps[i + 1] / eps[i] for i in range(1, len(eps) - 2)]
    assert np.allclose(ratios, 0.25)
    assert len(eps) == 2 + int(np.ceil(np.log(27.0 / 0.05) / np.log(2)))


def test_agrees_with_independent_dense_sinkhorn():
    """the eps-scaling loop lands on the fixed point of plain Sinkhorn at eps = blur^2."""
    x, y, _ = clouds(8, 32)
    a = S.sinkhorn_divergence(x, y)
    b = S.dense_sinkhorn_reference(x, y, 0.05 ** 2, iters=3000)
    assert abs(a - b) < 2e-2 * abs(b), (a, b)   # one pass per scale: geomloss stops ~1% short of the fixed point


def test_fp32_restatement_close_to_fp64():
    x, y, z = clouds()
    for p, q in ((x, y), (y, z), (x, z)):
        a64 = S.sinkhorn_divergence(p, q, np.float64)
        a32 = S.sinkhorn_divergence(p, q, np.float32)
        assert abs(a32 - a64) < 1e-4 * abs(a64), (a32, a64)


def test_unequal_sizes_and_rewards():
    x, y, z = clouds(8, 16)
    assert S.sinkhorn_divergence(x[:5], y) > 0
    feat = np.concatenate([x, y, z])          # 24 rows: 8 per domain
    m = 2
    # interleave so that rows j::M hold 4 points of each domain
    order = np.arange(24).reshape(3, 4, 2).transpose(1, 0, 2).reshape(-1)
    feat = feat[order]
    dom = np.repeat(np.arange(3), 8)[order]
    dc = np.eye(3, dtype=np.float32)[dom] * 0.9 + 0.03
    inc, vals = S.diversity_rewards(feat, dc, m)
    assert inc.shape == (2,) and vals.shape == (2, 3) and (vals > 0).all()
    assert np.allclose(inc, vals.sum(1))
    r = S.normalize_rewards(np.array([1.0, 2.0, 4.0]))
    assert abs(r.mean()) < 1e-12


def test_pinned_to_exact_optimal_transport():
    """An independent pin of the QUANTITY (geomloss itself is absent): at blur = 0.05 the debiased Sinkhorn divergence is
    the entropic (eps = blur^2 = 0.0025) approximation of the optimal-transport cost under the cosine ground cost, which
    for uniform clouds of equal size is an assignment problem -- solved exactly here by scipy's Hungarian solver, code
    that shares nothing with the oracle.  Exact for one and two points per cloud (the plan is a permutation and the
    entropic terms cancel in the debiasing); within 2 % beyond (measured 0.3-1.2 %: the entropic blur)."""
    from scipy.optimize import linear_sum_assignment
    for n, d, tol in ((1, 16, 1e-12), (2, 16, 1e-9), (8, 128, 2e-2), (8, 32, 2e-2), (16, 64, 2e-2), (64, 128, 2e-2)):
        x, y = feature_cloud(n, d, 0), feature_cloud(n, d, 2)
        C = S.cosine_cost(x, y)
        r, c = linear_sum_assignment(C)
        exact = C[r, c].sum() / n
        got = S.sinkhorn_divergence(x, y)
        assert abs(got - exact) <= tol * exact, (n, d, got, exact)
        assert got <= exact * (1 + 1e-9)        # debiased entropic OT never exceeds the unregularised cost here


def test_invariant_under_point_order_and_feature_scale():
    """properties of the reference's call the restatement must keep: clouds are SETS (row order is irrelevant) and the
    cosine cost ignores the length of a feature vector -- only the bounding-box diameter that starts the epsilon schedule
    sees the scale (geomloss max_diameter), so the pinned-diameter value is scale free."""
    x, y, _ = clouds(8, 32)
    rng = np.random.RandomState(0)
    a = S.sinkhorn_divergence(x, y)
    assert abs(S.sinkhorn_divergence(x[rng.permutation(8)], y[rng.permutation(8)]) - a) < 1e-12
    b = S.sinkhorn_divergence(x, y, diameter=3.0)
    c = S.sinkhorn_divergence(4.0 * x, 0.5 * y, diameter=3.0)       # powers of two: exact in float32
    assert abs(b - c) < 1e-12 * abs(b)
