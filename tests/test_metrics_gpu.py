"""GPU: HD95 kernels (C ABI) against the scipy/numpy restatement of medpy.metric.binary.hd95 -- bit-exact float64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import hd95 as O  # noqa: E402
blobs = O.random_blobs


@pytest.fixture(scope="module")
def M():
    assert torch.cuda.is_available()
    from aadg_b200.ops import metrics
    return metrics


@pytest.mark.parametrize("h,w", [(64, 64), (97, 131), (512, 512)])
def test_hd95_matches_oracle_bit_exact(M, h, w):
    rng = np.random.RandomState(h + w)
    res, ref = [], []
    for i in range(6):
        a, b = blobs(rng, h, w, 1 + i % 3), blobs(rng, h, w, 1 + (i + 1) % 3)
        if i == 3:                                  # speckle noise: many tiny surfaces
            a = a ^ (rng.rand(h, w) < 0.05)
        if i == 4:                                  # mask touching every image border
            b[:] = True
            b[h // 3:h // 2, w // 3:w // 2] = False
        if not a.any():
            a[h // 2, w // 2] = True
        if not b.any():
            b[h // 3, w // 3] = True
        res.append(a)
        ref.append(b)
    r = torch.from_numpy(np.stack(res)).cuda()
    t = torch.from_numpy(np.stack(ref)).cuda()
    got = M.hd95(r, t).cpu().numpy()
    for pct in (95.0, 100.0, 37.5):
        v, st = M.surface_distance_percentile(r, t, pct)
        assert (st == 0).all()
        want = np.array([O.hd95(a, b, pct) for a, b in zip(res, ref)])
        assert np.array_equal(v.cpu().numpy(), want), (pct, v.cpu().numpy(), want)
    assert np.array_equal(got, np.array([O.hd95(a, b) for a, b in zip(res, ref)]))


def test_hd95_empty_masks_and_validation_rule(M):
    h = w = 48
    rng = np.random.RandomState(3)
    pred = np.stack([np.stack([blobs(rng, h, w, 2), np.zeros((h, w), bool)]) for _ in range(3)])   # class 1 predicted empty
    gt = np.stack([np.stack([blobs(rng, h, w, 2) | True, blobs(rng, h, w, 2) | (np.arange(w) < 5)]) for _ in range(3)])
    p, g = torch.from_numpy(pred).cuda(), torch.from_numpy(gt).cuda()
    v, st = M.surface_distance_percentile(p, g)
    assert st[:, 1].eq(1).all() and st[:, 0].eq(0).all() and torch.isnan(v[:, 1]).all()
    with pytest.raises(RuntimeError):
        M.hd95(p, g)
    with pytest.raises(RuntimeError):
        M.hd95(g[:, :1], torch.zeros_like(g[:, :1]))
    got = M.validation_hd95(p.float(), g.float()).cpu().numpy()            # search_dg.py:246-260
    want0 = np.mean([O.hd95(pred[i, 0], gt[i, 0]) for i in range(3)])
    assert got[0] == pytest.approx(want0, abs=1e-12) and got[1] == 100.0
    with pytest.raises(RuntimeError):
        M.hd95(p.cpu(), g.cpu())
