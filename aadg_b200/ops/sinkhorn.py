"""Sinkhorn diversity reward: torch-facing wrappers over aadg_sinkhorn_* (csrc/sinkhorn.cu).

`SamplesLoss` keeps the constructor/call shape of geomloss.SamplesLoss as the reference uses it
(search_dg.py:116,158-160): `SamplesLoss("sinkhorn", cost=<cosine KeOps formula>, backend="online")`
then `loss(x[N,d], y[M,d]) -> 0-d tensor`, forward value only.  There is no CPU path."""
import numpy as np
import torch

from .. import _lib

COSINE_COST = "( IntCst(1) - (X | Y) / ( Norm2(X) * Norm2(Y) ) )"


def _check(x):
    if not (isinstance(x, torch.Tensor) and x.is_cuda):
        raise RuntimeError("aadg_b200.ops.sinkhorn: inputs must be CUDA tensors (no CPU path)")
    if x.dtype != torch.float32 or x.dim() != 2:
        raise ValueError("clouds must be float32 [N, d]")
    return x.detach().contiguous()


def small_max_points():
    return _lib.lib().aadg_sinkhorn_small_max_points()


def divergence_batched(points, problems):
    """points float32 [R,d] (CUDA); problems int [P,4] rows (x_off, x_n, y_off, y_n) -> float32 [P],
    one launch, no host sync."""
    points = _check(points)
    prob = torch.as_tensor(np.ascontiguousarray(problems, dtype=np.int32)).to(points.device, non_blocking=True)
    out = torch.empty(prob.shape[0], dtype=torch.float32, device=points.device)
    L = _lib.lib()
    with torch.cuda.device(points.device):
        _lib.check(L.aadg_sinkhorn_small_batched(_lib.ptr(points), _lib.ptr(prob), prob.shape[0],
                                                 points.shape[1], _lib.ptr(out), _lib.stream_ptr()))
    return out


def divergence_large(x, y, diameter=None):
    """One divergence on big clouds (streamed cost matrices). Returns (0-d tensor, n_eps)."""
    x, y = _check(x), _check(y)
    if x.shape[1] != y.shape[1]:
        raise ValueError("feature dimensions differ")
    L = _lib.lib()
    n, m, d = x.shape[0], y.shape[0], x.shape[1]
    ws = _lib.workspace(L.aadg_sinkhorn_large_workspace_bytes(n, m, d), x.device)
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    nit = np.zeros(1, np.int32)
    with torch.cuda.device(x.device):
        _lib.check(L.aadg_sinkhorn_large(_lib.ptr(x), n, _lib.ptr(y), m, d,
                                         float(diameter) if diameter else 0.0, _lib.ptr(out),
                                         nit.ctypes.data, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return out[0], int(nit[0])


def large_setup_only(x, y, diameter=None):
    """benchmark helper: norms + diameter + cost matrices of divergence_large, nothing else."""
    x, y = _check(x), _check(y)
    L = _lib.lib()
    n, m, d = x.shape[0], y.shape[0], x.shape[1]
    ws = _lib.workspace(L.aadg_sinkhorn_large_workspace_bytes(n, m, d), x.device)
    with torch.cuda.device(x.device):
        _lib.check(L.aadg_sinkhorn_large_setup(_lib.ptr(x), n, _lib.ptr(y), m, d, float(diameter) if diameter else 0.0,
                                               _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))


def divergence(x, y, diameter=None):
    x, y = _check(x), _check(y)
    if max(x.shape[0], y.shape[0]) <= small_max_points() and diameter is None:
        pts = torch.cat([x, y])
        return divergence_batched(pts, [[0, x.shape[0], x.shape[0], y.shape[0]]])[0]
    return divergence_large(x, y, diameter)[0]


class SamplesLoss:
    """geomloss.SamplesLoss call shape for the one configuration the reference uses."""

    def __init__(self, loss="sinkhorn", p=2, blur=.05, reach=None, diameter=None, scaling=.5, truncate=5,
                 cost=None, kernel=None, cluster_scale=None, debias=True, potentials=False, verbose=False,
                 backend="auto"):
        if loss != "sinkhorn" or p != 2 or blur != .05 or reach is not None or scaling != .5 \
                or not debias or potentials:
            raise NotImplementedError("only the reference's configuration is implemented: "
                                      "sinkhorn, p=2, blur=.05, scaling=.5, debiased, balanced")
        if cost is not None and "".join(cost.split()) != "".join(COSINE_COST.split()):
            raise NotImplementedError("only the cosine cost of search_dg.py:116 is implemented")
        if cost is None:
            raise NotImplementedError("the squared-Euclidean default cost is not on the reference's path")
        self.diameter = diameter

    def __call__(self, x, y):
        return divergence(x, y, self.diameter)

    forward = __call__


def _pair_order(nd):
    """the reference's call order for three domains is (1,2), (2,3), (1,3) (search_dg.py:158-160); otherwise lexicographic"""
    if nd == 3:
        return [(0, 1), (1, 2), (0, 2)]
    return [(a, b) for a in range(nd) for b in range(a + 1, nd)]


def _diversity_rewards_general(f, dc, M, rewards, pairs):
    """clouds of any size: the `[j::M]` / argmax split on the host side of the ABI (one device->host read of the domain
    ids), every pair through `divergence` (one-launch kernel up to small_max_points(), streamed large path above)."""
    nd = dc.shape[1]
    dom = torch.argmax(dc, dim=1).cpu().numpy()
    order = _pair_order(nd)
    for j in range(M):
        rows = np.arange(j, f.shape[0], M)
        clouds = []
        for k in range(nd):
            idx = rows[dom[rows] == k]
            if idx.size == 0:
                raise RuntimeError("diversity_rewards: policy %d has no sample of domain %d in this batch (the "
                                   "reference's sinkhorn(empty, .) is undefined)" % (j, k))
            clouds.append(f[torch.as_tensor(idx, device=f.device)])
        vals = [divergence(clouds[a], clouds[b]) for a, b in order]
        pairs[j] = torch.stack(vals)
        if nd == 3:
            rewards[j] += (vals[0] + vals[2]) + vals[1]        # (d12 + d13) + d23
        else:
            t = vals[0]
            for v in vals[1:]:
                t = t + v
            rewards[j] += t
    return rewards, pairs


def diversity_rewards(domain_feature, domain_code, M, rewards=None, max_cloud=None):
    """search_dg.py:150-162.  domain_feature float32 [n,d], domain_code float32 [n,D] (soft one-hot), rows ordered
    (b*D+d)*M+j.  Adds (d12+d13)+d23 to rewards[j] (created if None); returns (rewards [M], pair_values [M, pairs]).

    max_cloud = an upper bound on the number of rows of one (policy, domain) cloud, when the caller knows it (the
    search engine does: it is the largest per-domain source-image count of the global batch); default n // M, the worst
    case.  Up to small_max_points() everything is ONE launch without a host sync; above it the clouds go through the
    general path (streamed large-N kernels).  An EMPTY cloud makes the fused kernel store NaN and flag its status word:
    `check_rewards(rewards)` at the epoch boundary raises on it (a NaN reward must never reach the controller)."""
    f, dc = _check(domain_feature), _check(domain_code)
    n, d = f.shape
    nd = dc.shape[1]
    if dc.shape[0] != n:
        raise ValueError("domain_code rows != feature rows")
    if rewards is None:
        rewards = torch.zeros(M, dtype=torch.float32, device=f.device)
    pairs = torch.empty((M, nd * (nd - 1) // 2), dtype=torch.float32, device=f.device)
    if max_cloud is None:
        max_cloud = (n + M - 1) // M
    if max_cloud > small_max_points():
        return _diversity_rewards_general(f, dc, M, rewards, pairs)
    L = _lib.lib()
    ws = _lib.workspace(L.aadg_sinkhorn_rewards_workspace_bytes(M, nd), f.device)
    with _lib.on_device(f.device):
        _lib.check(L.aadg_sinkhorn_diversity_rewards(_lib.ptr(f), _lib.ptr(dc), n, d, nd, M, _lib.ptr(rewards),
                                                     _lib.ptr(pairs), _lib.ptr(ws), ws.numel(),
                                                     _lib.stream_ptr()))
    return rewards, pairs


def check_rewards(rewards):
    """raise if a reward is not finite (an empty or oversize domain cloud reached the fused kernel)"""
    if not bool(torch.isfinite(rewards).all()):
        raise RuntimeError("diversity rewards are not finite: a (policy, domain) cloud was empty or larger than %d "
                           "points in some step; rewards=%s" % (small_max_points(), rewards.tolist()))
    return rewards


def normalize_rewards(rewards):
    """search_dg.py:214 (raises instead of normalising NaN rewards into the controller update)."""
    check_rewards(rewards)
    return (rewards - torch.mean(rewards)) / (torch.std(rewards) + 1e-5)
