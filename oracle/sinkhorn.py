"""ORACLE (test infrastructure, never on the product path).  **Parity unpinned.**

NumPy restatement of `geomloss==0.2.4` `SamplesLoss("sinkhorn", cost=<cosine KeOps formula>,
backend="online")` as the reference calls it (search_dg.py:116,158-160; search_dg_2d.py:116,159-161;
requirements.txt:1,11).  geomloss / pykeops are third-party, not vendored under /root/reference and
not installable here (no network), so this follows their published algorithm
(geomloss/sinkhorn_divergence.py: max_diameter, epsilon_schedule, scaling_parameters, sinkhorn_loop,
sinkhorn_cost; geomloss/sinkhorn_samples.py: softmin_online, keops_lse; geomloss/samples_loss.py:
SamplesLoss.process_args / generate_weights) with the defaults that apply at the call site:
p=2, blur=0.05, reach=None (balanced, dampening 1), diameter=None, scaling=0.5, debias=True,
potentials=False, uniform weights.  No golden vector exists for this path (the reference has no
tests and geomloss is absent), hence "parity unpinned"; self-consistency pins are in
tests/test_oracle_sinkhorn.py (symmetry, S(x,x)=0, agreement with an independent dense
log-domain Sinkhorn run to convergence at the final temperature).

dtype=np.float64 is the truth the CUDA kernels are compared with; dtype=np.float32 mimics KeOps'
fp32 arithmetic and is used to calibrate the tolerance.
"""
import numpy as np

BLUR, P, SCALING = 0.05, 2, 0.5


def cosine_cost(x, y, dtype=np.float64):
    """KeOps formula '( IntCst(1) - (X | Y) / ( Norm2(X) * Norm2(Y) ) )' — search_dg.py:116.
    No epsilon: an all-zero row gives NaN, like the reference."""
    x = x.astype(dtype)
    y = y.astype(dtype)
    nx = np.sqrt((x * x).sum(1))
    ny = np.sqrt((y * y).sum(1))
    with np.errstate(invalid="ignore", divide="ignore"):
        return (dtype(1) - (x @ y.T) / (nx[:, None] * ny[None, :])).astype(dtype)


def max_diameter(x, y):
    """geomloss max_diameter: norm of the bounding-box diagonal of the FEATURES, in the input dtype
    (the reference feeds float32 tensors and calls .item())."""
    mins = np.minimum(x.min(0), y.min(0))
    maxs = np.maximum(x.max(0), y.max(0))
    d = (maxs - mins).astype(x.dtype)
    return float(np.sqrt((d * d).sum(dtype=x.dtype)))


def epsilon_schedule(diameter, blur=BLUR, p=P, scaling=SCALING):
    """geomloss epsilon_schedule: [diam^p] + exp(arange(p ln diam, p ln blur, p ln scaling)) + [blur^p]."""
    return ([diameter ** p]
            + [float(np.exp(e)) for e in np.arange(p * np.log(diameter), p * np.log(blur), p * np.log(scaling))]
            + [blur ** p])


def softmin(eps, C, f, dtype):
    """softmin_online: -eps * LSE_j( f_j - C_ij * (1/eps) ), 1/eps rounded to `dtype` like
    torch.Tensor([1/eps]).type_as(x)."""
    pinv = dtype(1.0 / eps)
    v = f[None, :] - C * pinv
    m = v.max(1)
    lse = m + np.log(np.exp(v - m[:, None]).sum(1, dtype=dtype))
    return (dtype(-eps) * lse).astype(dtype)


def sinkhorn_divergence(x, y, dtype=np.float64, diameter=None, return_info=False):
    """Debiased Sinkhorn divergence S_eps(alpha, beta) with uniform weights.  x [N,d], y [M,d]."""
    n, m = len(x), len(y)
    C_xx, C_yy = cosine_cost(x, x, dtype), cosine_cost(y, y, dtype)
    C_xy = cosine_cost(x, y, dtype)
    C_yx = C_xy.T
    if diameter is None:
        diameter = max_diameter(x, y)
    eps_s = epsilon_schedule(diameter)
    a_log = np.full(n, np.log(dtype(1) / dtype(n)), dtype)   # log_weights(ones/N)
    b_log = np.full(m, np.log(dtype(1) / dtype(m)), dtype)
    eps = eps_s[0]
    a_x = softmin(eps, C_xx, a_log, dtype)
    b_y = softmin(eps, C_yy, b_log, dtype)
    a_y = softmin(eps, C_yx, a_log, dtype)
    b_x = softmin(eps, C_xy, b_log, dtype)
    for eps in eps_s:
        e = dtype(eps)
        at_x = softmin(eps, C_xx, a_log + a_x / e, dtype)
        bt_y = softmin(eps, C_yy, b_log + b_y / e, dtype)
        at_y = softmin(eps, C_yx, a_log + b_x / e, dtype)
        bt_x = softmin(eps, C_xy, b_log + a_y / e, dtype)
        a_x, b_y = dtype(.5) * (a_x + at_x), dtype(.5) * (b_y + bt_y)
        a_y, b_x = dtype(.5) * (a_y + at_y), dtype(.5) * (b_x + bt_x)
    e = dtype(eps)
    a_x = softmin(eps, C_xx, a_log + a_x / e, dtype)
    b_y = softmin(eps, C_yy, b_log + b_y / e, dtype)
    a_y, b_x = softmin(eps, C_yx, a_log + b_x / e, dtype), softmin(eps, C_xy, b_log + a_y / e, dtype)
    alpha = np.full(n, dtype(1) / dtype(n), dtype)
    beta = np.full(m, dtype(1) / dtype(m), dtype)
    s = float(np.dot(alpha, b_x - a_x) + np.dot(beta, a_y - b_y))
    if return_info:
        return s, dict(diameter=diameter, eps=eps_s, a_x=a_x, b_y=b_y, a_y=a_y, b_x=b_x)
    return s


def diversity_rewards(domain_feature, dc, M, dtype=np.float64, n_domains=3):
    """search_dg.py:150-162: for each policy j, rows j::M are split by argmax(dc) into the domain
    clouds and the three pairwise divergences are summed (d12 + d13 + d23).  Returns ([M] rewards
    increment, [M, pairs] individual values) — pairs ordered (1,2), (2,3), (1,3) like the calls."""
    feat = np.asarray(domain_feature)
    dom = np.argmax(np.asarray(dc), axis=1)
    inc = np.zeros(M, np.float64)
    vals = np.zeros((M, 3), np.float64)
    for j in range(M):
        sub, sd = feat[j::M], dom[j::M]
        clouds = [sub[sd == k] for k in range(n_domains)]
        d12 = sinkhorn_divergence(clouds[0], clouds[1], dtype)
        d23 = sinkhorn_divergence(clouds[1], clouds[2], dtype)
        d13 = sinkhorn_divergence(clouds[0], clouds[2], dtype)
        vals[j] = (d12, d23, d13)
        inc[j] = (d12 + d13) + d23
    return inc, vals


def normalize_rewards(rewards):
    """search_dg.py:214: (r - mean) / (std + 1e-5), torch.std = unbiased."""
    r = np.asarray(rewards, np.float64)
    return (r - r.mean()) / (r.std(ddof=1) + 1e-5)


def dense_sinkhorn_reference(x, y, eps, iters=5000):
    """Independent check: plain alternating (non-symmetrised) log-domain Sinkhorn to convergence at a
    single temperature, float64; returns the debiased divergence  OT(a,b) - (OT(a,a)+OT(b,b))/2."""
    def ot(C, n, m):
        a = np.full(n, -np.log(n))
        b = np.full(m, -np.log(m))
        f, g = np.zeros(n), np.zeros(m)
        for _ in range(iters):
            v = (g + eps * b)[None, :] - C
            mx = v.max(1)
            f = -(mx + eps * np.log(np.exp((v - mx[:, None]) / eps).sum(1)))
            v = (f + eps * a)[:, None] - C
            mx = v.max(0)
            g = -(mx + eps * np.log(np.exp((v - mx[None, :]) / eps).sum(0)))
        return f.mean() + g.mean()
    Cxy, Cxx, Cyy = cosine_cost(x, y), cosine_cost(x, x), cosine_cost(y, y)
    return ot(Cxy, len(x), len(y)) - 0.5 * (ot(Cxx, len(x), len(x)) + ot(Cyy, len(y), len(y)))
