"""GPU: the whole search loop (controller -> policies -> hot loop -> rewards -> PPO) on synthetic data, and
one search step checked piece by piece against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import statistical

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_search_demo_two_epochs():
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import search_demo
    hist = search_demo.main(["--epochs", "2", "--steps", "2", "--size", "64", "--items", "2"])
    assert len(hist) == 2
    for h in hist:
        assert np.isfinite(h["seg_loss"]) and np.isfinite(h["dis_loss"]) and np.isfinite(h["controller_loss"])
        assert len(h["rewards"]) == 6 and abs(float(np.mean(h["rewards"]))) < 1e-3      # z-normalised


def test_step_rewards_match_oracle_on_the_same_features():
    """the rewards the engine accumulates equal the oracle's on the engine's own discriminator features."""
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine
    from aadg_b200.nn import DeepLabV3Plus
    from aadg_b200.ops import sinkhorn as SK
    from aadg_b200.synth import fundus_batch, random_policies
    from oracle import sinkhorn as OS

    class Cfg:
        class CONTROLLER:
            EXCLUDE_OPS = []
            L = 2
            NUM_MAGS = 10
            EXCLUDE_OPS_NUM = 0
        SEED = 0
    model = DeepLabV3Plus(encoder_name="resnet18", classes=2)
    eng = SearchEngine(model, n_domains=3, M=6, crop=64)
    eng.set_policies(parse_policies(random_policies(seed=3), Cfg), epoch=0)
    imgs, masks = fundus_batch(12, 64, 64, seed=8)
    captured = {}
    orig = SK.diversity_rewards

    def spy(feat, dc, M, rewards=None, max_cloud=None):
        captured["feat"], captured["dc"] = feat.cpu().numpy(), dc.cpu().numpy()
        assert max_cloud == 4                     # 12 sources over 3 domains
        return orig(feat, dc, M, rewards, max_cloud=max_cloud)
    import aadg_b200.host.search as S
    S.SK.diversity_rewards = spy
    try:
        eng.step(torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda(), [i % 3 for i in range(12)])
    finally:
        S.SK.diversity_rewards = orig
    want, _ = OS.diversity_rewards(captured["feat"], captured["dc"], 6)
    assert np.allclose(eng.rewards.cpu().numpy(), want, rtol=1e-4)
    assert captured["feat"].shape == (72, 128) and captured["dc"].shape == (72, 3)


def test_pretrain_step_learns_without_policies():
    """warm-up phase (search_dg.py pretrain): un-augmented images through scale/crop + normalise, segmentation and
    discriminator steps; the loss falls over a few steps on one batch"""
    from aadg_b200.host.search import SearchEngine
    from aadg_b200.nn import DeepLabV3Plus
    from aadg_b200.synth import fundus_batch
    model = DeepLabV3Plus(encoder_name="resnet18", classes=2)
    eng = SearchEngine(model, n_domains=3, M=6, crop=64)
    imgs, masks = fundus_batch(6, 64, 64, seed=4)
    x, m = torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda()
    losses = []
    for _ in range(8):
        out = eng.pretrain_step(x, m, [0, 1, 2, 0, 1, 2])
        assert out["n_images"] == 6
        losses.append(float(out["seg_loss"]))
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    assert float(eng.rewards.abs().sum()) == 0.0


class _Cfg:
    class CONTROLLER:
        EXCLUDE_OPS = []
        L = 2
        NUM_MAGS = 10
        EXCLUDE_OPS_NUM = 0
        PENALTY = 0.00001
        LOSS = "reinforce"
        T = 2
        C = 2.5
    SEED = 0


def test_warmup_to_search_transition():
    """search_dg.py:336-337 + scheduler.py:11: after the warm-up steps the first set_policies() copies the live
    discriminator into its EMA twin and drops the model's learning rate by 10x (the discriminator's stays)."""
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine
    from aadg_b200.nn import DeepLabV3Plus
    from aadg_b200.synth import fundus_batch, random_policies
    model = DeepLabV3Plus(encoder_name="resnet18", classes=2)
    eng = SearchEngine(model, n_domains=3, M=6, crop=64, lr=1e-3)
    imgs, masks = fundus_batch(6, 64, 64, seed=4)
    x, m = torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda()
    for _ in range(3):
        eng.pretrain_step(x, m, [0, 1, 2, 0, 1, 2])
    dis = eng.discriminator
    assert not torch.equal(dis.dis[0].weight, dis.mom_dis[0].weight)       # the live branch trained, the twin did not
    assert abs(float(model.store.hyper[0]) - 1e-3) < 1e-9
    eng.set_policies(parse_policies(random_policies(seed=3), _Cfg), epoch=3)
    for q, k in dis._pairs():
        assert torch.equal(q, k)
    assert abs(float(model.store.hyper[0]) - 1e-4) < 1e-9 and abs(eng.lr - 1e-4) < 1e-12
    assert eng.dis_optimizer.param_groups[0]["lr"] == 1e-3
    out = eng.step(x, m, [0, 1, 2, 0, 1, 2])
    assert np.isfinite(float(out["seg_loss"])) and torch.isfinite(eng.normalized_rewards()).all()
    eng.set_policies(parse_policies(random_policies(seed=4), _Cfg), epoch=4)     # idempotent: no second drop
    assert abs(float(model.store.hyper[0]) - 1e-4) < 1e-9


@statistical
def test_cuda_graph_step_matches_eager_step():
    """the captured step (graph=True) and the eager step walk the same trajectory: same decisions, same dropout seeds
    and Adam step counts from device memory; differences are the summation order of the gradient atomics"""
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine
    from aadg_b200.nn import DeepLabV3Plus
    from aadg_b200.synth import fundus_batch, random_policies
    imgs, masks = fundus_batch(6, 64, 64, seed=11)
    x, m = torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda()
    curves = []
    for graph in (False, False, True):      # two eager runs measure the engine's own run-to-run spread
        model = DeepLabV3Plus(encoder_name="resnet18", classes=2, seed=3)
        eng = SearchEngine(model, n_domains=3, M=6, crop=64, graph=graph, seed=21)
        eng.set_policies(parse_policies(random_policies(seed=3), _Cfg), epoch=0)
        losses = [float(eng.step(x, m, [0, 1, 2, 0, 1, 2])["seg_loss"]) for _ in range(6)]
        curves.append((np.array(losses), eng.rewards.cpu().numpy().copy(), int(model.store.step_dev.item()), model.steps,
                       int(model.seed_dev.item())))
        if graph:
            assert len(eng._graphs) == 1 and list(eng._graphs.values())[0][0] is not None
            assert list(eng._graphs.values())[0][0].launches > 100
    (le, re_, se, te, de), (l2, r2, _, _, _), (lg, rg, sg, tg, dg) = curves
    rel = lambda u, v: float(np.max(np.abs(u - v) / np.abs(v)))       # noqa: E731
    print("GRAPH vs EAGER losses", le.tolist(), lg.tolist(), "| loss rel: graph-vs-eager %.2e eager-vs-eager %.2e | rewards rel: "
          "graph-vs-eager %.2e eager-vs-eager %.2e" % (rel(lg, le), rel(l2, le), rel(rg, re_), rel(r2, re_)))
    assert (se, te, de) == (sg, tg, dg) == (6, 6, 0x5EED0000 + 3 + 6)
    # step 0 runs eagerly in both engines; the engine is not bitwise reproducible (fp32 atomics in the batch-norm
    # statistics, re-quantised by bf16 storage), so the yardstick is the spread of two EAGER runs (x4 + small floors):
    # the rewards are Sinkhorn divergences between clouds of two points per domain, the most sensitive output there is
    assert abs(le[0] - lg[0]) <= 1e-3 * abs(le[0])
    # (both sides of each comparison are single draws of the same noise -- measured over several boxes: loss 0.4-3.1e-2,
    # rewards 0.05-0.27 for graph-vs-eager and eager-vs-eager alike -- so the floors sit 2-3x above the largest value seen)
    assert rel(lg, le) <= max(4 * rel(l2, le), 8e-2), (rel(lg, le), rel(l2, le))
    assert rel(rg, re_) <= max(4 * rel(r2, re_), 0.6), (rel(rg, re_), rel(r2, re_))
    assert lg[-1] < lg[0]


def test_diversity_rewards_general_path_for_large_clouds():
    """clouds larger than the one-launch kernel's limit go through the streamed path: same rewards as the
    float64 oracle; an empty domain raises instead of poisoning the rewards with NaN"""
    from aadg_b200.ops import sinkhorn as SK
    from aadg_b200.synth import feature_cloud
    from oracle import sinkhorn as OS
    M, per = 2, SK.small_max_points() + 6
    rng = np.random.RandomState(0)
    feats, dcs = [], []
    for i in range(3 * per):
        d = i % 3
        for j in range(M):
            feats.append(feature_cloud(1, 32, d, seed=1000 * j + i)[0] + 0.05 * j)
            dc = np.full(3, 0.05, np.float32)
            dc[d] = 0.9
            dcs.append(dc)
    f = torch.from_numpy(np.stack(feats).astype(np.float32)).cuda()
    dc = torch.from_numpy(np.stack(dcs)).cuda()
    rewards, pairs = SK.diversity_rewards(f, dc, M)
    want, _ = OS.diversity_rewards(f.cpu().numpy(), dc.cpu().numpy(), M)
    assert np.allclose(rewards.cpu().numpy(), want, rtol=2e-4), (rewards, want)
    dc2 = dc.clone()
    dc2[:, 2] = 0.0                                   # nobody belongs to domain 2 any more
    with pytest.raises(RuntimeError):
        SK.diversity_rewards(f, dc2, M)
    small = SK.diversity_rewards(f[:12], dc2[:12], M)[0]          # fused kernel: NaN + status, caught at the boundary
    with pytest.raises(RuntimeError):
        SK.normalize_rewards(small)


def test_reinforce_loss_runs_on_the_fused_controller():
    """losses.py:96-114 with FusedController: graph-less sampled log-probs are rebuilt with a graph"""
    from aadg_b200.host.controller import FusedController
    from aadg_b200.host.losses import search_loss
    ctl = FusedController(_Cfg, seed=2).cuda()
    crit = search_loss(_Cfg)
    opt = torch.optim.Adam(ctl.parameters(), lr=3.5e-4)
    crit.register_optimizer(opt)
    pol, _, _, logp, ent = ctl(6)
    before = [p.detach().clone() for p in ctl.parameters()]
    loss, score, e = crit(ctl, pol, logp, ent, torch.linspace(-1, 1, 6, device="cuda"))
    assert torch.isfinite(loss) and any(not torch.equal(a, b) for a, b in zip(before, ctl.parameters()))
    sd = ctl.state_dict()
    assert sd["_extra_state"]["calls"] == 1
    other = FusedController(_Cfg, seed=2).cuda()
    other.load_state_dict(sd)
    assert other.calls == 1


def test_policy_call_shape_replays_the_reference_draws():
    """data/policy.py:15-43 call shapes: `Policy(policy)(img, mask)` draws from the GLOBAL random / np.random state in the
    reference's order, so it equals the bank applied to the rows `replay_sample` resolves under the same seeds (the
    path pinned to reference-run goldens in test_u8_gpu.py)."""
    import random
    from aadg_b200.data import decisions as D
    from aadg_b200.data.policy import MultiPolicy, Policy, parse_policies
    from aadg_b200.ops import u8 as U8
    from aadg_b200.synth import fundus_batch, random_policies
    parsed = parse_policies(random_policies(seed=9), _Cfg)
    imgs, masks = fundus_batch(1, 64, 80, seed=2)
    x, m = torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda()
    random.seed(5)
    np.random.seed(5)
    pols = [Policy(p) for p in parsed]
    got = [pol(x[0], m[0]) for pol in pols]
    py, npr = random.Random(5), np.random.RandomState(5)
    rows, _ = D.replay_sample(parsed, 0, x.shape[2], x.shape[1], x.shape[2], (1, 1.5), py, npr, D.PolicyState(len(parsed)),
                              scale_crop=False)
    want, want_m = U8.apply_policy(x, m, rows, want_masks=True)
    for j, (gi, gm) in enumerate(got):
        assert gi.shape == x[0].shape and torch.equal(gi, want[j]) and torch.equal(gm, want_m[j])
    multi = MultiPolicy(parsed, rng=(random.Random(5), np.random.RandomState(5)))
    outs = multi(x[0])
    assert len(outs) == len(parsed) and all(torch.equal(o[0], want[j]) and o[1] is None for j, o in enumerate(outs))


def test_resident_pool_step_reads_sources_in_place():
    """SURVEY 8f N4 on the device: a step fed from ResidentPools by INDEX (zero-copy: the uint8 bank reads the pool
    entries in place) produces bit-identical augmented batches to the same step fed the gathered sources, the gather
    itself equals numpy fancy indexing in the reference's collate order (b*D + d), and the whole step runs from it."""
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.data.pool import ResidentPools
    from aadg_b200.host.search import SearchEngine
    from aadg_b200.nn import DeepLabV3Plus
    from aadg_b200.synth import fundus_batch, random_policies
    sizes = {"A": 5, "B": 4, "C": 6}
    imgs, msks = {}, {}
    for k, (name, n) in enumerate(sizes.items()):
        imgs[name], msks[name] = fundus_batch(n, 64, 64, seed=40 + k)
    pools = ResidentPools(imgs, msks, device="cuda")
    np.random.seed(5)
    idx = pools.sample_indices(2)                                  # [B=2, D=3]
    gi, gm, dom = pools.gather(idx)
    want_i = np.stack([imgs[n][idx[b, d]] for b in range(2) for d, n in enumerate(sizes)])
    want_m = np.stack([msks[n][idx[b, d]] for b in range(2) for d, n in enumerate(sizes)])
    assert np.array_equal(gi.cpu().numpy(), want_i) and np.array_equal(gm.cpu().numpy(), want_m) and dom == [0, 1, 2] * 2
    flat, dom2 = pools.flat_indices(idx)
    assert dom2 == dom and np.array_equal(pools.images[torch.from_numpy(flat).cuda()].cpu().numpy(), want_i)
    parsed = parse_policies(random_policies(seed=3), _Cfg)
    for crop in (None, 64):
        model = DeepLabV3Plus(encoder_name="resnet18", classes=2, seed=3)
        eng = SearchEngine(model, n_domains=3, M=6, crop=crop, seed=21)
        eng.set_policies(parsed, epoch=0)
        rows, _ = eng.decision_rows(6, 64, 64)
        a_i, a_l, _ = eng._augment(gi, gm, rows, "search")
        rows_pool = rows.copy()
        rows_pool["src"] = flat[rows["src"]]
        b_i, b_l, _ = eng._augment(pools.images, pools.masks, rows_pool, "search")
        assert torch.equal(a_i, b_i) and torch.equal(a_l, b_l)
        out = eng.step(pools.images, pools.masks, dom, src_index=flat)
        assert np.isfinite(float(out["seg_loss"])) and out["n_images"] == 36
        out = eng.pretrain_step(pools.images, pools.masks, dom, src_index=flat)
        assert np.isfinite(float(out["seg_loss"])) and out["n_images"] == 6
    with pytest.raises(IndexError):
        eng.step(pools.images, pools.masks, dom, src_index=np.array([0, 1, 2, 3, 4, 99]))
