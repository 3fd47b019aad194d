"""GPU: the whole search loop (controller -> policies -> hot loop -> rewards -> PPO) on synthetic data, and
one search step checked piece by piece against the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

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

    def spy(feat, dc, M, rewards=None):
        captured["feat"], captured["dc"] = feat.cpu().numpy(), dc.cpu().numpy()
        return orig(feat, dc, M, rewards)
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
