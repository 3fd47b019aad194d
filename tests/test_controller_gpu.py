"""GPU: the one-launch controller (csrc/controller.cu) against the torch mirror of models/controller.py (same
weights): log-probabilities, entropies, probabilities, PPO gradients; counter-based sampling against its numpy oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class Cfg:
    class CONTROLLER:
        EXCLUDE_OPS = []
        L = 2
        T = 2.0
        C = 2.5
        NUM_MAGS = 10
        EXCLUDE_OPS_NUM = 0
        PENALTY = 0.0
    SEED = 0


@pytest.fixture(scope="module")
def pair():
    assert torch.cuda.is_available()
    from aadg_b200.host.controller import Controller, FusedController
    torch.manual_seed(5)
    ref = Controller(Cfg).cuda()
    with torch.no_grad():
        for p in ref.parameters():           # larger weights than the 0.1 init: sharper, more varied distributions
            p.mul_(4.0)
    fused = FusedController(Cfg, seed=77).cuda()
    fused.load_state_dict(ref.state_dict())
    return ref, fused


def test_evaluate_matches_torch_and_backward(pair):
    ref, fused = pair
    m = 6
    g = torch.Generator().manual_seed(1)
    steps = ref.Q * ref.L * 2
    pol = torch.stack([torch.randint(0, ref.NUM_MAGS if t & 1 else ref.NUM_OPS, (m,), generator=g) for t in range(steps)], 1).cuda()
    want = ref.evaluate(pol, m)
    got = fused.evaluate(pol, m)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5), (got, want)
    w = torch.randn(m, device="cuda")
    ref.zero_grad()
    fused.zero_grad()
    (want * w).sum().backward()
    (got * w).sum().backward()
    for (n1, p1), (n2, p2) in zip(ref.named_parameters(), fused.named_parameters()):
        assert n1 == n2 and p2.grad is not None
        scale = p1.grad.abs().max().item() + 1e-8
        err = (p1.grad - p2.grad).abs().max().item()
        assert err <= 1e-4 * scale + 1e-7, (n1, err, scale)


def test_sample_shapes_consistency_and_oracle(pair):
    from oracle.controller_sampling import sample_actions
    ref, fused = pair
    m = 6
    fused.calls = 3
    pol, op_probs, mag_probs, logp, ent = fused.sample(m)
    steps = ref.Q * ref.L * 2
    assert pol.shape == (m, steps) and pol.dtype == torch.int64
    assert (pol[:, 0::2] >= 0).all() and (pol[:, 0::2] < ref.NUM_OPS).all() and (pol[:, 1::2] < ref.NUM_MAGS).all()
    assert op_probs.shape == (ref.NUM_OPS,) and mag_probs.shape == (ref.NUM_MAGS,)
    assert abs(op_probs.sum().item() - 1) < 1e-5 and abs(mag_probs.sum().item() - 1) < 1e-5
    # the log-probabilities / entropies returned with the sample are those of the torch module for these actions
    assert torch.allclose(logp, ref.evaluate(pol, m), rtol=1e-5, atol=1e-5)
    ents = []

    def choose(step, kind, lp):
        ents.append(-(lp * lp.exp()).sum(1))
        return pol[:, step]
    ref._walk(m, choose)
    assert torch.allclose(ent, torch.stack(ents, -1).sum(-1), rtol=1e-5, atol=1e-5)
    # same (seed, call) -> same draw; next call -> a different one
    fused.calls = 3
    pol2 = fused.sample(m)[0]
    assert torch.equal(pol, pol2)
    assert not torch.equal(pol, fused.sample(m)[0])
    # oracle replay of the categorical draws from the kernel's own probabilities (bit-exact policy sampling)
    _, _, probs, _, pol3 = fused._walk(pol.clone(), mode=1, want_probs=True)
    want = sample_actions(probs.cpu().numpy(), ref.NUM_OPS, ref.NUM_MAGS, fused.seed, 3)
    assert np.array_equal(pol.cpu().numpy(), want)


def test_ppo_update_runs_with_the_reference_loop(pair):
    """losses.py:127-157 shape: five evaluate / backward / Adam rounds on the fused controller reduce the surrogate."""
    ref, fused = pair
    from aadg_b200.host.controller import FusedController
    ctl = FusedController(Cfg, seed=1).cuda()
    ctl.load_state_dict(fused.state_dict())
    opt = torch.optim.Adam(ctl.parameters(), lr=3.5e-2)
    m = 6
    pol, _, _, logp, ent = ctl.sample(m)
    reward = torch.tensor([1.0, -0.5, 0.3, 0.8, -1.2, 0.1], device="cuda")
    prev = logp.detach()
    losses = []
    for _ in range(5):
        cur = ctl.evaluate(pol, m)
        ratios = torch.exp(cur - prev)
        loss = (-torch.min(ratios * reward, torch.clamp(ratios, 0.8, 1.2) * reward)).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
