"""Scalar losses around the hot path, reference call shapes (losses.py:52-68,96-157)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class CrossEntropy(nn.Module):
    """soft-target cross entropy (losses.py:52-68)."""

    def forward(self, input, target):
        return torch.mean(torch.sum(-target.detach() * F.log_softmax(input, dim=1), dim=1))


class Reinforce(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.penalty = cfg.CONTROLLER.PENALTY

    def register_optimizer(self, optimizer):
        self.optimizer = optimizer

    def forward(self, controller, policies, log_probs, entropies, reward):
        if not log_probs.requires_grad:
            # FusedController.sample() returns graph-less values (one kernel launch): rebuild log-probabilities and
            # entropies of the SAME policies with a graph through the torch mirror of the walk (once per epoch)
            from .controller import evaluate_with_entropy
            log_probs, entropies = evaluate_with_entropy(controller, policies, reward.size(0))
        score = torch.mean(-log_probs * reward)
        ent = torch.mean(entropies)
        loss = score - self.penalty * ent
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss, score, ent


class ProximalPolicyOptimization(nn.Module):
    """clip 0.2, five inner updates (losses.py:117-157)."""

    def __init__(self, cfg):
        super().__init__()
        self.clip, self.n_updates_per_iteration, self.penalty = 0.2, 5, cfg.CONTROLLER.PENALTY

    def register_optimizer(self, optimizer):
        self.optimizer = optimizer

    def forward(self, controller, policies, log_probs, entropies, reward):
        prev = log_probs.detach()
        total = 0
        for _ in range(self.n_updates_per_iteration):
            ratios = torch.exp(controller.evaluate(policies, reward.size(0)) - prev)
            loss = (-torch.min(ratios * reward, torch.clamp(ratios, 1 - self.clip, 1 + self.clip) * reward)).mean()
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
            total = total + loss.detach()
        mean = total / self.n_updates_per_iteration
        return mean, mean, torch.mean(entropies)


def search_loss(config):
    if config.CONTROLLER.LOSS == "reinforce":
        return Reinforce(config)
    if config.CONTROLLER.LOSS == "ppo":
        return ProximalPolicyOptimization(config)
    raise NotImplementedError("{} is unavailable".format(config.CONTROLLER.LOSS))
