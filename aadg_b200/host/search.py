"""One step of the reference's search loop (`search_dg.train`, search_dg.py:123-206) on the CUDA engine.

    augment (DGMultiPolicy + Normalize_dg/ToTensor, uint8 bank)  ->  model forward  ->  momentum
    discriminator features  ->  Sinkhorn diversity rewards  ->  BCE backward + Adam  ->  discriminator step

Row order of every per-image tensor is the reference's collate order (b*D + d)*M + j
(data/transform.py:323-340).  With torch.distributed initialised, each rank owns a shard of the source
images: gradients are summed with one NCCL all-reduce of the flat gradient buffer and the 128-d
discriminator features (+ domain codes) are all-gathered so every rank computes the same rewards."""
import random

import numpy as np
import torch
import torch.distributed as dist

from ..data import decisions as D
from ..ops import sinkhorn as SK
from ..ops import u8 as U8
from ..nn.network import dice_from_counts
from .discriminator import MomentumFeatureDiscriminator
from .losses import CrossEntropy


def gather_rows(*tensors):
    """all-gather row blocks rank-major: every rank ends with the same [world*n, ...] tensors (the
    cross-domain feature exchange feeding the Sinkhorn reward)."""
    world = dist.get_world_size()
    out = []
    for t in tensors:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t.contiguous())
        out.append(torch.cat(parts))
    return out


def average_(flat):
    """gradient all-reduce (sum) and division by the world size, in place."""
    dist.all_reduce(flat)
    flat.mul_(1.0 / dist.get_world_size())
    return flat


def shard_sources(n_sources, rank, world):
    """contiguous block of source-image indices owned by `rank` (units are independent)."""
    per = (n_sources + world - 1) // world
    return list(range(min(rank * per, n_sources), min((rank + 1) * per, n_sources)))


class SearchEngine:
    def __init__(self, model, n_domains=3, M=6, lr=1e-3, weight_decay=0.0, dataset="optic", seed=1023,
                 crop=None, scale_range=(1, 1.5)):
        self.model = model
        self.M, self.n_domains, self.dataset = M, n_domains, dataset
        self.lr, self.wd = lr, weight_decay
        dev = model.device
        enc_c = model.encoder.out_channels[-1]
        torch.manual_seed(seed)
        self.discriminator = MomentumFeatureDiscriminator(n_domains, enc_c).to(dev)
        self.discriminator.synchronize_parameters()
        self.dis_optimizer = torch.optim.Adam(self.discriminator.parameters(), lr=lr)
        self.dis_criterion = CrossEntropy()
        self.rewards = torch.zeros(M, dtype=torch.float32, device=dev)
        self.seed, self.epoch, self.step_idx = seed, 0, 0
        self.crop, self.scale_range = crop, scale_range
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.policies = None
        self._soft_rng = random.Random(seed * 7919 + self.rank)

    def set_policies(self, parsed_policies, epoch=None):
        """install the epoch's policies (search_dg.py:339-341) and reset the reward accumulator."""
        self.policies = parsed_policies
        assert len(parsed_policies) == self.M
        if epoch is not None:
            self.epoch = epoch
        self.step_idx = 0
        self.rewards.zero_()

    def decision_rows(self, n_src, width, height):
        rows, raws = D.philox_rows(self.policies, n_src, width, height, self.crop or width, self.scale_range,
                                   self.seed + 1000003 * self.rank, self.epoch, self.step_idx,
                                   scale_crop=self.crop is not None)
        return rows

    def domain_codes(self, src_domains):
        """ToTensor's random soft domain label (data/transform.py:260-274), one per source image,
        repeated for its M copies (data/transform.py:234)."""
        dc = np.stack([D.soft_label(self._soft_rng, int(d), self.n_domains) for d in src_domains])
        return np.repeat(dc, self.M, axis=0).astype(np.float32)

    def step(self, src_images, src_masks, src_domains, rows=None, dc=None):
        """src_images uint8 [S,H,W,3] (CUDA), src_masks uint8 [S,H,W], src_domains int [S] (host).
        Returns dict(seg_loss, dis_loss, dice [classes], n_images) of 0-d / small CUDA tensors."""
        s, h, w, _ = src_images.shape
        if rows is None:
            rows = self.decision_rows(s, w, h)
        if dc is None:
            dc = self.domain_codes(src_domains)
        dc_dev = torch.from_numpy(dc).to(src_images.device, non_blocking=True)
        if self.crop is not None:      # DGMultiPolicy -> DGRandomScaleCrop -> Normalize_dg -> ToTensor
            images, labels = U8.policy_scale_crop_normalize(src_images, src_masks, rows, self.crop, self.dataset)
        else:                          # DGMultiPolicy -> Normalize_dg -> ToTensor (one fused pass)
            images, labels = U8.policy_normalize(src_images, src_masks, rows, dataset=self.dataset)
        model = self.model
        model.store.zero_grad()
        out = model.loss_step(images, labels)
        feature = out["pooled"]
        # discriminator (search_dg.py:134-138)
        dis = self.discriminator
        _, domain_feature = dis(feature, momentum=True, return_feature=True)
        dis_loss = self.dis_criterion(dis(feature, momentum=False), dc_dev)
        # diversity reward over ALL ranks' features (search_dg.py:150-162)
        if self.world > 1:
            domain_feature, dc_all = gather_rows(domain_feature, dc_dev)
            average_(model.store.grads)
        else:
            dc_all = dc_dev
        SK.diversity_rewards(domain_feature, dc_all, self.M, self.rewards)
        model.store.adam_step(self.lr, weight_decay=self.wd)
        self.dis_optimizer.zero_grad()
        dis_loss.backward()
        if self.world > 1:
            for p in dis.parameters():
                if p.grad is not None:
                    average_(p.grad)
        self.dis_optimizer.step()
        self.step_idx += 1
        return dict(seg_loss=out["loss"], dis_loss=dis_loss.detach(), dice=dice_from_counts(out["counts"]),
                    n_images=images.shape[0])

    def pretrain_step(self, src_images, src_masks, src_domains):
        """One step of the warm-up phase (`pretrain`, search_dg.py:22-100): the un-augmented `sample['image']` (only
        DGRandomScaleCrop / Normalize_dg / ToTensor), segmentation step, live discriminator step on the detached pooled
        features; no policies, no rewards.  Same argument / return shapes as `step`."""
        s, h, w, _ = src_images.shape
        _, raws = D.philox_rows([[[]]], s, w, h, self.crop or w, self.scale_range, self.seed + 1000003 * self.rank,
                                self.epoch, self.step_idx, scale_crop=self.crop is not None)
        dc = np.stack([D.soft_label(self._soft_rng, int(d), self.n_domains) for d in src_domains]).astype(np.float32)
        dc_dev = torch.from_numpy(dc).to(src_images.device, non_blocking=True)
        if self.crop is not None:
            images, labels = U8.policy_scale_crop_normalize(src_images, src_masks, raws, self.crop, self.dataset)
        else:
            images, labels = U8.policy_normalize(src_images, src_masks, raws, dataset=self.dataset)
        model = self.model
        model.store.zero_grad()
        out = model.loss_step(images, labels)
        dis_loss = self.dis_criterion(self.discriminator(out["pooled"].detach(), momentum=False), dc_dev)
        if self.world > 1:
            average_(model.store.grads)
        model.store.adam_step(self.lr, weight_decay=self.wd)
        self.dis_optimizer.zero_grad()
        dis_loss.backward()
        if self.world > 1:
            for p in self.discriminator.parameters():
                if p.grad is not None:
                    average_(p.grad)
        self.dis_optimizer.step()
        self.step_idx += 1
        return dict(seg_loss=out["loss"], dis_loss=dis_loss.detach(), dice=dice_from_counts(out["counts"]),
                    n_images=images.shape[0])

    def normalized_rewards(self):
        """search_dg.py:214"""
        return SK.normalize_rewards(self.rewards)

    def end_epoch(self):
        self.discriminator.momentum_update()      # search_dg.py:346
