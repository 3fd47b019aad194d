"""Policy decode and application — the reference's `data/policy.py` call shapes on a batched GPU bank.

* `parse_policies(policies, config, logger)`  — same decode as data/policy.py:64-97 (bit-exact
  integer indexing): `name = augment_list()[p[i][2L*j+2k]]`, `mag = p[i][2L*j+2k+1]/(NUM_MAGS-1)`.
* `Policy` / `MultiPolicy` / `DGMultiPolicy`   — same constructors; `__call__` takes a *batch*
  sample: `sample['image']` uint8 `[S,H,W,3]` (CUDA tensor or array), `sample['label']` uint8
  `[S,H,W]`, and adds `sample['aug_images']` / `['aug_labels']`.  Random decisions come from a
  decision table (data/decisions.py) instead of global RNG state.
"""
import random as _random

import numpy as np

from .basic import augment_list
from . import decisions as _dec


def parse_policies(policies, config, logger=None):
    """int array [M, Q*L*2] -> [[[(op_name, mag)]*L]*Q]*M   (reference data/policy.py:64-97)."""
    exclude_ops = list(config.CONTROLLER.EXCLUDE_OPS)
    L = config.CONTROLLER.L
    num_mags = config.CONTROLLER.NUM_MAGS
    exclude_num = config.CONTROLLER.EXCLUDE_OPS_NUM

    ops = augment_list()
    if exclude_ops:
        ops = [op for op in ops if op[0] not in exclude_ops]
        if logger:
            logger.info(exclude_ops)
    elif exclude_num > 0:
        # data/policy.py:76-83: seeded shuffle, drop the head, remember it in the config
        for _ in range(exclude_num):
            seed = np.random.randint(0, 65536) * config.SEED
            _random.seed(seed)
            _random.shuffle(ops)
            dropped = ops.pop(0)
            config.CONTROLLER.EXCLUDE_OPS.append(dropped[0])
            if logger:
                logger.info(dropped[0])

    policies = np.asarray(policies)
    m, width = policies.shape
    q = width // (L * 2)
    parsed = []
    for i in range(m):
        subs = []
        for j in range(q):
            base = 2 * L * j
            subs.append([(ops[policies[i][base + 2 * k]][0], policies[i][base + 2 * k + 1] / (num_mags - 1))
                         for k in range(L)])
        parsed.append(subs)
    return parsed


def _apply_one_policy(policy_obj, img, mask, py, npr):
    """the body of `Policy.__call__` on CUDA tensors: one row per image resolved through the reference's draws"""
    import torch
    from ..ops import u8 as _u8  # CUDA bank; fails loudly when the extension is missing
    single = img.dim() == 3
    imgs = img[None] if single else img
    masks = None if mask is None else (mask[None] if single else mask)
    s, h, w, _ = imgs.shape
    rows = np.zeros(s, _dec.ROW_DTYPE)
    for i in range(s):
        policy_obj.calls += 1
        _dec.replay_policy_call(rows[i], policy_obj.policy, policy_obj.calls, i, w, h, py, npr)
        rows[i]["scale_w"], rows[i]["scale_h"] = w, h
    if masks is None:
        out, outm = _u8.apply_policy(imgs, None, rows), None
    else:
        out, outm = _u8.apply_policy(imgs, masks, rows, want_masks=True)
    policy_obj.last_rows = rows
    if single:
        return out[0], (None if outm is None else outm[0])
    return out, outm


class Policy:
    """One searched policy = Q sub-policies of L (op, mag) pairs (data/policy.py:7-30).

    `policy(img, mask)` keeps the reference's call shape on CUDA uint8 tensors (`[H,W,3]` + `[H,W]`, or a batch
    `[S,H,W,3]` + `[S,H,W]`): one randomly chosen sub-policy is applied by the CUDA bank.  The draws come from `rng`
    = (random.Random-like, RandomState-like); the default is the GLOBAL `random` / `np.random` state the reference
    uses, consumed in its order (queue draw, sub-policy choice, Cutout centre / mirror sign), so that a seeded
    reference `Policy` and this one make identical decisions."""

    def __init__(self, policy, rng=None):
        self.policy = policy
        self.calls = 0  # stands in for the CutMix queue length (see decisions.PolicyState)
        self.rng = rng
        self.last_rows = None

    def __call__(self, img, mask=None):
        py, npr = self.rng if self.rng is not None else (_random, np.random)
        return _apply_one_policy(self, img, mask, py, npr)


class MultiPolicy:
    """data/policy.py:33-43: `multi(img) -> [policy(img) for policy in policies]` (each entry an (img, mask) pair,
    mask None: the reference's MultiPolicy passes no mask and would raise inside Policy.__call__)."""

    def __init__(self, policies, rng=None):
        self.policies = [Policy(p, rng) for p in policies]

    def __call__(self, img, mask=None):
        return [policy(img, mask) for policy in self.policies]


class DGMultiPolicy:
    """M policies applied to every source image of a batch (data/policy.py:45-61).

    crop / scale_range / dataset select the fused DGRandomScaleCrop + Normalize_dg + ToTensor
    epilogue (data/transform.py:97-236); with crop=None the epilogue only normalises.
    rng: None -> Philox rows keyed (seed, epoch, step); or a (random.Random, RandomState) pair for
    reference replay."""

    def __init__(self, policies, crop=None, scale_range=(1, 1.5), dataset="optic", seed=1023,
                 rng=None):
        self.policies = [Policy(p) for p in policies]
        self.parsed = policies
        self.crop = crop
        self.scale_range = scale_range
        self.dataset = dataset
        self.seed = seed
        self.epoch = 0
        self.step = 0
        self.rng = rng
        self._state = _dec.PolicyState(len(policies))

    def set_epoch(self, epoch):
        self.epoch = epoch
        self.step = 0

    def rows_for(self, n_src, width, height):
        crop = self.crop if self.crop is not None else width
        sc = self.crop is not None
        if self.rng is None:
            rows, raws = _dec.philox_rows(self.parsed, n_src, width, height, crop, self.scale_range,
                                          self.seed, self.epoch, self.step, scale_crop=sc)
        else:
            py, npr = self.rng
            rr, rw = [], []
            for s in range(n_src):
                r, raw = _dec.replay_sample(self.parsed, s, width, height, crop, self.scale_range,
                                            py, npr, self._state, scale_crop=sc)
                rr.append(r)
                rw.append(raw)
            rows, raws = np.concatenate(rr), np.stack(rw)
        self.step += 1
        return rows, raws

    def __call__(self, sample):
        from ..ops import u8 as _u8  # CUDA bank; fails loudly when the extension is missing
        return _u8.apply_dg_multipolicy(self, sample)
