"""A run.py-style search loop on synthetic data (reference call stack: SURVEY.md 3.1): per epoch the controller
samples M policies (one launch), the engine runs the hot loop with them on batches gathered from device-resident
domain pools, rewards are normalised, the controller is updated with PPO (one forward + one backward launch per
round), the momentum discriminator is refreshed and a held-out batch is validated with Dice and HD95 on the GPU.

    python scripts/search_demo.py --epochs 2 --steps 3 --size 128 --backbone resnet18
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aadg_b200.data.policy import parse_policies  # noqa: E402
from aadg_b200.host.config import optic_search_config  # noqa: E402
from aadg_b200.data.pool import ResidentPools  # noqa: E402
from aadg_b200.host.controller import FusedController  # noqa: E402
from aadg_b200.host.losses import search_loss  # noqa: E402
from aadg_b200.host.search import SearchEngine  # noqa: E402
from aadg_b200.nn import DeepLabV3Plus  # noqa: E402
from aadg_b200.nn.network import dice_from_counts  # noqa: E402
from aadg_b200.ops import metrics  # noqa: E402
from aadg_b200.ops import nn as K  # noqa: E402
from aadg_b200.ops import u8 as U8  # noqa: E402
from aadg_b200.synth import fundus_batch  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--items", type=int, default=2)
    ap.add_argument("--backbone", default="resnet18")
    a = ap.parse_args(argv)
    cfg = optic_search_config(a.backbone)
    dev = torch.device("cuda")
    torch.manual_seed(cfg.SEED)
    M, D = cfg.CONTROLLER.M, len(cfg.DATASET.DG.TRAIN)
    model = DeepLabV3Plus(encoder_name=cfg.MODEL.BACKBONE, encoder_weights=None, in_channels=3, classes=2,
                          aux_params=dict(pooling="avg"))
    controller = FusedController(cfg, seed=cfg.SEED).to(dev)
    controller_opt = torch.optim.Adam(controller.parameters(), lr=0.00035)      # scheduler.py:7
    criterion = search_loss(cfg)
    criterion.register_optimizer(controller_opt)
    eng = SearchEngine(model, n_domains=D, M=M, lr=cfg.TRAIN.LR, weight_decay=cfg.TRAIN.WD, crop=a.size)
    # data/optic.py image pools, resident on the device: 6 synthetic images per source domain + a held-out domain
    pool_imgs, pool_masks = {}, {}
    for d in range(D):
        pool_imgs["domain%d" % d], pool_masks["domain%d" % d] = fundus_batch(6, a.size, a.size, seed=100 + d)
    pools = ResidentPools(pool_imgs, pool_masks, device=dev)
    val_imgs, val_masks = fundus_batch(4, a.size, a.size, seed=999)
    val_x, val_t = U8.normalize_to_tensor(torch.from_numpy(val_imgs).to(dev), torch.from_numpy(val_masks).to(dev))
    rng = np.random.RandomState(cfg.SEED)
    history = []
    for epoch in range(a.epochs):
        policies, op_probs, mag_probs, log_probs, entropies = controller(M)            # search_dg.py:339
        parsed = parse_policies(policies.cpu().numpy(), cfg)                           # search_dg.py:340
        eng.set_policies(parsed, epoch=epoch)                                          # search_dg.py:341
        for step in range(a.steps):                                                    # search_dg.train()
            imgs, masks, domains = pools.batch(a.items, rng)                           # data/optic.py:78-90
            out = eng.step(imgs, masks, domains)
        rewards = eng.normalized_rewards()                                             # search_dg.py:214
        eng.end_epoch()                                                                # search_dg.py:346
        loss, score, ent = criterion(controller, policies, log_probs, entropies, rewards)   # search_dg.py:347
        # validate(): search_dg.py:230-267 -- hard masks at 0.75, samplewise Dice and HD95 per class, all on the GPU
        model.eval()
        logits, _ = model(val_x)
        model.train()
        seg_hard = (torch.sigmoid(logits) > 0.75)
        hd = metrics.validation_hd95(seg_hard, val_t > 0.5)
        history.append(dict(epoch=epoch, seg_loss=float(out["seg_loss"]), dis_loss=float(out["dis_loss"]),
                            val_hd95=hd.cpu().numpy().round(3).tolist(),
                            dice=out["dice"].cpu().numpy().round(4).tolist(), rewards=rewards.cpu().numpy().round(3).tolist(),
                            controller_loss=float(loss)))
        print(history[-1])
    return history


if __name__ == "__main__":
    main()
