"""A run.py-style search loop on synthetic data (reference call stack: SURVEY.md 3.1): per epoch the controller
samples M policies, the engine runs the hot loop with them, rewards are normalised, the controller is updated
with PPO and the momentum discriminator is refreshed.

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
from aadg_b200.host.controller import Controller  # noqa: E402
from aadg_b200.host.losses import search_loss  # noqa: E402
from aadg_b200.host.search import SearchEngine  # noqa: E402
from aadg_b200.nn import DeepLabV3Plus  # noqa: E402
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
    controller = Controller(cfg).to(dev)
    controller_opt = torch.optim.Adam(controller.parameters(), lr=0.00035)      # scheduler.py:7
    criterion = search_loss(cfg)
    criterion.register_optimizer(controller_opt)
    eng = SearchEngine(model, n_domains=D, M=M, lr=cfg.TRAIN.LR, weight_decay=cfg.TRAIN.WD, crop=a.size)
    history = []
    for epoch in range(a.epochs):
        policies, op_probs, mag_probs, log_probs, entropies = controller(M)            # search_dg.py:339
        parsed = parse_policies(policies.cpu().numpy(), cfg)                           # search_dg.py:340
        eng.set_policies(parsed, epoch=epoch)                                          # search_dg.py:341
        for step in range(a.steps):                                                    # search_dg.train()
            imgs, masks = fundus_batch(a.items * D, a.size, a.size, seed=1000 * epoch + step)
            out = eng.step(torch.from_numpy(imgs).to(dev), torch.from_numpy(masks).to(dev),
                           [i % D for i in range(a.items * D)])
        rewards = eng.normalized_rewards()                                             # search_dg.py:214
        eng.end_epoch()                                                                # search_dg.py:346
        loss, score, ent = criterion(controller, policies, log_probs, entropies, rewards)   # search_dg.py:347
        history.append(dict(epoch=epoch, seg_loss=float(out["seg_loss"]), dis_loss=float(out["dis_loss"]),
                            dice=out["dice"].cpu().numpy().round(4).tolist(), rewards=rewards.cpu().numpy().round(3).tolist(),
                            controller_loss=float(loss)))
        print(history[-1])
    return history


if __name__ == "__main__":
    main()
