"""CPU: BASELINE config 1 — single-domain synthetic 128x128, DeepLabV3+/ResNet-18, fixed random augmentation,
2 train steps — as a plumbing check of the ORACLE pipeline (the product has no CPU path by design; this is the
CPU-runnable case the GPU engine is compared against)."""
import numpy as np
import torch

from aadg_b200.data import decisions as D
from aadg_b200.data.policy import parse_policies
from aadg_b200.host.config import get_config
from aadg_b200.synth import fundus_batch, random_policies
from oracle import u8_policy as OP
from oracle.segnet_torch import DeepLabV3PlusTorch


def test_two_train_steps_on_cpu():
    torch.manual_seed(0)
    torch.set_num_threads(4)
    cfg = get_config()
    imgs, masks = fundus_batch(2, 128, 128, seed=1)
    parsed = parse_policies(random_policies(m=1, seed=1), cfg)          # one fixed random policy
    rows, _ = D.philox_rows(parsed, 2, 128, 128, 128, (1, 1.5), seed=1, scale_crop=True)
    out = OP.apply_rows(imgs, masks, rows, crop=128, dataset="optic")
    x, y = torch.from_numpy(out["images"]), torch.from_numpy(out["labels"])
    assert x.shape == (2, 3, 128, 128) and y.shape == (2, 2, 128, 128)
    model = DeepLabV3PlusTorch("resnet18", 2).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    losses = []
    for _ in range(2):
        logits, feat = model(x)
        loss = torch.nn.functional.binary_cross_entropy(torch.sigmoid(logits), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert feat.shape == (2, 512) and all(np.isfinite(losses))
