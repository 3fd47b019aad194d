"""cProfile of the host side of the search step (what the CPU spends enqueueing one step).  GPU box only."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aadg_b200.data.policy import parse_policies  # noqa: E402
from aadg_b200.host.search import SearchEngine  # noqa: E402
from aadg_b200.nn import DeepLabV3Plus  # noqa: E402
from aadg_b200.synth import fundus_batch, random_policies  # noqa: E402
from bench import Cfg  # noqa: E402

imgs, masks = fundus_batch(24, 512, 512, seed=1023)
d_imgs, d_masks = torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda()
model = DeepLabV3Plus(encoder_name="resnet50", encoder_weights=None, in_channels=3, classes=2,
                      aux_params=dict(pooling="avg"), seed=1023)
eng = SearchEngine(model, n_domains=3, M=6, lr=1e-3, dataset="optic", seed=1023, crop=512, scale_range=(1, 1.5))
eng.set_policies(parse_policies(random_policies(m=6, seed=1023), Cfg), epoch=0)
domains = [i % 3 for i in range(24)]
for _ in range(3):
    eng.step(d_imgs, d_masks, domains)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(4):
    eng.step(d_imgs, d_masks, domains)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
st.sort_stats("cumulative").print_stats(45)
