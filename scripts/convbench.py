"""Per-launch table of the tensor-core convolutions of one 144-image search step (CUDA events around every
launch): time, TFLOP/s, algorithmic bytes, the roofline bound max(flops/peak, bytes/hbm) and the time lost
against it.  `python scripts/convbench.py [--backbone resnet50] [--size 512] [--items 8]`"""
import argparse
import collections
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aadg_b200.nn import DeepLabV3Plus, Unet  # noqa: E402
from aadg_b200.ops import conv as C  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--backbone", default="resnet50")
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--n", type=int, default=144)
ap.add_argument("--top", type=int, default=40)
ap.add_argument("--arch", default="deeplabv3plus", choices=["deeplabv3plus", "unet"])
ap.add_argument("--classes", type=int, default=2)
a = ap.parse_args()
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
TF, BW = pk["bf16_tflops_sustained"] * 1e12, pk["hbm_gbs"] * 1e9

model = (DeepLabV3Plus if a.arch == "deeplabv3plus" else Unet)(
    encoder_name=a.backbone, encoder_weights=None, in_channels=3, classes=a.classes, aux_params=dict(pooling="avg"), seed=1)
x = torch.randn(a.n, 3, a.size, a.size, device="cuda")
t = (torch.rand(a.n, a.classes, a.size, a.size, device="cuda") > 0.5).float()
for _ in range(2):
    model.store.zero_grad()
    model.loss_step(x, t)
    model.store.adam_step(1e-3)
torch.cuda.synchronize()
C.TIMING = []
model.store.zero_grad()
model.loss_step(x, t)
torch.cuda.synchronize()
rec = C.TIMING
C.TIMING = None
agg = collections.OrderedDict()
for kind, flops, e0, e1, g in rec:
    n, h, w, cin, ho, wo, cout, r, stride, dil = g
    ms = e0.elapsed_time(e1)
    nbytes = 2.0 * n * h * w * cin + 2.0 * n * ho * wo * cout + (2 if kind != "wgrad" else 4) * r * r * cin * cout
    key = (kind,) + g
    d = agg.setdefault(key, [0, 0.0, flops, nbytes])
    d[0] += 1
    d[1] += ms
rows = []
for key, (cnt, ms, flops, nbytes) in agg.items():
    bound = max(flops / TF, nbytes / BW) * 1e3 * cnt
    rows.append((ms - bound, key, cnt, ms, flops * cnt / ms / 1e9, nbytes * cnt / ms / 1e6, bound))
tot = sum(r[3] for r in rows)
totb = sum(r[6] for r in rows)
print("conv launches %d  total %.2f ms  roofline bound %.2f ms" % (len(rec), tot, totb))
for kind in ("fprop", "dgrad", "wgrad"):
    print("  %s %.2f ms (bound %.2f)" % (kind, sum(r[3] for r in rows if r[1][0] == kind),
                                         sum(r[6] for r in rows if r[1][0] == kind)))
print("%-6s %-42s %3s %8s %8s %8s %8s %8s" % ("kind", "n,h,w,cin -> ho,wo,cout k/s/d", "x", "ms", "TFLOP/s", "GB/s", "bound", "lost"))
for lost, key, cnt, ms, tf, gbs, bound in sorted(rows, key=lambda r: -r[0])[:a.top]:
    kind, n, h, w, cin, ho, wo, cout, r, stride, dil = key
    print("%-6s %-42s %3d %8.3f %8.0f %8.0f %8.3f %8.3f" % (
        kind, "%d,%d,%d,%d -> %d,%d,%d %d/%d/%d" % (n, h, w, cin, ho, wo, cout, r, stride, dil), cnt, ms, tf, gbs, bound, lost))
