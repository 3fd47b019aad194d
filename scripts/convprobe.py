"""Run a handful of representative convolution launches (shapes of the ResNet-50 DeepLabV3+ step) once each after a
warm-up, for `ncu --set full` captures:  ncu ... -k regex:'igemm|wgrad' python scripts/convprobe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aadg_b200.ops import conv as C  # noqa: E402

BF = torch.bfloat16
N = 144
CASES = [  # kind, h, cin, cout, k, stride, dil
    ("fprop", 128, 64, 256, 1, 1, 1),
    ("dgrad", 128, 64, 256, 1, 1, 1),       # dy has 256 channels -> dx 64
    ("fprop", 128, 64, 64, 3, 1, 1),
    ("wgrad", 128, 64, 64, 3, 1, 1),
    ("fprop", 32, 512, 2048, 1, 1, 1),
    ("wgrad", 32, 512, 2048, 1, 1, 1),
    ("fprop", 32, 512, 512, 3, 1, 2),
    ("fprop", 256, 168, 64, 1, 1, 1),
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for kind, h, cin, cout, k, stride, dil in CASES:
    pad = dil * (k // 2)
    x = torch.randn(N, h, h, cin, device="cuda").to(BF)
    w = (torch.randn(k * k, cout, cin, device="cuda") * 0.05).to(BF)
    ho = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
    dy = torch.randn(N, ho, ho, cout, device="cuda").to(BF)
    wt = w.permute(0, 2, 1).contiguous()
    dw = torch.zeros(k * k, cout, cin, device="cuda")
    for _ in range(1 + reps):
        if kind == "fprop":
            C.fprop(x, w, k, k, stride, pad, dil)
            ssum, ssq = torch.zeros(cout, device="cuda"), torch.zeros(cout, device="cuda")
            C.fprop(x, w, k, k, stride, pad, dil, stats=(ssum, ssq))      # fused batch-norm statistics variant
        elif kind == "dgrad":
            C.dgrad(dy, wt, k, k, stride, pad, dil, (h, h))
        else:
            C.wgrad(x, dy, k, k, stride, pad, dil, out=dw)
    torch.cuda.synchronize()
    del x, dy
    torch.cuda.empty_cache()
print("done")
