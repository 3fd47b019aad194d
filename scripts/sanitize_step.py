"""A tiny but complete pass over the hot-path kernels for compute-sanitizer (memcheck / racecheck / initcheck):

    PYTORCH_NO_CUDA_MEMORY_CACHING=1 compute-sanitizer --tool initcheck python scripts/sanitize_step.py [encoder] [size] [n]

one search step (uint8 bank + scale/crop, tcgen05 convolutions forward / dgrad / wgrad, batch-norm, depthwise, loss, Adam,
fused Sinkhorn rewards) on a small configuration, then a second step (the CUDA-graph path is excluded: the sanitizer
instruments eager launches)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class Cfg:
    class CONTROLLER:
        EXCLUDE_OPS = []
        L = 2
        NUM_MAGS = 10
        EXCLUDE_OPS_NUM = 0
    SEED = 0


def main():
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine
    from aadg_b200.nn import DeepLabV3Plus, Unet
    from aadg_b200.synth import fundus_batch, random_policies
    enc = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    arch = sys.argv[4] if len(sys.argv) > 4 else "deeplabv3plus"
    ctor = DeepLabV3Plus if arch == "deeplabv3plus" else Unet
    model = ctor(encoder_name=enc, encoder_weights=None, in_channels=3, classes=2, aux_params=dict(pooling="avg"))
    eng = SearchEngine(model, n_domains=3, M=6, crop=size, seed=3)
    eng.set_policies(parse_policies(random_policies(seed=3), Cfg), epoch=0)
    imgs, masks = fundus_batch(n, size, size, seed=8)
    x, m = torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda()
    for _ in range(2):
        out = eng.step(x, m, [i % 3 for i in range(n)])
    torch.cuda.synchronize()
    print("SANITIZE_STEP done: %s/%s %d^2 n=%d loss %.5f rewards %s" % (arch, enc, size, n, float(out["seg_loss"]),
                                                                       np.round(eng.rewards.cpu().numpy(), 4).tolist()))


if __name__ == "__main__":
    main()
