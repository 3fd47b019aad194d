"""Per-kernel device-time summary of one eager search step (torch.profiler / CUPTI kernel records; nsys is not installed
and an ncu launch list costs ~0.1 s per launch).  Same engine construction as bench.py.

    python scripts/kernel_summary.py [--arch deeplabv3plus] [--backbone mobilenet_v2] [--size 256] [--items 8] [--steps 2]

Prints, per kernel name: launches per step, device time per step, share; then the sum of kernel times and the span of
the step on the device (span - sum = idle gaps between kernels: launch latency the CUDA graph removes).
"""
import argparse
import os
import sys

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


def short(name):
    name = name.replace("void ", "").replace("aadg::", "")
    depth, out = 0, []
    for ch in name:                  # drop the argument list, keep template arguments
        if ch == "(" and depth == 0:
            break
        out.append(ch)
    return "".join(out)[:64]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="deeplabv3plus")
    ap.add_argument("--backbone", default="mobilenet_v2")
    ap.add_argument("--dataset", default="optic")
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--items", type=int, default=8)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--top", type=int, default=45)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine, shutdown
    from aadg_b200.nn import DeepLabV3Plus, Unet
    from aadg_b200.synth import fundus_batch, random_policies, vessel_batch
    d = 3 if a.dataset == "optic" else 4
    s = a.items * d
    make = fundus_batch if a.dataset == "optic" else vessel_batch
    imgs, masks = make(s, a.size, a.size, seed=1023)
    x, m = torch.from_numpy(imgs).to(dev), torch.from_numpy(masks).to(dev)
    ctor = DeepLabV3Plus if a.arch == "deeplabv3plus" else Unet
    model = ctor(encoder_name=a.backbone, encoder_weights=None, in_channels=3, classes=2 if a.dataset == "optic" else 1,
                 aux_params=dict(pooling="avg"), device=dev, seed=1023)
    eng = SearchEngine(model, n_domains=d, M=6, lr=1e-3, dataset=a.dataset, seed=1023, crop=a.size, scale_range=(1, 1.5),
                       graph=False)
    eng.set_policies(parse_policies(random_policies(m=6, seed=1023), Cfg), epoch=0)
    domains = [i % d for i in range(s)]
    for _ in range(3):
        eng.step(x, m, domains)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            eng.step(x, m, domains)
        torch.cuda.synchronize()
    rows, first, last = {}, None, None
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CUDA or e.time_range is None:
            continue
        t0, t1 = e.time_range.start, e.time_range.end
        first = t0 if first is None else min(first, t0)
        last = t1 if last is None else max(last, t1)
        k = short(e.name)
        c, t = rows.get(k, (0, 0.0))
        rows[k] = (c + 1, t + (t1 - t0))
    total = sum(t for _, t in rows.values())
    n = sum(c for c, _ in rows.values())
    print("KERNEL_SUMMARY %s/%s %s %d^2, %d images per step, %d eager steps: %.1f device records per step, sum of kernel "
          "times %.3f ms per step, device span %.3f ms per step" %
          (a.arch, a.backbone, a.dataset, a.size, 6 * s, a.steps, n / a.steps, total / a.steps / 1e3,
           (last - first) / a.steps / 1e3))
    for k, (c, t) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:a.top]:
        print("%6.2f %%  %9.3f ms  %7.1f  avg %8.1f us  %s" % (100.0 * t / total, t / a.steps / 1e3, c / a.steps, t / c, k))
    sys.stdout.flush()
    shutdown([eng])


if __name__ == "__main__":
    main()
