"""BASELINE.json config 5: microbenchmarks of the two HBM-bound kernel families against the measured HBM
peak (MEASURED_PEAKS.json): the uint8 augment bank at batch 8..512 @512x512 and the streamed Sinkhorn at
N = 1k..64k, d = 256.  One JSON line per point.  `python scripts/microbench.py [aug] [sinkhorn] [--max-n 65536]`"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aadg_b200.data import decisions as D  # noqa: E402
from aadg_b200.data.policy import parse_policies  # noqa: E402
from aadg_b200.ops import sinkhorn as SK  # noqa: E402
from aadg_b200.ops import u8 as U8  # noqa: E402
from aadg_b200.synth import fundus_batch, random_policies, feature_cloud  # noqa: E402


class Cfg:
    class CONTROLLER:
        EXCLUDE_OPS = []
        L = 2
        NUM_MAGS = 10
        EXCLUDE_OPS_NUM = 0
    SEED = 0


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def timeit(fn, iters, flush=None):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                      # 256 MB write: evicts the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def bench_aug(batches=(8, 16, 32, 64, 128, 256, 512)):
    pk = peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    parsed = parse_policies(random_policies(seed=1023), Cfg)
    h = w = 512
    for n_out in batches:
        s = max(1, n_out // 6)
        imgs, masks = fundus_batch(s, h, w, seed=7)
        rows, _ = D.philox_rows(parsed, s, w, h, w, (1, 1.5), seed=1, scale_crop=False)
        rows = rows[:n_out] if len(rows) >= n_out else np.concatenate([rows] * (n_out // len(rows) + 1))[:n_out]
        d_imgs, d_masks = torch.from_numpy(imgs).cuda(), torch.from_numpy(masks).cuda()
        out_i = torch.empty((n_out, 3, h, w), dtype=torch.float32, device="cuda")
        ms = timeit(lambda: U8.policy_normalize(d_imgs, d_masks, rows, want_labels=False, out_images=out_i), 10, flush)
        stat_ops = {0, 2, 5}
        n_stat_src = len({int(r["src"]) for r in rows if any(int(o) in stat_ops for o in r["op"][:int(r["n_ops"])])})
        alg = n_out * (3 + 12) * h * w + n_stat_src * 3 * h * w
        print(json.dumps({"bench": "aug_u8_bank", "batch": n_out, "size": 512, "ms": ms, "images_per_s": n_out / ms * 1e3,
                          "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6, "peak_gbs": pk,
                          "frac": alg / ms / 1e6 / pk, "l2": "flushed before every launch"}))


def bench_sinkhorn(max_n):
    pk = peak()
    for n in (1024, 2048, 4096, 8192, 16384, 32768, 65536):
        if n > max_n:
            break
        x = torch.from_numpy(feature_cloud(n, 256, 0)).cuda()
        y = torch.from_numpy(feature_cloud(n, 256, 1)).cuda()
        val, n_eps = SK.divergence_large(x, y)
        torch.cuda.synchronize()
        iters = 3 if n >= 16384 else 5
        ms = timeit(lambda: SK.divergence_large(x, y), iters)
        ms_setup = timeit(lambda: SK.large_setup_only(x, y), iters)
        sweeps = n_eps + 2
        alg = sweeps * 4.0 * n * n * 4
        ms_it = max(ms - ms_setup, 1e-6)
        print(json.dumps({"bench": "sinkhorn_large", "n": n, "d": 256, "ms_total": ms, "ms_cost_build": ms_setup,
                          "ms_iterations": ms_it, "n_eps": n_eps, "softmin_sweeps": sweeps,
                          "iters_per_s": sweeps / ms_it * 1e3, "iters_per_s_incl_cost_build": sweeps / ms * 1e3,
                          "algorithmic_bytes": alg, "achieved_gbs": alg / ms_it / 1e6, "peak_gbs": pk,
                          "frac": alg / ms_it / 1e6 / pk, "frac_incl_cost_build": alg / ms / 1e6 / pk,
                          "value": float(val),
                          "l2": "4 cost matrices = %.1f MB vs 126 MB L2" % (4 * n * n * 4 / 1e6)}))
    # the reference's native shape: 18 problems of 8 points, d = 128, one launch
    pts = torch.from_numpy(np.concatenate([feature_cloud(8, 128, k, seed=k) for k in range(18)])).cuda()
    probs = [[8 * (3 * j + a), 8, 8 * (3 * j + b), 8] for j in range(6) for a, b in ((0, 1), (1, 2), (0, 2))]
    ms = timeit(lambda: SK.divergence_batched(pts, probs), 20)
    print(json.dumps({"bench": "sinkhorn_native_18x(8x8,d128)", "launches": 1, "us_per_step": ms * 1e3}))


if __name__ == "__main__":
    args = sys.argv[1:]
    max_n = int(args[args.index("--max-n") + 1]) if "--max-n" in args else 65536
    if not args or "aug" in args:
        bench_aug(tuple(int(v) for v in args[args.index("--batches") + 1].split(",")) if "--batches" in args
                  else (8, 16, 32, 64, 128, 256, 512))
    if not args or "sinkhorn" in args:
        bench_sinkhorn(max_n)
