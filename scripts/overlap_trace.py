"""Evidence that the bucketed gradient all-reduce runs CONCURRENTLY with the backward pass (SURVEY 2a replacement map:
per-bucket ncclAllReduce issued from the backward pass; reference models/__init__.py:39 DDP bucketing).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/overlap_trace.py

nsys is not available here, so the device timeline comes from torch.profiler (CUPTI kernel records: name, stream,
start, duration) over two eager search steps on every rank.  Rank 0 prints, for each NCCL kernel of the second step,
its interval and the engine kernels (other streams) whose intervals intersect it, plus the totals: NCCL time, NCCL time
hidden under engine kernels, and what is left exposed."""
import os
import sys

import torch
import torch.distributed as dist

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
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine, shutdown
    from aadg_b200.nn import DeepLabV3Plus
    from aadg_b200.synth import fundus_batch, random_policies
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    s = int(sys.argv[2]) if len(sys.argv) > 2 else 12        # source images per rank (x6 augmented copies)
    imgs, masks = fundus_batch(s, size, size, seed=1023 + rank)
    x, m = torch.from_numpy(imgs).to(dev), torch.from_numpy(masks).to(dev)
    model = DeepLabV3Plus(encoder_name="resnet50", classes=2, device=dev, seed=1023)
    eng = SearchEngine(model, n_domains=3, M=6, crop=size, seed=1023, graph=False)
    eng.set_policies(parse_policies(random_policies(seed=1023), Cfg), epoch=0)
    domains = [(rank * s + i) % 3 for i in range(s)]
    for _ in range(2):
        eng.step(x, m, domains)
    torch.cuda.synchronize()
    dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        eng.step(x, m, domains)
        torch.cuda.synchronize()
    if rank == 0:
        ks = []
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None:
                name = e.name
                if "memcpy" in name.lower() or "memset" in name.lower() or name.startswith("nccl:"):
                    continue          # copies, and the profiler's own "nccl:all_reduce" annotation ranges (not kernels)
                ks.append((e.time_range.start, e.time_range.end, name))
        ks.sort()
        t0 = ks[0][0]
        nccl = [k for k in ks if "nccl" in k[2].lower()]
        eng_k = [k for k in ks if "nccl" not in k[2].lower()]
        print("OVERLAP_TRACE world %d, %d images of %dx%d per rank, eager step: %d device kernels, %d of them NCCL" %
              (world, 6 * s, size, size, len(ks), len(nccl)))
        total = hidden = 0.0
        for (a, b, name) in nccl:
            dur = b - a
            cover = []
            for (c, d, other) in eng_k:
                if d <= a:
                    continue
                if c >= b:
                    break
                cover.append((max(a, c), min(b, d), other))
            # union of the covering intervals
            cov = 0.0
            end = a
            for (c, d, _o) in sorted(cover):
                c = max(c, end)
                if d > c:
                    cov += d - c
                    end = d
            total += dur
            hidden += cov
            if dur >= 20:      # us: the gradient buckets and the feature all-gather, not the tiny scalar collectives
                names = {}
                for (_c, _d, o) in cover:
                    short = o.split("(")[0].replace("void ", "").replace("aadg::", "")[:40]
                    names[short] = names.get(short, 0) + 1
                top = sorted(names.items(), key=lambda kv: -kv[1])[:4]
                print("  %-44s start %9.1f us  dur %8.1f us  covered by engine kernels %5.1f %%  concurrent: %s" %
                      (name[:44], a - t0, dur, 100.0 * cov / max(dur, 1e-9), top))
        print("OVERLAP_TRACE totals: NCCL kernel time %.1f us per step, hidden under engine kernels %.1f us (%.1f %%), exposed "
              "%.1f us; step span %.1f us" % (total, hidden, 100.0 * hidden / max(total, 1e-9), total - hidden,
                                              ks[-1][1] - t0))
    sys.stdout.flush()
    shutdown([eng])


if __name__ == "__main__":
    main()
