#!/usr/bin/env python
"""Headline benchmark: joint search+train images/sec of one AADG search step (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W             # our arm (sm_100a engine); torchrun for N > 1
    python bench.py --gpus N --scaling strong                  # BASELINE config 3: config 2's 24 sources sharded over N GPUs
    python bench.py --arch unet --backbone resnet34 --dataset vessel --size 1024 --items 1     # config 4 (per GPU)
    python bench.py --impl reference --steps K --warmup W      # the reference algorithm's CPU path (oracle port)

A step = augment (uint8 bank, M=6 policies x L=2 ops, DGRandomScaleCrop, Normalize_dg/ToTensor) -> DeepLabV3+/ResNet-50
forward -> momentum-discriminator features -> 18 Sinkhorn divergences -> BCE backward -> Adam (+ the discriminator
step) on B*D*M = 8*3*6 = 144 synthetic 512x512 fundus images per GPU (weak scaling: every rank owns its own 24 source
images; strong scaling: the 24 source images are sharded; gradients all-reduced in buckets from the backward pass,
features all-gathered over NCCL).  The model part of the step replays a captured CUDA graph (--graph 0: eager).
Prints ONE JSON line (rank 0): value / e2e / roofline (conv family vs the measured sustained bf16 peak, with the ncu
DRAM traffic when profiles/r02_conv_traffic.json matches the launch list) / extra (BASELINE's second metric: Sinkhorn
iterations/s at N = 65536 and the uint8 bank at batch 512 vs the measured HBM peak) / cpu_baseline.  DESIGN.md
"Measurement" explains every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "joint search+train images/sec"
UNIT = "images/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--backbone", default="resnet50")
    ap.add_argument("--items", type=int, default=8, help="TRAIN.BATCH_SIZE: items per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-scale-crop", action="store_true", help="skip DGRandomScaleCrop (fused policy+normalise pass)")
    ap.add_argument("--arch", default="deeplabv3plus", choices=["deeplabv3plus", "unet"])
    ap.add_argument("--dataset", default="optic", choices=["optic", "vessel"],
                    help="optic: 3 source domains, 2 classes (config 2); vessel: 4 domains, 1 class (config 4)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU owns --items items (144 images at the defaults); strong: the --items items "
                         "(BASELINE config 3: config 2's 24 source images) are sharded over the GPUs")
    ap.add_argument("--graph", type=int, default=1, help="1: the model part of the step replays a captured CUDA graph")
    ap.add_argument("--extras", type=int, default=1,
                    help="1: add BASELINE's second metric to the line (`extra`): Sinkhorn iterations/s at N=65536 d=256 and "
                         "the uint8 bank at batch 512 @512x512, both against the measured HBM peak")
    ap.add_argument("--ref-sources", type=int, default=2,
                    help="--impl reference: source images per timed step (x6 augmented copies each)")
    return ap.parse_args()


class Cfg:
    class CONTROLLER:
        EXCLUDE_OPS = []
        L = 2
        NUM_MAGS = 10
        EXCLUDE_OPS_NUM = 0
    SEED = 0


def n_source_domains(a):
    return 3 if a.dataset == "optic" else 4


def sources_per_gpu(a, n_gpus):
    total = a.items * n_source_domains(a)
    if a.scaling == "strong":
        if total % n_gpus:
            raise SystemExit("--scaling strong: %d source images do not split over %d GPUs" % (total, n_gpus))
        return total // n_gpus
    return total


def workload_config(a, n_gpus):
    d, m = n_source_domains(a), 6
    name = "config2: OD/OC 3-source-domain %dx%d fundus, DeepLabV3+/%s" if (a.dataset, a.arch) == ("optic", "deeplabv3plus") \
        else ("config4-style: " + a.dataset + " %d-source-domain" % d + " %dx%d, " + a.arch + "/%s")
    return {"workload": (name + ", Sinkhorn diversity reward, search step (augment->fwd->rewards->bwd->Adam)") %
                        (a.size, a.size, a.backbone),
            "items_per_gpu": sources_per_gpu(a, n_gpus) / d, "domains": d, "policies_M": m,
            "images_per_step_per_gpu": sources_per_gpu(a, n_gpus) * m,
            "global_images_per_step": sources_per_gpu(a, n_gpus) * m * n_gpus, "image_size": a.size, "sub_policy_ops_L": 2,
            "scaling": a.scaling, "cuda_graph": bool(a.graph),
            "scale_crop": "none" if a.no_scale_crop else
            "DGRandomScaleCrop(%d, scale 1-1.5, p=0.8) on the device, bit-exact Pillow resize" % a.size,
            "parallelism": "dp%d (source images sharded; bucketed NCCL grad all-reduce overlapped with backward + one "
                           "feature all-gather)" % n_gpus,
            "l2": "per-step working set (tens of GB of activations) >> 126 MB L2, no explicit flush needed"}


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm on host cores (oracle port; /root/reference does not exist on the box)
# ---------------------------------------------------------------------------------------------------
def _worker_init():
    """augmentation workers are single-threaded numpy (the model's torch threads belong to the parent)"""
    try:
        import torch as _t
        _t.set_num_threads(1)
    except Exception:
        pass


def _aug_one_source(job):
    """worker: the oracle's DGMultiPolicy -> DGRandomScaleCrop -> Normalize_dg -> ToTensor for ONE source image
    (its M = 6 augmented copies), the unit a DataLoader worker of the reference processes (data/optic.py:79-91)"""
    import numpy as _np
    from oracle import u8_policy as OP
    img, mask, rows, crop, dataset = job
    rows = rows.copy()
    rows["src"] = 0
    out = OP.apply_rows(img[None], mask[None], rows, crop=crop, dataset=dataset)
    return _np.ascontiguousarray(out["images"]), _np.ascontiguousarray(out["labels"])


class CpuReference:
    """The reference algorithm's CPU path for the same workload, as REAL timed steps on a bounded sample:
    one step = `n_src` source images -> 6 augmented copies each through the oracle's uint8 bank + scale/crop +
    normalise (one process per source image, like the reference's DataLoader workers), then forward + BCE + backward
    of the torch DeepLabV3+/UNet oracle over ALL of those images in micro-batches of one source's 6 copies (gradients
    accumulated, one Adam step per step, every host thread), then the 18 Sinkhorn divergences of a step at the native
    shape.  images/s = images actually processed / wall time of the step; nothing is extrapolated."""

    def __init__(self, a, n_src):
        import multiprocessing as mp
        import torch
        from aadg_b200.data.policy import parse_policies
        from aadg_b200.synth import fundus_batch, random_policies, vessel_batch, feature_cloud
        from oracle.segnet_torch import DeepLabV3PlusTorch, UnetTorch
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)          # explicit: torchrun exports OMP_NUM_THREADS=1
        self.torch, self.a, self.n_src = torch, a, n_src
        d = n_source_domains(a)
        make = fundus_batch if a.dataset == "optic" else vessel_batch
        self.imgs, self.masks = make(n_src, a.size, a.size, seed=1023)
        self.parsed = parse_policies(random_policies(seed=1023), Cfg)
        torch.manual_seed(0)
        classes = 2 if a.dataset == "optic" else 1
        self.model = (DeepLabV3PlusTorch if a.arch == "deeplabv3plus" else UnetTorch)(a.backbone, classes).train()
        self.opt = torch.optim.Adam(self.model.parameters(), lr=1e-3)
        self.clouds = [feature_cloud(8, 128, k, seed=k) for k in range(d)]
        self.pool = mp.get_context("fork").Pool(min(self.cores, n_src), initializer=_worker_init) if n_src > 1 else None
        self.step_idx = 0
        self.split = {}

    def step(self):
        import torch.nn.functional as F
        from aadg_b200.data import decisions as D
        from oracle import sinkhorn as OS
        torch, a = self.torch, self.a
        t0 = time.perf_counter()
        rows, _ = D.philox_rows(self.parsed, self.n_src, a.size, a.size, a.size, (1, 1.5), seed=1023, step=self.step_idx,
                                scale_crop=not a.no_scale_crop)
        jobs = [(self.imgs[s], self.masks[s], rows[s * 6:(s + 1) * 6], None if a.no_scale_crop else a.size, a.dataset)
                for s in range(self.n_src)]
        outs = self.pool.map(_aug_one_source, jobs) if self.pool else [_aug_one_source(j) for j in jobs]
        t1 = time.perf_counter()
        self.opt.zero_grad()
        for im, lb in outs:                         # micro-batch = one source image's 6 copies
            logits, feat = self.model(torch.from_numpy(im))
            loss = F.binary_cross_entropy(torch.sigmoid(logits), torch.from_numpy(lb)) / len(outs)
            loss.backward()
        self.opt.step()
        t2 = time.perf_counter()
        # 18 divergences serve a 144-image step: this step's share, at least one full policy (3 pairs)
        n_pol = max(1, round(6 * self.n_src * 6 / 144.0))
        nd = len(self.clouds)
        for _ in range(n_pol):
            for p in range(nd):
                for q in range(p + 1, nd):
                    OS.sinkhorn_divergence(self.clouds[p], self.clouds[q], np.float32)
        t3 = time.perf_counter()
        self.step_idx += 1
        self.split = {"augment_s": t1 - t0, "model_s": t2 - t1, "sinkhorn_s": t3 - t2}
        return 6 * self.n_src, t3 - t0

    def describe(self):
        a = self.a
        return {"cores": self.cores, "kind": "port",
                "sample": "per step: %d source images -> %d augmented %dx%d copies (oracle uint8 bank + scale/crop, %d "
                          "worker processes), %s/%s forward+BCE+backward over all %d images in micro-batches of 6 + one "
                          "Adam step (torch CPU fp32, %d threads), %d policies' Sinkhorn divergences N=8 d=128 (numpy fp32); "
                          "measured wall time, last step: augment %.2f s, model %.2f s, sinkhorn %.3f s" %
                          (self.n_src, 6 * self.n_src, a.size, a.size, min(self.cores, self.n_src), a.arch, a.backbone,
                           6 * self.n_src, self.cores, max(1, round(6 * self.n_src * 6 / 144.0)),
                           self.split.get("augment_s", 0), self.split.get("model_s", 0), self.split.get("sinkhorn_s", 0))}

    def close(self):
        if self.pool:
            self.pool.terminate()


def cpu_baseline_leg(a, budget_s=25.0):
    """`cpu_baseline` of our arm's line: the same CPU path on a bounded sample (~10-30 s): one warm-up step and as many
    timed steps as fit the budget (at least one)."""
    ref = CpuReference(a, 1)
    try:
        ref.step()
        n_img, secs = 0, 0.0
        while True:
            n, t = ref.step()
            n_img, secs = n_img + n, secs + t
            if secs + t > budget_s:
                break
        info = ref.describe()
    finally:
        ref.close()
    return n_img / secs, info


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ref = CpuReference(a, a.ref_sources)
    try:
        for _ in range(a.warmup):
            ref.step()
        t0 = time.perf_counter()
        n_img = 0
        for _ in range(a.steps):
            n, _t = ref.step()
            n_img += n
        secs = time.perf_counter() - t0
        info = ref.describe()
    finally:
        ref.close()
    value = n_img / secs
    cfg = workload_config(a, a.gpus)        # the GPU arm's config, verbatim (the driver compares them)
    info["sample"] = ("a bounded sample of the workload: %d images per timed step (the GPU arm's step has %d per GPU); " % (
        6 * a.ref_sources, cfg["images_per_step_per_gpu"])) + info["sample"]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * secs / a.steps, "images_per_timed_step": 6 * a.ref_sources,
            "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": dict(info, value=value, unit=UNIT),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def measure_extras(dev):
    """BASELINE's second metric on this GPU (config 5's two headline points), measured here so that it is driver-run:
    Sinkhorn epsilon-iterations/s at N = M = 65536, d = 256 (streamed cost matrices, HBM-bound) and the uint8 bank at
    batch 512 @512x512 (policy + normalise, u8 -> f32), both as achieved algorithmic GB/s against the measured HBM peak."""
    import torch
    from aadg_b200.data import decisions as D
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.ops import sinkhorn as SK
    from aadg_b200.ops import u8 as U8
    from aadg_b200.synth import fundus_batch, random_policies, feature_cloud
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        pk_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pk, pk_src = 6650.0, "fallback 6 650 GB/s"

    def timeit(fn, iters, flush=None):
        fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(iters):
            if flush is not None:
                flush.zero_()                  # 256 MB write: evicts the 126 MB L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms))
    out = {"hbm_peak_gbs": pk, "peak_source": pk_src}
    with torch.cuda.device(dev):
        n, d = 65536, 256
        free = torch.cuda.mem_get_info(dev)[0]
        while n > 8192 and 4 * n * n * 4 * 1.1 > free:
            n //= 2
        x = torch.from_numpy(feature_cloud(n, d, 0)).to(dev)
        y = torch.from_numpy(feature_cloud(n, d, 1)).to(dev)
        val, n_eps = SK.divergence_large(x, y)
        ms = timeit(lambda: SK.divergence_large(x, y), 3)
        ms_setup = timeit(lambda: SK.large_setup_only(x, y), 3)
        sweeps = n_eps + 2
        alg = sweeps * 4.0 * n * n * 4
        ms_it = max(ms - ms_setup, 1e-6)
        out["sinkhorn"] = {"n": n, "d": d, "softmin_sweeps": sweeps, "iters_per_s": sweeps / ms_it * 1e3,
                           "iters_per_s_incl_cost_build": sweeps / ms * 1e3, "ms_cost_build": ms_setup,
                           "ms_iterations": ms_it, "achieved_gbs": alg / ms_it / 1e6, "frac": alg / ms_it / 1e6 / pk,
                           "algorithmic_bytes_per_sweep": 4.0 * n * n * 4, "value": float(val),
                           "l2": "4 cost matrices = %.1f GB >> 126 MB L2" % (4 * n * n * 4 / 1e9)}
        del x, y
        SK._lib._workspaces.clear()
        torch.cuda.empty_cache()
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        parsed = parse_policies(random_policies(seed=1023), Cfg)
        h = w = 512
        n_out = 512
        s_ = n_out // 6
        imgs, masks = fundus_batch(s_, h, w, seed=7)
        rows, _ = D.philox_rows(parsed, s_, w, h, w, (1, 1.5), seed=1, scale_crop=False)
        rows = np.concatenate([rows] * (n_out // len(rows) + 1))[:n_out]
        d_imgs, d_masks = torch.from_numpy(imgs).to(dev), torch.from_numpy(masks).to(dev)
        out_i = torch.empty((n_out, 3, h, w), dtype=torch.float32, device=dev)
        ms = timeit(lambda: U8.policy_normalize(d_imgs, d_masks, rows, want_labels=False, out_images=out_i), 10, flush)
        stat_ops = {0, 2, 5}
        n_stat = len({int(r["src"]) for r in rows if any(int(o) in stat_ops for o in r["op"][:int(r["n_ops"])])})
        alg = n_out * (3 + 12) * h * w + n_stat * 3 * h * w
        out["aug_u8_bank"] = {"batch": n_out, "size": 512, "ms": ms, "images_per_s": n_out / ms * 1e3,
                              "algorithmic_bytes": alg, "achieved_gbs": alg / ms / 1e6, "frac": alg / ms / 1e6 / pk,
                              "l2": "flushed before every launch"}
    return out


def run_ours(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # exactly ONE line on stdout: libraries that print there (NCCL's version banner) are sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    from aadg_b200 import _lib
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.host.search import SearchEngine
    from aadg_b200.nn import DeepLabV3Plus, Unet
    from aadg_b200.ops import conv as C
    from aadg_b200.synth import fundus_batch, random_policies, vessel_batch

    d, m = n_source_domains(a), 6
    s = sources_per_gpu(a, world)
    total_src = s * world
    make = fundus_batch if a.dataset == "optic" else vessel_batch
    if a.scaling == "strong":          # one global batch, this rank's contiguous block of it
        g_imgs, g_masks = make(total_src, a.size, a.size, seed=1023)
        imgs, masks = g_imgs[rank * s:(rank + 1) * s], g_masks[rank * s:(rank + 1) * s]
    else:                              # every rank owns its own batch
        imgs, masks = make(s, a.size, a.size, seed=1023 + rank)
    h_imgs = torch.from_numpy(np.ascontiguousarray(imgs)).pin_memory()
    h_masks = torch.from_numpy(np.ascontiguousarray(masks)).pin_memory()
    d_imgs, d_masks = h_imgs.to(dev), h_masks.to(dev)
    domains = [(rank * s + i) % d for i in range(s)]        # global row order b*D + d

    ctor = DeepLabV3Plus if a.arch == "deeplabv3plus" else Unet
    model = ctor(encoder_name=a.backbone, encoder_weights=None, in_channels=3, classes=2 if a.dataset == "optic" else 1,
                 aux_params=dict(pooling="avg"), device=dev, seed=1023)
    eng = SearchEngine(model, n_domains=d, M=m, lr=1e-3, dataset=a.dataset, seed=1023,
                       crop=None if a.no_scale_crop else a.size, scale_range=(1, 1.5), graph=bool(a.graph))
    eng.set_policies(parse_policies(random_policies(m=m, seed=1023), Cfg), epoch=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_issue = {}

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        host_issue["ms"] = (time.perf_counter() - t0) * 1e3 / steps     # CPU time to enqueue one step (no sync)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def step_resident():
        return eng.step(d_imgs, d_masks, domains)

    result_host = torch.empty(2 + (2 if a.dataset == "optic" else 1), dtype=torch.float32).pin_memory()

    def step_e2e():
        xi = h_imgs.to(dev, non_blocking=True)
        xm = h_masks.to(dev, non_blocking=True)
        out = eng.step(xi, xm, domains)
        res = torch.cat([out["seg_loss"].reshape(1), out["dis_loss"].reshape(1), out["dice"].float()])
        result_host.copy_(res, non_blocking=False)

    for _ in range(a.warmup):
        step_resident()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    calls0 = _lib.CALLS
    ms = timed(step_resident, a.steps)
    launches = (_lib.CALLS - calls0)
    host_ms = host_issue.get("ms")
    for _ in range(1):
        step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    clk = clocks.stop() if rank == 0 else None
    # conv family roofline (tensor pipe): algorithmic FLOPs / summed device time of those launches, every launch bracketed
    # by CUDA events on the launching stream.  A graph replay cannot be bracketed per kernel, so this pass runs the same
    # step EAGERLY (same kernels, same shapes, same stream) right after the timed regions; its step time is reported too.
    roof_steps = max(1, min(a.steps, 3))
    eng.use_graph = False
    step_resident()
    C.TIMING = []                      # (kind, flops, start event, end event, geometry) per tensor-core conv launch
    ms_eager = timed(step_resident, roof_steps)
    conv_records = C.TIMING
    C.TIMING = None
    eng.use_graph = bool(a.graph)
    tflops_achieved = None
    conv_ms = 0.0
    if conv_records:
        fl = sum(r[1] for r in conv_records)
        conv_ms = sum(r[2].elapsed_time(r[3]) for r in conv_records)
        tflops_achieved = fl / (conv_ms * 1e-3) / 1e12

    n_img = s * m
    value = n_img * world * a.steps / (ms * 1e-3)
    e2e = n_img * world * a.steps / (ms_e2e * 1e-3)
    extra = None
    if a.extras:
        # free the step's memory first: N = 65536 needs 4 x 17.2 GB of cost matrices
        del eng, model
        _lib._workspaces.clear()
        import gc
        gc.collect()
        torch.cuda.empty_cache()
        try:
            extra = measure_extras(dev)
        except Exception as e:
            extra = {"failed": repr(e)}
        if world > 1:       # config 5 at N GPUs: independent replicas, aggregate = sum
            agg = torch.tensor([extra.get("sinkhorn", {}).get("iters_per_s", 0.0),
                                extra.get("aug_u8_bank", {}).get("images_per_s", 0.0)], device=dev, dtype=torch.float64)
            dist.all_reduce(agg)
            extra["aggregate_over_gpus"] = {"n_gpus": world, "sinkhorn_iters_per_s": agg[0].item(),
                                            "aug_images_per_s": agg[1].item(), "note": "independent replicas (no collective)"}
    from aadg_b200.host.search import shutdown
    engines = [] if a.extras else [eng]
    if rank != 0:
        shutdown(engines)
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    n_conv = len(conv_records) / roof_steps if conv_records else None
    traffic, traffic_note = None, "no ncu capture of this exact launch list is committed"
    try:   # dram__bytes_read+write per conv launch from this round's committed ncu capture -- used only if it matches
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_conv_traffic.json")))
        if tj.get("conv_launches_per_step") == n_conv and tj.get("workload") == workload_config(a, 1)["workload"]:
            traffic = tj["dram_bytes_per_launch_avg"]
            traffic_note = ("avg DRAM bytes (dram__bytes_read + write) per conv call from an ncu capture of the same command's "
                            "conv kernels (profiles/r02_conv_traffic.json; used because its launch list matches this run's)")
    except Exception:
        pass
    roof = {"bound": "tensor", "kernel": "aadg::tc::igemm_p_kernel / wgrad_kernel (all conv fprop+dgrad+wgrad launches)",
            "achieved": tflops_achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (tflops_achieved / peak_tf) if tflops_achieved else None, "traffic": traffic,
            "traffic_note": traffic_note,
            "flops_per_launch_avg": (sum(r[1] for r in conv_records) / len(conv_records)) if conv_records else None,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
            "conv_ms_per_step": conv_ms / roof_steps, "conv_share_of_step": conv_ms / ms_eager if ms_eager else None,
            "conv_launches_per_step": n_conv,
            "measured_in": "%d eager steps with per-launch CUDA events after the timed regions (%.2f ms/step eager)" %
                           (roof_steps, ms_eager / roof_steps)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(a, world), "clocks": clk,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": int(h_imgs.numel() + h_masks.numel()) + 160 * n_img + 4 * d * n_img,
                    "d2h_bytes_per_step": int(result_host.numel()) * 4},
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms, "ms_per_step_eager": ms_eager / roof_steps,
            "roofline": roof}
    if extra is not None:
        line["extra"] = extra
    if world == 1 and not a.no_cpu_baseline:
        try:
            v, info = cpu_baseline_leg(a)
            line["cpu_baseline"] = dict(info, value=v, unit=UNIT)
        except Exception as e:       # the baseline is a reported number, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: %r" % e}
    emit(line)
    shutdown(engines)


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
