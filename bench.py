#!/usr/bin/env python
"""Headline benchmark: joint search+train images/sec of one AADG search step (BASELINE.json config 2).

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a engine)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm's CPU path

A step = augment (uint8 bank, M=6 policies x L=2 ops, Normalize_dg/ToTensor) -> DeepLabV3+/ResNet-50
forward -> momentum-discriminator features -> 18 Sinkhorn divergences -> BCE backward -> Adam (+ the
discriminator step) on B*D*M = 8*3*6 = 144 synthetic 512x512 fundus images per GPU (weak scaling:
every rank owns its own 24 source images; gradients all-reduced, features all-gathered over NCCL).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
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


def workload_config(a, n_gpus):
    d, m = n_source_domains(a), 6
    name = "config2: OD/OC 3-source-domain %dx%d fundus, DeepLabV3+/%s" if (a.dataset, a.arch) == ("optic", "deeplabv3plus") \
        else ("config4-style: " + a.dataset + " %d-source-domain" % d + " %dx%d, " + a.arch + "/%s")
    return {"workload": (name + ", Sinkhorn diversity reward, search step (augment->fwd->rewards->bwd->Adam)") %
                        (a.size, a.size, a.backbone),
            "items_per_gpu": a.items, "domains": d, "policies_M": m, "images_per_step_per_gpu": a.items * d * m,
            "global_images_per_step": a.items * d * m * n_gpus, "image_size": a.size, "sub_policy_ops_L": 2,
            "scale_crop": "none" if a.no_scale_crop else
            "DGRandomScaleCrop(%d, scale 1-1.5, p=0.8) on the device, bit-exact Pillow resize" % a.size,
            "parallelism": "dp%d (source images sharded; NCCL grad all-reduce + feature all-gather)" % n_gpus,
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
def cpu_joint_step_sample(a, threads=None):
    """A bounded sample of the same workload on the CPU: 1 source image -> M=6 augmented copies through the
    oracle's uint8 bank + normalise (numpy, 1 core), forward+backward+Adam of the torch DeepLabV3+ oracle on
    2 of them (all cores), 18 Sinkhorn divergences at the native shape (numpy fp32).  Returns (images/s, info)."""
    import torch
    from aadg_b200.data import decisions as D
    from aadg_b200.data.policy import parse_policies
    from aadg_b200.synth import fundus_batch, random_policies, feature_cloud
    from oracle import sinkhorn as OS
    from oracle import u8_policy as OP
    from oracle.segnet_torch import DeepLabV3PlusTorch
    if threads:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    size = a.size
    imgs, masks = fundus_batch(1, size, size, seed=1023)
    parsed = parse_policies(random_policies(seed=1023), Cfg)
    rows, _ = D.philox_rows(parsed, 1, size, size, size, (1, 1.5), seed=1023, scale_crop=not a.no_scale_crop)
    t0 = time.perf_counter()
    out = OP.apply_rows(imgs, masks, rows, crop=None if a.no_scale_crop else size, dataset="optic")
    t_aug = (time.perf_counter() - t0) / len(rows)
    torch.manual_seed(0)
    model = DeepLabV3PlusTorch(a.backbone, 2).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    nb = 2
    x = torch.from_numpy(out["images"][:nb])
    y = torch.from_numpy(out["labels"][:nb])
    t0 = time.perf_counter()
    logits, feat = model(x)
    loss = torch.nn.functional.binary_cross_entropy(torch.sigmoid(logits), y)
    opt.zero_grad()
    loss.backward()
    opt.step()
    t_model = (time.perf_counter() - t0) / nb
    clouds = [feature_cloud(8, 128, k, seed=k) for k in range(3)]
    t0 = time.perf_counter()
    for _ in range(6):
        for p, q in ((0, 1), (1, 2), (0, 2)):
            OS.sinkhorn_divergence(clouds[p], clouds[q], np.float32)
    t_sink = (time.perf_counter() - t0) / 144.0
    per_img = t_aug + t_model + t_sink
    info = {"cores": cores, "kind": "port",
            "sample": "1 source image -> 6 augmented %dx%d copies (oracle uint8 bank + scale/crop, numpy, 1 core): %.3f s/img; "
                      "DeepLabV3+/%s fwd+bwd+Adam on 2 images (torch CPU fp32, %d threads): %.3f s/img; 18 Sinkhorn "
                      "divergences N=8 d=128 (numpy fp32): %.4f s per 144-image step" %
                      (size, size, t_aug, a.backbone, cores, t_model, t_sink * 144)}
    return 1.0 / per_img, info


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    info = None
    for i in range(a.warmup + a.steps):
        v, info = cpu_joint_step_sample(a)
        if i >= a.warmup:
            vals.append(v)
        if i == 0 and a.warmup + a.steps > 2:
            pass
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1000.0 * 144 / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(a, 1),
            "cpu_baseline": dict(info, value=value, unit=UNIT),
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
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
    s = a.items * d
    make = fundus_batch if a.dataset == "optic" else vessel_batch
    imgs, masks = make(s, a.size, a.size, seed=1023 + rank)
    h_imgs = torch.from_numpy(imgs).pin_memory()
    h_masks = torch.from_numpy(masks).pin_memory()
    d_imgs, d_masks = h_imgs.to(dev), h_masks.to(dev)
    domains = [i % d for i in range(s)]                     # row order b*D + d

    ctor = DeepLabV3Plus if a.arch == "deeplabv3plus" else Unet
    model = ctor(encoder_name=a.backbone, encoder_weights=None, in_channels=3, classes=2 if a.dataset == "optic" else 1,
                 aux_params=dict(pooling="avg"), device=dev, seed=1023)
    eng = SearchEngine(model, n_domains=d, M=m, lr=1e-3, dataset=a.dataset, seed=1023,
                       crop=None if a.no_scale_crop else a.size, scale_range=(1, 1.5))
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
    C.TIMING = []                      # (kind, flops, start event, end event) per tensor-core conv launch
    calls0 = _lib.CALLS
    ms = timed(step_resident, a.steps)
    launches = (_lib.CALLS - calls0)
    host_ms = host_issue.get("ms")
    conv_records = C.TIMING
    C.TIMING = None
    clk = clocks.stop() if rank == 0 else None
    # conv family roofline (tensor pipe): algorithmic FLOPs / summed device time of those launches
    tflops_achieved = None
    conv_ms = 0.0
    if conv_records:
        fl = sum(r[1] for r in conv_records)
        conv_ms = sum(r[2].elapsed_time(r[3]) for r in conv_records)
        tflops_achieved = fl / (conv_ms * 1e-3) / 1e12
    for _ in range(1):
        step_e2e()
    ms_e2e = timed(step_e2e, a.steps)

    n_img = a.items * d * m
    value = n_img * world * a.steps / (ms * 1e-3)
    e2e = n_img * world * a.steps / (ms_e2e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    traffic = None
    try:   # dram__bytes_read+write of the conv launches of ONE step, from the committed ncu capture (not live)
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_conv_traffic.json")))
        traffic = tj["dram_bytes_per_launch_avg"]
    except Exception:
        pass
    roof = {"bound": "tensor", "kernel": "aadg::tc::igemm_kernel / wgrad_kernel (all conv fprop+dgrad+wgrad launches)",
            "achieved": tflops_achieved, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": (tflops_achieved / peak_tf) if tflops_achieved else None, "traffic": traffic,
            "traffic_note": "avg DRAM bytes per conv launch from profiles/r01_conv_traffic.json (ncu), same command",
            "flops_per_launch_avg": (sum(r[1] for r in conv_records) / len(conv_records)) if conv_records else None,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
            "conv_ms_per_step": conv_ms / a.steps, "conv_share_of_step": conv_ms / ms if ms else None,
            "conv_launches_per_step": len(conv_records) / a.steps if conv_records else None}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": workload_config(a, world), "clocks": clk,
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e / a.steps,
                    "h2d_bytes_per_step": int(h_imgs.numel() + h_masks.numel()) + 160 * n_img + 4 * d * n_img,
                    "d2h_bytes_per_step": int(result_host.numel()) * 4},
            "gpu_launches": launches, "host_enqueue_ms_per_step": host_ms, "roofline": roof}
    if world == 1 and not a.no_cpu_baseline:
        try:
            v, info = cpu_joint_step_sample(a)
            line["cpu_baseline"] = dict(info, value=v, unit=UNIT)
        except Exception as e:       # the baseline is a reported number, never a reason to lose the GPU line
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": None, "kind": "port", "sample": "failed: %r" % e}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
