"""Time the HBM-bound layer kernels (csrc/nn_elem.cu, csrc/nn_loss.cu) at the shapes of the 144-image
DeepLabV3+/ResNet-50 step and report algorithmic GB/s against the measured HBM peak.  One JSON line per
kernel x shape.  `python scripts/nnbench.py [bn] [dw] [misc]`"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from aadg_b200.ops import nn as K  # noqa: E402

BF16 = torch.bfloat16
N = int(os.environ.get("NNBENCH_N", "144"))


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def report(name, shape, ms, nbytes):
    pk = peak()
    print(json.dumps({"kernel": name, "shape": list(shape), "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1),
                      "gbs": round(nbytes / ms / 1e6, 0), "frac_of_measured_hbm": round(nbytes / ms / 1e6 / pk, 3)}),
          flush=True)


def bench_bn():
    for (h, c) in ((256, 64), (128, 64), (128, 256), (64, 512), (32, 1024), (32, 2048), (128, 48)):
        shape = (N, h, h, c)
        x = torch.randn(shape, device="cuda").to(BF16)
        dy = torch.randn(shape, device="cuda").to(BF16)
        y = torch.empty_like(x)
        dx = torch.empty_like(x)
        res = torch.randn(shape, device="cuda").to(BF16)
        dres = torch.empty_like(x)
        bits = torch.empty(x.numel() // 8, dtype=torch.uint8, device="cuda")
        s = torch.zeros(6, c, device="cuda")
        gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
        dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
        nb = x.numel() * 2
        K.bn_stats(x, s[0], s[1])
        K.bn_finalize(s[0], s[1], gamma, beta, x.numel() // c, 1e-5, 0.1, s[2], s[3], s[4], s[5], rm, rv)
        report("bn_stats", shape, timeit(lambda: K.bn_stats(x, s[0], s[1])), nb)
        report("bn_apply relu", shape, timeit(lambda: K.bn_apply(x, s[4], s[5], y)), 2 * nb)
        report("bn_apply res+relu+bits", shape,
               timeit(lambda: K.bn_apply(x, s[4], s[5], y, res=res, relu_bits=bits)), 3 * nb + nb // 16)
        report("bn_backward recompute-mask (reduce+apply)", shape,
               timeit(lambda: K.bn_backward(dy, x, None, s[2], s[3], gamma, dg, db, dx, shift=s[5])), 5 * nb)
        report("bn_backward bits+dres (reduce+apply)", shape,
               timeit(lambda: K.bn_backward(dy, x, bits, s[2], s[3], gamma, dg, db, dx, dres=dres)),
               6 * nb + nb // 8)
        del x, dy, y, dx, res, dres, bits
        torch.cuda.empty_cache()


def bench_dw():
    for (h, c, dil) in ((128, 304, 1), (128, 256, 1), (32, 256, 1), (32, 2048, 12), (32, 2048, 36)):
        shape = (N, h, h, c)
        x = torch.randn(shape, device="cuda").to(BF16)
        dy = torch.randn(shape, device="cuda").to(BF16)
        y = torch.empty_like(x)
        w = torch.randn(9, c, device="cuda")
        dw = torch.zeros(9, c, device="cuda")
        nb = x.numel() * 2
        report("dwconv3x3 fwd dil%d" % dil, shape, timeit(lambda: K.dwconv3x3(x, w, dil, y)), 2 * nb)
        report("dwconv3x3 dgrad dil%d" % dil, shape, timeit(lambda: K.dwconv3x3(dy, w, dil, y, backward_data=True)), 2 * nb)
        report("dwconv3x3 wgrad dil%d" % dil, shape, timeit(lambda: K.dwconv3x3_wgrad(x, dy, dil, dw)), 2 * nb)
        del x, dy, y
        torch.cuda.empty_cache()


def bench_misc():
    size = 512
    img = torch.randn(N, 3, size, size, device="cuda")
    col = K.im2col_stem(img, 7, 7, 2, 3, 168, row_pitch=24)
    report("im2col_stem 7x7 s2 -> 168 (row pitch 24)", img.shape,
           timeit(lambda: K.im2col_stem(img, 7, 7, 2, 3, 168, row_pitch=24)), img.numel() * 4 + col.numel() * 2)
    del col
    f1 = torch.randn(N, 256, 256, 64, device="cuda").to(BF16)
    pooled, arg = K.maxpool_fwd(f1)
    report("maxpool fwd", f1.shape, timeit(lambda: K.maxpool_fwd(f1)), f1.numel() * 2 + pooled.numel() * 3)
    d = torch.randn_like(pooled)
    report("maxpool bwd", f1.shape, timeit(lambda: K.maxpool_bwd(d, arg, f1.shape)), f1.numel() * 2 + pooled.numel() * 3)
    del f1, pooled, arg, d
    a = torch.randn(N, 32, 32, 256, device="cuda").to(BF16)
    up = torch.empty(N, 128, 128, 304, device="cuda", dtype=BF16)
    report("upsample x4 fwd", a.shape, timeit(lambda: K.upsample_fwd(a, up[..., :256])), a.numel() * 2 + N * 128 * 128 * 512)
    da = torch.empty_like(a)
    report("upsample x4 bwd", a.shape, timeit(lambda: K.upsample_bwd(up[..., :256], da)), a.numel() * 2 + N * 128 * 128 * 512)
    del a, up, da
    dec = torch.randn(N, 128, 128, 256, device="cuda").to(BF16)
    hw, hb = torch.randn(2, 256, device="cuda") * 0.05, torch.zeros(2, device="cuda")
    z = K.seg_head_fwd(dec, hw, hb)
    report("seg_head fwd", dec.shape, timeit(lambda: K.seg_head_fwd(dec, hw, hb)), dec.numel() * 2 + z.numel() * 4)
    target = (torch.rand(N, 2, size, size, device="cuda") > 0.5).float()
    loss = torch.zeros(1, dtype=torch.float64, device="cuda")
    counts = torch.zeros(N, 2, 3, dtype=torch.int32, device="cuda")
    report("seg_loss fwd", target.shape, timeit(lambda: K.seg_loss_fwd(z, target, 0.5, loss, counts)),
           target.numel() * 4 + z.numel() * 4)
    report("seg_loss bwd", target.shape, timeit(lambda: K.seg_loss_bwd(z, target, 1e-6)), target.numel() * 4 + 2 * z.numel() * 4)
    dz = K.seg_loss_bwd(z, target, 1e-6)
    ddec = torch.empty_like(dec)
    dw_, db_ = torch.zeros_like(hw), torch.zeros_like(hb)
    report("seg_head bwd", dec.shape, timeit(lambda: K.seg_head_bwd(dz, dec, hw, ddec, dw_, db_)),
           dec.numel() * 4 + dz.numel() * 4)
    a2, b2 = torch.randn(N, 128, 128, 256, device="cuda").to(BF16), torch.randn(N, 128, 128, 256, device="cuda").to(BF16)
    report("add_", a2.shape, timeit(lambda: K.add_(a2, b2)), a2.numel() * 6)


if __name__ == "__main__":
    args = sys.argv[1:]
    if not args or "bn" in args:
        bench_bn()
    if not args or "dw" in args:
        bench_dw()
    if not args or "misc" in args:
        bench_misc()
