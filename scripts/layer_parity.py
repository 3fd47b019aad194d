"""Layer-by-layer implementation parity on REAL activations: every convolution / depthwise convolution / batch-norm of
the engine is fed the bf16-storage oracle's own input for that layer (oracle/segnet_bf16.py, exactly bf16-representable)
and compared with the oracle's output.  With identical rounding points the only legitimate residual is fp32 summation
order (a few 1e-5 relative L2 after bf16 rounding); anything near 1e-3 or above is a semantic difference.  Then the free
-running forward is compared stage by stage to show how the residual grows through the net.

    python scripts/layer_parity.py [encoder] [size] [n] [arch]      (GPU)
"""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BF = torch.bfloat16


def nhwc(x):
    return x.permute(0, 2, 3, 1).clone(memory_format=torch.contiguous_format)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def l2(got, want):
    return ((got.float() - want.float()).norm() / (want.float().norm() + 1e-20)).item()


def main():
    from test_parity_gpu import make_data, make_models
    from aadg_b200.nn import network as NW
    from aadg_b200.ops import conv as C
    from aadg_b200.ops import nn as K
    enc = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    arch = sys.argv[4] if len(sys.argv) > 4 else "deeplabv3plus"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    classes, dataset = (2, "optic") if arch == "deeplabv3plus" else (1, "vessel")
    x, target = make_data(n, size, classes, dataset=dataset)
    ref, twin, net = make_models(arch, enc, classes)
    mods = dict(twin.named_modules())
    io = {}
    for name, m in mods.items():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.BatchNorm2d)):
            m.register_forward_hook(lambda mod, inp, out, name=name: io.__setitem__(name, (inp[0].detach(), out.detach())))
    with torch.no_grad():
        logits_t, pooled_t = twin(x)

    rows = []
    seen = set()

    def walk(o):
        if id(o) in seen:
            return
        seen.add(id(o))
        if isinstance(o, NW.ConvBN):
            cname = o.w.name[:-len(".weight")]
            if cname in io:
                xin, want = io[cname]
                got = C.fprop(nhwc(xin).to(BF), o.w.bf16, o.k, o.k, o.stride, o.pad, o.dil)
                # the same convolution in float64 (bf16-rounded weights), rounded to bf16: tells the oracle's own fp32
                # algorithm error (cuDNN may pick Winograd / FFT) from the engine's
                m = mods[cname]
                w64 = m.weight.detach().to(BF).double()
                exact = torch.nn.functional.conv2d(xin.double(), w64, None, m.stride, m.padding, m.dilation).to(BF).float()
                rows.append((l2(nchw(got), want), "conv", cname, tuple(xin.shape), "k%d s%d d%d | engine vs fp64: %.2e, "
                             "oracle vs fp64: %.2e" % (o.k, o.stride, o.dil, l2(nchw(got), exact), l2(want, exact))))
                walk_bn(o.bn, want, relu=o.relu, relu6=o.relu6)
        elif isinstance(o, NW.Depthwise3x3):
            cname = o.w.name[:-len(".weight")]
            if cname in io:
                xin, want = io[cname]
                xi = nhwc(xin).to(BF)
                ho, wo = want.shape[2:]
                got = torch.empty((xi.shape[0], ho, wo, xi.shape[3]), dtype=BF, device=xi.device)
                K.dwconv3x3(xi, o.w.data, o.dil, got, stride=o.stride)
                rows.append((l2(nchw(got), want), "dwconv", cname, tuple(xin.shape), "s%d d%d" % (o.stride, o.dil)))
        if isinstance(o, (list, tuple)):
            for i in o:
                walk(i)
        elif hasattr(o, "__dict__") and not isinstance(o, (NW.ParamStore, NW.Param, torch.Tensor)):
            for v in vars(o).values():
                walk(v)

    def walk_bn(bn, conv_out, relu, relu6):
        if bn.name not in io:
            return
        xin, want = io[bn.name]                       # want: BN output (rounded unless it feeds a residual add)
        xi = nhwc(conv_out).to(BF)
        c = xi.shape[-1]
        buf = torch.zeros(6, c, device=xi.device)
        K.bn_stats(xi, buf[0], buf[1])
        K.bn_finalize(buf[0], buf[1], bn.gamma.data, bn.beta.data, xi.numel() // c, NW.BN_EPS, NW.BN_MOMENTUM, buf[2], buf[3],
                      buf[4], buf[5], None, None)
        y = torch.empty_like(xi)
        K.bn_apply(xi, buf[4], buf[5], y, relu=False)
        rounded = bool(torch.equal(want, want.to(BF).float()))
        # float64 statistics of the same tensor: how far are the engine's (and torch's) mean / invstd from exact?
        x64 = conv_out.double()
        mean64 = x64.mean((0, 2, 3))
        inv64 = (x64.var((0, 2, 3), unbiased=False) + NW.BN_EPS).rsqrt()
        e_mean = ((buf[2].double() - mean64).abs() * inv64).max().item()          # in units of sigma
        e_inv = ((buf[3].double() - inv64).abs() / inv64).max().item()
        rows.append((l2(nchw(y), want), "bn" if rounded else "bn(unrounded oracle: expect ~1e-3)", bn.name, tuple(xin.shape),
                     "| engine stats vs fp64: mean %.1e sigma, invstd rel %.1e" % (e_mean, e_inv)))
        # fused statistics of the convolution epilogue vs the separate pass, on the same tensor
    walk(net.encoder)
    walk(net.decoder)
    # stems (packed as 1x1 GEMMs over im2col patches)
    if hasattr(net.encoder, "stem_w"):
        cname = net.encoder.stem_w.name[:-len(".weight")]
        xin, want = io[cname]
        m = mods[cname]
        exact = torch.nn.functional.conv2d(xin.double(), m.weight.detach().to(BF).double(), None, m.stride, m.padding,
                                           m.dilation).to(BF).float()
        if enc == "mobilenet_v2":
            col = K.im2col_stem(xin.contiguous(), 3, 3, 2, 1, 3 * NW.MBV2_STEM_RP, row_pitch=NW.MBV2_STEM_RP)
        else:
            col = K.im2col_stem(xin.contiguous(), 7, 7, 2, 3, NW.STEM_KP, row_pitch=NW.STEM_RP)
        got = C.fprop(col, net.encoder.stem_w.bf16, 1, 1)
        rows.append((l2(nchw(got), want), "stem conv", cname, tuple(xin.shape), "| engine vs fp64: %.2e, oracle vs fp64: %.2e" %
                     (l2(nchw(got), exact), l2(want, exact))))
    rows.sort(reverse=True)
    print("== teacher-forced per-layer residuals (%s/%s %d^2 n=%d), worst first ==" % (arch, enc, size, n))
    for r in rows[:40]:
        print("  %.3e  %-10s %-45s %s %s" % r)
    print("  ... %d layers, median %.3e" % (len(rows), sorted(r[0] for r in rows)[len(rows) // 2]))

    # free-running forward, stage by stage
    feats_t = None
    with torch.no_grad():
        feats_t = twin.encoder(x.to(BF).float())
        feats_r = ref.encoder(x)
    net.train()
    feats_e = net.encoder.forward(x.contiguous(), True)
    print("== free-running encoder features: engine vs bf16-storage oracle | bf16-storage oracle vs fp32 oracle ==")
    for i, f in enumerate(feats_e):
        print("  feat[%d] %-22s %.3e | %.3e" % (i + 1, tuple(f.shape), l2(nchw(f), feats_t[i + 1]), l2(feats_t[i + 1], feats_r[i + 1])))

    # ---- determinism / uninitialised-memory probe -------------------------------------------------------------------
    # (a) the same forward twice: bitwise equal?  (fp32 atomics in the statistics may flip a last bit; anything larger is
    # a race)  (b) the allocator's free blocks poisoned with NaN before a run: a NaN anywhere downstream means a kernel
    # reads memory no kernel wrote
    def run_once():
        net.store.zero_grad()
        out = net.loss_step(x, target, want_logits=True)
        return out["loss"].item(), out["pooled"].clone(), out["logits"].clone(), net.store.grads.clone()
    net.dropout_enabled = False
    a = run_once()
    b = run_once()
    print("== determinism: two identical loss_step calls ==")
    print("  loss %.9f vs %.9f (rel %.2e)  pooled L2 %.2e  logits L2 %.2e  grads L2 %.2e" %
          (a[0], b[0], abs(a[0] - b[0]) / abs(a[0]), l2(b[1], a[1]), l2(b[2], a[2]), l2(b[3], a[3])))
    del b
    torch.cuda.empty_cache()
    free = torch.cuda.mem_get_info()[0]
    poison = torch.full((int(min(free * 0.6, 24 << 30)) // 2,), float("nan"), dtype=BF, device="cuda")
    del poison                      # the cached block (all NaN) is what the next torch.empty() calls carve up
    c = run_once()
    print("== NaN-poisoned allocator: loss %.9f (rel to clean %.2e) finite: loss %s pooled %s logits %s grads %s ==" %
          (c[0], abs(c[0] - a[0]) / abs(a[0]), c[0] == c[0], bool(torch.isfinite(c[1]).all()),
           bool(torch.isfinite(c[2]).all()), bool(torch.isfinite(c[3]).all())))


if __name__ == "__main__":
    main()
