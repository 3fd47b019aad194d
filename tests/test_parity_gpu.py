"""GPU: whole-network parity at the benchmarked shapes (search_dg.py:132,140-142,170-172).

Two oracles (both stock torch fp32 layers, identical weights):
  * `oracle/segnet_bf16.py` — the fp32 oracle with the engine's bf16 STORAGE points made explicit.  Same algorithm as
    the engine, so the ReLU masks coincide ("teacher-forced" by construction) and the bounds are tight: this is the
    implementation-parity test.
  * `oracle/segnet_torch.py` — the plain fp32 oracle.  engine-vs-fp32 = implementation error + the precision cost of
    bf16 activations.  At random initialisation a ResNet-50 amplifies a 2^-9 perturbation of its INPUT IMAGE to 3 % of
    the pooled feature (measured, DESIGN.md "Precision"), so the fp32 comparison is made on conditioned weights (a few
    fp32 Adam steps of the oracle) and over a 50-step training trajectory.
Tolerances are written next to each assert; measured values are printed (pytest -s) and recorded in DESIGN.md.
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import statistical

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


@pytest.fixture(scope="module", autouse=True)
def _fp32_math():
    assert torch.cuda.is_available()
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def l2err(got, want):
    return ((got.float() - want.float()).norm() / (want.float().norm() + 1e-20)).item()


def make_data(n, size, classes, seed=0, dataset="optic"):
    from aadg_b200.synth import fundus_batch, vessel_batch
    rng = np.random.RandomState(seed)
    if dataset == "optic":
        imgs, masks = fundus_batch(n, size, size, seed=seed + 5)
    else:
        imgs, masks = vessel_batch(n, size, size, seed=seed + 5)
    for i in range(n):                       # per-sample colour / contrast variety
        imgs[i] = np.clip(imgs[i].astype(np.float32) * rng.uniform(0.5, 1.3) + rng.uniform(-40, 40, 3), 0, 255)
    x = (torch.from_numpy(imgs).cuda().permute(0, 3, 1, 2).float() / 127.5 - 1.0).contiguous()
    m = torch.from_numpy(masks).cuda()
    if dataset == "optic":
        target = torch.stack([(m <= 50).float(), (m <= 200).float()], 1)[:, :classes].contiguous()
    else:
        target = (m != 0).float().unsqueeze(1).contiguous()
    return x, target


def make_models(arch, encoder, classes, seed=0, presteps=0, x=None, target=None):
    """(fp32 oracle, bf16-storage oracle, engine) with identical weights; `presteps` fp32 Adam steps of the oracle on
    (x, target) first, so that the comparison is made on conditioned weights rather than on a random initialisation."""
    from aadg_b200.nn import DeepLabV3Plus, Unet
    from oracle import segnet_bf16
    from oracle.segnet_torch import DeepLabV3PlusTorch, UnetTorch
    torch.manual_seed(seed)
    ref = (DeepLabV3PlusTorch if arch == "deeplabv3plus" else UnetTorch)(encoder, classes).cuda().train()
    for m in ref.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    if presteps:
        opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
        for _ in range(presteps):
            logits, _ = ref(x)
            loss = F.binary_cross_entropy(torch.sigmoid(logits), target)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        ref.zero_grad(set_to_none=True)
    twin = segnet_bf16.install(copy.deepcopy(ref))
    ctor = DeepLabV3Plus if arch == "deeplabv3plus" else Unet
    net = ctor(encoder_name=encoder, encoder_weights=None, in_channels=3, classes=classes,
               aux_params=dict(pooling="avg"))
    net.load_state_dict(ref.state_dict())
    net.dropout_enabled = False
    return ref, twin, net


def grad_table(net, ref):
    """[(name, cosine, norm ratio)] of every parameter gradient of the engine against the oracle's"""
    from test_nn_gpu import _my_grad_as_torch
    rg = {k: v.grad for k, v in ref.named_parameters() if v.grad is not None}
    rows = []
    for name, p in net.named_params().items():
        if name not in rg:
            continue
        g, w = _my_grad_as_torch(name, p).reshape(-1).double(), rg[name].reshape(-1).double()
        if w.norm() < 1e-12:
            continue
        rows.append((name, (g @ w / (g.norm() * w.norm() + 1e-30)).item(), (g.norm() / w.norm()).item()))
    return rows


def grad_table_torch(model, ref):
    """the same table for two torch modules with identical parameter names"""
    rg = {k: v.grad for k, v in ref.named_parameters() if v.grad is not None}
    rows = []
    for name, p in model.named_parameters():
        if p.grad is None or name not in rg:
            continue
        g, w = p.grad.reshape(-1).double(), rg[name].reshape(-1).double()
        if w.norm() < 1e-12:
            continue
        rows.append((name, (g @ w / (g.norm() * w.norm() + 1e-30)).item(), (g.norm() / w.norm()).item()))
    return rows


def oracle_step(model, x, target):
    model.zero_grad(set_to_none=True)
    logits, pooled = model(x)
    loss = F.binary_cross_entropy(torch.sigmoid(logits), target)
    loss.backward()
    return logits.detach(), pooled.detach(), loss.item()


CASES = [
    # arch, encoder, classes, size, n, dataset
    ("deeplabv3plus", "resnet18", 2, 128, 8, "optic"),
    ("deeplabv3plus", "resnet50", 2, 512, 8, "optic"),          # BASELINE config 2 geometry (512^2, ResNet-50)
    ("deeplabv3plus", "mobilenet_v2", 2, 256, 144, "optic"),    # the reference-native step: 144 x 256^2, MobileNetV2
    ("unet", "resnet34", 1, 512, 4, "vessel"),                  # config 4 family at 512^2
]


def nhwc(x):
    return x.permute(0, 2, 3, 1).clone(memory_format=torch.contiguous_format)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


def layer_residuals(arch, encoder, classes, size, n, dataset, fp64=False):
    """Every convolution / depthwise convolution / batch-norm(+activation) of the engine, fed the bf16-storage oracle's
    OWN input for that layer (exactly bf16-representable) and compared with the oracle's output: [(relative L2 residual,
    kind, layer name, input shape, note)].  With identical rounding points the legitimate residual is the summation
    order inside one layer: a few 1e-5 .. 2e-4 after bf16 rounding (a fraction of a percent of the elements land on the
    other side of a rounding boundary); a semantic difference (wrong tap, border, stride, statistics) shows as >= 1e-3."""
    from aadg_b200.nn import network as NW
    from aadg_b200.ops import conv as C
    from aadg_b200.ops import nn as K
    x, target = make_data(n, size, classes, dataset=dataset)
    ref, twin, net = make_models(arch, encoder, classes)
    mods = dict(twin.named_modules())
    io = {}
    for name, m in mods.items():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.BatchNorm2d)):
            m.register_forward_hook(lambda mod, inp, out, name=name: io.__setitem__(name, (inp[0].detach(), out.detach().clone())))
    with torch.no_grad():
        twin(x)
    rows, seen = [], set()

    def exact_conv(m, xin):
        w64 = m.weight.detach().to(BF).double() if m.groups == 1 else m.weight.detach().double()
        return F.conv2d(xin.double(), w64, None, m.stride, m.padding, m.dilation, m.groups).to(BF).float()

    def check_bn(bn, conv_out, relu, relu6):
        if bn.name not in io:
            return
        want = io[bn.name][1]                         # BN output, rounded to bf16 unless it feeds a residual add
        xi = nhwc(conv_out).to(BF)
        c = xi.shape[-1]
        buf = torch.zeros(6, c, device=xi.device)
        K.bn_stats(xi, buf[0], buf[1])
        K.bn_finalize(buf[0], buf[1], bn.gamma.data, bn.beta.data, xi.numel() // c, NW.BN_EPS, NW.BN_MOMENTUM, buf[2], buf[3],
                      buf[4], buf[5], None, None)
        y = torch.empty_like(xi)
        K.bn_apply(xi, buf[4], buf[5], y, relu=False)
        rounded = bool(torch.equal(want, want.to(BF).float()))
        rows.append((l2err(nchw(y), want.to(BF).float()), "bn" if rounded else "bn (oracle unrounded here)", bn.name,
                     tuple(conv_out.shape), ""))

    def walk(o):
        if id(o) in seen:
            return
        seen.add(id(o))
        if isinstance(o, NW.ConvBN):
            cname = o.w.name[:-len(".weight")]
            if cname in io:
                xin, want = io[cname]
                xi = nhwc(xin).to(BF)
                P = o._pack_factor(xi)
                got = o._fprop_packed(xi, P, False) if P else C.fprop(xi, o.w.bf16, o.k, o.k, o.stride, o.pad, o.dil)
                note = "k%d s%d d%d%s" % (o.k, o.stride, o.dil, " pixel-packed x%d" % P if P else "")
                if fp64:
                    ex = exact_conv(mods[cname], xin)
                    note += " | engine vs fp64 %.2e, oracle vs fp64 %.2e" % (l2err(nchw(got), ex), l2err(want, ex))
                rows.append((l2err(nchw(got), want), "conv", cname, tuple(xin.shape), note))
                check_bn(o.bn, want, o.relu, o.relu6)
        elif isinstance(o, NW.Depthwise3x3):
            cname = o.w.name[:-len(".weight")]
            if cname in io:
                xin, want = io[cname]
                xi = nhwc(xin).to(BF)
                got = torch.empty((xi.shape[0], want.shape[2], want.shape[3], xi.shape[3]), dtype=BF, device=xi.device)
                K.dwconv3x3(xi, o.w.data, o.dil, got, stride=o.stride)
                rows.append((l2err(nchw(got), want), "dwconv", cname, tuple(xin.shape), "s%d d%d" % (o.stride, o.dil)))
        elif isinstance(o, NW.DepthwiseBN):
            dname = o.dw.w.name[:-len(".weight")]
            if dname in io:
                check_bn(o.bn, io[dname][1], True, True)
        if isinstance(o, (list, tuple)):
            for i in o:
                walk(i)
        elif hasattr(o, "__dict__") and not isinstance(o, (NW.ParamStore, NW.Param, torch.Tensor)):
            for v in vars(o).values():
                walk(v)
    walk(net.encoder)
    walk(net.decoder)
    if hasattr(net.encoder, "stem_w"):
        cname = net.encoder.stem_w.name[:-len(".weight")]
        xin, want = io[cname]
        if encoder == "mobilenet_v2":      # 3x3 / stride-2 stem: a 1x1 GEMM over im2col patches
            col = K.im2col_stem(xin.contiguous(), 3, 3, 2, 1, 3 * NW.MBV2_STEM_RP, row_pitch=NW.MBV2_STEM_RP)
            got = C.fprop(col, net.encoder.stem_w.bf16, 1, 1)
        else:                              # 7x7 / stride-2 stem: window convolution over the space-to-depth image
            _, got = net.encoder.stem_fprop(xin.contiguous())
        rows.append((l2err(nchw(got), want), "stem conv", cname, tuple(xin.shape), ""))
        check_bn(net.encoder.stem_bn, want, True, encoder == "mobilenet_v2")
    rows.sort(reverse=True)
    return rows, (ref, twin, net, x, target)


LAYER_CASES = [
    ("deeplabv3plus", "resnet50", 2, 512, 4, "optic"),          # every layer geometry of the benchmarked step
    ("deeplabv3plus", "mobilenet_v2", 2, 256, 16, "optic"),     # the reference-native backbone at its native size
    ("unet", "resnet34", 1, 512, 2, "vessel"),                  # config 4 family, pixel-packed 16/32-channel layers included
]


@pytest.mark.parametrize("arch,encoder,classes,size,n,dataset", LAYER_CASES)
def test_every_layer_teacher_forced_vs_bf16_storage_oracle(arch, encoder, classes, size, n, dataset):
    """implementation parity, layer by layer at the real shapes: each of the engine's forward layers reproduces the
    oracle's output for the oracle's own input to 5e-4 relative L2 (measured: median 3e-5, worst 2e-4)."""
    rows, _ = layer_residuals(arch, encoder, classes, size, n, dataset)
    assert len(rows) >= 40, len(rows)
    print("LAYERS %s/%s %d^2 n=%d: %d layers, worst %.2e (%s %s), median %.2e" %
          (arch, encoder, size, n, len(rows), rows[0][0], rows[0][1], rows[0][2], rows[len(rows) // 2][0]))
    bad = [r for r in rows if r[0] > 5e-4]
    assert not bad, bad[:6]


def backward_layer_residuals(arch, encoder, classes, size, n, dataset):
    """The BACKWARD kernels of every convolution / depthwise convolution of the engine at the real shapes, teacher-forced:
    fed the bf16-storage oracle's own layer input x and its own gradient dy of the layer output (rounded to bf16, as the
    engine stores gradients) and compared with the exact (float64) data and weight gradients of those very tensors:
    [(relative L2 residual, kind, layer, shapes)].  Identical inputs and masks on both sides, so the residual is fp32
    accumulation order: weight gradients are fp32 outputs; data gradients are compared after rounding the exact result
    to bf16 too, so only elements that land on the other side of a rounding boundary differ.  A wrong tap, border, stride,
    split or tile choice shows as >= 1e-2."""
    from aadg_b200.nn import network as NW
    from aadg_b200.ops import conv as C
    from aadg_b200.ops import nn as K
    x, target = make_data(n, size, classes, dataset=dataset)
    ref, twin, net = make_models(arch, encoder, classes)
    mods = dict(twin.named_modules())
    io, gout = {}, {}
    for name, m in mods.items():
        if isinstance(m, torch.nn.Conv2d):
            m.register_forward_hook(lambda mod, inp, out, name=name: io.__setitem__(name, inp[0].detach()))
            m.register_full_backward_hook(lambda mod, gi, go, name=name: gout.__setitem__(name, go[0].detach()))
    logits, _ = twin(x)
    F.binary_cross_entropy(torch.sigmoid(logits), target).backward()
    rows, seen = [], set()

    def exact(m, xin, dy, need_dx):
        w64 = (m.weight.detach().to(BF) if m.groups == 1 else m.weight.detach()).double()
        dw = torch.nn.grad.conv2d_weight(xin.double(), w64.shape, dy.double(), m.stride, m.padding, m.dilation, m.groups)
        dx = torch.nn.grad.conv2d_input(xin.shape, w64, dy.double(), m.stride, m.padding, m.dilation, m.groups) \
            if need_dx else None
        return dw, dx

    def walk(o):
        if id(o) in seen:
            return
        seen.add(id(o))
        if isinstance(o, NW.ConvBN):
            cname = o.w.name[:-len(".weight")]
            if cname in gout:
                m = mods[cname]
                xin, dy = io[cname], gout[cname].to(BF).float()
                if float(dy.abs().max()) > 0:
                    dw64, dx64 = exact(m, xin, dy, o.need_dgrad)
                    xi, dyi = nhwc(xin).to(BF), nhwc(dy).to(BF)
                    P = o._pack_factor(xi)
                    nb, h, w_, _ = xi.shape
                    if P:
                        if o.packed is None:
                            o.packed = NW.PackedTaps(o.cout, o.cin, P, xi.device)
                        x4, d4 = xi.view(nb, h, w_ // P, 64), dyi.view(nb, h, w_ // P, P * o.cout)
                        dw = torch.zeros((9, o.cout, o.cin), dtype=torch.float32, device=xi.device)
                        o.packed.fold_grad(C.wgrad(x4, d4, 3, 3, 1, 1, 1), dw)
                        dx = C.dgrad(d4, o.packed.expand_t(o.w.bf16), 3, 3, 1, 1, 1, (h, w_ // P)).view(nb, h, w_, o.cin) \
                            if o.need_dgrad else None
                    else:
                        dw = C.wgrad(xi, dyi, o.k, o.k, o.stride, o.pad, o.dil)
                        dx = C.dgrad(dyi, o.w.bf16_t, o.k, o.k, o.stride, o.pad, o.dil, (h, w_)) if o.need_dgrad else None
                    note = "%s k%d s%d d%d%s" % (tuple(xin.shape), o.k, o.stride, o.dil, " pixel-packed x%d" % P if P else "")
                    dw_t = dw.reshape(o.k, o.k, o.cout, o.cin).permute(2, 3, 0, 1)
                    rows.append((l2err(dw_t, dw64), "wgrad", cname, note))
                    if dx is not None:
                        rows.append((l2err(nchw(dx), dx64.to(BF)), "dgrad", cname, note))
        elif isinstance(o, NW.Depthwise3x3):
            cname = o.w.name[:-len(".weight")]
            if cname in gout:
                m = mods[cname]
                xin, dy = io[cname], gout[cname].to(BF).float()
                dw64, dx64 = exact(m, xin, dy, True)
                xi, dyi = nhwc(xin).to(BF), nhwc(dy).to(BF)
                dw = torch.zeros((9, o.c), dtype=torch.float32, device=xi.device)
                K.dwconv3x3_wgrad(xi, dyi, o.dil, dw, stride=o.stride)
                dx = torch.empty(xi.shape, dtype=BF, device=xi.device)
                K.dwconv3x3(dyi, o.w.data, o.dil, dx, backward_data=True, stride=o.stride)
                note = "%s s%d d%d" % (tuple(xin.shape), o.stride, o.dil)
                rows.append((l2err(dw.t().reshape(o.c, 1, 3, 3), dw64), "dw wgrad", cname, note))
                rows.append((l2err(nchw(dx), dx64.to(BF)), "dw dgrad", cname, note))
        if isinstance(o, (list, tuple)):
            for i in o:
                walk(i)
        elif hasattr(o, "__dict__") and not isinstance(o, (NW.ParamStore, NW.Param, torch.Tensor)):
            for v in vars(o).values():
                walk(v)
    walk(net.encoder)
    walk(net.decoder)
    if hasattr(net.encoder, "stem_w"):               # stem weight gradient
        cname = net.encoder.stem_w.name[:-len(".weight")]
        m = mods[cname]
        xin, dy = io[cname], gout[cname].to(BF).float()
        dw64, _ = exact(m, xin.to(BF).float(), dy, False)
        if encoder == "mobilenet_v2":
            col = K.im2col_stem(xin.contiguous(), 3, 3, 2, 1, 3 * NW.MBV2_STEM_RP, row_pitch=NW.MBV2_STEM_RP)
            dw = NW.stem_unpack_rows(C.wgrad(col, nhwc(dy).to(BF), 1, 1, 1, 0, 1), 3, 3, NW.MBV2_STEM_RP)
        else:
            xw, _ = net.encoder.stem_fprop(xin.contiguous())
            dw = NW.stem_unpack(net.encoder.stem_wgrad(xw, nhwc(dy).to(BF)))
        rows.append((l2err(dw, dw64), "stem wgrad", cname, str(tuple(xin.shape))))
    rows.sort(reverse=True)
    return rows


@pytest.mark.parametrize("arch,encoder,classes,size,n,dataset", LAYER_CASES)
def test_every_layer_backward_teacher_forced_vs_float64(arch, encoder, classes, size, n, dataset):
    """backward implementation parity, layer by layer at the real shapes (tile selection, splits, halo and packed variants
    included): weight gradients within 2e-3 and (bf16-rounded) data gradients within 1e-3 relative L2 of the exact
    float64 gradients of the same bf16 tensors."""
    rows = backward_layer_residuals(arch, encoder, classes, size, n, dataset)
    assert len(rows) >= 60, len(rows)
    worst = {}
    for r in rows:
        worst.setdefault(r[1], r)
    print("LAYERS-BWD %s/%s %d^2 n=%d: %d checks; worst per kind: %s" %
          (arch, encoder, size, n, len(rows), {k: ("%.2e" % v[0], v[2]) for k, v in worst.items()}))
    tol = {"wgrad": 2e-3, "stem wgrad": 2e-3, "dw wgrad": 2e-3, "dgrad": 1e-3, "dw dgrad": 1e-3}
    bad = [r for r in rows if r[0] > tol[r[1]]]
    assert not bad, bad[:6]


@pytest.mark.parametrize("arch,encoder,classes,size,n,dataset", CASES)
@statistical
def test_engine_within_the_bf16_noise_envelope_at_random_init(arch, encoder, classes, size, n, dataset):
    """Whole network at random initialisation.  bf16 storage re-quantises every layer, so ANY difference in fp32
    summation order (the engine's own two runs differ: its statistics and weight gradients use fp32 atomics) grows
    within a few layers to the bf16 noise floor and no further.  The engine therefore cannot sit closer to the
    bf16-storage oracle than that floor; the test asserts that it sits INSIDE it: its distance to the bf16-storage oracle
    is no larger than that oracle's own distance to the fp32 oracle (measured: about half), and the loss agrees with
    the fp32 oracle as well as the bf16-storage oracle does.  Printed: the engine's run-to-run spread."""
    x, target = make_data(n, size, classes, dataset=dataset)
    ref, twin, net = make_models(arch, encoder, classes)
    with torch.no_grad():
        logits_r, pooled_r = ref(x)
        loss_r = F.binary_cross_entropy(torch.sigmoid(logits_r), target).item()
        logits_t, pooled_t = twin(x)
        loss_t = F.binary_cross_entropy(torch.sigmoid(logits_t), target).item()
    net.store.zero_grad()
    out = net.loss_step(x, target, want_logits=True)
    loss_e, pooled_e, logits_e = out["loss"].item(), out["pooled"].clone(), out["logits"].clone()
    net.store.zero_grad()
    out2 = net.loss_step(x, target, want_logits=True)
    env_pool, env_logit, env_loss = l2err(pooled_t, pooled_r), l2err(logits_t, logits_r), abs(loss_t - loss_r) / loss_r
    e_pool, e_logit = l2err(pooled_e, pooled_t), l2err(logits_e, logits_t)
    e_loss_t, e_loss_r = abs(loss_e - loss_t) / loss_t, abs(loss_e - loss_r) / loss_r
    print("ENVELOPE %s/%s %d^2 n=%d: engine vs bf16-oracle pooled %.2e logits %.2e loss %.2e | bf16-oracle vs fp32 pooled "
          "%.2e logits %.2e loss %.2e | engine vs fp32 loss %.2e | engine run-to-run pooled %.2e logits %.2e loss %.2e" %
          (arch, encoder, size, n, e_pool, e_logit, e_loss_t, env_pool, env_logit, env_loss, e_loss_r,
           l2err(out2["pooled"], pooled_e), l2err(out2["logits"], logits_e), abs(out2["loss"].item() - loss_e) / loss_e))
    assert e_pool <= env_pool and e_logit <= env_logit, (e_pool, env_pool, e_logit, env_logit)
    # the loss is ONE scalar draw of that noise (random sign, pixel-averaged): bounded by a multiple of the oracle pair's
    # (largest value seen in 8 runs per case: 3.2x the pair's at ResNet-50 @512^2, so x4 would fail about one run in 50)
    assert e_loss_r <= max(6.0 * env_loss, 2e-3), (e_loss_r, env_loss)
    assert e_loss_t <= max(6.0 * env_loss, 2e-3), (e_loss_t, env_loss)


@pytest.mark.parametrize("arch,encoder,classes,size,n,dataset", CASES[:3])
@statistical
def test_engine_vs_fp32_oracle_on_conditioned_weights(arch, encoder, classes, size, n, dataset):
    """the north-star tolerance against the PLAIN fp32 oracle, on weights conditioned by 12 fp32 Adam steps of the
    oracle.  ResNets: loss within 5e-4 relative, logits within 2e-2 relative L2, Dice (samplewise F1) within 1e-3.
    MobileNetV2 stores un-activated 16..320-channel bottleneck tensors in bf16: the bf16-storage ORACLE itself is 3e-3 from
    fp32 there (printed), and the engine is held to 1.5x that oracle's own error instead.  Gradients: cosine >= 0.95 for
    every parameter whose gradient is well conditioned (cosine(bf16-storage oracle, fp32 oracle) >= 0.99; batch-norm biases
    that a following convolution + batch-norm cancels analytically carry pure rounding noise in all three)."""
    from aadg_b200.nn.network import dice_from_counts
    from oracle.segnet_torch import f1_samplewise
    x, target = make_data(n, size, classes, dataset=dataset)
    sub = slice(0, min(n, 16))     # conditioning uses a sub-batch (cheap); the comparison uses the whole batch
    ref, twin, net = make_models(arch, encoder, classes, presteps=12, x=x[sub], target=target[sub])
    logits, pooled, loss = oracle_step(ref, x, target)
    logits_t, pooled_t, loss_twin = oracle_step(twin, x, target)
    net.store.zero_grad()
    out = net.loss_step(x, target, want_logits=True)
    e_loss = abs(out["loss"].item() - loss) / abs(loss)
    t_loss = abs(loss_twin - loss) / abs(loss)
    e_pool, e_logit = l2err(out["pooled"], pooled), l2err(out["logits"], logits)
    t_logit = l2err(logits_t, logits)
    dice = dice_from_counts(out["counts"])
    want_dice = [f1_samplewise(torch.sigmoid(logits)[:, k], target[:, k]).item() for k in range(classes)]
    twin_dice = [f1_samplewise(torch.sigmoid(logits_t)[:, k], target[:, k]).item() for k in range(classes)]
    e_dice = max(abs(dice[k].item() - want_dice[k]) for k in range(classes))
    t_dice = max(abs(twin_dice[k] - want_dice[k]) for k in range(classes))
    rows = grad_table(net, ref)
    twin_cos = {name: cos for name, cos, _ in grad_table_torch(twin, ref)}
    # every parameter gradient: the engine is as close to the fp32 oracle as the bf16-storage oracle is (x3 + 0.02 in
    # 1 - cosine).  Gradients that bf16 storage leaves well conditioned (twin cosine >= 0.99) are additionally held to 0.97.
    offenders = sorted(((name, round(cos, 4), round(twin_cos.get(name, 1.0), 4)) for name, cos, _ in rows
                        if (1 - cos) > 3 * (1 - twin_cos.get(name, 1.0)) + 0.02), key=lambda r: r[1])
    # The ASPP pooling branch normalises n values per channel (its map is 1x1): a channel whose n pre-BN values nearly
    # coincide has invstd up to eps^-1/2 = 316 and a gradient row that is amplified noise.  Repeating one engine step 300
    # times on fixed weights (scripts/stress_repeat.py, profiles/r02_stress_repeat_resnet18.log, n = 8) gives that
    # convolution's weight gradient a median run-to-run cosine of 0.9991 with a tail down to 0.948, 88 % of the deviation
    # in ONE output channel; against the fp32 oracle one suite run in ~15 saw 0.81.  Those parameters are held to a
    # sanity floor here and count fully in the whole-gradient cosine below.
    chaotic = lambda nm: nm.startswith("decoder.aspp.0.convs.4.")      # noqa: E731
    pool_rows = [r for r in rows if chaotic(r[0])]
    rows_judged = [r for r in rows if not chaotic(r[0])]
    offenders = [o for o in offenders if not chaotic(o[0])]
    judged = sorted((r for r in rows_judged if twin_cos.get(r[0], 0.0) >= 0.99), key=lambda r: r[1])

    def flat(table_names, getter):
        return torch.cat([getter(nm).reshape(-1).double() for nm in table_names])
    from test_nn_gpu import _my_grad_as_torch
    names = [r[0] for r in rows]
    rgrad = dict(ref.named_parameters())
    tgrad = dict(twin.named_parameters())
    g_ref = flat(names, lambda nm: rgrad[nm].grad)
    g_eng = flat(names, lambda nm: _my_grad_as_torch(nm, net.named_params()[nm]))
    g_twin = flat(names, lambda nm: tgrad[nm].grad)
    cos_e = (g_eng @ g_ref / (g_eng.norm() * g_ref.norm())).item()
    cos_t = (g_twin @ g_ref / (g_twin.norm() * g_ref.norm())).item()
    print("PARITY fp32-oracle (conditioned) %s/%s %d^2 n=%d: loss %.5f rel %.2e (bf16-storage oracle alone: %.2e) pooled "
          "%.2e logits %.2e (oracle alone %.2e) dice abs %.2e (oracle alone %.2e) | whole gradient cosine vs fp32: engine "
          "%.5f, bf16-storage oracle %.5f | %d of %d gradients well conditioned under bf16 storage, engine min cos there "
          "%.4f | offenders %s" % (arch, encoder, size, n, loss, e_loss, t_loss, e_pool, e_logit, t_logit, e_dice, t_dice,
                                  cos_e, cos_t, len(judged), len(rows), judged[0][1] if judged else float("nan"),
                                  offenders[:4]))
    # (the loss is one scalar draw of the storage noise: x3; the logits are an L2 norm over millions of draws: x1.5).
    # Loss and Dice errors of engine and bf16-storage oracle are two single draws: when the oracle's happens to be small
    # the multiple alone is not a bound, so each case also has a floor at ~2x the largest value EITHER side showed in ten
    # runs (MobileNetV2: loss 3.1e-3, Dice 1.2e-2 for the kernel-free oracle itself; ResNets: the hard limits below)
    mbv2 = encoder == "mobilenet_v2"
    assert e_loss <= max(6e-3 if mbv2 else 5e-4, 3.0 * t_loss), (e_loss, t_loss)
    assert e_logit <= max(2e-2, 1.5 * t_logit), (e_logit, t_logit)
    assert e_dice <= max(2.5e-2 if mbv2 else 1e-3, 2.0 * t_dice), (e_dice, t_dice)
    if encoder != "mobilenet_v2":
        assert e_loss <= 5e-4 and e_logit <= 2e-2 and e_dice <= 1e-3
    assert (1 - cos_e) <= 2 * (1 - cos_t) + 2e-3, (cos_e, cos_t)
    assert len(offenders) <= 0.02 * len(rows), offenders[:8]
    assert not judged or judged[0][1] >= 0.97, judged[:4]
    assert all(r[1] >= 0.3 for r in pool_rows), pool_rows


@statistical
def test_training_trajectory_50_steps_vs_fp32_oracle():
    """50 Adam steps from identical weights on identical data (search_dg.py:140-142,164-172): the engine's loss curve
    stays within 4 % of the fp32 oracle's and its Dice within 0.04 at every step (measured: 1.7-2.6 % / 0.008-0.025;
    bf16-storage oracle on the CPU: 1.5 % / 0.02, DESIGN.md "Precision"); both learn (loss falls by > 10x)."""
    from aadg_b200.nn.network import dice_from_counts
    from oracle.segnet_torch import f1_samplewise
    x, target = make_data(8, 128, 2)
    ref, twin, net = make_models("deeplabv3plus", "resnet18", 2)
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    curve = []
    for step in range(50):
        logits, _ = ref(x)
        loss = F.binary_cross_entropy(torch.sigmoid(logits), target)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        d_ref = f1_samplewise(torch.sigmoid(logits.detach())[:, 1], target[:, 1]).item()
        net.store.zero_grad()
        out = net.loss_step(x, target)
        net.store.adam_step(1e-3)
        curve.append((loss.item(), out["loss"].item(), d_ref, dice_from_counts(out["counts"])[1].item()))
    rel = [abs(b - a) / a for a, b, _, _ in curve]
    dd = [abs(d - c) for _, _, c, d in curve]
    print("TRAJECTORY resnet18 128^2 n=8: loss %.4f -> oracle %.5f / engine %.5f; max rel %.3e at step %d; max dice diff "
          "%.4f; step0 rel %.2e" % (curve[0][0], curve[-1][0], curve[-1][1], max(rel), int(np.argmax(rel)), max(dd), rel[0]))
    assert rel[0] <= 1e-3, rel[0]
    # maxima over 50 steps seen in eight runs: loss 1.7-2.6 %, Dice 0.008-0.025; the bounds leave 1.5x on the largest
    assert max(rel) <= 4e-2, (max(rel), int(np.argmax(rel)))
    assert max(dd) <= 4e-2, max(dd)
    assert curve[-1][1] < 0.1 * curve[0][1]


@statistical
def test_autograd_surface_runs_the_reference_training_lines():
    """`seg_output, feature = model(input)` ... `model_optimizer.zero_grad(); seg_loss.backward(); model_optimizer.step()`
    (search_dg.py:132,140-142,170-172) verbatim, with torch.optim.Adam over `model.parameters()`, against the engine's own
    fused loss_step + Adam on an identical model.  The engine is not bitwise reproducible (fp32 atomics re-quantised by
    bf16 storage), so "same" is measured against the spread of a THIRD identical model stepped the fused way: the autograd
    surface may differ from the fused path by at most 4x what two fused runs differ from each other (+ small floors)."""
    from aadg_b200.nn import DeepLabV3Plus
    M = 2
    x, target = make_data(8, 128, 2)
    a, b, c = [DeepLabV3Plus(encoder_name="resnet18", encoder_weights=None, in_channels=3, classes=2,
                             aux_params=dict(pooling="avg"), seed=4) for _ in range(3)]
    assert torch.equal(a.store.params, b.store.params) and torch.equal(a.store.params, c.store.params)
    model_optimizer = torch.optim.Adam(b.parameters(), lr=1e-3)
    model_criterion = torch.nn.BCELoss()

    def cosine(u, v):
        u, v = u.double(), v.double()
        return (u @ v / (u.norm() * v.norm())).item()
    for step in range(3):
        seg_output, feature = b(x)
        assert seg_output.requires_grad and seg_output.shape == (8, 2, 128, 128) and feature.shape == (8, 512)
        seg_soft = torch.sigmoid(seg_output)
        seg_loss = torch.mean(torch.stack([model_criterion(seg_soft[j::M], target[j::M]) for j in range(M)]))
        model_optimizer.zero_grad()
        seg_loss.backward()
        a.store.zero_grad()
        out = a.loss_step(x, target)
        c.store.zero_grad()
        out_c = c.loss_step(x, target)
        la, lb, lc = out["loss"].item(), seg_loss.item(), out_c["loss"].item()
        rel, spread = abs(lb - la) / la, abs(lc - la) / la
        cos, cos_spread = cosine(a.store.grads, b.store.grads), cosine(a.store.grads, c.store.grads)
        ratio = (b.store.grads.double().norm() / a.store.grads.double().norm()).item()
        ratio_spread = abs((c.store.grads.double().norm() / a.store.grads.double().norm()).item() - 1)
        print("AUTOGRAD step %d: loss rel %.2e (fused run-to-run %.2e) grad cos %.6f (run-to-run %.6f) norm ratio %.5f" %
              (step, rel, spread, cos, cos_spread, ratio))
        # `rel` and `spread` are single draws of the same noise (measured over ten runs on different boxes: step 0
        # 0.7-3.7e-4, step 1 0.2-3.7e-3, step 2 4.5-9.9e-3 for both), so besides the x4 yardstick the floors sit three
        # times above the largest value seen; a wrong gradient moves the step-1 loss by percents
        assert rel <= max(4 * spread, (1.5e-3, 1.2e-2, 3e-2)[step]), (step, rel, spread)
        assert 1 - cos <= max(4 * (1 - cos_spread), 1e-3 if step == 0 else 2e-2), (step, cos, cos_spread)
        assert abs(ratio - 1) <= max(4 * ratio_spread, 0.1 if step == 0 else 0.25), (step, ratio, ratio_spread)
        for name, leaf in b.named_parameters():
            assert leaf.grad is not None and leaf.grad.data_ptr() == b.named_params()[name].grad.data_ptr()
        model_optimizer.step()
        a.store.adam_step(1e-3)
        c.store.adam_step(1e-3)
    e_b, e_c = l2err(b.store.params, a.store.params), l2err(c.store.params, a.store.params)
    print("AUTOGRAD params after 3 steps: surface vs fused %.2e, fused run-to-run %.2e" % (e_b, e_c))
    assert e_b <= max(4 * e_c, 1e-3), (e_b, e_c)
    # the pooled feature is differentiable too (the reference detaches it; a caller need not)
    b.zero_grad()
    seg_output, feature = b(x)
    feature.square().mean().backward()
    g = b.store.grads
    assert torch.isfinite(g).all() and float(g.abs().sum()) > 0
    dec_w = b.named_params()["decoder.block2.1.weight"].grad
    assert float(dec_w.abs().sum()) == 0.0            # nothing flowed through the decoder
