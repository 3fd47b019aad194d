import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_nn_gpu import _pair
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
BF = torch.bfloat16
l2 = lambda a, b: ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()
nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous().to(BF)
nchw = lambda t: t.float().permute(0, 3, 1, 2)
for enc, size, n in [("resnet50", 128, 4)]:
    ref, net, x, target = _pair(enc, 2, size, n)
    with torch.no_grad():
        feats = ref.encoder(x)
    # stage-level teacher forcing
    f1 = feats[1]
    pooled = ref.encoder.maxpool(f1)
    ins = [pooled, feats[2], feats[3], feats[4]]
    for li in range(4):
        xin = nhwc(ins[li])
        with torch.no_grad():
            want = getattr(ref.encoder, "layer%d" % (li + 1))(nchw(xin))
        y = xin
        for bi, blk in enumerate(net.encoder.blocks[li]):
            y_prev = y
            y = blk.forward(y, True)
            with torch.no_grad():
                wb = getattr(ref.encoder, "layer%d" % (li + 1))[bi](nchw(y_prev))
            print("  layer%d.%d block-level relL2" % (li + 1, bi), l2(nchw(y), wb))
        print("layer%d stage relL2" % (li + 1), l2(nchw(y), want))
    # decoder teacher forcing
    fb = [None, nhwc(feats[1]), nhwc(feats[2]), nhwc(feats[3]), nhwc(feats[4]), nhwc(feats[5])]
    with torch.no_grad():
        for m in ref.modules():
            if isinstance(m, torch.nn.Dropout): m.p = 0.0
        want = ref.decoder(*[None if f is None else nchw(f) for f in fb])
        a_ref = ref.decoder.aspp[0](nchw(fb[5]))
        cat_ref = torch.cat([c(nchw(fb[5])) for c in ref.decoder.aspp[0].convs], 1)
    got = net.decoder.forward(fb, True, None)
    print("decoder relL2", l2(nchw(got), want))
