import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_nn_gpu import _pair, nchw, nhwc, l2err
from aadg_b200.ops import nn as K
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
BF = torch.bfloat16
ref, net, x, target = _pair("resnet18", 2, 128, 4)
with torch.no_grad():
    feats = ref.encoder(x)
f5 = nhwc(feats[5]).to(BF)
dec = net.decoder
aspp = ref.decoder.aspp[0]
n, h, w, c = f5.shape
for i in range(5):
    xr = nchw(f5).requires_grad_(True)
    o = aspp.convs[i](xr)
    dy = torch.randn_like(o).to(BF)
    o.backward(dy.float())
    if i == 0:
        y = dec.b0.forward(f5, True)
        dx = dec.b0.backward(nhwc(dy).to(BF))
    elif i < 4:
        y = dec.br[i - 1].forward(f5, True)
        dx = dec.br[i - 1].backward(nhwc(dy).to(BF))
    else:
        pooled = K.f32_to_bf16(K.global_sum(f5, 1.0 / (h * w))).view(n, 1, 1, c)
        pv = dec.bp.forward(pooled, True)
        y = torch.empty(n, h, w, 256, device="cuda", dtype=BF)
        K.broadcast_pixels(pv, y)
        dpv = K.f32_to_bf16(K.global_sum(nhwc(dy).to(BF), 1.0)).view(n, 1, 1, 256)
        dpooled = dec.bp.backward(dpv)
        dx = torch.empty(f5.shape, dtype=BF, device="cuda")
        K.broadcast_pixels((dpooled.float() / (h * w)).to(BF), dx)
    print("branch", i, "fwd", l2err(nchw(y), o), "dx", l2err(nchw(dx), xr.grad), "norm dx", xr.grad.norm().item())
