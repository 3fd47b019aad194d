import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_nn_gpu import _pair, nchw, nhwc, l2err
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
BF = torch.bfloat16
ref, net, x, target = _pair("resnet18", 2, 128, 4)
with torch.no_grad():
    feats = ref.encoder(x)
    xin = nhwc(ref.encoder.maxpool(feats[1])).to(BF)
blk_r = ref.encoder.layer1[0]
blk = net.encoder.blocks[0][0]
xr = nchw(xin).requires_grad_(True)
o1 = blk_r.relu(blk_r.bn1(blk_r.conv1(xr))); o1.retain_grad()
pre2 = blk_r.conv2(o1); pre2.retain_grad()
o2 = blk_r.bn2(pre2); o2.retain_grad()
out = F.relu(o2 + xr)
y = blk.forward(xin, True)
print("fwd", l2err(nchw(y), out))
dy = (torch.randn_like(out) * (out > 0)).to(BF)
out.backward(dy.float())
d = nhwc(dy).to(BF)
# manual backward mirroring BasicBlock.backward
x_c2, pre_c2, y_c2, _ = blk.c2.ctx
d1, dres = blk.c2.backward(d, want_dres=True)
print("dres vs dy*mask", l2err(nchw(dres), o2.grad))
print("d1 (grad wrt o1) ", l2err(nchw(d1), o1.grad))
x_c1, pre_c1, y_c1, _ = blk.c1.ctx
dx_only = blk.c1.backward(d1)
print("dx main-path only", l2err(nchw(dx_only), xr.grad - o2.grad))
print("dx total", l2err(nchw(dx_only).float() + nchw(dres).float(), xr.grad))
# dpre2 check
print("norms", xr.grad.norm().item(), o2.grad.norm().item(), (xr.grad - o2.grad).norm().item())
