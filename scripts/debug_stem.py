import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_nn_gpu import _pair, nchw, nhwc, l2err
from aadg_b200.ops import nn as K, conv as C
from aadg_b200.nn.network import STEM_KP
torch.backends.cudnn.allow_tf32 = False
BF = torch.bfloat16
for size, n in [(64, 3), (64, 4), (128, 3), (128, 8), (96, 2)]:
    ref, net, x, target = _pair("resnet18", 2, size, n)
    with torch.no_grad():
        want = ref.encoder.conv1(x)
    col = K.im2col_stem(x, 7, 7, 2, 3, STEM_KP)
    refcol = F.unfold(x, 7, padding=3, stride=2)
    ho = want.shape[2]
    refcol = refcol.view(n, 3, 49, ho, ho).permute(0, 3, 4, 2, 1).reshape(n, ho, ho, 147)
    print(size, n, "im2col equal", torch.equal(col[..., :147], refcol.to(BF)))
    pre = C.fprop(col, net.encoder.stem_w.bf16, 1, 1)
    print("   conv relL2", l2err(nchw(pre), want))
    w = net.encoder.stem_w.bf16[0].float()     # [64,192]
    manual = (col.float().reshape(-1, 192) @ w.t()).reshape(n, ho, ho, 64)
    print("   conv vs manual matmul", l2err(pre, manual), " manual vs torch", l2err(nchw(manual), want))
