import sys, torch
import torch.nn.functional as F
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from test_nn_gpu import _pair
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for enc, size, n in [("resnet18", 128, 4), ("resnet50", 128, 4), ("resnet50", 256, 6)]:
    ref, net, x, target = _pair(enc, 2, size, n)
    masks, pooled = ref(x)
    loss = F.binary_cross_entropy(torch.sigmoid(masks), target)
    loss.backward()
    net.store.zero_grad()
    out = net.loss_step(x, target, want_logits=True)
    l2 = lambda a, b: ((a.float() - b.float()).norm() / b.float().norm()).item()
    print(enc, size, n, "pooled relL2", l2(out["pooled"], pooled), "logits relL2", l2(out["logits"], masks),
          "loss", out["loss"].item(), loss.item())
    sd_grads = {k: v.grad for k, v in ref.named_parameters()}
    worst = []
    for name, p in net.named_params().items():
        g = p.grad.detach()
        if name == "encoder.conv1.weight":
            g = __import__('aadg_b200.nn.network', fromlist=['stem_unpack']).stem_unpack(g)
        elif p.kind in ("conv", "conv_nt"):
            k = int(round(p.shape[0] ** 0.5))
            g = g.reshape(k, k, p.shape[1], p.shape[2]).permute(2, 3, 0, 1)
        elif len(p.shape) == 2 and p.shape[0] == 9:
            g = g.t().reshape(p.shape[1], 1, 3, 3)
        w = sd_grads[name].reshape(-1).double(); g = g.reshape(-1).double()
        cos = (g @ w / (g.norm() * w.norm() + 1e-30)).item()
        worst.append((cos, (g.norm() / (w.norm() + 1e-30)).item(), name, w.norm().item()))
    worst.sort()
    for t in worst[:8]: print("   ", t)
    import statistics
    print("    median cos", statistics.median([w[0] for w in worst]))
