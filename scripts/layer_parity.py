"""Layer-by-layer implementation parity on REAL activations (see tests/test_parity_gpu.py::layer_residuals): every
convolution / depthwise convolution / batch-norm of the engine is fed the bf16-storage oracle's own input for that layer
and compared with the oracle's output; with `fp64` each convolution is also compared with a float64 evaluation (tells
the oracle's own fp32 algorithm error from the engine's).  Then the free-running forward stage by stage, the engine's
run-to-run spread and an uninitialised-memory probe (allocator free blocks poisoned with NaN).

    python scripts/layer_parity.py [encoder] [size] [n] [arch]      (GPU)
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

BF = torch.bfloat16


def main():
    from test_parity_gpu import l2err, layer_residuals, nchw
    enc = sys.argv[1] if len(sys.argv) > 1 else "resnet18"
    size = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    arch = sys.argv[4] if len(sys.argv) > 4 else "deeplabv3plus"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    classes, dataset = (2, "optic") if arch == "deeplabv3plus" else (1, "vessel")
    rows, (ref, twin, net, x, target) = layer_residuals(arch, enc, classes, size, n, dataset, fp64=True)
    print("== teacher-forced per-layer residuals (%s/%s %d^2 n=%d), worst first ==" % (arch, enc, size, n))
    for r in rows[:30]:
        print("  %.3e  %-10s %-45s %s %s" % r)
    print("  ... %d layers, median %.3e" % (len(rows), rows[len(rows) // 2][0]))
    with torch.no_grad():
        feats_t = twin.encoder(x.to(BF).float())
        feats_r = ref.encoder(x)
    feats_e = net.encoder.forward(x.contiguous(), True)
    print("== free-running encoder features: engine vs bf16-storage oracle | bf16-storage oracle vs fp32 oracle ==")
    for i, f in enumerate(feats_e):
        print("  feat[%d] %-22s %.3e | %.3e" % (i + 1, tuple(f.shape), l2err(nchw(f), feats_t[i + 1]), l2err(feats_t[i + 1], feats_r[i + 1])))

    def run_once():
        net.store.zero_grad()
        out = net.loss_step(x, target, want_logits=True)
        return out["loss"].item(), out["pooled"].clone(), out["logits"].clone(), net.store.grads.clone()
    a = run_once()
    b = run_once()
    print("== run-to-run spread of two identical loss_step calls (fp32 atomics -> bf16 re-quantisation) ==")
    print("  loss %.9f vs %.9f (rel %.2e)  pooled L2 %.2e  logits L2 %.2e  grads L2 %.2e" %
          (a[0], b[0], abs(a[0] - b[0]) / abs(a[0]), l2err(b[1], a[1]), l2err(b[2], a[2]), l2err(b[3], a[3])))
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
