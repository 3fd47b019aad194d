"""Generate tests/golden/f32_bank_grad.npz: GRADIENTS of the reference's tensor bank, from torch autograd through the
reference's own code (data/functional.py with a stub `kornia`; the 13 ops that do not touch Kornia) composed exactly as
data/operations.py:73-100 composes them in training mode:

    out = clamp(mask * fn(x, mag) + (1 - mask) * x, 0, 1);   loss = sum(out * G)

for a fixed random G; stored: d loss / d x, d loss / d mag [B], d loss / d mask [B].  Needs /root/reference; run in the
build container only.  Inputs are the images of tests/golden/f32_bank.npz.
"""
import os
import sys
import types

import numpy as np
import torch

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
stub = types.ModuleType("kornia")
for name in ("rgb_to_hsv", "hsv_to_rgb", "shear", "translate", "rotate"):
    setattr(stub, name, lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("kornia is not installed")))
sys.modules["kornia"] = stub

import data.functional as RF  # noqa: E402  (reference)

OUT = os.path.join(ROOT, "tests", "golden", "f32_bank_grad.npz")
WITH_MAG = ("solarize", "posterize", "contrast", "saturate", "brightness", "sharpness", "sample_pairing")
NO_MAG = ("hflip", "vflip", "invert", "gray", "auto_contrast", "equalize")


def main():
    g = np.load(os.path.join(ROOT, "tests", "golden", "f32_bank.npz"))
    x0 = torch.from_numpy(g["imgs_u8"]).permute(0, 3, 1, 2).float() / 255
    b = x0.shape[0]
    rng = np.random.RandomState(3)
    G = torch.from_numpy(rng.randn(*x0.shape).astype(np.float32))
    mags0 = torch.tensor([0.2, 0.3, 0.65, 0.9])
    masks0 = torch.tensor([0.05, 0.5, 0.93, 1.0])           # RelaxedBernoulli samples live in (0, 1]
    out = {"G": G.numpy(), "mags": mags0.numpy(), "masks": masks0.numpy(), "pairing_perm": g["pairing_perm"]}
    for name in NO_MAG + WITH_MAG:
        x = x0.clone().requires_grad_(True)
        mag = mags0.clone().requires_grad_(True)
        mask = masks0.clone().requires_grad_(True)
        if name == "sample_pairing":
            torch.manual_seed(7)                              # the permutation of f32_bank.npz (pairing_perm)
        y = getattr(RF, name)(x, mag) if name in WITH_MAG else getattr(RF, name)(x)
        m4 = mask.view(b, 1, 1, 1)
        o = (m4 * y + (1 - m4) * x).clamp(0, 1)
        (o * G).sum().backward()
        out["out_" + name] = o.detach().numpy()
        out["gx_" + name] = x.grad.numpy()
        out["gmask_" + name] = mask.grad.numpy()
        if name in WITH_MAG:
            out["gmag_" + name] = (mag.grad if mag.grad is not None else torch.zeros(b)).numpy()
    np.savez_compressed(OUT, torch_version=torch.__version__, **out)
    print("wrote", OUT, len(out), "arrays")


if __name__ == "__main__":
    main()
