"""Generate tests/golden/f32_bank.npz by running the REFERENCE's tensor bank (data/functional.py,
data/operations.py — dead code there, named by the north star) with a stub `kornia` module: the 13 ops that
do not touch Kornia run unmodified.  Needs /root/reference; run in the build container only.

Inputs are uint8-quantised images / 255 (what the bank would see after ToTensor), stored as the uint8 array.
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
import data.operations as RO  # noqa: E402  (reference)
from aadg_b200.synth import fundus_batch  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "f32_bank.npz")


def main():
    imgs, _ = fundus_batch(4, 40, 56, seed=21)
    imgs[3] = np.random.RandomState(5).randint(0, 256, imgs[3].shape).astype(np.uint8)
    x = torch.from_numpy(imgs).permute(0, 3, 1, 2).float() / 255
    mags = torch.tensor([0.0, 0.3, 0.65, 1.0])
    out = {"imgs_u8": imgs, "mags": mags.numpy()}
    for name in ("hflip", "vflip", "invert", "gray", "auto_contrast", "equalize"):
        out["fn_" + name] = getattr(RF, name)(x.clone()).numpy()
    for name in ("solarize", "posterize", "contrast", "saturate", "brightness", "sharpness"):
        out["fn_" + name] = getattr(RF, name)(x.clone(), mags.clone()).numpy()
    # sample_pairing draws a permutation from the global torch RNG
    torch.manual_seed(7)
    out["pairing_perm"] = torch.randperm(4).numpy()
    torch.manual_seed(7)
    out["fn_sample_pairing"] = RF.sample_pairing(x.clone(), mags.clone()).numpy()
    # _Operation.forward, training mode (soft mask) and eval mode (hard mask), masks recorded
    for cls, mag in (("Solarize", 0.4), ("Sharpness", 0.8), ("Invert", None)):
        op = getattr(RO, cls)() if mag is None else getattr(RO, cls)(initial_magnitude=mag)
        op.train()
        torch.manual_seed(11)
        mask = op.get_mask(4)
        sign = torch.randint(2, (4,), dtype=torch.float32).mul_(2).sub_(1) if op.flip_magnitude else torch.ones(4)
        out["op_train_sign_" + cls] = sign.numpy()
        torch.manual_seed(11)
        out["op_train_" + cls] = op(x.clone()).detach().numpy()
        out["op_train_mask_" + cls] = mask.detach().numpy().reshape(4)
        op.eval()
        torch.manual_seed(13)
        mask = op.get_mask(4)
        sign = torch.randint(2, (4,), dtype=torch.float32).mul_(2).sub_(1) if op.flip_magnitude else torch.ones(4)
        out["op_eval_sign_" + cls] = sign.numpy()
        torch.manual_seed(13)
        out["op_eval_" + cls] = op(x.clone()).detach().numpy()
        out["op_eval_mask_" + cls] = mask.numpy().reshape(4)
        out["op_mag_" + cls] = np.float32(0.0 if mag is None else float(op.magnitude))
    np.savez_compressed(OUT, torch_version=torch.__version__, **out)
    print("wrote", OUT, sorted(out))


if __name__ == "__main__":
    main()
