"""Generate tests/golden/u8_*.npz by running the REFERENCE itself (needs /root/reference; run in the
build container, never on the GPU box).  Inputs are the seeded synthetic images of
aadg_b200/synth.py, so only seeds + outputs are stored.

  PYTHONDONTWRITEBYTECODE=1 python scripts/make_golden_u8.py
"""
import os
import random
import sys

import numpy as np

sys.dont_write_bytecode = True
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

from PIL import Image  # noqa: E402
import PIL  # noqa: E402
import data.basic as rbasic  # noqa: E402  (reference)
import data.policy as rpolicy  # noqa: E402  (reference)
import data.transform as rtransform  # noqa: E402  (reference)

from aadg_b200.synth import fundus_batch, vessel_batch, random_policies  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class _Cfg:
    class CONTROLLER:
        EXCLUDE_OPS = []
        L = 2
        NUM_MAGS = 10
        EXCLUDE_OPS_NUM = 0
    SEED = 0


def single_ops():
    """every live op x every magnitude on two images; Cutout centres from np.random.seed(case)."""
    imgs, masks = fundus_batch(2, 48, 64, seed=11)
    rng = np.random.RandomState(3)
    imgs[1] = rng.randint(0, 256, imgs[1].shape).astype(np.uint8)
    outs = np.zeros((2, 10, 10) + imgs[0].shape, np.uint8)
    for i in range(2):
        for o, (fn, lo, hi) in enumerate(rbasic.augment_list()):
            for lv in range(10):
                np.random.seed(1000 * i + 10 * o + lv)
                r, _ = rbasic.apply_augment(Image.fromarray(imgs[i]), Image.fromarray(masks[i]),
                                            fn.__name__, lv / 9)
                outs[i, o, lv] = np.asarray(r)
    np.savez_compressed(os.path.join(OUT, "u8_single_ops.npz"), outs=outs, seed=11, noise_seed=3,
                        pillow=PIL.__version__)


def geometric_ops():
    imgs, masks = fundus_batch(1, 40, 56, seed=12)
    names = ["ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate"]
    outs = np.zeros((5, 10, 2) + imgs[0].shape, np.uint8)
    outm = np.zeros((5, 10, 2) + masks[0].shape, np.uint8)
    for o, name in enumerate(names):
        lo, hi = {"ShearX": (-.3, .3), "ShearY": (-.3, .3), "TranslateX": (-.45, .45),
                  "TranslateY": (-.45, .45), "Rotate": (-30, 30)}[name]
        for lv in range(10):
            for mi, mirror in enumerate((False, True)):
                rbasic.random.random = (lambda m=mirror: 0.9 if m else 0.1)
                r, rm = getattr(rbasic, name)(Image.fromarray(imgs[0]), Image.fromarray(masks[0]),
                                              (lv / 9) * (hi - lo) + lo)
                outs[o, lv, mi] = np.asarray(r)
                outm[o, lv, mi] = np.asarray(rm)
    import importlib
    importlib.reload(random)
    flip, _ = rbasic.Flip(Image.fromarray(imgs[0]), Image.fromarray(masks[0]), 0)
    np.savez_compressed(os.path.join(OUT, "u8_geometric_ops.npz"), outs=outs, outm=outm,
                        flip=np.asarray(flip), seed=12, pillow=PIL.__version__)


def pipeline(tag, dataset, height, width, crop, scale_range, seed, n_src, vessel=False):
    """DGMultiPolicy -> DGRandomScaleCrop -> Normalize_dg -> ToTensor on n_src samples, RNGs seeded
    once (random.seed / np.random.seed), processed sequentially like a num_workers=0 loader."""
    gen = vessel_batch if vessel else fundus_batch
    imgs, masks = gen(n_src, height, width, seed=seed)
    pol = random_policies(seed=seed)
    parsed = rpolicy.parse_policies(pol, _Cfg, None)
    multi = rpolicy.DGMultiPolicy(parsed)
    scale = rtransform.DGRandomScaleCrop(crop, scale_range=list(scale_range))
    norm = rtransform.Normalize_dg(dataset)
    tot = rtransform.ToTensor(dataset)
    random.seed(seed)
    np.random.seed(seed)
    post_policy, aug_u8, lab, dcs, raw_u8, raw_lab = [], [], [], [], [], []
    for s in range(n_src):
        sample = {"image": Image.fromarray(imgs[s]), "label": Image.fromarray(masks[s]),
                  "dc": s % 3, "img_name": str(s)}
        sample = multi(sample)
        post_policy.append(np.stack([np.asarray(a) for a in sample["aug_images"]]))
        sample = scale(sample)
        cropped = np.stack([np.asarray(a) for a in sample["aug_images"]])
        raw_crop = np.asarray(sample["image"])
        sample = tot(norm(sample))
        ai = sample["aug_images"].numpy()
        assert np.array_equal(ai, (cropped.astype(np.float32) / np.float32(127.5) - np.float32(1)).transpose(0, 3, 1, 2))
        assert np.array_equal(sample["image"].numpy(), (raw_crop.astype(np.float32) / np.float32(127.5) - np.float32(1)).transpose(2, 0, 1))
        aug_u8.append(cropped)
        lab.append(sample["aug_labels"].numpy().astype(np.uint8))
        dcs.append(sample["dc"].numpy())
        raw_u8.append(raw_crop)
        raw_lab.append(sample["label"].numpy().astype(np.uint8))
    np.savez_compressed(
        os.path.join(OUT, "u8_pipeline_%s.npz" % tag), policies=pol, seed=seed, n_src=n_src,
        height=height, width=width, crop=crop, scale_range=np.asarray(scale_range, np.float64),
        dataset=dataset, vessel=vessel, post_policy=np.stack(post_policy),
        aug_u8=np.stack(aug_u8), aug_labels=np.stack(lab), dc=np.stack(dcs),
        raw_u8=np.stack(raw_u8), raw_labels=np.stack(raw_lab), pillow=PIL.__version__)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    single_ops()
    geometric_ops()
    pipeline("optic64", "optic", 64, 64, 64, (1, 1.5), 1023, 13)
    pipeline("optic_rect", "optic", 72, 56, 48, (1, 1.5), 77, 5)
    pipeline("rvs64", "vessel", 64, 64, 48, (0.5, 2), 4242, 5, vessel=True)
    print("golden files written to", OUT)
