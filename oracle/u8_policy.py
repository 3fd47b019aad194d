"""ORACLE (test infrastructure, never on the product path).

Applies a decision table (rows of `aadg_aug_row_t`, see aadg_b200/data/decisions.py) with the NumPy
restatements of the reference's uint8 bank and train transform.  This is the CPU statement of what
`DGMultiPolicy -> DGRandomScaleCrop -> Normalize_dg -> ToTensor -> train_dg_collate_fn` produce
(reference data/policy.py:51-61, data/transform.py:114-236,323-340) once every random draw is fixed.

Pinned by tests/test_oracle_u8.py against tests/golden/u8_pipeline_*.npz (outputs of the reference
itself under seeded RNGs, scripts/make_golden_u8.py).
"""
import numpy as np

from . import u8_bank as B
from . import u8_transform as T


def parse_policies(policies, L=2, num_mags=10, exclude_ops=()):
    """data/policy.py:85-97 (the decode; exclusion by explicit names only)."""
    names = [n for n in B.OP_NAMES[:10] if n not in exclude_ops]
    policies = np.asarray(policies)
    out = []
    for p in policies:
        q = len(p) // (2 * L)
        out.append([[(names[p[2 * L * j + 2 * k]], p[2 * L * j + 2 * k + 1] / (num_mags - 1))
                     for k in range(L)] for j in range(q)])
    return out


def apply_chain(img, mask, row):
    """The L ops of the chosen sub-policy, sequentially (data/policy.py:24-28)."""
    for k in range(int(row["n_ops"])):
        name = B.OP_NAMES[int(row["op"][k])]
        ip = row["iparam"][k]
        f = row["fparam"][k]
        if name == "AutoContrast":
            img = B.autocontrast(img)
        elif name == "Invert":
            img = B.invert(img)
        elif name == "Equalize":
            img = B.equalize(img)
        elif name == "Solarize":
            img = B.apply_lut(img, B.lut_solarize(int(ip[0])))
        elif name == "Posterize":
            i = np.arange(256)
            img = B.apply_lut(img, np.tile((i & int(ip[0])).astype(np.uint8), (3, 1)))
        elif name == "Contrast":
            img = B.contrast(img, f)
        elif name == "Color":
            img = B.color(img, f)
        elif name == "Brightness":
            img = B.brightness(img, f)
        elif name == "Sharpness":
            img = B.sharpness(img, f)
        elif name == "Cutout":
            rect = tuple(int(v) for v in ip[:4])
            img = B.cutout(img, rect)
            m = mask.copy()
            x0, y0, x1, y1 = rect
            if x1 >= x0 and y1 >= y0:
                m[max(y0, 0):y1 + 1, max(x0, 0):x1 + 1] = 0
            mask = m
        elif name == "Flip":
            img = B.flip(img)
        else:
            fx = tuple(int(v) for v in ip)
            img, mask = B.affine_nearest(img, fx), B.affine_nearest(mask, fx)
    return img, mask


def apply_rows(src_imgs, src_masks, rows, crop=None, dataset="optic"):
    """Returns dict(aug_u8 [n,H,W,3], aug_mask_u8 [n,H,W], images f32 [n,3,th,tw],
    labels f32 [n,C,th,tw]).  The label of every copy comes from the ORIGINAL mask, scaled and
    cropped with the copy's own decisions (data/transform.py:127-131)."""
    aug, augm, ims, lbs = [], [], [], []
    for row in rows:
        s = int(row["src"])
        img, m = apply_chain(src_imgs[s], src_masks[s], row)
        aug.append(img)
        augm.append(m)
        mask = src_masks[s]
        if crop is not None:
            dec = dict(do_scale=int(row["do_scale"]), w=int(row["scale_w"]), h=int(row["scale_h"]),
                       x1=int(row["crop_x"]), y1=int(row["crop_y"]))
            img, mask = T.scale_crop(img, mask, dec, crop, crop)
        fi, fm = T.to_tensor_pair(T.normalize_image(img), T.normalize_mask(mask, dataset))
        ims.append(fi)
        lbs.append(fm)
    return dict(aug_u8=np.stack(aug), aug_mask_u8=np.stack(augm), images=np.stack(ims),
                labels=np.stack(lbs))
