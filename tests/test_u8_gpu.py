"""GPU: the CUDA uint8 bank (through the C ABI) against the oracle, the reference-generated golden
files and size-independent properties.  Bit-exact everywhere (integer / byte work)."""
import os
import random

import numpy as np
import pytest
import torch

from aadg_b200.data import decisions as D
from aadg_b200.data.basic import AADG_OPS, OP_ID
from aadg_b200.data.policy import parse_policies
from aadg_b200.synth import fundus_batch, vessel_batch, random_policies
from oracle import u8_bank as B
from oracle import u8_policy as P

pytestmark = pytest.mark.gpu


class Cfg:
    class CONTROLLER:
        EXCLUDE_OPS = []
        L = 2
        NUM_MAGS = 10
        EXCLUDE_OPS_NUM = 0
    SEED = 0


@pytest.fixture(scope="module")
def u8():
    assert torch.cuda.is_available(), "these tests need the B200"
    from aadg_b200.ops import u8 as mod
    return mod


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def make_row(src, ops, w, h, rng):
    """ops: [(name, level)] -> one decision row with random Cutout centres / mirrors."""
    row = np.zeros(1, D.ROW_DTYPE)[0]
    row["src"] = src
    row["n_ops"] = len(ops)
    row["scale_w"], row["scale_h"] = w, h
    for k, (name, level) in enumerate(ops):
        D.resolve_op(row, k, name, level, w, h, ux=rng.uniform(w), uy=rng.uniform(h),
                     mirror=bool(rng.rand() > 0.5))
    return row


def test_single_ops_golden(u8, golden_dir):
    g = np.load(os.path.join(golden_dir, "u8_single_ops.npz"))
    imgs, masks = fundus_batch(2, 48, 64, seed=int(g["seed"]))
    imgs[1] = np.random.RandomState(int(g["noise_seed"])).randint(0, 256, imgs[1].shape).astype(np.uint8)
    rows, want = [], []
    for i in range(2):
        for o in range(10):
            name = AADG_OPS[o][0]
            for lv in range(10):
                row = np.zeros(1, D.ROW_DTYPE)[0]
                row["src"], row["n_ops"] = i, 1
                r = np.random.RandomState(1000 * i + 10 * o + lv)
                ux = uy = None
                if name == "Cutout" and D.level_to_value(name, lv / 9) > 0:
                    ux, uy = r.uniform(64), r.uniform(48)
                D.resolve_op(row, 0, name, lv / 9, 64, 48, ux, uy)
                rows.append(row)
                want.append(g["outs"][i, o, lv])
    out = u8.apply_policy(dev(imgs), dev(masks), np.stack(rows)).cpu().numpy()
    for k, (row, w) in enumerate(zip(rows, want)):
        assert np.array_equal(out[k], w), (AADG_OPS[int(row["op"][0])][0], k % 10)


def test_geometric_ops_golden(u8, golden_dir):
    g = np.load(os.path.join(golden_dir, "u8_geometric_ops.npz"))
    imgs, masks = fundus_batch(1, 40, 56, seed=int(g["seed"]))
    rows, want, wantm = [], [], []
    for o, name in enumerate(["ShearX", "ShearY", "TranslateX", "TranslateY", "Rotate"]):
        for lv in range(10):
            for mi, mirror in enumerate((False, True)):
                row = np.zeros(1, D.ROW_DTYPE)[0]
                row["n_ops"] = 1
                D.resolve_op(row, 0, name, lv / 9, 56, 40, mirror=mirror)
                rows.append(row)
                want.append(g["outs"][o, lv, mi])
                wantm.append(g["outm"][o, lv, mi])
    row = np.zeros(1, D.ROW_DTYPE)[0]
    row["n_ops"] = 1
    D.resolve_op(row, 0, "Flip", 0.0, 56, 40)
    rows.append(row)
    want.append(g["flip"])
    wantm.append(masks[0])
    out, outm = u8.apply_policy(dev(imgs), dev(masks), np.stack(rows), want_masks=True)
    out, outm = out.cpu().numpy(), outm.cpu().numpy()
    assert np.array_equal(out, np.stack(want))
    assert np.array_equal(outm, np.stack(wantm))


def replay_rows(g):
    seed, n_src = int(g["seed"]), int(g["n_src"])
    h, w, crop = int(g["height"]), int(g["width"]), int(g["crop"])
    parsed = parse_policies(g["policies"], Cfg)
    py, npr = random.Random(seed), np.random.RandomState(seed)
    state = D.PolicyState(len(parsed))
    rows = []
    for s in range(n_src):
        r, raw = D.replay_sample(parsed, s, w, h, crop, tuple(g["scale_range"]), py, npr, state)
        D.soft_label(py, s % 3, 3)
        rows.append(r)
    return np.concatenate(rows)


@pytest.mark.parametrize("tag", ["optic64", "optic_rect", "rvs64"])
def test_post_policy_golden(u8, golden_dir, tag):
    """decision replay + CUDA bank == the reference's DGMultiPolicy output under the same seeds."""
    g = np.load(os.path.join(golden_dir, "u8_pipeline_%s.npz" % tag))
    gen = vessel_batch if bool(g["vessel"]) else fundus_batch
    imgs, masks = gen(int(g["n_src"]), int(g["height"]), int(g["width"]), seed=int(g["seed"]))
    rows = replay_rows(g)
    out = u8.apply_policy(dev(imgs), dev(masks), rows).cpu().numpy()
    assert np.array_equal(out.reshape(g["post_policy"].shape), g["post_policy"])


def replay_all(g):
    seed, n_src = int(g["seed"]), int(g["n_src"])
    h, w, crop = int(g["height"]), int(g["width"]), int(g["crop"])
    parsed = parse_policies(g["policies"], Cfg)
    py, npr = random.Random(seed), np.random.RandomState(seed)
    state = D.PolicyState(len(parsed))
    rows, raws = [], []
    for s in range(n_src):
        r, raw = D.replay_sample(parsed, s, w, h, crop, tuple(g["scale_range"]), py, npr, state)
        D.soft_label(py, s % 3, 3)
        rows.append(r)
        raws.append(raw)
    return np.concatenate(rows), np.stack(raws)


@pytest.mark.parametrize("tag", ["optic64", "optic_rect", "rvs64"])
def test_full_train_transform_golden(u8, golden_dir, tag):
    """policy -> DGRandomScaleCrop -> Normalize_dg -> ToTensor == the reference pipeline, bit for bit
    (Pillow BILINEAR/NEAREST resize, padding, crop), for the augmented copies and the raw image."""
    g = np.load(os.path.join(golden_dir, "u8_pipeline_%s.npz" % tag))
    gen = vessel_batch if bool(g["vessel"]) else fundus_batch
    imgs, masks = gen(int(g["n_src"]), int(g["height"]), int(g["width"]), seed=int(g["seed"]))
    rows, raws = replay_all(g)
    crop, ds = int(g["crop"]), str(g["dataset"])
    im, lb = u8.policy_scale_crop_normalize(dev(imgs), dev(masks), rows, crop, ds)
    want = (g["aug_u8"].astype(np.float32) / np.float32(127.5) - np.float32(1)).transpose(0, 1, 4, 2, 3)
    assert np.array_equal(im.cpu().numpy().reshape(want.shape), want)
    assert np.array_equal(lb.cpu().numpy().reshape(g["aug_labels"].shape), g["aug_labels"].astype(np.float32))
    im, lb = u8.scale_crop_normalize(dev(imgs), dev(masks), raws, crop, ds, image_by_row=False)
    wraw = (g["raw_u8"].astype(np.float32) / np.float32(127.5) - np.float32(1)).transpose(0, 3, 1, 2)
    assert np.array_equal(im.cpu().numpy(), wraw)
    assert np.array_equal(lb.cpu().numpy(), g["raw_labels"].astype(np.float32))


def test_scale_crop_vs_oracle_up_and_down(u8):
    """down-scaling (antialias window > 3 taps), padding and every crop offset path vs the oracle."""
    from oracle import u8_transform as T
    h, w, crop = 60, 72, 48
    imgs, masks = vessel_batch(2, h, w, seed=4)
    rng = np.random.RandomState(3)
    rows = np.zeros(12, D.ROW_DTYPE)
    for i, row in enumerate(rows):
        row["src"] = i % 2
        row["do_scale"] = int(i % 4 != 0)
        sw, sh = (int(rng.uniform(0.5, 2) * w), int(rng.uniform(0.5, 2) * h)) if row["do_scale"] else (w, h)
        row["scale_w"], row["scale_h"] = sw, sh
        pad = D.crop_padding(sw, sh, crop, crop)
        row["pad"] = pad
        row["crop_x"] = rng.randint(0, sw + 2 * pad - crop + 1)
        row["crop_y"] = rng.randint(0, sh + 2 * pad - crop + 1)
    im, lb = u8.scale_crop_normalize(dev(imgs), dev(masks), rows, crop, "vessel", image_by_row=False)
    for i, row in enumerate(rows):
        dec = dict(do_scale=int(row["do_scale"]), w=int(row["scale_w"]), h=int(row["scale_h"]),
                   x1=int(row["crop_x"]), y1=int(row["crop_y"]))
        wi, wm = T.scale_crop(imgs[row["src"]], masks[row["src"]], dec, crop, crop)
        fi, fm = T.to_tensor_pair(T.normalize_image(wi), T.normalize_mask(wm, "vessel"))
        assert np.array_equal(im[i].cpu().numpy(), fi), i
        assert np.array_equal(lb[i].cpu().numpy(), fm), i


CHAINS = [
    [("Sharpness", 1.0), ("Sharpness", 0.0)],
    [("Sharpness", 7 / 9), ("Equalize", 0.0)],
    [("Color", 2 / 9), ("Contrast", 1.0)],
    [("Invert", 0.0), ("AutoContrast", 0.0)],
    [("Posterize", 0.0), ("Equalize", 0.0), ("Contrast", 0.0)],
    [("Cutout", 1.0), ("Sharpness", 1.0), ("Cutout", 5 / 9), ("AutoContrast", 0.0)],
    [("Rotate", 1.0), ("Sharpness", 0.0), ("Color", 1.0)],
    [("ShearX", 0.0), ("TranslateY", 1.0), ("Equalize", 0.0)],
    [("Sharpness", 1.0), ("Rotate", 0.0), ("Sharpness", 1 / 9), ("Flip", 0.0)],
    [("Flip", 0.0), ("Cutout", 1.0), ("ShearY", 8 / 9), ("Contrast", 8 / 9)],
    [("Brightness", 1.0), ("Solarize", 4 / 9), ("Equalize", 0.0), ("AutoContrast", 0.0)],
    [("Equalize", 0.0), ("Equalize", 0.0)],
    [],
]


@pytest.mark.parametrize("size", [(37, 53), (64, 64), (130, 272)])
def test_chains_vs_oracle(u8, size):
    h, w = size
    imgs, masks = fundus_batch(3, h, w, seed=5)
    imgs[2] = np.random.RandomState(9).randint(0, 256, imgs[2].shape).astype(np.uint8)
    rng = np.random.RandomState(17)
    chains = list(CHAINS)
    names = [n for n, _, _ in AADG_OPS]
    for _ in range(40):
        n = rng.randint(1, 5)
        chains.append([(names[rng.randint(0, 16)], rng.randint(0, 10) / 9) for _ in range(n)])
    rows = np.stack([make_row(i % 3, c, w, h, rng) for i, c in enumerate(chains)])
    out, outm = u8.apply_policy(dev(imgs), dev(masks), rows, want_masks=True)
    out, outm = out.cpu().numpy(), outm.cpu().numpy()
    for i, c in enumerate(chains):
        wi, wm = P.apply_chain(imgs[i % 3], masks[i % 3], rows[i])
        assert np.array_equal(out[i], wi), (i, c, int(np.abs(out[i].astype(int) - wi).max()))
        assert np.array_equal(outm[i], wm), (i, c)


@pytest.mark.parametrize("dataset", ["optic", "vessel"])
def test_policy_normalize_vs_oracle(u8, dataset):
    h, w = 48, 64
    gen = fundus_batch if dataset == "optic" else vessel_batch
    imgs, masks = gen(4, h, w, seed=3)
    parsed = parse_policies(random_policies(seed=8), Cfg)
    rows, _ = D.philox_rows(parsed, 4, w, h, w, (1, 1.5), seed=1, scale_crop=False)
    im, lb = u8.policy_normalize(dev(imgs), dev(masks), rows, dataset=dataset)
    want = P.apply_rows(imgs, masks, rows, crop=None, dataset=dataset)
    assert np.array_equal(im.cpu().numpy(), want["images"])
    assert np.array_equal(lb.cpu().numpy(), want["labels"])


def test_full_size_properties(u8):
    """512x512 (BASELINE config 2 size): identities that hold for any image."""
    h = w = 512
    imgs, masks = fundus_batch(6, h, w, seed=1023)
    d_imgs, d_masks = dev(imgs), dev(masks)
    rng = np.random.RandomState(0)

    def run(chains):
        rows = np.stack([make_row(i % 6, c, w, h, rng) for i, c in enumerate(chains)])
        return u8.apply_policy(d_imgs, d_masks, rows), rows

    out, _ = run([[("Invert", 0), ("Invert", 0)]] * 6)
    assert torch.equal(out, d_imgs)
    out, _ = run([[("Flip", 0), ("Flip", 0)]] * 6)
    assert torch.equal(out, d_imgs)
    once, _ = run([[("AutoContrast", 0)]] * 6)
    twice, _ = run([[("AutoContrast", 0), ("AutoContrast", 0)]] * 6)
    assert torch.equal(once, twice)                       # idempotent: second pass has lo=0, hi=255
    once, _ = run([[("Posterize", 0.5)]] * 6)
    twice, _ = run([[("Posterize", 0.5), ("Posterize", 0.5)]] * 6)
    assert torch.equal(once, twice)
    # Brightness/Contrast/Color/Sharpness with factor 1.0 (level 4.5/9) are the identity
    out, _ = run([[("Brightness", 0.5), ("Contrast", 0.5), ("Color", 0.5), ("Sharpness", 0.5)]] * 6)
    assert torch.equal(out, d_imgs)
    # equalize flattens the histogram: every channel's CDF is within one step of the diagonal
    eq, _ = run([[("Equalize", 0)]] * 6)
    want = np.stack([B.equalize(im) for im in imgs])
    assert np.array_equal(eq.cpu().numpy(), want)
    # normalised float output == (uint8 output)/127.5 - 1 and labels follow the original masks
    chains = [[("Sharpness", 1.0), ("Equalize", 0)], [("Color", 0.0), ("Cutout", 1.0)]] * 3
    o8, rows = run(chains)
    f32, lab = u8.policy_normalize(d_imgs, d_masks, rows, dataset="optic")
    # (numpy on the host: torch's CUDA division by a scalar multiplies by the reciprocal)
    ref = (o8.cpu().numpy().astype(np.float32) / np.float32(127.5) - np.float32(1)).transpose(0, 3, 1, 2)
    assert np.array_equal(f32.cpu().numpy(), ref)
    m = d_masks[torch.arange(6, device="cuda") % 6]
    assert torch.equal(lab[:, 1], (m <= 200).float())
    assert torch.equal(lab[:, 0], (m <= 50).float())
    assert o8.shape == (6, 512, 512, 3)


def test_empty_and_errors(u8):
    imgs, masks = fundus_batch(1, 16, 16, seed=2)
    out = u8.apply_policy(dev(imgs), dev(masks), np.zeros(0, D.ROW_DTYPE))
    assert out.shape == (0, 16, 16, 3)
    bad = np.zeros(1, D.ROW_DTYPE)
    bad["src"] = 3
    with pytest.raises(RuntimeError, match="src"):
        u8.apply_policy(dev(imgs), dev(masks), bad)
    bad["src"] = 0
    bad["n_ops"] = 1
    bad["op"][0, 0] = 99
    with pytest.raises(RuntimeError, match="op"):
        u8.apply_policy(dev(imgs), dev(masks), bad)
