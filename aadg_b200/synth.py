"""Seeded synthetic inputs: fundus / vessel images, masks, feature clouds.

Pure numpy so that tests, bench.py and the oracle see byte-identical inputs on any box.
The reference ships no data and no seeds (run.py:48 only defaults --seed 1023), so these
generators define the workload for every parity test and benchmark line.
"""
import numpy as np

DEFAULT_SEED = 1023  # reference run.py:48


def fundus_batch(n, height, width, seed=DEFAULT_SEED):
    """uint8 images [n,H,W,3] and masks [n,H,W] with the optic convention used by
    data/transform.py:156-163 of the reference: 255 = background, 128 = disc ring, 0 = cup."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    imgs = np.empty((n, height, width, 3), np.uint8)
    masks = np.empty((n, height, width), np.uint8)
    for i in range(n):
        cy = height * (0.5 + 0.12 * (rng.rand() - 0.5))
        cx = width * (0.5 + 0.12 * (rng.rand() - 0.5))
        ry = height * (0.20 + 0.06 * rng.rand())
        rx = width * (0.20 + 0.06 * rng.rand())
        cup = 0.35 + 0.3 * rng.rand()
        d = ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2
        vign = 1.0 - 0.55 * (((yy - height / 2) / height) ** 2 + ((xx - width / 2) / width) ** 2) * 4
        base = np.stack([150 + 40 * rng.rand(), 70 + 30 * rng.rand(), 30 + 20 * rng.rand()]).astype(np.float32)
        img = vign[..., None] * base[None, None, :]
        disc = np.exp(-np.maximum(d - 0.6, 0) * 2.5)
        img += disc[..., None] * np.array([70, 90, 60], np.float32)
        img += (d < cup * cup)[..., None] * np.array([25, 40, 45], np.float32)
        img += rng.randint(-8, 9, size=img.shape)
        imgs[i] = np.clip(img, 0, 255).astype(np.uint8)
        m = np.full((height, width), 255, np.uint8)
        m[d < 1.0] = 128
        m[d < cup * cup] = 0
        masks[i] = m
    return imgs, masks


def vessel_batch(n, height, width, seed=DEFAULT_SEED):
    """uint8 images [n,H,W,3] and binary masks {0,255} [n,H,W] of thin random-walk vessels
    (RVS convention, data/transform.py:169-171: mask != 0 -> 1)."""
    rng = np.random.RandomState(seed + 7)
    imgs = np.empty((n, height, width, 3), np.uint8)
    masks = np.zeros((n, height, width), np.uint8)
    for i in range(n):
        img = np.empty((height, width, 3), np.float32)
        img[:] = np.array([160, 80, 40], np.float32) + 20 * (rng.rand(3) - 0.5)
        for _ in range(12):
            y, x = rng.randint(height), rng.randint(width)
            ang = rng.rand() * 2 * np.pi
            for _ in range(max(height, width)):
                ang += 0.25 * (rng.rand() - 0.5)
                y += np.sin(ang)
                x += np.cos(ang)
                iy, ix = int(y), int(x)
                if not (1 <= iy < height - 1 and 1 <= ix < width - 1):
                    break
                masks[i, iy - 1:iy + 1, ix - 1:ix + 1] = 255
        img[masks[i] > 0] *= 0.55
        img += rng.randint(-6, 7, size=img.shape)
        imgs[i] = np.clip(img, 0, 255).astype(np.uint8)
    return imgs, masks


def random_policies(m=6, q=5, l=2, num_ops=10, num_mags=10, seed=DEFAULT_SEED):
    """int64 [M, Q*L*2] laid out (op, mag) pairs like Controller.sample (models/controller.py:110)."""
    rng = np.random.RandomState(seed)
    pol = np.empty((m, q * l * 2), np.int64)
    pol[:, 0::2] = rng.randint(0, num_ops, (m, q * l))
    pol[:, 1::2] = rng.randint(0, num_mags, (m, q * l))
    return pol


def feature_cloud(n, d, domain, seed=DEFAULT_SEED):
    """float32 [n,d] cloud: leaky_relu(randn + 0.3*domain, 0.2) — mimics the 128-d hidden of
    MomentumFeatureDiscriminator (models/discriminator.py:28-29)."""
    rng = np.random.RandomState(seed + domain)
    x = rng.randn(n, d).astype(np.float32) + np.float32(0.3 * domain)
    return np.where(x > 0, x, np.float32(0.2) * x).astype(np.float32)
