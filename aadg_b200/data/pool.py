"""Resident dataset pools: every source-domain image and mask lives in device memory as uint8 and a training step
gathers its B x D source images with one indexed copy -- no DataLoader workers, no pinned-memory staging, no
per-step host-to-device transfer of pixels.

Mirrors the reference's in-memory pools and their sampling semantics (data/optic.py:20-103, data/vessel.py:126-156,
data/dataloader.py:27-30): `len(pool)` is the largest domain's image count, an item ignores its index and draws ONE
image per source domain with `np.random.choice(len(domain), 1)[0]` (optic.py:81-88), a batch is B items, the epoch
has len // B steps (drop_last=True), and rows are ordered item-major, domain-minor (b*D + d) like
train_dg_collate_fn (data/transform.py:323-340).  File decoding / resizing to the pool resolution happens once,
outside the hot path (the caller hands in arrays)."""
import numpy as np
import torch


class ResidentPools:
    def __init__(self, images_by_domain, masks_by_domain, device="cuda"):
        """images_by_domain: {domain name: uint8 [n, H, W, 3]}, masks_by_domain: {domain name: uint8 [n, H, W]};
        domain order = insertion order (the reference's `list(self.image_pool.keys()).index(key)` domain code)."""
        self.names = list(images_by_domain)
        if not self.names or list(masks_by_domain) != self.names:
            raise ValueError("images and masks must list the same, non-empty set of domains in the same order")
        self.sizes = [int(len(images_by_domain[k])) for k in self.names]
        if min(self.sizes) < 1:
            raise ValueError("every domain needs at least one image")
        shapes = {tuple(np.asarray(images_by_domain[k]).shape[1:]) for k in self.names}
        if len(shapes) != 1 or len(next(iter(shapes))) != 3 or next(iter(shapes))[2] != 3:
            raise ValueError("all pools must hold uint8 [n, H, W, 3] images of one resolution")
        self.offsets = np.concatenate([[0], np.cumsum(self.sizes)]).astype(np.int64)
        self.device = torch.device(device)
        imgs = np.concatenate([np.asarray(images_by_domain[k], dtype=np.uint8) for k in self.names])
        msks = np.concatenate([np.asarray(masks_by_domain[k], dtype=np.uint8) for k in self.names])
        if msks.shape != imgs.shape[:3]:
            raise ValueError("masks must be uint8 [n, H, W] matching the images")
        self.images = torch.from_numpy(imgs).to(self.device)
        self.masks = torch.from_numpy(msks).to(self.device)

    @property
    def n_domains(self):
        return len(self.names)

    def __len__(self):
        return max(self.sizes)                       # data/optic.py:72-77

    def steps_per_epoch(self, batch_size):
        return len(self) // batch_size               # DataLoader(drop_last=True)

    def nbytes(self):
        return self.images.numel() + self.masks.numel()

    def sample_indices(self, batch_size, rng=np.random):
        """[B, D] pool-local indices drawn in the reference's order: for every item, for every domain,
        rng.choice(len(domain), 1)[0] (data/optic.py:81-84).  Pass np.random (after np.random.seed) to replay a
        num_workers=0 reference run exactly, or any RandomState."""
        out = np.empty((batch_size, self.n_domains), np.int64)
        for b in range(batch_size):
            for d, n in enumerate(self.sizes):
                out[b, d] = rng.choice(n, 1)[0]
        return out

    def flat_indices(self, indices):
        """indices [B, D] pool-local -> (flat [B*D] int64 indices into `self.images` / `self.masks`, domains list[B*D])
        in collate order b*D + d.  The zero-copy form of a batch: SearchEngine.step(pools.images, pools.masks, domains,
        src_index=flat) lets the augmentation kernels read the step's sources straight out of the resident pool."""
        idx = np.asarray(indices, np.int64)
        if idx.ndim != 2 or idx.shape[1] != self.n_domains:
            raise ValueError("indices must be [B, %d]" % self.n_domains)
        if (idx < 0).any() or (idx >= np.asarray(self.sizes)[None, :]).any():
            raise IndexError("pool index out of range")
        domains = [d for _ in range(idx.shape[0]) for d in range(self.n_domains)]
        return (idx + self.offsets[None, :-1]).reshape(-1), domains

    def gather(self, indices):
        """indices [B, D] pool-local -> (images uint8 [B*D, H, W, 3], masks uint8 [B*D, H, W], domains list[B*D]) in
        collate order b*D + d; one device gather per tensor."""
        idx = np.asarray(indices, np.int64)
        if idx.ndim != 2 or idx.shape[1] != self.n_domains:
            raise ValueError("indices must be [B, %d]" % self.n_domains)
        if (idx < 0).any() or (idx >= np.asarray(self.sizes)[None, :]).any():
            raise IndexError("pool index out of range")
        flat = torch.from_numpy((idx + self.offsets[None, :-1]).reshape(-1)).to(self.device, non_blocking=True)
        domains = [d for _ in range(idx.shape[0]) for d in range(self.n_domains)]
        return self.images.index_select(0, flat), self.masks.index_select(0, flat), domains

    def batch(self, batch_size, rng=np.random):
        return self.gather(self.sample_indices(batch_size, rng))

    def epoch(self, batch_size, rng=np.random):
        for _ in range(self.steps_per_epoch(batch_size)):
            yield self.batch(batch_size, rng)
