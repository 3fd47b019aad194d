"""uint8 augmentation bank: torch-facing wrappers over aadg_u8_* (csrc/aug_u8.cu).

Inputs are CUDA uint8 tensors `[S,H,W,3]` (masks `[S,H,W]`) and a decision table (numpy structured
array, aadg_b200/data/decisions.py ROW_DTYPE == aadg_aug_row_t).  There is no CPU path."""
import numpy as np
import torch

from .. import _lib
from ..data.decisions import ROW_DTYPE

DATASETS = {"optic": 0, "vessel": 1, "rvs": 1}


def _prep(src_images, src_masks, rows):
    if not (isinstance(src_images, torch.Tensor) and src_images.is_cuda):
        raise RuntimeError("aadg_b200.ops.u8: src_images must be a CUDA uint8 tensor (no CPU path)")
    if src_images.dtype != torch.uint8 or src_images.dim() != 4 or src_images.shape[-1] != 3:
        raise ValueError("src_images must be uint8 [S,H,W,3]")
    src_images = src_images.contiguous()
    if src_masks is not None:
        if src_masks.dtype != torch.uint8 or tuple(src_masks.shape) != tuple(src_images.shape[:3]):
            raise ValueError("src_masks must be uint8 [S,H,W]")
        src_masks = src_masks.contiguous()
    rows = np.ascontiguousarray(rows, dtype=ROW_DTYPE)
    return src_images, src_masks, rows


def apply_policy(src_images, src_masks, rows, want_masks=False):
    """Post-policy uint8 images [n_rows,H,W,3] (and edited masks) — DGMultiPolicy's 'aug_images'."""
    src_images, src_masks, rows = _prep(src_images, src_masks, rows)
    s, h, w, _ = src_images.shape
    n = len(rows)
    dev = src_images.device
    out = torch.empty((n, h, w, 3), dtype=torch.uint8, device=dev)
    outm = torch.empty((n, h, w), dtype=torch.uint8, device=dev) if want_masks else None
    L = _lib.lib()
    ws = _lib.workspace(L.aadg_u8_workspace_bytes(max(n, 1), s, h, w), dev)
    with torch.cuda.device(dev):
        _lib.check(L.aadg_u8_apply_policy(_lib.ptr(src_images), _lib.ptr(src_masks), _lib.ptr(rows), n, s,
                                          h, w, _lib.ptr(out), _lib.ptr(outm), _lib.ptr(ws), ws.numel(),
                                          _lib.stream_ptr()))
    return (out, outm) if want_masks else out


def policy_normalize(src_images, src_masks, rows, dataset="optic", want_images=True, want_labels=True,
                     out_images=None, out_labels=None):
    """float32 [n_rows,3,H,W] in [-1,1] and labels float32 [n_rows,C,H,W] in one pass."""
    src_images, src_masks, rows = _prep(src_images, src_masks, rows)
    s, h, w, _ = src_images.shape
    n = len(rows)
    dev = src_images.device
    ds = DATASETS[dataset]
    c = 2 if ds == 0 else 1
    if want_images and out_images is None:
        out_images = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev)
    if want_labels and out_labels is None:
        if src_masks is None:
            raise ValueError("labels need src_masks")
        out_labels = torch.empty((n, c, h, w), dtype=torch.float32, device=dev)
    L = _lib.lib()
    ws = _lib.workspace(L.aadg_u8_workspace_bytes(max(n, 1), s, h, w), dev)
    with torch.cuda.device(dev):
        _lib.check(L.aadg_u8_policy_normalize(_lib.ptr(src_images), _lib.ptr(src_masks), _lib.ptr(rows), n,
                                              s, h, w, ds, _lib.ptr(out_images if want_images else None),
                                              _lib.ptr(out_labels if want_labels else None), _lib.ptr(ws),
                                              ws.numel(), _lib.stream_ptr()))
    return out_images, out_labels


def normalize_to_tensor(src_images, src_masks, dataset="optic"):
    """The test-time transform (Normalize_dg + ToTensor, data/transform.py:138-236 without any policy): images
    float32 [S,3,H,W] in [-1,1] and labels float32 [S,C,H,W], one row per source image with zero operations."""
    from ..data.decisions import ROW_DTYPE
    rows = np.zeros(src_images.shape[0], ROW_DTYPE)
    rows["src"] = np.arange(src_images.shape[0])
    return policy_normalize(src_images, src_masks, rows, dataset=dataset)


def scale_crop_normalize(images, masks, rows, crop, dataset="optic", image_by_row=True, want_labels=True,
                         out_images=None, out_labels=None):
    """DGRandomScaleCrop + Normalize_dg + ToTensor for every row (data/transform.py:97-236): images uint8
    [n,H,W,3] (post-policy, one per row) or the sources (image_by_row=False); masks = ORIGINAL masks [S,H,W].
    Returns (float32 [n,3,crop,crop], float32 [n,C,crop,crop] or None)."""
    images, masks, rows = _prep(images, masks, rows) if masks is None or masks.shape[0] == images.shape[0] \
        else (images.contiguous(), masks.contiguous(), np.ascontiguousarray(rows, dtype=ROW_DTYPE))
    _, h, w, _ = images.shape
    n = len(rows)
    dev = images.device
    ds = DATASETS[dataset]
    c = 2 if ds == 0 else 1
    n_src = masks.shape[0] if masks is not None else images.shape[0]
    out_i = out_images if out_images is not None else torch.empty((n, 3, crop, crop), dtype=torch.float32, device=dev)
    assert out_i.shape == (n, 3, crop, crop) and out_i.dtype == torch.float32 and out_i.is_contiguous()
    out_l = None
    if want_labels and masks is not None:
        out_l = out_labels if out_labels is not None else torch.empty((n, c, crop, crop), dtype=torch.float32, device=dev)
        assert out_l.shape == (n, c, crop, crop) and out_l.dtype == torch.float32 and out_l.is_contiguous()
    mw = int(max([w] + [int(r["scale_w"]) for r in rows if r["do_scale"]]))
    mh = int(max([h] + [int(r["scale_h"]) for r in rows if r["do_scale"]]))
    L = _lib.lib()
    ws = _lib.workspace(L.aadg_u8_scale_crop_workspace_bytes(max(n, 1), mw, mh), dev)
    with torch.cuda.device(dev):
        _lib.check(L.aadg_u8_scale_crop_normalize(_lib.ptr(images), int(image_by_row), _lib.ptr(masks), _lib.ptr(rows), n,
                                                  n_src, h, w, crop, crop, ds, _lib.ptr(out_i), _lib.ptr(out_l),
                                                  _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return out_i, out_l


def policy_scale_crop_normalize(src_images, src_masks, rows, crop, dataset="optic", out_images=None, out_labels=None):
    """The reference's whole train transform for the augmented copies: DGMultiPolicy -> DGRandomScaleCrop ->
    Normalize_dg -> ToTensor -> collate (data/policy.py:51-61, data/transform.py:97-236,323-340)."""
    post = apply_policy(src_images, src_masks, rows)
    return scale_crop_normalize(post, src_masks, rows, crop, dataset, image_by_row=True, out_images=out_images,
                                out_labels=out_labels)


def apply_dg_multipolicy(policy, sample):
    """DGMultiPolicy.__call__ on a batch sample (reference data/policy.py:51-61): adds
    'aug_images' uint8 [S*M,H,W,3] and 'aug_labels' uint8 [S*M,H,W], row index s*M + j."""
    imgs, masks = sample["image"], sample["label"]
    rows, raws = policy.rows_for(imgs.shape[0], imgs.shape[2], imgs.shape[1])
    out, outm = apply_policy(imgs, masks, rows, want_masks=True)
    sample = dict(sample)
    sample["aug_images"], sample["aug_labels"] = out, outm
    sample["aug_rows"], sample["raw_rows"] = rows, raws
    return sample
