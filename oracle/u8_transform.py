"""ORACLE (test infrastructure, never on the product path).

NumPy restatement of the stages that follow the policy in the reference's train transform
(`/root/reference/data/transform.py:97-236`): DGRandomScaleCrop (Pillow BILINEAR / NEAREST resize,
ImageOps.expand padding, crop), Normalize_dg and ToTensor.  All random decisions are passed in
explicitly (see oracle/decisions.py for the draw order).

Pillow's resize arithmetic (Resample.c, Geometry.c ImagingScaleAffine) is restated at the
integer level; pinned by tests/test_oracle_u8.py against the reference-generated golden files and
against Pillow itself.
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2  # Resample.c


def bilinear_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BILINEAR (triangle) filter.
    Returns (xmin[out], count[out], k[out, ksize]) with k int32 fixed point (22 bits)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        if lo < 0:
            lo = 0
        hi = int(center + support + 0.5)
        if hi > in_size:
            hi = in_size
        n = hi - lo
        w = np.zeros(n, np.float64)
        ww = 0.0
        for x in range(n):
            t = (x + lo - center + 0.5) * ss
            if t < 0.0:
                t = -t
            w[x] = 1.0 - t if t < 1.0 else 0.0
            ww += w[x]
        for x in range(n):
            if ww != 0.0:
                w[x] /= ww
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        xmin[xx] = lo
        cnt[xx] = n
    return xmin, cnt, kk


def _clip8(ss):
    return np.clip(ss >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear(img, out_w, out_h):
    """Image.resize((w,h), BILINEAR) on RGB uint8: horizontal pass then vertical pass, uint8
    intermediate (Resample.c ImagingResampleInner).  A pass whose size is unchanged is skipped."""
    h, w, c = img.shape
    cur = img
    if out_w != w:
        xmin, cnt, kk = bilinear_coeffs(w, out_w)
        acc = np.full((h, out_w, c), 1 << (PRECISION_BITS - 1), np.int64)
        for t in range(kk.shape[1]):
            idx = np.minimum(xmin + t, w - 1)
            acc += cur[:, idx, :].astype(np.int64) * (kk[:, t] * (t < cnt))[None, :, None]
        cur = _clip8(acc)
    if out_h != h:
        ymin, cnt, kk = bilinear_coeffs(h, out_h)
        acc = np.full((out_h, cur.shape[1], c), 1 << (PRECISION_BITS - 1), np.int64)
        for t in range(kk.shape[1]):
            idx = np.minimum(ymin + t, h - 1)
            acc += cur[idx, :, :].astype(np.int64) * (kk[:, t] * (t < cnt))[:, None, None]
        cur = _clip8(acc)
    return cur


def nearest_index(in_size, out_size):
    """Geometry.c ImagingScaleAffine: source index = (int)o with o accumulated in double."""
    scale = float(in_size) / out_size
    o = 0.0 + scale * 0.5
    idx = np.empty(out_size, np.int64)
    for i in range(out_size):
        idx[i] = int(o) if o >= 0.0 else -1
        o += scale
    return idx


def resize_nearest(mask, out_w, out_h):
    """Image.resize((w,h), NEAREST) on an L image; same size -> copy."""
    h, w = mask.shape[:2]
    if out_w == w and out_h == h:
        return mask.copy()
    xi = nearest_index(w, out_w)
    yi = nearest_index(h, out_h)
    out = np.zeros((out_h, out_w) + mask.shape[2:], mask.dtype)
    okx = (xi >= 0) & (xi < w)
    oky = (yi >= 0) & (yi < h)
    sub = mask[np.clip(yi, 0, h - 1)][:, np.clip(xi, 0, w - 1)]
    sub[~oky, :] = 0
    sub[:, ~okx] = 0
    out[:] = sub
    return out


def crop_padding(w, h, tw, th, padding=0):
    """RandomCrop.__call__ (data/transform.py:35-41): border added on all four sides, or 0."""
    if padding > 0 or w < tw or h < th:
        return int(max(padding, max((tw - w) // 2 + 5, (th - h) // 2 + 5)))
    return 0


def pad_crop(arr, pad, x1, y1, tw, th):
    """ImageOps.expand(border=pad, fill=0) then crop((x1,y1,x1+tw,y1+th))."""
    if pad:
        shape = (arr.shape[0] + 2 * pad, arr.shape[1] + 2 * pad) + arr.shape[2:]
        big = np.zeros(shape, arr.dtype)
        big[pad:pad + arr.shape[0], pad:pad + arr.shape[1]] = arr
        arr = big
    return arr[y1:y1 + th, x1:x1 + tw].copy()


def scale_crop(img, mask, dec, tw, th):
    """DGRandomScaleCrop.scale + RandomCrop for one (image, original mask) pair
    (data/transform.py:104-112,35-55).  dec: dict(do_scale, w, h, x1, y1)."""
    if dec["do_scale"]:
        img = resize_bilinear(img, dec["w"], dec["h"])
        mask = resize_nearest(mask, dec["w"], dec["h"])
    h, w = img.shape[:2]
    pad = crop_padding(w, h, tw, th)
    w2, h2 = w + 2 * pad, h + 2 * pad
    if w2 == tw and h2 == th:
        if pad == 0:
            return img, mask
        return pad_crop(img, pad, 0, 0, tw, th), pad_crop(mask, pad, 0, 0, tw, th)
    return pad_crop(img, pad, dec["x1"], dec["y1"], tw, th), pad_crop(mask, pad, dec["x1"], dec["y1"], tw, th)


def normalize_image(img):
    """Normalize_dg.normalize (data/transform.py:150-152): float32 img/127.5 - 1, HWC."""
    out = img.astype(np.float32)
    out /= np.float32(127.5)
    out -= np.float32(1.0)
    return out


def normalize_mask(mask, dataset):
    """Normalize_dg.normalize mask branch (data/transform.py:153-172) + to_multilabel (:244-249).
    optic: >200 background -> [0,0]; (50,201) disc ring -> [0,1]; else cup -> [1,1]  (HWC, 2 ch).
    vessel: mask != 0 -> 1 (HWC, 1 ch)."""
    m = mask.astype(np.uint8)
    if dataset == "optic":
        out = np.zeros(m.shape + (2,), np.float32)
        ring = (m > 50) & (m < 201)
        cup = ~(m > 200) & ~ring
        out[ring] = (0, 1)
        out[cup] = (1, 1)
        return out
    return (m != 0).astype(np.float32)[..., None]


def to_tensor_pair(img_f32_hwc, mask_hwc):
    """ToTensor (data/transform.py:217-236): HWC -> CHW float32."""
    return (np.ascontiguousarray(img_f32_hwc.transpose(2, 0, 1)),
            np.ascontiguousarray(mask_hwc.astype(np.uint8).astype(np.float32).transpose(2, 0, 1)))
