"""ORACLE (test infrastructure, never on the product path).

NumPy restatement of the reference's float tensor bank (`/root/reference/data/functional.py:110-271`,
`data/operations.py:73-108`, `data/kernels.py:9-13`) — dead code in the reference, named by the north star.
Images float32 `[B,3,H,W]` in [0,1]; `mag` float32 `[B]`.  Every op ends with clamp(0,1) like
`tensor_function` (functional.py:47-72).

Pinned by tests/test_oracle_f32.py against tests/golden/f32_bank.npz — outputs of the reference itself for
the 13 ops that run without Kornia (scripts/make_golden_f32.py).  The six Kornia-backed ops (shear_x/y,
translate_x/y, rotate, hue) are **parity unpinned**: Kornia is absent and unversioned in the reference
(functional.py:4, not in requirements.txt); they follow SURVEY.md App. A.2 and are checked against torch's
affine_grid/grid_sample and colorsys.

equalize: the reference histograms all B*C planes with ONE torch.histc over `x*255 + 256*plane`
(functional.py:242-262), whose bin width (n*256-1)/(n*256) drifts by up to one bin for later planes; this
restatement histograms each plane on floor(x*255) — the documented intent (the Pillow algorithm it cites) —
and the golden test quantifies the agreement.
"""
import numpy as np

F = np.float32


def _clamp(x):
    return np.clip(x, F(0), F(1)).astype(F)


def _mag(mag, b):
    m = np.asarray(mag, F).reshape(-1)
    return (np.repeat(m, b) if m.size == 1 else m).reshape(b, 1, 1, 1)


def gray_luma(img):
    """functional.py:85-87 (sic: 0.110 for blue)"""
    return (F(0.299) * img[:, 0:1] + F(0.587) * img[:, 1:2] + F(0.110) * img[:, 2:3]).astype(F)


def blend(img1, img2, alpha):
    """functional.py:76-82: alpha = 1 returns img1"""
    return _clamp(img2 + alpha * (img1 - img2))


def hflip(img, mag=None):
    return _clamp(img[..., ::-1])


def vflip(img, mag=None):
    return _clamp(img[..., ::-1, :])


def invert(img, mag=None):
    return _clamp(F(1) - img)


def solarize(img, mag):
    m = _mag(mag, len(img))
    return _clamp(np.where(img < m, img, F(1) - img))


def posterize(img, mag):
    """functional.py:176-184: (long(x*255) << s) >> s is the identity on the integer part"""
    return _clamp((img * F(255)).astype(np.int64).astype(F) / F(255))


def gray(img, mag=None):
    return _clamp(np.repeat(gray_luma(img), 3, axis=1))


def contrast(img, mag):
    b = len(img)
    mean = np.floor(gray_luma(img * F(255)).reshape(b, -1).mean(1, dtype=F) + F(0.5)).reshape(b, 1, 1, 1) / F(255)
    return blend(img, mean.astype(F), F(1) - _mag(mag, b))


def auto_contrast(img, mag=None):
    b, c, h, w = img.shape
    r = (_clamp(img) * F(255)).reshape(b * c, h * w)
    lo, hi = r.min(1, keepdims=True), r.max(1, keepdims=True)
    lut = np.floor((np.arange(256, dtype=F)[None, :] - lo) * (F(255) / (hi - lo + F(0.1))))
    out = np.take_along_axis(lut, r.astype(np.int64), axis=1)
    return _clamp((out / F(255)).reshape(img.shape).astype(F))


def saturate(img, mag):
    return blend(img, gray_luma(img), F(1) - _mag(mag, len(img)))


def brightness(img, mag):
    return blend(img, np.zeros_like(img), F(1) - _mag(mag, len(img)))


def sample_pairing(img, mag, perm):
    m = _mag(mag, len(img))
    return _clamp((F(1) - m) * img + m * img[perm])


def equalize(img, mag=None):
    b, c, h, w = img.shape
    idx = (_clamp(img) * F(255)).astype(np.int64).reshape(b * c, h * w)
    out = np.empty((b * c, h * w), F)
    for p in range(b * c):
        hist = np.bincount(idx[p], minlength=256).astype(F)
        cdf = np.cumsum(hist, dtype=F)
        step = np.floor((cdf[-1] - hist[-1]) / F(255))
        cdf_ex = np.concatenate([[F(0)], cdf])[:256] + np.floor(step / F(2))
        lut = np.floor(cdf_ex / (step + F(0.1)))
        out[p] = lut[idx[p]] / F(255)
    return _clamp(out.reshape(img.shape))


SHARP_K = (np.array([[1, 1, 1], [1, 5, 1], [1, 1, 1]], F) / F(13)).astype(F)


def blur3x3_reflect(img, k=SHARP_K):
    """functional.py:98-106: reflect-pad 1, depthwise 3x3 correlation"""
    p = np.pad(img, ((0, 0), (0, 0), (1, 1), (1, 1)), mode="reflect")
    h, w = img.shape[2:]
    out = np.zeros_like(img)
    for r in range(3):
        for s in range(3):
            out += k[r, s] * p[:, :, r:r + h, s:s + w]
    return out.astype(F)


def sharpness(img, mag):
    return blend(img, blur3x3_reflect(img), F(1) - _mag(mag, len(img)))


# ---- Kornia-backed ops: parity unpinned ---------------------------------------------
def affine_matrices(kind, mag, h, w):
    """forward pixel-space matrices M (dst = M src) [B,2,3]"""
    m = np.asarray(mag, np.float64).reshape(-1)
    out = np.zeros((len(m), 2, 3))
    out[:, 0, 0] = out[:, 1, 1] = 1
    if kind == "shear_x":
        out[:, 0, 1] = m
    elif kind == "shear_y":
        out[:, 1, 0] = m
    elif kind == "translate_x":
        out[:, 0, 2] = m * w
    elif kind == "translate_y":
        out[:, 1, 2] = m * h
    elif kind == "rotate":
        a = np.deg2rad(m)
        c, s = np.cos(a), np.sin(a)
        cx, cy = (w - 1) / 2.0, (h - 1) / 2.0
        out[:, 0, 0], out[:, 0, 1], out[:, 0, 2] = c, s, (1 - c) * cx - s * cy
        out[:, 1, 0], out[:, 1, 1], out[:, 1, 2] = -s, c, s * cx + (1 - c) * cy
    else:
        raise KeyError(kind)
    return out


def warp_affine(img, M):
    """dst(p) = bilinear_zero_pad(src, M^-1 p), pixel-centre coordinates"""
    b, c, h, w = img.shape
    out = np.zeros_like(img)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    for i in range(b):
        A = np.vstack([M[i], [0, 0, 1]])
        inv = np.linalg.inv(A)
        sx = (inv[0, 0] * xs + inv[0, 1] * ys + inv[0, 2]).astype(F)
        sy = (inv[1, 0] * xs + inv[1, 1] * ys + inv[1, 2]).astype(F)
        x0, y0 = np.floor(sx).astype(np.int64), np.floor(sy).astype(np.int64)
        fx, fy = (sx - x0).astype(F), (sy - y0).astype(F)
        acc = np.zeros((c, h, w), F)
        for dy, wy in ((0, 1 - fy), (1, fy)):
            for dx, wx in ((0, 1 - fx), (1, fx)):
                xi, yi = x0 + dx, y0 + dy
                ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
                v = img[i][:, np.clip(yi, 0, h - 1), np.clip(xi, 0, w - 1)]
                acc += np.where(ok, wy * wx, 0).astype(F)[None] * v
        out[i] = acc
    return _clamp(out)


def geometric(img, kind, mag):
    return warp_affine(img, affine_matrices(kind, mag, img.shape[2], img.shape[3]))


def rgb_to_hsv(img):
    r, g, b = img[:, 0], img[:, 1], img[:, 2]
    mx, mn = img.max(1), img.min(1)
    d = mx - mn
    s = np.where(mx > 0, d / np.where(mx > 0, mx, 1), 0).astype(F)
    dz = np.where(d > 0, d, 1)
    hr = ((g - b) / dz) % 6
    hg = (b - r) / dz + 2
    hb = (r - g) / dz + 4
    hh = np.where(mx == r, hr, np.where(mx == g, hg, hb))
    hh = np.where(d > 0, hh / 6.0, 0) % 1.0
    return np.stack([hh.astype(F), s, mx.astype(F)], 1)


def hsv_to_rgb(hsv):
    h, s, v = hsv[:, 0], hsv[:, 1], hsv[:, 2]
    i = np.floor(h * 6).astype(np.int64)
    f = (h * 6 - i).astype(F)
    p, q, t = v * (1 - s), v * (1 - f * s), v * (1 - (1 - f) * s)
    i = i % 6
    r = np.choose(i, [v, q, p, p, t, v])
    g = np.choose(i, [t, v, v, q, p, p])
    b = np.choose(i, [p, p, t, v, v, q])
    return np.stack([r, g, b], 1).astype(F)


def hue(img, mag):
    hsv = rgb_to_hsv(img)
    hsv[:, 0] = (hsv[:, 0] + _mag(mag, len(img)).reshape(-1, 1, 1)) % F(1)
    return _clamp(hsv_to_rgb(hsv))


def operation(img, fn, mag, mask, **kw):
    """_Operation.forward, training form (operations.py:73-100): clamp(mask*op(x) + (1-mask)*x)."""
    m = np.asarray(mask, F).reshape(-1, 1, 1, 1)
    y = fn(img, mag, **kw) if mag is not None else fn(img)
    return _clamp(m * y + (F(1) - m) * img)
