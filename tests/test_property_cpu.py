"""CPU property tests (hypothesis): the uint8-bank oracle against Pillow ITSELF over op x magnitude x image content
(SURVEY.md §4(5)).  The reference's live ops are thin wrappers over these Pillow calls (data/basic.py:70-120), so any
divergence of the numpy restatement from Pillow's C routines on unusual content (constant images, two-level images,
saturated gradients, odd sizes) shows up here, beyond the fixed golden files."""
import numpy as np
import pytest

hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st, HealthCheck  # noqa: E402

PIL = pytest.importorskip("PIL")
import PIL.Image  # noqa: E402
import PIL.ImageEnhance  # noqa: E402
import PIL.ImageOps  # noqa: E402

from oracle import u8_bank as B  # noqa: E402
from oracle import u8_transform as T  # noqa: E402

PILLOW = {
    "AutoContrast": lambda im, v: PIL.ImageOps.autocontrast(im),
    "Invert": lambda im, v: PIL.ImageOps.invert(im),
    "Equalize": lambda im, v: PIL.ImageOps.equalize(im),
    "Solarize": lambda im, v: PIL.ImageOps.solarize(im, v),
    "Posterize": lambda im, v: PIL.ImageOps.posterize(im, int(v)),
    "Contrast": lambda im, v: PIL.ImageEnhance.Contrast(im).enhance(v),
    "Color": lambda im, v: PIL.ImageEnhance.Color(im).enhance(v),
    "Brightness": lambda im, v: PIL.ImageEnhance.Brightness(im).enhance(v),
    "Sharpness": lambda im, v: PIL.ImageEnhance.Sharpness(im).enhance(v),
}


@st.composite
def images(draw, min_side=3, max_side=40):
    h = draw(st.integers(min_side, max_side))
    w = draw(st.integers(min_side, max_side))
    kind = draw(st.sampled_from(["noise", "constant", "two_level", "gradient", "narrow", "saturated"]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.RandomState(seed)
    if kind == "noise":
        img = rng.randint(0, 256, (h, w, 3))
    elif kind == "constant":
        img = np.broadcast_to(rng.randint(0, 256, (1, 1, 3)), (h, w, 3))
    elif kind == "two_level":
        lo, hi = sorted(rng.randint(0, 256, 2))
        img = np.where(rng.rand(h, w, 1) > 0.5, hi, lo) * np.ones((1, 1, 3), int)
    elif kind == "gradient":
        img = (np.arange(w)[None, :, None] * 255 // max(w - 1, 1) + np.arange(h)[:, None, None]) % 256 * np.ones((1, 1, 3), int)
    elif kind == "narrow":
        base = rng.randint(0, 250)
        img = base + rng.randint(0, 6, (h, w, 3))
    else:
        img = np.where(rng.rand(h, w, 3) > 0.5, 255, rng.randint(0, 256, (h, w, 3)))
    return np.ascontiguousarray(img.astype(np.uint8))


@settings(max_examples=150, deadline=None, suppress_health_check=list(HealthCheck))
@given(img=images(), op=st.sampled_from(sorted(PILLOW)), level=st.integers(0, 9))
def test_oracle_op_equals_pillow(img, op, level):
    v = B.level_to_value(op, level / 9)
    want = np.asarray(PILLOW[op](PIL.Image.fromarray(img), v))
    got, _ = B.apply_op(img, np.zeros(img.shape[:2], np.uint8), op, level / 9, {})
    assert np.array_equal(got, want), (op, level, img.shape)


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(img=images(4, 32), ops=st.lists(st.tuples(st.sampled_from(sorted(PILLOW)), st.integers(0, 9)), min_size=2, max_size=3))
def test_oracle_chain_equals_pillow_chain(img, ops):
    """sub-policies are chains (data/policy.py:24-28): intermediate uint8 rounding must match op by op"""
    pil = PIL.Image.fromarray(img)
    cur = img
    for op, level in ops:
        pil = PILLOW[op](pil, B.level_to_value(op, level / 9))
        cur, _ = B.apply_op(cur, np.zeros(img.shape[:2], np.uint8), op, level / 9, {})
    assert np.array_equal(cur, np.asarray(pil)), ops


@settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))
@given(img=images(3, 48), sx=st.floats(0.4, 2.2), sy=st.floats(0.4, 2.2))
def test_oracle_resize_equals_pillow(img, sx, sy):
    """DGRandomScaleCrop's resizes (data/transform.py:104-112): Pillow BILINEAR (antialiased two-pass fixed point) and
    NEAREST at arbitrary up/down/anisotropic scales"""
    h, w = img.shape[:2]
    nw, nh = max(1, int(sx * w)), max(1, int(sy * h))
    want = np.asarray(PIL.Image.fromarray(img).resize((nw, nh), PIL.Image.BILINEAR))
    assert np.array_equal(T.resize_bilinear(img, nw, nh), want), (w, h, nw, nh)
    mask = np.ascontiguousarray(img[..., 0])
    wantm = np.asarray(PIL.Image.fromarray(mask).resize((nw, nh), PIL.Image.NEAREST))
    assert np.array_equal(T.resize_nearest(mask, nw, nh), wantm), (w, h, nw, nh)
