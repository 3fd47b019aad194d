"""GPU: the CUDA float tensor bank against the reference-generated golden file (13 ops) and the oracle
(all 19 ops), float32 tolerance 2e-6 (pure arithmetic) / 2e-5 (interpolation)."""
import os

import numpy as np
import pytest
import torch

from oracle import f32_bank as B

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def f32():
    assert torch.cuda.is_available()
    from aadg_b200.ops import f32 as mod
    return mod


def load(golden_dir):
    g = np.load(os.path.join(golden_dir, "f32_bank.npz"))
    x = g["imgs_u8"].transpose(0, 3, 1, 2).astype(np.float32) / np.float32(255)
    return g, x


def run(f32, op, x, mag=None, mask=None, perm=None):
    return f32.apply(op, torch.from_numpy(np.ascontiguousarray(x)).cuda(), mag, mask, perm).cpu().numpy()


NAMES = {"hflip": "HorizontalFlip", "vflip": "VerticalFlip", "invert": "Invert", "gray": "Gray",
         "auto_contrast": "AutoContrast", "equalize": "Equalize", "solarize": "Solarize", "posterize": "Posterize",
         "contrast": "Contrast", "saturate": "Saturate", "brightness": "Brightness", "sharpness": "Sharpness"}


def test_golden_functional_ops(f32, golden_dir):
    g, x = load(golden_dir)
    for fn, cls in NAMES.items():
        mag = g["mags"] if fn in ("solarize", "posterize", "contrast", "saturate", "brightness", "sharpness") else None
        got = run(f32, cls, x, mag)
        assert np.abs(got - g["fn_" + fn]).max() <= 2e-6, fn
    got = run(f32, "SamplePairing", x, g["mags"], None, g["pairing_perm"])
    assert np.abs(got - g["fn_sample_pairing"]).max() <= 2e-6


def test_golden_operation_forward(f32, golden_dir):
    g, x = load(golden_dir)
    for cls in ("Solarize", "Sharpness", "Invert"):
        mag = None if cls == "Invert" else np.full(4, g["op_mag_" + cls], np.float32)
        for mode in ("train", "eval"):
            if mag is not None:
                mag = np.abs(mag) * g["op_%s_sign_%s" % (mode, cls)]
            got = run(f32, cls, x, mag, g["op_%s_mask_%s" % (mode, cls)])
            assert np.abs(got - g["op_%s_%s" % (mode, cls)]).max() <= 3e-6, (cls, mode)


@pytest.mark.parametrize("shape", [(3, 37, 53), (2, 64, 128)])
def test_all_ops_vs_oracle(f32, shape):
    b, h, w = shape
    rng = np.random.RandomState(1)
    x = (rng.randint(0, 256, (b, 3, h, w)) / 255.0).astype(np.float32)
    mag = rng.rand(b).astype(np.float32)
    mask = rng.rand(b).astype(np.float32)
    perm = rng.permutation(b).astype(np.int32)
    for fn, cls in NAMES.items():
        m = mag if fn in ("solarize", "posterize", "contrast", "saturate", "brightness", "sharpness") else None
        want = B.operation(x, getattr(B, fn), m, mask)
        assert np.abs(run(f32, cls, x, m, mask) - want).max() <= 3e-6, fn
    want = B.operation(x, B.sample_pairing, mag, mask, perm=perm)
    assert np.abs(run(f32, "SamplePairing", x, mag, mask, perm) - want).max() <= 3e-6
    for kind, cls, scale in (("shear_x", "ShearX", 0.3), ("shear_y", "ShearY", 0.3), ("translate_x", "TranslateX", 0.45),
                             ("translate_y", "TranslateY", 0.45), ("rotate", "Rotate", 30.0)):
        mg = ((mag * 2 - 1) * scale).astype(np.float32)
        want = B.operation(x, lambda im, m_, k=kind: B.geometric(im, k, m_), mg, mask)
        assert np.abs(run(f32, cls, x, mg, mask) - want).max() <= 3e-5, kind
    want = B.operation(x, B.hue, (mag * 2).astype(np.float32), mask)
    assert np.abs(run(f32, "Hue", x, (mag * 2).astype(np.float32), mask) - want).max() <= 2e-5


def test_operation_modules(f32):
    from aadg_b200.data import operations as O
    torch.manual_seed(0)
    x = torch.rand(6, 3, 32, 32, device="cuda")
    assert sorted(O.__all__) == sorted(f32.OPS)
    for name in O.__all__:
        op = getattr(O, name)().cuda()
        y = op(x)
        assert y.shape == x.shape and float(y.min()) >= 0 and float(y.max()) <= 1
        op.eval()
        y = op(x)
        changed = (y != x).flatten(1).any(1)
        assert y.shape == x.shape and (changed.sum() <= 6)
    inv = O.Invert(initial_probability=1.0, probability_range=None).cuda().eval()
    assert torch.allclose(inv(x), 1 - x)
    rot = O.Rotate(initial_magnitude=1.0)
    assert abs(float(rot.magnitude) - 30.0) < 1e-6 and rot.flip_magnitude
    with pytest.raises(RuntimeError):
        f32.apply("Invert", torch.rand(1, 3, 4, 4))
