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


# ---- gradients (the differentiable bank: data/operations.py:73-108, functional.py:21-46) --------------------------------
GRAD_FNS = dict(NAMES, sample_pairing="SamplePairing")
GRAD_WITH_MAG = ("solarize", "posterize", "contrast", "saturate", "brightness", "sharpness", "sample_pairing")


def test_golden_gradients_of_the_reference_bank(f32, golden_dir):
    """d/dx, d/dmagnitude, d/dmask of clamp(mask*fn(x,mag)+(1-mask)*x) from ONE CUDA kernel per op, against torch
    autograd run through the reference's own data/functional.py (scripts/make_golden_f32_grad.py): straight-through
    estimators (Solarize / Posterize -> magnitude only, AutoContrast / Equalize -> image), inclusive clamp gates, the
    reflect-padded blur's transpose, the pairing permutation's scatter.  float32: 1e-5 per element, 2e-4 on the sums."""
    g, x = load(golden_dir)
    gg = np.load(os.path.join(golden_dir, "f32_bank_grad.npz"))
    xt = torch.from_numpy(x).cuda()
    G = torch.from_numpy(gg["G"]).cuda()
    mask = torch.from_numpy(gg["masks"]).cuda()
    for fn, cls in GRAD_FNS.items():
        mag = torch.from_numpy(gg["mags"]).cuda() if fn in GRAD_WITH_MAG else None
        perm = torch.from_numpy(gg["pairing_perm"].astype(np.int32)).cuda() if fn == "sample_pairing" else None
        out = f32.apply(cls, xt, mag, mask, perm)
        assert np.abs(out.cpu().numpy() - gg["out_" + fn]).max() <= 3e-6, fn
        gx, gmag, gmask = f32.backward(cls, xt, G, mag, mask, perm)
        want = gg["gx_" + fn]
        assert np.abs(gx.cpu().numpy() - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), fn
        wm = gg["gmask_" + fn]
        assert np.allclose(gmask.cpu().numpy(), wm, rtol=2e-4, atol=2e-4 * np.abs(wm).max() + 1e-4), (fn, gmask, wm)
        if fn in GRAD_WITH_MAG:
            wg = gg["gmag_" + fn]
            assert np.allclose(gmag.cpu().numpy(), wg, rtol=2e-4, atol=2e-4 * np.abs(wg).max() + 1e-4), (fn, gmag, wg)


def _torch_geometric(x, kind, mag):
    """torch restatement of the pixel-space warp of SURVEY.md App. A.2 (Kornia absent: parity unpinned), differentiable
    in x and mag through F.grid_sample(align_corners=True, zeros)"""
    b, _, h, w = x.shape
    ys, xs = torch.meshgrid(torch.arange(h, device=x.device, dtype=torch.float32),
                            torch.arange(w, device=x.device, dtype=torch.float32), indexing="ij")
    xs, ys = xs[None], ys[None]
    m = mag.view(b, 1, 1)
    if kind == "shear_x":
        sx, sy = xs - m * ys, ys.expand(b, h, w)
    elif kind == "shear_y":
        sx, sy = xs.expand(b, h, w), ys - m * xs
    elif kind == "translate_x":
        sx, sy = xs - m * w, ys.expand(b, h, w)
    elif kind == "translate_y":
        sx, sy = xs.expand(b, h, w), ys - m * h
    else:
        a = m * (np.pi / 180.0)
        cx, cy = (w - 1) * 0.5, (h - 1) * 0.5
        u, v = xs - cx, ys - cy
        sx, sy = torch.cos(a) * u - torch.sin(a) * v + cx, torch.sin(a) * u + torch.cos(a) * v + cy
    grid = torch.stack([2 * sx / (w - 1) - 1, 2 * sy / (h - 1) - 1], dim=-1)
    return torch.nn.functional.grid_sample(x, grid, mode="bilinear", padding_mode="zeros", align_corners=True)


def _torch_hue(x, mag):
    """Kornia 0.2-era rgb_to_hsv / hsv_to_rgb with h in [0,1), differentiable"""
    r, g, b = x[:, 0], x[:, 1], x[:, 2]
    mx, mn = x.max(1).values, x.min(1).values
    d = mx - mn
    safe = torch.where(d > 0, d, torch.ones_like(d))
    hr = ((g - b) / safe) % 6
    hg = (b - r) / safe + 2
    hb = (r - g) / safe + 4
    h = torch.where(mx == r, hr, torch.where(mx == g, hg, hb)) / 6
    h = torch.where(d > 0, h % 1, torch.zeros_like(h))
    s = torch.where(mx > 0, d / torch.where(mx > 0, mx, torch.ones_like(mx)), torch.zeros_like(mx))
    v = mx
    h = (h + mag.view(-1, 1, 1)) % 1
    h6 = h * 6
    fi = torch.floor(h6).detach()
    f = h6 - fi
    p, q, t = v * (1 - s), v * (1 - f * s), v * (1 - (1 - f) * s)
    i = fi.long() % 6
    sel = lambda *c: sum(torch.where(i == k, c[k], torch.zeros_like(v)) for k in range(6))   # noqa: E731
    return torch.stack([sel(v, q, p, p, t, v), sel(t, v, v, q, p, p), sel(p, p, t, v, v, q)], 1)


def test_gradients_of_the_kornia_backed_ops_vs_torch_autograd(f32):
    """the six Kornia-backed ops (parity unpinned): forward AND gradients against torch autograd through a torch
    restatement of the same pixel-space warp / HSV round trip"""
    torch.manual_seed(3)
    b, h, w = 3, 24, 40
    x0 = torch.rand(b, 3, h, w, device="cuda")
    G = torch.randn(b, 3, h, w, device="cuda")
    mask0 = torch.tensor([0.3, 0.8, 1.0], device="cuda")
    # magnitudes chosen so that no source coordinate lands on an integer (there the bilinear cell, hence d/dmag, is a
    # matter of the last float bit: 0.3 * y is an integer on whole rows)
    cases = [("shear_x", "ShearX", [0.21, -0.31, 0.07]), ("shear_y", "ShearY", [-0.11, 0.29, 0.19]),
             ("translate_x", "TranslateX", [0.21, -0.44, 0.06]), ("translate_y", "TranslateY", [-0.3, 0.1, 0.45]),
             ("rotate", "Rotate", [17.0, -30.0, 4.5]), ("hue", "Hue", [0.3, 1.7, 0.95])]
    for kind, cls, mags in cases:
        x = x0.clone().requires_grad_(True)
        mag = torch.tensor(mags, device="cuda", requires_grad=True)
        mask = mask0.clone().requires_grad_(True)
        y = (_torch_hue(x, mag) if kind == "hue" else _torch_geometric(x, kind, mag)).clamp(0, 1)
        m4 = mask.view(b, 1, 1, 1)
        o = (m4 * y + (1 - m4) * x).clamp(0, 1)
        (o * G).sum().backward()
        out = f32.apply(cls, x0, mag.detach(), mask0)
        assert (out - o.detach()).abs().max().item() <= 5e-5, kind
        gx, gmag, gmask = f32.backward(cls, x0, G, mag.detach(), mask0)
        scale = x.grad.abs().max().item()
        bad = ((gx - x.grad).abs() > 2e-3 * scale).float().mean().item()        # interpolation cells at .0 fractions
        assert bad <= 2e-3, (kind, bad)
        assert torch.allclose(gmask, mask.grad, rtol=2e-3, atol=2e-3 * mask.grad.abs().max().item() + 1e-3), (kind, gmask, mask.grad)
        assert torch.allclose(gmag, mag.grad, rtol=5e-3, atol=5e-3 * mag.grad.abs().max().item() + 1e-2), (kind, gmag, mag.grad)


def test_operation_modules_learn_probability_and_magnitude(f32):
    """operations.py:73-108 end to end: in training mode the loss reaches `_probability` (through the RelaxedBernoulli
    sample) and `_magnitude`; the gradient equals torch autograd's through the reference composition with the same
    mask / sign draws."""
    from aadg_b200.data import operations as O
    x = torch.rand(4, 3, 16, 20, device="cuda")
    G = torch.randn(4, 3, 16, 20, device="cuda")
    for name, has_mag in (("Brightness", True), ("Solarize", True), ("Invert", False), ("Rotate", True)):
        op = getattr(O, name)().cuda().train()
        torch.manual_seed(5)
        out = op(x)
        assert out.requires_grad
        (out * G).sum().backward()
        assert op._probability.grad is not None and torch.isfinite(op._probability.grad).all()
        assert float(op._probability.grad.abs().sum()) > 0
        if has_mag:
            assert op._magnitude.grad is not None and torch.isfinite(op._magnitude.grad).all()
            assert float(op._magnitude.grad.abs().sum()) > 0
        op.eval()
        with torch.no_grad():
            assert not op(x).requires_grad
    # Brightness against the composed torch expression, same draws
    op = O.Brightness(initial_magnitude=0.3).cuda().train()
    torch.manual_seed(9)
    out = op(x)
    (out * G).sum().backward()
    ref = O.Brightness(initial_magnitude=0.3).cuda().train()
    torch.manual_seed(9)
    mask = ref.get_mask(4).reshape(4, 1, 1, 1)
    y = (x * (1 - ref.magnitude.view(-1, 1, 1, 1))).clamp(0, 1)
    o = (mask * y + (1 - mask) * x).clamp(0, 1)
    (o * G).sum().backward()
    assert torch.allclose(out, o, atol=2e-6)
    assert torch.allclose(op._magnitude.grad, ref._magnitude.grad, rtol=1e-3, atol=1e-3)
    assert torch.allclose(op._probability.grad, ref._probability.grad, rtol=1e-3, atol=1e-3)
