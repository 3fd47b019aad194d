"""GPU: tcgen05 convolution kernels (C ABI) against torch fp32 convolutions of the same bf16-rounded
operands.  Tolerance: fp32 accumulation of bf16 products, bf16 output rounding (2^-8 relative)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def conv():
    assert torch.cuda.is_available()
    from aadg_b200.ops import conv as mod
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return mod


def rand_case(n, h, w, cin, cout, r, s, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, r, s, device="cuda", generator=g) / (cin * r * s) ** 0.5).bfloat16()
    return x, wt


def to_taps(wt):       # [Cout,Cin,R,S] -> [R*S,Cout,Cin]
    co, ci, r, s = wt.shape
    return wt.permute(2, 3, 0, 1).reshape(r * s, co, ci).contiguous()


def to_taps_t(wt):     # [Cout,Cin,R,S] -> [R*S,Cin,Cout]
    co, ci, r, s = wt.shape
    return wt.permute(2, 3, 1, 0).reshape(r * s, ci, co).contiguous()


CASES = [
    # n, h, w, cin, cout, r, s, stride, pad, dil
    (2, 16, 16, 64, 64, 1, 1, 1, 0, 1),
    (2, 16, 16, 64, 128, 3, 3, 1, 1, 1),
    (3, 20, 12, 128, 64, 3, 3, 1, 1, 1),
    (2, 32, 32, 64, 256, 1, 1, 1, 0, 1),
    (2, 32, 32, 128, 128, 3, 3, 2, 1, 1),
    (2, 32, 32, 256, 512, 1, 1, 2, 0, 1),
    (2, 16, 16, 512, 512, 3, 3, 1, 2, 2),
    (1, 32, 32, 256, 256, 3, 3, 1, 12, 12),
    (5, 8, 8, 2048, 256, 1, 1, 1, 0, 1),
    (2, 64, 64, 304, 256, 3, 3, 1, 1, 1),
    (2, 64, 64, 256, 48, 1, 1, 1, 0, 1),
    (1, 130, 70, 64, 64, 3, 3, 1, 1, 1),
    (2, 17, 19, 192, 64, 1, 1, 1, 0, 1),
    (1, 30, 30, 64, 64, 7, 7, 2, 3, 1),
    # enough 256-column tiles for the BN = 256 kernel
    (8, 64, 64, 256, 512, 1, 1, 1, 0, 1),
    (5, 64, 64, 64, 256, 3, 3, 1, 1, 1),
    (6, 60, 50, 320, 256, 1, 1, 1, 0, 1),
    (4, 32, 32, 256, 128, 3, 3, 1, 1, 1),       # 256-wide weight-gradient tiles
    # halo-reuse kernel (Cin <= 64, rows of 128 pixels): partial tiles, dilation, small / several channel blocks
    (2, 40, 200, 64, 64, 3, 3, 1, 1, 1),
    (1, 20, 128, 32, 48, 3, 3, 1, 2, 2),
    (1, 16, 256, 16, 128, 3, 3, 1, 1, 1),
    (3, 9, 130, 64, 64, 3, 3, 1, 1, 1),
]


def ref_conv(x, wt, stride, pad, dil):
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), stride=stride, padding=pad, dilation=dil)
    return y.permute(0, 2, 3, 1).contiguous()


def close(got, want, what):
    err = (got.float() - want).abs().max().item()
    scale = want.abs().max().item() + 1e-6
    assert err <= 1.2e-2 * scale, (what, err, scale)
    # bf16 output rounding dominates: mean error must be far smaller
    assert (got.float() - want).abs().mean().item() <= 2.5e-3 * scale, what


@pytest.mark.parametrize("case", CASES)
def test_fprop(conv, case):
    n, h, w, cin, cout, r, s, stride, pad, dil = case
    x, wt = rand_case(n, h, w, cin, cout, r, s)
    y = conv.fprop(x, to_taps(wt), r, s, stride, pad, dil)
    close(y, ref_conv(x, wt, stride, pad, dil), case)


@pytest.mark.parametrize("case", CASES)
def test_dgrad(conv, case):
    n, h, w, cin, cout, r, s, stride, pad, dil = case
    x, wt = rand_case(n, h, w, cin, cout, r, s, seed=1)
    ho, wo = conv.out_size(h, w, r, s, stride, pad, dil)
    dy = torch.randn(n, ho, wo, cout, device="cuda").bfloat16()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    F.conv2d(xr, wt.float(), stride=stride, padding=pad, dilation=dil).backward(dy.float().permute(0, 3, 1, 2))
    want = xr.grad.permute(0, 2, 3, 1).contiguous()
    got = conv.dgrad(dy, to_taps_t(wt), r, s, stride, pad, dil, (h, w))
    close(got, want, case)


@pytest.mark.parametrize("case", CASES)
def test_wgrad(conv, case):
    n, h, w, cin, cout, r, s, stride, pad, dil = case
    x, wt = rand_case(n, h, w, cin, cout, r, s, seed=2)
    ho, wo = conv.out_size(h, w, r, s, stride, pad, dil)
    dy = torch.randn(n, ho, wo, cout, device="cuda").bfloat16()
    wr = wt.float().requires_grad_(True)
    F.conv2d(x.float().permute(0, 3, 1, 2), wr, stride=stride, padding=pad, dilation=dil).backward(
        dy.float().permute(0, 3, 1, 2))
    want = wr.grad.permute(2, 3, 0, 1).reshape(r * s, cout, cin)
    got = conv.wgrad(x, dy, r, s, stride, pad, dil)
    err = (got - want).abs().max().item()
    assert err <= 2e-3 * (want.abs().max().item() + 1e-6), (case, err)


def test_channel_slices_and_accumulate(conv):
    """concat buffers: read a channel slice, write into a channel slice, accumulate into the output."""
    x_full = torch.randn(2, 16, 16, 192, device="cuda").bfloat16()
    x = x_full[..., 64:192]
    _, wt = rand_case(1, 1, 1, 128, 64, 3, 3, seed=3)
    out_full = torch.zeros(2, 16, 16, 256, device="cuda", dtype=torch.bfloat16)
    out = out_full[..., 128:192]
    conv.fprop(x, to_taps(wt), 3, 3, 1, 1, 1, out=out)
    want = ref_conv(x.contiguous(), wt, 1, 1, 1)
    close(out, want, "slice")
    assert out_full[..., :128].abs().max() == 0 and out_full[..., 192:].abs().max() == 0
    conv.fprop(x, to_taps(wt), 3, 3, 1, 1, 1, out=out, accumulate=True)
    close(out, 2 * want, "accumulate")


@pytest.mark.parametrize("case", CASES)
def test_fprop_fused_statistics(conv, case):
    """the epilogue's per-channel sum / sum of squares equal those of the bf16 tensor it stored (partial tiles,
    channel tails and the 64/128/256-column kernels included) and the output is unchanged"""
    n, h, w, cin, cout, r, s, stride, pad, dil = case
    x, wt = rand_case(n, h, w, cin, cout, r, s, seed=4)
    ssum = torch.zeros(cout, device="cuda")
    ssq = torch.zeros(cout, device="cuda")
    y = conv.fprop(x, to_taps(wt), r, s, stride, pad, dil, stats=(ssum, ssq))
    y0 = conv.fprop(x, to_taps(wt), r, s, stride, pad, dil)
    assert torch.equal(y, y0)
    yf = y.double().reshape(-1, cout)
    want_sum, want_sq = yf.sum(0), (yf * yf).sum(0)
    tol = 1e-4 * (yf.abs().sum(0) + 1e-3)
    assert ((ssum.double() - want_sum).abs() <= tol).all(), (case, (ssum.double() - want_sum).abs().max().item())
    assert ((ssq.double() - want_sq).abs() <= 1e-4 * want_sq + 1e-6).all(), case


def test_halo_kernel_channel_slices_and_accumulate(conv):
    """the halo-reuse kernels (rows of 128 pixels, Cin <= 64) with strided channel slices on both sides and the
    reduce-add epilogue; data gradient accumulated into an existing tensor"""
    x_full = torch.randn(2, 12, 200, 96, device="cuda").bfloat16()
    x = x_full[..., 32:96]
    _, wt = rand_case(1, 1, 1, 64, 48, 3, 3, seed=6)
    out_full = torch.zeros(2, 12, 200, 128, device="cuda", dtype=torch.bfloat16)
    out = out_full[..., 64:112]
    conv.fprop(x, to_taps(wt), 3, 3, 1, 1, 1, out=out)
    want = ref_conv(x.contiguous(), wt, 1, 1, 1)
    close(out, want, "halo slice")
    assert out_full[..., :64].abs().max() == 0 and out_full[..., 112:].abs().max() == 0
    conv.fprop(x, to_taps(wt), 3, 3, 1, 1, 1, out=out, accumulate=True)
    close(out, 2 * want, "halo accumulate")
    dy = torch.randn(2, 12, 200, 48, device="cuda").bfloat16()
    xr = x.contiguous().float().permute(0, 3, 1, 2).requires_grad_(True)
    F.conv2d(xr, wt.float(), padding=1).backward(dy.float().permute(0, 3, 1, 2))
    base = torch.randn(2, 12, 200, 64, device="cuda").bfloat16()
    got = base.clone()
    conv.dgrad(dy, to_taps_t(wt), 3, 3, 1, 1, 1, (12, 200), out=got, accumulate=True)
    close(got, base.float() + xr.grad.permute(0, 2, 3, 1), "halo dgrad accumulate")
