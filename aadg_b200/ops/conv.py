"""Convolutions on tcgen05 tensor cores: torch-facing wrappers over aadg_conv_* (csrc/conv_tc.cu).

Tensors are bf16 NHWC views (`[N,H,W,C]`, channel stride = x.stride(2), which may exceed C when the
tensor is a channel slice of a concat buffer).  Weights: bf16 `[R*S, Cout, Cin]` (fprop/wgrad) and
`[R*S, Cin, Cout]` (dgrad).  There is no CPU path."""
import torch

from .. import _lib


# bench.py sets TIMING = [] to bracket every launch with CUDA events: (kind, algorithmic flops, start, end, geometry);
# `flops=` lets a caller that runs a reformulated problem (pixel packing) report the ORIGINAL layer's FLOPs
TIMING = None


class _timed:
    def __init__(self, kind, flops, geom=None):
        self.kind, self.flops, self.geom = kind, flops, geom

    def __enter__(self):
        if TIMING is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if TIMING is not None:
            self.e1.record()
            TIMING.append((self.kind, self.flops, self.e0, self.e1, self.geom))


def _nhwc(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.bfloat16 and t.dim() == 4):
        raise RuntimeError("%s must be a CUDA bf16 [N,H,W,C] tensor (no CPU path)" % name)
    n, h, w, c = t.shape
    ld = t.stride(2)
    if t.stride(3) != 1 or t.stride(1) != w * ld or t.stride(0) != h * w * ld:
        raise ValueError("%s must be NHWC-contiguous up to a channel stride" % name)
    return n, h, w, c, ld


def out_size(h, w, r, s, stride, pad, dil):
    return ((h + 2 * pad - dil * (r - 1) - 1) // stride + 1, (w + 2 * pad - dil * (s - 1) - 1) // stride + 1)


def fprop(x, wgt, r, s, stride=1, pad=0, dil=1, out=None, accumulate=False, stats=None, flops=None):
    """x [N,H,W,Cin] bf16, wgt [R*S,Cout,Cin] bf16 -> y [N,Ho,Wo,Cout] bf16 (or written into `out`).
    stats = (sum, sumsq) fp32 [Cout] tensors (zeroed by the caller): the batch-norm statistics of y are accumulated
    in the convolution's epilogue instead of a separate pass over y."""
    n, h, w, cin, ldx = _nhwc(x, "x")
    cout = wgt.shape[1]
    assert wgt.shape == (r * s, cout, cin) and wgt.dtype == torch.bfloat16 and wgt.is_contiguous()
    ho, wo = out_size(h, w, r, s, stride, pad, dil)
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    _, oh, ow, oc, ldy = _nhwc(out, "out")
    assert (oh, ow, oc) == (ho, wo, cout)
    with _lib.on_device(x.device), _timed("fprop", flops or 2.0 * n * ho * wo * cout * cin * r * s,
                                              (n, h, w, cin, ho, wo, cout, r, stride, dil)):
        if stats is not None:
            assert not accumulate
            ssum, ssq = stats
            assert ssum.dtype == torch.float32 and ssq.dtype == torch.float32 and ssum.numel() == cout == ssq.numel()
            _lib.check(_lib.lib().aadg_conv_fprop_stats_bf16(x.data_ptr(), n, h, w, cin, ldx, wgt.data_ptr(), cout, r, s,
                                                             stride, pad, dil, out.data_ptr(), ho, wo, ldy, 0,
                                                             ssum.data_ptr(), ssq.data_ptr(), _lib.stream_ptr()))
        else:
            _lib.check(_lib.lib().aadg_conv_fprop_bf16(x.data_ptr(), n, h, w, cin, ldx, wgt.data_ptr(), cout, r, s, stride,
                                                       pad, dil, out.data_ptr(), ho, wo, ldy, 0, int(accumulate),
                                                       _lib.stream_ptr()))
    return out


def dgrad(dy, wgt_t, r, s, stride, pad, dil, in_hw, out=None, accumulate=False, flops=None):
    """dy [N,Ho,Wo,Cout] bf16, wgt_t [R*S,Cin,Cout] bf16 -> dx [N,H,W,Cin] bf16."""
    n, ho, wo, cout, lddy = _nhwc(dy, "dy")
    cin = wgt_t.shape[1]
    assert wgt_t.shape == (r * s, cin, cout) and wgt_t.dtype == torch.bfloat16 and wgt_t.is_contiguous()
    h, w = in_hw
    if out is None:
        out = torch.empty((n, h, w, cin), dtype=torch.bfloat16, device=dy.device)
    _, xh, xw, xc, lddx = _nhwc(out, "out")
    assert (xh, xw, xc) == (h, w, cin)
    with _lib.on_device(dy.device), _timed("dgrad", flops or 2.0 * n * ho * wo * cout * cin * r * s,
                                               (n, h, w, cin, ho, wo, cout, r, stride, dil)):
        _lib.check(_lib.lib().aadg_conv_dgrad_bf16(dy.data_ptr(), n, ho, wo, cout, lddy, wgt_t.data_ptr(), cin, r, s,
                                                   stride, pad, dil, out.data_ptr(), h, w, lddx, 0, int(accumulate),
                                                   _lib.stream_ptr()))
    return out


def wgrad(x, dy, r, s, stride, pad, dil, out=None, flops=None):
    """x [N,H,W,Cin], dy [N,Ho,Wo,Cout] bf16 -> dw [R*S,Cout,Cin] fp32 (accumulated into `out`)."""
    n, h, w, cin, ldx = _nhwc(x, "x")
    _, ho, wo, cout, lddy = _nhwc(dy, "dy")
    if out is None:
        out = torch.zeros((r * s, cout, cin), dtype=torch.float32, device=x.device)
    assert out.shape == (r * s, cout, cin) and out.dtype == torch.float32 and out.is_contiguous()
    with _lib.on_device(x.device), _timed("wgrad", flops or 2.0 * n * ho * wo * cout * cin * r * s,
                                              (n, h, w, cin, ho, wo, cout, r, stride, dil)):
        _lib.check(_lib.lib().aadg_conv_wgrad_bf16(x.data_ptr(), n, h, w, cin, ldx, dy.data_ptr(), ho, wo, cout, lddy,
                                                   r, s, stride, pad, dil, out.data_ptr(), _lib.stream_ptr()))
    return out


def fprop_windows(x, wgt, r, s, ho, wo, out=None, stats=None, flops=None):
    """valid stride-1 convolution over OVERLAPPING channel windows (aadg_conv_fprop_windows_bf16): x is a bf16 view
    [N,H,W,cin] whose pixel stride x.stride(2) may be smaller than cin (torch.as_strided over the space-to-depth stem
    buffer); wgt bf16 [r*s, cout, cin] -> y bf16 [N,ho,wo,cout]."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4):
        raise RuntimeError("x must be a CUDA bf16 [N,H,W,C] view (no CPU path)")
    n, h, w, cin = x.shape
    ldx = x.stride(2)
    assert x.stride(3) == 1 and x.stride(1) == w * ldx and x.stride(0) == h * w * ldx
    cout = wgt.shape[1]
    assert wgt.shape == (r * s, cout, cin) and wgt.dtype == torch.bfloat16 and wgt.is_contiguous()
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device=x.device)
    _, oh, ow, oc, ldy = _nhwc(out, "out")
    assert (oh, ow, oc) == (ho, wo, cout)
    ssum, ssq = stats if stats is not None else (None, None)
    with _lib.on_device(x.device), _timed("fprop", flops or 2.0 * n * ho * wo * cout * cin * r * s,
                                              (n, h, w, cin, ho, wo, cout, r, 1, 1)):
        _lib.check(_lib.lib().aadg_conv_fprop_windows_bf16(x.data_ptr(), n, h, w, cin, ldx, wgt.data_ptr(), cout, r, s,
                                                           out.data_ptr(), ho, wo, ldy,
                                                           ssum.data_ptr() if ssum is not None else None,
                                                           ssq.data_ptr() if ssq is not None else None, _lib.stream_ptr()))
    return out


def wgrad_windows(x, dy, r, s, out=None, flops=None):
    """weight gradient of fprop_windows: dw fp32 [r*s, cout, cin] (accumulated into `out`)."""
    n, h, w, cin = x.shape
    ldx = x.stride(2)
    _, ho, wo, cout, lddy = _nhwc(dy, "dy")
    if out is None:
        out = torch.zeros((r * s, cout, cin), dtype=torch.float32, device=x.device)
    assert out.shape == (r * s, cout, cin) and out.dtype == torch.float32 and out.is_contiguous()
    with _lib.on_device(x.device), _timed("wgrad", flops or 2.0 * n * ho * wo * cout * cin * r * s,
                                              (n, h, w, cin, ho, wo, cout, r, 1, 1)):
        _lib.check(_lib.lib().aadg_conv_wgrad_windows_bf16(x.data_ptr(), n, h, w, cin, ldx, dy.data_ptr(), ho, wo, cout,
                                                           lddy, r, s, out.data_ptr(), _lib.stream_ptr()))
    return out
