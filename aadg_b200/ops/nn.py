"""Thin torch-facing calls into the layer kernels (csrc/nn_elem.cu, csrc/nn_loss.cu).  All tensors
are CUDA; activations bf16 `[N,H,W,C]` views whose channel stride may exceed C.  No CPU path."""
import torch

from .. import _lib

BF16 = torch.bfloat16


def _ld(t):
    assert t.is_cuda and t.dtype == BF16 and t.stride(-1) == 1, "bf16 CUDA channel-last tensor expected"
    return t.stride(-2)


def _pix(t):
    n = 1
    for s in t.shape[:-1]:
        n *= s
    return n


def _call(name, *args):
    _lib.check(getattr(_lib.lib(), name)(*args, _lib.stream_ptr()))


def p(t):
    return None if t is None else t.data_ptr()


def bn_stats(x, sum_, sumsq):
    _call("aadg_bn_stats", p(x), _pix(x), x.shape[-1], _ld(x), p(sum_), p(sumsq))


def bn_finalize(sum_, sumsq, gamma, beta, count, eps, momentum, mean, invstd, scale, shift, run_mean, run_var,
                reset_sums=False):
    _call("aadg_bn_finalize", p(sum_), p(sumsq), p(gamma), p(beta), gamma.numel(), float(count), eps, momentum,
          p(mean), p(invstd), p(scale), p(shift), p(run_mean), p(run_var), int(reset_sums))


def _seed(dropout_seed):
    """(flag bits, value) of a dropout seed: None, an int (passed by value), or an int64 CUDA tensor holding it (flag
    64: the kernel reads the seed from device memory, so a captured CUDA graph follows the host's seed updates)"""
    if dropout_seed is None:
        return 0, 0
    if isinstance(dropout_seed, torch.Tensor):
        assert dropout_seed.is_cuda and dropout_seed.dtype == torch.int64 and dropout_seed.numel() == 1
        return 2 | 64, dropout_seed.data_ptr()
    return 2, int(dropout_seed)


def bn_apply(x, scale, shift, y, res=None, relu=True, dropout_seed=None, relu_bits=None, relu6=False):
    sf, sv = _seed(dropout_seed)
    flags = (1 if relu else 0) | sf | (32 if (relu and relu6) else 0)
    _call("aadg_bn_apply", p(x), _ld(x), p(scale), p(shift), p(res), _ld(res) if res is not None else 0, p(y), _ld(y),
          _pix(x), x.shape[-1], flags, sv, p(relu_bits))


def _bn_bwd_flags(y, relu, dropout_seed, grads_zeroed, relu6):
    bits = y is not None and y.dtype == torch.uint8        # relu bit mask written by bn_apply(relu_bits=...)
    sf, sv = _seed(dropout_seed)
    flags = (1 if relu else 0) | sf | (4 if (relu and y is None) else 0) | \
        (8 if bits else 0) | (16 if grads_zeroed else 0) | (32 if (relu and relu6) else 0)
    return bits, flags, sv


def bn_backward(dy, x, y, mean, invstd, gamma, dgamma, dbeta, dx, relu=True, dropout_seed=None, dres=None,
                dres_accumulate=False, shift=None, dy2=None, grads_zeroed=False, relu6=False):
    """y=None with relu=True recomputes the ReLU mask from x and the forward `shift` (no residual case).
    dy2: a second gradient tensor added to dy on load."""
    bits, flags, sv = _bn_bwd_flags(y, relu, dropout_seed, grads_zeroed, relu6)
    tail = (p(x), _ld(x), p(y), _ld(y) if (y is not None and not bits) else 0, p(mean), p(invstd),
            p(gamma), p(shift), _pix(x), x.shape[-1], flags, sv, p(dgamma), p(dbeta), p(dx), _ld(dx),
            p(dres), _ld(dres) if dres is not None else 0, int(dres_accumulate))
    if dy2 is None:
        _call("aadg_bn_backward", p(dy), _ld(dy), *tail)
    else:
        assert dy2.shape == dy.shape
        _call("aadg_bn_backward2", p(dy), _ld(dy), p(dy2), _ld(dy2), *tail)


def bn_backward_reduce(dy, x, y, mean, invstd, gamma, sums, relu=True, dropout_seed=None, shift=None, dy2=None,
                       relu6=False):
    """first half of the backward (SyncBN): sums fp32 [2,C] += (sum g*xhat, sum g) of this rank's pixels"""
    bits, flags, sv = _bn_bwd_flags(y, relu, dropout_seed, True, relu6)
    assert sums.dtype == torch.float32 and sums.shape == (2, x.shape[-1]) and sums.is_contiguous()
    _call("aadg_bn_backward_reduce", p(dy), _ld(dy), p(dy2), _ld(dy2) if dy2 is not None else 0, p(x), _ld(x), p(y),
          _ld(y) if (y is not None and not bits) else 0, p(mean), p(invstd), p(gamma), p(shift), _pix(x), x.shape[-1],
          flags, sv, p(sums[0]), p(sums[1]))


def bn_backward_apply(dy, x, y, mean, invstd, gamma, sums, count, dx, relu=True, dropout_seed=None, dres=None,
                      dres_accumulate=False, shift=None, dy2=None, relu6=False):
    """second half: dx (and dres) from the GLOBAL sums [2,C] and the global pixel count"""
    bits, flags, sv = _bn_bwd_flags(y, relu, dropout_seed, True, relu6)
    _call("aadg_bn_backward_apply", p(dy), _ld(dy), p(dy2), _ld(dy2) if dy2 is not None else 0, p(x), _ld(x), p(y),
          _ld(y) if (y is not None and not bits) else 0, p(mean), p(invstd), p(gamma), p(shift), _pix(x), x.shape[-1],
          flags, sv, p(sums[0]), p(sums[1]), 1.0 / float(count), p(dx), _ld(dx), p(dres),
          _ld(dres) if dres is not None else 0, int(dres_accumulate))


def add_(a, b):
    _call("aadg_add_bf16", p(a), _ld(a), p(b), _ld(b), _pix(a), a.shape[-1])


def maxpool_fwd(x):
    n, h, w, c = x.shape
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    y = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    arg = torch.empty((n, ho, wo, c), dtype=torch.uint8, device=x.device)
    _call("aadg_maxpool3x3s2_fwd", p(x), n, h, w, c, p(y), p(arg))
    return y, arg


def maxpool_bwd(dy, arg, in_shape):
    n, h, w, c = in_shape
    dx = torch.empty(in_shape, dtype=BF16, device=dy.device)
    _call("aadg_maxpool3x3s2_bwd", p(dy), p(arg), n, h, w, c, p(dx))
    return dx


def upsample_fwd(x, y):
    n, h, w, c = x.shape
    _call("aadg_upsample_bilinear_fwd", p(x), n, h, w, c, _ld(x), p(y), y.shape[1], y.shape[2], _ld(y))


def upsample_bwd(dy, dx):
    n, h, w, c = dx.shape
    _call("aadg_upsample_bilinear_bwd", p(dy), n, dy.shape[1], dy.shape[2], c, _ld(dy), p(dx), h, w, _ld(dx))


def global_sum(x, scale):
    n, h, w, c = x.shape
    out = torch.empty((n, c), dtype=torch.float32, device=x.device)
    _call("aadg_global_sum", p(x), n, h * w, c, _ld(x), p(out), float(scale))
    return out


def broadcast_pixels(v, y):
    n, h, w, c = y.shape
    _call("aadg_broadcast_pixels", p(v), n, c, p(y), h * w, _ld(y))


def broadcast_add_pixels(v, y, scale=1.0):
    """y bf16 [n,h,w,c] += scale * v fp32 [n,c]"""
    n, h, w, c = y.shape
    assert v.dtype == torch.float32 and v.shape == (n, c) and v.is_contiguous()
    _call("aadg_broadcast_add_pixels", p(v), n, c, p(y), h * w, _ld(y), float(scale))


def f32_to_bf16(x, scale=1.0):
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    _call("aadg_f32_to_bf16", p(x), p(y), x.numel(), float(scale))
    return y


def dwconv3x3(x, w, dil, y, backward_data=False, stride=1, accumulate=False):
    """forward: x [n,h,w,c] -> y [n,ho,wo,c]; backward_data: x is dy [n,ho,wo,c], y is dx [n,h,w,c];
    accumulate (stride 1): y += result."""
    if stride == 1:
        n, h, wd, c = x.shape
        _call("aadg_dwconv3x3", p(x), n, h, wd, c, _ld(x), p(w), dil, int(backward_data) | (2 if accumulate else 0), p(y), _ld(y))
        return
    assert not accumulate, "strided depthwise: accumulate is not implemented"
    big, small = (y, x) if backward_data else (x, y)
    n, h, wd, c = big.shape
    _call("aadg_dwconv3x3_strided", p(x), n, h, wd, c, _ld(x), p(w), dil, stride, int(backward_data), p(y), small.shape[1],
          small.shape[2], _ld(y))


def dwconv3x3_wgrad(x, dy, dil, dw, stride=1):
    n, h, wd, c = x.shape
    if stride == 1:
        _call("aadg_dwconv3x3_wgrad", p(x), n, h, wd, c, _ld(x), p(dy), _ld(dy), dil, p(dw))
    else:
        _call("aadg_dwconv3x3_strided_wgrad", p(x), n, h, wd, c, _ld(x), p(dy), dy.shape[1], dy.shape[2], _ld(dy), dil,
              stride, p(dw))


def stem_s2d(img):
    """fp32 NCHW image [n,3,h,w] -> (buffer, windows): the bf16 space-to-depth stem input [n, h/2+3, w/2+3, 16] with its
    zero border (aadg_stem_s2d) and the overlapping view [n, h/2+3, w/2+3, 64] (pixel stride 16) whose rows are the
    four-pixel windows the stem convolution contracts over (ops.conv.fprop_windows)."""
    n, c, h, w = img.shape
    assert c == 3 and img.dtype == torch.float32 and img.is_contiguous()
    hs, ws = h // 2 + 3, w // 2 + 3
    flat = torch.empty(n * hs * ws * 16 + 64, dtype=torch.bfloat16, device=img.device)
    flat[-64:].zero_()                      # the spare tail the last rows' windows end in (kept finite)
    _call("aadg_stem_s2d", p(img), n, h, w, p(flat))
    buf = flat[:n * hs * ws * 16].view(n, hs, ws, 16)
    windows = torch.as_strided(flat, (n, hs, ws, 64), (hs * ws * 16, ws * 16, 16, 1))
    return buf, windows


def im2col_stem(img, r, s, stride, pad, kp, row_pitch=None):
    """fp32 NCHW image -> bf16 patches [n,ho,wo,kp]; k = (r*S+s)*3+c, or r*row_pitch + s*3 + c when row_pitch is given."""
    img = img.contiguous()
    n, _, h, w = img.shape
    ho, wo = (h + 2 * pad - r) // stride + 1, (w + 2 * pad - s) // stride + 1
    col = torch.empty((n, ho, wo, kp), dtype=BF16, device=img.device)
    if row_pitch is None:
        _call("aadg_im2col_stem", p(img), n, h, w, r, s, stride, pad, kp, p(col))
    else:
        _call("aadg_im2col_stem_rows", p(img), n, h, w, r, s, stride, pad, row_pitch, kp, p(col))
    return col


def adam_step(params, grads, m, v, lr, beta1, beta2, eps, wd, step):
    _call("aadg_adam_step", p(params), p(grads), p(m), p(v), params.numel(), lr, beta1, beta2, eps, wd, step)


def adam_step_dev(params, grads, m, v, hyper, step):
    """Adam with its scalars in device memory: hyper float32 [6] = (lr, beta1, beta2, eps, weight_decay, grad_scale),
    step int64 [1] (1-based)"""
    assert hyper.dtype == torch.float32 and hyper.numel() == 6 and step.dtype == torch.int64 and step.numel() == 1
    _call("aadg_adam_step_dev", p(params), p(grads), p(m), p(v), params.numel(), p(hyper), p(step))


def weight_prep(master, wb, wbt, descs, n):
    _call("aadg_weight_prep", p(master), p(wb), p(wbt), p(descs), n)


def seg_head_fwd(a, w, b):
    n, h, wd, c = a.shape
    k = w.shape[0]
    z = torch.empty((n, h, wd, k), dtype=torch.float32, device=a.device)
    _call("aadg_seg_head_fwd", p(a), n * h * wd, c, _ld(a), p(w), p(b), k, p(z))
    return z


def seg_loss_fwd(z, target, thr, loss_sum, counts, logits_out=None):
    n, h, w, k = z.shape
    _call("aadg_seg_loss_fwd", p(z), n, h, w, k, p(target), target.shape[2], target.shape[3], float(thr),
          p(loss_sum), p(counts), p(logits_out))


def seg_loss_bwd(z, target, grad_scale):
    n, h, w, k = z.shape
    dz = torch.empty_like(z)
    _call("aadg_seg_loss_bwd", p(z), n, h, w, k, p(target), target.shape[2], target.shape[3], float(grad_scale), p(dz))
    return dz


def upsample_logits_bwd(dlogits, z_shape):
    """dlogits fp32 [n,k,H,W] -> dz fp32 [n,h,w,k] (transpose of the head's align_corners bilinear up-sampling)"""
    n, h, w, k = z_shape
    dlogits = dlogits.contiguous()
    assert dlogits.is_cuda and dlogits.dtype == torch.float32 and dlogits.shape[:2] == (n, k)
    dz = torch.empty((n, h, w, k), dtype=torch.float32, device=dlogits.device)
    _call("aadg_upsample_logits_bwd", p(dlogits), n, h, w, k, dlogits.shape[2], dlogits.shape[3], p(dz))
    return dz


def seg_head_bwd(dz, a, w, da, dw, db):
    n, h, wd, c = a.shape
    _call("aadg_seg_head_bwd", p(dz), p(a), n * h * wd, c, _ld(a), p(w), w.shape[0], p(da), _ld(da), p(dw), p(db))


def nearest2x_fwd(x, y):
    n, h, w, c = x.shape
    _call("aadg_upsample_nearest2x_fwd", p(x), n, h, w, c, _ld(x), p(y), _ld(y))


def nearest2x_bwd(dy, dx):
    n, h, w, c = dx.shape
    _call("aadg_upsample_nearest2x_bwd", p(dy), n, h, w, c, _ld(dy), p(dx), _ld(dx))


def copy_(x, y):
    _call("aadg_copy_bf16", p(x), _ld(x), p(y), _ld(y), _pix(x), x.shape[-1])


def seg_head3x3_fwd(a, w, b):
    n, h, wd, c = a.shape
    k = w.shape[0]
    z = torch.empty((n, h, wd, k), dtype=torch.float32, device=a.device)
    _call("aadg_seg_head3x3_fwd", p(a), n, h, wd, c, _ld(a), p(w), p(b), k, p(z))
    return z


def seg_head3x3_bwd(dz, a, w, da, dw, db):
    n, h, wd, c = a.shape
    _call("aadg_seg_head3x3_bwd", p(dz), p(a), n, h, wd, c, _ld(a), p(w), w.shape[0], p(da), _ld(da), p(dw), p(db))
