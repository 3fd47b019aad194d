"""Float tensor augmentation bank: torch-facing wrappers over aadg_f32_op / aadg_f32_op_backward (csrc/aug_f32.cu).
`apply` is the plain forward; `differentiable` is the same call behind a torch.autograd.Function whose backward is the
CUDA backward kernel (gradients to the image, the per-sample magnitude and the per-sample mask)."""
import torch

from .. import _lib

OPS = ["ShearX", "ShearY", "TranslateX", "TranslateY", "HorizontalFlip", "VerticalFlip", "Rotate", "Invert",
       "Solarize", "Posterize", "Gray", "Contrast", "AutoContrast", "Saturate", "Brightness", "Hue",
       "SamplePairing", "Equalize", "Sharpness"]
OP_ID = {n: i for i, n in enumerate(OPS)}


def apply(op, x, mag=None, mask=None, perm=None):
    """out = clamp(mask*op(x, mag) + (1-mask)*x, 0, 1).  x float32 [B,3,H,W] CUDA; mag/mask float32 [B] or None."""
    if not (isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
        raise RuntimeError("aadg_b200.ops.f32: x must be a CUDA float32 [B,3,H,W] tensor (no CPU path)")
    x = x.contiguous()
    b, _, h, w = x.shape
    dev = x.device

    def vec(t, dtype):
        if t is None:
            return None
        t = torch.as_tensor(t, dtype=dtype, device=dev).reshape(-1)
        return (t.expand(b) if t.numel() == 1 else t).contiguous()
    mag, mask, perm = vec(mag, torch.float32), vec(mask, torch.float32), vec(perm, torch.int32)
    out = torch.empty_like(x)
    L = _lib.lib()
    ws = _lib.workspace(L.aadg_f32_workspace_bytes(b), dev)
    with torch.cuda.device(dev):
        _lib.check(L.aadg_f32_op(OP_ID[op] if isinstance(op, str) else int(op), x.data_ptr(), b, h, w, _lib.ptr(mag),
                                 _lib.ptr(mask), _lib.ptr(perm), out.data_ptr(), _lib.ptr(ws), ws.numel(),
                                 _lib.stream_ptr()))
    return out


def _vec(t, b, dev, dtype):
    if t is None:
        return None
    t = torch.as_tensor(t, dtype=dtype, device=dev).reshape(-1)
    return (t.expand(b) if t.numel() == 1 else t).contiguous()


def backward(op, x, grad_out, mag=None, mask=None, perm=None):
    """(d/dx [B,3,H,W], d/dmag [B] or None, d/dmask [B] or None) of `apply(op, x, mag, mask, perm)` for the upstream
    gradient grad_out."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and x.shape[1] == 3):
        raise RuntimeError("aadg_b200.ops.f32: x must be a CUDA float32 [B,3,H,W] tensor (no CPU path)")
    x, grad_out = x.contiguous(), grad_out.contiguous().float()
    b, _, h, w = x.shape
    dev = x.device
    mag, mask, perm = _vec(mag, b, dev, torch.float32), _vec(mask, b, dev, torch.float32), _vec(perm, b, dev, torch.int32)
    gx = torch.empty_like(x)
    gmag = torch.empty(b, dtype=torch.float32, device=dev) if mag is not None else None
    gmask = torch.empty(b, dtype=torch.float32, device=dev) if mask is not None else None
    L = _lib.lib()
    ws = _lib.workspace(L.aadg_f32_workspace_bytes(b), dev)
    with _lib.on_device(dev):
        _lib.check(L.aadg_f32_op_backward(OP_ID[op] if isinstance(op, str) else int(op), x.data_ptr(), grad_out.data_ptr(),
                                          b, h, w, _lib.ptr(mag), _lib.ptr(mask), _lib.ptr(perm), gx.data_ptr(),
                                          _lib.ptr(gmag), _lib.ptr(gmask), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return gx, gmag, gmask


class _BankFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op, x, mag, mask, perm):
        ctx.op, ctx.perm = op, perm
        ctx.save_for_backward(x, mag, mask)
        return apply(op, x, mag, mask, perm)

    @staticmethod
    def backward(ctx, grad_out):
        x, mag, mask = ctx.saved_tensors
        gx, gmag, gmask = backward(ctx.op, x, grad_out, mag, mask, ctx.perm)
        return None, gx, gmag, gmask, None


def differentiable(op, x, mag=None, mask=None, perm=None):
    """`apply` with autograd: mag / mask float32 [B] tensors (or None) may require grad, and so may x."""
    b = x.shape[0]
    mag = None if mag is None else (mag.reshape(-1).expand(b) if mag.numel() == 1 else mag.reshape(-1))
    mask = None if mask is None else (mask.reshape(-1).expand(b) if mask.numel() == 1 else mask.reshape(-1))
    return _BankFunction.apply(op, x, mag, mask, perm)
