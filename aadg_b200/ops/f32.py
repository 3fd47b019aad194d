"""Float tensor augmentation bank: torch-facing wrapper over aadg_f32_op (csrc/aug_f32.cu)."""
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
