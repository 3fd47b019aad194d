"""Validation metrics on the GPU (csrc/hd95.cu): the reference computes medpy.metric.binary.hd95 per image and
class on the CPU inside validate() (search_dg.py:246-260).  No CPU path."""
import torch

from .. import _lib


def surface_distance_percentile(result, reference, percentile=95.0):
    """result, reference: CUDA bool/uint8 tensors [..., H, W] (non-zero = foreground).  Returns (values float64 [...],
    status int32 [...]): status 0 ok, 1 = result empty, 2 = reference empty (value NaN)."""
    if not (result.is_cuda and reference.is_cuda):
        raise RuntimeError("hd95: CUDA tensors expected (no CPU path)")
    if result.shape != reference.shape or result.dim() < 2:
        raise ValueError("hd95: result and reference must have the same [..., H, W] shape")
    lead = result.shape[:-2]
    h, w = result.shape[-2:]
    r = (result != 0).to(torch.uint8).reshape(-1, h, w).contiguous()
    t = (reference != 0).to(torch.uint8).reshape(-1, h, w).contiguous()
    n = r.shape[0]
    out = torch.empty(n, dtype=torch.float64, device=r.device)
    status = torch.empty(n, dtype=torch.int32, device=r.device)
    lib = _lib.lib()
    nbytes = lib.aadg_hd95_workspace_bytes(n, h, w)
    ws = _lib.workspace(nbytes, r.device)
    with torch.cuda.device(r.device):
        _lib.check(lib.aadg_hd95(r.data_ptr(), t.data_ptr(), n, h, w, float(percentile), out.data_ptr(),
                                 status.data_ptr(), ws.data_ptr(), ws.numel(), _lib.stream_ptr()))
    return out.reshape(lead), status.reshape(lead)


def hd95(result, reference):
    """medpy.metric.binary.hd95(result, reference) for one pair or a batch; raises RuntimeError like medpy when a mask
    holds no object."""
    v, st = surface_distance_percentile(result, reference, 95.0)
    bad = int((st != 0).sum())
    if bad:
        first = int(st.flatten()[st.flatten() != 0][0])
        raise RuntimeError("The %s supplied array does not contain any binary object." % ("first" if first == 1 else "second"))
    return v


def validation_hd95(seg_hard, mask_gt, empty_value=100.0):
    """The per-batch HD95 of the reference's validate(): seg_hard, mask_gt [N, C, H, W]; an empty prediction scores
    `empty_value` (search_dg.py:249-252,255-258); returns float64 [C] = mean over the batch (what the AverageMeter
    is updated with, total_*_hd / input.size(0))."""
    v, st = surface_distance_percentile(seg_hard, mask_gt, 95.0)
    if bool((st == 2).any()):
        raise RuntimeError("The second supplied array does not contain any binary object.")
    v = torch.where(st == 1, torch.full_like(v, empty_value), v)
    return v.mean(0)
