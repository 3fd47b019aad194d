"""validate() of the reference (search_dg.py:219-275 / train_dg.py) on the CUDA engine: per batch the model runs in
eval mode, predictions are thresholded at 0.75, the samplewise Dice (torchmetrics F1, class 1) comes from the fused
TP/FP/FN counts and HD95 from the GPU kernels -- no mask ever travels to the host.  Meters are updated like the
reference's AverageMeters (value of the batch, weight = batch size)."""
import math

import torch

from ..nn.network import dice_from_counts
from ..ops import metrics


class AverageMeter:
    """utils.AverageMeter of the reference: running weighted mean"""

    def __init__(self):
        self.sum, self.count = 0.0, 0

    def update(self, val, n=1):
        self.sum += float(val) * n
        self.count += n

    @property
    def avg(self):
        return self.sum / max(self.count, 1)


def validate(model, batches, threshold=0.75, empty_hd=100.0):
    """batches: iterable of (input float32 [N,3,H,W], mask_gt float32 [N,C,H,W]) CUDA tensors (what
    test_dg_collate_fn + .cuda() hand to the reference's loop).  Returns dict(dsc=[C], hd=[C]) of meter averages; for
    the optic disc data C = 2 and the reference logs them as (cup, disc)."""
    was_training = model.training
    model.eval()
    dsc, hd = None, None
    try:
        for x, mask_gt in batches:
            out = model.evaluate_batch(x, mask_gt, thr=threshold)
            n, c = mask_gt.shape[:2]
            if dsc is None:
                dsc, hd = [AverageMeter() for _ in range(c)], [AverageMeter() for _ in range(c)]
            d = dice_from_counts(out["counts"]).cpu()
            # sigmoid(z) > t  <=>  z > log(t / (1 - t))
            seg_hard = out["logits"] > math.log(threshold / (1.0 - threshold))
            h = metrics.validation_hd95(seg_hard, mask_gt > 0.5, empty_value=empty_hd).cpu()
            for k in range(c):
                dsc[k].update(d[k].item(), n)
                hd[k].update(h[k].item(), n)
    finally:
        model.train(was_training)
    return dict(dsc=[m.avg for m in dsc], hd=[m.avg for m in hd])
