"""ORACLE (test infrastructure, never on the product path).

The torch oracle of `oracle/segnet_torch.py` with the ENGINE'S STORAGE POINTS made explicit: the same stock
fp32 torch layers, but every tensor the CUDA engine keeps in HBM as bf16 is rounded to bf16 at that point
(straight-through in backward), and the tensor-core convolutions see bf16-rounded weights.  It separates the two
things a whole-network comparison mixes:

  engine  vs  bf16-storage oracle   = implementation parity (same algorithm; residual = fp32 summation order and the
                                      bf16 rounding of the GRADIENT tensors, which autograd keeps in fp32 here)
  bf16-storage oracle  vs  fp32 oracle = the precision cost of bf16 activations, independent of any kernel

Because the rounding points coincide, the ReLU / ReLU6 masks of the two sides coincide (they are "teacher-forced" by
construction) and the backward comparison is free of mask flips.

Rounding points (aadg_b200/nn/network.py): the input image (the stem input is written as bf16), every convolution output
(`pre`), every ReLU / ReLU6 output, every batch-norm output that is stored without an activation (downsample
branches, MobileNetV2 linear bottlenecks without identity), the block output of a MobileNetV2 identity block, the
pooled ASPP vector, both bilinear up-samplings of the decoder.  Not rounded: batch-norm outputs that feed a residual
add (the engine adds in fp32 and rounds once after the ReLU), the segmentation head and the loss (fp32).
Weights: every dense convolution uses round_bf16(w) (depthwise and head weights stay fp32, as in the engine).
"""
import torch
import torch.nn as nn
import torchvision


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def rb(x):
    return _RoundSTE.apply(x)


def _round_out(_m, _inp, out):
    return rb(out)


def _conv_forward_bf16_weight(m):
    def forward(x):
        return m._conv_forward(x, rb(m.weight), m.bias)
    return forward


def install(model, head_name="segmentation_head", keep_fp32=(), round_weights=True, round_activations=True,
            round_input=True):
    """turn a DeepLabV3PlusTorch / UnetTorch into its bf16-storage twin (in place); returns the model.

    The defaults are the engine's storage points.  The keyword arguments exist for the precision-attribution study
    (`tests/tools/precision_attribution.py`): modules whose qualified name starts with one of `keep_fp32` keep fp32 storage
    and fp32 weights; `round_weights` / `round_activations` / `round_input` switch one class of rounding points off."""
    from torchvision.models.mobilenetv2 import InvertedResidual
    from torchvision.models.resnet import BasicBlock, Bottleneck
    keep = tuple(keep_fp32)
    kept = {mod for name, mod in model.named_modules() if any(name == k or name.startswith(k + ".") for k in keep)}
    unrounded_bn = set()
    for mod in model.modules():
        if isinstance(mod, Bottleneck):
            unrounded_bn.add(mod.bn3)
        elif isinstance(mod, BasicBlock):
            unrounded_bn.add(mod.bn2)
        elif isinstance(mod, InvertedResidual):
            if mod.use_res_connect:
                unrounded_bn.add(mod.conv[-1])
                if round_activations and mod not in kept:
                    mod.register_forward_hook(_round_out)
    head = getattr(model, head_name)
    head_mods = set(head.modules())
    for mod in model.modules():
        if mod in head_mods or mod in kept:
            continue
        if isinstance(mod, nn.Conv2d):
            if mod.groups == 1 and round_weights:
                mod.forward = _conv_forward_bf16_weight(mod)
            if round_activations:
                mod.register_forward_hook(_round_out)
        elif not round_activations:
            continue
        elif isinstance(mod, (nn.ReLU, nn.ReLU6)):
            mod.inplace = False
            mod.register_forward_hook(_round_out)
        elif isinstance(mod, nn.BatchNorm2d) and mod not in unrounded_bn:
            mod.register_forward_hook(_round_out)         # harmless before a ReLU: rounding commutes with it
        elif isinstance(mod, (nn.UpsamplingBilinear2d, nn.AdaptiveAvgPool2d)) and mod is not getattr(model, "pool", None):
            mod.register_forward_hook(_round_out)
    if round_input:
        model.register_forward_pre_hook(lambda _m, args: (rb(args[0]),) + tuple(args[1:]))
    return model
