"""The reference's float operation classes (data/operations.py:16-399) over the CUDA tensor bank.

Same class names, constructor arguments, `magnitude` / `probability` properties and `forward(input)`:
    training: out = clamp(mask*op(x, mag) + (1-mask)*x), mask ~ RelaxedBernoulli(temperature, p)
    eval:     mask ~ Bernoulli(p); the op replaces the selected rows
Differentiable like the reference's: in training mode the output carries a graph to the image, to `_magnitude` and
(through the RelaxedBernoulli sample) to `_probability`; the backward pass is one CUDA kernel per op
(aadg_f32_op_backward) with torch autograd's semantics through the reference code, straight-through estimators
included (functional.py:21-46).  `forward` does not modify its input (the reference's eval branch writes into it in
place)."""
import torch
from torch import nn
from torch.distributions import RelaxedBernoulli, Bernoulli

from ..ops import f32 as _f32


class _Operation(nn.Module):
    op_name = None

    def __init__(self, initial_magnitude=None, initial_probability=0.5, magnitude_range=None, probability_range=None,
                 temperature=0.1, flip_magnitude=False, magnitude_scale=1, debug=False):
        super().__init__()
        self.magnitude_range = None
        if initial_magnitude is None:
            self._magnitude = None
        elif magnitude_range is None:
            self.register_buffer("_magnitude", torch.empty(1).fill_(initial_magnitude))
        else:
            self._magnitude = nn.Parameter(torch.empty(1).fill_(initial_magnitude))
            assert 0 <= magnitude_range[0] < magnitude_range[1] <= 1
            self.magnitude_range = magnitude_range
        self.probability_range = probability_range
        if probability_range is None:
            self.register_buffer("_probability", torch.empty(1).fill_(initial_probability))
        else:
            assert 0 <= initial_probability <= 1 and 0 <= probability_range[0] < probability_range[1] <= 1
            self._probability = nn.Parameter(torch.empty(1).fill_(initial_probability))
        assert 0 < temperature and 0 < magnitude_scale
        self.register_buffer("temperature", torch.empty(1).fill_(temperature))
        self.flip_magnitude = flip_magnitude and (self._magnitude is not None)
        self.magnitude_scale = magnitude_scale
        self.debug = debug

    @property
    def magnitude(self):
        if self._magnitude is None:
            return None
        mag = self._magnitude
        if self.magnitude_range is not None:
            mag = mag.clamp(*self.magnitude_range)
        return mag * self.magnitude_scale

    @property
    def probability(self):
        if self.probability_range is None:
            return self._probability
        return self._probability.clamp(*self.probability_range)

    def get_mask(self, batch_size=None):
        size = (batch_size, 1, 1)
        if self.training:
            return RelaxedBernoulli(self.temperature, self.probability).rsample(size)
        return Bernoulli(self.probability).sample(size)

    def forward(self, input):
        b = input.size(0)
        mask = self.get_mask(b).reshape(b)
        mag = self.magnitude
        if mag is not None:
            mag = mag.reshape(-1).expand(b) if mag.numel() == 1 else mag
            if self.flip_magnitude:
                mag = torch.randint(2, (b,), dtype=torch.float32, device=input.device).mul_(2).sub_(1) * mag
        perm = torch.randperm(b, device=input.device) if self.op_name == "SamplePairing" else None
        if self.training and torch.is_grad_enabled():
            return _f32.differentiable(self.op_name, input, mag, mask, perm)
        return _f32.apply(self.op_name, input, None if mag is None else mag.detach(), mask.detach(), perm)


def _make(name, has_mag=True, flip=False, scale=1, default_mag=0.5):
    if has_mag:
        def __init__(self, initial_magnitude=default_mag, initial_probability=0.5, magnitude_range=(0, 1),
                     probability_range=(0, 1), temperature=0.1, magnitude_scale=scale, debug=False):
            _Operation.__init__(self, initial_magnitude, initial_probability, magnitude_range, probability_range,
                                temperature, flip_magnitude=flip, magnitude_scale=magnitude_scale, debug=debug)
    else:
        def __init__(self, initial_probability=0.5, probability_range=(0, 1), temperature=0.1, debug=False):
            _Operation.__init__(self, None, initial_probability, None, probability_range, temperature, debug=debug)
    return type(name, (_Operation,), {"__init__": __init__, "op_name": name})


ShearX = _make("ShearX", flip=True, scale=0.3)
ShearY = _make("ShearY", flip=True, scale=0.3)
TranslateX = _make("TranslateX", flip=True, scale=0.45)
TranslateY = _make("TranslateY", flip=True, scale=0.45)
HorizontalFlip = _make("HorizontalFlip", has_mag=False)
VerticalFlip = _make("VerticalFlip", has_mag=False)
Rotate = _make("Rotate", flip=True, scale=30)
Invert = _make("Invert", has_mag=False)
Solarize = _make("Solarize")
Posterize = _make("Posterize")
Gray = _make("Gray", has_mag=False)
Contrast = _make("Contrast")
AutoContrast = _make("AutoContrast", has_mag=False)
Saturate = _make("Saturate")
Brightness = _make("Brightness")
Hue = _make("Hue", scale=2)
SamplePairing = _make("SamplePairing")
Equalize = _make("Equalize", has_mag=False)
Sharpness = _make("Sharpness", flip=True)     # the reference flips its sign too (operations.py:389-399)

__all__ = _f32.OPS
