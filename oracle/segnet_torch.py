"""ORACLE (test infrastructure, never on the product path).  **Parity unpinned** for module order.

Plain torch.nn restatement of `segmentation_models_pytorch==0.2.0` DeepLabV3Plus as the reference
builds it (models/__init__.py:17-23: smp.DeepLabV3Plus(encoder_name, encoder_weights, in_channels=3,
classes, aux_params=dict(pooling='avg'))) with the patched classification head of models/heads.py:14-25
(AdaptiveAvgPool2d(1) + flatten; `model(x) -> (masks, pooled_features)`).  smp is not installed here and
not vendored under /root/reference, so the module structure follows its published source
(smp/deeplabv3/{model,decoder}.py, smp/encoders/{resnet,_utils}.py, smp/base/{model,heads}.py) and keeps
its state_dict key names; every layer is a stock torch layer, which pins the numerics of each op.
Encoders: torchvision mobilenet_v2 (the reference's only reachable backbone, models/__init__.py:16) and
resnet18/34/50 without the classifier, stage 5 dilated (output stride 16).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision


class SeparableConv2d(nn.Sequential):
    def __init__(self, cin, cout, k, stride=1, padding=0, dilation=1, bias=True):
        super().__init__(nn.Conv2d(cin, cin, k, stride=stride, padding=padding, dilation=dilation, groups=cin, bias=False),
                         nn.Conv2d(cin, cout, 1, bias=bias))


class ASPPConv(nn.Sequential):
    def __init__(self, cin, cout, rate):
        super().__init__(nn.Conv2d(cin, cout, 3, padding=rate, dilation=rate, bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class ASPPSeparableConv(nn.Sequential):
    def __init__(self, cin, cout, rate):
        super().__init__(SeparableConv2d(cin, cout, 3, padding=rate, dilation=rate, bias=False), nn.BatchNorm2d(cout),
                         nn.ReLU())


class ASPPPooling(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(nn.AdaptiveAvgPool2d(1), nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU())

    def forward(self, x):
        size = x.shape[-2:]
        for m in self:
            x = m(x)
        return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


class ASPP(nn.Module):
    def __init__(self, cin, cout, rates, separable=True):
        super().__init__()
        mods = [nn.Sequential(nn.Conv2d(cin, cout, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU())]
        conv = ASPPSeparableConv if separable else ASPPConv
        mods += [conv(cin, cout, r) for r in rates]
        mods.append(ASPPPooling(cin, cout))
        self.convs = nn.ModuleList(mods)
        self.project = nn.Sequential(nn.Conv2d(5 * cout, cout, 1, bias=False), nn.BatchNorm2d(cout), nn.ReLU(),
                                     nn.Dropout(0.5))

    def forward(self, x):
        return self.project(torch.cat([c(x) for c in self.convs], dim=1))


class Decoder(nn.Module):
    def __init__(self, enc_channels, oc=256, rates=(12, 24, 36)):
        super().__init__()
        self.aspp = nn.Sequential(ASPP(enc_channels[-1], oc, rates, separable=True),
                                  SeparableConv2d(oc, oc, 3, padding=1, bias=False), nn.BatchNorm2d(oc), nn.ReLU())
        self.up = nn.UpsamplingBilinear2d(scale_factor=4)
        self.block1 = nn.Sequential(nn.Conv2d(enc_channels[-4], 48, 1, bias=False), nn.BatchNorm2d(48), nn.ReLU())
        self.block2 = nn.Sequential(SeparableConv2d(48 + oc, oc, 3, padding=1, bias=False), nn.BatchNorm2d(oc), nn.ReLU())

    def forward(self, *features):
        a = self.up(self.aspp(features[-1]))
        h = self.block1(features[-4])
        return self.block2(torch.cat([a, h], dim=1))


class ResNetEncoder(nn.Module):
    """smp ResNetEncoder: torchvision ResNet minus fc/avgpool; keys conv1, bn1, layer1..4."""

    def __init__(self, name, dilated=True):
        super().__init__()
        net = getattr(torchvision.models, name)(weights=None)
        self.conv1, self.bn1, self.relu, self.maxpool = net.conv1, net.bn1, net.relu, net.maxpool
        self.layer1, self.layer2, self.layer3, self.layer4 = net.layer1, net.layer2, net.layer3, net.layer4
        # make_dilated(stage_list=[5], dilation_list=[2]) -> replace_strides_with_dilation(layer4, 2)
        for m in (self.layer4.modules() if dilated else []):
            if isinstance(m, nn.Conv2d):
                m.stride = (1, 1)
                m.dilation = (2, 2)
                kh, kw = m.kernel_size
                m.padding = ((kh // 2) * 2, (kw // 2) * 2)
        exp = 4 if name == "resnet50" else 1
        self.out_channels = (3, 64, 64 * exp, 128 * exp, 256 * exp, 512 * exp)

    def forward(self, x):
        f0 = x
        f1 = self.relu(self.bn1(self.conv1(x)))
        f2 = self.layer1(self.maxpool(f1))
        f3 = self.layer2(f2)
        f4 = self.layer3(f3)
        f5 = self.layer4(f4)
        return [f0, f1, f2, f3, f4, f5]


class MobileNetV2Encoder(nn.Module):
    """smp MobileNetV2Encoder: torchvision mobilenet_v2 `features` (keys features.0 .. features.18), stages
    [:2], [2:4], [4:7], [7:14], [14:]; make_dilated(stage_list=[5], dilation_list=[2]) on the last stage."""

    def __init__(self, dilated=True):
        super().__init__()
        self.features = torchvision.models.mobilenet_v2(weights=None).features
        for m in (self.features[14:].modules() if dilated else []):
            if isinstance(m, nn.Conv2d):
                m.stride = (1, 1)
                m.dilation = (2, 2)
                kh, kw = m.kernel_size
                m.padding = ((kh // 2) * 2, (kw // 2) * 2)
        self.out_channels = (3, 16, 24, 32, 96, 1280)

    def forward(self, x):
        feats = [x]
        for lo, hi in ((0, 2), (2, 4), (4, 7), (7, 14), (14, 19)):
            x = self.features[lo:hi](x)
            feats.append(x)
        return feats


class DeepLabV3PlusTorch(nn.Module):
    def __init__(self, encoder_name="resnet50", classes=2):
        super().__init__()
        self.encoder = MobileNetV2Encoder() if encoder_name == "mobilenet_v2" else ResNetEncoder(encoder_name)
        self.decoder = Decoder(self.encoder.out_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(256, classes, 1), nn.UpsamplingBilinear2d(scale_factor=4),
                                               nn.Identity())
        self.pool = nn.AdaptiveAvgPool2d(1)

    def forward(self, x):
        feats = self.encoder(x)
        masks = self.segmentation_head(self.decoder(*feats))
        return masks, torch.flatten(self.pool(feats[-1]), 1)


class Conv2dReLU(nn.Sequential):
    def __init__(self, cin, cout):
        super().__init__(nn.Conv2d(cin, cout, 3, padding=1, bias=False), nn.BatchNorm2d(cout), nn.ReLU())


class UnetDecoderBlock(nn.Module):
    def __init__(self, cin, cskip, cout):
        super().__init__()
        self.conv1 = Conv2dReLU(cin + cskip, cout)
        self.conv2 = Conv2dReLU(cout, cout)

    def forward(self, x, skip=None):
        x = F.interpolate(x, scale_factor=2, mode="nearest")
        if skip is not None:
            x = torch.cat([x, skip], dim=1)
        return self.conv2(self.conv1(x))


class UnetDecoder(nn.Module):
    def __init__(self, enc_channels, dec_channels=(256, 128, 64, 32, 16)):
        super().__init__()
        enc = list(enc_channels[1:])[::-1]
        in_ch = [enc[0]] + list(dec_channels[:-1])
        skip_ch = enc[1:] + [0]
        self.blocks = nn.ModuleList([UnetDecoderBlock(i, s, o) for i, s, o in zip(in_ch, skip_ch, dec_channels)])

    def forward(self, *features):
        feats = features[1:][::-1]
        x, skips = feats[0], feats[1:]
        for i, blk in enumerate(self.blocks):
            x = blk(x, skips[i] if i < len(skips) else None)
        return x


class UnetTorch(nn.Module):
    """smp.Unet(encoder_name, classes) with the patched classification head (pool + flatten)."""

    def __init__(self, encoder_name="resnet34", classes=1):
        super().__init__()
        self.encoder = MobileNetV2Encoder(dilated=False) if encoder_name == "mobilenet_v2" else \
            ResNetEncoder(encoder_name, dilated=False)
        self.decoder = UnetDecoder(self.encoder.out_channels)
        self.segmentation_head = nn.Sequential(nn.Conv2d(16, classes, 3, padding=1), nn.Identity(), nn.Identity())
        self.pool = nn.AdaptiveAvgPool2d(1)

    def forward(self, x):
        feats = self.encoder(x)
        return self.segmentation_head(self.decoder(*feats)), torch.flatten(self.pool(feats[-1]), 1)


def f1_samplewise(prob, target, thr=0.5):
    """torchmetrics 0.4.1 F1(num_classes=2, average=None, mdmc_average='samplewise')[1] for one channel:
    mean over samples of the hard Dice of (prob > thr) vs target."""
    pred = (prob > thr).flatten(1).double()
    t = (target > 0.5).flatten(1).double()
    tp = (pred * t).sum(1)
    den = pred.sum(1) + t.sum(1)
    return torch.where(den > 0, 2 * tp / den.clamp(min=1), torch.zeros_like(den)).mean()
