"""Momentum feature discriminator (reference models/discriminator.py:20-59): two tiny MLPs, the EMA
twin's 128-d hidden layer is the Sinkhorn point cloud.  Two small GEMMs per step: left to torch/cuBLAS."""
import torch
import torch.nn as nn


class MomentumFeatureDiscriminator(nn.Module):
    def __init__(self, num_classes, in_channels, m=0.999):
        super().__init__()
        self.m = m
        self.dis = nn.Sequential(nn.Linear(in_channels, 128), nn.LeakyReLU(0.2, inplace=True))
        self.fc = nn.Linear(128, num_classes)
        self.mom_dis = nn.Sequential(nn.Linear(in_channels, 128), nn.LeakyReLU(0.2, inplace=True))
        self.mom_fc = nn.Linear(128, num_classes)

    def _pairs(self):
        return list(zip(self.dis.parameters(), self.mom_dis.parameters())) + \
            list(zip(self.fc.parameters(), self.mom_fc.parameters()))

    @torch.no_grad()
    def momentum_update(self):
        for q, k in self._pairs():
            k.mul_(self.m).add_(q, alpha=1.0 - self.m)

    @torch.no_grad()
    def synchronize_parameters(self):
        for q, k in self._pairs():
            k.copy_(q)

    def forward(self, x, momentum=False, return_feature=False):
        if momentum:
            with torch.no_grad():
                fe = self.mom_dis(x)
                out = self.mom_fc(fe)
        else:
            fe = self.dis(x)
            out = self.fc(fe)
        return (out, fe) if return_feature else out
