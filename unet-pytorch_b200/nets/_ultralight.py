"""Shared parameter containers of the UltraLightweightUnet family (reference: nets/UltraLightweightUnet.py:6-54,
nets/UltraLightweightUnet_large.py:6-52, nets/UltraLightweightUnet_large_optimized.py:5-48).  The three reference files
differ only in stage widths, the floor of the block's mid channels, whether SE blocks follow the encoder stages and
the bridge Dropout2d probability; the module tree and state_dict keys below are the reference's."""
import torch.nn as nn

from ..graph import ULU_VARIANTS, UltraLightUnetEngine
from ._function import EngineModuleMixin


def _container_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is a parameter container here; call the network's forward (CUDA engine)")


class DepthwiseSeparableConv(nn.Module):
    """depthwise 3x3 (groups = channels, with bias) then pointwise 1x1."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1):
        super().__init__()
        if (kernel_size, stride, padding) != (3, 1, 1):
            raise NotImplementedError("the CUDA depthwise kernel is 3x3, stride 1, padding 1 (the only use in the reference)")
        self.depthwise = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1, groups=in_channels)
        self.pointwise = nn.Conv2d(in_channels, out_channels, kernel_size=1)

    forward = _container_forward


def light_conv_block(mid_min):
    class LightConvBlock(nn.Module):
        """conv1x1 -> BN -> ReLU -> DepthwiseSeparableConv -> BN -> ReLU, mid = max(mid_min, out // 2)."""

        def __init__(self, in_channels, out_channels):
            super().__init__()
            mid_channels = max(mid_min, out_channels // 2)
            self.conv = nn.Sequential(
                nn.Conv2d(in_channels, mid_channels, kernel_size=1), nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True),
                DepthwiseSeparableConv(mid_channels, out_channels), nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))

        forward = _container_forward
    return LightConvBlock


def light_se_block(rule):
    class LightSEBlock(nn.Module):
        """global average pool -> Linear -> ReLU -> Linear -> Sigmoid -> channel scale."""

        def __init__(self, channels):
            super().__init__()
            self.avg_pool = nn.AdaptiveAvgPool2d(1)
            reduced = rule(channels)
            self.fc = nn.Sequential(nn.Linear(channels, reduced), nn.ReLU(inplace=True), nn.Linear(reduced, channels), nn.Sigmoid())

        forward = _container_forward
    return LightSEBlock


class UltraLightBase(nn.Module, EngineModuleMixin):
    VARIANT = None
    MODULE_DROPOUT = 0.0        # p of the registered nn.Dropout2d (the base variant registers one but never applies it)

    def __init__(self, num_classes=21):
        super().__init__()
        widths, mid_min, se_rule, _ = ULU_VARIANTS[self.VARIANT]
        Block = light_conv_block(mid_min)
        self.enc1 = Block(3, widths[0])
        self.enc2 = Block(widths[0], widths[1])
        self.enc3 = Block(widths[1], widths[2])
        self.enc4 = Block(widths[2], widths[3])
        self.bridge = Block(widths[3], widths[4])
        self.dec4 = Block(widths[4] + widths[3], widths[3])
        self.dec3 = Block(widths[3] + widths[2], widths[2])
        self.dec2 = Block(widths[2] + widths[1], widths[1])
        self.dec1 = Block(widths[1] + widths[0], widths[0])
        self.final = nn.Conv2d(widths[0], num_classes, 1)
        self.dropout = nn.Dropout2d(self.MODULE_DROPOUT)
        self.pool = nn.MaxPool2d(2, 2)
        if se_rule is not None:
            SE = light_se_block(se_rule)
            self.se1, self.se2, self.se3, self.se4 = SE(widths[0]), SE(widths[1]), SE(widths[2]), SE(widths[3])
        self.num_classes = num_classes
        self._init_engine_state()

    def _make_engine(self, device):
        return UltraLightUnetEngine(self.num_classes, self.VARIANT, device=device)

    def forward(self, x):
        return self._engine_forward(x)


def count_parameters(model):
    """nets/UltraLightweightUnet.py:110-111"""
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
