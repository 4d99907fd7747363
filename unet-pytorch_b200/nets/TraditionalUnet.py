"""Drop-in for the reference's nets/TraditionalUnet.py::TraditionalUnet (lines 5-115): same constructor, module
tree, state_dict keys (conv / BatchNorm2d weights, biases and running statistics) and forward contract, executed
by the sm_100a engine (conv3x3 -> BatchNorm -> ReLU blocks, 32-64-128-256 channels, 3 decoder stages)."""
import torch.nn as nn

from ..engine import TraditionalUnetEngine
from ._function import EngineModuleMixin


def _container_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is a parameter container here; call TraditionalUnet.forward (CUDA engine)")


class DoubleConv(nn.Module):
    """(conv3x3 + bias, BatchNorm2d, ReLU) x 2 -- nets/TraditionalUnet.py:5-18"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels), nn.ReLU(inplace=True))

    forward = _container_forward


class Down(nn.Module):
    """MaxPool2d(2) + DoubleConv -- nets/TraditionalUnet.py:21-30"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels))

    forward = _container_forward


class Up(nn.Module):
    """UpsamplingBilinear2d(2) of x1, cat([x2, x1]), DoubleConv -- nets/TraditionalUnet.py:33-42"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.up = nn.UpsamplingBilinear2d(scale_factor=2)
        self.conv = DoubleConv(in_channels, out_channels)

    forward = _container_forward


class TraditionalUnet(nn.Module, EngineModuleMixin):
    def __init__(self, in_channels=3, num_classes=21):
        super().__init__()
        self.inc = DoubleConv(in_channels, 32)
        self.down1 = Down(32, 64)
        self.down2 = Down(64, 128)
        self.down3 = Down(128, 256)
        self.up1 = Up(256 + 128, 128)
        self.up2 = Up(128 + 64, 64)
        self.up3 = Up(64 + 32, 32)
        self.outc = nn.Conv2d(32, num_classes, kernel_size=1)
        self.in_channels, self.num_classes = in_channels, num_classes
        self._initialize_weights()
        self._init_engine_state()

    def _initialize_weights(self):          # nets/TraditionalUnet.py:69-77
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_engine(self, device):
        return TraditionalUnetEngine(self.num_classes, in_channels=self.in_channels, device=device)

    def forward(self, x):
        return self._engine_forward(x)

    def _set_encoder_grad(self, flag):
        for part in (self.inc, self.down1, self.down2, self.down3):
            for param in part.parameters():
                param.requires_grad = flag

    def freeze_encoder(self):               # nets/TraditionalUnet.py:95-104
        self._set_encoder_grad(False)

    def unfreeze_encoder(self):             # nets/TraditionalUnet.py:106-115
        self._set_encoder_grad(True)
