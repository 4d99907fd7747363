"""Drop-in for the reference's nets/LightWeightUnet.py::LightweightUnet (lines 125-177): same constructor, module tree,
state_dict keys, kaiming fan_out init, freeze/unfreeze_backbone and forward contract (logits at H/2 x W/2; the losses in
nets/unet_training.py resize them to the label size), executed by the sm_100a graph engine: dense conv3x3 + BatchNorm +
ReLU blocks with squeeze-excite residual blocks, widths 24-48-96-192-384 zero-padded to multiples of 64 channels."""
import torch.nn as nn

from ..graph import LightweightUnetEngine
from ._function import EngineModuleMixin


def _container_forward(self, *a, **k):
    raise RuntimeError(f"{type(self).__name__} is a parameter container here; call LightweightUnet.forward (CUDA engine)")


class ConvBlock(nn.Module):
    """conv3x3 + bias, BatchNorm2d, ReLU -- nets/LightWeightUnet.py:5-15"""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1),
                                  nn.BatchNorm2d(out_channels, momentum=0.1), nn.ReLU(inplace=True))

    forward = _container_forward


class SEBlock(nn.Module):
    """avg-pool, Linear(c, c // r), ReLU, Linear(c // r, c), Sigmoid, channel scale -- nets/LightWeightUnet.py:18-33"""

    def __init__(self, channels, reduction=4):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channels, channels // reduction), nn.ReLU(inplace=True),
                                nn.Linear(channels // reduction, channels), nn.Sigmoid())

    forward = _container_forward


class ResidualBlock(nn.Module):
    """conv-BN-ReLU-conv-BN-SE, += input, ReLU -- nets/LightWeightUnet.py:36-55"""

    def __init__(self, channels):
        super().__init__()
        self.conv1 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn1 = nn.BatchNorm2d(channels, momentum=0.1)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(channels, channels, kernel_size=3, padding=1)
        self.bn2 = nn.BatchNorm2d(channels, momentum=0.1)
        self.se = SEBlock(channels)

    forward = _container_forward


class LightweightVGG(nn.Module):
    """five stages of ConvBlock + ResidualBlock + MaxPool2d(2), Dropout2d(0.1) on every feature -- :58-112"""

    def __init__(self, in_channels=3, pretrained=False):
        super().__init__()
        widths = (24, 48, 96, 192, 384)
        cin = in_channels
        for k, w in enumerate(widths, start=1):
            setattr(self, f"stage{k}", nn.Sequential(ConvBlock(cin, w), ResidualBlock(w), nn.MaxPool2d(kernel_size=2, stride=2)))
            cin = w
        self.dropout = nn.Dropout2d(0.1)

    forward = _container_forward


class LightweightUnetUp(nn.Module):
    """up2x(inputs2), cat([inputs1, up]), ConvBlock, ResidualBlock, Dropout2d(0.1) -- :115-129"""

    def __init__(self, in_size, out_size):
        super().__init__()
        self.up = nn.UpsamplingBilinear2d(scale_factor=2)
        self.conv = nn.Sequential(ConvBlock(in_size, out_size), ResidualBlock(out_size))
        self.dropout = nn.Dropout2d(0.1)

    forward = _container_forward


class LightweightUnet(nn.Module, EngineModuleMixin):
    def __init__(self, num_classes=21, pretrained=False, backbone="lightweight_vgg", in_channels=3):
        super().__init__()
        if backbone == "lightweight_vgg":
            self.backbone = LightweightVGG(in_channels=in_channels, pretrained=pretrained)
        else:
            raise ValueError("Unsupported backbone - `{}`, Only lightweight_vgg is supported.".format(backbone))
        self.up_concat4 = LightweightUnetUp(576, 192)
        self.up_concat3 = LightweightUnetUp(288, 96)
        self.up_concat2 = LightweightUnetUp(144, 48)
        self.up_concat1 = LightweightUnetUp(72, 24)
        self.final_conv = nn.Sequential(ConvBlock(24, 24), nn.Dropout2d(0.1), ResidualBlock(24), nn.Conv2d(24, num_classes, 1))
        self.backbone_name = backbone
        self.num_classes, self.in_channels = num_classes, in_channels
        self._initialize_weights()
        self._init_engine_state()

    def _initialize_weights(self):          # nets/LightWeightUnet.py:150-158
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_engine(self, device):
        return LightweightUnetEngine(self.num_classes, in_channels=self.in_channels, device=device)

    def forward(self, inputs):
        return self._engine_forward(inputs)

    def freeze_backbone(self):              # :171-173
        for param in self.backbone.parameters():
            param.requires_grad = False

    def unfreeze_backbone(self):            # :175-177
        for param in self.backbone.parameters():
            param.requires_grad = True


def count_parameters(model):
    return sum(p.numel() for p in model.parameters() if p.requires_grad)
