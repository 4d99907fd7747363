"""Drop-in for the reference's nets/unet.py::Unet (lines 24-94): same constructor, attributes, state_dict keys
and forward contract (NCHW fp32 image in, NCHW fp32 logits out), executed by the sm_100a engine."""
import torch.nn as nn

from ..engine import VGGUnetEngine
from ..graph import ResNet50UnetEngine
from ._function import EngineModuleMixin
from .resnet import resnet50
from .vgg import VGG16


class unetUp(nn.Module):
    """Parameter container mirroring nets/unet.py:8-22 (conv1 over cat[skip, up2x(x)], conv2)."""

    def __init__(self, in_size, out_size):
        super().__init__()
        self.conv1 = nn.Conv2d(in_size, out_size, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(out_size, out_size, kernel_size=3, padding=1)
        self.up = nn.UpsamplingBilinear2d(scale_factor=2)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, inputs1, inputs2):
        raise RuntimeError("unetUp is a parameter container here; call Unet.forward (CUDA engine) instead")


class Unet(nn.Module, EngineModuleMixin):
    def __init__(self, num_classes=21, pretrained=False, backbone="vgg"):
        super().__init__()
        if backbone == "vgg":
            self.vgg = VGG16(pretrained=pretrained)
            in_filters = [192, 384, 768, 1024]
        elif backbone == "resnet50":
            self.resnet = resnet50(pretrained=pretrained)
            in_filters = [192, 512, 1024, 3072]
        else:
            raise ValueError("Unsupported backbone - `{}`, Use vgg, resnet50.".format(backbone))
        out_filters = [64, 128, 256, 512]
        self.up_concat4 = unetUp(in_filters[3], out_filters[3])
        self.up_concat3 = unetUp(in_filters[2], out_filters[2])
        self.up_concat2 = unetUp(in_filters[1], out_filters[1])
        self.up_concat1 = unetUp(in_filters[0], out_filters[0])
        if backbone == "resnet50":
            self.up_conv = nn.Sequential(                   # nets/unet.py:47-54
                nn.UpsamplingBilinear2d(scale_factor=2),
                nn.Conv2d(out_filters[0], out_filters[0], kernel_size=3, padding=1), nn.ReLU(),
                nn.Conv2d(out_filters[0], out_filters[0], kernel_size=3, padding=1), nn.ReLU())
        else:
            self.up_conv = None
        self.final = nn.Conv2d(out_filters[0], num_classes, 1)
        self.backbone = backbone
        self.num_classes = num_classes
        self._init_engine_state()

    def _make_engine(self, device):
        if self.backbone == "resnet50":
            return ResNet50UnetEngine(self.num_classes, device=device)
        return VGGUnetEngine(self.num_classes, in_channels=3, device=device)

    def forward(self, inputs):
        return self._engine_forward(inputs)

    def _backbone(self):
        return self.vgg if self.backbone == "vgg" else self.resnet

    def freeze_backbone(self):
        for param in self._backbone().parameters():
            param.requires_grad = False

    def unfreeze_backbone(self):
        for param in self._backbone().parameters():
            param.requires_grad = True
