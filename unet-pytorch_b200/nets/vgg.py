"""Parameter containers of the VGG16 encoder, key-compatible with the reference (nets/vgg.py:47-75).

`features` keeps the reference's nn.Sequential indexing (convs at 0,2,5,7,10,12,14,17,19,21,24,26,28) so
state_dict keys, `weights_init`'s class-name matching and optimizers see the same tensors.  The modules are
never *called* on the hot path: Unet.forward hands their parameters to the CUDA engine.
"""
import torch.nn as nn

from ..engine import VGG16_CFG


class VGG(nn.Module):
    def __init__(self, in_channels=3):
        super().__init__()
        layers = []
        for item in VGG16_CFG + ["M"]:          # the reference builds (and skips at run time) a 5th pool
            if item == "M":
                layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
            else:
                i, cin, cout = item
                layers += [nn.Conv2d(in_channels if i == 0 else cin, cout, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
        self.features = nn.Sequential(*layers)
        for m in self.modules():                # same init family as nets/vgg.py:33-38
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.constant_(m.bias, 0)

    def forward(self, x):
        raise RuntimeError("VGG is a parameter container here; call Unet.forward (CUDA engine) instead")


def VGG16(pretrained=False, in_channels=3, **kwargs):
    if pretrained:
        raise RuntimeError("pretrained=True needs a download (nets/vgg.py:70 of the reference); load a state_dict instead")
    return VGG(in_channels=in_channels)
