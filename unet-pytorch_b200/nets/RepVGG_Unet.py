"""Drop-in for the reference's nets/RepVGG_Unet.py::ImprovedSegNet (lines 149-206) with `use_repvgg=True`: the module tree and
state_dict keys are the reference's (training form: `conv.3.conv1 / bn1 / conv2 / bn2`; after `switch_to_deploy()`:
`conv.3.reparam_conv`), the forward runs on the CUDA graph engine (graph.py::improved_segnet_program).

`switch_to_deploy` folds each RepVGGBlock's two conv + BatchNorm branches into one 3x3 conv with bias -- the algebra of
nets/RepVGG_Unet.py:62-99: a BatchNorm in eval mode is the affine map y = (x - mean) * gamma / sqrt(var + eps) + beta, so
conv followed by it is a conv with weights scaled per output channel and a bias; the 1x1 kernel sits at the centre tap of the
3x3 one; parallel branches add.  The `use_repvgg=False` variant (FusedMBConv with ReLU6) is not part of the instruction set."""
import torch
import torch.nn as nn

from ..graph import ImprovedSegNetEngine
from ._function import EngineModuleMixin
from ._ultralight import _container_forward, light_se_block

LightSEBlock = light_se_block(lambda c: max(8, c // 4))


class RepVGGBlock(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, groups=1):
        super().__init__()
        if (kernel_size, stride, padding, groups) != (3, 1, 1, 1):
            raise NotImplementedError("RepVGGBlock: the CUDA path covers the reference's only use (3x3, stride 1, padding 1)")
        if in_channels == out_channels:
            raise NotImplementedError("RepVGGBlock identity branch (in == out channels) is not used by ImprovedSegNet")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.conv2 = nn.Conv2d(in_channels, out_channels, 1, 1, 0, bias=False)
        self.bn2 = nn.BatchNorm2d(out_channels)
        self.identity = False
        self.relu = nn.ReLU(inplace=True)
        self.deploy = False

    forward = _container_forward

    @staticmethod
    def _affine(bn):
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        return scale, bn.bias - bn.running_mean * scale

    def get_equivalent_kernel_bias(self):
        s3, b3 = self._affine(self.bn1)
        s1, b1 = self._affine(self.bn2)
        kernel = self.conv1.weight * s3.view(-1, 1, 1, 1)
        kernel[:, :, 1, 1] += self.conv2.weight[:, :, 0, 0] * s1.view(-1, 1)
        return kernel, b3 + b1

    def switch_to_deploy(self):
        with torch.no_grad():
            kernel, bias = self.get_equivalent_kernel_bias()
            self.reparam_conv = nn.Conv2d(self.in_channels, self.out_channels, 3, 1, 1).to(kernel.device)
            self.reparam_conv.weight.copy_(kernel)
            self.reparam_conv.bias.copy_(bias)
        self.deploy = True          # like the reference, the training branches stay registered (same state_dict keys); unused


class LightweightConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, use_repvgg=True):
        super().__init__()
        if not use_repvgg:
            raise NotImplementedError("ImprovedSegNet(use_repvgg=False) (FusedMBConv, ReLU6) is not built on the CUDA engine")
        mid_channels = max(16, out_channels // 2)
        self.conv = nn.Sequential(nn.Conv2d(in_channels, mid_channels, 1), nn.BatchNorm2d(mid_channels), nn.ReLU(inplace=True),
                                  RepVGGBlock(mid_channels, out_channels))

    forward = _container_forward


class ImprovedSegNet(nn.Module, EngineModuleMixin):
    def __init__(self, num_classes=21, use_repvgg=True):
        super().__init__()
        w = (44, 88, 176, 352, 704)
        self.enc1 = LightweightConvBlock(3, w[0], use_repvgg)
        self.enc2 = LightweightConvBlock(w[0], w[1], use_repvgg)
        self.enc3 = LightweightConvBlock(w[1], w[2], use_repvgg)
        self.enc4 = LightweightConvBlock(w[2], w[3], use_repvgg)
        self.bridge = LightweightConvBlock(w[3], w[4], use_repvgg)
        self.dec4 = LightweightConvBlock(w[4] + w[3], w[3], use_repvgg)
        self.dec3 = LightweightConvBlock(w[3] + w[2], w[2], use_repvgg)
        self.dec2 = LightweightConvBlock(w[2] + w[1], w[1], use_repvgg)
        self.dec1 = LightweightConvBlock(w[1] + w[0], w[0], use_repvgg)
        self.se1, self.se2, self.se3, self.se4 = LightSEBlock(w[0]), LightSEBlock(w[1]), LightSEBlock(w[2]), LightSEBlock(w[3])
        self.final = nn.Conv2d(w[0], num_classes, 1)
        self.dropout = nn.Dropout2d(0.15)
        self.pool = nn.MaxPool2d(2, 2)
        self.num_classes = num_classes
        self._deployed = False
        self._init_engine_state()

    def _make_engine(self, device):
        return ImprovedSegNetEngine(self.num_classes, deploy=self._deployed, device=device)

    def forward(self, x):
        return self._engine_forward(x)

    def switch_to_deploy(self):
        """nets/RepVGG_Unet.py:201-206: every RepVGGBlock becomes one conv3x3 + bias; the engines are rebuilt for the new
        parameter set."""
        for m in self.modules():
            if isinstance(m, RepVGGBlock) and not m.deploy:
                m.switch_to_deploy()
        self._deployed = True
        for e in self._engines.values():
            e.release()
        self._init_engine_state()
