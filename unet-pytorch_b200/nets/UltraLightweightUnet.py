"""Drop-in for the reference's nets/UltraLightweightUnet.py::UltraLightweightUnet (lines 57-108): widths
32-64-128-256-512, no SE blocks, Dropout2d(0.1) registered but never applied in forward (as in the reference)."""
from ._ultralight import DepthwiseSeparableConv, UltraLightBase, count_parameters, light_conv_block, light_se_block  # noqa: F401

LightConvBlock = light_conv_block(8)
LightSEBlock = light_se_block(lambda c: max(4, c // 8))          # defined by the reference file (lines 38-54), unused by the net


class UltraLightweightUnet(UltraLightBase):
    VARIANT = "ultralight"
    MODULE_DROPOUT = 0.1
