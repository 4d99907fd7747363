"""Drop-in for the reference's nets/UltraLightweightUnet_large.py::UltraLightweightUnet_large (lines 55-113): widths
64-128-256-512-1024, SE after every encoder stage, Dropout2d(0.2) on the bridge."""
from ._ultralight import DepthwiseSeparableConv, UltraLightBase, count_parameters, light_conv_block, light_se_block  # noqa: F401

LightConvBlock = light_conv_block(16)
LightSEBlock = light_se_block(lambda c: max(8, c // 4))


class UltraLightweightUnet_large(UltraLightBase):
    VARIANT = "ultralight_large"
    MODULE_DROPOUT = 0.2
