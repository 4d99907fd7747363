"""Drop-in for the reference's nets/UltraLightweightUnet_large_optimized.py::UltraLightweightUnet_large_optimized
(lines 51-109): widths 44-88-176-352-704 (channel counts that are not multiples of 64 run zero-padded on the tensor
cores), SE after every encoder stage, Dropout2d(0.15) on the bridge."""
from ._ultralight import DepthwiseSeparableConv, UltraLightBase, count_parameters, light_conv_block, light_se_block  # noqa: F401

LightConvBlock = light_conv_block(16)
LightSEBlock = light_se_block(lambda c: max(8, c // 4))


class UltraLightweightUnet_large_optimized(UltraLightBase):
    VARIANT = "ultralight_large_optimized"
    MODULE_DROPOUT = 0.15
