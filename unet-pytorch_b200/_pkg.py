"""Public surface of the package (see unet_pytorch_b200/__init__.py for why this is not __init__)."""
from . import _lib, ops  # noqa: F401
from .engine import TraditionalUnetEngine, UNetEngine, VGGUnetEngine, vgg_unet_param_shapes  # noqa: F401
from .graph import GraphEngine, ImprovedSegNetEngine, LightweightUnetEngine, ResNet50UnetEngine, UltraLightUnetEngine  # noqa: F401
from .nets.unet import Unet  # noqa: F401
from .nets.TraditionalUnet import TraditionalUnet  # noqa: F401
from .nets.LightWeightUnet import LightweightUnet  # noqa: F401
from .nets.UltraLightweightUnet import UltraLightweightUnet  # noqa: F401
from .nets.UltraLightweightUnet_large import UltraLightweightUnet_large  # noqa: F401
from .nets.UltraLightweightUnet_large_optimized import UltraLightweightUnet_large_optimized  # noqa: F401
from .nets.RepVGG_Unet import ImprovedSegNet  # noqa: F401
from .nets.unet_training import CE_Loss, Dice_loss, Focal_Loss, ce_dice_loss  # noqa: F401
from .utils.utils_metrics import (f_score, fast_hist, fast_hist_device, per_Accuracy, per_class_iu,  # noqa: F401
                                  per_class_PA_Recall, per_class_Precision)
from .trainer import UnetTrainer, FlatBuckets, GradientSync  # noqa: F401,E402
from . import synthetic  # noqa: F401,E402
