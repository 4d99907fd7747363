"""Import alias: the package directory is `unet-pytorch_b200/` (not a valid Python identifier), so
`import unet_pytorch_b200` forwards to it."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "unet-pytorch_b200"))
from ._pkg import *  # noqa: F401,F403,E402
