from . import psnr  # noqa: F401
from .psnr import PeakSignalNoiseRatio  # noqa: F401
