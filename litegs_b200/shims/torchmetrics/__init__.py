"""Minimal ``torchmetrics`` stand-in: only ``torchmetrics.image.psnr.PeakSignalNoiseRatio`` (and the ``image`` namespace the
reference imports from: litegs/training/trainer.py:4,171-199)."""
from . import image  # noqa: F401
