"""Drop-in for the reference's ``fused_ssim`` package (litegs/submodules/fused_ssim): ``import fused_ssim`` resolves
here when the repo root is on sys.path (trainer.py:3,145 calls fused_ssim.fused_l1_ssim_loss)."""
from litegs_b200.ssim import (FusedL1SSIMLossMap, FusedSSIMMap, allowed_padding, fused_l1_ssim_loss, fused_ssim, fusedl1ssim_loss,  # noqa: F401
                              fusedl1ssim_loss_backward, fusedssim, fusedssim_backward, l1_ssim_loss_and_grad)
