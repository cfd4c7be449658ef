import sys, torch
sys.path.insert(0, ".")
from litegs_b200 import ssim
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
img = torch.rand((1, 3, 1080, 1920), generator=g).to(dev); gt = torch.rand((1, 3, 1080, 1920), generator=g).to(dev)
for _ in range(3):
    ssim.l1_ssim_loss_and_grad(img, gt, 0.2)
torch.cuda.synchronize()
