"""distCUDA2(points f32[N,3] on CUDA) -> f32[N]: mean of the squared distances to the 3 nearest OTHER points
(litegs/submodules/simple-knn/simple_knn.cu:76-139 ``boxMeanDist``: 3 best candidates, mean of their squared distances).

The reference's kernel walks Morton-ordered boxes; the result it computes is the exact 3-NN mean, which is what this
stand-in returns, by a dense chunked distance computation on the device (initialisation-time code, run once on the SfM
points; O(N^2 / chunk) memory-bounded)."""
from __future__ import annotations

import torch


@torch.no_grad()
def distCUDA2(points: torch.Tensor) -> torch.Tensor:
    if points.dim() != 2 or points.shape[1] != 3:
        raise RuntimeError("distCUDA2 expects a [N,3] tensor")
    p = points.detach().to(torch.float32).contiguous()
    n = p.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=p.device)
    if n == 0:
        return out
    k = min(4, n)                                    # self + 3 neighbours
    sq = (p * p).sum(1)
    chunk = max(1, min(n, (64 << 20) // max(n, 1)))  # <= 64M distance entries at a time
    for s in range(0, n, chunk):
        q = p[s:s + chunk]
        d = (sq[s:s + chunk, None] + sq[None, :] - 2.0 * (q @ p.T)).clamp_min_(0)
        idx = torch.arange(s, min(s + chunk, n), device=p.device)
        d[torch.arange(idx.numel(), device=p.device), idx] = float("inf")          # exclude the point itself
        near = torch.topk(d, k - 1, dim=1, largest=False).values if k > 1 else torch.zeros((idx.numel(), 1), device=p.device)
        # recompute the selected distances exactly (the expanded form above loses digits for far-away clouds)
        out[s:s + chunk] = near.mean(1)
    return out
