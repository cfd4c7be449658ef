"""PeakSignalNoiseRatio with torchmetrics' semantics for the call the reference makes
(``psnr.PeakSignalNoiseRatio(data_range=(0.0, 1.0)).cuda()`` then ``metric(img, gt)``, trainer.py:171,198):
a tuple data_range clamps both inputs to it and uses its width; PSNR = 10 log10(range^2 / mse) over ALL elements of the
batch (torchmetrics' default reduction ``elementwise_mean`` with ``dim=None``); ``forward`` returns the value of the batch
and also accumulates for ``compute()``."""
from __future__ import annotations

import torch


class PeakSignalNoiseRatio(torch.nn.Module):
    def __init__(self, data_range=None, base: float = 10.0, reduction: str = "elementwise_mean", dim=None, **_ignored):
        super().__init__()
        if dim is not None or reduction != "elementwise_mean":
            raise NotImplementedError("this stand-in implements the default reduction over all elements only")
        self.clamp = None
        if isinstance(data_range, (tuple, list)):
            self.clamp = (float(data_range[0]), float(data_range[1]))
            self.range = float(data_range[1]) - float(data_range[0])
        elif data_range is not None:
            self.range = float(data_range)
        else:
            self.range = None
        self.base = float(base)
        self.register_buffer("sum_sq", torch.zeros((), dtype=torch.float64), persistent=False)
        self.register_buffer("total", torch.zeros((), dtype=torch.float64), persistent=False)
        self._min = None
        self._max = None

    def _psnr(self, sum_sq, total, rng):
        mse = sum_sq / total
        return (10.0 / torch.log(torch.tensor(self.base, dtype=mse.dtype, device=mse.device))) * torch.log(rng * rng / mse)

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if self.clamp is not None:
            preds = preds.clamp(*self.clamp)
            target = target.clamp(*self.clamp)
        d = (preds - target).double()
        s, n = (d * d).sum(), torch.tensor(float(d.numel()), dtype=torch.float64, device=d.device)
        if self.range is None:          # torchmetrics: range inferred from the targets seen so far
            lo, hi = target.min().double(), target.max().double()
            self._min = lo if self._min is None else torch.minimum(self._min, lo)
            self._max = hi if self._max is None else torch.maximum(self._max, hi)
            rng = hi - lo
        else:
            rng = torch.tensor(self.range, dtype=torch.float64, device=d.device)
        self.sum_sq = self.sum_sq.to(d.device) + s
        self.total = self.total.to(d.device) + n
        return self._psnr(s, n, rng).to(torch.float32)

    def update(self, preds, target):
        self.forward(preds, target)

    def compute(self) -> torch.Tensor:
        rng = torch.tensor(self.range, dtype=torch.float64, device=self.sum_sq.device) if self.range is not None else (self._max - self._min)
        return self._psnr(self.sum_sq, self.total, rng).to(torch.float32)

    def reset(self):
        self.sum_sq.zero_(); self.total.zero_()
        self._min = self._max = None
