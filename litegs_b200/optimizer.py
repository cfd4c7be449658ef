"""Optimizer step of the multi-view / data-parallel path.

The reference steps ``SparseGaussianAdam`` (litegs/training/optimizer.py:9-44): six ``adamUpdate`` launches per iteration,
one per parameter group, each reading a chunk-compacted gradient (GR/compact.cu:320-417: Adam WITHOUT bias correction,
eps 1e-15, betas 0.9/0.999, only visible chunks update).  Here the gradients of all views of a step already sit in one
dense buffer (``dist.GradAccumulator``, all-reduced over ranks), so the step is ONE kernel over all six tensors
(csrc/optim.cu) that also clears the consumed gradient rows and chunk marks.

``get_optimizer`` mirrors the reference's group layout and learning rates (optimizer.py:74-95); ``Scheduler`` is its
log-linear position learning-rate schedule (optimizer.py:46-72).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict

import torch

from . import _lib
from .dist import PARAM_ORDER, GradAccumulator
from .fused import _ptr, _stream


class FusedAdam:
    """Adam (no bias correction) over the six parameter tensors from a GradAccumulator.

    params: dict name -> contiguous float32 CUDA tensor [..., C, S], updated in place.   lr: dict name -> float."""

    def __init__(self, params: Dict[str, torch.Tensor], lr: Dict[str, float], eps: float = 1e-15, betas=(0.9, 0.999)):
        for k in PARAM_ORDER:
            t = params[k]
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise RuntimeError(f"FusedAdam: parameter '{k}' must be a contiguous float32 CUDA tensor")
        self.params = {k: params[k] for k in PARAM_ORDER}
        self.lr = {k: float(lr[k]) for k in PARAM_ORDER}
        self.eps, self.betas = float(eps), (float(betas[0]), float(betas[1]))
        C, S = params["xyz"].shape[-2:]
        self.C, self.S = int(C), int(S)
        self.rows = [int(params[k].numel() // (C * S)) for k in PARAM_ORDER]
        n_rows = sum(self.rows)
        dev = params["xyz"].device
        self.exp_avg = torch.zeros((n_rows, C, S), dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros((n_rows, C, S), dtype=torch.float32, device=dev)
        self._rows_c = (ctypes.c_int * 6)(*self.rows)

    def state_for(self, name: str):
        """(exp_avg, exp_avg_sq) views shaped like parameter `name` (the reference keeps them per parameter)."""
        r0 = sum(self.rows[: PARAM_ORDER.index(name)])
        r1 = r0 + self.rows[PARAM_ORDER.index(name)]
        shp = self.params[name].shape
        return self.exp_avg[r0:r1].view(shp), self.exp_avg_sq[r0:r1].view(shp)

    @torch.no_grad()
    def step(self, acc: GradAccumulator, clear_grad: bool = True, all_chunks: bool = False):
        """One step from acc.buf (already all-reduced).  Only chunks marked in acc.touched update unless all_chunks."""
        if tuple(acc.buf.shape) != tuple(self.exp_avg.shape):
            raise RuntimeError("FusedAdam.step: accumulator and optimizer were built for different parameter shapes")
        dev = self.exp_avg.device
        ptrs = (ctypes.c_void_p * 6)(*[self.params[k].data_ptr() for k in PARAM_ORDER])
        lrs = (ctypes.c_float * 6)(*[self.lr[k] for k in PARAM_ORDER])
        with torch.cuda.device(dev):
            _lib.call("lgs_adam_step_dense", ptrs, self._rows_c, lrs, _ptr(acc.buf), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                      None if all_chunks else _ptr(acc.touched), self.C, self.S, self.betas[0], self.betas[1], self.eps,
                      int(bool(clear_grad)), _stream(dev))
        if clear_grad and all_chunks:
            acc.touched.zero_()


class Scheduler:
    """optimizer.py:46-72: log-linear interpolation of the position learning rate, other groups constant."""

    def __init__(self, optimizer: FusedAdam, lr_init: float, lr_final: float, max_epochs: int = 10000):
        self.opt, self.lr_init, self.lr_final, self.max_epochs = optimizer, float(lr_init), float(lr_final), int(max_epochs)
        self.last_epoch = 0
        self.opt.lr["xyz"] = self._lr()

    def _lr(self) -> float:
        if self.lr_init == 0.0 and self.lr_final == 0.0:
            return 0.0
        t = min(max(self.last_epoch / self.max_epochs, 0.0), 1.0)
        return math.exp(math.log(self.lr_init) * (1 - t) + math.log(self.lr_final) * t)

    def step(self):
        self.last_epoch += 1
        self.opt.lr["xyz"] = self._lr()


def get_optimizer(params: Dict[str, torch.Tensor], spatial_lr_scale: float, position_lr_init=0.00016, position_lr_final=0.0000016,
                  position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005, rotation_lr=0.001):
    """optimizer.py:74-95 with the reference's default OptimizationParams (arguments.py:82-88)."""
    lr = {"xyz": position_lr_init * spatial_lr_scale, "sh_0": feature_lr, "sh_rest": feature_lr / 10.0, "opacity": opacity_lr,
          "scale": scaling_lr, "rot": rotation_lr}
    opt = FusedAdam(params, lr, eps=1e-15)
    sched = Scheduler(opt, position_lr_init * spatial_lr_scale, position_lr_final * spatial_lr_scale, position_lr_max_steps)
    return opt, sched
