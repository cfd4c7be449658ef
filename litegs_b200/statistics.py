"""Per-Gaussian statistics of the render path, as the densifier of the reference consumes them.

``StatisticsHelper`` mirrors ``litegs/utils/statistic_helper.py`` -- same attribute and method names, same accumulation
rules (:82-156), same read-outs (``get_mean`` :215-223, ``get_var`` :225-243 with its ``count + 1`` denominators,
``get_global_culling`` :245-248) -- on top of this library's ``gpu_driven_pipeline_sparse_op`` (device-side visible count, no
host synchronisation).  ``litegs_b200.wrapper`` / ``render`` feed it exactly where the reference's wrapper does
(wrapper.py:501-506, 733-737; render/__init__.py:24-25, 84-85):

    fragment_weight : sum, sum of squares and count of the per-view blending weight of each Gaussian     (prune score)
    fragment_err    : d_opacity, err_square_sum * gmax^2 and the fragment count                          (split / clone score)
    visible_count   : number of views in which the Gaussian owned at least one tile
    tile blend count: per view name, the tiles in heaviest-first order (fed back as ``specific_tiles``)

The densification POLICY (what to split, clone or prune from these numbers: litegs/training/densify.py) stays out of scope.
When the unmodified reference package runs on this library it uses its own ``StatisticsHelperInst``; this one serves the
operator surface of ``litegs_b200`` itself (Level A and the fused Level B)."""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch


class MeanStdData:
    def __init__(self, data_shape, cluster_shape, device):
        self.sum = torch.zeros((*data_shape, *cluster_shape), device=device)
        self.square_sum = torch.zeros((*data_shape, *cluster_shape), device=device)
        self.count = torch.zeros(tuple(cluster_shape), device=device, dtype=torch.int32)


def _uncluster(t: torch.Tensor) -> torch.Tensor:
    return t.reshape(*t.shape[:-2], t.shape[-2] * t.shape[-1])


class StatisticsHelper:
    def __init__(self, chunk_num: int = 0, chunk_size: int = 0):
        self.cached_tiles_blend_count: Dict[str, torch.Tensor] = {}
        self.cached_sorted_tile_list: Dict[str, torch.Tensor] = {}
        self.cur_sample: Optional[str] = None
        self.device = None
        # hooks the operator surface calls; by default they are this object's own update methods
        self.on_fragment_weight: Optional[Callable] = self._on_fragment_weight
        self.on_fragment_err: Optional[Callable] = self._on_fragment_err
        self.on_visible: Optional[Callable] = self.update_visible_count
        self.on_compact_mask: Optional[Callable] = self.set_compact_mask
        self.on_blend_count: Optional[Callable] = self.update_tile_blend_count
        self.reset(chunk_num, chunk_size, lambda epoch: False)

    # ---- life cycle (statistic_helper.py:26-49, 250-262) ----------------------------------------------------
    def reset(self, chunk_num: int, chunk_size: int, statistics_check_handle: Optional[Callable[[int], bool]] = None, device=None):
        self.bStart = False
        if statistics_check_handle is not None:
            self.is_statistics_enabled = statistics_check_handle
        self.chunk_num, self.chunk_size = int(chunk_num), int(chunk_size)
        self.mean_and_std: Dict[str, MeanStdData] = {}
        self.max_and_min: Dict[str, list] = {}
        if device is not None:
            self.device = torch.device(device)
        self.visible_count = None           # allocated on first use (the reference allocates on 'cuda' at import time)
        self.compact_mask: Optional[torch.Tensor] = None
        self.valid_length: Optional[torch.Tensor] = None

    class _Guard:
        def __init__(self, owner):
            self.owner = owner

        def __enter__(self):
            if self.owner is not None:
                self.owner.bStart = True

        def __exit__(self, *a):
            if self.owner is not None:
                self.owner.bStart = False

    def try_start(self, epoch: int):
        return self._Guard(self if self.is_statistics_enabled(epoch) else None)

    def enabled(self, active: bool = True):
        return self._Guard(self if active else None)

    # ---- feeding (statistic_helper.py:60-156) -----------------------------------------------------------------
    @torch.no_grad()
    def set_compact_mask(self, compact_mask: torch.Tensor, valid_length: Optional[torch.Tensor] = None):
        self.compact_mask, self.valid_length = compact_mask, valid_length
        self.device = compact_mask.device

    def _sparse_add(self, dst: torch.Tensor, src: torch.Tensor):
        """dst[:, ids[j], :] += src[:, j, :] for j < *valid_length (device count); dst [E,C,S], src [E,A,S]."""
        from . import fused
        fused.gpu_driven_pipeline_sparse_op(dst, src, self.compact_mask, self.valid_length, "add")

    @torch.no_grad()
    def update_visible_count(self, visible_mask: torch.Tensor):
        dev = visible_mask.device
        if self.visible_count is None:
            self.visible_count = torch.zeros((self.chunk_num, self.chunk_size), dtype=torch.int32, device=dev)
        if self.compact_mask is None:
            self.visible_count += visible_mask.sum(0).reshape(self.chunk_num, self.chunk_size).to(torch.int32)
        elif self.valid_length is None:
            self.visible_count[self.compact_mask] += visible_mask.sum(0).reshape(-1, self.chunk_size).to(torch.int32)
        else:   # GPU-driven pipeline: the tail of the compacted mask is dirty and must be ignored (statistic_helper.py:88-92)
            self._sparse_add(self.visible_count.view(1, -1, self.chunk_size),
                             visible_mask.sum(0, dtype=torch.int32).reshape(1, -1, self.chunk_size))

    @torch.no_grad()
    def update_mean_std(self, key: str, tensor_sum: torch.Tensor, square_sum: torch.Tensor, count, bCompacted: Optional[bool] = None):
        if bCompacted is None:
            bCompacted = self.compact_mask is not None
        S = self.chunk_size
        if bCompacted:
            tensor_sum = tensor_sum.reshape(*tensor_sum.shape[:-1], -1, S)
            square_sum = square_sum.reshape(*square_sum.shape[:-1], -1, S)
            if isinstance(count, torch.Tensor):
                count = count.reshape(-1, S)
        elif isinstance(count, torch.Tensor):
            count = count.squeeze()
        data = self.mean_and_std.get(key)
        if data is None:
            if bCompacted:
                data = MeanStdData(list(tensor_sum.shape[:-2]), [self.chunk_num, S], tensor_sum.device)
            else:
                data = MeanStdData(list(tensor_sum.shape[:-1]), [tensor_sum.shape[-1]], tensor_sum.device)
            self.mean_and_std[key] = data
        if not bCompacted:
            data.sum += tensor_sum; data.square_sum += square_sum; data.count += count
            return
        if self.valid_length is None:
            data.sum[..., self.compact_mask, :] += tensor_sum
            data.square_sum[..., self.compact_mask, :] += square_sum
            data.count[self.compact_mask, :] += count
            return
        C, A = data.sum.shape[-2], tensor_sum.shape[-2]
        self._sparse_add(data.sum.view(-1, C, S), tensor_sum.reshape(-1, A, S).contiguous())
        self._sparse_add(data.square_sum.view(-1, C, S), square_sum.reshape(-1, A, S).contiguous())
        self._sparse_add(data.count.view(-1, C, S), count.reshape(-1, A, S).to(torch.int32).contiguous())

    def _on_fragment_weight(self, fragment_weight, fragment_count):
        self.update_mean_std("fragment_weight", fragment_weight, fragment_weight * fragment_weight, fragment_count, None)   # wrapper.py:505

    def _on_fragment_err(self, d_opacity, err_square, fragment_count):
        self.update_mean_std("fragment_err", d_opacity, err_square, fragment_count, None)                                   # wrapper.py:506

    @torch.no_grad()
    def update_tile_blend_count(self, pixel_blend_count: torch.Tensor, tilesize_h: int, tilesize_w: int):
        """Per-tile maximum of last_contributor and the tiles in descending order of it (statistic_helper.py:66-79): the
        reference's heaviest-first tile schedule for the NEXT time this sample is rendered."""
        N, _, H, W = pixel_blend_count.shape
        th, tw = int(tilesize_h), int(tilesize_w)
        gy, gx = (H + th - 1) // th, (W + tw - 1) // tw
        c = pixel_blend_count.detach().view(torch.uint16).to(torch.int32).reshape(N, gy, th, gx, tw).permute(1, 3, 0, 2, 4).reshape(gy * gx, -1)
        tiles = c.max(dim=1).values
        if self.cur_sample is not None:
            self.cached_tiles_blend_count[self.cur_sample] = tiles
            self.cached_sorted_tile_list[self.cur_sample] = tiles.sort(descending=True)[1].int() + 1

    # ---- read-outs (statistic_helper.py:215-248) ----------------------------------------------------------------
    @torch.no_grad()
    def get_mean(self, key: str):
        data = self.mean_and_std.get(key)
        if data is None:
            return None
        return _uncluster(data.sum / (data.count + 1e-9)), data.count.reshape(-1)

    @torch.no_grad()
    def get_var(self, key: str):
        data = self.mean_and_std.get(key)
        if data is None:
            return None
        mean = data.sum / (data.count + 1)
        var = (data.square_sum / (data.count + 1) - mean ** 2).clamp_min(0)
        if self.compact_mask is not None:
            var = _uncluster(var)
        return var, data.count.reshape(-1)

    @torch.no_grad()
    def get_global_culling(self) -> torch.Tensor:
        if self.visible_count is None:
            raise RuntimeError("no view was rendered with statistics enabled")
        return _uncluster(self.visible_count == 0)


StatisticsHelperInst = StatisticsHelper(0, 0)
