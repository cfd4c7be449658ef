"""Minimal stand-in for the switch the reference's ``StatisticsHelperInst`` exposes to the render path.

Only what the hot path reads is mirrored (litegs/utils/statistic_helper.py:26-30,242-259): ``bStart`` turns
on the per-splat fragment statistics outputs of the raster kernels, ``cached_sorted_tile_list`` /
``cur_sample`` select a heaviest-first tile order.  The densification policy that consumes the
statistics is out of scope; hooks receive the raw tensors."""
from __future__ import annotations

from typing import Callable, Optional


class _Statistics:
    def __init__(self):
        self.bStart = False
        self.cur_sample: Optional[str] = None
        self.cached_sorted_tile_list: dict = {}
        self.on_fragment_weight: Optional[Callable] = None
        self.on_fragment_err: Optional[Callable] = None
        self.on_visible: Optional[Callable] = None
        self.on_compact_mask: Optional[Callable] = None
        self.on_blend_count: Optional[Callable] = None

    class _Guard:
        def __init__(self, owner, active):
            self.owner, self.active = owner, active

        def __enter__(self):
            if self.active:
                self.owner.bStart = True

        def __exit__(self, *a):
            if self.active:
                self.owner.bStart = False

    def enabled(self, active: bool = True):
        return self._Guard(self, active)


StatisticsHelperInst = _Statistics()
