"""Fused per-view render pipeline ("Level B") above the C ABI.

One view = render_preprocess + render of the reference (litegs/render/__init__.py:11-94) collapsed to

    cull_chunks -> project_forward -> [one 16-byte D2H: sizes + depth-key range] -> depth radix sort (N keys) ->
    gathered scan -> emit_pairs -> tile radix sort (tile bits only) -> tile_range -> raster_forward

and the backward to  raster_backward -> project_backward  (two kernels + one memset).  Everything runs on
the current CUDA stream; the only host synchronisation is the 16-byte read-back of (visible chunks, pair
count, depth-key range) that sizes the sorts, exactly one per view (the reference pays two: GR/compact.cu:527-549 and
GR/binning.cu:137-163, hidden behind last epoch's feedback values when available).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .fused import CONFIG, _on, _ptr, _stream

_F32, _I32, _I64, _U8 = torch.float32, torch.int32, torch.int64, torch.uint8

_ws_bytes_cache: dict = {}
LAST_VIEW_SIZES: list = []      # (pairs, depth-key bits) of the views that went through the synchronising forward (capacity probe)


def _query_bytes(fn: str, *args) -> int:
    key = (fn,) + args
    v = _ws_bytes_cache.get(key)
    if v is None:
        n = ctypes.c_size_t(0)
        _lib.call(fn, *args, ctypes.byref(n))
        v = _ws_bytes_cache[key] = int(n.value)
    return v


def _round_up(n: int, m: int) -> int:
    return ((n + m - 1) // m) * m


@dataclass
class ViewState:
    """Everything the backward of one view needs (the reference keeps the same set alive through
    ctx.save_for_backward in wrapper.py:469,815)."""
    sh_degree: int
    hw: tuple
    tile: tuple
    n_chunks_visible: int
    n_pairs: int
    chunk_ids: torch.Tensor          # i64[M], first n_chunks_visible valid, ascending
    counters: torch.Tensor           # i32[4] = (visible chunks, pairs, ~min depth key, max depth key) on device
    view: torch.Tensor
    proj: torch.Tensor
    packed: torch.Tensor             # f32[1, Nv, 12]
    tile_count: torch.Tensor         # i32[Nv]
    sorted_pid: torch.Tensor         # i32[1, D]
    ranges: torch.Tensor             # i32[1, tiles+2]
    T: torch.Tensor                  # f32[1,1,Hp,Wp]
    last: torch.Tensor               # i16[1,1,Hp,Wp] (unsigned 16-bit counts)
    tile_order: Optional[torch.Tensor] = None   # i32[1,tiles]: tile ids, heaviest backward work first


class _Pinned:
    """Per-device pinned int32[4] used for the size read-back (visible chunks, pairs, ~min depth key, max depth key)."""
    _bufs: dict = {}

    @classmethod
    def get(cls, dev) -> torch.Tensor:
        b = cls._bufs.get(dev)
        if b is None:
            b = cls._bufs[dev] = torch.zeros(4, dtype=_I32).pin_memory()
        return b


def render_view_forward(params: dict, cluster_origin: torch.Tensor, cluster_extend: torch.Tensor, frustumplane: torch.Tensor,
                        view_matrix: torch.Tensor, proj_matrix: torch.Tensor, sh_degree: int, hw: tuple, tile: tuple,
                        enable_statistic: bool = False, specific_tiles: Optional[torch.Tensor] = None, clamp_zero: bool = False):
    """Forward of one view.  params: xyz[3,C,S] scale[3,C,S] rot[4,C,S] sh_0[1,3,C,S] sh_rest[R,3,C,S]
    opacity[1,C,S] (raw, clustered; float32 CUDA, contiguous).  Returns (img f32[1,3,Hp,Wp] padded to whole
    tiles, ViewState, (fragment_count, fragment_weight) or None)."""
    xyz = params["xyz"]
    dev = xyz.device
    C, S = xyz.shape[-2:]
    H, W = int(hw[0]), int(hw[1])
    th, tw = int(tile[0]), int(tile[1])
    if view_matrix.shape[0] != 1:
        raise RuntimeError("the fused pipeline renders one view per call (loop over views on the host)")
    for k in ("xyz", "scale", "rot", "sh_0", "sh_rest", "opacity"):
        t = params[k]
        if not (t.is_cuda and t.dtype == _F32 and t.is_contiguous()):
            raise RuntimeError(f"params['{k}'] must be a contiguous float32 CUDA tensor")
    gx, gy = (W + tw - 1) // tw, (H + th - 1) // th
    ntile = gx * gy
    Hp, Wp = gy * th, gx * tw
    M = C
    with _on(dev):
        st = _stream(dev)
        counters = torch.empty(4, dtype=_I32, device=dev)
        vis = torch.empty(M, dtype=_U8, device=dev)
        ids = torch.empty(M, dtype=_I64, device=dev)
        _lib.call("lgs_frustum_culling_aabb", _ptr(cluster_origin), _ptr(cluster_extend), _ptr(frustumplane), M, 1, _ptr(vis),
                  ctypes.c_void_p(counters.data_ptr()), _ptr(ids), st)
        Nmax = M * S
        packed = torch.empty((1, Nmax, 12), dtype=_F32, device=dev)
        dkey = torch.empty(Nmax, dtype=_I32, device=dev)
        iota = torch.empty(Nmax, dtype=_I32, device=dev)
        tcount = torch.empty(Nmax, dtype=_I32, device=dev)
        _lib.call("lgs_project_forward", int(sh_degree), _ptr(ids), ctypes.c_void_p(counters.data_ptr()), _ptr(view_matrix),
                  _ptr(proj_matrix), _ptr(xyz), _ptr(params["scale"]), _ptr(params["rot"]), _ptr(params["sh_0"]),
                  _ptr(params["sh_rest"]), _ptr(params["opacity"]), C, S, M, H, W, th, tw, _ptr(packed), _ptr(dkey), _ptr(iota),
                  _ptr(tcount), ctypes.c_void_p(counters.data_ptr() + 4), st)
        pinned = _Pinned.get(dev)
        pinned.copy_(counters, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        nvis, D = int(pinned[0]), int(pinned[1])
        kmin, kmax = ~int(pinned[2]) & 0xFFFFFFFF, int(pinned[3]) & 0xFFFFFFFF
        Nv = nvis * S
        LAST_VIEW_SIZES.append((D, max(1, (kmax - kmin).bit_length()) if (D > 0 and kmax >= kmin) else 1))
        if len(LAST_VIEW_SIZES) > 4096:
            del LAST_VIEW_SIZES[:2048]

        ranges = torch.empty((1, ntile + 2), dtype=_I32, device=dev)
        if D > 0:
            # depth order of the Nv live slots: stable LSD radix sort on the float bits of view z, rebased to the
            # smallest key of a splat that owns pairs and limited to the bits of the key range (z in [1.3, 4.7) spans 31
            # bits of float pattern but 24 bits of range: 3 passes instead of 4); where a splat without pairs lands in
            # the order is irrelevant, it emits nothing
            depth_bits = max(1, (kmax - kmin).bit_length()) if kmax >= kmin else 1
            nb = _query_bytes("lgs_sort_pairs_u32_workspace_bytes", _round_up(Nv, 1 << 16))
            ws = torch.empty(nb, dtype=_U8, device=dev)
            dkey_s = torch.empty(Nv, dtype=_I32, device=dev)
            order = torch.empty(Nv, dtype=_I32, device=dev)
            _lib.call("lgs_sort_pairs_u32_rebased", _ptr(dkey), _ptr(dkey_s), _ptr(iota), _ptr(order), Nv, kmin, depth_bits, _ptr(ws),
                      ctypes.c_size_t(nb), st)
            nb2 = _query_bytes("lgs_scan_gathered_workspace_bytes", _round_up(Nv, 1 << 16))
            ws2 = ws if nb2 <= nb else torch.empty(nb2, dtype=_U8, device=dev)
            offsets = dkey_s     # the sorted keys are dead now: reuse their storage for the scan
            _lib.call("lgs_scan_gathered", _ptr(tcount), _ptr(order), Nv, _ptr(offsets), _ptr(ws2), ctypes.c_size_t(max(nb, nb2)), st)
            bits = ntile.bit_length()          # floor(log2(tiles)) + 1, GR/binning.cu:199-202
            u16 = (ntile + 1) < 65536          # 16-bit tile keys whenever they fit (up to 4K at 8x16)
            kdt = torch.int16 if u16 else _I32
            sfx = "_u16" if u16 else "_u32"
            keys = torch.empty(D, dtype=kdt, device=dev)
            vals = torch.empty(D, dtype=_I32, device=dev)
            _lib.call("lgs_emit_pairs_u16" if u16 else "lgs_emit_pairs", _ptr(packed), _ptr(offsets), _ptr(order), Nv, D, H, W, th, tw,
                      _ptr(keys), _ptr(vals), st)
            nb3 = _query_bytes(f"lgs_sort_pairs{sfx}_workspace_bytes", _round_up(D, 1 << 18))
            ws3 = ws if nb3 <= nb else torch.empty(nb3, dtype=_U8, device=dev)
            keys_s = torch.empty((1, D), dtype=kdt, device=dev)
            sorted_pid = torch.empty((1, D), dtype=_I32, device=dev)
            _lib.call(f"lgs_sort_pairs{sfx}", _ptr(keys), _ptr(keys_s), _ptr(vals), _ptr(sorted_pid), D, 0, bits, _ptr(ws3),
                      ctypes.c_size_t(max(nb, nb3)), st)
            _lib.call("lgs_tile_range_u16" if u16 else "lgs_tile_range", _ptr(keys_s), 1, D, ntile, 1, _ptr(ranges), st)
        else:
            sorted_pid = torch.zeros((1, 1), dtype=_I32, device=dev)
            _lib.call("lgs_tile_range", None, 1, 0, ntile, 1, _ptr(ranges), st)

        img = torch.empty((1, 3, Hp, Wp), dtype=_F32, device=dev)
        T = torch.empty((1, 1, Hp, Wp), dtype=_F32, device=dev)
        last = torch.empty((1, 1, Hp, Wp), dtype=torch.int16, device=dev)
        n_sel = 0
        if specific_tiles is not None:
            n_sel = specific_tiles.shape[1]
            img.zero_(); T.fill_(1.0); last.zero_()
        stats = None
        fc = fw = None
        if enable_statistic:
            fc = torch.zeros((1, 1, Nmax), dtype=_I32, device=dev)
            fw = torch.zeros((1, 1, Nmax), dtype=_F32, device=dev)
            stats = (fc, fw)
        # per-tile trip count of the backward (deepest list position any pixel consumed) -> heaviest-first tile order
        order = None
        work = torch.empty((1, ntile), dtype=_I32, device=dev) if (CONFIG["tile_order"] and specific_tiles is None and D > 0) else None
        _lib.call("lgs_rasterize_forward_packed", _ptr(sorted_pid), _ptr(ranges), _ptr(packed), _ptr(specific_tiles), n_sel, 1,
                  Nmax, sorted_pid.shape[1], H, W, th, tw, int(bool(enable_statistic)), int(bool(clamp_zero)), _ptr(img), _ptr(T),
                  _ptr(last), _ptr(fc), _ptr(fw), _ptr(work), st)
        if work is not None:
            order = torch.empty((1, ntile), dtype=_I32, device=dev)
            _lib.call("lgs_tile_order", _ptr(work), 1, ntile, _ptr(order), st)
    state = ViewState(sh_degree=int(sh_degree), hw=(H, W), tile=(th, tw), n_chunks_visible=nvis, n_pairs=D, chunk_ids=ids,
                      counters=counters, view=view_matrix, proj=proj_matrix, packed=packed, tile_count=tcount,
                      sorted_pid=sorted_pid, ranges=ranges, T=T, last=last, tile_order=order)
    return img, state, stats


def render_view_backward(params: dict, state: ViewState, d_img: torch.Tensor, d_trans: Optional[torch.Tensor] = None,
                         enable_statistic: bool = False, specific_tiles: Optional[torch.Tensor] = None,
                         accumulate_into: Optional[dict] = None, clamped_img: Optional[torch.Tensor] = None):
    """Backward of one view: d_img f32[1,3,Hp,Wp] (padded) -> compacted parameter gradients
    (xyz[3,A,S], scale[3,A,S], rot[4,A,S], sh_0[1,3,A,S], sh_rest[R,3,A,S], opacity[1,A,S]) with
    A = state.n_chunks_visible, plus packed_grad (whose slot 9 carries the statistics term).

    accumulate_into: dict of DENSE contiguous gradient tensors shaped like the parameters; when given, this
    view's gradients are added into them by the kernel itself and no compacted tensors are produced
    (returns (None, packed_grad)).  An optional "_touched" entry (f32[C]) receives 1 at every visible chunk."""
    xyz = params["xyz"]
    dev = xyz.device
    C, S = xyz.shape[-2:]
    H, W = state.hw
    th, tw = state.tile
    A = state.n_chunks_visible
    R = params["sh_rest"].shape[0]
    Nmax = state.packed.shape[1]
    if not (d_img.is_cuda and d_img.dtype == _F32):
        raise RuntimeError("d_img must be a float32 CUDA tensor")
    d_img = d_img if d_img.is_contiguous() else d_img.contiguous()
    if d_trans is not None:
        d_trans = d_trans if d_trans.is_contiguous() else d_trans.contiguous()
    with _on(dev):
        st = _stream(dev)
        pg = torch.empty((1, Nmax, 12), dtype=_F32, device=dev)
        if specific_tiles is None:
            specific_tiles = state.tile_order          # every tile, longest lists first (None = index order)
        n_sel = 0 if specific_tiles is None else specific_tiles.shape[1]
        _lib.call("lgs_rasterize_backward", _ptr(state.sorted_pid), _ptr(state.ranges), _ptr(state.packed), _ptr(specific_tiles), n_sel,
                  _ptr(state.T), _ptr(state.last), _ptr(d_img), _ptr(d_trans), _ptr(clamped_img), None, 1, Nmax, state.sorted_pid.shape[1],
                  H, W, th, tw,
                  int(bool(enable_statistic)), _ptr(pg), None, None, None, None, None, None, st)
        if accumulate_into is not None:
            for k in ("xyz", "scale", "rot", "sh_0", "sh_rest", "opacity"):
                t = accumulate_into[k]
                if not (t.is_cuda and t.dtype == _F32 and t.is_contiguous() and tuple(t.shape) == tuple(params[k].shape)):
                    raise RuntimeError(f"accumulate_into['{k}'] must be a contiguous float32 CUDA tensor shaped like the parameter")
            if A > 0:
                d = accumulate_into
                _lib.call("lgs_project_backward", state.sh_degree, _ptr(state.chunk_ids), ctypes.c_void_p(state.counters.data_ptr()),
                          _ptr(state.view), _ptr(state.proj), _ptr(xyz), _ptr(params["scale"]), _ptr(params["rot"]),
                          _ptr(params["opacity"]), C, S, A, R, H, W, int(CONFIG["true_sigmoid_grad"]), _ptr(pg), None, 2,
                          _ptr(d["xyz"]), _ptr(d["scale"]), _ptr(d["rot"]), _ptr(d["sh_0"]), _ptr(d["sh_rest"]), _ptr(d["opacity"]),
                          _ptr(d.get("_touched")), st)   # "_touched": chunk marks for the fused optimizer step
            return None, pg
        g_pos = torch.empty((3, A, S), dtype=_F32, device=dev)
        g_sc = torch.empty((3, A, S), dtype=_F32, device=dev)
        g_rot = torch.empty((4, A, S), dtype=_F32, device=dev)
        g_s0 = torch.empty((1, 3, A, S), dtype=_F32, device=dev)
        K = (state.sh_degree + 1) ** 2
        g_sr = torch.zeros((R, 3, A, S), dtype=_F32, device=dev) if R > K - 1 else torch.empty((R, 3, A, S), dtype=_F32, device=dev)
        g_op = torch.empty((1, A, S), dtype=_F32, device=dev)
        if A > 0:
            _lib.call("lgs_project_backward", state.sh_degree, _ptr(state.chunk_ids), ctypes.c_void_p(state.counters.data_ptr()),
                      _ptr(state.view), _ptr(state.proj), _ptr(xyz), _ptr(params["scale"]), _ptr(params["rot"]), _ptr(params["opacity"]),
                      C, S, A, R, H, W, int(CONFIG["true_sigmoid_grad"]), _ptr(pg), None, 0, _ptr(g_pos), _ptr(g_sc), _ptr(g_rot),
                      _ptr(g_s0), _ptr(g_sr), _ptr(g_op), None, st)
    return [g_pos, g_sc, g_rot, g_s0, g_sr, g_op], pg


# ---------------------------------------------------------------------------------------------------
# GPU-driven path: preallocated workspace, no host synchronisation, optionally replayed as CUDA graphs
# ---------------------------------------------------------------------------------------------------

GRAPHS_ENABLED = os.environ.get("LGS_GRAPHS", "1") != "0"      # replay the per-view forward / backward as CUDA graphs
SYNC_FREE = os.environ.get("LGS_SYNC_FREE", "1") != "0"        # render_views: GPU-driven sizing on preallocated workspaces


class ViewWorkspace:
    """Every buffer one view needs, allocated once for fixed capacities, plus (optionally) the view's forward and backward
    captured as CUDA graphs.

    The forward is enqueued without reading anything back: the sorts, the scan, the pair emission and the tile ranges take the
    live counts from device memory (``lgs_*_dev`` entry points), the number of (tile, splat) pairs is bounded by
    ``pair_capacity`` and the depth sort by ``planned_depth_bits``.  ``lgs_view_params`` raises a device-side flag when either
    prediction was too small; :meth:`check` (called by the owner when it synchronises anyway, e.g. at the end of a step) reads it
    and raises :class:`CapacityExceeded` so that the caller can grow the workspace and redo the step.  This is the reference's
    "size from the previous epoch, write feedback for the next" protocol (GR/compact.cu:527-549, GR/binning.cu:137-163,
    data.py:238) moved onto the device.

    A workspace serves ONE view at a time: its state (lists, transmittance, counts) is consumed by the backward of the same view
    before the next forward on it; ``render_views`` keeps one workspace per stream slot."""

    def __init__(self, params: dict, hw: tuple, tile: tuple, pair_capacity: int, planned_depth_bits: int = 24, use_graphs: bool = True):
        xyz = params["xyz"]
        dev = xyz.device
        C, S = xyz.shape[-2:]
        H, W = int(hw[0]), int(hw[1])
        th, tw = int(tile[0]), int(tile[1])
        gx, gy = (W + tw - 1) // tw, (H + th - 1) // th
        self.dev, self.C, self.S, self.hw, self.tile = dev, C, S, (H, W), (th, tw)
        self.ntile, self.Hp, self.Wp = gx * gy, gy * th, gx * tw
        self.Nmax = C * S
        self.cap = int(max(1024, pair_capacity))
        self.planned_bits = int(planned_depth_bits)
        self.use_graphs = bool(use_graphs)
        self.u16 = (self.ntile + 1) < 65536
        kdt = torch.int16 if self.u16 else _I32
        N, D = self.Nmax, self.cap
        e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dev)
        self.counters = torch.zeros(4, dtype=_I32, device=dev)
        self.vparams = torch.zeros(8, dtype=_I32, device=dev)
        self.vis, self.chunk_ids = e(C, _U8), e(C, _I64)
        self.packed, self.dkey, self.iota, self.tcount = e((1, N, 12), _F32), e(N, _I32), e(N, _I32), e(N, _I32)
        self.dkey_s, self.order = e(N, _I32), e(N, _I32)
        self.keys, self.vals, self.keys_s, self.sorted_pid = e(D, kdt), e(D, _I32), e((1, D), kdt), e((1, D), _I32)
        self.ranges = e((1, self.ntile + 2), _I32)
        self.img, self.T = e((1, 3, self.Hp, self.Wp), _F32), e((1, 1, self.Hp, self.Wp), _F32)
        self.last = e((1, 1, self.Hp, self.Wp), torch.int16)
        self.work, self.tile_order = e((1, self.ntile), _I32), e((1, self.ntile), _I32)
        self.pg = e((1, N, 12), _F32)
        self.d_img = e((1, 3, self.Hp, self.Wp), _F32)
        self.cam_view, self.cam_proj, self.cam_planes = e((1, 4, 4), _F32), e((1, 4, 4), _F32), e((1, 6, 4), _F32)
        nb = max(_query_bytes("lgs_sort_pairs_u32_workspace_bytes", _round_up(N, 1 << 16)),
                 _query_bytes("lgs_scan_gathered_workspace_bytes", _round_up(N, 1 << 16)),
                 _query_bytes(f"lgs_sort_pairs_{'u16' if self.u16 else 'u32'}_workspace_bytes", _round_up(D, 1 << 18)))
        self.ws, self.ws_bytes = e(nb, _U8), nb
        self.sticky = torch.zeros(4, dtype=_I32, device=dev)      # |flags, max pairs, max depth bits, views since the last check
        self.sticky_host = torch.zeros(4, dtype=_I32).pin_memory()
        self.sticky_event = None
        self._graphs = {}               # ("fwd"|"bwd", pointer signature) -> torch.cuda.CUDAGraph
        self._eager_runs = {}           # same key -> eager runs so far (the first run of a signature is never captured)
        self.views_done = 0

    # -- enqueue ---------------------------------------------------------------------------------------------------
    def _forward_kernels(self, params, cluster_origin, cluster_extend, sh_degree, clamp_zero):
        dev, st = self.dev, _stream(self.dev)
        H, W = self.hw
        th, tw = self.tile
        C, S, N, D = self.C, self.S, self.Nmax, self.cap
        cnt = self.counters.data_ptr()
        vp = self.vparams.data_ptr()
        _lib.call("lgs_frustum_culling_aabb", _ptr(cluster_origin), _ptr(cluster_extend), _ptr(self.cam_planes), C, 1, _ptr(self.vis),
                  ctypes.c_void_p(cnt), _ptr(self.chunk_ids), st)
        _lib.call("lgs_project_forward", int(sh_degree), _ptr(self.chunk_ids), ctypes.c_void_p(cnt), _ptr(self.cam_view), _ptr(self.cam_proj),
                  _ptr(params["xyz"]), _ptr(params["scale"]), _ptr(params["rot"]), _ptr(params["sh_0"]), _ptr(params["sh_rest"]),
                  _ptr(params["opacity"]), C, S, C, H, W, th, tw, _ptr(self.packed), _ptr(self.dkey), _ptr(self.iota), _ptr(self.tcount),
                  ctypes.c_void_p(cnt + 4), st)
        _lib.call("lgs_view_params", ctypes.c_void_p(cnt), S, D, self.planned_bits, ctypes.c_void_p(vp), _ptr(self.sticky), st)
        n_dev, d_dev, bias_dev = ctypes.c_void_p(vp), ctypes.c_void_p(vp + 4), ctypes.c_void_p(vp + 8)
        wsz = ctypes.c_size_t(self.ws_bytes)
        _lib.call("lgs_sort_pairs_u32_dev", _ptr(self.dkey), _ptr(self.dkey_s), _ptr(self.iota), _ptr(self.order), N, n_dev, bias_dev,
                  self.planned_bits, _ptr(self.ws), wsz, st)
        offsets = self.dkey_s           # the sorted keys are dead: their storage receives the scan
        _lib.call("lgs_scan_gathered_dev", _ptr(self.tcount), _ptr(self.order), N, n_dev, _ptr(offsets), _ptr(self.ws), wsz, st)
        _lib.call("lgs_emit_pairs_dev", _ptr(self.packed), _ptr(offsets), _ptr(self.order), N, n_dev, D, H, W, th, tw, 16 if self.u16 else 32,
                  _ptr(self.keys), _ptr(self.vals), d_dev, st)
        bits = self.ntile.bit_length()
        _lib.call("lgs_sort_pairs_u16_dev" if self.u16 else "lgs_sort_pairs_u32k_dev", _ptr(self.keys), _ptr(self.keys_s), _ptr(self.vals),
                  _ptr(self.sorted_pid), D, d_dev, 0, bits, _ptr(self.ws), wsz, st)
        _lib.call("lgs_tile_range_u16_dev" if self.u16 else "lgs_tile_range_dev", _ptr(self.keys_s), D, d_dev, self.ntile,
                  int(CONFIG["fix_last_tile"]), _ptr(self.ranges), st)
        order = CONFIG["tile_order"]
        _lib.call("lgs_rasterize_forward_packed", _ptr(self.sorted_pid), _ptr(self.ranges), _ptr(self.packed), None, 0, 1, N, D, H, W, th, tw,
                  0, int(bool(clamp_zero)), _ptr(self.img), _ptr(self.T), _ptr(self.last), None, None, _ptr(self.work) if order else None, st)
        if order:
            _lib.call("lgs_tile_order", _ptr(self.work), 1, self.ntile, _ptr(self.tile_order), st)

    def _backward_kernels(self, params, sh_degree, accumulate_into, use_clamp):
        st = _stream(self.dev)
        H, W = self.hw
        th, tw = self.tile
        C, S, N, D = self.C, self.S, self.Nmax, self.cap
        tiles = self.tile_order if CONFIG["tile_order"] else None
        _lib.call("lgs_rasterize_backward", _ptr(self.sorted_pid), _ptr(self.ranges), _ptr(self.packed), _ptr(tiles),
                  self.ntile if tiles is not None else 0, _ptr(self.T), _ptr(self.last), _ptr(self.d_img), None,
                  _ptr(self.img) if use_clamp else None, None, 1, N, D, H, W, th, tw, 0, _ptr(self.pg), None, None, None, None, None, None, st)
        d = accumulate_into
        R = params["sh_rest"].shape[0]
        # A = all chunks: project_backward returns at once for chunks past the (device) visible count
        _lib.call("lgs_project_backward", int(sh_degree), _ptr(self.chunk_ids), ctypes.c_void_p(self.counters.data_ptr()), _ptr(self.cam_view),
                  _ptr(self.cam_proj), _ptr(params["xyz"]), _ptr(params["scale"]), _ptr(params["rot"]), _ptr(params["opacity"]), C, S, C, R,
                  H, W, int(CONFIG["true_sigmoid_grad"]), _ptr(self.pg), None, 2, _ptr(d["xyz"]), _ptr(d["scale"]), _ptr(d["rot"]),
                  _ptr(d["sh_0"]), _ptr(d["sh_rest"]), _ptr(d["opacity"]), _ptr(d.get("_touched")), st)

    def _run(self, kind, sig, fn):
        """Eager the first time a pointer signature is seen, captured into a CUDA graph the second time, replayed afterwards."""
        key = (kind, sig)
        graphs = self.use_graphs and GRAPHS_ENABLED          # GRAPHS_ENABLED: module-wide switch (stage timing, A/B)
        g = self._graphs.get(key) if graphs else None
        if g is not None:
            g.replay()
            return
        seen = self._eager_runs.get(key, 0)
        cur = torch.cuda.current_stream(self.dev)
        # graphs cannot be captured on the legacy default stream
        if not graphs or seen < 1 or cur.cuda_stream == 0:
            self._eager_runs[key] = seen + 1
            fn()
            return
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=cur, capture_error_mode="thread_local"):
            fn()
        self._graphs[key] = g
        g.replay()

    def forward(self, params, cluster_origin, cluster_extend, cam, sh_degree, clamp_zero=True):
        """cam: dict(view, proj, frustumplane) of device tensors.  Returns the padded image (a view of the workspace)."""
        self.cam_view.copy_(cam["view"], non_blocking=True)
        self.cam_proj.copy_(cam["proj"], non_blocking=True)
        self.cam_planes.copy_(cam["frustumplane"], non_blocking=True)
        sig = (tuple(params[k].data_ptr() for k in ("xyz", "scale", "rot", "sh_0", "sh_rest", "opacity")), cluster_origin.data_ptr(),
               cluster_extend.data_ptr(), int(sh_degree), bool(clamp_zero), bool(CONFIG["tile_order"]))
        self._run("fwd", sig, lambda: self._forward_kernels(params, cluster_origin, cluster_extend, sh_degree, clamp_zero))
        self.views_done += 1
        return self.img

    def backward(self, params, d_img, sh_degree, accumulate_into, use_clamp=True):
        """d_img f32[1,3,H,W] or [1,3,Hp,Wp]: gradient of the loss w.r.t. the (clamped) image."""
        H, W = self.hw
        if d_img.shape[-2:] == (self.Hp, self.Wp):
            self.d_img.copy_(d_img, non_blocking=True)
        else:
            if self.Hp != H or self.Wp != W:
                self.d_img.zero_()
            self.d_img[..., :H, :W].copy_(d_img, non_blocking=True)
        sig = (tuple(params[k].data_ptr() for k in ("xyz", "scale", "rot", "sh_0", "sh_rest", "opacity")),
               tuple(accumulate_into[k].data_ptr() for k in ("xyz", "scale", "rot", "sh_0", "sh_rest", "opacity")),
               0 if accumulate_into.get("_touched") is None else accumulate_into["_touched"].data_ptr(), int(sh_degree), bool(use_clamp),
               bool(CONFIG["tile_order"]))
        self._run("bwd", sig, lambda: self._backward_kernels(params, sh_degree, accumulate_into, use_clamp))

    # -- feedback --------------------------------------------------------------------------------------------------
    def post_flags(self):
        """Enqueue (on the current stream) the copy of the sticky overflow word to pinned memory and reset it on the device."""
        self.sticky_host.copy_(self.sticky, non_blocking=True)
        self.sticky.zero_()
        self.sticky_event = torch.cuda.Event()
        self.sticky_event.record(torch.cuda.current_stream(self.dev))

    def check(self, wait: bool = False):
        """Result of the last post_flags(): None if it has not landed yet (and wait is False); otherwise a dict, or
        CapacityExceeded if a view since the previous check overflowed a prediction (its lists were truncated)."""
        ev = self.sticky_event
        if ev is None:
            return None
        if not ev.query():
            if not wait:
                return None
            ev.synchronize()
        self.sticky_event = None
        flags, pairs, bits, views = (int(x) for x in self.sticky_host)
        if flags:
            raise CapacityExceeded(pairs=pairs, pair_capacity=self.cap, depth_bits=bits, planned_depth_bits=self.planned_bits)
        return {"max_pairs": pairs, "max_depth_bits": bits, "views": views}


class CapacityExceeded(RuntimeError):
    def __init__(self, pairs, pair_capacity, depth_bits, planned_depth_bits):
        super().__init__(f"view workspace too small: {pairs} (tile, splat) pairs for a capacity of {pair_capacity}, depth keys span {depth_bits} "
                         f"bits with {planned_depth_bits} planned; grow the workspace and redo the step")
        self.pairs, self.pair_capacity, self.depth_bits, self.planned_depth_bits = pairs, pair_capacity, depth_bits, planned_depth_bits


def probe_view_sizes(params, cluster_origin, cluster_extend, cams, sh_degree, hw, tile):
    """One synchronising forward per camera (the cold path) -> (max pairs, max depth-key bits): the first-epoch sizing step of
    the reference's feedback protocol, used to dimension a ViewWorkspace."""
    max_pairs, max_bits = 0, 1
    p = {k: params[k].detach() for k in ("xyz", "scale", "rot", "sh_0", "sh_rest", "opacity")}
    with torch.no_grad():
        for cam in cams:
            _, st, _ = render_view_forward(p, cluster_origin, cluster_extend, cam["frustumplane"], cam["view"], cam["proj"], sh_degree, hw, tile)
            max_pairs = max(max_pairs, st.n_pairs)
            c = st.counters.cpu()
            kmin, kmax = ~int(c[2]) & 0xFFFFFFFF, int(c[3]) & 0xFFFFFFFF
            if st.n_pairs > 0 and kmax >= kmin:
                max_bits = max(max_bits, max(1, (kmax - kmin).bit_length()))
    return max_pairs, max_bits
