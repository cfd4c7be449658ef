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
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .fused import CONFIG, _on, _ptr, _stream

_F32, _I32, _I64, _U8 = torch.float32, torch.int32, torch.int64, torch.uint8

_ws_bytes_cache: dict = {}


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
