"""The render operator: ``render_preprocess`` + ``render`` with the signatures of the reference's
``litegs/render/__init__.py:11-94`` (Level A: op by op through ``litegs_b200.wrapper``), and
``render_view`` -- the same computation as ONE differentiable call on the fused pipeline (Level B).

Both levels produce the same image and the same six parameter gradients (tests/test_gpu_pipeline.py);
Level B is what ``bench.py`` measures.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch
import torch.cuda.nvtx as nvtx

from . import pipeline, wrapper
from .compacted import CompactedTensor
from .statistics import StatisticsHelperInst


def uncluster(*tensors):
    """[..., chunks, chunk_size] -> [..., chunks*chunk_size] views (reference scene/cluster.py:24-28)."""
    return tuple(t.reshape(*t.shape[:-2], t.shape[-2] * t.shape[-1]) for t in tensors)


# ---------------------------------------------------------------------------------------------------
# Level A: the reference's two-call surface
# ---------------------------------------------------------------------------------------------------

def render_preprocess(cluster_origin, cluster_extend, frustumplane, view_matrix,
                      xyz, scale, rot, sh_0, sh_rest, opacity,
                      feedback_buffer, idx_tensor, pp, actived_sh_degree: int):
    """Chunk frustum culling -> compaction -> activation (+ SH->RGB).

    Returns (visible_chunkid, visible_chunks_num, culled_xyz [4,N], culled_scale [3,N], culled_rot [4,N],
    color [V,3,N], culled_opacity [1,N]) exactly as render/__init__.py:11-48."""
    visible_chunkid = None
    visible_chunks_num = None
    if pp.cluster_size:
        if cluster_origin is None or cluster_extend is None:
            raise RuntimeError("cluster_origin / cluster_extend are required when cluster_size > 0 "
                               "(compute them with litegs_b200.scene.cluster_aabb or the caller's scene code)")
        _, visible_chunks_num, visible_chunkid = wrapper.litegs_fused.frustum_culling_aabb(cluster_origin, cluster_extend, frustumplane,
                                                                                 feedback_buffer, idx_tensor)
        if StatisticsHelperInst.bStart and StatisticsHelperInst.on_compact_mask is not None:
            StatisticsHelperInst.on_compact_mask(visible_chunkid, visible_chunks_num)
        culled = wrapper.CullCompactActivateWithSparseGrad.apply(pp.sparse_grad, actived_sh_degree, visible_chunkid,
                                                                 visible_chunks_num, view_matrix, xyz, scale, rot, sh_0, sh_rest, opacity)
        culled_xyz, culled_scale, culled_rot, color, culled_opacity = uncluster(*culled)
    else:
        nvtx.range_push("Activate")
        ones = torch.ones((1, xyz.shape[-1]), dtype=xyz.dtype, device=xyz.device)
        culled_xyz = torch.cat((xyz, ones), dim=0)
        culled_scale = scale.exp()
        culled_rot = torch.nn.functional.normalize(rot, dim=0)
        culled_opacity = opacity.sigmoid()
        with torch.no_grad():
            R = view_matrix[..., :3, :3]
            t = view_matrix[..., 3:4, :3]
            camera_center = (-t @ R.transpose(-1, -2)).squeeze(1)
            dirs = torch.nn.functional.normalize(culled_xyz[:3] - camera_center.unsqueeze(-1), dim=-2)
        color = wrapper.SphericalHarmonicToRGB.call_fused(actived_sh_degree, sh_0, sh_rest, dirs)
        nvtx.range_pop()
    return visible_chunkid, visible_chunks_num, culled_xyz, culled_scale, culled_rot, color, culled_opacity


def render(view_matrix, proj_matrix, xyz, scale, rot, color, opacity,
           valid_length, feedback_binning_allocate_size, idx_tensor,
           actived_sh_degree: int, output_shape, pp):
    """Projection -> binning -> rasterisation; returns (img, transmittance, depth, normal, primitive_visible)
    as render/__init__.py:50-94."""
    nvtx.range_push("Proj")
    view_pos, ndc_pos = wrapper.MVPTransform.apply(xyz, view_matrix, proj_matrix, valid_length)
    transform_matrix = wrapper.CreateTransformMatrix.call_fused(scale, rot, valid_length)
    J = wrapper.CreateRaySpaceTransformMatrix.call_fused(view_pos, proj_matrix, output_shape, valid_length)
    cov2d = wrapper.CreateCov2dDirectly.call_fused(J, view_matrix, transform_matrix, valid_length)
    _, _, inv_cov2d = wrapper.EighAndInverse2x2Matrix.call_fused(cov2d, valid_length)
    view_depth = view_pos[:, 2, :]
    nvtx.range_pop()

    tile_start_index, sorted_pointId, primitive_visible = wrapper.Binning.call_fused(
        ndc_pos, view_depth, inv_cov2d, opacity, valid_length, feedback_binning_allocate_size, idx_tensor,
        output_shape, pp.tile_size)

    tiles = None
    cached = StatisticsHelperInst.cached_sorted_tile_list.get(StatisticsHelperInst.cur_sample)
    if cached is not None:
        tiles = cached.unsqueeze(0)
    H, W = int(output_shape[0]), int(output_shape[1])
    img, transmitance, depth, normal, last = wrapper.GaussiansRasterFunc.apply(
        sorted_pointId, tile_start_index, ndc_pos, inv_cov2d, color, opacity, tiles, H, W,
        pp.tile_size[0], pp.tile_size[1], pp.enable_transmitance, pp.enable_depth)
    if StatisticsHelperInst.bStart and StatisticsHelperInst.on_blend_count is not None:
        StatisticsHelperInst.on_blend_count(last, pp.tile_size[0], pp.tile_size[1])

    img = img[..., :H, :W].clamp(0, 1).contiguous()
    if transmitance is not None:
        transmitance = transmitance[..., :H, :W].contiguous()
    if depth is not None:
        depth = depth[..., :H, :W].contiguous()
    return img, transmitance, depth, normal, primitive_visible


# ---------------------------------------------------------------------------------------------------
# Level B: one differentiable call per view
# ---------------------------------------------------------------------------------------------------

@torch.no_grad()
def _feed_statistics(state, stats, packed_grad, tile):
    """What Level A feeds the StatisticsHelper op by op (render/__init__.py:24-25,84-85; wrapper.py:501-506,733-737), from the
    fused pipeline's state: compact mask + device-side visible count, visible splats, fragment weight / count, fragment error
    (d_opacity = sum s0 / o and the err_square term of the gradient record), per-tile blend counts."""
    SH = StatisticsHelperInst
    if SH.on_compact_mask is not None:
        SH.on_compact_mask(state.chunk_ids, state.counters[:1])
    if SH.on_visible is not None:
        SH.on_visible((state.tile_count > 0).reshape(1, -1))
    fc, fw = stats
    if SH.on_fragment_weight is not None:
        SH.on_fragment_weight(fw, fc)
    if SH.on_fragment_err is not None and packed_grad is not None:
        o = state.packed[..., 5]
        d_op = torch.where(o > 0, packed_grad[..., 8] / o.clamp_min(1e-30), torch.zeros_like(o))
        SH.on_fragment_err(d_op.unsqueeze(0), packed_grad[..., 9].unsqueeze(0), fc)
    if SH.on_blend_count is not None:
        SH.on_blend_count(state.last, tile[0], tile[1])


class _RenderViewFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, scale, rot, sh_0, sh_rest, opacity, cluster_origin, cluster_extend, frustumplane,
                view_matrix, proj_matrix, sh_degree, H, W, th, tw, sparse_grad, enable_transmitance, accumulate_into):
        params = dict(xyz=xyz, scale=scale, rot=rot, sh_0=sh_0, sh_rest=sh_rest, opacity=opacity)
        stat = bool(StatisticsHelperInst.bStart)
        ctx.set_materialize_grads(False)       # an unused transmittance output must not cost a zero-filled gradient image
        # the kernel writes clamp(c,0,1) directly (render/__init__.py:87 does it as a separate pass) ...
        img, state, stats = pipeline.render_view_forward(params, cluster_origin, cluster_extend, frustumplane, view_matrix,
                                                         proj_matrix, sh_degree, (H, W), (th, tw), enable_statistic=stat, clamp_zero=True)
        ctx.state = state
        ctx.stats = stats
        ctx.stat = stat
        ctx.sparse = bool(sparse_grad)
        ctx.trans = bool(enable_transmitance)
        ctx.accumulate_into = accumulate_into
        ctx.save_for_backward(xyz, scale, rot, sh_0, sh_rest, opacity, img)
        ctx.mark_non_differentiable(state.last)
        return img, state.T, state.last

    @staticmethod
    def backward(ctx, g_img, g_T, _g_last):
        xyz, scale, rot, sh_0, sh_rest, opacity, img_out = ctx.saved_tensors
        params = dict(xyz=xyz, scale=scale, rot=rot, sh_0=sh_0, sh_rest=sh_rest, opacity=opacity)
        state = ctx.state
        if g_img is None:
            g_img = torch.zeros((1, 3, *state.T.shape[-2:]), dtype=torch.float32, device=xyz.device)
        # ... and the backward kernel applies that clamp's gradient mask from the saved image
        grads, pg = pipeline.render_view_backward(params, state, g_img, g_T if (ctx.trans and g_T is not None) else None,
                                                  enable_statistic=ctx.stat,
                                                  accumulate_into=ctx.accumulate_into, clamped_img=img_out)
        if ctx.stat:
            _feed_statistics(state, ctx.stats, pg, state.tile)
        if grads is None:          # gradients went straight into the caller's dense buffers
            ctx.state = None
            return (None,) * 19
        C, S = xyz.shape[-2:]
        ids = state.chunk_ids[: state.n_chunks_visible]
        out = []
        for g in grads:
            ct = CompactedTensor((*g.shape[:-2], C, S), ids, g)
            out.append(ct if ctx.sparse else ct.to_dense())
        ctx.state = None
        return (*out, None, None, None, None, None, None, None, None, None, None, None, None, None)


def render_view(cluster_origin, cluster_extend, frustumplane, view_matrix, proj_matrix,
                xyz, scale, rot, sh_0, sh_rest, opacity, actived_sh_degree: int, output_shape, pp, accumulate_into=None):
    """render_preprocess + render of one view on the fused pipeline.

    Same inputs as the two reference calls (raw clustered parameters, chunk AABBs, camera); returns the five values of the
    reference's render() with the contributor counts in the last slot: (img [1,3,H,W] clamped to [0,1], transmittance or
    None, depth=None, normal=None, last_contributor [1,1,Hp,Wp]).  Gradients reach the six parameter tensors as CompactedTensor (pp.sparse_grad) or dense
    tensors -- or, with ``accumulate_into`` (dict of dense gradient tensors, e.g. ``GradAccumulator.grads()``), are
    ADDED into those buffers by the backward kernel itself and ``param.grad`` stays untouched (multi-view batches,
    data-parallel training)."""
    if not pp.cluster_size:
        raise RuntimeError("render_view needs the clustered layout (cluster_size > 0); use render_preprocess + render otherwise")
    H, W = int(output_shape[0]), int(output_shape[1])
    th, tw = int(pp.tile_size[0]), int(pp.tile_size[1])
    img, T, last = _RenderViewFn.apply(xyz, scale, rot, sh_0, sh_rest, opacity, cluster_origin, cluster_extend, frustumplane,
                                       view_matrix, proj_matrix, int(actived_sh_degree), H, W, th, tw, pp.sparse_grad,
                                       pp.enable_transmitance, accumulate_into)
    img = img[..., :H, :W]          # already clamped to [0,1] by the kernel
    trans = T[..., :H, :W] if pp.enable_transmitance else None
    return img, trans, None, None, last


# ---------------------------------------------------------------------------------------------------
# multi-view micro-batch: views pipelined over CUDA streams, gradients summed in dense buffers
# ---------------------------------------------------------------------------------------------------

_side_streams: dict = {}
# render_views drives the pipeline's forward/backward directly and uses autograd only for the caller's loss (default);
# LGS_VIEWS_AUTOGRAD=1 routes every view through the render_view autograd Function instead (A/B switch)
_DIRECT_VIEWS = os.environ.get("LGS_VIEWS_AUTOGRAD", "0") != "1"


def _streams(dev, n):
    key = (dev, n)
    if key not in _side_streams:
        _side_streams[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
    return _side_streams[key]


def render_views(n_views: int, camera_fn, loss_fn, cluster_origin, cluster_extend,
                 xyz, scale, rot, sh_0, sh_rest, opacity, actived_sh_degree: int, output_shape, pp,
                 accumulate_into: dict, n_streams: int = 4, loss_and_grad_fn=None):
    """Forward + backward of a batch of views with the gradients summed into ``accumulate_into`` (dense tensors shaped
    like the parameters, e.g. ``GradAccumulator.grads()``).  This is the per-rank body of a data-parallel step.

    ``camera_fn(i)`` -> dict(view, proj, frustumplane) and ``loss_fn(i, img)`` -> scalar loss are called with view i's
    stream current (so H2D copies issued inside them are ordered correctly).  ``loss_fn`` may instead return
    ``(loss, d_img)`` -- the scalar and dloss/dimg computed outside autograd, e.g. ``ssim.l1_ssim_loss_and_grad(img.detach(),
    gt)`` -- in which case the image gradient is fed straight to the rasterizer's backward.  Consecutive views alternate over
    ``n_streams`` CUDA streams: view i+1's forward (bandwidth-bound projection / sort kernels and the one host
    read-back) overlaps view i's backward (issue-bound raster kernel).  Views only interact through the dense
    accumulate, which is ordered by an event.  Returns the list of (detached) per-view losses.

    ``loss_and_grad_fn(i, img) -> (loss, d_img)`` (with ``loss_fn=None``) skips autograd altogether: the pipeline's forward
    and backward are called directly (no autograd Function, no engine hop: ~0.1 ms less host time per view), the image
    handed to the function is the kernel's clamp(0,1) output and its gradient goes straight to the raster backward."""
    dev = xyz.device
    losses = []
    H, W = int(output_shape[0]), int(output_shape[1])
    th, tw = int(pp.tile_size[0]), int(pp.tile_size[1])
    direct = loss_and_grad_fn is not None or _DIRECT_VIEWS
    if direct:
        params = dict(xyz=xyz.detach(), scale=scale.detach(), rot=rot.detach(), sh_0=sh_0.detach(), sh_rest=sh_rest.detach(),
                      opacity=opacity.detach())
        stat = bool(StatisticsHelperInst.bStart)

    def one_direct(i, wait_ev):
        cam = camera_fn(i)
        img_p, state, stats = pipeline.render_view_forward(params, cluster_origin, cluster_extend, cam["frustumplane"], cam["view"],
                                                           cam["proj"], int(actived_sh_degree), (H, W), (th, tw), enable_statistic=stat,
                                                           clamp_zero=True)
        if loss_and_grad_fn is not None:
            loss, d_img = loss_and_grad_fn(i, img_p[..., :H, :W])
        else:                          # autograd only through the user's loss, never through the render kernels
            leaf = img_p[..., :H, :W].detach().requires_grad_(True)
            loss = loss_fn(i, leaf)
            if isinstance(loss, tuple):
                loss, d_img = loss
            else:
                (d_img,) = torch.autograd.grad(loss, leaf)
        if d_img.shape[-2:] != img_p.shape[-2:]:                        # image padded to whole tiles: pad the gradient with zeros
            d_img = torch.nn.functional.pad(d_img, (0, img_p.shape[-1] - W, 0, img_p.shape[-2] - H))
        if wait_ev is not None:
            torch.cuda.current_stream(dev).wait_event(wait_ev)
        _, pg_ = pipeline.render_view_backward(params, state, d_img, None, enable_statistic=stat, accumulate_into=accumulate_into,
                                               clamped_img=img_p)
        if stat:
            _feed_statistics(state, stats, pg_, (th, tw))
        losses.append(loss.detach())

    def one_autograd(i, wait_ev):
        cam = camera_fn(i)
        img = render_view(cluster_origin, cluster_extend, cam["frustumplane"], cam["view"], cam["proj"], xyz, scale, rot, sh_0, sh_rest,
                          opacity, actived_sh_degree, output_shape, pp, accumulate_into=accumulate_into)[0]
        loss = loss_fn(i, img)
        if wait_ev is not None:          # the previous view's accumulate (other stream) must have landed
            torch.cuda.current_stream(dev).wait_event(wait_ev)
        if isinstance(loss, tuple):
            loss, d_img = loss
            img.backward(d_img)
        else:
            loss.backward()
        losses.append(loss.detach())

    # GPU-driven path (default): one preallocated ViewWorkspace per stream slot, no host synchronisation inside the batch, the
    # per-view forward / backward replayed as CUDA graphs.  The first batch of a configuration goes through the synchronising
    # path and measures the capacities (the reference's cold first epoch); statistics runs keep the synchronising path.
    slots = None
    if direct and pipeline.SYNC_FREE and not stat:
        slots = _view_slots(params, (H, W), (th, tw), max(1, n_streams))
        if slots.big:
            slots = None
    probe = {"pairs": 0, "bits": 1} if (slots is not None and slots.ws is None) else None

    def one_ws(i, wait_ev, ws):
        cam = camera_fn(i)
        img_p = ws.forward(params, cluster_origin, cluster_extend, cam, int(actived_sh_degree), clamp_zero=True)
        if loss_and_grad_fn is not None:
            loss, d_img = loss_and_grad_fn(i, img_p[..., :H, :W])
        else:
            leaf = img_p[..., :H, :W].detach().requires_grad_(True)
            loss = loss_fn(i, leaf)
            if isinstance(loss, tuple):
                loss, d_img = loss
            else:
                (d_img,) = torch.autograd.grad(loss, leaf)
        if wait_ev is not None:
            torch.cuda.current_stream(dev).wait_event(wait_ev)
        ws.backward(params, d_img, int(actived_sh_degree), accumulate_into, use_clamp=True)
        losses.append(loss.detach())

    def one_probe(i, wait_ev):
        n0 = len(pipeline.LAST_VIEW_SIZES)
        one_direct(i, wait_ev)
        for pairs, bits in pipeline.LAST_VIEW_SIZES[n0:]:
            probe["pairs"] = max(probe["pairs"], pairs); probe["bits"] = max(probe["bits"], bits)
        del pipeline.LAST_VIEW_SIZES[:]

    if slots is not None and slots.ws is not None:
        slots.check()
        one = None
    elif probe is not None:
        one = one_probe
    else:
        one = one_direct if direct else one_autograd
    if n_streams <= 1:
        for i in range(n_views):
            if one is None:
                one_ws(i, None, slots.ws[0])
            else:
                one(i, None)
        if one is None:
            slots.ws[0].post_flags()
        elif probe is not None:
            slots.size(probe["pairs"], probe["bits"])
        return losses
    cur = torch.cuda.current_stream(dev)
    side = _streams(dev, n_streams)
    for s in side:
        s.wait_stream(cur)
    prev = None
    for i in range(n_views):
        s = side[i % n_streams]
        with torch.cuda.stream(s):
            if one is None:
                one_ws(i, prev, slots.ws[i % n_streams])
            else:
                one(i, prev)
            prev = torch.cuda.Event()
            prev.record(s)
    if one is None:
        for k, s in enumerate(side[: min(n_streams, n_views)]):
            with torch.cuda.stream(s):
                slots.ws[k].post_flags()
    elif probe is not None:
        slots.size(probe["pairs"], probe["bits"])
    for s in side:
        cur.wait_stream(s)
    return losses


class _ViewSlots:
    """The per-configuration state of render_views' GPU-driven path: capacities measured by the first (synchronising) batch and
    one ViewWorkspace per stream slot, created from them."""

    def __init__(self, params, hw, tile):
        self.params_like, self.hw, self.tile = params, hw, tile
        self.ws = None
        self.big = False
        self.cap, self.bits = 0, 24

    def size(self, max_pairs, max_bits):
        """Capacities from the measured maxima: 30 % head-room on the pairs, the depth range rounded up to whole 8-bit passes (a
        view that needs more is flagged and the step redone)."""
        self.cap = int(max_pairs * 1.3) + 65536
        self.bits = min(32, 8 * ((max_bits + 7) // 8))
        H, W = self.hw
        ntile = ((W + self.tile[1] - 1) // self.tile[1]) * ((H + self.tile[0] - 1) // self.tile[0])
        # very large pair lists with 8-bit tile digits (4K frames): the library's onesweep, which needs the count on the host,
        # beats the own passes (csrc/sort.cu: sort_impl_for) and such views are device-bound anyway -> stay on the synchronising path
        self.ws = None if (ntile.bit_length() > 14 and max_pairs > (8 << 20)) else []
        self.big = self.ws is None

    def ensure(self, n_slots):
        """One workspace per stream slot (created on demand once the capacities are known)."""
        while self.ws is not None and len(self.ws) < n_slots:
            self.ws.append(pipeline.ViewWorkspace(self.params_like, self.hw, self.tile, self.cap, self.bits))

    def check(self):
        """Overflow flags of the previous batch (if they have landed): on overflow drop the workspaces -- the next batch measures
        again -- and tell the caller to redo the step."""
        try:
            for w in self.ws:
                w.check(wait=False)
        except pipeline.CapacityExceeded:
            self.ws = None
            raise


_slot_cache: dict = {}


def _view_slots(params, hw, tile, n_slots):
    xyz = params["xyz"]
    key = (xyz.device, tuple(xyz.shape[-2:]), hw, tile, params["sh_rest"].shape[0])
    ent = _slot_cache.get(key)
    if ent is None:
        ent = _slot_cache[key] = _ViewSlots(params, hw, tile)
    ent.ensure(n_slots)
    return ent


def check_views(wait: bool = True):
    """Explicit form of the lazy overflow check of render_views' GPU-driven path: (after a synchronisation) raises
    pipeline.CapacityExceeded if a view of the last batch overflowed its workspace -- redo that step."""
    for ent in _slot_cache.values():
        if ent.ws is not None:
            try:
                for w in ent.ws:
                    w.check(wait=wait)
            except pipeline.CapacityExceeded:
                ent.ws = None
                raise


def reset_view_workspaces():
    """Forget every cached workspace / CUDA graph (call after the scene's size changed, e.g. densification)."""
    _slot_cache.clear()
