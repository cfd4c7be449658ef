"""The GPU-driven path of render_views (no host synchronisation inside a batch, preallocated ViewWorkspace per stream slot,
forward / backward replayed as CUDA graphs) against the synchronising path it replaces: same losses, same accumulated
gradients; capacity overflow is flagged on the device and reported."""
import numpy as np
import pytest
import torch

from litegs_b200 import pipeline, render, scene
from litegs_b200.arguments import PipelineParams
from litegs_b200.dist import GradAccumulator
from tests.util import PARAM_KEYS, scaled_err

pytestmark = pytest.mark.gpu


def _setup(cuda, n=8000, hw=(72, 96), seed=6, cube=1.5):
    p = scene.make_scene(n, sh_degree=3, cube=cube, seed=seed, log_scale_range=(0.02, 0.08))
    P = {k: torch.from_numpy(p[k]).to(cuda) for k in PARAM_KEYS}
    A = [torch.from_numpy(p[k]).to(cuda) for k in ("cluster_origin", "cluster_extend")]
    cams = [{k: torch.from_numpy(x).to(cuda) for k, x in scene.make_camera(v, 12, hw[1], hw[0]).items()} for v in range(12)]
    w = torch.from_numpy(np.random.default_rng(0).normal(size=(1, 3, *hw)).astype(np.float32)).to(cuda)
    return P, A, cams, w


def _batch(P, A, cams, w, hw, pp, acc, views, n_streams):
    acc.zero_()
    losses = render.render_views(len(views), lambda i: cams[views[i]], lambda i, img: (img * w).sum() * (1.0 + 0.1 * views[i]), A[0], A[1],
                                 P["xyz"], P["scale"], P["rot"], P["sh_0"], P["sh_rest"], P["opacity"], 3, hw, pp, acc.grads(), n_streams=n_streams)
    torch.cuda.synchronize()
    return [float(x) for x in losses], {k: v.clone() for k, v in acc.grads().items()}


@pytest.mark.parametrize("n_streams", [1, 3])
@pytest.mark.parametrize("hw", [(72, 96), (70, 100)])            # whole tiles / padded image
def test_gpu_driven_batches_match_the_synchronising_path(cuda, n_streams, hw):
    P, A, cams, w = _setup(cuda, hw=hw)
    pp = PipelineParams(tile_size=(8, 16))
    acc = GradAccumulator(P)
    render.reset_view_workspaces()
    keep = pipeline.SYNC_FREE
    try:
        pipeline.SYNC_FREE = False
        ref_a = _batch(P, A, cams, w, hw, pp, acc, [0, 1, 2, 3, 4, 5], n_streams)
        ref_b = _batch(P, A, cams, w, hw, pp, acc, [6, 7, 8, 9, 10, 11], n_streams)
        pipeline.SYNC_FREE = True
        # batch 1 measures the capacities on the synchronising path, 2 runs eagerly on the workspaces, 3 captures the graphs,
        # 4.. replay them -- with other cameras in between (a replay must pick up the new camera, not the captured one)
        got = [_batch(P, A, cams, w, hw, pp, acc, v, n_streams) for v in ([0, 1, 2, 3, 4, 5], [0, 1, 2, 3, 4, 5], [6, 7, 8, 9, 10, 11],
                                                                           [0, 1, 2, 3, 4, 5], [6, 7, 8, 9, 10, 11], [0, 1, 2, 3, 4, 5])]
        render.check_views(wait=True)
    finally:
        pipeline.SYNC_FREE = keep
        render.reset_view_workspaces()
    for i, g in enumerate(got):
        ref = ref_b if i in (2, 4) else ref_a
        assert np.allclose(g[0], ref[0], rtol=1e-6), i
        for k in PARAM_KEYS:
            assert scaled_err(g[1][k].cpu().numpy(), ref[1][k].cpu().numpy()) < 1e-5, (i, k)
        assert torch.equal(g[1]["_touched"] > 0, ref[1]["_touched"] > 0)


def test_workspace_overflow_is_flagged(cuda):
    hw, tile = (72, 96), (8, 16)
    P, A, cams, w = _setup(cuda, hw=hw)
    acc = GradAccumulator(P)
    pairs, bits = pipeline.probe_view_sizes(P, A[0], A[1], cams[:2], 3, hw, tile)
    assert pairs > 2000 and 1 <= bits <= 32
    for cap, planned, what in ((pairs // 2, 32, "pairs"), (pairs * 2, max(1, bits - 2), "depth bits")):
        ws = pipeline.ViewWorkspace(P, hw, tile, pair_capacity=max(1024, cap), planned_depth_bits=planned, use_graphs=False)
        img = ws.forward(P, A[0], A[1], cams[0], 3)
        ws.backward(P, torch.ones_like(img), 3, acc.grads())
        ws.post_flags()
        torch.cuda.synchronize()
        with pytest.raises(pipeline.CapacityExceeded) as e:
            ws.check(wait=True)
        assert (e.value.pairs > e.value.pair_capacity) if what == "pairs" else (e.value.depth_bits > e.value.planned_depth_bits)
    # and a sufficient workspace reports the measured sizes
    ws = pipeline.ViewWorkspace(P, hw, tile, pair_capacity=pairs + 1000, planned_depth_bits=32, use_graphs=False)
    ws.forward(P, A[0], A[1], cams[0], 3); ws.post_flags(); torch.cuda.synchronize()
    r = ws.check(wait=True)
    assert r["views"] == 1 and 0 < r["max_pairs"] <= pairs


def test_workspace_forward_equals_synchronising_forward(cuda):
    """Same image, transmittance, contributor counts and tile lists as pipeline.render_view_forward (bit for bit)."""
    hw, tile = (70, 100), (8, 16)
    P, A, cams, w = _setup(cuda, hw=hw)
    pairs, bits = pipeline.probe_view_sizes(P, A[0], A[1], cams, 3, hw, tile)
    ws = pipeline.ViewWorkspace(P, hw, tile, pair_capacity=int(pairs * 1.3), planned_depth_bits=32, use_graphs=False)
    for cam in cams[:4]:
        img_ref, st, _ = pipeline.render_view_forward(P, A[0], A[1], cam["frustumplane"], cam["view"], cam["proj"], 3, hw, tile, clamp_zero=True)
        img = ws.forward(P, A[0], A[1], cam, 3)
        torch.cuda.synchronize()
        D = st.n_pairs
        assert int(ws.vparams[5]) == D
        assert torch.equal(ws.ranges, st.ranges) and torch.equal(ws.sorted_pid[:, :D], st.sorted_pid)
        assert torch.equal(img, img_ref) and torch.equal(ws.T, st.T) and torch.equal(ws.last, st.last)
