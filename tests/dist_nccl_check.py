"""N-rank NCCL numeric check (launched by tests/test_gpu_dist.py under torch.distributed.run, one rank per GPU):
the all-reduced dense gradient buffer of a view batch sharded over the ranks equals the single-rank sum of the same views.

Every rank renders its round-robin shard of the views (render_views -> GradAccumulator), the buffers are summed with ONE NCCL
all-reduce (GradAccumulator.all_reduce, the only collective of the data path), and rank 0 compares the result with the
buffer it gets by rendering ALL views itself.  Differences are limited to the order of fp32 additions."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    from litegs_b200 import dist as lgs_dist, render, scene
    from litegs_b200.arguments import PipelineParams
    hw, tile, deg, n_views = (270, 480), (8, 16), 3, 16
    p = scene.make_scene(200_000, sh_degree=3, cube=2.0, seed=3, log_scale_range=(0.004, 0.03))     # cube 2: every view culls some chunks
    P = {k: torch.from_numpy(p[k]).to(dev) for k in lgs_dist.PARAM_ORDER}
    A = [torch.from_numpy(p[k]).to(dev) for k in ("cluster_origin", "cluster_extend")]
    pp = PipelineParams(tile_size=tile)
    cams = [{k: torch.from_numpy(v).to(dev) for k, v in scene.make_camera(i, n_views, hw[1], hw[0]).items()} for i in range(n_views)]
    w = torch.from_numpy(np.random.default_rng(0).normal(size=(1, 3, *hw)).astype(np.float32)).to(dev)

    def render_into(acc, views):
        return render.render_views(len(views), lambda j: cams[views[j]], lambda j, img: (img * w).sum() * (1.0 + 0.01 * views[j]), A[0], A[1],
                                   P["xyz"], P["scale"], P["rot"], P["sh_0"], P["sh_rest"], P["opacity"], deg, hw, pp, acc.grads())
    mine = lgs_dist.shard_views(n_views, rank, world)
    acc = lgs_dist.GradAccumulator(P)
    losses = render_into(acc, mine)
    acc.all_reduce(async_op=True)          # side stream, as the training loop / bench use it
    acc.wait()
    torch.cuda.synchronize()
    loss_sum = torch.stack(losses).sum().reshape(1).double()
    dist.all_reduce(loss_sum)
    out = {"world": world, "views": n_views,
           "allreduce_impl": "own NVLS multimem kernel (csrc/nvls.cu)" if acc._nvls is not None else "ncclAllReduce"}
    if rank == 0:
        ref = lgs_dist.GradAccumulator(P, symmetric=False)      # rank-local: a symmetric allocation would be a collective
        ref_losses = render_into(ref, list(range(n_views)))
        torch.cuda.synchronize()
        a, b = acc.buf.double(), ref.buf.double()
        out["buffer_max_abs_diff_over_max"] = float((a - b).abs().max() / b.abs().max())
        rows = acc.rows
        out["per_parameter"] = {k: float((a[rows[k]] - b[rows[k]]).abs().max() / (b[rows[k]].abs().max() + 1e-300)) for k in lgs_dist.PARAM_ORDER}
        out["marks_equal"] = bool(torch.equal(acc.touched > 0, ref.touched > 0))
        out["marks_visible_chunks"] = int((ref.touched > 0).sum())
        out["chunks"] = int(ref.touched.numel())
        out["loss_sum_rel_diff"] = float(abs(loss_sum.item() - torch.stack(ref_losses).double().sum().item()) / abs(loss_sum.item()))
        out["nccl_version"] = ".".join(str(x) for x in torch.cuda.nccl.version())
        out["ok"] = bool(out["buffer_max_abs_diff_over_max"] < 1e-5 and out["marks_equal"] and out["loss_sum_rel_diff"] < 1e-6
                         and 0 < out["marks_visible_chunks"])
        print("NCCL_CHECK " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not out["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
