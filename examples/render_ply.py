"""Render a 3DGS point cloud (.ply, the layout shared by every 3DGS code base) with the fused forward pipeline.

    python examples/render_ply.py --ply point_cloud.ply --out /tmp/renders --views 8 [--colmap /path/to/colmap]
    python examples/render_ply.py --make /tmp/demo.ply --out /tmp/renders        # writes a synthetic cloud first

Cameras: the poses of a COLMAP model when --colmap is given, else the Fibonacci lattice of scene.make_camera.  Forward only
(the reference's example_metrics.py path): project -> bin -> sort -> composite, no gradients kept.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from litegs_b200 import colmap, pipeline, ply, scene  # noqa: E402
from litegs_b200.dist import PARAM_ORDER  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ply", default=None)
    ap.add_argument("--make", default=None, help="write a synthetic cloud to this .ply first")
    ap.add_argument("--colmap", default=None)
    ap.add_argument("--out", default="renders")
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--height", type=int, default=540)
    ap.add_argument("--sh-degree", type=int, default=3)
    a = ap.parse_args()
    path = a.ply
    if a.make:
        sc = scene.make_scene(200_000, sh_degree=3, seed=0)
        ply.params_to_ply(a.make, sc)
        path = a.make
    if path is None:
        ap.error("give --ply or --make")
    dev = torch.device("cuda:0")
    g = ply.params_from_ply(path, a.sh_degree)
    P = {k: torch.from_numpy(g[k]).to(dev) for k in PARAM_ORDER}
    A = [torch.from_numpy(g[k]).to(dev) for k in ("cluster_origin", "cluster_extend")]
    cams = []
    if a.colmap:
        cs, ims, _ = colmap.read_model(a.colmap)
        for im in sorted(ims.values(), key=lambda v: v.name)[: a.views]:
            c = cs[im.camera_id]
            cams.append((colmap.camera_from_colmap(im.qvec, im.tvec, c.params, c.width, c.height), (c.height, c.width), im.name))
    else:
        cams = [(scene.make_camera(i, a.views, a.width, a.height), (a.height, a.width), f"view_{i:04d}.png") for i in range(a.views)]
    import PIL.Image
    os.makedirs(a.out, exist_ok=True)
    imgs = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        for cam, hw, _ in cams:
            c = {k: torch.from_numpy(v).to(dev) for k, v in cam.items()}
            img, _, _ = pipeline.render_view_forward(P, A[0], A[1], c["frustumplane"], c["view"], c["proj"], a.sh_degree, hw, (8, 16),
                                                     clamp_zero=True)
            imgs.append(img[0, :, : hw[0], : hw[1]])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    for (cam, hw, name), img in zip(cams, imgs):
        PIL.Image.fromarray((img.permute(1, 2, 0) * 255.0 + 0.5).clamp(0, 255).to(torch.uint8).cpu().numpy()).save(
            os.path.join(a.out, os.path.splitext(name)[0] + ".png"))
    print(f"{g['n_points']} Gaussians, {len(cams)} views rendered in {dt * 1e3:.1f} ms ({len(cams) / dt:.0f} views/s forward only) -> {a.out}")


if __name__ == "__main__":
    main()
