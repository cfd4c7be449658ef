"""Host-side (Python) cost of one view through render.render_views, on a scene small enough that the GPU is never
the limit.  usage: python profiles/host_profile.py [gaussians] [steps]"""
import cProfile
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
from litegs_b200 import dist as lgs_dist, render, scene  # noqa: E402
from litegs_b200.arguments import PipelineParams  # noqa: E402

KEYS = ("xyz", "scale", "rot", "sh_0", "sh_rest", "opacity")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    H, W, vpr = 240, 320, 8
    dev = torch.device("cuda:0")
    sc = scene.make_scene(n, sh_degree=3, seed=0)
    P = {k: torch.from_numpy(sc[k]).to(dev).requires_grad_(True) for k in KEYS}
    A = [torch.from_numpy(sc[k]).to(dev) for k in ("cluster_origin", "cluster_extend")]
    cams = [{k: torch.from_numpy(v).to(dev) for k, v in scene.make_camera(j, 64, W, H).items()} for j in range(vpr)]
    w = torch.randn((1, 3, H, W), device=dev)
    acc = lgs_dist.GradAccumulator(P)
    pp = PipelineParams(tile_size=(8, 16), sparse_grad=True)

    def step():
        acc.zero_()
        render.render_views(vpr, lambda j: cams[j], lambda j, img: (img * w).sum(), A[0], A[1], P["xyz"], P["scale"], P["rot"], P["sh_0"],
                            P["sh_rest"], P["opacity"], 3, (H, W), pp, acc.grads(), n_streams=3)

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{dt / (steps * vpr) * 1e6:.1f} us of wall clock per view ({n} Gaussians, {W}x{H})")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)
    st.sort_stats("cumulative").print_stats(22)


if __name__ == "__main__":
    main()
