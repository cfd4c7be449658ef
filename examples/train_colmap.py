"""From a COLMAP directory to a trained set of Gaussians on the litegs_b200 stack (BASELINE.json config 5 in spirit).

    python examples/train_colmap.py --make /tmp/synth_colmap          # writes a synthetic dataset first (renders of a hidden scene)
    python examples/train_colmap.py --data /path/to/colmap --iters 2000

Reads ``sparse/0/{cameras,images,points3D}.bin`` and ``images/*`` (litegs_b200.colmap; same files and conventions as the
reference's ``litegs/io_manager/colmap.py`` + ``litegs/data.py``), initialises Gaussians from the SfM points the way
``litegs/scene/point.py:7-19`` does, and trains appearance + geometry with the fused L1+SSIM loss and the fused Adam step.
No densification (that policy is out of scope): the point count stays what COLMAP delivered.
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from litegs_b200 import colmap, dist as lgs_dist, optimizer, render, scene, ssim  # noqa: E402
from litegs_b200.arguments import PipelineParams  # noqa: E402
from litegs_b200.dist import PARAM_ORDER  # noqa: E402


def make_dataset(root, n_gaussians=60_000, n_views=24, hw=(270, 480), n_points=20_000, seed=0, dev=None, log_scale_range=(0.01, 0.04)):
    """A hidden scene rendered from the lattice cameras -> COLMAP model + PNGs.  The 'SfM points' are a subsample of the
    hidden Gaussians' centres with their band-0 colours (what a real reconstruction would roughly deliver)."""
    dev = dev or torch.device("cuda:0")
    H, W = hw
    pp = PipelineParams(tile_size=(8, 16), sparse_grad=True)
    truth = scene.make_scene(n_gaussians, sh_degree=3, seed=seed, log_scale_range=log_scale_range)
    T = {k: torch.from_numpy(truth[k]).to(dev) for k in PARAM_ORDER}
    A = [torch.from_numpy(truth[k]).to(dev) for k in ("cluster_origin", "cluster_extend")]

    def render_fn(i, cam):
        c = {k: torch.from_numpy(v).to(dev) for k, v in cam.items()}
        with torch.no_grad():
            img = render.render_view(A[0], A[1], c["frustumplane"], c["view"], c["proj"], T["xyz"], T["scale"], T["rot"], T["sh_0"],
                                     T["sh_rest"], T["opacity"], 3, (H, W), pp)[0]
        return (img[0].permute(1, 2, 0) * 255.0 + 0.5).clamp(0, 255).to(torch.uint8).cpu().numpy()

    rng = np.random.default_rng(seed + 7)
    xyz = truth["xyz"].reshape(3, -1).T
    sel = rng.choice(xyz.shape[0], size=min(n_points, xyz.shape[0]), replace=False)
    rgb = np.clip((truth["sh_0"].reshape(3, -1).T[sel] * colmap.SH_C0 + 0.5) * 255.0, 0, 255).astype(np.uint8)
    colmap.write_synthetic_dataset(root, xyz[sel].astype(np.float64), rgb, n_views, W, H, render_fn=render_fn)
    return root


def load_dataset(root, image_dir="images", dev=None):
    import PIL.Image
    dev = dev or torch.device("cuda:0")
    cams, images, pts = colmap.read_model(root)
    frames = []
    for im in sorted(images.values(), key=lambda v: v.name):
        c = cams[im.camera_id]
        if c.model != "PINHOLE":
            continue                                     # as the reference (colmap.py:222-224)
        cam = colmap.camera_from_colmap(im.qvec, im.tvec, c.params, c.width, c.height)
        gt = np.array(PIL.Image.open(os.path.join(root, image_dir, im.name)).convert("RGB"), np.uint8)
        frames.append(({k: torch.from_numpy(v).to(dev) for k, v in cam.items()},
                       torch.from_numpy(gt).to(dev).permute(2, 0, 1)[None].float().div_(255.0).contiguous(), (c.height, c.width)))
    P = list(pts.values())
    return frames, np.stack([p.xyz for p in P]), np.stack([p.rgb for p in P])


def train(root, iters=300, views_per_step=8, log=print):
    from litegs_b200 import fused
    keep = fused.CONFIG["true_sigmoid_grad"]
    fused.CONFIG["true_sigmoid_grad"] = True               # our own loops train with the true sigmoid derivative
    try:
        return _train(root, iters, views_per_step, log)
    finally:
        fused.CONFIG["true_sigmoid_grad"] = keep


def _train(root, iters, views_per_step, log):
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    frames, xyz, rgb = load_dataset(root, dev=dev)
    H, W = frames[0][2]
    g = colmap.gaussians_from_points(xyz, rgb, sh_degree=3)
    P = {k: torch.from_numpy(g[k]).to(dev) for k in PARAM_ORDER}
    pp = PipelineParams(tile_size=(8, 16), sparse_grad=True)
    acc = lgs_dist.GradAccumulator(P)
    extent = float(np.linalg.norm(xyz.max(0) - xyz.min(0)) * 0.5)
    opt, sched = optimizer.get_optimizer(P, spatial_lr_scale=extent)
    hist = []
    t0 = time.perf_counter()
    for it in range(iters):
        idx = [(it * views_per_step + j) % len(frames) for j in range(views_per_step)]
        # positions and shapes move, so the chunk AABBs used for culling are refreshed from the parameters now and then
        if it % 50 == 0:
            A = list(scene.cluster_aabb_torch(P["xyz"], P["scale"], P["rot"]))
        losses = render.render_views(views_per_step, lambda i: frames[idx[i]][0], None, A[0], A[1], P["xyz"], P["scale"], P["rot"],
                                     P["sh_0"], P["sh_rest"], P["opacity"], 3, (H, W), pp, acc.grads(),
                                     loss_and_grad_fn=lambda i, img: ssim.l1_ssim_loss_and_grad(img.contiguous(), frames[idx[i]][1], 0.2,
                                                                                              upstream=1.0 / views_per_step))
        opt.step(acc)
        sched.step()
        hist.append(float(torch.stack(losses).mean()))
        if it % 50 == 0 or it == iters - 1:
            log(f"iter {it:5d}  loss {hist[-1]:.5f}")
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    with torch.no_grad():
        mse = []
        for cam, gt, _ in frames[:8]:
            img = render.render_view(A[0], A[1], cam["frustumplane"], cam["view"], cam["proj"], P["xyz"], P["scale"], P["rot"], P["sh_0"],
                                     P["sh_rest"], P["opacity"], 3, (H, W), pp)[0]
            mse.append(float(((img - gt) ** 2).mean()))
    psnr = -10.0 * np.log10(np.mean(mse))
    log(f"{iters} iterations x {views_per_step} views in {dt:.1f} s ({iters * views_per_step / dt:.0f} views/s incl. loss + optimizer); "
        f"{xyz.shape[0]} Gaussians, PSNR over 8 training views {psnr:.2f} dB")
    return hist, psnr


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--make", default=None, help="write a synthetic COLMAP dataset to this directory first and train on it")
    ap.add_argument("--data", default=None)
    ap.add_argument("--iters", type=int, default=300)
    a = ap.parse_args()
    root = a.data
    if a.make:
        root = make_dataset(a.make)
    if root is None:
        ap.error("give --data or --make")
    h, _ = train(root, a.iters)
    assert h[-1] < h[0]
