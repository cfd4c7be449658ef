"""Shared helpers for the parity tests: small seeded scenes pushed through the CPU oracle."""
import numpy as np

import oracle
from litegs_b200 import scene

PARAM_KEYS = ("xyz", "scale", "rot", "sh_0", "sh_rest", "opacity")


def small_scene(n=2000, hw=(96, 128), tile=(16, 16), sh_degree=3, seed=0, log_scale_range=(0.02, 0.08), view=0, n_views=8,
                chunk=128):
    p = scene.make_scene(n, sh_degree=sh_degree, chunk=chunk, seed=seed, log_scale_range=log_scale_range)
    cam = scene.make_camera(view, n_views, hw[1], hw[0])
    params = {k: p[k] for k in PARAM_KEYS}
    return params, (p["cluster_origin"], p["cluster_extend"]), cam


def oracle_projected(params, aabb, cam, hw, sh_degree):
    """Oracle outputs up to the inputs of binning/raster: activated+projected per-Gaussian tensors (numpy)."""
    vis, nvis, ids = oracle.frustum_culling_aabb(aabb[0], aabb[1], cam["frustumplane"])
    act = oracle.cull_compact_activate(sh_degree, ids, nvis, cam["view"], params["xyz"], params["scale"], params["rot"],
                                       params["sh_0"], params["sh_rest"], params["opacity"])
    xyz, scale, rot, color, opacity = [a.reshape(*a.shape[:-2], -1) for a in act]
    inter = oracle.project(xyz, scale, rot, cam["view"], cam["proj"], hw)
    return dict(ids=ids, nvis=nvis, vis=vis, act=act, xyz=xyz, scale=scale, rot=rot, color=color, opacity=opacity, **inter)


def rel_err(a, b):
    """max |a-b| / max(1, |b|) -- the Tier-1 metric of SURVEY 8c."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def scaled_err(a, b):
    """max |a-b| / max|b| -- for gradients whose magnitude is far from 1."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    m = np.max(np.abs(b)) if b.size else 0.0
    return float(np.max(np.abs(a - b)) / max(m, 1e-30)) if a.size else 0.0
