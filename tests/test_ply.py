"""CPU tests of litegs_b200/ply.py: the standard 3DGS point-cloud layout (what litegs/io_manager/ply.py:7-45 writes via
plyfile): header, property order, channel-major SH, round trips, fewer SH bands, ASCII files, clustering."""
import numpy as np
import pytest

from litegs_b200 import ply, scene


def _raw(n=300, deg=3, seed=0):
    rng = np.random.default_rng(seed)
    K = (deg + 1) ** 2
    return (rng.normal(size=(3, n)).astype(np.float32), rng.normal(size=(3, n)).astype(np.float32), rng.normal(size=(4, n)).astype(np.float32),
            rng.normal(size=(1, 3, n)).astype(np.float32), rng.normal(size=(K - 1, 3, n)).astype(np.float32), rng.normal(size=(1, n)).astype(np.float32))


def test_header_and_layout_are_the_standard_ones(tmp_path):
    v = _raw(n=5)
    p = str(tmp_path / "a" / "pc.ply")
    ply.save_ply(p, *v)
    raw = open(p, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", "element vertex 5"]
    props = [ln.split()[2] for ln in lines if ln.startswith("property float ")]
    assert props == (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)] +
                     ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"])
    tab = np.frombuffer(body, "<f4").reshape(5, len(props))
    xyz, scale, rot, sh0, shr, op = v
    assert np.array_equal(tab[:, :3], xyz.T) and not tab[:, 3:6].any() and np.array_equal(tab[:, 6:9], sh0[0].T)
    # f_rest is channel-major: f_rest_[c * 15 + k] = sh_rest[k, c]   (litegs/io_manager/ply.py:12,39-41)
    assert np.array_equal(tab[:, 9 + 1 * 15 + 4], shr[4, 1]) and np.array_equal(tab[:, 9 + 2 * 15 + 14], shr[14, 2])
    assert np.array_equal(tab[:, 54], op[0]) and np.array_equal(tab[:, 55:58], scale.T) and np.array_equal(tab[:, 58:62], rot.T)


def test_round_trip_and_lower_sh_degree(tmp_path):
    v = _raw(n=257)
    p = str(tmp_path / "pc.ply")
    ply.save_ply(p, *v)
    got = ply.load_ply(p, 3)
    assert all(np.array_equal(a, b) and a.dtype == np.float32 for a, b in zip(got, v))
    v1 = _raw(n=40, deg=1)
    p1 = str(tmp_path / "pc1.ply")
    ply.save_ply(p1, *v1)
    g1 = ply.load_ply(p1, 3)                       # degree-1 file into a degree-3 model: higher bands are zero
    assert np.array_equal(g1[4][:3], v1[4]) and not g1[4][3:].any() and g1[4].shape == (15, 3, 40)
    with pytest.raises(ValueError):
        ply.load_ply(p, 1)                          # 45 rest coefficients do not fit degree 1


def test_ascii_ply_and_clustered_round_trip(tmp_path):
    v = _raw(n=10, deg=0)
    names = ply._names(3, 0)
    cols = np.concatenate([v[0].T, np.zeros((10, 3)), v[3][0].T, v[5].T, v[1].T, v[2].T], 1)
    p = str(tmp_path / "ascii.ply")
    with open(p, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 10\n" + "".join(f"property float {a}\n" for a in names) + "end_header\n")
        for r in cols:
            f.write(" ".join(repr(float(x)) for x in r) + "\n")
    g = ply.load_ply(p, 0)
    assert np.allclose(g[0], v[0]) and np.allclose(g[5], v[5]) and g[4].shape == (0, 3, 10)
    sc = scene.make_scene(1000, sh_degree=3, seed=4)
    q = str(tmp_path / "scene.ply")
    ply.params_to_ply(q, sc, n_points=1000)
    back = ply.params_from_ply(q, 3, chunk=128, morton=False)
    assert back["n_points"] == 1000
    for k in ply.PARAM_KEYS:
        a = sc[k].reshape(*sc[k].shape[:-2], -1)[..., :1000]
        b = back[k].reshape(*back[k].shape[:-2], -1)[..., :1000]
        assert np.array_equal(a, b), k
    assert back["cluster_origin"].shape == (3, back["xyz"].shape[-2])


def test_chunk_aabbs_on_torch_match_the_numpy_form():
    """scene.cluster_aabb_torch (device-agnostic, used by the training example to refresh the culling boxes) against
    scene.cluster_aabb (reference semantics litegs/scene/cluster.py:29-46) on a random clustered scene."""
    import torch
    sc = scene.make_scene(3000, sh_degree=0, seed=8, log_scale_range=(0.01, 0.2))
    o, e = scene.cluster_aabb(sc["xyz"], sc["scale"], sc["rot"])
    to, te = scene.cluster_aabb_torch(torch.from_numpy(sc["xyz"]), torch.from_numpy(sc["scale"]), torch.from_numpy(sc["rot"]))
    assert to.shape == o.shape and te.shape == e.shape and to.dtype == torch.float32
    assert np.abs(to.numpy() - o).max() < 1e-5 and np.abs(te.numpy() - e).max() < 1e-5
    assert np.array_equal(sc["cluster_origin"], o) and np.array_equal(sc["cluster_extend"], e)
