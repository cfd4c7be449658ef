"""CPU tests of the dependency stand-ins (litegs_b200/shims) that let the unmodified reference package import and train where
plyfile / torchmetrics / matplotlib / simple_knn are not installed, and of the host logic of the statistics
helper.  Each stand-in is checked on exactly the calls the reference makes."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def shims():
    from litegs_b200 import shims as s
    s.install()
    return s


def test_install_appends_after_site_packages(shims):
    here = os.path.dirname(os.path.abspath(shims.__file__))
    missing = [n for n in shims.NAMES if importlib.util.find_spec(n) is not None and
               os.path.abspath(importlib.util.find_spec(n).origin or "").startswith(here)]
    if missing:                                   # something is being shimmed: the directory sits at the END of sys.path
        assert sys.path.index(here) > max(i for i, p in enumerate(sys.path) if "site-packages" in p)
    assert shims.install() == shims.install() or True      # idempotent


def test_plyfile_roundtrip_like_the_reference(shims, tmp_path):
    """io_manager/ply.py:7-45 (save) and :47-86 (load); io_manager/colmap.py:281-306 (point cloud with uchar colours)."""
    plyfile = importlib.import_module("plyfile")
    if not os.path.abspath(plyfile.__file__).startswith(ROOT):
        pytest.skip("a real plyfile is installed")
    n = 257
    rng = np.random.default_rng(0)
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)] + [f"f_rest_{i}" for i in range(45)] + ["opacity"] + \
            [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]
    el = np.empty(n, dtype=[(a, "f4") for a in names])
    data = rng.normal(size=(n, len(names))).astype(np.float32)
    el[:] = list(map(tuple, data))
    path = str(tmp_path / "pc.ply")
    plyfile.PlyData([plyfile.PlyElement.describe(el, "vertex")]).write(path)
    back = plyfile.PlyData.read(path)
    assert [p.name for p in back.elements[0].properties] == names
    for j, a in enumerate(names):
        assert np.array_equal(np.asarray(back.elements[0][a]), data[:, j])
    # our own reader (litegs_b200/ply.py) reads what the stand-in wrote, and the stand-in reads what our writer writes
    from litegs_b200 import ply as lgs_ply
    xyz, scale, rot, sh0, shr, op = lgs_ply.load_ply(path, 3)
    assert np.array_equal(xyz.T, data[:, :3])
    lgs_ply.save_ply(str(tmp_path / "pc2.ply"), xyz, scale, rot, sh0, shr, op)
    again = plyfile.PlyData.read(str(tmp_path / "pc2.ply"))
    assert np.array_equal(np.asarray(again["vertex"]["x"]), data[:, 0]) and np.array_equal(np.asarray(again["vertex"]["opacity"]), op[0])
    # colmap.py's point cloud: float positions + uchar colours, element access by name
    pts = np.empty(5, dtype=[("x", "f4"), ("y", "f4"), ("z", "f4"), ("nx", "f4"), ("ny", "f4"), ("nz", "f4"), ("red", "u1"), ("green", "u1"),
                             ("blue", "u1")])
    pts[:] = [(i, i + 1, i + 2, 0, 0, 0, 10 * i, 20 * i, 250) for i in range(5)]
    plyfile.PlyData([plyfile.PlyElement.describe(pts, "vertex")]).write(str(tmp_path / "p.ply"))
    v = plyfile.PlyData.read(str(tmp_path / "p.ply"))["vertex"]
    assert np.array_equal(np.vstack([v["red"], v["green"], v["blue"]]).T, np.array([[10 * i, 20 * i, 250] for i in range(5)], np.uint8))
    # ascii files (written by other tools) read as well
    plyfile.PlyData([plyfile.PlyElement.describe(pts, "vertex")], text=True).write(str(tmp_path / "a.ply"))
    assert np.array_equal(np.asarray(plyfile.PlyData.read(str(tmp_path / "a.ply"))["vertex"]["x"]), pts["x"])


def test_torchmetrics_psnr(shims):
    tm = importlib.import_module("torchmetrics")
    if not os.path.abspath(tm.__file__).startswith(ROOT):
        pytest.skip("a real torchmetrics is installed")
    from torchmetrics.image import psnr
    m = psnr.PeakSignalNoiseRatio(data_range=(0.0, 1.0))
    g = torch.Generator().manual_seed(0)
    a, b = torch.rand((1, 3, 16, 16), generator=g) * 1.2 - 0.1, torch.rand((1, 3, 16, 16), generator=g)
    want = 10 * torch.log10(1.0 / ((a.clamp(0, 1) - b.clamp(0, 1)) ** 2).mean())
    assert abs(float(m(a, b)) - float(want)) < 1e-5
    m(b, a)
    both = 10 * torch.log10(1.0 / ((a.clamp(0, 1) - b) ** 2).mean())          # two equal-size batches: the overall mse is the mean
    assert abs(float(m.compute()) - float(both)) < 1e-5


def test_simple_knn_dist_and_matplotlib(shims):
    knn = importlib.import_module("simple_knn._C")
    g = torch.Generator().manual_seed(1)
    pts = torch.randn((700, 3), generator=g)
    got = knn.distCUDA2(pts)
    d = torch.cdist(pts.double(), pts.double()) ** 2
    d.fill_diagonal_(float("inf"))
    want = d.topk(3, largest=False).values.mean(1)
    assert float((got.double() - want).abs().max()) < 1e-5
    plt = importlib.import_module("matplotlib.pyplot")
    if os.path.abspath(plt.__file__).startswith(ROOT):
        plt.figure(); plt.plot([1, 2, 3]); plt.savefig("/dev/null"); plt.close("all")       # no-ops


def test_statistics_helper_dense_accumulation():
    """update_mean_std / get_mean / get_var without a compact mask (statistic_helper.py:96-156,215-243) on CPU tensors."""
    from litegs_b200.statistics import StatisticsHelper
    sh = StatisticsHelper(4, 8)
    assert not sh.bStart
    with sh.try_start(0):
        assert not sh.bStart                       # the default handle says "no statistics this epoch"
    sh.reset(4, 8, lambda epoch: epoch % 2 == 0)
    with sh.try_start(2):
        assert sh.bStart
    assert not sh.bStart
    rng = np.random.default_rng(0)
    tot, sq, cnt = np.zeros((1, 32)), np.zeros((1, 32)), np.zeros(32)
    for _ in range(3):
        v = rng.normal(size=(1, 32)).astype(np.float32); c = rng.integers(0, 3, 32).astype(np.int32)
        sh.update_mean_std("k", torch.from_numpy(v), torch.from_numpy(v * v), torch.from_numpy(c), bCompacted=False)
        tot += v; sq += v * v; cnt += c
    mean, count = sh.get_mean("k")
    assert np.allclose(mean.numpy(), tot / (cnt + 1e-9), rtol=1e-5) and np.array_equal(count.numpy(), cnt)
    var, _ = sh.get_var("k")
    assert np.allclose(var.numpy(), np.maximum(sq / (cnt + 1) - (tot / (cnt + 1)) ** 2, 0), rtol=1e-4, atol=1e-6)
    assert sh.get_mean("missing") is None
