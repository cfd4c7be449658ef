"""COLMAP sparse-model I/O and the conversions between COLMAP cameras/points and this package's scene layout -- the
data format on the input side of the render path (BASELINE.json config 5 needs a synthetic scene
on disk in this format).

Binary layouts follow COLMAP's ``src/base/reconstruction.cc`` (the reader the reference uses is
``litegs/io_manager/colmap.py:78-160,250-279``): little endian,

    cameras.bin   u64 n | per camera: i32 id, i32 model, u64 width, u64 height, f64 params[model]
    images.bin    u64 n | per image : i32 id, f64 qvec[4] (w,x,y,z), f64 tvec[3], i32 camera_id, name '\\0',
                                      u64 m, m x (f64 x, f64 y, i64 point3D_id)
    points3D.bin  u64 n | per point : u64 id, f64 xyz[3], u8 rgb[3], f64 error, u64 t, t x (i32 image_id, i32 point2D_idx)

Camera conventions are the reference's (``litegs/data.py:33-51,72-77``): ``view_matrix = [R|t]^T`` (row vectors) from the
COLMAP world->camera rotation ``R(qvec)`` and translation ``tvec``; PINHOLE intrinsics (fx, fy, cx, cy) give
``proj_matrix`` through fx / (W/2), fy / (H/2) with z_near 0.01, z_far 5000.  Pure numpy; nothing here touches the GPU.
"""
from __future__ import annotations

import collections
import os
import struct
from typing import Dict, List, Tuple

import numpy as np

from . import scene

Camera = collections.namedtuple("Camera", ["id", "model", "width", "height", "params"])
Image = collections.namedtuple("Image", ["id", "qvec", "tvec", "camera_id", "name", "xys", "point3D_ids"])
Point3D = collections.namedtuple("Point3D", ["id", "xyz", "rgb", "error", "track"])

MODEL_PARAMS = {0: ("SIMPLE_PINHOLE", 3), 1: ("PINHOLE", 4), 2: ("SIMPLE_RADIAL", 4), 3: ("RADIAL", 5), 4: ("OPENCV", 8),
                5: ("OPENCV_FISHEYE", 8), 6: ("FULL_OPENCV", 12), 7: ("FOV", 5), 8: ("SIMPLE_RADIAL_FISHEYE", 4),
                9: ("RADIAL_FISHEYE", 5), 10: ("THIN_PRISM_FISHEYE", 12)}
MODEL_IDS = {name: mid for mid, (name, _) in MODEL_PARAMS.items()}
SH_C0 = 0.28209479177387814


# ---- rotations ----------------------------------------------------------------------------------------------------------

def qvec_to_rotmat(q) -> np.ndarray:
    """COLMAP quaternion (w, x, y, z) -> world->camera rotation (litegs/utils/__init__.py:7-17)."""
    w, x, y, z = [float(v) for v in q]
    return np.array([[1 - 2 * y * y - 2 * z * z, 2 * x * y - 2 * w * z, 2 * z * x + 2 * w * y],
                     [2 * x * y + 2 * w * z, 1 - 2 * x * x - 2 * z * z, 2 * y * z - 2 * w * x],
                     [2 * z * x - 2 * w * y, 2 * y * z + 2 * w * x, 1 - 2 * x * x - 2 * y * y]], dtype=np.float64)


def rotmat_to_qvec(R) -> np.ndarray:
    """Inverse of qvec_to_rotmat (eigenvector form, as COLMAP's rotmat2qvec); w >= 0."""
    Rxx, Ryx, Rzx, Rxy, Ryy, Rzy, Rxz, Ryz, Rzz = np.asarray(R, np.float64).flat
    K = np.array([[Rxx - Ryy - Rzz, 0, 0, 0],
                  [Ryx + Rxy, Ryy - Rxx - Rzz, 0, 0],
                  [Rzx + Rxz, Rzy + Ryz, Rzz - Rxx - Ryy, 0],
                  [Ryz - Rzy, Rzx - Rxz, Rxy - Ryx, Rxx + Ryy + Rzz]]) / 3.0
    vals, vecs = np.linalg.eigh(K)
    q = vecs[[3, 0, 1, 2], np.argmax(vals)]
    return -q if q[0] < 0 else q


# ---- binary files -------------------------------------------------------------------------------------------------------

def write_cameras_bin(path: str, cameras: Dict[int, Camera]) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(cameras)))
        for c in cameras.values():
            mid = MODEL_IDS[c.model] if isinstance(c.model, str) else int(c.model)
            n = MODEL_PARAMS[mid][1]
            if len(c.params) != n:
                raise ValueError(f"camera {c.id}: model {MODEL_PARAMS[mid][0]} takes {n} parameters, got {len(c.params)}")
            f.write(struct.pack("<iiQQ", int(c.id), mid, int(c.width), int(c.height)))
            f.write(struct.pack("<" + "d" * n, *[float(p) for p in c.params]))


def read_cameras_bin(path: str) -> Dict[int, Camera]:
    out = {}
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            cid, mid, w, h = struct.unpack("<iiQQ", f.read(24))
            name, k = MODEL_PARAMS[mid]
            out[cid] = Camera(cid, name, w, h, np.array(struct.unpack("<" + "d" * k, f.read(8 * k))))
    return out


def write_images_bin(path: str, images: Dict[int, Image]) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(images)))
        for im in images.values():
            f.write(struct.pack("<idddddddi", int(im.id), *[float(v) for v in im.qvec], *[float(v) for v in im.tvec], int(im.camera_id)))
            f.write(im.name.encode("utf-8") + b"\x00")
            xys = np.asarray(im.xys, np.float64).reshape(-1, 2)
            ids = np.asarray(im.point3D_ids, np.int64).reshape(-1)
            f.write(struct.pack("<Q", len(ids)))
            for (x, y), pid in zip(xys, ids):
                f.write(struct.pack("<ddq", float(x), float(y), int(pid)))


def read_images_bin(path: str) -> Dict[int, Image]:
    out = {}
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            v = struct.unpack("<idddddddi", f.read(64))
            name = b""
            while True:
                ch = f.read(1)
                if ch == b"\x00" or ch == b"":
                    break
                name += ch
            (m,) = struct.unpack("<Q", f.read(8))
            raw = np.frombuffer(f.read(24 * m), dtype=np.dtype([("x", "<f8"), ("y", "<f8"), ("id", "<i8")]))
            out[v[0]] = Image(v[0], np.array(v[1:5]), np.array(v[5:8]), v[8], name.decode("utf-8"),
                              np.stack([raw["x"], raw["y"]], 1) if m else np.zeros((0, 2)), raw["id"].copy())
    return out


def write_points3d_bin(path: str, points: Dict[int, Point3D]) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(points)))
        for p in points.values():
            f.write(struct.pack("<QdddBBBd", int(p.id), *[float(v) for v in p.xyz], *[int(v) for v in p.rgb], float(p.error)))
            tr = np.asarray(p.track, np.int32).reshape(-1, 2)
            f.write(struct.pack("<Q", len(tr)))
            f.write(tr.astype("<i4").tobytes())


def read_points3d_bin(path: str) -> Dict[int, Point3D]:
    out = {}
    with open(path, "rb") as f:
        (n,) = struct.unpack("<Q", f.read(8))
        for _ in range(n):
            v = struct.unpack("<QdddBBBd", f.read(43))
            (t,) = struct.unpack("<Q", f.read(8))
            tr = np.frombuffer(f.read(8 * t), dtype="<i4").reshape(-1, 2).copy()
            out[v[0]] = Point3D(v[0], np.array(v[1:4]), np.array(v[4:7], np.uint8), v[7], tr)
    return out


def write_model(root: str, cameras, images, points) -> str:
    """root/sparse/0/{cameras,images,points3D}.bin (the directory layout load_colmap_result expects, colmap.py:322-325)."""
    d = os.path.join(root, "sparse", "0")
    os.makedirs(d, exist_ok=True)
    write_cameras_bin(os.path.join(d, "cameras.bin"), cameras)
    write_images_bin(os.path.join(d, "images.bin"), images)
    write_points3d_bin(os.path.join(d, "points3D.bin"), points)
    return d


def read_model(root: str):
    d = os.path.join(root, "sparse", "0")
    return (read_cameras_bin(os.path.join(d, "cameras.bin")), read_images_bin(os.path.join(d, "images.bin")),
            read_points3d_bin(os.path.join(d, "points3D.bin")))


# ---- cameras ------------------------------------------------------------------------------------------------------------

def camera_from_colmap(qvec, tvec, pinhole_params, width: int, height: int, z_near: float = 0.01, z_far: float = 5000.0):
    """COLMAP pose + PINHOLE intrinsics -> dict(view [1,4,4], proj [1,4,4], frustumplane [1,6,4]) in the reference's
    row-vector convention (litegs/data.py:33-51 for the projection, :72-77 for the view matrix)."""
    R = qvec_to_rotmat(qvec)
    Rt = np.zeros((4, 4), np.float64)
    Rt[:3, :3] = R
    Rt[:3, 3] = np.asarray(tvec, np.float64)
    Rt[3, 3] = 1.0
    V = np.float32(Rt).T.copy()
    fx, fy = float(pinhole_params[0]), float(pinhole_params[1])
    rx, ry = fx / (width * 0.5), fy / (height * 0.5)
    P = np.array([[rx, 0, 0, 0], [0, ry, 0, 0], [0, 0, z_far / (z_far - z_near), -z_far * z_near / (z_far - z_near)], [0, 0, 1, 0]],
                 dtype=np.float32).T.copy()
    return dict(view=V[None], proj=P[None], frustumplane=scene.frustum_planes(V, P)[None].copy())


def camera_to_colmap(cam: dict, width: int, height: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Inverse of camera_from_colmap for a centred pinhole: (qvec, tvec, [fx, fy, cx, cy])."""
    Rt = np.asarray(cam["view"], np.float64).reshape(4, 4).T
    P = np.asarray(cam["proj"], np.float64).reshape(4, 4).T
    return (rotmat_to_qvec(Rt[:3, :3]), Rt[:3, 3].copy(),
            np.array([P[0, 0] * width * 0.5, P[1, 1] * height * 0.5, width * 0.5, height * 0.5]))


# ---- points -> initial Gaussians ----------------------------------------------------------------------------------------

def mean_knn_dist2(xyz: np.ndarray, k: int = 3) -> np.ndarray:
    """Mean squared distance to the k nearest other points (what simple_knn.distCUDA2 returns), xyz [N,3]."""
    from scipy.spatial import cKDTree
    n = xyz.shape[0]
    kk = min(k + 1, n)
    d, _ = cKDTree(xyz).query(xyz, k=kk)
    d = np.asarray(d, np.float64).reshape(n, kk)
    return (d[:, 1:] ** 2).sum(1) / max(kk - 1, 1) if kk > 1 else np.zeros(n)


def gaussians_from_points(xyz: np.ndarray, rgb_u8: np.ndarray, sh_degree: int = 3, chunk: int = 128, morton: bool = True):
    """SfM points -> the reference's initial Gaussians (litegs/scene/point.py:7-19: isotropic log-scale from the mean
    squared 3-NN distance clamped at 1e-7, identity rotation, SH band 0 from the colour, opacity logit of 0.1), Morton
    sorted and clustered like scene.make_scene.  Returns the same dict layout (parameters + cluster_origin/extend)."""
    xyz = np.asarray(xyz, np.float64).reshape(-1, 3)
    n = xyz.shape[0]
    order = scene.morton_order(xyz.T.astype(np.float32)) if morton else np.arange(n)
    xyz = xyz[order]
    col = np.asarray(rgb_u8, np.float64).reshape(-1, 3)[order] / 255.0
    dist2 = np.maximum(mean_knn_dist2(xyz), 1e-7)
    K = (sh_degree + 1) ** 2
    P = {
        "xyz": xyz.T.astype(np.float32),
        "scale": np.repeat(np.log(np.sqrt(dist2))[None], 3, 0).astype(np.float32),
        "rot": np.concatenate([np.ones((1, n)), np.zeros((3, n))]).astype(np.float32),
        "sh_0": ((col.T - 0.5) / SH_C0)[None].astype(np.float32),
        "sh_rest": np.zeros((K - 1, 3, n), np.float32),
        "opacity": np.full((1, n), np.log(0.1 / 0.9), np.float32),
    }
    out = {k: scene.cluster(v, chunk) for k, v in P.items()}
    out["cluster_origin"], out["cluster_extend"] = scene.cluster_aabb(out["xyz"], out["scale"], out["rot"])
    return out


# ---- a synthetic scene on disk ------------------------------------------------------------------------------------------

def write_synthetic_dataset(root: str, xyz: np.ndarray, rgb_u8: np.ndarray, n_views: int, width: int, height: int, render_fn=None,
                            image_dir: str = "images") -> List[str]:
    """Writes root/sparse/0/*.bin for the Fibonacci-lattice cameras of scene.make_camera and, when ``render_fn(i, cam) ->
    uint8 [H,W,3]`` is given, root/images/view_XXXX.png (PIL).  Returns the image names in COLMAP order."""
    cams = {1: None}
    images, names = {}, []
    for i in range(n_views):
        cam = scene.make_camera(i, n_views, width, height)
        q, t, intr = camera_to_colmap(cam, width, height)
        if cams[1] is None:
            cams[1] = Camera(1, "PINHOLE", width, height, intr)
        name = f"view_{i:04d}.png"
        images[i + 1] = Image(i + 1, q, t, 1, name, np.zeros((0, 2)), np.zeros((0,), np.int64))
        names.append(name)
        if render_fn is not None:
            import PIL.Image
            os.makedirs(os.path.join(root, image_dir), exist_ok=True)
            PIL.Image.fromarray(np.asarray(render_fn(i, cam), np.uint8)).save(os.path.join(root, image_dir, name))
    pts = {j + 1: Point3D(j + 1, xyz[j], rgb_u8[j], 0.0, np.zeros((0, 2), np.int32)) for j in range(xyz.shape[0])}
    write_model(root, cams, images, pts)
    return names
