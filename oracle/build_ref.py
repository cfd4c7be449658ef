"""Recipe that compiles the UNMODIFIED reference extension (litegs/submodules/gaussian_raster) for sm_100a.

TEST INFRASTRUCTURE ONLY.  Sources are compiled where they lie under /root/reference (nothing is
copied into this repository); the only outputs are ``oracle/_ref/litegs_fused_ref.so`` and ``oracle/_ref/fused_ssim_cuda_ref.so`` (git-ignored,
NOT gpurun-ignored, so it travels to the GPU box).  It is the "reference itself run here" that pins the
CPU oracle (tests/test_gpu_vs_reference.py, tests/golden/make_golden.py) and the CUDA baseline that
``bench.py`` times next to ours ("ref_cuda").  Flags are the reference's own: -O3 --use_fast_math
(GR/setup.py:33-36) plus the arch.  The reference's build system (setup.py / CMake) is not run.
"""
from __future__ import annotations

import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_GR = "/root/reference/litegs/submodules/gaussian_raster"
NAME = "litegs_fused_ref"
SOURCES = ["binning.cu", "compact.cu", "cuda_errchk.cpp", "ext_cuda.cpp", "raster.cu", "transform.cu"]


REF_SSIM = "/root/reference/litegs/submodules/fused_ssim"
SSIM_NAME = "fused_ssim_cuda_ref"
SSIM_SOURCES = ["ssim.cu", "ext.cpp"]


def so_path(name: str = NAME):
    if not os.path.isdir(OUT):
        return None
    for f in os.listdir(OUT):
        if f.startswith(name) and f.endswith(".so"):
            return os.path.join(OUT, f)
    return None


def build(quiet: bool = False):
    """Build oracle/_ref/litegs_fused_ref.so if the reference tree is present; returns the path or None."""
    if so_path() is not None:
        return so_path()
    if not os.path.isdir(REF_GR):
        if not quiet:
            print("[build_ref] /root/reference not present: nothing to build")
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ["CXX"] = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++")
    os.environ["CC"] = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else os.environ.get("CC", "gcc")
    from torch.utils import cpp_extension
    # the reference's own setup.py strips torch's "no half operators" defines (GR/setup.py:4-16)
    for flag in ("-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_BFLOAT16_CONVERSIONS__",
                 "-D__CUDA_NO_HALF2_OPERATORS__"):
        if flag in cpp_extension.COMMON_NVCC_FLAGS:
            cpp_extension.COMMON_NVCC_FLAGS.remove(flag)
    cpp_extension.load(
        name=NAME,
        sources=[os.path.join(REF_GR, s) for s in SOURCES],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "--use_fast_math", "-gencode", "arch=compute_100a,code=sm_100a"],
        build_directory=OUT,
        verbose=not quiet,
        is_python_module=True,
    )
    return so_path()


def build_ssim(quiet: bool = False):
    """Build oracle/_ref/fused_ssim_cuda_ref.so from the reference's fused_ssim/{ssim.cu,ext.cpp} with the reference's own
    nvcc flags (fused_ssim/setup.py:33: --maxrregcount=32 --use_fast_math) plus the arch; returns the path or None."""
    if so_path(SSIM_NAME) is not None:
        return so_path(SSIM_NAME)
    if not os.path.isdir(REF_SSIM):
        if not quiet:
            print("[build_ref] /root/reference not present: nothing to build")
        return None
    out = os.path.join(OUT, "ssim_build")
    os.makedirs(out, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ["CXX"] = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else os.environ.get("CXX", "g++")
    os.environ["CC"] = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else os.environ.get("CC", "gcc")
    from torch.utils import cpp_extension
    cpp_extension.load(
        name=SSIM_NAME,
        sources=[os.path.join(REF_SSIM, s) for s in SSIM_SOURCES],
        extra_cflags=["-O3"],
        extra_cuda_cflags=["-O3", "--maxrregcount=32", "--use_fast_math", "-gencode", "arch=compute_100a,code=sm_100a"],
        build_directory=out,
        verbose=not quiet,
        is_python_module=True,
    )
    import shutil
    for f in os.listdir(out):
        if f.startswith(SSIM_NAME) and f.endswith(".so"):
            shutil.copy2(os.path.join(out, f), os.path.join(OUT, f))
    return so_path(SSIM_NAME)


REF_ROOT = "/root/reference"
PY_OUT = os.path.join(os.path.dirname(HERE), "baseline", "_ref")


def stage_python(quiet: bool = False):
    """Stage the UNMODIFIED reference Python package (``litegs/`` without its native submodules, plus the example scripts)
    under ``baseline/_ref`` (the location the bench contract reserves for an unmodified reference install) -- git-ignored,
    NOT gpurun-ignored -- so that the drop-in test
    (tests/test_gpu_dropin.py) and the ``ref_cuda`` bench leg can import the reference's own ``wrapper.py`` /
    ``render/__init__.py`` / ``training/trainer.py`` on the GPU box, where /root/reference does not exist.  Files are byte
    copies (checked by the test against a manifest of SHA-256 digests written here); nothing is copied into tracked paths."""
    import hashlib
    import json
    import shutil
    if not os.path.isdir(os.path.join(REF_ROOT, "litegs")):
        if not quiet:
            print("[build_ref] /root/reference not present: nothing to stage")
        return PY_OUT if os.path.isdir(PY_OUT) else None
    os.makedirs(PY_OUT, exist_ok=True)
    manifest = {}
    for top in ("litegs", "example_train.py", "example_metrics.py"):
        src = os.path.join(REF_ROOT, top)
        if os.path.isfile(src):
            shutil.copy2(src, os.path.join(PY_OUT, top))
            manifest[top] = hashlib.sha256(open(src, "rb").read()).hexdigest()
            continue
        for d, dirs, files in os.walk(src):
            dirs[:] = [x for x in dirs if x not in ("submodules", "__pycache__")]
            rel = os.path.relpath(d, REF_ROOT)
            os.makedirs(os.path.join(PY_OUT, rel), exist_ok=True)
            for f in files:
                if f.endswith(".py"):
                    shutil.copy2(os.path.join(d, f), os.path.join(PY_OUT, rel, f))
                    manifest[os.path.join(rel, f)] = hashlib.sha256(open(os.path.join(d, f), "rb").read()).hexdigest()
    json.dump(manifest, open(os.path.join(PY_OUT, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return PY_OUT


def reference_python_path():
    """Directory to put on sys.path to import the reference's ``litegs`` package: the mounted tree if present, else the
    staged copy, else None."""
    if os.path.isdir(os.path.join(REF_ROOT, "litegs")):
        return REF_ROOT
    return PY_OUT if os.path.isdir(os.path.join(PY_OUT, "litegs")) else None


def _load(name: str):
    p = so_path(name)
    if p is None:
        return None
    import torch  # noqa: F401  (registers the ATen symbols the extension links against)
    spec = importlib.util.spec_from_file_location(name, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Import the prebuilt reference module (needs torch; used on the GPU box). Returns None if absent."""
    return _load(NAME)


def load_ssim():
    """Import the prebuilt reference fused_ssim_cuda module (fusedssim, fusedssim_backward, fusedl1ssim_loss,
    fusedl1ssim_loss_backward). Returns None if absent."""
    return _load(SSIM_NAME)


if __name__ == "__main__":
    print(build(quiet="-q" in sys.argv))
    print(build_ssim(quiet="-q" in sys.argv))
    print(stage_python(quiet="-q" in sys.argv))
