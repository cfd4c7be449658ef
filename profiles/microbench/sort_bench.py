"""Times lgs_sort_pairs_u16 / _u32 (both implementations) on pipeline-shaped inputs.
usage: python profiles/microbench/sort_bench.py [reps]"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from litegs_b200 import _lib  # noqa: E402


def run(keys, vals, begin, end, impl, reps):
    n = keys.numel()
    sfx = "_u16" if keys.dtype == torch.int16 else "_u32"
    nb = ctypes.c_size_t(0)
    _lib.call(f"lgs_sort_pairs{sfx}_workspace_bytes", n, ctypes.byref(nb))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=keys.device)
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    _lib.call("lgs_set_sort_impl", impl)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    f = lambda: _lib.call(f"lgs_sort_pairs{sfx}", ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(ko.data_ptr()),
                          ctypes.c_void_p(vals.data_ptr()), ctypes.c_void_p(vo.data_ptr()), n, begin, end, ctypes.c_void_p(ws.data_ptr()),
                          ctypes.c_size_t(nb.value), st)
    for _ in range(3 if reps > 1 else 0):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    n = 10_000_000
    start = torch.randint(1, 16000, (n // 7 + 1,), generator=g, dtype=torch.int64)
    k16 = (start[:, None] + torch.arange(7)[None, :]).reshape(-1)[:n].to(torch.int16).to(dev)
    v = torch.arange(n, dtype=torch.int32, device=dev)
    m = 1_000_064
    z = (torch.rand(m, generator=g) * 2.0 + 2.0).view(torch.int32).to(dev)
    vz = torch.arange(m, dtype=torch.int32, device=dev)
    for name, impl in (("lgs", 1), ("cub", 0)):
        print(f"{name}: tile sort 10M x 14 bit {run(k16, v, 0, 14, impl, reps):8.1f} us   depth sort 1M x 32 bit "
              f"{run(z, vz, 0, 32, impl, reps):8.1f} us   1M x 24 bit {run(z, vz, 0, 24, impl, reps):8.1f} us")


if __name__ == "__main__":
    main()
