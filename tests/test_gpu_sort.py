"""GPU tests of the stable LSD radix sort behind lgs_sort_pairs_u16/_u32 (both implementations) against
torch's stable sort on the same bit range: keys AND payload order must be identical (integer work: bit-exact).
Covers the shapes the pipeline uses (14/16 tile bits on u16 keys, 24/32 depth bits on u32 keys), ragged tails,
single-element and empty inputs, a sub-range of bits, and a constant key (every key in one digit)."""
import ctypes

import pytest
import torch

from litegs_b200 import _lib

pytestmark = pytest.mark.gpu


def _sort(keys, vals, begin, end, impl):
    dev = keys.device
    n = keys.numel()
    u16 = keys.dtype == torch.int16
    sfx = "_u16" if u16 else "_u32"
    nb = ctypes.c_size_t(0)
    _lib.call(f"lgs_sort_pairs{sfx}_workspace_bytes", max(n, 1), ctypes.byref(nb))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=dev)
    ko, vo = torch.full_like(keys, -1), torch.full_like(vals, -1)
    # impl: 1 = own sort, histogram / row-scan / scatter passes (default) | 2 = own sort, onesweep form | 0 = cub
    _lib.call("lgs_set_sort_impl", 0 if impl == 0 else 1)
    _lib.call("lgs_set_radix_form", 1 if impl == 2 else 0)
    try:
        _lib.call(f"lgs_sort_pairs{sfx}", ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(ko.data_ptr()), ctypes.c_void_p(vals.data_ptr()),
                  ctypes.c_void_p(vo.data_ptr()), n, begin, end, ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(nb.value),
                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    finally:
        _lib.call("lgs_set_sort_impl", 1)
        _lib.call("lgs_set_radix_form", 0)
    torch.cuda.synchronize()
    return ko, vo


def _expect(keys, vals, begin, end):
    k = keys.to(torch.int64) & (0xFFFF if keys.dtype == torch.int16 else 0xFFFFFFFF)
    digit = (k >> begin) & ((1 << (end - begin)) - 1)
    order = torch.sort(digit, stable=True).indices
    return keys[order], vals[order]


CASES = [  # dtype, n, begin, end, key generator
    (torch.int16, 1, 0, 14, "uniform"), (torch.int16, 100, 0, 14, "uniform"), (torch.int16, 2048, 0, 14, "uniform"),
    (torch.int16, 2049, 0, 16, "uniform"), (torch.int16, 1_000_003, 0, 14, "uniform"), (torch.int16, 5_000_001, 0, 14, "runs"),
    (torch.int16, 4_194_305, 0, 13, "uniform"), (torch.int16, 300_000, 3, 11, "uniform"), (torch.int16, 70_000, 0, 9, "uniform"),
    (torch.int32, 1, 0, 32, "uniform"), (torch.int32, 4097, 0, 32, "uniform"), (torch.int32, 1_000_064, 0, 32, "depth"),
    (torch.int32, 1_000_064, 0, 24, "depth"), (torch.int32, 5_000_000, 0, 32, "depth"), (torch.int32, 123_457, 5, 22, "uniform"),
    (torch.int32, 50_000, 0, 32, "constant"), (torch.int32, 50_000, 0, 0, "uniform"),
]


@pytest.mark.parametrize("impl", [1, 2, 0], ids=["lgs", "lgs_onesweep", "cub"])
@pytest.mark.parametrize("dtype,n,begin,end,kind", CASES)
def test_sort_pairs_matches_stable_sort(cuda, impl, dtype, n, begin, end, kind):
    g = torch.Generator(device="cpu").manual_seed(n + 31 * end)
    hi = 1 << (16 if dtype == torch.int16 else 32)
    if kind == "uniform":
        k = torch.randint(0, hi, (n,), generator=g, dtype=torch.int64)
    elif kind == "runs":                       # what emit produces: short runs of consecutive tile ids
        start = torch.randint(1, 16000, (n // 7 + 1,), generator=g, dtype=torch.int64)
        k = (start[:, None] + torch.arange(7)[None, :]).reshape(-1)[:n]
    elif kind == "depth":                      # float bits of view-space z in [0.2, 6): top byte almost constant
        z = torch.rand(n, generator=g) * 5.8 + 0.2
        k = z.view(torch.int32).to(torch.int64)
        k[::17] = 0xFFFFFFFF                   # culled splats carry the all-ones key
    else:
        k = torch.full((n,), 0x40490FDB, dtype=torch.int64)
    if dtype == torch.int16:
        keys = (k & 0xFFFF).to(torch.int32).to(torch.int16)     # wraps to the signed view of the same bits
    else:
        keys = torch.where(k >= (1 << 31), k - (1 << 32), k).to(torch.int32)
    vals = torch.arange(n, dtype=torch.int32)
    keys, vals = keys.to(cuda), vals.to(cuda)
    ko, vo = _sort(keys, vals, begin, end, impl)
    ek, ev = _expect(keys, vals, begin, end)
    assert torch.equal(vo, ev), f"payload order differs at {int((vo != ev).nonzero()[0])}"
    assert torch.equal(ko, ek)


@pytest.mark.parametrize("impl", [1, 2, 0], ids=["lgs", "lgs_onesweep", "cub"])
def test_rebased_depth_sort_orders_keys_inside_the_range(cuda, impl):
    """lgs_sort_pairs_u32_rebased: keys inside [bias, bias + 2^bits) come out in full-key stable order; the keys outside
    (culled splats, all ones) may land anywhere but must all still be present."""
    g = torch.Generator(device="cpu").manual_seed(7)
    n = 700_001
    z = torch.rand(n, generator=g) * 3.4 + 1.3                       # crosses the 2.0 and 4.0 exponent boundaries
    k = z.view(torch.int32).to(torch.int64)
    k[::11] = 0xFFFFFFFF
    inside = k != 0xFFFFFFFF
    kmin, kmax = int(k[inside].min()), int(k[inside].max())
    bits = max(1, (kmax - kmin).bit_length())
    assert bits <= 24 < (kmin ^ kmax).bit_length()
    keys = torch.where(k >= (1 << 31), k - (1 << 32), k).to(torch.int32).to(cuda)
    vals = torch.arange(n, dtype=torch.int32, device=cuda)
    nb = ctypes.c_size_t(0)
    _lib.call("lgs_sort_pairs_u32_workspace_bytes", n, ctypes.byref(nb))
    ws = torch.empty(nb.value, dtype=torch.uint8, device=cuda)
    ko, vo = torch.empty_like(keys), torch.empty_like(vals)
    _lib.call("lgs_set_sort_impl", 0 if impl == 0 else 1)
    _lib.call("lgs_set_radix_form", 1 if impl == 2 else 0)
    try:
        _lib.call("lgs_sort_pairs_u32_rebased", ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(ko.data_ptr()),
                  ctypes.c_void_p(vals.data_ptr()), ctypes.c_void_p(vo.data_ptr()), n, kmin, bits, ctypes.c_void_p(ws.data_ptr()),
                  ctypes.c_size_t(nb.value), None)
    finally:
        _lib.call("lgs_set_sort_impl", 1)
        _lib.call("lgs_set_radix_form", 0)
    torch.cuda.synchronize()
    vo_c, ko_c = vo.cpu().long(), ko.cpu()
    assert torch.equal(torch.sort(vo_c).values, torch.arange(n))      # a permutation
    assert torch.equal(ko_c.long() & 0xFFFFFFFF, k[vo_c])              # keys travel with their payload
    got = vo_c[inside[vo_c]]                                           # order of the keys that matter
    want = torch.sort(torch.where(inside, k, torch.full_like(k, 1 << 40)), stable=True).indices[: int(inside.sum())]
    assert torch.equal(got, want)


def test_sort_pairs_empty_and_bad_range(cuda):
    k = torch.zeros(8, dtype=torch.int16, device=cuda)
    v = torch.zeros(8, dtype=torch.int32, device=cuda)
    ws = torch.empty(1 << 20, dtype=torch.uint8, device=cuda)
    args = lambda n, b, e: (ctypes.c_void_p(k.data_ptr()), ctypes.c_void_p(k.data_ptr()), ctypes.c_void_p(v.data_ptr()),
                            ctypes.c_void_p(v.data_ptr()), n, b, e, ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel()), None)
    _lib.call("lgs_sort_pairs_u16", *args(0, 0, 14))            # n = 0: no-op
    with pytest.raises(_lib.LiteGSB200Error):
        _lib.call("lgs_sort_pairs_u16", *args(8, 0, 17))
    with pytest.raises(_lib.LiteGSB200Error):
        _lib.call("lgs_sort_pairs_u16", *args(8, 0, 14)[:8] + (ctypes.c_size_t(16), None))
