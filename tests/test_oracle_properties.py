"""Property tests (hypothesis) of oracle pieces whose CUDA counterparts use a different algorithm:

* tile ranges: the oracle streams the sorted key list and marks boundaries (GR/binning.cu:228-265); the kernel gives every
  tile one lower_bound (csrc/binning.cu: tile_range_bsearch_kernel).  Here the search formulation is restated in Python
  and must equal the oracle on arbitrary sorted key lists, with and without the last-tile fix;
* SSIM: symmetry in its arguments, SSIM(x, x) = 1, value range, and zero L1+SSIM loss / gradient for identical images.
"""
import bisect

import numpy as np
from hypothesis import given, settings, strategies as st

import oracle


def ranges_by_search(keys, max_tile, fix_last):
    L = len(keys)
    out = []
    for t in range(max_tile + 2):
        lo = bisect.bisect_left(keys, t)
        r = -1
        if lo < L and keys[lo] == t:
            r = lo
        elif t >= 1 and lo > 0 and keys[lo - 1] == t - 1 and (lo < L or fix_last):
            r = lo
        if t == max_tile + 1:
            r = L
        out.append(r)
    return np.array(out, np.int32)


@settings(max_examples=300, deadline=None)
@given(st.integers(1, 60).flatmap(lambda mt: st.tuples(st.just(mt), st.lists(st.integers(1, mt), min_size=1, max_size=200))),
       st.booleans())
def test_tile_range_equals_one_lower_bound_per_tile(case, fix_last):
    max_tile, keys = case
    keys = sorted(keys)
    got = oracle.tileRange(np.array([keys], np.int32), max_tile, fix_last=fix_last)[0]
    assert np.array_equal(got, ranges_by_search(keys, max_tile, fix_last))


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 3), st.integers(1, 3), st.integers(1, 40), st.integers(1, 40), st.integers(0, 2 ** 31 - 1))
def test_ssim_properties(B, CH, H, W, seed):
    rng = np.random.default_rng(seed)
    x, y = rng.random((B, CH, H, W)), rng.random((B, CH, H, W))
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mxy = oracle.fusedssim(C1, C2, x, y, False)[0]
    myx = oracle.fusedssim(C1, C2, y, x, False)[0]
    assert np.abs(mxy - myx).max() < 1e-12                          # symmetric
    assert mxy.max() <= 1.0 + 1e-12 and mxy.min() >= -1.0 - 1e-12   # Cauchy-Schwarz on the windowed moments
    mxx = oracle.fusedssim(C1, C2, x, x, False)[0]
    assert np.abs(mxx - 1.0).max() < 1e-12
    lm, d0, d1, d2 = oracle.fusedl1ssim_loss(0.2, C1, C2, x, x, True)
    g = oracle.fusedl1ssim_loss_backward(0.2, C1, C2, x, x, np.full(x.shape, 1.0 / x.size), d0, d1, d2)
    assert np.abs(lm).max() < 1e-12 and np.abs(g).max() < 1e-12     # identical images: zero loss, stationary
