"""CPU tests of the optimizer row's host logic and of its oracle: the oracle's Adam restatement against the reference
formula written with numpy (GR/compact.cu:320-344: no bias correction, eps outside the square root), and the learning-rate
schedule against the reference's closed form (training/optimizer.py:46-72)."""
import math
import types

import numpy as np

import oracle


def test_oracle_adam_is_the_reference_update_on_visible_chunks_only():
    rng = np.random.default_rng(0)
    R, C, S = 4, 7, 8
    p = rng.normal(size=(R, C, S)); m = rng.normal(size=(R, C, S)) * 0.1; v = rng.random((R, C, S)) * 0.01
    ids = np.array([1, 4, 5, 0], dtype=np.int64)          # last entry is an allocated-but-invalid tail slot
    nvis = 3
    g = rng.normal(size=(R, len(ids), S))
    lr, b1, b2, eps = 1e-2, 0.9, 0.999, 1e-15
    p0, m0, v0 = p.copy(), m.copy(), v.copy()
    oracle.adamUpdate(p, g, m, v, ids, np.array([nvis], np.int32), lr, b1, b2, eps)
    for a in range(nvis):
        c = ids[a]
        e1 = b1 * m0[:, c] + (1 - b1) * g[:, a]
        e2 = b2 * v0[:, c] + (1 - b2) * g[:, a] ** 2
        assert np.allclose(m[:, c], e1, rtol=0, atol=1e-15) and np.allclose(v[:, c], e2, rtol=0, atol=1e-15)
        assert np.allclose(p[:, c], p0[:, c] - lr * e1 / (np.sqrt(e2) + eps), rtol=0, atol=1e-14)
    untouched = [c for c in range(C) if c not in ids[:nvis]]
    assert np.array_equal(p[:, untouched], p0[:, untouched]) and np.array_equal(m[:, untouched], m0[:, untouched])
    assert np.array_equal(v[:, untouched], v0[:, untouched])          # moments of invisible chunks do not decay


def test_scheduler_is_the_reference_log_linear_schedule():
    from litegs_b200.optimizer import Scheduler
    opt = types.SimpleNamespace(lr={"xyz": 0.0})
    lr0, lr1, n = 1.6e-4, 1.6e-6, 100
    s = Scheduler(opt, lr0, lr1, max_epochs=n)
    assert math.isclose(opt.lr["xyz"], lr0, rel_tol=1e-12)
    for k in range(1, 151):
        s.step()
        t = min(k / n, 1.0)
        assert math.isclose(opt.lr["xyz"], math.exp(math.log(lr0) * (1 - t) + math.log(lr1) * t), rel_tol=1e-12)
    z = Scheduler(types.SimpleNamespace(lr={"xyz": 1.0}), 0.0, 0.0)
    assert z.opt.lr["xyz"] == 0.0
