"""world_size-2 gloo test (CPU) of the data-parallel host logic: view sharding, placement of compacted per-view
gradients into the dense buffer, the all-reduce, and the parameter views of the result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from litegs_b200.dist import PARAM_ORDER, GradAccumulator, param_rows, shard_views


def test_shard_views_partition():
    for world in (1, 2, 3, 8):
        allv = sorted(v for r in range(world) for v in shard_views(64, r, world))
        assert allv == list(range(64))
        sizes = [len(shard_views(64, r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_views(8, 2, 2)


def test_param_rows_layout():
    shapes = {"xyz": (3, 5, 4), "scale": (3, 5, 4), "rot": (4, 5, 4), "sh_0": (1, 3, 5, 4), "sh_rest": (15, 3, 5, 4), "opacity": (1, 5, 4)}
    rows = param_rows(shapes)
    assert rows["xyz"] == slice(0, 3) and rows["sh_rest"] == slice(13, 58) and rows["opacity"] == slice(58, 59)


def _fake_params(C=6, S=4, R=3):
    return {"xyz": torch.zeros(3, C, S), "scale": torch.zeros(3, C, S), "rot": torch.zeros(4, C, S), "sh_0": torch.zeros(1, 3, C, S),
            "sh_rest": torch.zeros(R, 3, C, S), "opacity": torch.zeros(1, C, S)}


def _view_grads(view, C=6, S=4, R=3):
    """Deterministic compacted gradients of a fake view: visible chunks depend on the view index."""
    rng = np.random.default_rng(view)
    ids = np.sort(rng.choice(C, size=rng.integers(1, C + 1), replace=False)).astype(np.int64)
    A = len(ids) + 1                         # one allocated-but-invalid tail chunk, as with the feedback sizing
    g = {"xyz": rng.normal(size=(3, A, S)), "scale": rng.normal(size=(3, A, S)), "rot": rng.normal(size=(4, A, S)),
         "sh_0": rng.normal(size=(1, 3, A, S)), "sh_rest": rng.normal(size=(R, 3, A, S)), "opacity": rng.normal(size=(1, A, S))}
    g = {k: torch.from_numpy(v.astype(np.float32)) for k, v in g.items()}
    ids_t = torch.from_numpy(np.concatenate([ids, [0]]))
    return g, ids_t, torch.tensor([len(ids)], dtype=torch.int32)


def _expected(n_views, C=6, S=4, R=3):
    p = _fake_params(C, S, R)
    out = {k: torch.zeros_like(v) for k, v in p.items()}
    for v in range(n_views):
        g, ids, n = _view_grads(v, C, S, R)
        n = int(n)
        for k in PARAM_ORDER:
            out[k][..., ids[:n], :] += g[k][..., :n, :]
    return out


def _worker(rank, world, port, n_views, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        acc = GradAccumulator(_fake_params())
        acc.zero_()
        for v in shard_views(n_views, rank, world):
            g, ids, n = _view_grads(v)
            acc.add_view(g, ids, n)
        acc.all_reduce()
        acc.wait()
        q.put((rank, {k: t.clone().numpy() for k, t in acc.grads().items()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_views = 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_views, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = _expected(n_views)
    for r in range(2):
        for k in PARAM_ORDER:
            assert np.allclose(results[r][k], exp[k].numpy(), atol=1e-6), (r, k)
    # identical on both ranks (replicated optimiser steps stay in lock-step)
    for k in PARAM_ORDER:
        assert np.array_equal(results[0][k], results[1][k])
    # the chunk marks ride in the same all-reduce: non-zero exactly at the chunks visible in some view on some rank
    seen = np.zeros(6, bool)
    for v in range(n_views):
        _, ids, n = _view_grads(v)
        seen[ids[: int(n)].numpy()] = True
    for r in range(2):
        assert np.array_equal(results[r]["_touched"] > 0, seen), r
