"""CPU, world_size 2 over gloo: the k-sharding / gather / all-reduce host logic of pysktb_b200.dist."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from pysktb_b200 import dist as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k = np.random.default_rng(0).random((37, 3))
        fake = lambda ks: np.stack([ks.sum(axis=1), ks[:, 0] * 2.0, ks[:, 1] - ks[:, 2]])   # (3, nk_local)
        full = D.solve_kpath_sharded(None, list(k), gather=True, solver=fake)
        local, (lo, hi) = D.solve_kpath_sharded(None, list(k), gather=False, solver=fake)
        hist = D.allreduce_sum(np.full(5, float(rank + 1)))
        q.put((rank, full, lo, hi, local, hist))
    finally:
        dist.destroy_process_group()


def test_sharded_solve_matches_single_rank():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k = np.random.default_rng(0).random((37, 3))
    ref = np.stack([k.sum(axis=1), k[:, 0] * 2.0, k[:, 1] - k[:, 2]])
    spans = []
    for rank, full, lo, hi, local, hist in got:
        assert np.array_equal(full, ref)                       # multi-rank result == single-rank result, every rank
        assert np.array_equal(local, ref[:, lo:hi])
        assert np.array_equal(hist, np.full(5, 3.0))
        spans.append((lo, hi))
    assert sorted(spans) == [(0, 19), (19, 37)]                # contiguous, sizes differ by at most one


def test_shard_range_covers_everything():
    from pysktb_b200.greens import shard_range
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
