"""N > 1 host logic on CPU: world_size 2, gloo backend (no GPU needed).  Covers the population partitioning that
bench.py and the multi-GPU trainer use, the max-over-ranks timing reduction and the shared-table all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dql_multirotor_landing_b200 import parallel


def test_partition_is_disjoint_and_complete():
    for total in (1, 7, 8, 740, 5920):
        for world in (1, 2, 3, 4, 8):
            parts = [parallel.partition_populations(total, world, r) for r in range(world)]
            ids = [i for p in parts for i in p]
            assert ids == list(range(total))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_sweep_axes_depend_only_on_global_id():
    speeds = [0.4, 0.8, 1.2, 1.6]
    full = parallel.sweep_axes(range(64), 4, speeds, 2)
    for world in (2, 4):
        got = ([], [], [])
        for r in range(world):
            s, v, a = parallel.sweep_axes(parallel.partition_populations(64, world, r), 4, speeds, 2)
            got[0].extend(s); got[1].extend(v); got[2].extend(a)
        assert got == full
    assert len({(s, v, a) for s, v, a in zip(*full)}) == 32 and full[0][:4] == [0, 1, 2, 3]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 1. timing reduction: every rank ends up with the maximum
        mx = parallel.max_over_ranks([1.0 + rank, 10.0 - rank])
        # 2. shared-table merge: rank r visited cell c dcount[r][c] times and moved it by dq[r][c]
        rng = np.random.default_rng(100 + rank)
        cells = 2835
        dcount = rng.integers(0, 4, size=cells).astype(np.float32)
        dq = rng.standard_normal(cells).astype(np.float32) * (dcount > 0)
        delta = torch.zeros((1, 3, cells), dtype=torch.float32)
        delta[0, 0] = torch.from_numpy(dq * dcount)
        delta[0, 1] = torch.from_numpy(dcount)
        parallel.merge_deltas(delta)
        out[rank] = dict(mx=mx, delta=delta.numpy().copy(), dq=dq, dcount=dcount)
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["mx"] == r1["mx"] == [2.0, 10.0]
    assert np.array_equal(r0["delta"], r1["delta"])
    assert np.array_equal(r0["delta"][0, 1], r0["dcount"] + r1["dcount"])
    np.testing.assert_allclose(r0["delta"][0, 0], r0["dq"] * r0["dcount"] + r1["dq"] * r1["dcount"], rtol=1e-6, atol=1e-6)
    # the merged table value: visit-weighted mean of the replicas' deltas where anybody visited
    tot = r0["delta"][0, 1]
    merged = np.where(tot > 0, r0["delta"][0, 0] / np.maximum(tot, 1), 0.0)
    only0 = (r0["dcount"] > 0) & (r1["dcount"] == 0)
    np.testing.assert_allclose(merged[only0], r0["dq"][only0], rtol=1e-6)
