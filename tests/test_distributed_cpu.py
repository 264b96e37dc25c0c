"""N > 1 host logic on CPU: world_size 2, gloo backend (no GPU needed).  Covers the population partitioning that
bench.py and the multi-GPU trainer use, the max-over-ranks timing reduction and the shared-table exchange (all-gather)."""
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
        # 2. shared-table exchange: every rank contributes its packed int32 words [agents, SHARED_WORDS]; all ranks end up
        #    with the same rank-major stack (the reduction in rank order happens in shared_apply_kernel on the GPU)
        from dql_multirotor_landing_b200 import constants as K
        rng = np.random.default_rng(100 + rank)
        packed = torch.from_numpy(rng.integers(-2 ** 31, 2 ** 31 - 1, size=(3, K.SHARED_WORDS), dtype=np.int64).astype(np.int32))
        gathered = torch.zeros((world, 3, K.SHARED_WORDS), dtype=torch.int32)
        parallel.gather_packed(packed, gathered)
        out[rank] = dict(mx=mx, packed=packed.numpy().copy(), gathered=gathered.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    r0, r1 = out[0], out[1]
    assert r0["mx"] == r1["mx"] == [2.0, 10.0]
    assert np.array_equal(r0["gathered"], r1["gathered"])                    # identical on every rank
    assert np.array_equal(r0["gathered"][0], r0["packed"]) and np.array_equal(r0["gathered"][1], r1["packed"])   # rank-major, bit for bit


def test_gather_packed_without_process_group():
    from dql_multirotor_landing_b200 import constants as K
    packed = torch.arange(2 * K.SHARED_WORDS, dtype=torch.int32).reshape(2, K.SHARED_WORDS)
    gathered = torch.zeros((1, 2, K.SHARED_WORDS), dtype=torch.int32)
    assert torch.equal(parallel.gather_packed(packed, gathered)[0], packed)
