"""Shared-table exchange through the C-ABI alone (dqlb200_shared_sync_nccl on a raw ncclComm_t) against the Python path
(SharedTableSync.sync = C pack + torch.distributed all-gather + C apply): same seeds, same steps, both must leave bit-identical
tables and trainer states on every rank.  Run with one process per GPU:
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/nccl_sync_check.py      (or plainly: 1 rank)
Prints one JSON line per rank; exit code 0 iff identical.  Used by tests/test_gpu_parity.py::test_shared_sync_nccl_*."""
import ctypes as C
import json
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine
from dql_multirotor_landing_b200.parallel import SharedTableSync


class NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


def raw_comm(rank: int, world: int):
    """A communicator of this process' own (ncclCommInitRank), the unique id distributed with torch.distributed when world > 1."""
    nccl = C.CDLL("libnccl.so.2")            # the copy PyTorch has already loaded, or the system's
    uid = NcclUniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    if world > 1:
        t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).clone()
        dist.broadcast(t, 0)
        C.memmove(C.byref(uid), bytes(t.tolist()), 128)
    comm = C.c_void_p()
    nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, NcclUniqueId, C.c_int]
    assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0
    return nccl, comm


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")      # only carries the unique id and the Python reference path's all-gather (via CPU copies)
    R, n_r, M, rounds = 4, 64, 5, 6
    tp = K.TrainerParameters(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=10 ** 9)

    def engine():
        e = Engine(R, n_r, device=local, threads_per_block=64, seeds=[7] * R, population_ids=[rank * R + p for p in range(R)],
                   replicas_per_population=R, tp=tp)
        e.reset(0)
        return e
    a, b = engine(), engine()
    sa, sb = SharedTableSync(a, pooled_promotion=True), SharedTableSync(b, pooled_promotion=True)
    nccl, comm = raw_comm(rank, world)
    for _ in range(rounds):
        a.train(M); b.train(M)
        # reference path: C pack / apply around an all-gather done by torch.distributed (gloo here: through host copies)
        if a.R > 1:
            a.lib.dqlb200_replica_merge(a.handle, a.merge_snapshot.data_ptr(), 0, a._stream())
        sa.pack()
        if world > 1:
            parts = [torch.zeros_like(sa.packed.cpu()) for _ in range(world)]
            dist.all_gather(parts, sa.packed.cpu())
            sa.gathered.copy_(torch.stack(parts).to(sa.gathered.device))
        else:
            sa.gathered[0].copy_(sa.packed)
        sa.apply(sa.gathered)
        # the C-ABI path: one call on the raw communicator
        sb.sync_nccl(comm.value, world)
        torch.cuda.synchronize()
    a.check_errors(); b.check_errors()
    same = bool(torch.equal(a.tables, b.tables) and torch.equal(a.pop_state, b.pop_state) and torch.equal(sa.snapshot, sb.snapshot))
    visited = int((a.tables[0, 2] != 0).sum())
    print(json.dumps({"rank": rank, "world": world, "identical": same, "cells_visited": visited,
                      "working_step": int(a.population_state()[0]["working_step"])}), flush=True)
    nccl.ncclCommDestroy.argtypes = [C.c_void_p]
    nccl.ncclCommDestroy(comm)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0 if same and visited > 0 else 1)


if __name__ == "__main__":
    main()
