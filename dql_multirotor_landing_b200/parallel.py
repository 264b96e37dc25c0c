"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The path shards by POPULATION: every rank owns a disjoint set of agents (Q-table pairs) with their envs, so the
data path needs no collective (bench.py, `"scaling": "weak"`).  The only exchange step is the optional
shared-table mode: one agent replicated on G ranks, merged every `sync_every` global steps by ONE collective: an all-reduce
of the ranks' Q-deltas and counts realised as all-gather of the raw [Q_a bits | count | trainer counters] words (22.7 KB per
agent and rank, NCCL over NVLink on GPUs, gloo in the CPU tests) + a reduction in RANK ORDER inside the apply kernel.  Counts
and counters are integers end to end (exact for any number of visits per sync); the float32 sum has one defined order, so the
merged tables are bit-identical on every rank, from run to run and for any NCCL algorithm -- which an fp32 SUM all-reduce is
not.  See csrc/table_kernels.cuh: shared_pack_kernel / shared_apply_kernel.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def partition_populations(total: int, world_size: int, rank: int) -> range:
    """Contiguous block partition of population ids 0..total-1; sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(total, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def sweep_axes(population_ids: Sequence[int], seeds_per_point: int, speeds: Sequence[float], n_alpha: int
               ) -> Tuple[List[int], List[float], List[int]]:
    """seed x platform-speed x learning-rate sweep (BASELINE config 5): global population id -> (seed, v_mp, alpha variant).
    The mapping depends only on the GLOBAL id, so a population's trajectory is independent of how many GPUs run."""
    seeds, v_mp, alpha = [], [], []
    for g in population_ids:
        point, seed = divmod(g, seeds_per_point)
        seeds.append(seed)
        v_mp.append(speeds[point % len(speeds)])
        alpha.append((point // len(speeds)) % n_alpha)
    return seeds, v_mp, alpha


last_bind_note = "not called"      # why the last bind_to_gpu_numa_node call returned what it did (reported by bench.py)


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that pinned host buffers allocated
    afterwards are first-touched on that node and the host<->device copies of the ranks do not cross sockets.  Returns the
    node id, or None when nothing was changed; `last_bind_note` says why (no sysfs entry, a single-node host, ...)."""
    import os
    global last_bind_note
    try:
        import ctypes
        rt = ctypes.CDLL("libcudart.so.12")
        buf = ctypes.create_string_buffer(32)
        rc = rt.cudaDeviceGetPCIBusId(buf, 32, device_index)
        if rc != 0:
            last_bind_note = f"cudaDeviceGetPCIBusId failed ({rc})"
            return None
        bus = buf.value.decode().lower()
        node_file = f"/sys/bus/pci/devices/{bus}/numa_node"
        if not os.path.exists(node_file):
            last_bind_note = f"{node_file} does not exist (virtualised PCI topology)"
            return None
        node = int(open(node_file).read().strip())
        if node < 0:
            last_bind_note = f"{node_file} = {node}: the platform reports no NUMA affinity for the GPU"
            return None
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = set(cpus) & os.sched_getaffinity(0)
        if not allowed:
            last_bind_note = f"node {node} has no CPU this process may run on"
            return None
        os.sched_setaffinity(0, allowed)
        last_bind_note = f"bound to node {node} ({len(allowed)} CPUs; the host has {len(nodes)} NUMA node(s))"
        return node
    except Exception as exc:      # never fatal: the binding is an optimisation
        last_bind_note = f"{type(exc).__name__}: {exc}"
        return None


def max_over_ranks(values: Sequence[float], device: Optional[torch.device] = None) -> List[float]:
    """Device-timed numbers are reported as the maximum over ranks."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def gather_packed(packed: torch.Tensor, gathered: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of every rank's packed int32 buffer [agents, SHARED_WORDS] into gathered [world, agents, SHARED_WORDS]
    (rank-major, the layout shared_apply_kernel reduces in rank order) over `group` (default: all ranks); returns gathered."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_gather_into_tensor(gathered.view(-1), packed.view(-1), group=group)
    else:
        gathered[0].copy_(packed)
    return gathered


class SharedTableSync:
    """Shared-table mode: the agents of `engine` (groups of engine.R consecutive populations) are replicated on every rank.
    `sync()` = local replica merge (R > 1) + pack + ONE all-gather (22.7 KB per agent and rank) + apply (the rank-ordered
    reduction).  With `pooled_promotion` the curriculum promotion is decided from the windows of ALL ranks.  `group`: the ranks
    that share the agents (default: every rank; BASELINE config 4 on 4+ GPUs uses one group per axis)."""

    def __init__(self, engine, pooled_promotion: bool = False, group=None):
        import ctypes as C
        from . import _ffi
        from . import constants as K
        self._C, self._ffi, self.engine = C, _ffi, engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        self.snapshot = engine.tables[:: engine.R].clone().contiguous()
        n_agents = self.snapshot.shape[0]
        self.packed = torch.zeros((n_agents, K.SHARED_WORDS), dtype=torch.int32, device=engine.device)
        self.gathered = torch.zeros((self.world, n_agents, K.SHARED_WORDS), dtype=torch.int32, device=engine.device)
        self.pooled_promote = 0
        if pooled_promotion:
            self.pooled_promote = K.promote_threshold(engine.tp.successive_successful_episodes * engine.R * self.world, engine.tp.success_rate)
        if engine.R > 1:
            engine._ensure_merge_snapshot()

    def pack(self):
        e = self.engine
        self._ffi.check(e.lib.dqlb200_shared_pack(e.handle, self._C.c_void_p(self.packed.data_ptr()), e._stream()))

    def apply(self, gathered: torch.Tensor):
        e, C = self.engine, self._C
        self._ffi.check(e.lib.dqlb200_shared_apply(e.handle, C.c_void_p(self.snapshot.data_ptr()), C.c_void_p(gathered.data_ptr()), int(gathered.shape[0]),
                                                   self.pooled_promote, e._stream()))

    def sync_nccl(self, nccl_comm: int, n_ranks: int):
        """The same exchange as ONE C call on a raw ncclComm_t (address as int) -- what a non-Python caller uses
        (dqlb200_shared_sync_nccl: replica merge -> pack -> ncclAllGather -> apply); `gathered` must hold n_ranks entries."""
        e, C = self.engine, self._C
        if self.gathered.shape[0] != n_ranks:
            self.gathered = torch.zeros((n_ranks,) + tuple(self.packed.shape), dtype=torch.int32, device=e.device)
        self._ffi.check(e.lib.dqlb200_shared_sync_nccl(e.handle, C.c_void_p(self.snapshot.data_ptr()), C.c_void_p(self.packed.data_ptr()),
                                                       C.c_void_p(self.gathered.data_ptr()), C.c_void_p(nccl_comm), int(n_ranks),
                                                       int(e.pooled_promote), int(self.pooled_promote), e._stream()))

    def sync(self):
        """tables <- snapshot + visit-weighted mean of every rank's dQ_a (rank order); counts <- snapshot + sum of dcounts."""
        e = self.engine
        if e.R > 1:      # local copies first; with pooled promotion no rank decides alone
            self._ffi.check(e.lib.dqlb200_replica_merge(e.handle, e.merge_snapshot.data_ptr(), 0 if self.pooled_promote else e.pooled_promote, e._stream()))
        self.pack()
        self.apply(gather_packed(self.packed, self.gathered, self.group))
