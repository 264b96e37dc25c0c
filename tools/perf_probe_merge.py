"""Replica-merge timing probe (development aid; BASELINE configs 3 / 4: one agent spread over R replicas, merged every M steps).

    python tools/perf_probe_merge.py                 throughput of (train, merge) graphs for the shapes DESIGN.md section 3 quotes
    python tools/perf_probe_merge.py split           cost of one pair split up: train launch alone, merge alone, the pair from Python, the pair in a graph
    python tools/perf_probe_merge.py total_envs,R,tpb,M[,steps]
"""
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine


def make(R, n_r, tpb):
    eng = Engine(R, n_r, threads_per_block=tpb, seeds=[42] * R, population_ids=list(range(R)), replicas_per_population=R,
                 tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10 ** 12))
    eng.reset(0)
    return eng


def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps      # microseconds per call


def throughput(total_envs, R, tpb, M, steps=256):
    eng = make(R, total_envs // R, tpb)
    eng.train_merged(2 * M, M); torch.cuda.synchronize()
    us = min(timed(lambda: eng.train_merged(steps, M), 1) for _ in range(3))
    print(json.dumps(dict(total_envs=total_envs, R=R, n_r=total_envs // R, tpb=tpb, merge_every=M, us_per_step=round(us / steps, 2),
                          env_steps_per_s=f"{total_envs * steps / (us * 1e-6):.3e}")), flush=True)
    eng.close()


def split():
    out = {}
    for R, n_r, tpb in ((512, 128, 128), (1024, 64, 64), (128, 512, 128), (64, 1024, 128)):
        eng = make(R, n_r, tpb)
        eng.train_merged(600, 4); torch.cuda.synchronize()
        out[f"R{R}x{n_r}"] = dict(train_k1_only_us=round(timed(lambda: eng.train(1), 200), 2), merge_only_us=round(timed(eng.replica_merge, 200), 2),
                                  pair_python_us=round(timed(lambda: (eng.train(1), eng.replica_merge()), 200), 2),
                                  pair_graph_us=round(timed(lambda: eng.train_merged(64, 1), 4) / 64, 2))
        eng.close()
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    args = sys.argv[1:]
    if not args:
        for a in [(65536, 512, 128, 1), (65536, 512, 128, 4), (65536, 512, 128, 16), (65536, 1024, 64, 1), (65536, 2048, 32, 1),
                  (262144, 1024, 128, 1), (262144, 1024, 128, 16)]:
            throughput(*a)
    for a in args:
        if a == "split":
            split()
        else:
            throughput(*[int(x) for x in a.split(",")])
