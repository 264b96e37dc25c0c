"""Warm timing of the replica merge alone and of the train launch alone (config 3 layout: 512 replicas x 128 envs)."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine
out = {}
for R, n_r in ((512, 128), (128, 512), (64, 1024)):
    eng = Engine(R, n_r, threads_per_block=128, seeds=[42] * R, population_ids=list(range(R)), replicas_per_population=R,
                 tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10 ** 12))
    eng.reset(0); eng.train_merged(64, 1); torch.cuda.synchronize()
    def timed(fn, reps=50):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return round(e0.elapsed_time(e1) * 1e3 / reps, 2)
    out[f"R{R}x{n_r}"] = dict(merge_only_us=timed(eng.replica_merge), train_k1_only_us=timed(lambda: eng.train(1)),
                              pair_python_us=timed(lambda: (eng.train(1), eng.replica_merge())), pair_graph_us=timed(lambda: eng.train_merged(16, 1), 10) / 16)
    eng.close()
print(json.dumps(out, indent=1))
