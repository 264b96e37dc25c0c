"""Development aid: split the cost of one (train K=1, merge) pair."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

def timed(fn, n):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) * 1e3 / n, 2)

for R, n_r, tpb in ((512, 128, 128), (1024, 64, 64)):
    eng = Engine(R, n_r, threads_per_block=tpb, seeds=[42] * R, population_ids=list(range(R)), replicas_per_population=R,
                 tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12))
    eng.reset(0); eng.train_merged(600, 4); torch.cuda.synchronize()
    out = dict(R=R, n_r=n_r)
    out["train_k1_us"] = timed(lambda: eng.train(1), 200)
    out["merge_only_us"] = timed(lambda: eng.replica_merge(), 200)
    out["train1_then_merge_us"] = timed(lambda: (eng.train(1), eng.replica_merge()), 200)
    out["graph_pair_us"] = timed(lambda: eng.train_merged(50, 1), 4) / 50
    print(json.dumps(out), flush=True)
    eng.close()
