"""Quick device timing sweep (development aid, not the bench): env-steps/s of train_kernel for a few layouts."""
import sys, pathlib, time, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

def run(P, n_p, tpb, k, reps=3):
    eng = Engine(P, n_p, threads_per_block=tpb, seeds=list(range(P)), tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12))
    eng.reset(0)
    eng.train(k); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.train(k); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    sps = P * n_p * k / (best * 1e-3)
    print(json.dumps(dict(P=P, n_p=n_p, tpb=tpb, k=k, ms=round(best, 3), env_steps_per_s=f"{sps:.3e}")), flush=True)
    eng.close()

if __name__ == "__main__":
    for (P, n_p, tpb, k) in [(740, 1434, 128, 32), (888, 1195, 128, 32), (888, 1184, 128, 32), (1036, 1024, 128, 32), (444, 2390, 256, 32), (1776, 598, 64, 32),
                             (740, 1434, 128, 1), (888, 1195, 128, 1)]:
        run(P, n_p, tpb, k)
