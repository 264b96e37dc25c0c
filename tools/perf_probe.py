"""Quick device timing sweep (development aid, not the bench): env-steps/s of train_kernel for a few layouts.
usage: perf_probe.py P,n_p,tpb,k [P,n_p,tpb,k ...]"""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

def run(P, n_p, tpb, k, mode=1, reps=5):
    """mode 0: no L2 flush, 1: dirty flush (256 MiB write), 2: clean flush (write, then a 256 MiB read)"""
    eng = Engine(P, n_p, threads_per_block=tpb, seeds=list(range(P)), tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12))
    eng.reset(0)
    eng.train(600); torch.cuda.synchronize()          # past the first episodes: counts grow, episodes desynchronise
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    src = torch.ones(64 << 20, dtype=torch.float32, device="cuda")
    best = 1e9
    for _ in range(reps):
        if mode >= 1:
            flush.zero_()
        if mode == 2:
            src.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.train(k); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    sps = P * n_p * k / (best * 1e-3)
    print(json.dumps(dict(P=P, n_p=n_p, tpb=tpb, k=k, mode=mode, us=round(best * 1e3, 1), env_steps_per_s=f"{sps:.3e}")), flush=True)
    eng.close()

if __name__ == "__main__":
    for a in sys.argv[1:]:
        run(*[int(x) for x in a.split(",")])
