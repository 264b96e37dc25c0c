"""Config 3/4 probe: one agent, 65,536 (or 262,144) envs spread over R replicas, merged every M global steps."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

def run(total_envs, R, tpb, M, steps=256):
    n_r = total_envs // R
    eng = Engine(R, n_r, threads_per_block=tpb, seeds=[42] * R, population_ids=list(range(R)), replicas_per_population=R,
                 tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12))
    eng.reset(0)
    eng.train_merged(2 * M, M); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.train_merged(steps, M); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(json.dumps(dict(total_envs=total_envs, R=R, n_r=n_r, tpb=tpb, merge_every=M, us_per_step=round(best * 1e3 / steps, 2),
                          env_steps_per_s=f"{total_envs * steps / (best * 1e-3):.3e}")), flush=True)
    eng.close()

if __name__ == "__main__":
    for args in [(65536, 512, 128, 1), (65536, 512, 128, 4), (65536, 512, 128, 16), (65536, 1024, 64, 1), (65536, 2048, 32, 1),
                 (262144, 1024, 128, 1), (262144, 1024, 128, 16)]:
        run(*args)
