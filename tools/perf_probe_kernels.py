"""Development aid: device times of the secondary kernels (reset, greedy evaluation, un-fused env step, agent ops)."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine, greedy_policy

def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best * 1e-3

P, n_p = 888, 5120
eng = Engine(P, n_p, threads_per_block=128, seeds=list(range(P)), tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12))
n = eng.n_total
out = {}
s = timed(lambda: eng.reset(0))
out["reset_kernel"] = dict(envs=n, us=round(s * 1e6, 1), write_gb_per_s=round(n * 48 / s / 1e9, 1))
eng.train(8)
act, st = eng.agent_select(0, 8)
s = timed(lambda: eng.agent_select(0, 8))
out["agent_select_kernel"] = dict(envs=n, us=round(s * 1e6, 1), env_per_s=f"{n / s:.3e}")
s = timed(lambda: eng.env_step(0, 9, act, auto_reset=True))
out["env_step_kernel"] = dict(envs=n, us=round(s * 1e6, 1), env_steps_per_s=f"{n / s:.3e}", state_gb_per_s=round(n * 96 / s / 1e9, 1))
o = eng.env_step(0, 10, act, auto_reset=True)
s = timed(lambda: eng.agent_update(st, act, o["next_state"], o["reward"]))
out["agent_update_kernel"] = dict(envs=n, populations=P, us=round(s * 1e6, 1), updates_per_s=f"{n / s:.3e}")
print(json.dumps(out))
s = timed(lambda: eng.env_step(0, 11, act, auto_reset=False))
print(json.dumps(dict(env_step_no_auto_reset_us=round(s * 1e6, 1), env_steps_per_s=f"{n / s:.3e}")))
