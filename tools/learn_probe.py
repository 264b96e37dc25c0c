"""Learning-curve sanity (SURVEY.md A.4 behavioural golden): success rate of the last 100 episodes vs episodes."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

def run(n_envs, n_sub, total_steps, chunk, P=8):
    tp = K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12)
    eng = Engine(P, n_envs, threads_per_block=32 if n_envs <= 32 else 128, seeds=list(range(42, 42 + P)), tp=tp, dp=K.DynamicsParameters(n_sub=n_sub))
    eng.reset(0)
    done, rows = 0, []
    prev_ep = prev_su = 0
    while done < total_steps:
        eng.train(chunk); done += chunk
        ps = eng.population_state()
        ep, su = int(ps["total_episodes"].sum()), int(ps["total_successes"].sum())
        rows.append((done, ep // P, round((su - prev_su) / max(ep - prev_ep, 1), 3), round(float(ps["window_sum"].mean()) / 100, 3)))
        prev_ep, prev_su = ep, su
    print(json.dumps(dict(n_envs=n_envs, n_sub=n_sub, curve_steps_episodesPerPop_rateInChunk_window=rows)))

if __name__ == "__main__":
    run(1, 1, 400000, 50000)
    run(1, 4, 400000, 50000)
    run(1024, 1, 40000, 5000)
