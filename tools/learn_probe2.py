"""Learning curves vs population size (development aid): does the batched trainer reach the promotion threshold?"""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

def run(n_envs, total_steps, chunk, P=8, **tpkw):
    tp = K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12, **tpkw)
    eng = Engine(P, n_envs, threads_per_block=32 if n_envs <= 32 else 128, seeds=list(range(42, 42 + P)), tp=tp)
    eng.reset(0)
    done, rows = 0, []
    prev_ep = prev_su = 0
    prev_hist = np.zeros(9)
    while done < total_steps:
        eng.train(chunk); done += chunk
        ps = eng.population_state()
        ep, su = int(ps["total_episodes"].sum()), int(ps["total_successes"].sum())
        hist = ps["termination_hist"].sum(axis=0).astype(float)
        dh = hist - prev_hist
        rows.append((done, ep // (P * n_envs), round((su - prev_su) / max(ep - prev_ep, 1), 3), [round(x, 3) for x in (dh / max(dh.sum(), 1))[[2, 4, 8]]]))
        prev_ep, prev_su, prev_hist = ep, su, hist
    print(json.dumps(dict(n_envs=n_envs, curve_steps_episodesPerEnv_rate_succ_flyx_timeout=rows)), flush=True)

if __name__ == "__main__":
    run(1, 600000, 100000)
    run(64, 300000, 50000)
    run(1024, 300000, 50000)
