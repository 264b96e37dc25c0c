"""Learning curves in replica-merge mode (development aid)."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

def run(R, n_r, merge_every, total_steps, chunk):
    tp = K.TrainerParameters(success_rate=2.0, max_num_episodes=10**15)
    eng = Engine(R, n_r, threads_per_block=128, seeds=[42] * R, population_ids=list(range(R)), replicas_per_population=R, tp=tp)
    eng.reset(0)
    done, rows = 0, []
    prev_ep = prev_su = 0
    while done < total_steps:
        eng.train_merged(chunk, merge_every); done += chunk
        ps = eng.population_state()
        ep, su = int(ps["total_episodes"].sum()), int(ps["total_successes"].sum())
        rows.append((done, ep // (R * n_r), round((su - prev_su) / max(ep - prev_ep, 1), 3)))
        prev_ep, prev_su = ep, su
    print(json.dumps(dict(R=R, n_r=n_r, merge_every=merge_every, curve_steps_episodesPerEnv_rate=rows)), flush=True)

if __name__ == "__main__":
    for R, m in ((8, 8), (64, 8), (512, 8), (512, 1), (512, 64)):
        run(R, 128, m, 300000, 50000)
