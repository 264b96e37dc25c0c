"""Development aid: success-rate plateau at curriculum step 0 vs platform speed (replica-merge mode, merge every step)."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

def run(v_mp, R=128, n_r=128, M=1, total=250000, chunk=50000):
    tp = K.TrainerParameters(success_rate=2.0, max_num_episodes=10**15)
    eng = Engine(R, n_r, threads_per_block=128, seeds=[42] * R, population_ids=list(range(R)), replicas_per_population=R, v_mp=[v_mp] * R, tp=tp)
    eng.reset(0)
    done, rows, prev_ep, prev_su = 0, [], 0, 0
    while done < total:
        eng.train_merged(chunk, M); done += chunk
        ps = eng.population_state()
        ep, su = int(ps["total_episodes"].sum()), int(ps["total_successes"].sum())
        rows.append((ep // (R * n_r), round((su - prev_su) / max(ep - prev_ep, 1), 3)))
        prev_ep, prev_su = ep, su
    print(json.dumps(dict(v_mp=v_mp, R=R, M=M, curve_episodesPerEnv_rate=rows)), flush=True)

if __name__ == "__main__":
    for v in (0.0, 0.4, 0.8, 1.0, 1.2, 1.6):
        run(v)
