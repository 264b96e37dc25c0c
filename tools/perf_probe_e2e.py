"""Development aid: end-to-end (host buffers) throughput of dqlb200_train_host."""
import sys, pathlib, json, time
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine
P, n_p, k = 888, 1280, 64
eng = Engine(P, n_p, threads_per_block=128, seeds=list(range(P)), tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12))
eng.reset(0); eng.train(300); torch.cuda.synchronize()
env_h = eng.env_state.cpu().pin_memory(); tab_h = eng.tables.cpu().pin_memory(); ps_h = eng.pop_state.cpu().pin_memory()
import os
def call_ms(kk, levels=2, n=10):
    for _ in range(2): eng.train_host(kk, env_h, tab_h, ps_h, table_levels=levels)
    t0 = time.perf_counter()
    for _ in range(n): eng.train_host(kk, env_h, tab_h, ps_h, table_levels=levels)
    return (time.perf_counter() - t0) / n
for chunks in os.environ.get("CHUNKS", "8").split(","):
    os.environ["DQLB200_HOST_CHUNKS"] = chunks
    row = {"chunks": int(chunks)}
    for kk in (0, 1, 16, 64):
        s = call_ms(kk)
        row[f"k{kk}_ms"] = round(s * 1e3, 3)
    row["env_steps_per_s_k64"] = f"{P * n_p * 64 / (row['k64_ms'] * 1e-3):.3e}"
    print(json.dumps(row), flush=True)
# raw components
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return round((time.perf_counter() - t0) / n * 1e3, 3)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): eng.env_state.copy_(env_h, non_blocking=True)
    with torch.cuda.stream(s2): env_h2.copy_(eng.env_state, non_blocking=True)
env_h2 = torch.empty_like(env_h).pin_memory()
print(json.dumps(dict(h2d_env_ms=t(lambda: eng.env_state.copy_(env_h, non_blocking=True)), d2h_env_ms=t(lambda: env_h2.copy_(eng.env_state, non_blocking=True)),
                      both_dirs_ms=t(both), train64_ms=t(lambda: eng.train(k)), env_mb=env_h.numel() * 4 / 1e6)))
