"""Device timing of train_kernel (development aid, not the bench).  Three shapes, all at curriculum step 0 after 600 warm-up steps:
  k1      888 x 1280 envs, one global step per launch, L2 flushed between launches (the bench headline shape)
  k32     the same envs, 32 fused global steps per launch (state L2-resident)
  stream  888 x 5120 envs (218 MB of state > L2), 40 single-step launches back to back in one timed region, no flush
usage: [DQLB200_LIB=build_variants/libX.so] python tools/perf_probe.py [k1] [k32] [stream] [P,n_p,tpb,k,mode ...]"""
import json
import os
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

TAG = os.path.basename(os.environ.get("DQLB200_LIB", "libdqlb200.so"))


def engine(P, n_p, tpb):
    eng = Engine(P, n_p, threads_per_block=tpb, seeds=list(range(P)), tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12))
    eng.reset(0)
    eng.train(600)
    torch.cuda.synchronize()          # past the first episodes: counts grow, episodes desynchronise
    return eng


def run(P, n_p, tpb, k, mode=1, reps=7, name=None):
    """mode 0: no L2 flush, 1: dirty flush (256 MiB write) before every timed launch"""
    eng = engine(P, n_p, tpb)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    times = []
    for _ in range(reps):
        if mode >= 1:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.train(k); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    times.sort()
    med = times[len(times) // 2]
    print(json.dumps(dict(lib=TAG, shape=name or f"{P}x{n_p}", tpb=tpb, k=k, flush=mode, us_min=round(times[0] * 1e3, 1), us_med=round(med * 1e3, 1),
                          env_steps_per_s=f"{P * n_p * k / (med * 1e-3):.3e}", hbm_frac=round(96 * P * n_p / (med * 1e-3 / k) / 6.544e12, 3))), flush=True)
    eng.check_errors()
    eng.close()


def stream(P=888, n_p=5120, tpb=128, launches=40, reps=3):
    eng = engine(P, n_p, tpb)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(launches):
            eng.train(1)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / launches)
    print(json.dumps(dict(lib=TAG, shape=f"stream {P}x{n_p}", tpb=tpb, k=1, us_per_launch=round(best * 1e3, 1),
                          env_steps_per_s=f"{P * n_p / (best * 1e-3):.3e}", hbm_frac=round(96 * P * n_p / (best * 1e-3) / 6.544e12, 3))), flush=True)
    eng.check_errors()
    eng.close()


if __name__ == "__main__":
    for a in sys.argv[1:] or ["k1", "k32", "stream"]:
        if a == "k1":
            run(888, 1280, 128, 1, 1, reps=15, name="k1")
        elif a == "k32":
            run(888, 1280, 128, 32, 0, name="k32")
        elif a == "stream":
            stream()
        else:
            run(*[int(x) for x in a.split(",")])
