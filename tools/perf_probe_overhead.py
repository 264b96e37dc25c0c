"""Development aid: where does the fixed per-launch time of train_kernel go?  (event floor, prologue/epilogue only, 1 slot)"""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

def timed(fn, flush, reps=7):
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return round(best * 1e3, 1)

P, n_p, tpb = 888, 1280, 128
eng = Engine(P, n_p, threads_per_block=tpb, seeds=list(range(P)), tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12))
eng.reset(0); eng.train(300); torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tiny = torch.zeros(1024, device="cuda")
out = {"event_floor_tiny_kernel_us": timed(lambda: tiny.add_(1.0), flush), "train_k1_us": timed(lambda: eng.train(1), flush)}
ps = eng.pop_state.view(P, -1)
saved = ps.clone()
fin = np.zeros(1, K.POPULATION_STATE_DTYPE); fin["finished"] = 1
off = K.POPULATION_STATE_DTYPE.fields["finished"][1]
ps[:, off:off + 4] = torch.tensor(np.frombuffer(np.int32(1).tobytes(), np.uint8).copy(), device="cuda")
out["train_k1_finished_populations_us"] = timed(lambda: eng.train(1), flush)
ps.copy_(saved)
out["train_k1_again_us"] = timed(lambda: eng.train(1), flush)
import ctypes as C
lib, hnd = eng.lib, eng.handle
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
out["empty_kernel_888x128_35KB_smem_us"] = timed(lambda: lib.dqlb200_bench_launch_floor(hnd, 888, 128, 36000, st()), flush)
out["empty_kernel_888x128_no_smem_us"] = timed(lambda: lib.dqlb200_bench_launch_floor(hnd, 888, 128, 0, st()), flush)
out["empty_kernel_148x128_no_smem_us"] = timed(lambda: lib.dqlb200_bench_launch_floor(hnd, 148, 128, 0, st()), flush)
print(json.dumps(out))
