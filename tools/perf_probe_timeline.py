"""Timeline of ONE K = 1 launch of train_kernel (development aid): %globaltimer stamps of every CTA at a few points, from a probe
build of the library (-DDQL_TIMING).  Prints, relative to the first CTA's entry, the median / max over CTAs of every stamp.
    DQL_BUILD_OUT=build_variants/libtiming.so DQL_NVCC_EXTRA=-DDQL_TIMING python -m dql_multirotor_landing_b200.build
    DQLB200_LIB=build_variants/libtiming.so python tools/perf_probe_timeline.py"""
import ctypes as C
import json
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

NAMES = ["entry", "tables staged (1st barrier)", "snapshot built (loop starts)", "end-of-step barrier passed", "loop left", "exit",
         "slot 0 done (warp 0)", "first tile landed"]
P, n_p = (int(sys.argv[2]) if len(sys.argv) > 2 else 888), 1280      # argv: [clean|dirty] [populations]
eng = Engine(P, n_p, threads_per_block=128, seeds=list(range(P)), tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10 ** 12))
eng.reset(0); eng.train(600); torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
buf = np.zeros(4096 * 8, np.uint64)
rows = []
clean = len(sys.argv) > 1 and sys.argv[1] == "clean"
flush2 = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if clean else None
for rep in range(5):
    flush.zero_()
    if clean:
        _ = int(flush2.view(torch.int64).sum())       # read 256 MiB: the L2 ends up holding CLEAN lines
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.train(1); e1.record(); torch.cuda.synchronize()
    eng.lib.dqlb200_debug_timing(buf.ctypes.data_as(C.c_void_p), C.c_int(buf.size))
    t = buf.reshape(4096, 8)[:P].astype(np.int64)
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3      # microseconds
    rows.append({"event_us": round(e0.elapsed_time(e1) * 1e3, 1),
                 **{NAMES[i]: [round(float(np.median(rel[:, i])), 2), round(float(rel[:, i].max()), 2)] for i in (0, 7, 1, 2, 6, 3, 4, 5)}})
print(json.dumps(rows[-2:], indent=1))
# spread over CTAs of the last launch: percentiles of every stamp, and the duration of the 10 slots per CTA
pct = lambda x: [round(float(v), 2) for v in np.percentile(x, [0, 10, 50, 90, 99, 100])]
print("percentiles 0/10/50/90/99/100 over CTAs (us)")
for i in (1, 2, 7, 6, 3, 5):
    print(f"  {NAMES[i]:34s} {pct(rel[:, i])}")
print(f"  {'slots 1-9 (stamp 3 - stamp 6)':34s} {pct(rel[:, 3] - rel[:, 6])}")
print(f"  {'staging (stamp 1 - entry)':34s} {pct(rel[:, 1] - rel[:, 0])}")
loc = buf[::-1][:P]                     # written from the tail of the buffer: (smid << 32) | warp slot of warp 0
smid, wslot = (loc >> np.uint64(32)).astype(np.int64), (loc & np.uint64(0xFFFFFFFF)).astype(np.int64)
dur = rel[:, 5] - rel[:, 0]
rank_in_sm = np.zeros(P, np.int64)      # 0 = the CTA of its SM that exits first
for sm in np.unique(smid):
    idx = np.where(smid == sm)[0]
    rank_in_sm[idx[np.argsort(dur[idx])]] = np.arange(len(idx))
print("CTAs per SM:", np.bincount(np.bincount(smid)).tolist(), "(index = CTAs on an SM)")
for r in range(int(rank_in_sm.max()) + 1):
    m = rank_in_sm == r
    print(f"  exit rank {r} within its SM: median duration {np.median(dur[m]):6.2f} us, median blockIdx {int(np.median(np.where(m)[0])):4d}, median warp slot of warp 0 {int(np.median(wslot[m])):3d}")
print("  correlation(duration, blockIdx) =", round(float(np.corrcoef(dur, np.arange(P))[0, 1]), 3), " correlation(duration, warp slot) =", round(float(np.corrcoef(dur, wslot)[0, 1]), 3))
per_sm = np.array([dur[smid == sm].max() for sm in np.unique(smid)])
print("  slowest CTA per SM: percentiles", pct(per_sm), " fastest CTA per SM:", pct(np.array([dur[smid == sm].min() for sm in np.unique(smid)])))
order = np.argsort(rel[:, 5])
print("slowest CTAs (blockIdx):", order[-12:].tolist(), " fastest:", order[:12].tolist())

