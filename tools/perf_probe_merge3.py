import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine
R, n_r = 512, 128
eng = Engine(R, n_r, threads_per_block=128, seeds=[42] * R, population_ids=list(range(R)), replicas_per_population=R,
             tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10**12))
eng.reset(0); eng.train_merged(200, 4, graph=False); torch.cuda.synchronize()
for _ in range(6):
    eng.train(1); eng.replica_merge()
torch.cuda.synchronize()
