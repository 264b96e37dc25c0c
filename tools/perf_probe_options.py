"""Cost of the observation-realism / model-fidelity options in the generic instance of train_kernel:
env-steps/s of 32 fused global steps for one wave of populations x 1280 envs (888 populations = 6 CTAs/SM, 740 = 5 CTAs/SM for
the extended variant), default configuration (production instance) next to each option."""
import sys, pathlib, json
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import torch
from dql_multirotor_landing_b200 import constants as K
from dql_multirotor_landing_b200.engine import Engine

CASES = {
    "default (production instance)": {},
    "generic instance, no option (c_d = 0.21)": dict(c_d=0.21),
    "observation noise": dict(noise_pos_sd=0.25, noise_vel_sd=0.1),
    "n_sub = 4": dict(n_sub=4),
    "n_sub = 4 + kalman_reference": dict(n_sub=4, accel_mode="kalman_reference"),
    "n_sub = 4 + second_order": dict(n_sub=4, dynamics_model="second_order"),
    "n_sub = 4 + second_order + kalman + noise": dict(n_sub=4, dynamics_model="second_order", accel_mode="kalman", noise_pos_sd=0.25, noise_vel_sd=0.1),
}
n_p, k = 1280, 32
out = {}
for name, dp in CASES.items():
    # one wave of CTAs: 6 per SM for the production / generic variants, 5 per SM for the extended variant (102 registers, 41 KB)
    P = 148 * (5 if ("accel_mode" in dp or "dynamics_model" in dp) else 6)
    eng = Engine(P, n_p, threads_per_block=128, seeds=list(range(P)), dp=K.DynamicsParameters(**dp),
                 tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10 ** 12))
    eng.reset(0)
    eng.train(300); torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.train(k); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    eng.check_errors()
    out[name] = dict(populations=P, us_per_global_step=round(best * 1e3 / k, 1), env_steps_per_s=float(f"{P * n_p * k / (best * 1e-3):.3e}"),
                     default_instance=bool(eng.lib.dqlb200_uses_default_instance(eng.handle)))
    eng.close()
print(json.dumps(out, indent=1))
