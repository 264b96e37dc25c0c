"""Starts a training session -- the counterpart of the reference's scripts/training.py (training.sh / training_x.sh /
training_y.sh -> launch/training.launch): no ROS node, no Gazebo; the batched Trainer drives the CUDA path.

    python scripts/training.py                       # x axis, 4 096 envs on one table pair
    python scripts/training.py --direction y --num-envs 65536 --success-rate 0.8
"""
import argparse
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))

from dql_multirotor_landing_b200 import constants as K          # noqa: E402
from dql_multirotor_landing_b200.trainer import Trainer         # noqa: E402

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--direction", choices=["x", "y"], default="x", help="launch/training.launch direction:=x|y")
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--success-rate", type=float, default=0.96, help="promotion threshold (PKG/trainer.py:25)")
    ap.add_argument("--max-global-steps", type=int, default=None)
    ap.add_argument("--save-path", default=None)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--noise", action="store_true", help="manager_node's default observation noise (0.25 m, 0.1 m/s)")
    ap.add_argument("--kalman", action="store_true", help="the observation node's Kalman-filtered acceleration estimate")
    ap.add_argument("--second-order", action="store_true", help="second-order attitude + vertical PID dynamics")
    a = ap.parse_args()
    dp = K.DynamicsParameters(n_sub=4 if (a.kalman or a.second_order) else 1,
                              noise_pos_sd=0.25 if a.noise else 0.0, noise_vel_sd=0.1 if a.noise else 0.0,
                              accel_mode="kalman_reference" if a.kalman else "exact",
                              dynamics_model="second_order" if a.second_order else "first_order")
    trainer = Trainer(seed=a.seed, success_rate=a.success_rate, save_path=a.save_path, num_envs=a.num_envs, device=a.device,
                      direction=a.direction, max_global_steps=a.max_global_steps, dynamics=dp)
    info = trainer.curriculum_training()
    trainer.log(info, clean=False)
