"""Drop-in mirrors of the reference's gym environments (PKG/landing_simulation_env.py:30-400), batched on the device.

``TrainingLandingEnv`` / ``SimulationLandingEnv`` keep the reference's constructor signatures and the gym calling
convention (``reset() -> state``, ``step(action) -> (state, reward, done, info)``; for the simulation env
``reset() -> (state_x, state_y)``, ``step(ax, ay) -> (state_x, state_y, done, info)``), and add ``num_envs``: with the default
``num_envs=1`` states are 5-tuples, rewards floats and ``info`` holds the reference's keys, exactly like the reference; with
``num_envs > 1`` every quantity is a NumPy array over the environments.  Gazebo/ROS is replaced by the analytic stand-in
inside ``dqlb200_env_reset`` / ``dqlb200_env_step`` (csrc/env_kernels.cuh: env_reset_kernel, env_step_kernel) -- the same device
functions, in the same order, as the fused training kernel.  ``make("Landing-Training-v0", ...)`` /
``make("Landing-Simulation-v0", ...)`` stand in for ``gym.make`` (the ids are registered at PKG/landing_simulation_env.py:432-440).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from . import _ffi
from . import constants as K
from .mdp import state_tuple


class AbstractLandingEnv:
    """PKG/landing_simulation_env.py:30-140 (services, publishers and subscribers have no counterpart: no simulator)."""

    _simulation = False

    def __init__(self, t_max: int = 20, initial_curriculum_step: int = 0, f_ag: float = 22.92, p_max: float = 4.5, z_init: float = 2.0,
                 *, num_envs: int = 1, seed: int = 42, device: int = 0, platform_speed: float = 1.6, direction: str = "x",
                 dynamics: Optional[K.DynamicsParameters] = None, auto_reset: bool = False):
        from .engine import Engine
        self._working_curriculum_step = initial_curriculum_step
        self._f_ag, self._t_max, self._p_max, self._z_init = f_ag, t_max, p_max, z_init
        self._flyzone_x = self._flyzone_y = (-p_max, p_max)          # PKG/landing_simulation_env.py:82-87
        self._flyzone_z = (0, p_max)
        self.num_envs, self._auto_reset = num_envs, auto_reset
        dp = dynamics or K.DynamicsParameters(z_init=z_init, v_mp=platform_speed)
        self._engine = Engine(1, num_envs, device=device, threads_per_block=32 if num_envs <= 32 else 128, seeds=[seed],
                              v_mp=[dp.v_mp], axes=[direction], mp=K.MdpParameters(f_ag=f_ag, t_max=t_max, p_max=p_max), dp=dp)
        dev = self._engine.device
        n = num_envs
        self._t = 0                     # step counter = Philox counter word 1 (births of resets, like the fused kernel's t)
        self._fresh = True              # a new env object comes with a new MDP object (PKG/landing_simulation_env.py:157-164)
        self._actions = torch.zeros(n, dtype=torch.int8, device=dev)
        self._state = torch.zeros(n, dtype=torch.int16, device=dev)
        self._reward = torch.zeros(n, dtype=torch.float64, device=dev)
        self._code = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._done = torch.zeros(n, dtype=torch.uint8, device=dev)
        self._obs = torch.zeros((n, 5), dtype=torch.float32, device=dev)
        self._steps = torch.zeros(n, dtype=torch.int32, device=dev)
        self._cum = torch.zeros(n, dtype=torch.float64, device=dev)

    # ------------------------------------------------------------------------------------------------
    def _states(self):
        ids = self._state.cpu().numpy().astype(np.int64)
        if self.num_envs == 1:
            return state_tuple(int(ids[0]))
        t = ids % 7; r = ids // 7
        a = r % 3; r //= 3
        v = r % 3; r //= 3
        p = r % 3; r //= 3
        return np.stack([r, p, v, a, t], axis=1)

    def _reset(self, mask=None):
        e = self._engine
        m = None
        if mask is not None:
            m = torch.as_tensor(np.ascontiguousarray(mask, dtype=np.uint8), device=e.device)
        _ffi.check(e.lib.dqlb200_env_reset(e.handle, self._working_curriculum_step, self._t, m.data_ptr() if m is not None else None,
                                           int(self._fresh), int(self._simulation), self._state.data_ptr(), e._stream()))
        self._fresh = False

    def _step(self, actions):
        e = self._engine
        a = np.broadcast_to(np.asarray(actions, dtype=np.int8), (self.num_envs,))
        if ((a < 0) | (a > 2)).any():
            raise ValueError("actions must be 0 (increase), 1 (decrease) or 2 (hold)")
        self._actions.copy_(torch.from_numpy(np.array(a, dtype=np.int8)))
        _ffi.check(e.lib.dqlb200_env_step(e.handle, self._working_curriculum_step, self._t, self._actions.data_ptr(), int(self._auto_reset),
                                          int(self._simulation), self._state.data_ptr(), self._reward.data_ptr(), self._code.data_ptr(),
                                          self._done.data_ptr(), self._obs.data_ptr(), self._steps.data_ptr(), self._cum.data_ptr(), None, e._stream()))
        e.check_errors()
        self._t += 1

    def _info(self, with_reward: bool) -> Dict[str, Any]:
        code, steps = self._code.cpu().numpy(), self._steps.cpu().numpy()
        if self.num_envs == 1:
            info: Dict[str, Any] = {}
            c = int(code[0])
            if c >= 2:           # PKG/mdp.py:427-439, 832-845
                info["Termination condition"] = K.TERMINATION_STRINGS[c]
                info["Number of steps"] = int(steps[0])
                if with_reward:
                    cum = float(self._cum.cpu().numpy()[0])
                    info["Cumulative reward"] = cum                      # without the last step's reward (quirk Q12)
                    info["Mean reward"] = cum / int(steps[0])
            if with_reward:
                info["Current reward"] = float(self._reward.cpu().numpy()[0])   # PKG/landing_simulation_env.py:274
            return info
        info = {"code": code.copy(), "Number of steps": steps.copy(),
                "Termination condition": np.asarray([K.TERMINATION_STRINGS.get(int(c), "") for c in code], dtype=object)}
        if with_reward:
            info["Cumulative reward"] = self._cum.cpu().numpy()
            info["Current reward"] = self._reward.cpu().numpy()
        return info

    @property
    def observation(self) -> np.ndarray:
        """Continuous observation of the last step, [num_envs, 5]: rel_p_x, rel_v_x, rel_a_x, pitch, z (the stand-in's
        counterpart of the /landing_simulation/observation message, PKG/landing_simulation_env.py:120-140)."""
        return self._obs.cpu().numpy()

    def close(self):
        self._engine.close()


class TrainingLandingEnv(AbstractLandingEnv):
    """PKG/landing_simulation_env.py:143-282."""

    def __init__(self, initial_curriculum_step: int = 0, *, t_max: int = 20, f_ag: float = 22.92, p_max: float = 4.5, z_init: float = 2.0, **extras):
        # z_init = 2.0 like the reference's AbstractLandingEnv / TrainingLandingEnv (PKG/landing_simulation_env.py:37,150); the Trainer
        # passes 4.0 explicitly (PKG/trainer.py:41,180), SimulationLandingEnv defaults to 4 (:293)
        super().__init__(t_max=t_max, initial_curriculum_step=initial_curriculum_step, f_ag=f_ag, p_max=p_max, z_init=z_init, **extras)

    def reset(self, mask=None):
        """New episode for every env (or those selected by `mask`): R1 initial-state law + one hover period + first
        discretisation (PKG/landing_simulation_env.py:167-243).  Returns the state(s)."""
        self._reset(mask)
        return self._states()

    def step(self, action_x, action_y=2):
        """PKG/landing_simulation_env.py:245-282.  action_y is accepted and ignored like in the reference's TrainingMdp."""
        self._step(action_x)
        done = self._done.cpu().numpy().astype(bool)
        reward = self._reward.cpu().numpy()
        if self.num_envs == 1:
            return self._states(), float(reward[0]), bool(done[0]), self._info(True)
        return self._states(), reward, done, self._info(True)


class SimulationLandingEnv(AbstractLandingEnv):
    """PKG/landing_simulation_env.py:285-400: greedy roll-outs, SimulationMdp semantics, x and y states."""

    _simulation = True

    def __init__(self, initial_curriculum_step: int = 4, *, t_max: int = 20, f_ag: float = 22.92, p_max: float = 4.5, z_init: float = 4, **extras):
        super().__init__(t_max=t_max, initial_curriculum_step=initial_curriculum_step, f_ag=f_ag, p_max=p_max, z_init=z_init, **extras)

    def _state_y(self):
        # the reference keeps the drone on the platform's y (PKG/landing_simulation_env.py:336-340) and never applies the y
        # action (PKG/mdp.py:863-876): the y observation is identically zero, its state is the all-centre state of level w
        w = self._working_curriculum_step
        s = (w, 1, 1, 1, 3)
        return s if self.num_envs == 1 else np.tile(np.asarray(s), (self.num_envs, 1))

    def reset(self, mask=None):
        self._reset(mask)
        return self._states(), self._state_y()

    def step(self, action_x, action_y=2):
        self._step(action_x)
        done = self._done.cpu().numpy().astype(bool)
        return self._states(), self._state_y(), (bool(done[0]) if self.num_envs == 1 else done), self._info(False)


def make(env_id: str, **kwargs):
    """gym.make for the two ids the reference registers (PKG/landing_simulation_env.py:432-440)."""
    if env_id == "Landing-Training-v0":
        return TrainingLandingEnv(**kwargs)
    if env_id == "Landing-Simulation-v0":
        return SimulationLandingEnv(**kwargs)
    raise ValueError(f"unknown environment id {env_id!r}")
