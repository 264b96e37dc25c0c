"""Drop-in mirrors of the reference's MDP classes (PKG/mdp.py) backed by the CUDA facade kernel.

Same class names, constructor signatures, method names, return types and ValueErrors as
PKG/mdp.py:11-886; the numerics run on the GPU in float64 through ``dqlb200_mdp_facade_step``
(csrc/facade_kernels.cuh: facade_kernel) -- there is no CPU implementation here.  These objects exist so that code
written against the reference (and parity tests that read like reference tests) keep working; the
throughput path is ``Trainer`` / ``Engine`` (one fused kernel for thousands of envs).
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from . import _ffi
from . import constants as K
from .msg import Action, Observation


class ContinuousObservation:
    """PKG/mdp.py:11-32."""

    def __init__(self, observation: Optional[Observation] = None, pitch: float = 0.0, roll: float = 0.0,
                 abs_p_z: float = 0.0, contact: bool = False) -> None:
        observation = observation if observation is not None else Observation()
        self.rel_p_x, self.rel_p_y, self.rel_p_z = observation.rel_p_x, observation.rel_p_y, observation.rel_p_z
        self.rel_v_x, self.rel_v_y, self.rel_v_z = observation.rel_v_x, observation.rel_v_y, observation.rel_v_z
        self.rel_a_x, self.rel_a_y, self.rel_a_z = observation.rel_a_x, observation.rel_a_y, observation.rel_a_z
        self.contact = observation.contact
        self.pitch, self.roll, self.abs_p_z = pitch, roll, abs_p_z


class CheckResult(enum.Enum):
    """PKG/mdp.py:68-77."""
    TERMINAL_CONTACT = "SUCCESS: Touched platform"
    TERMINAL_SUCCESS = "SUCCESS: Goal state reached"
    TERMINAL_FLYZONE_X = "FAILURE: Drone moved too far from platform in x direction"
    TERMINAL_FLYZONE_Y = "FAILURE: Drone moved too far from platform in y direction"
    TERMINAL_FLYZONE_Z = "FAILURE: Drone moved too far from platform in z direction"
    TERMINAL_MINIMUM_ALTITUDE = "FAILURE: Reached minimum altitude"
    TERMINAL_TIMEOUT = "FAILURE: Maximum episode duration"
    NON_TERMINAL_SUCCESS = "non-terminal success"
    NON_TERMINAL = "non-terminal"


_CODE_TO_RESULT = {
    0: CheckResult.NON_TERMINAL, 1: CheckResult.NON_TERMINAL_SUCCESS, 2: CheckResult.TERMINAL_SUCCESS,
    3: CheckResult.TERMINAL_CONTACT, 4: CheckResult.TERMINAL_FLYZONE_X, 5: CheckResult.TERMINAL_FLYZONE_Y,
    6: CheckResult.TERMINAL_FLYZONE_Z, 7: CheckResult.TERMINAL_MINIMUM_ALTITUDE, 8: CheckResult.TERMINAL_TIMEOUT,
}


def state_tuple(sid: int) -> Tuple[int, int, int, int, int]:
    t = sid % 7; sid //= 7
    a = sid % 3; sid //= 3
    v = sid % 3; sid //= 3
    p = sid % 3; sid //= 3
    return (sid, p, v, a, t)


def state_id(s) -> int:
    return (((s[0] * 3 + s[1]) * 3 + s[2]) * 3 + s[3]) * 7 + s[4]


class _FacadeHandle:
    """One libdqlb200 handle per distinct parameter set (the config carries the MDP constants)."""
    _cache: Dict[tuple, "_FacadeHandle"] = {}

    def __init__(self, mp: K.MdpParameters, device: int):
        if not torch.cuda.is_available():
            raise _ffi.Dqlb200Error("no CUDA device: dql_multirotor_landing_b200 has no CPU fallback")
        self.lib = _ffi.load()
        self.device = torch.device("cuda", device)
        self.cfg = K.build_config(1, 1, 32, mp)
        lut = K.alpha_lut()
        dphase, r, rw, rw2 = K.platform_constants(2.0, 1.6, mp.f_ag, 1)
        pp = (K.PopulationParams * 1)(K.PopulationParams(0, 0, 0, dphase, r, rw, rw2, 0, 9.81, 0))
        self.handle = C.c_void_p()
        _ffi.check(self.lib.dqlb200_create(C.byref(self.cfg), lut.ctypes.data_as(C.POINTER(C.c_float)), pp, device, C.byref(self.handle)))

    @classmethod
    def get(cls, mp: K.MdpParameters, device: int = 0) -> "_FacadeHandle":
        key = (tuple(sorted(vars(mp).items())), device)
        if key not in cls._cache:
            cls._cache[key] = cls(mp, device)
        return cls._cache[key]

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)


class _MdpRecord:
    """Device record of one MDP object: 12 doubles (layout in include/dqlb200.h: dqlb200_mdp_facade_step)."""

    def __init__(self, fh: _FacadeHandle, w: int, simulation: bool):
        self.fh, self.w, self.sim = fh, w, simulation
        dev = fh.device
        self.state = torch.zeros(12, dtype=torch.float64, device=dev)
        self.state[8] = -1.0
        self.state[9] = -1.0
        self.obs = torch.zeros(6, dtype=torch.float64, device=dev)
        self.contact = torch.zeros(1, dtype=torch.uint8, device=dev)
        self.action = torch.zeros(1, dtype=torch.int8, device=dev)
        self.out_state = torch.zeros(1, dtype=torch.int16, device=dev)
        self.out_code = torch.zeros(1, dtype=torch.uint8, device=dev)
        self.out_reward = torch.zeros(1, dtype=torch.float64, device=dev)

    def run(self, ops: int):
        if self.sim:
            ops |= _ffi.OP_SIMULATION
        fh = self.fh
        _ffi.check(fh.lib.dqlb200_mdp_facade_step(fh.handle, self.w, ops, 1, self.obs.data_ptr(), self.contact.data_ptr(),
                                                  self.action.data_ptr(), self.state.data_ptr(), self.out_state.data_ptr(),
                                                  self.out_code.data_ptr(), self.out_reward.data_ptr(), fh.stream()))
        _ffi.check(fh.lib.dqlb200_check_errors(fh.handle, fh.stream()))      # raises ValueError like PKG/mdp.py:170,353,442-452

    def set_obs(self, rel_p, rel_v, rel_a, angle, z, rel_p_other, contact):
        self.obs.copy_(torch.tensor([rel_p, rel_v, rel_a, angle, z, rel_p_other], dtype=torch.float64))
        self.contact.fill_(1 if contact else 0)

    def host(self) -> np.ndarray:
        return self.state.cpu().numpy()


class AbstractMdp:
    """Common constructor (PKG/mdp.py:87-147)."""

    def __init__(self, working_curriculum_step: int, f_ag: float, t_max: int, p_max: float = 4.5, *, w_p: float = -100.0,
                 w_v: float = -10.0, w_theta: float = -1.55, w_dur: float = -6.0, w_fail: float = -2.6, w_succ: float = 2.6,
                 n_theta: int = 3, v_max: float = 3.39411, a_max: float = 1.28, theta_max: float = np.deg2rad(21.37723),
                 delta_theta: float = np.deg2rad(7.12574), beta: float = 1 / 3, sigma_a: float = 0.416,
                 minimum_altitude: float = 0.1, device: int = 0) -> None:
        if not (0 <= working_curriculum_step < K.MAX_CURRICULUM):
            raise ValueError("working_curriculum_step must be in 0..4")
        self._working_curriculum_step = working_curriculum_step
        self._f_ag, self._t_max, self._p_max = f_ag, t_max, p_max
        self._flyzone_x = (-p_max, p_max)
        self._flyzone_y = (-p_max, p_max)
        self._flyzone_z = (0.0, p_max)
        self._params = K.MdpParameters(f_ag=f_ag, t_max=t_max, p_max=p_max, w_p=w_p, w_v=w_v, w_theta=w_theta, w_dur=w_dur,
                                       w_fail=w_fail, w_succ=w_succ, n_theta=n_theta, v_max=v_max, a_max=a_max,
                                       theta_max=float(theta_max), delta_theta=float(delta_theta), beta=beta, sigma_a=sigma_a,
                                       minimum_altitude=minimum_altitude)
        self._theta_max, self._delta_theta = float(theta_max), float(delta_theta)
        self._discrete_angles = np.linspace(-theta_max, theta_max, (n_theta * 2) + 1)
        self._delta_t = 1 / f_ag
        self._fh = _FacadeHandle.get(self._params, device)
        self._info: Dict[str, Any] = {}

    @staticmethod
    def _termination_info(info: Dict[str, Any], code: int, steps: int):
        info["Termination condition"] = _CODE_TO_RESULT[code].value
        info["Number of steps"] = steps


class TrainingMdp(AbstractMdp):
    """PKG/mdp.py:206-569."""

    def __init__(self, working_curriculum_step: int, f_ag: float, t_max: int, p_max: float = 4.5, *, minimum_altitude: float = 0.2, **kw) -> None:
        super().__init__(working_curriculum_step, f_ag, t_max, p_max, minimum_altitude=minimum_altitude, **kw)
        self._rec = _MdpRecord(self._fh, working_curriculum_step, simulation=False)
        self._current_continuous_action = Action(pitch=0, roll=0, yaw=0, v_z=-0.1)
        self._current_discrete_state: Optional[Tuple[int, int, int, int, int]] = None
        self._previous_discrete_state: Optional[Tuple[int, int, int, int, int]] = None
        self._current_continuous_observation = ContinuousObservation()

    # -- reference API ----------------------------------------------------------------------------
    def reset(self):
        """PKG/mdp.py:562-569 (+194-200): the shaping potentials are NOT cleared (quirk Q11)."""
        self._rec.run(_ffi.OP_RESET)
        self._info = {}
        self._current_continuous_observation = ContinuousObservation()
        self._current_discrete_state = None
        self._previous_discrete_state = None
        self._current_continuous_action = Action(pitch=0, roll=0, yaw=0, v_z=-0.1)

    def discrete_state(self, current_continuous_observation: ContinuousObservation) -> Tuple[int, int, int, int, int]:
        """PKG/mdp.py:257-333."""
        o = current_continuous_observation
        self._current_continuous_observation = o
        self._rec.set_obs(o.rel_p_x, o.rel_v_x, o.rel_a_x, o.pitch, o.abs_p_z, o.rel_p_y, o.contact)
        self._rec.run(_ffi.OP_OBSERVE)
        self._previous_discrete_state = self._current_discrete_state
        self._current_discrete_state = state_tuple(int(self._rec.out_state.item()))
        return self._current_discrete_state

    def continuous_action(self, action_x: int, action_y: int = 2):
        """PKG/mdp.py:543-560."""
        if action_y != 2:
            raise ValueError("Cannot move in the y direction while training")
        self._rec.action.fill_(int(action_x))
        self._rec.run(_ffi.OP_ACTION)
        self._current_continuous_action.pitch = float(self._rec.state[0].item())
        return self._current_continuous_action

    def check(self) -> Dict[str, Any]:
        """PKG/mdp.py:335-439."""
        if not self._current_discrete_state:
            raise ValueError("Cannot check an empty state\nYou must call `discrete_state` before calling check.")
        self._rec.run(_ffi.OP_CHECK)
        st = self._rec.host()
        code, steps = int(st[7]), int(st[5])
        self._check_result = _CODE_TO_RESULT[code]
        if code >= 2:
            self._termination_info(self._info, code, steps)
            self._info["Cumulative reward"] = float(st[4])           # before this step's reward (quirk Q12)
            self._info["Mean reward"] = float(st[4]) / steps
        return self._info

    def reward(self) -> float:
        """PKG/mdp.py:441-541."""
        if not self._previous_discrete_state:
            raise ValueError("Previous state missing.\nYou must call `reset` and `discrete_state`and then `step`before calling check.")
        self._rec.run(_ffi.OP_REWARD)
        return float(self._rec.out_reward.item())

    # -- observable internals some callers/tests read ---------------------------------------------
    @property
    def _step_count(self) -> int:
        return int(self._rec.host()[5])

    @property
    def _cumulative_reward(self) -> float:
        return float(self._rec.host()[4])

    @property
    def _curriculum_check(self) -> int:
        return int(self._rec.host()[6])


class SimulationMdp(AbstractMdp):
    """PKG/mdp.py:572-886: x and y discretisation, terminal chain without goal logic, no reward."""

    def __init__(self, working_curriculum_step: int, f_ag: float, t_max: int, *, p_max: float = 4.5, minimum_altitude: float = 0.2, **kw) -> None:
        super().__init__(working_curriculum_step, f_ag, t_max, p_max, minimum_altitude=minimum_altitude, **kw)
        self._rec = _MdpRecord(self._fh, working_curriculum_step, simulation=True)      # x axis + check
        self._rec_y = _MdpRecord(self._fh, working_curriculum_step, simulation=True)    # y axis (observe only)
        self._current_continuous_action = Action(pitch=0, roll=0, yaw=0, v_z=-0.4)
        self._current_discrete_state_x = self._previous_discrete_state_x = None
        self._current_discrete_state_y = self._previous_discrete_state_y = None
        self._current_continuous_observation = ContinuousObservation()

    def reset(self):
        """PKG/mdp.py:879-886."""
        self._rec.run(_ffi.OP_RESET)
        self._rec_y.run(_ffi.OP_RESET)
        self._info = {}
        self._current_continuous_observation = ContinuousObservation()
        self._current_discrete_state_x = self._previous_discrete_state_x = None
        self._current_discrete_state_y = self._previous_discrete_state_y = None
        self._current_continuous_action = Action(pitch=0.0, roll=0.0, yaw=0, v_z=-0.4)

    def discrete_state(self, current_continuous_observation: ContinuousObservation):
        """PKG/mdp.py:625-632."""
        self._previous_discrete_state_x, self._previous_discrete_state_y = self._current_discrete_state_x, self._current_discrete_state_y
        self._current_continuous_observation = current_continuous_observation
        return self.discrete_state_x(), self.discrete_state_y()

    def discrete_state_x(self):
        """PKG/mdp.py:634-707."""
        o = self._current_continuous_observation
        self._rec.set_obs(o.rel_p_x, o.rel_v_x, o.rel_a_x, o.pitch, o.abs_p_z, o.rel_p_y, o.contact)
        self._rec.run(_ffi.OP_OBSERVE)
        self._current_discrete_state_x = state_tuple(int(self._rec.out_state.item()))
        return self._current_discrete_state_x

    def discrete_state_y(self):
        """PKG/mdp.py:709-782."""
        o = self._current_continuous_observation
        self._rec_y.set_obs(o.rel_p_y, o.rel_v_y, o.rel_a_y, o.roll, o.abs_p_z, o.rel_p_x, o.contact)
        self._rec_y.run(_ffi.OP_OBSERVE)
        self._current_discrete_state_y = state_tuple(int(self._rec_y.out_state.item()))
        return self._current_discrete_state_y

    def continuous_action(self, action_x: int, action_y: int):
        """PKG/mdp.py:847-877 (the y branch is dead code in the reference: `if False and ...`)."""
        self._rec.action.fill_(int(action_x))
        self._rec.run(_ffi.OP_ACTION)
        self._current_continuous_action.pitch = float(self._rec.state[0].item())
        return self._current_continuous_action

    def check(self) -> Dict[str, Any]:
        """PKG/mdp.py:784-845."""
        if not self._current_discrete_state_x or not self._current_discrete_state_y:
            raise ValueError("Cannot check an empty state\nYou must call `discrete_state` before calling check.")
        self._rec.run(_ffi.OP_CHECK)
        st = self._rec.host()
        code, steps = int(st[7]), int(st[5])
        self._check_result = _CODE_TO_RESULT[code]
        if code >= 2:
            self._termination_info(self._info, code, steps)
        return self._info

    def reward(self) -> float:
        return 0.0
