"""Drop-in mirror of the reference's DoubleQLearningAgent (PKG/double_q_learning.py:32-146).

Same constructor, attributes (`Q_table_a`, `Q_table_b`, `state_action_counter`: float64 NumPy arrays of shape
(curriculum_steps, 3, 3, 3, 7, 3)), methods and `.npy` files.  The arithmetic of `predict` / `update` /
`transfer_learning` runs on the GPU in float64 (csrc/facade_kernels.cuh: agent_facade_kernel); the NumPy arrays are
host mirrors, kept coherent lazily.  The batched trainer (`trainer.Trainer`) uses float32 device tables
instead and writes its result back into these attributes.

Coherence rule of the mirrors: the three arrays are STABLE objects (a GPU update is copied into them in place, so a
reference taken once -- `qa = agent.Q_table_a` -- keeps showing the current values after the next attribute access or
`save`).  Reading an attribute marks the host side as possibly modified (the reference's arrays are mutable in place):
it is uploaded before the next GPU call.  An in-place write through a reference taken BEFORE a later `update()` is only
picked up if the attribute is read again first; `counter(state_action)` reads one count without forcing an upload.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import Tuple, Union

import numpy as np
import torch

from . import ASSETS_PATH, _ffi
from . import constants as K
from .mdp import _FacadeHandle, state_id

State = Tuple[int, int, int, int, int]
StateAction = Tuple[int, int, int, int, int, int]


class DoubleQLearningAgent:
    """Agent that learns and makes decisions."""

    def __init__(self, curriculum_steps: int = 5, *, device: int = 0) -> None:
        if not (1 <= curriculum_steps <= K.MAX_CURRICULUM):
            raise ValueError("curriculum_steps must be in 1..5")
        self.curriculum_steps = curriculum_steps
        shape = (curriculum_steps, 3, 3, 3, 7, 3)
        self._host = [np.zeros(shape), np.zeros(shape), np.zeros(shape)]
        self._device_index = device
        self._dev = None              # torch.float64 [3, MAX_CELLS], created on first GPU call
        self._host_dirty = False      # host arrays newer than the device copy
        self._dev_dirty = False       # device copy newer than the host arrays
        self._fh = None

    # -- attribute mirrors --------------------------------------------------------------------------
    def _pull(self):
        if self._dev_dirty:
            n = self.curriculum_steps * K.CELLS_PER_LEVEL
            h = self._dev.cpu().numpy()
            for i in range(3):      # in place: references handed out earlier stay live (like the reference's own arrays)
                if self._host[i].dtype == np.float64 and self._host[i].flags.writeable:
                    self._host[i][...] = h[i, :n].reshape(self._host[i].shape)
                else:
                    self._host[i] = h[i, :n].reshape(self._host[i].shape).copy()
            self._dev_dirty = False

    def _get(self, i):
        self._pull()
        self._host_dirty = True       # the caller may mutate the array in place, like with the reference
        return self._host[i]

    def _set(self, i, v):
        self._pull()
        self._host[i] = np.asarray(v)
        self._host_dirty = True

    def counter(self, state_action: StateAction) -> float:
        """state_action_counter[state_action] without marking the host mirror as modified (Trainer.alpha reads one count per step;
        through the attribute every read would force a re-upload of all three tables before the next GPU call)."""
        self._pull()
        return float(self._host[2][tuple(state_action)])

    Q_table_a = property(lambda self: self._get(0), lambda self, v: self._set(0, v))
    Q_table_b = property(lambda self: self._get(1), lambda self, v: self._set(1, v))
    state_action_counter = property(lambda self: self._get(2), lambda self, v: self._set(2, v))

    def _push(self):
        if self._fh is None:
            self._fh = _FacadeHandle.get(K.MdpParameters(), self._device_index)
            self._dev = torch.zeros((3, K.MAX_CELLS), dtype=torch.float64, device=self._fh.device)
            self._host_dirty = True
        if self._host_dirty:
            n = self.curriculum_steps * K.CELLS_PER_LEVEL
            h = np.zeros((3, K.MAX_CELLS))
            for i in range(3):
                if self._host[i].shape != (self.curriculum_steps, 3, 3, 3, 7, 3):
                    raise ValueError(f"table {i} has shape {self._host[i].shape}")
                h[i, :n] = np.asarray(self._host[i], np.float64).reshape(-1)
            self._dev.copy_(torch.from_numpy(h))
            self._host_dirty = False

    def _call(self, op, state, action=None, next_state=None, alpha=None, reward=None, gamma=0.0, want_action=False):
        self._push()
        fh, dev = self._fh, self._fh.device
        t = lambda v, dt: torch.tensor([v], dtype=dt, device=dev) if v is not None else None
        s, a, s2 = t(state, torch.int32), t(action, torch.int32), t(next_state, torch.int32)
        al, rw = t(alpha, torch.float64), t(reward, torch.float64)
        out = torch.zeros(1, dtype=torch.int32, device=dev) if want_action else None
        p = lambda x: x.data_ptr() if x is not None else None
        _ffi.check(fh.lib.dqlb200_agent_facade(fh.handle, op, 1, self._dev.data_ptr(), p(s), p(a), p(s2), p(al), p(rw),
                                               float(gamma), p(out), fh.stream()))
        if op != _ffi.AGENT_PREDICT:
            self._dev_dirty = True
        return int(out.item()) if want_action else None

    # -- reference API ------------------------------------------------------------------------------
    def save(self, save_path: Path):
        """PKG/double_q_learning.py:42-53."""
        save_path = Path(save_path)
        self._pull()
        for name, arr in (("Q_table_a.npy", self._host[0]), ("Q_table_b.npy", self._host[1]), ("state_action_count.npy", self._host[2])):
            with open(save_path / name, "wb") as f:
                np.save(f, np.asarray(arr, np.float64))

    @staticmethod
    def load(save_path: Path = ASSETS_PATH):
        """PKG/double_q_learning.py:55-75."""
        save_path = Path(save_path)
        with open(save_path / "Q_table_a.npy", "rb") as f:
            qa = np.load(f)
        with open(save_path / "Q_table_b.npy", "rb") as f:
            qb = np.load(f)
        with open(save_path / "state_action_count.npy", "rb") as f:
            sac = np.load(f)
        if qa.shape != qb.shape != sac.shape:
            raise ValueError(f"The shapes of Q table a {qa.shape}, Q table b {qb.shape}"
                             + f"and State action count {sac.shape} cannot be different")
        agent = DoubleQLearningAgent(len(qa))
        agent.Q_table_a, agent.Q_table_b, agent.state_action_counter = qa, qb, sac
        return agent

    def transfer_learning(self, current_curriculum_step: int, transfer_learning_ratio: float):
        """PKG/double_q_learning.py:77-89 (slot k from slot k-1; k = 0 reads the last slot, quirk Q7)."""
        step = int(current_curriculum_step) % self.curriculum_steps
        self._call(_ffi.AGENT_TRANSFER, step, next_state=(step - 1) % self.curriculum_steps, alpha=float(transfer_learning_ratio))

    def update(self, current_state_action: StateAction, next_state: State, alpha: float, gamma: float, reward):
        """PKG/double_q_learning.py:91-108: the table-pick draw is consumed, table A is updated either way (quirk Q1)."""
        np.random.uniform(0, 1)
        self._call(_ffi.AGENT_UPDATE, state_id(current_state_action[:5]), action=int(current_state_action[5]),
                   next_state=state_id(next_state), alpha=float(alpha), reward=float(reward), gamma=gamma)

    def guess(self, state: State, exploration_rate: float):
        """PKG/double_q_learning.py:110-117: both draws are always consumed (quirk Q4)."""
        explore = np.random.uniform(0, 1) < exploration_rate
        return int(np.where(explore, np.random.randint(3), self.predict(state)))

    def predict(self, state: State):
        """PKG/double_q_learning.py:119-124."""
        return self._call(_ffi.AGENT_PREDICT, state_id(state), want_action=True)
