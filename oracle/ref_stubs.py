"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the UNMODIFIED reference modules (mdp.py, double_q_learning.py,
trainer.py) from /root/reference by installing four stub modules for the
ROS-only dependencies they pull in at import time:

  * ``rospkg``                                   (PKG/__init__.py:3-7 -> ASSETS_PATH)
  * ``dql_multirotor_landing.msg``               (PKG/mdp.py:8 -> Action, Observation;
                                                  field lists from msg/Action.msg,
                                                  msg/Observation.msg)
  * ``gym``                                      (PKG/trainer.py:10)
  * ``dql_multirotor_landing.landing_simulation_env`` (PKG/trainer.py:16)

/root/reference exists only in the build container, not on the GPU box, so
this module is used exclusively by ``oracle/gen_golden.py`` (which writes the
committed fixtures under tests/golden/) and by CPU tests that skip when the
reference tree is absent.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DQL_REFERENCE_ROOT", "/root/reference")
_PKG_PARENT = os.path.join(REFERENCE_ROOT, "src", "dql_multirotor_landing", "src")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(_PKG_PARENT, "dql_multirotor_landing"))


class Observation:
    """Plain-Python stand-in for the catkin-generated Observation message."""

    __slots__ = (
        "rel_p_x", "rel_p_y", "rel_p_z", "rel_v_x", "rel_v_y", "rel_v_z",
        "rel_a_x", "rel_a_y", "rel_a_z", "contact",
    )

    def __init__(self, rel_p_x=0.0, rel_p_y=0.0, rel_p_z=0.0, rel_v_x=0.0, rel_v_y=0.0,
                 rel_v_z=0.0, rel_a_x=0.0, rel_a_y=0.0, rel_a_z=0.0, contact=False):
        self.rel_p_x, self.rel_p_y, self.rel_p_z = rel_p_x, rel_p_y, rel_p_z
        self.rel_v_x, self.rel_v_y, self.rel_v_z = rel_v_x, rel_v_y, rel_v_z
        self.rel_a_x, self.rel_a_y, self.rel_a_z = rel_a_x, rel_a_y, rel_a_z
        self.contact = contact


class Action:
    """Plain-Python stand-in for the catkin-generated Action message."""

    __slots__ = ("roll", "pitch", "yaw", "v_z")

    def __init__(self, roll=0.0, pitch=0.0, yaw=0.0, v_z=0.0):
        self.roll, self.pitch, self.yaw, self.v_z = roll, pitch, yaw, v_z


_installed = None


def install():
    """Install the stubs and return a namespace with the reference modules."""
    global _installed
    if _installed is not None:
        return _installed
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")

    rospkg = types.ModuleType("rospkg")

    class RosPack:
        def get_path(self, name):
            return os.path.join(REFERENCE_ROOT, "src", "dql_multirotor_landing")

    rospkg.RosPack = RosPack
    sys.modules["rospkg"] = rospkg

    gym = types.ModuleType("gym")

    class Env:
        pass

    def make(*a, **k):
        raise RuntimeError("gym.make is stubbed: the oracle drives the MDP directly")

    gym.Env, gym.make = Env, make
    sys.modules.setdefault("gym", gym)

    if _PKG_PARENT not in sys.path:
        sys.path.insert(0, _PKG_PARENT)
    pkg = importlib.import_module("dql_multirotor_landing")

    msg = types.ModuleType("dql_multirotor_landing.msg")
    msg.Observation, msg.Action = Observation, Action
    sys.modules["dql_multirotor_landing.msg"] = msg
    pkg.msg = msg

    env = types.ModuleType("dql_multirotor_landing.landing_simulation_env")

    class TrainingLandingEnv:
        pass

    env.TrainingLandingEnv = TrainingLandingEnv
    sys.modules["dql_multirotor_landing.landing_simulation_env"] = env

    ns = types.SimpleNamespace()
    ns.mdp = importlib.import_module("dql_multirotor_landing.mdp")
    ns.dql = importlib.import_module("dql_multirotor_landing.double_q_learning")
    ns.trainer = importlib.import_module("dql_multirotor_landing.trainer")
    ns.Observation, ns.Action = Observation, Action
    ns.assets = os.path.join(REFERENCE_ROOT, "assets")
    _installed = ns
    return ns
