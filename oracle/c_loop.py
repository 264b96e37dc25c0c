"""TEST/BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

ctypes driver of oracle/c/loop.c: the reference trainer loop for one environment in plain C (float32 tables, the Philox draw
contract, the deterministic stand-in).  tests/test_oracle_c.py holds it identical to the fixtures of the unmodified reference;
bench.py times it as a second, much stricter CPU baseline next to the Python port."""
from __future__ import annotations

import ctypes as C
import pathlib
import subprocess
import time

import numpy as np

from .dynamics import StandInDet, StandInParams, derive

CDIR = pathlib.Path(__file__).resolve().parent / "c"


class Params(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("h", "half_h2", "k_theta", "g", "c_d", "r", "rw", "rw2")] + [
        ("dphase", C.c_uint32), ("n_sub", C.c_int32), ("accel_mode", C.c_int32), ("kf_q", C.c_float), ("kf_r", C.c_float),
        ("second_order", C.c_int32), ("pid_ticks", C.c_int32)] + [(n, C.c_float) for n in (
            "att_kr", "att_kw", "inv_m", "inv_mg", "g_abs", "pid_kp", "pid_ki", "pid_lo", "pid_hi", "pid_windup", "pid_dt",
            "bw_inv_denom", "bw_k2", "z_init")]


class LoopParams(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("dz", "z_touch", "half_platform", "p_max_f", "two_p_max_f", "sigma_x")] + [
        (n, C.c_double) for n in ("f_ag", "t_max", "p_max", "alpha_min", "omega", "gamma")]


def load() -> C.CDLL:
    so = CDIR / "libstandin_oracle.so"
    if not so.exists():
        subprocess.run(["make", "-C", str(CDIR)], check=True, capture_output=True)
    lib = C.CDLL(str(so))
    lib.mdp_sizeof.restype = C.c_size_t
    lib.oracle_single_env_loop.argtypes = [C.c_void_p] * 3 + [C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 11
    return lib


def run_single_env_c(n_steps: int, seed: int = 42, w: int = 0):
    """n_steps env-steps incl. Q-updates of one env at curriculum step w; returns (steps, episodes, seconds)."""
    lib = load()
    sp = StandInParams()
    d = derive(sp)
    p = Params(h=d.h, half_h2=d.half_h2, k_theta=d.k_theta, g=d.g, c_d=d.c_d, r=d.r, rw=d.rw, rw2=d.rw2, dphase=d.dphase, n_sub=d.n_sub,
               z_init=d.z_init)
    lp = LoopParams(dz=d.dz, z_touch=d.z_touch, half_platform=d.half_platform, p_max_f=d.p_max, two_p_max_f=d.two_p_max, sigma_x=d.sigma_x,
                    f_ag=sp.f_ag, t_max=20.0, p_max=sp.p_max, alpha_min=0.02949, omega=0.51, gamma=0.99)
    qa, qb, count = np.zeros(2835, np.float32), np.zeros(2835, np.float32), np.zeros(2835, np.float64)
    obs = np.zeros((n_steps, 5), np.float32)
    ints = [np.zeros(n_steps, np.int32) for _ in range(6)]
    reward = np.zeros(n_steps, np.float64)
    buf = C.create_string_buffer(lib.mdp_sizeof())
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    t0 = time.perf_counter()
    episodes = lib.oracle_single_env_loop(C.byref(p), C.byref(lp), C.cast(buf, C.c_void_p), seed, 0, w, 0, n_steps, ptr(qa), ptr(qb), ptr(count),
                                          ptr(obs), *[ptr(a) for a in ints[:5]], ptr(reward), ptr(ints[5]))
    return n_steps, int(episodes), time.perf_counter() - t0
