"""TEST/BENCH INFRASTRUCTURE ONLY -- never imported by the product package.

BASELINE.md section 4 "CPU-baseline plan": the reference's single-env hot loop (PKG/trainer.py:191-212 with
PKG/landing_simulation_env.py:245-282 ordering) on the analytic stand-in, timed on host cores.  The MDP and
agent are the oracle restatements (oracle/mdp_oracle.py, oracle/agent_oracle.py, float64 tables like the
reference); the dynamics are the float64 'textbook' form of the stand-in (math.sin/tan -- the cheapest
faithful CPU form), the RNG is NumPy's MT19937 like the reference (PKG/trainer.py:45).  This loop is what
bench.py reports as `cpu_baseline` (kind "port") and runs on every host core for `--impl reference`.
"""
from __future__ import annotations

import math
import time

import numpy as np

from .agent_oracle import AgentOracle, exploration_rate
from .dynamics import StandInParams
from .mdp_oracle import MdpParams, TrainingMdpOracle


def run_single_env(n_steps: int, seed: int = 42, w: int = 0, sp: StandInParams = None):
    """Runs n_steps env-steps (incl. Q-updates); returns (steps, episodes, seconds)."""
    sp = sp or StandInParams()
    rng = np.random.RandomState(seed)
    agent = AgentOracle(5, np.float64)
    mdp = TrainingMdpOracle(w, sp.f_ag, 20, sp.p_max, MdpParams(), sp.v_z)
    h = 1.0 / sp.f_ag
    om = sp.v_mp / sp.r_mp
    k_th = -math.expm1(-h / sp.tau_theta)
    t_plat = rng.uniform(0.0, 2.0 * math.pi / om)
    episodes, ep = 0, 0
    t0 = time.perf_counter()

    def reset():
        nonlocal x, v, th, t_plat
        x_init = rng.normal(0.0, sp.p_max / 3) if w == 0 else rng.uniform(-sp.p_max, sp.p_max)
        x_mp = sp.r_mp * math.sin(om * t_plat)
        x = x_mp + min(max(x_init, -sp.p_max), sp.p_max)
        v = th = 0.0
        t_plat += h
        mdp.reset()
        return mdp.observe(sp.r_mp * math.sin(om * t_plat) - x, sp.r_mp * om * math.cos(om * t_plat) - v,
                           -sp.r_mp * om * om * math.sin(om * t_plat), th, sp.z_init, False)

    x = v = th = 0.0
    s = reset()
    for _ in range(n_steps):
        eps = exploration_rate(ep, w)
        explore = rng.uniform(0, 1) < eps
        k = rng.randint(3)
        a = int(k) if explore else agent.predict(s)
        th_sp = mdp.act(a)
        th = th + (th_sp - th) * k_th
        acc = sp.g * math.tan(th) - sp.c_d * v
        x = x + v * h + 0.5 * acc * h * h
        v = v + acc * h
        t_plat += h
        sn, cs = math.sin(om * t_plat), math.cos(om * t_plat)
        rel_p = sp.r_mp * sn - x
        z = sp.z_init + (mdp.step_count + 1) * sp.v_z * h
        s2 = mdp.observe(rel_p, sp.r_mp * om * cs - v, -sp.r_mp * om * om * sn - acc, th, z,
                         (z <= sp.z_touch) and abs(rel_p) <= sp.half_platform)
        _code, done = mdp.check()
        r = mdp.reward()
        sa = s + (a,)
        rng.uniform(0, 1)                                   # table-pick draw (quirk Q1)
        agent.update(sa, s2, agent.alpha(sa), r)
        if done:
            ep += 1
            episodes += 1
            s = reset()
        else:
            s = s2
    return n_steps, episodes, time.perf_counter() - t0


def _worker(args):
    n_steps, seed = args
    return run_single_env(n_steps, seed)


def run_all_cores(steps_per_proc: int, n_procs: int, seed0: int = 0, pool=None):
    """`--impl reference`: n_procs independent single-env loops, one per host core."""
    import multiprocessing as mp
    own = pool is None
    if own:
        pool = mp.get_context("fork").Pool(n_procs)
    try:
        t0 = time.perf_counter()
        res = pool.map(_worker, [(steps_per_proc, seed0 + i) for i in range(n_procs)])
        wall = time.perf_counter() - t0
    finally:
        if own:
            pool.close()
            pool.join()
    return sum(r[0] for r in res), wall
