"""TEST INFRASTRUCTURE ONLY.  Generates the committed fixtures under tests/golden/ by running the
UNMODIFIED reference modules (imported from /root/reference through oracle/ref_stubs.py).

Run in the build container only (the reference tree does not exist on the GPU box):

    python -m oracle.gen_golden

Fixtures:
  discretise.npz   random + near-threshold fp32 observations -> TrainingMdp/SimulationMdp state tuples, w = 0..4
  mdp_trace_w*.npz forced-action episodes: reference TrainingMdp check codes, rewards (float64), states
  replay_*.npz     the reference trainer loop (guess/update/alpha/exploration_rate + TrainingMdp), single env,
                   Philox draws injected through np.random, float32 tables (NEP 50) and float64 tables
  schedules.npz    Trainer.alpha / exploration_rate / transfer_learning_ratio tables
  sim_trace.npz    SimulationMdp greedy episodes with the committed assets policy
  sim2d_trace.npz  two-axis SimulationMdp episodes (x and y states, FLYZONE_Y, contact on both axes), three platform cases
  kalman_accel.npz the reference KalmanFilter3D (PKG/filters.py) driven the way ObservationUtils drives it, on stand-in velocities
  curriculum_ref.npz the UNMODIFIED Trainer.curriculum_training() (PKG/trainer.py:169-245) driven through a gym.make stub whose env wraps
                   the reference TrainingMdp in the order of PKG/landing_simulation_env.py:167-282 on the stand-in: success window,
                   promotions, a max-episodes advance (window kept), transfer after every step incl. the last (quirk Q7), final tables
  second_order.npz the reference PID node (PKG/pid.py, Butterworth filter included) on the v_z errors of a second-order stand-in
                   run, and the reference AttitudeController moment (PKG/attitude_controller.py:124-156) for pure pitch states
"""
from __future__ import annotations

import os
import pathlib
import tempfile

import numpy as np

from . import philox, ref_stubs
from .agent_oracle import state_id
from .dynamics import StandInDet, StandInParams
from .mdp_oracle import TERMINATION_STRINGS, NON_TERMINAL, NON_TERMINAL_SUCCESS

GOLDEN = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"
F_AG, T_MAX, P_MAX = 22.92, 20, 4.5
_STR2CODE = {v: k for k, v in TERMINATION_STRINGS.items()}


def _ref_code(mdp, ns) -> int:
    cr = mdp._check_result
    if cr == ns.mdp.CheckResult.NON_TERMINAL:
        return NON_TERMINAL
    if cr == ns.mdp.CheckResult.NON_TERMINAL_SUCCESS:
        return NON_TERMINAL_SUCCESS
    return _STR2CODE[cr.value]


def _obs(ns, rel_p, rel_v, rel_a, pitch, z, contact):
    o = ns.Observation(rel_p_x=float(rel_p), rel_v_x=float(rel_v), rel_a_x=float(rel_a), contact=bool(contact))
    return ns.mdp.ContinuousObservation(o, float(pitch), 0.0, float(z))


def probe_values(rng, n_random=4000):
    """fp32 probes: broad random values plus +-2 ulp neighbourhoods of every threshold candidate."""
    from .mdp_oracle import LIMITS_P, LIMITS_V, MdpParams, linspace7
    prm = MdpParams()
    cands_p, cands_v, cands_a, cands_t = [], [], [], []
    for w in range(5):
        for l in range(w + 1):
            for sign in (-1, 1):
                cands_p += [sign * LIMITS_P[l] * P_MAX, sign * LIMITS_P[l] * prm.beta * P_MAX]
                cands_v += [sign * LIMITS_V[l] * prm.v_max, sign * LIMITS_V[l] * prm.beta * prm.v_max]
                if l < w:
                    cands_p.append(sign * LIMITS_P[l] * (LIMITS_P[l + 1] / LIMITS_P[l]) * P_MAX)
                    cands_v.append(sign * LIMITS_V[l] * (LIMITS_V[l + 1] / LIMITS_V[l]) * prm.v_max)
    for sign in (-1, 1):
        cands_a += [sign * prm.a_max, sign * prm.sigma_a * prm.a_max, sign * prm.sigma_a * prm.beta * prm.a_max]
    ang = linspace7(prm.theta_max, prm.n_theta)
    cands_t += ang + [(ang[i] + ang[i + 1]) / 2 for i in range(6)]

    def around(c):
        x = np.float32(c)
        out = [x]
        lo = hi = x
        for _ in range(3):
            lo = np.nextafter(lo, np.float32(-np.inf)); hi = np.nextafter(hi, np.float32(np.inf))
            out += [lo, hi]
        return out

    def expand(cands, scale):
        vals = [v for c in cands for v in around(c)]
        vals += list((rng.standard_normal(n_random) * scale).astype(np.float32))
        vals += [np.float32(0.0), np.float32(-0.0)]
        return np.asarray(vals, np.float32)

    return expand(cands_p, 2.5), expand(cands_v, 2.0), expand(cands_a, 1.0), expand(cands_t, 0.25)


def gen_discretise(ns):
    rng = np.random.default_rng(1234)
    P, V, A, T = probe_values(rng)
    n = 6000
    # every probe value appears at least once; the other coordinates are random picks
    rows = []
    for arr, col in ((P, 0), (V, 1), (A, 2), (T, 3)):
        for x in arr:
            r = [rng.choice(P), rng.choice(V), rng.choice(A), rng.choice(T)]
            r[col] = x
            rows.append(r)
    obs = np.asarray(rows, np.float32)
    out_train = np.zeros((5, len(obs), 5), np.int8)
    out_sim = np.zeros((5, len(obs), 5), np.int8)
    for w in range(5):
        tm = ns.mdp.TrainingMdp(w, F_AG, T_MAX, P_MAX)
        sm = ns.mdp.SimulationMdp(w, F_AG, T_MAX)
        tm.reset(); sm.reset()
        for k, (p, v, a, t) in enumerate(obs):
            out_train[w, k] = tm.discrete_state(_obs(ns, p, v, a, t, 3.0, False))
            out_sim[w, k] = sm.discrete_state(_obs(ns, p, v, a, t, 3.0, False))[0]
    np.savez_compressed(GOLDEN / "discretise.npz", obs=obs, train=out_train, sim=out_sim)
    print("discretise:", obs.shape)


def gen_mdp_trace(ns, w, n_episodes=12, seed=7, sp=None, tag=None):
    """Forced random actions through the stand-in; the reference MDP consumes the fp32 observations."""
    rng = np.random.default_rng(seed + w)
    sp = sp or StandInParams()
    dyn = StandInDet(sp, 1)
    mdp = ns.mdp.TrainingMdp(w, F_AG, T_MAX, P_MAX)
    rec = {k: [] for k in ("obs", "contact", "action", "state", "code", "done", "reward", "theta_sp", "episode", "cum")}
    idx = np.asarray([0])
    t = 0                      # global step index: reset draws use birth = t (the RNG contract)
    for ep in range(n_episodes):
        words = philox.draws(99, 0, idx, t, philox.PURPOSE_RESET)
        dyn.reset(idx, words[0], words[1], words[2], normal_init=(w == 0))
        mdp.reset()
        dyn.advance(np.zeros(1, np.float32))
        rp, rv, ra, pit, z, c = (x[0] for x in dyn.observe(np.zeros(1)))
        s = mdp.discrete_state(_obs(ns, rp, rv, ra, pit, z, c))
        rec["obs"].append((rp, rv, ra, pit, z)); rec["contact"].append(c); rec["action"].append(255)
        rec["state"].append(state_id(s)); rec["code"].append(0); rec["done"].append(0); rec["reward"].append(0.0)
        rec["theta_sp"].append(0.0); rec["episode"].append(ep); rec["cum"].append(0.0)
        # a policy mix so that goal states, fly-zone exits and timeouts all occur
        mode = ep % 4
        done, k = False, 0
        while not done:
            if mode == 0:
                a = int(rng.integers(3))
            elif mode == 1:
                a = 2
            elif mode == 2:   # crude stabiliser -> reaches the goal bins
                a = 0 if (rp + 0.8 * rv) > 0.05 else (1 if (rp + 0.8 * rv) < -0.05 else 2)
                if abs(pit) > 0.2:
                    a = 1 if pit > 0 else 0
            else:
                a = int(rng.integers(2))
            act = mdp.continuous_action(a, 2)
            dyn.advance(np.asarray([act.pitch], np.float32))
            k += 1
            t += 1
            rp, rv, ra, pit, z, c = (x[0] for x in dyn.observe(np.asarray([k])))
            s = mdp.discrete_state(_obs(ns, rp, rv, ra, pit, z, c))
            info = mdp.check()
            r = mdp.reward()
            done = "Termination condition" in info
            rec["obs"].append((rp, rv, ra, pit, z)); rec["contact"].append(c); rec["action"].append(a)
            rec["state"].append(state_id(s)); rec["code"].append(_ref_code(mdp, ns)); rec["done"].append(int(done))
            rec["reward"].append(r); rec["theta_sp"].append(act.pitch); rec["episode"].append(ep)
            rec["cum"].append(mdp._cumulative_reward)
    out = dict(
        obs=np.asarray(rec["obs"], np.float32), contact=np.asarray(rec["contact"], np.uint8),
        action=np.asarray(rec["action"], np.uint8), state=np.asarray(rec["state"], np.uint16),
        code=np.asarray(rec["code"], np.uint8), done=np.asarray(rec["done"], np.uint8),
        reward=np.asarray(rec["reward"], np.float64), theta_sp=np.asarray(rec["theta_sp"], np.float64),
        episode=np.asarray(rec["episode"], np.int32), cum=np.asarray(rec["cum"], np.float64), w=np.int32(w),
        z_init=np.float64(sp.z_init), v_z=np.float64(sp.v_z), v_mp=np.float64(sp.v_mp), seed=np.int64(99),
    )
    np.savez_compressed(GOLDEN / f"mdp_trace_{tag or ('w%d' % w)}.npz", **out)
    codes, cnt = np.unique(out["code"][out["done"] == 1], return_counts=True)
    print(f"mdp_trace w={w}: {len(out['action'])} rows, terminal codes {dict(zip(codes.tolist(), cnt.tolist()))}")


class _DrawFeeder:
    """Feeds Philox words to the reference through np.random (order per SURVEY.md A.1 Q4)."""

    def __init__(self):
        self.queue = []

    def uniform(self, lo=0.0, hi=1.0, size=None):
        return float(philox.uniform01(self.queue.pop(0))) * (hi - lo) + lo

    def randint(self, n):
        return int(philox.random_action(self.queue.pop(0)))


def gen_replay(ns, w, n_steps, dtype, seed=42, tag="", q_init=None, ep0=0):
    """The reference trainer loop body (PKG/trainer.py:187-236) on one env."""
    feeder = _DrawFeeder()
    saved = (np.random.uniform, np.random.randint)
    np.random.uniform, np.random.randint = feeder.uniform, feeder.randint
    try:
        trainer = ns.trainer.Trainer(save_path=pathlib.Path(tempfile.mkdtemp()), seed=seed)
        agent = trainer._double_q_learning_agent
        if q_init is not None:
            agent.Q_table_a = q_init[0].astype(dtype).copy()
            agent.Q_table_b = q_init[1].astype(dtype).copy()
        else:
            agent.Q_table_a = agent.Q_table_a.astype(dtype)
            agent.Q_table_b = agent.Q_table_b.astype(dtype)
        qa0, qb0 = agent.Q_table_a.copy(), agent.Q_table_b.copy()
        sp = StandInParams()
        dyn = StandInDet(sp, 1)
        mdp = ns.mdp.TrainingMdp(w, F_AG, T_MAX, P_MAX)
        idx = np.asarray([0])
        rec = {k: [] for k in ("obs", "action", "state", "next_state", "code", "done", "reward", "episode", "alpha")}
        t, ep = 0, ep0

        def reset(birth):
            words = philox.draws(seed, 0, idx, birth, philox.PURPOSE_RESET)
            dyn.reset(idx, words[0], words[1], words[2], normal_init=(w == 0))
            mdp.reset()
            dyn.advance(np.zeros(1, np.float32))
            rp, rv, ra, pit, z, c = (x[0] for x in dyn.observe(np.zeros(1)))
            return mdp.discrete_state(_obs(ns, rp, rv, ra, pit, z, c))

        s = reset(0)
        k = 0
        while t < n_steps:
            words = philox.draws(seed, 0, idx, t, philox.PURPOSE_STEP)
            feeder.queue = [words[0][0], words[1][0], words[2][0]]
            a = agent.guess(s, trainer.exploration_rate(ep, w))
            act = mdp.continuous_action(a, 2)
            dyn.advance(np.asarray([act.pitch], np.float32))
            k += 1
            rp, rv, ra, pit, z, c = (x[0] for x in dyn.observe(np.asarray([k])))
            s2 = mdp.discrete_state(_obs(ns, rp, rv, ra, pit, z, c))
            info = mdp.check()
            r = mdp.reward()
            done = "Termination condition" in info
            sa = s + (a,)
            alpha = trainer.alpha(sa)
            agent.update(sa, s2, alpha, trainer._gamma, r)
            assert not feeder.queue
            rec["obs"].append((rp, rv, ra, pit, z)); rec["action"].append(a); rec["state"].append(state_id(s))
            rec["next_state"].append(state_id(s2)); rec["code"].append(_ref_code(mdp, ns)); rec["done"].append(int(done))
            rec["reward"].append(r); rec["episode"].append(ep); rec["alpha"].append(alpha)
            t += 1
            if done:
                ep += 1
                k = 0
                s = reset(t)
            else:
                s = s2
    finally:
        np.random.uniform, np.random.randint = saved
    out = dict(
        obs=np.asarray(rec["obs"], np.float32), action=np.asarray(rec["action"], np.uint8),
        state=np.asarray(rec["state"], np.uint16), next_state=np.asarray(rec["next_state"], np.uint16),
        code=np.asarray(rec["code"], np.uint8), done=np.asarray(rec["done"], np.uint8),
        reward=np.asarray(rec["reward"], np.float64), episode=np.asarray(rec["episode"], np.int32),
        alpha=np.asarray(rec["alpha"], np.float64), qa=agent.Q_table_a, qb=agent.Q_table_b,
        count=agent.state_action_counter, qa0=qa0, qb0=qb0, w=np.int32(w), seed=np.int64(seed), ep0=np.int32(ep0),
    )
    name = f"replay_w{w}_{np.dtype(dtype).name}{tag}.npz"
    np.savez_compressed(GOLDEN / name, **out)
    print(f"{name}: {n_steps} steps, {ep} episodes, visited cells {int((agent.state_action_counter > 0).sum())},"
          f" explore-free steps {int((np.asarray(rec['episode']) > 800).sum())}")


def gen_schedules(ns):
    trainer = ns.trainer.Trainer(save_path=pathlib.Path(tempfile.mkdtemp()))
    agent = trainer._double_q_learning_agent
    sa = (0, 1, 1, 1, 3, 2)
    alphas = []
    for c in range(0, 1200):
        agent.state_action_counter[sa] = c
        alphas.append(trainer.alpha(sa))
    eps = [trainer.exploration_rate(e, 0) for e in range(0, 2300)]
    eps1 = [trainer.exploration_rate(e, 1) for e in range(0, 10)]
    ratios = [trainer.transfer_learning_ratio(k) for k in range(5)]
    np.savez_compressed(GOLDEN / "schedules.npz", alpha=np.asarray(alphas), eps=np.asarray(eps), eps1=np.asarray(eps1),
                        ratios=np.asarray(ratios))
    print("schedules ok")


def gen_sim_trace(ns, n_episodes=6, seed=5):
    """Greedy SimulationMdp episodes with the committed policy (scripts/simulation.py:48-63)."""
    agent = ns.dql.DoubleQLearningAgent.load()
    sp = StandInParams(v_z=-0.4)
    dyn = StandInDet(sp, 1)
    mdp = ns.mdp.SimulationMdp(4, F_AG, T_MAX)
    idx = np.asarray([0])
    rec = {k: [] for k in ("obs", "contact", "action", "state", "code", "done", "episode")}
    for ep in range(n_episodes):
        words = philox.draws(seed, 0, idx + ep, 0, philox.PURPOSE_RESET)
        dyn.reset(idx, words[0], words[1], words[2], normal_init=False, simulation=True)
        mdp.reset()
        dyn.advance(np.zeros(1, np.float32))
        rp, rv, ra, pit, z, c = (x[0] for x in dyn.observe(np.zeros(1)))
        sx, _sy = mdp.discrete_state(_obs(ns, rp, rv, ra, pit, z, c))
        rec["obs"].append((rp, rv, ra, pit, z)); rec["contact"].append(c); rec["action"].append(255)
        rec["state"].append(state_id(sx)); rec["code"].append(0); rec["done"].append(0); rec["episode"].append(ep)
        done, k = False, 0
        while not done:
            a = agent.predict(sx)
            act = mdp.continuous_action(a, 2)
            dyn.advance(np.asarray([act.pitch], np.float32))
            k += 1
            rp, rv, ra, pit, z, c = (x[0] for x in dyn.observe(np.asarray([k])))
            sx, _sy = mdp.discrete_state(_obs(ns, rp, rv, ra, pit, z, c))
            info = mdp.check()
            done = "Termination condition" in info
            rec["obs"].append((rp, rv, ra, pit, z)); rec["contact"].append(c); rec["action"].append(a)
            rec["state"].append(state_id(sx)); rec["code"].append(_ref_code(mdp, ns)); rec["done"].append(int(done))
            rec["episode"].append(ep)
    out = dict(obs=np.asarray(rec["obs"], np.float32), contact=np.asarray(rec["contact"], np.uint8),
               action=np.asarray(rec["action"], np.uint8), state=np.asarray(rec["state"], np.uint16),
               code=np.asarray(rec["code"], np.uint8), done=np.asarray(rec["done"], np.uint8),
               episode=np.asarray(rec["episode"], np.int32), seed=np.int64(seed))
    np.savez_compressed(GOLDEN / "sim_trace.npz", **out)
    codes, cnt = np.unique(out["code"][out["done"] == 1], return_counts=True)
    print(f"sim_trace: {len(out['action'])} rows, terminal codes {dict(zip(codes.tolist(), cnt.tolist()))}")


def gen_sim2d_trace(ns, n_episodes=4, seed=9):
    """Two-axis greedy SimulationMdp episodes: the UNMODIFIED reference SimulationMdp discretises both axes and runs the
    terminal chain (PKG/mdp.py:625-845) on the observations of the two-axis stand-in; pitch set-points come from the
    reference's continuous_action, roll set-points (dead code in the reference, PKG/mdp.py:863-876) from the same
    min/max expressions when the case enables them."""
    from .dynamics import StandIn2D, sim2d_cases
    from .loop import mirrored_policy
    agent = ns.dql.DoubleQLearningAgent.load()
    lut_x = np.asarray([agent.predict(_state_tuple(sid)) for sid in range(945)], np.uint8)
    lut_y = mirrored_policy(lut_x)
    theta_max, delta_theta = np.deg2rad(21.37723), np.deg2rad(7.12574)
    cases = sim2d_cases()
    luts_y = {"reference": lut_y, "xy": lut_y, "eight": lut_y, "ywrong": lut_x}
    out = dict(seed=np.int64(seed), lut_x=lut_x)
    keys = ("rel_p_x", "rel_v_x", "rel_a_x", "pitch", "z", "rel_p_y", "rel_v_y", "rel_a_y", "roll")
    for ci, (name, p2) in enumerate(cases.items()):
        lut_y = luts_y[name]
        out[f"{name}_lut_y"] = lut_y
        rec = {k: [] for k in ("obs", "contact", "action_x", "action_y", "state_x", "state_y", "code", "done", "episode")}
        for ep in range(n_episodes):
            dyn = StandIn2D(p2, 1)
            mdp = ns.mdp.SimulationMdp(4, F_AG, T_MAX)
            w = philox.draws(seed, ci, np.asarray([ep]), 0, philox.PURPOSE_RESET)
            dyn.reset(*w)
            mdp.reset()
            dyn.advance(np.zeros(1, np.float32), np.zeros(1, np.float32))

            def look(k):
                o = {kk: v[0] for kk, v in dyn.observe(np.asarray([k])).items()}
                ob = ns.Observation(rel_p_x=float(o["rel_p_x"]), rel_v_x=float(o["rel_v_x"]), rel_a_x=float(o["rel_a_x"]),
                                    rel_p_y=float(o["rel_p_y"]), rel_v_y=float(o["rel_v_y"]), rel_a_y=float(o["rel_a_y"]),
                                    contact=bool(o["contact"]))
                sx, sy = mdp.discrete_state(ns.mdp.ContinuousObservation(ob, float(o["pitch"]), float(o["roll"]), float(o["z"])))
                return o, sx, sy

            def push(o, ax, ay, sx, sy, code, done):
                rec["obs"].append([o[k] for k in keys]); rec["contact"].append(o["contact"])
                rec["action_x"].append(ax); rec["action_y"].append(ay)
                rec["state_x"].append(state_id(sx)); rec["state_y"].append(state_id(sy))
                rec["code"].append(code); rec["done"].append(int(done)); rec["episode"].append(ep)

            o, sx, sy = look(0)
            push(o, 255, 255, sx, sy, 0, False)
            roll_sp, done, k = 0.0, False, 0
            while not done:
                ax, ay = int(agent.predict(sx)), int(lut_y[state_id(sy)])
                act = mdp.continuous_action(ax, ay)
                if p2.y_action_enabled:
                    if ay == 0:
                        roll_sp = min((roll_sp + delta_theta, theta_max))
                    elif ay == 1:
                        roll_sp = max((roll_sp - delta_theta, -theta_max))
                dyn.advance(np.asarray([act.pitch], np.float32), np.asarray([roll_sp], np.float32))
                k += 1
                o, sx, sy = look(k)
                info = mdp.check()
                done = "Termination condition" in info
                push(o, ax, ay, sx, sy, _ref_code(mdp, ns), done)
        out[f"{name}_obs"] = np.asarray(rec["obs"], np.float32)
        for k2, dt in (("contact", np.uint8), ("action_x", np.uint8), ("action_y", np.uint8), ("state_x", np.uint16),
                       ("state_y", np.uint16), ("code", np.uint8), ("done", np.uint8), ("episode", np.int32)):
            out[f"{name}_{k2}"] = np.asarray(rec[k2], dt)
        codes, cnt = np.unique(out[f"{name}_code"][out[f"{name}_done"] == 1], return_counts=True)
        print(f"sim2d_trace[{name}]: {len(rec['code'])} rows, terminal codes {dict(zip(codes.tolist(), cnt.tolist()))}")
    np.savez_compressed(GOLDEN / "sim2d_trace.npz", **out)


def gen_kalman(n_steps=700, seed=11):
    """PKG/filters.py:4-80 (unmodified, imported with a stub for geometry_msgs.msg.Vector3Stamped) called with the protocol of
    PKG/observation_utils.py:134-150: the first sample sets last_velocity / last_timestep and reports 0; every later sample
    calls filter(current, t, last_velocity, last_timestep) -- and nothing ever refreshes the anchor ("anchor" outputs).  The
    "consecutive" outputs come from the same reference class with the anchor refreshed after every call (the evident
    intention).  Inputs: the true relative velocity of the stand-in at every 100 Hz sub-step (n_sub = 4) over several
    episodes with random set-points, teleport resets included; y carries the same signal negated, z is constant."""
    import importlib
    import sys
    import types
    gm = types.ModuleType("geometry_msgs")
    gmm = types.ModuleType("geometry_msgs.msg")

    class _V3:
        def __init__(self, x=0.0, y=0.0, z=0.0):
            self.x, self.y, self.z = x, y, z

    class Vector3Stamped:
        def __init__(self):
            self.vector = _V3()

    gmm.Vector3Stamped = Vector3Stamped
    gm.msg = gmm
    sys.modules.setdefault("geometry_msgs", gm)
    sys.modules.setdefault("geometry_msgs.msg", gmm)
    ref_stubs.install()
    filters = importlib.import_module("dql_multirotor_landing.filters")

    rng = np.random.default_rng(seed)
    sp = StandInParams(n_sub=4, accel_mode="kalman_reference")
    dyn = StandInDet(sp, 1)
    samples = []
    orig = dyn.kf.sample
    dyn.kf.sample = lambda idx, rel_v: (samples.append(np.float32(np.asarray(rel_v).reshape(-1)[0])), orig(idx, rel_v))[1]
    idx = np.arange(1)
    step = 0
    while step < n_steps:
        w = rng.integers(0, 2 ** 32, size=3, dtype=np.uint64).astype(np.uint32)
        dyn.reset(idx, w[0:1], w[1:2], w[2:3], normal_init=True)
        dyn.advance(np.zeros(1, np.float32), idx)
        theta = 0.0
        for _ in range(int(rng.integers(40, 200))):
            theta = float(np.clip(theta + rng.choice([-1, 0, 1]) * np.deg2rad(7.12574), -0.3731, 0.3731))
            dyn.advance(np.asarray([theta], np.float32), idx)
            step += 1
    v = np.asarray(samples, np.float32)
    h = float(dyn.d.h)
    out = dict(rel_v=v, h=np.float32(h), q=np.float64(1e-4))
    for sd in (0.1, 0.25):
        for mode in ("anchor", "consecutive"):
            kf = filters.KalmanFilter3D(process_variance=1e-4, measurement_variance=sd)      # scripts/manager_node.py:96-98
            last_v, last_t, acc = None, None, []
            for k, vk in enumerate(v):
                cur, t = _V3(float(vk), -float(vk), 0.5), k * h
                if last_v is None:
                    last_v, last_t = cur, t
                    acc.append((0.0, 0.0, 0.0))
                    continue
                a = kf.filter(current_rel_v=cur, timestep=t, last_vel=last_v, last_timestep=last_t)
                acc.append((a.vector.x, a.vector.y, a.vector.z))
                if mode == "consecutive":
                    last_v, last_t = cur, t
            out[f"{mode}_sd{sd}"] = np.asarray(acc, np.float64)
    np.savez_compressed(GOLDEN / "kalman_accel.npz", **out)
    print("kalman_accel:", len(v), "samples")


def gen_second_order(n_steps=400, seed=13):
    """SURVEY 8f-4 pins.  (1) PKG/pid.py PID.output() -- unmodified; the object is created without running __init__ (which
    spins forever in run()), its fields are set as __init__ / load_params set them (PKG/pid.py:14-23,33-48) with the gains of
    launch/drone.launch:33-46, rospy.Time.now() is a controllable clock at the node's 1 kHz and the publisher records the
    efforts.  Input: the held v_z errors of a second-order stand-in episode sequence (10 node iterations per 100 Hz sample).
    (2) AttitudeController._compute_desired_moment() -- unmodified, with tf.transformations replaced by the two textbook
    functions it uses -- for pure pitch attitudes: the y moment must be -k_R sin(theta - theta_sp) - k_w omega."""
    import importlib
    import sys
    import types
    ref_stubs.install()

    class _Dur:
        def __init__(self, s): self.s = s
        def to_sec(self): return self.s

    class _Time:
        def __init__(self, s): self.s = s
        def __sub__(self, o): return _Dur(self.s - o.s)
        def to_sec(self): return self.s

    clock = [0.0]
    rospy = types.ModuleType("rospy")
    rospy.Time = types.SimpleNamespace(now=lambda: _Time(clock[0]))
    rospy.logerr = lambda *a, **k: None
    sys.modules["rospy"] = rospy
    std = types.ModuleType("std_msgs")
    stdm = types.ModuleType("std_msgs.msg")

    class Float64:
        def __init__(self, data=0.0): self.data = data

    stdm.Float64 = Float64
    std.msg = stdm
    sys.modules["std_msgs"], sys.modules["std_msgs.msg"] = std, stdm
    gm, gmm = types.ModuleType("geometry_msgs"), types.ModuleType("geometry_msgs.msg")
    gmm.Vector3Stamped = type("Vector3Stamped", (), {})
    gm.msg = gmm
    sys.modules.setdefault("geometry_msgs", gm)
    sys.modules.setdefault("geometry_msgs.msg", gmm)
    tf = types.ModuleType("tf")
    tft = types.ModuleType("tf.transformations")

    def rotation_matrix(angle, axis):
        x, y, z = np.asarray(axis, float) / np.linalg.norm(axis)
        c, s_ = np.cos(angle), np.sin(angle)
        Km = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
        M = np.eye(4)
        M[:3, :3] = np.eye(3) * c + s_ * Km + (1 - c) * np.outer([x, y, z], [x, y, z])
        return M

    def quaternion_matrix(q):
        x, y, z, w = q
        M = np.eye(4)
        M[:3, :3] = [[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]]
        return M

    tft.rotation_matrix, tft.quaternion_matrix = rotation_matrix, quaternion_matrix
    tf.transformations = tft
    sys.modules["tf"], sys.modules["tf.transformations"] = tf, tft
    pid_mod = importlib.import_module("dql_multirotor_landing.pid")
    att_mod = importlib.import_module("dql_multirotor_landing.attitude_controller")
    from collections import deque

    # ---- (1) the vertical PID on a second-order stand-in run ------------------------------------------------------------
    rng = np.random.default_rng(seed)
    sp = StandInParams(n_sub=4, dynamics_model="second_order")
    dyn = StandInDet(sp, 1)
    errors, thrusts = [], []
    orig = dyn.pid.thrust
    def rec(idx, e):
        T = orig(idx, e)
        errors.append(np.float32(np.asarray(e).reshape(-1)[0]))
        thrusts.append(np.float32(T.reshape(-1)[0]))
        return T
    dyn.pid.thrust = rec
    idx = np.arange(1)
    step = 0
    while step < n_steps:
        w = rng.integers(0, 2 ** 32, size=3, dtype=np.uint64).astype(np.uint32)
        dyn.reset(idx, w[0:1], w[1:2], w[2:3], normal_init=True)
        dyn.advance(np.zeros(1, np.float32), idx, hover=True)
        theta = 0.0
        for _ in range(int(rng.integers(40, 160))):
            theta = float(np.clip(theta + rng.choice([-1, 0, 1]) * np.deg2rad(7.12574), -0.3731, 0.3731))
            dyn.advance(np.asarray([theta], np.float32), idx)
            step += 1
    errors = np.asarray(errors, np.float32)
    node = object.__new__(pid_mod.PID)
    node.rate_hz = 1000.0
    node.error = deque([0.0, 0.0], maxlen=3)
    node.error_deriv = deque([0.0, 0.0, 0.0], maxlen=3)
    node.filter_error, node.filter_deriv = pid_mod.ButterworthFilter(), pid_mod.ButterworthFilter()
    node.error_integral = sp.mass * sp.g / sp.pid_ki          # a simulator that has been hovering
    node.current_state, node.setpoint = 0.0, 0.0
    node.Kp, node.Ki, node.Kd = 5.0, 10.0, 0.0               # launch/drone.launch:35-37
    node.upper_limit, node.lower_limit, node.windup_limit = 10.0, 0.0, 10.0
    efforts = []
    node.effort_pub = types.SimpleNamespace(publish=lambda m: efforts.append(float(m.data)))
    node.prev_time = _Time(0.0)
    dt = float(dyn.pid.dt)
    ref_T, tick = [], 0
    for e in errors:
        node.setpoint, node.current_state = float(e), 0.0      # error = setpoint - state
        for _ in range(sp.pid_ticks):
            tick += 1
            clock[0] = tick * dt
            node.output()
        ref_T.append(efforts[-1])
    out = dict(pid_error=errors, pid_thrust_ref=np.asarray(ref_T, np.float64), pid_thrust_oracle=np.asarray(thrusts, np.float32),
               pid_dt=np.float32(dt), pid_i0=np.float64(sp.mass * sp.g / sp.pid_ki))

    # ---- (2) the attitude moment for pure pitch states -------------------------------------------------------------------
    ctrl = att_mod.AttitudeController()
    n = 400
    th, th_sp, om = rng.uniform(-0.45, 0.45, n), rng.uniform(-0.3731, 0.3731, n), rng.uniform(-3, 3, n)
    M = np.zeros((n, 3))
    for i in range(n):
        ctrl.state = att_mod.StateMsg(roll=0.0, pitch=float(th_sp[i]), yaw_rate=0.0)
        ctrl.odometry = types.SimpleNamespace(orientation=np.array([0.0, np.sin(th[i] / 2), 0.0, np.cos(th[i] / 2)]),
                                              angular_velocity=np.array([0.0, om[i], 0.0]))
        M[i] = ctrl._compute_desired_moment()
    out.update(att_theta=th, att_theta_sp=th_sp, att_omega=om, att_moment_ref=M,
               att_inertia=np.diag(ctrl.drone.inertia).copy(), att_mass=np.float64(ctrl.drone.mass))
    np.savez_compressed(GOLDEN / "second_order.npz", **out)
    print("second_order:", len(errors), "pid samples; max |T_ref - T_oracle| =",
          float(np.abs(out["pid_thrust_ref"] - out["pid_thrust_oracle"]).max()))


class _StepFeeder:
    """np.random for the UNMODIFIED trainer loop: per agent step the reference draws uniform (explore), randint (random
    action) inside guess() and uniform (table pick, ignored: quirk Q1) inside update() -- words x, y, z of the step draw."""

    def __init__(self, seed):
        self.seed, self.calls = seed, 0

    def _word(self, kind):
        t, j = divmod(self.calls, 3)
        assert kind == ("u", "i", "u")[j], (kind, j)
        self.calls += 1
        return philox.draws(self.seed, 0, np.asarray([0]), t, philox.PURPOSE_STEP)[j][0]

    def uniform(self, lo=0.0, hi=1.0, size=None):
        return float(philox.uniform01(self._word("u"))) * (hi - lo) + lo

    def randint(self, n):
        return int(philox.random_action(self._word("i")))


def gen_curriculum(ns, seed=4, dtype=np.float32, tag="", **trainer_kw):
    """R14 (+ the transfer order of R13) pinned to the reference itself: Trainer.curriculum_training() runs UNMODIFIED; only
    gym.make (no Gazebo: a stand-in-driven env around the reference TrainingMdp), Trainer.save and Trainer.log (file and
    TensorBoard writers) are replaced."""
    kw = dict(successive_successful_episodes=5, success_rate=0.2, max_num_episodes=40)
    kw.update(trainer_kw)
    import gym
    ctx = dict(t=0, makes=[], rec={k: [] for k in ("obs", "action", "state", "next_state", "code", "done", "reward", "episode", "w")})
    sp = StandInParams()
    idx = np.asarray([0])

    class StandInTrainingEnv:
        """PKG/landing_simulation_env.py:143-282 with the simulator replaced by the analytic stand-in (one env)."""

        def __init__(self, initial_curriculum_step, t_max, f_ag, p_max, z_init):
            assert z_init == sp.z_init and f_ag == F_AG
            self.w = initial_curriculum_step
            self.mdp = ns.mdp.TrainingMdp(initial_curriculum_step, f_ag, t_max, p_max)      # :159-164
            self.dyn = StandInDet(sp, 1)
            self.k, self.state = 0, None

        def reset(self):                                                                      # :167-243
            self.mdp.reset()
            words = philox.draws(seed, 0, idx, ctx["t"], philox.PURPOSE_RESET)
            self.dyn.reset(idx, words[0], words[1], words[2], normal_init=(self.w == 0))
            self.dyn.advance(np.zeros(1, np.float32), idx, hover=True)
            rp, rv, ra, pit, z, c = (x[0] for x in self.dyn.observe(np.zeros(1)))
            self.k = 0
            self.state = self.mdp.discrete_state(_obs(ns, rp, rv, ra, pit, z, c))
            return self.state

        def step(self, action_x, action_y=2):                                                 # :245-282
            act = self.mdp.continuous_action(action_x, action_y)
            self.dyn.advance(np.asarray([act.pitch], np.float32))
            self.k += 1
            rp, rv, ra, pit, z, c = (x[0] for x in self.dyn.observe(np.asarray([self.k])))
            s2 = self.mdp.discrete_state(_obs(ns, rp, rv, ra, pit, z, c))
            info = self.mdp.check()
            reward = self.mdp.reward()
            done = "Termination condition" in info.keys()
            info["Current reward"] = reward
            r = ctx["rec"]
            r["obs"].append((rp, rv, ra, pit, z)); r["action"].append(action_x); r["state"].append(state_id(self.state))
            r["next_state"].append(state_id(s2)); r["code"].append(_ref_code(self.mdp, ns)); r["done"].append(int(done))
            r["reward"].append(reward); r["episode"].append(ctx["trainer"]._current_episode); r["w"].append(self.w)
            ctx["t"] += 1
            self.state = s2
            return s2, reward, done, info

        def close(self):
            pass

    def make(name, **k):
        assert name == "Landing-Training-v0"
        ctx["makes"].append((ctx["t"], k["initial_curriculum_step"]))
        return StandInTrainingEnv(**k)

    feeder = _StepFeeder(seed)
    T = ns.trainer.Trainer
    saved = (np.random.uniform, np.random.randint, gym.make, T.save, T.log)
    np.random.uniform, np.random.randint, gym.make = feeder.uniform, feeder.randint, make
    T.save = lambda self: None
    T.log = lambda self, info, clean=False: None
    try:
        trainer = T(save_path=pathlib.Path(tempfile.mkdtemp()), seed=seed, **kw)
        ctx["trainer"] = trainer
        agent = trainer._double_q_learning_agent
        agent.Q_table_a = agent.Q_table_a.astype(dtype)
        agent.Q_table_b = agent.Q_table_b.astype(dtype)
        trainer.curriculum_training()
    finally:
        np.random.uniform, np.random.randint, gym.make, T.save, T.log = saved
    r = ctx["rec"]
    ws = np.asarray(r["w"], np.int32)
    ep = np.asarray(r["episode"], np.int32)
    done = np.asarray(r["done"], np.uint8)
    # how every curriculum step ended: promoted (window cleared) or max_num_episodes reached (window kept)
    ends = []
    for (t_make, w) in ctx["makes"]:
        sel = np.flatnonzero(ws == w)
        n_ep = int(done[sel].sum())
        ends.append((w, int(sel[-1]) + 1, n_ep, int(n_ep < kw["max_num_episodes"]) if w < len(ctx["makes"]) else 0))
    last_sel = np.flatnonzero(ws == ctx["makes"][-1][1])
    codes = np.asarray(r["code"], np.uint8)
    out = dict(
        obs=np.asarray(r["obs"], np.float32), action=np.asarray(r["action"], np.uint8), state=np.asarray(r["state"], np.uint16),
        next_state=np.asarray(r["next_state"], np.uint16), code=codes, done=done, reward=np.asarray(r["reward"], np.float64),
        episode=ep, w=ws, step_end_t=np.asarray([e[1] for e in ends], np.int64), step_episodes=np.asarray([e[2] for e in ends], np.int64),
        qa=agent.Q_table_a, qb=agent.Q_table_b, count=agent.state_action_counter, seed=np.int64(seed),
        successive_successful_episodes=np.int32(kw["successive_successful_episodes"]), success_rate=np.float64(kw["success_rate"]),
        max_num_episodes=np.int64(kw["max_num_episodes"]), window_at_end=np.asarray(list(trainer._successes), np.int8),
    )
    name = f"curriculum_ref{tag}.npz"
    np.savez_compressed(GOLDEN / name, **out)
    print(f"{name}: {len(ws)} steps; per curriculum step (w, t_end, episodes): {[(e[0], e[1], e[2]) for e in ends]}; "
          f"window at end {list(trainer._successes)}")
    return out


def _state_tuple(sid: int):
    th = sid % 7; sid //= 7
    a = sid % 3; sid //= 3
    v = sid % 3; sid //= 3
    p = sid % 3; sid //= 3
    return (sid, p, v, a, th)


def main():
    ns = ref_stubs.install()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    gen_schedules(ns)
    gen_discretise(ns)
    for w in range(5):
        gen_mdp_trace(ns, w)
    # altitude terminals are unreachable with the training defaults (z: 4.0 -> 2.0 m): vary the stand-in
    gen_mdp_trace(ns, 0, 16, sp=StandInParams(z_init=0.9, v_z=-0.4, v_mp=0.4), tag="lowz")
    gen_mdp_trace(ns, 1, 4, sp=StandInParams(z_init=4.6), tag="highz")
    gen_replay(ns, 0, 6000, np.float32)
    gen_replay(ns, 0, 3000, np.float64)
    gen_replay(ns, 0, 8000, np.float32, tag="_ep1950", ep0=1950)
    gen_replay(ns, 0, 4000, np.float64, tag="_ep1950", ep0=1950)
    # later curriculum steps are greedy (eps = 0): start from the committed tables so the policy is non-trivial
    agent = ns.dql.DoubleQLearningAgent.load()
    q0 = (agent.Q_table_a, agent.Q_table_b)
    gen_replay(ns, 2, 3000, np.float32, q_init=q0)
    gen_replay(ns, 4, 3000, np.float32, q_init=q0)
    gen_curriculum(ns)
    gen_sim_trace(ns)
    gen_sim2d_trace(ns)
    gen_kalman()
    gen_second_order()
    total = sum(f.stat().st_size for f in GOLDEN.glob("*.npz"))
    print("golden bytes:", total)


if __name__ == "__main__":
    main()
