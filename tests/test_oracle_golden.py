"""The CPU restatement (oracle/) against the fixtures generated from the UNMODIFIED reference
(tests/golden/, made by oracle/gen_golden.py).  Integer outputs bit-exact, float64 rewards and
tables compared with == (same IEEE operations in the same order)."""
import numpy as np
import pytest

from oracle import philox
from oracle.agent_oracle import (AgentOracle, alpha_of, exploration_rate, state_from_id, state_id, transfer_ratio)
from oracle.dynamics import StandInParams
from oracle.loop import PopulationOracle, TrainerParams, eval_episode
from oracle.mdp_oracle import MdpParams, TrainingMdpOracle, discretise, linspace7

F_AG, T_MAX, P_MAX = 22.92, 20, 4.5


def test_philox_known_answers():
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
         (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, out in kat:
        got = philox.philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == out


def test_linspace_matches_numpy():
    prm = MdpParams()
    assert linspace7(prm.theta_max, 3) == list(np.linspace(-prm.theta_max, prm.theta_max, 7))
    assert prm.theta_max == np.deg2rad(21.37723) and prm.delta_theta == np.deg2rad(7.12574)


def test_schedules(golden_dir):
    g = np.load(golden_dir / "schedules.npz")
    assert [alpha_of(c) for c in range(len(g["alpha"]))] == list(g["alpha"])
    assert [exploration_rate(e, 0) for e in range(len(g["eps"]))] == list(g["eps"])
    assert [exploration_rate(e, 1) for e in range(len(g["eps1"]))] == list(g["eps1"])
    assert [transfer_ratio(k) for k in range(5)] == list(g["ratios"])
    assert alpha_of(1001) > 0.02949 and alpha_of(1002) == 0.02949


def test_discretise(golden_dir):
    g = np.load(golden_dir / "discretise.npz")
    prm = MdpParams()
    ang = linspace7(prm.theta_max, prm.n_theta)
    obs = g["obs"].astype(np.float64)
    for w in range(5):
        got = np.asarray([discretise(w, prm, P_MAX, ang, *row) for row in obs], np.int8)
        assert np.array_equal(got, g["train"][w])
        assert np.array_equal(got, g["sim"][w])


@pytest.mark.parametrize("name", ["w0", "w1", "w2", "w3", "w4", "lowz", "highz"])
def test_mdp_trace(golden_dir, name):
    g = np.load(golden_dir / f"mdp_trace_{name}.npz")
    w = int(g["w"])
    m = TrainingMdpOracle(w, F_AG, T_MAX, P_MAX)
    for i in range(len(g["action"])):
        o = g["obs"][i].astype(np.float64)
        if g["action"][i] == 255:
            m.reset()
            s = m.observe(o[0], o[1], o[2], o[3], o[4], bool(g["contact"][i]))
            assert state_id(s) == g["state"][i]
            continue
        th = m.act(int(g["action"][i]))
        assert th == g["theta_sp"][i]
        s = m.observe(o[0], o[1], o[2], o[3], o[4], bool(g["contact"][i]))
        code, done = m.check()
        r = m.reward()
        assert (state_id(s), code, int(done)) == (g["state"][i], g["code"][i], g["done"][i]), i
        assert r == g["reward"][i], i
        assert m.cumulative_reward == g["cum"][i]


@pytest.mark.parametrize("name,dtype", [
    ("replay_w0_float32", np.float32), ("replay_w0_float64", np.float64),
    ("replay_w0_float32_ep1950", np.float32), ("replay_w0_float64_ep1950", np.float64),
    ("replay_w2_float32", np.float32), ("replay_w4_float32", np.float32),
])
def test_replay_single_env(golden_dir, name, dtype):
    """PopulationOracle(n_envs=1) == the reference trainer loop, fully independent run
    (own dynamics, own Philox draws): observations, actions, states, flags, rewards and tables identical."""
    g = np.load(golden_dir / f"{name}.npz")
    w = int(g["w"])
    ag = AgentOracle(5, dtype)
    ag.qa[...] = g["qa0"]
    ag.qb[...] = g["qb0"]
    pop = PopulationOracle(1, seed=int(g["seed"]), population=0, w0=w, dtype=dtype, agent=ag,
                           tp=TrainerParams(max_num_episodes=10 ** 9, success_rate=2.0))
    pop.ep[0] = int(g["ep0"])
    n = len(g["action"])
    for t in range(n):
        tr = pop.step()
        assert np.array_equal(tr["obs"][0].view(np.uint32), g["obs"][t].view(np.uint32)), t
        assert (tr["action"][0], tr["state"][0], tr["next_state"][0], tr["code"][0], tr["done"][0]) == \
               (g["action"][t], g["state"][t], g["next_state"][t], g["code"][t], g["done"][t]), t
        assert tr["reward"][0] == g["reward"][t], t
        assert tr["episode"][0] == g["episode"][t]
    assert ag.qa.dtype == g["qa"].dtype
    assert np.array_equal(ag.qa, g["qa"]) and np.array_equal(ag.qb, g["qb"])
    assert np.array_equal(ag.count, g["count"])


def test_curriculum_bookkeeping_matches_reference_trainer(golden_dir):
    """R14 + the transfer order of R13, pinned to the reference itself: curriculum_ref.npz is the UNMODIFIED
    Trainer.curriculum_training() (PKG/trainer.py:169-245, PKG/double_q_learning.py:77-89) run to the end of the last
    curriculum step: a promotion (window cleared), a max-episodes advance (window NOT cleared), a promotion on the window
    carried over, two more max-episodes advances, the transfer applied after every step including the last (quirk Q7),
    epsilon restarting with every step.  PopulationOracle(n_envs=1) must reproduce every step; to keep the CPU suite short the
    Python statement replays the first three curriculum steps (promotion, max-episodes advance, promotion on the carried
    window: 5 200 of 15 649 steps) -- the C statement replays the whole run incl. the final tables (tests/test_oracle_c.py),
    and DQL_FULL_CURRICULUM=1 makes this test do so too."""
    import os
    g = np.load(golden_dir / "curriculum_ref.npz")
    tp = TrainerParams(successive_successful_episodes=int(g["successive_successful_episodes"]), success_rate=float(g["success_rate"]),
                       max_num_episodes=int(g["max_num_episodes"]))
    pop = PopulationOracle(1, seed=int(g["seed"]), population=0, w0=0, dtype=g["qa"].dtype.type, tp=tp)
    full = bool(os.environ.get("DQL_FULL_CURRICULUM"))
    n = len(g["action"]) if full else 5200
    for t in range(n):
        assert not pop.finished and pop.w == g["w"][t], t
        tr = pop.step()
        assert np.array_equal(tr["obs"][0].view(np.uint32), g["obs"][t].view(np.uint32)), t
        assert (tr["action"][0], tr["state"][0], tr["next_state"][0], tr["code"][0], tr["done"][0]) == \
               (g["action"][t], g["state"][t], g["next_state"][t], g["code"][t], g["done"][t]), t
        assert tr["reward"][0] == g["reward"][t], t
        assert tr["episode"][0] == g["episode"][t], t
    k = len(pop.promotions)
    assert k == (5 if full else 3) and pop.finished == full
    assert [t + 1 for (t, _w, _p) in pop.promotions] == list(g["step_end_t"][:k])
    # promoted <=> the step ended before max_num_episodes episodes
    assert [bool(p) for (_t, _w, p) in pop.promotions] == [int(e) < int(g["max_num_episodes"]) for e in g["step_episodes"][:k]]
    if not full:
        return
    assert list(pop.window) == list(g["window_at_end"])
    assert pop.agent.qa.dtype == g["qa"].dtype
    assert np.array_equal(pop.agent.qa, g["qa"]) and np.array_equal(pop.agent.qb, g["qb"])
    assert np.array_equal(pop.agent.count, g["count"])


def test_sim_trace(golden_dir):
    g = np.load(golden_dir / "sim_trace.npz")
    qa, qb = np.load(golden_dir.parent.parent / "assets" / "Q_table_a.npy"), np.load(golden_dir.parent.parent / "assets" / "Q_table_b.npy")
    policy = lambda s: int(np.argmax(np.add(qa[s], qb[s]) / 2))
    rows = []
    n_ep = int(g["episode"].max()) + 1
    for ep in range(n_ep):
        rows += eval_episode(policy, int(g["seed"]), 0, ep, StandInParams(v_z=-0.4))
    assert len(rows) == len(g["action"])
    for i, r in enumerate(rows):
        assert np.array_equal(np.asarray(r["obs"], np.float32).view(np.uint32), g["obs"][i].view(np.uint32)), i
        assert (r["action"], r["state"], r["code"], r["done"], int(r["contact"])) == \
               (g["action"][i], g["state"][i], g["code"][i], g["done"][i], g["contact"][i]), i


def test_state_id_roundtrip():
    for i in range(945):
        assert state_id(state_from_id(i)) == i


@pytest.mark.parametrize("case", ["reference", "xy", "eight", "ywrong"])
def test_two_axis_simulation_oracle_matches_reference_fixture(golden_dir, case):
    """the two-axis restatement (x and y discretisation, contact on both axes, FLYZONE_Y) reproduces what the
    unmodified reference SimulationMdp produced on the same two-axis stand-in trajectories (tests/golden/sim2d_trace.npz)."""
    from oracle.dynamics import sim2d_cases
    from oracle.loop import eval_episode_2d, mirrored_policy
    g = np.load(golden_dir / "sim2d_trace.npz")
    cases = sim2d_cases()
    ci = list(cases).index(case)
    lut_x, lut_y = g["lut_x"], g[f"{case}_lut_y"]
    if case != "ywrong":
        assert np.array_equal(lut_y, mirrored_policy(lut_x))
    keys = ("rel_p_x", "rel_v_x", "rel_a_x", "pitch", "z", "rel_p_y", "rel_v_y", "rel_a_y", "roll")
    ep_ids = g[f"{case}_episode"]
    for ep in np.unique(ep_ids):
        sel = np.nonzero(ep_ids == ep)[0]
        rows = eval_episode_2d(lut_x, lut_y, int(g["seed"]), ci, int(ep), cases[case])
        assert len(rows) == len(sel)
        obs = np.asarray([[r["obs"][k] for k in keys] for r in rows], np.float32)
        assert np.array_equal(obs.view(np.uint32), g[f"{case}_obs"][sel].view(np.uint32))
        for k in ("action_x", "action_y", "state_x", "state_y", "code", "done"):
            assert [r[k] for r in rows] == list(g[f"{case}_{k}"][sel]), k
    assert (g["ywrong_code"][g["ywrong_done"] == 1] == 5).all()        # FLYZONE_Y is covered


@pytest.mark.parametrize("mode,key", [("kalman_reference", "anchor"), ("kalman", "consecutive")])
@pytest.mark.parametrize("sd", [0.1, 0.25])
def test_kalman_acceleration_estimator_matches_reference_filter(golden_dir, mode, key, sd):
    """the fp32 estimator (oracle KalmanAccel == kf_sample of the kernels) against the UNMODIFIED reference
    KalmanFilter3D (PKG/filters.py:4-80, float64) driven with the call protocol of PKG/observation_utils.py:134-150 on the same
    velocity samples (tests/golden/kalman_accel.npz).  Tolerance 2e-5 m/s^2 absolute (measured 8e-6; fp32 vs float64 over 3 364 updates; the
    filter is a contraction, errors do not accumulate), i.e. 1.6e-5 of a_max = 1.28."""
    from oracle.dynamics import KalmanAccel
    g = np.load(golden_dir / "kalman_accel.npz")
    v, ref = g["rel_v"], g[f"{key}_sd{sd}"]
    kf = KalmanAccel(2, mode, g["h"], float(g["q"]), sd ** 2)
    out = np.zeros((len(v), 2), np.float32)
    for k, vk in enumerate(v):
        kf.sample(np.arange(2), np.asarray([vk, -vk], np.float32))
        out[k] = kf.x
    assert out[0, 0] == 0.0 and ref[0, 0] == 0.0            # the first observation reports 0 (observation_utils.py:140-143)
    assert np.array_equal(out[:, 0], -out[:, 1])            # the y channel of the fixture is the negated signal
    assert np.abs(ref[:, 2]).max() == 0.0                   # constant z velocity -> zero acceleration
    err = np.abs(out.astype(np.float64) - ref[:, :2]).max()
    assert err < 2e-5, err
    if key == "anchor":       # quirk Q13: against the growing time base the estimate decays toward the mean acceleration since start
        assert np.abs(ref[-200:, 0]).max() < 0.25 * np.abs(g["consecutive_sd0.1"][-200:, 0]).max()


def test_second_order_model_pins_against_reference_controllers(golden_dir):
    """the two controllers the second-order model restates, against the UNMODIFIED reference classes
    (tests/golden/second_order.npz, made by oracle/gen_golden.py: gen_second_order).
    (1) vertical PID: fp32 VerticalPid (== pid_thrust of the kernels) vs PKG/pid.py PID.output() (float64, Butterworth filter
        of PKG/filters.py:83-108 with its shifted output taps) on the same held errors at the node's 1 kHz: 5e-5 of the thrust
        (measured 1.4e-5 over 16 480 node iterations; the fp32 integral accumulates increments of 1e-4 on 0.67).
    (2) attitude: M_y of AttitudeController._compute_desired_moment() for pure pitch attitudes is -k_R sin(theta - theta_sp)
        - k_w omega, the torque the model applies to the inertia J (1e-12), and the other two moments vanish."""
    from oracle.dynamics import StandInParams, VerticalPid, det_sin_small
    g = np.load(golden_dir / "second_order.npz")
    sp = StandInParams(n_sub=4, dynamics_model="second_order")
    pid = VerticalPid(1, sp, float(g["pid_dt"]) * sp.pid_ticks)
    assert pid.dt == g["pid_dt"] and abs(float(pid.integ[0]) - float(g["pid_i0"])) < 1e-6
    out = np.asarray([pid.thrust(np.arange(1), np.asarray([e], np.float32))[0] for e in g["pid_error"]], np.float32)
    assert np.array_equal(out, g["pid_thrust_oracle"])                  # the fixture's own oracle column is reproducible
    ref = g["pid_thrust_ref"]
    assert ref.min() > 0.0 and ref.max() < 10.0 and np.ptp(ref) > 0.5   # inside the effort limits, and really moving
    assert np.abs(out - ref).max() < 5e-5 * np.abs(ref).max()
    th, th_sp, om, M = g["att_theta"], g["att_theta_sp"], g["att_omega"], g["att_moment_ref"]
    assert np.abs(M[:, 1] - (-sp.k_R * np.sin(th - th_sp) - sp.k_omega * om)).max() < 1e-12
    assert np.abs(M[:, [0, 2]]).max() == 0.0
    assert g["att_inertia"][1] == sp.inertia and float(g["att_mass"]) == sp.mass
    # the kernels' polynomial sine over the range of attitude errors
    x = np.linspace(-0.85, 0.85, 20001).astype(np.float32)
    assert np.abs(det_sin_small(x).astype(np.float64) - np.sin(x.astype(np.float64))).max() < 2e-7


def test_second_order_model_reduces_to_first_order_behaviour():
    """The second-order stand-in settles on the set-point like the first-order lag (tau = k_w / k_R) and holds the commanded
    descent rate: after 3 s at a constant set-point pitch, horizontal acceleration (g tan(theta)) and v_z agree within 2 %."""
    from oracle.dynamics import StandInDet
    a = StandInDet(StandInParams(n_sub=4), 1)
    b = StandInDet(StandInParams(n_sub=4, dynamics_model="second_order"), 1)
    for d in (a, b):
        d.reset(np.arange(1), [123456789], [987654321], [0], normal_init=True)
        d.advance(np.zeros(1, np.float32), hover=True) if d.so else d.advance(np.zeros(1, np.float32))
        for _ in range(69):
            d.advance(np.asarray([0.2], np.float32))
    assert abs(float(b.theta[0]) - 0.2) < 2e-3 and abs(float(a.theta[0]) - 0.2) < 2e-3
    assert abs(float(b.v_z[0]) - (-0.1)) < 2e-3
    acc_a = 9.81 * np.tan(0.2) - 0.2 * float(a.v_d[0])
    assert abs(float(b.a_d[0]) - (9.81 * np.tan(0.2) - 0.2 * float(b.v_d[0]))) < 0.02 * abs(acc_a)
    assert abs(float(b.v_d[0]) - float(a.v_d[0])) < 0.05 * abs(float(a.v_d[0]))
