"""GPU parity (-m gpu): the CUDA path, called through the C-ABI, against
  (a) the committed fixtures generated from the UNMODIFIED reference (tests/golden/), and
  (b) the CPU oracle (oracle/) on the same seeded inputs.
Bit-exact for observations (same fp32 operations), state indices, actions, terminal flags, rewards
(float64 ==) and -- with float32 tables -- the Q tables and counts.  Float64-table fixtures: Q within
1e-6 relative (north_star tolerance).  Nothing here reads /root/reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import philox
from oracle.agent_oracle import AgentOracle
from oracle.dynamics import StandInParams, det_normal, det_sincos_turns, standin_f64
from oracle.loop import PopulationOracle, TrainerParams, eval_episode
from oracle.mdp_oracle import MdpParams

NO_PROMOTION = dict(success_rate=2.0, max_num_episodes=10 ** 12)


def _engine(P, n_p, **kw):
    from dql_multirotor_landing_b200 import constants as K
    from dql_multirotor_landing_b200.engine import Engine
    tp = K.TrainerParameters(**kw.pop("tp", {}))
    dp = K.DynamicsParameters(**kw.pop("dp", {}))
    return Engine(P, n_p, tp=tp, dp=dp, **kw)


def test_fast_division_is_exact_for_every_fp32_numerator():
    """div_f32_by_const (3 instructions) == IEEE float64 division for all 2^32 fp32 numerators, both divisors."""
    eng = _engine(1, 1, threads_per_block=32)
    assert eng.selftest_division() == 0
    print("one-correction-step variant mismatches:", eng.selftest_one_step_mismatches)


@pytest.mark.parametrize("name", ["w0", "w1", "w2", "w3", "w4", "lowz", "highz"])
def test_mdp_trace_forced_actions(golden_dir, name):
    """R3-R8: forced action sequences; every integer output and the float64 reward equal the reference's."""
    g = np.load(golden_dir / f"mdp_trace_{name}.npz")
    w = int(g["w"])
    rows = np.nonzero(g["action"] != 255)[0]
    acts = g["action"][rows].astype(np.int8)
    eng = _engine(1, 1, threads_per_block=32, seeds=[int(g["seed"])], tp=NO_PROMOTION,
                  dp=dict(z_init=float(g["z_init"]), v_z_train=float(g["v_z"]), v_mp=float(g["v_mp"])))
    eng.reset(w)
    tr = eng.train(len(acts), trace=True, action_override=acts.reshape(-1, 1))
    eng.check_errors()
    assert np.array_equal(tr["obs"][:, 0].view(np.uint32), g["obs"][rows].view(np.uint32))
    assert np.array_equal(tr["action"][:, 0], g["action"][rows])
    assert np.array_equal(tr["next_state"][:, 0].astype(np.uint16), g["state"][rows])
    assert np.array_equal(tr["code"][:, 0], g["code"][rows])
    assert np.array_equal(tr["done"][:, 0], g["done"][rows])
    assert np.array_equal(tr["contact"][:, 0], g["contact"][rows])
    assert np.array_equal(tr["reward"][:, 0], g["reward"][rows])          # float64, bit for bit
    assert np.array_equal(tr["episode"][:, 0], g["episode"][rows])


@pytest.mark.parametrize("name", ["replay_w0_float32", "replay_w0_float32_ep1950", "replay_w2_float32", "replay_w4_float32"])
def test_replay_single_env_bit_exact(golden_dir, name):
    """Config 1: the reference trainer loop (guess/update/alpha/eps + TrainingMdp) on one env, float32 tables
    (NumPy >= 2 / NEP 50 arithmetic): identical actions, indices, flags, rewards, tables and counts."""
    g = np.load(golden_dir / f"{name}.npz")
    w, n = int(g["w"]), len(g["action"])
    eng = _engine(1, 1, threads_per_block=32, seeds=[int(g["seed"])], tp=NO_PROMOTION)
    eng.reset(w)
    eng.set_tables(0, g["qa0"], g["qb0"], np.zeros_like(g["count"]))
    eng.set_episode_index(int(g["ep0"]))
    tr = eng.train(n, trace=True)
    eng.check_errors()
    for key, col in (("action", "action"), ("state", "state"), ("next_state", "next_state"), ("code", "code"), ("done", "done")):
        assert np.array_equal(tr[key][:, 0].astype(np.int64), g[col].astype(np.int64)), key
    assert np.array_equal(tr["obs"][:, 0].view(np.uint32), g["obs"].view(np.uint32))
    assert np.array_equal(tr["reward"][:, 0], g["reward"])
    assert np.array_equal(tr["episode"][:, 0], g["episode"])
    qa, qb, cnt = eng.get_tables(0, np.float32)
    assert np.array_equal(qa.view(np.uint32), g["qa"].view(np.uint32))
    assert np.array_equal(qb.view(np.uint32), g["qb"].view(np.uint32))
    assert np.array_equal(cnt, g["count"])
    ps = eng.population_state()[0]
    assert ps["total_steps"] == n and ps["total_episodes"] == int(g["done"].sum())
    assert ps["total_successes"] == int((g["code"][g["done"] == 1] == 2).sum())


def test_replay_float64_reference_tolerance(golden_dir):
    """Same loop with the reference's default float64 tables: pure exploration (eps = 1) so the trajectory
    cannot depend on Q; device fp32 tables within 1e-6 relative of the float64 ones, counts identical."""
    g = np.load(golden_dir / "replay_w0_float64.npz")
    n = len(g["action"])
    eng = _engine(1, 1, threads_per_block=32, seeds=[int(g["seed"])], tp=NO_PROMOTION)
    eng.reset(0)
    tr = eng.train(n, trace=True)
    assert np.array_equal(tr["action"][:, 0], g["action"]) and np.array_equal(tr["done"][:, 0], g["done"])
    assert np.array_equal(tr["next_state"][:, 0].astype(np.uint16), g["next_state"])
    qa, _, cnt = eng.get_tables(0)
    assert np.array_equal(cnt, g["count"])
    np.testing.assert_allclose(qa, g["qa"], rtol=1e-6, atol=1e-6 * np.abs(g["qa"]).max())


@pytest.mark.parametrize("tpb,n_envs,axes", [(64, 70, "xx"), (256, 300, "xx"), (128, 130, "xy")])
def test_multi_env_population_vs_oracle(tpb, n_envs, axes):
    """Batched semantics S1 (oracle/loop.py): several envs share a table pair; ragged env count; two
    populations with different seeds and platform speeds.  Traces, tables, counts, counters identical.
    axes = "xy": the second population is a y-axis agent (roll; a = -g tan(angle), BASELINE config 4)."""
    steps = 120
    seeds, v_mp = [42, 7], [1.6, 0.8]
    g = [9.81 if a == "x" else -9.81 for a in axes]
    r_mp = [2.0, 2.0] if axes == "xx" else [3.0, 1.5]          # "xy": the two axes of the reference's eight trajectory
    eng = _engine(2, n_envs, threads_per_block=tpb, seeds=seeds, v_mp=v_mp, r_mp=r_mp, axes=list(axes), tp=NO_PROMOTION)
    eng.reset(0)
    tr = eng.train(steps, trace=True)
    eng.check_errors()
    ps = eng.population_state()
    for p in range(2):
        pop = PopulationOracle(n_envs, seed=seeds[p], population=p, w0=0, dtype=np.float32,
                               tp=TrainerParams(**NO_PROMOTION), sp=StandInParams(v_mp=v_mp[p], r_mp=r_mp[p], g=g[p]))
        sl = slice(p * n_envs, (p + 1) * n_envs)
        for t in range(steps):
            o = pop.step()
            assert np.array_equal(tr["obs"][t, sl].view(np.uint32), o["obs"].view(np.uint32)), (p, t)
            for key in ("action", "code", "done"):
                assert np.array_equal(tr[key][t, sl], o[key]), (p, t, key)
            assert np.array_equal(tr["next_state"][t, sl].astype(np.uint16), o["next_state"]), (p, t)
            assert np.array_equal(tr["reward"][t, sl], o["reward"]), (p, t)
        qa, qb, cnt = eng.get_tables(p, np.float32)
        assert np.array_equal(cnt, pop.agent.count)
        assert np.array_equal(qa.view(np.uint32), pop.agent.qa.view(np.uint32))
        assert ps[p]["total_episodes"] == pop.total_episodes and ps[p]["total_successes"] == pop.total_successes
        assert list(ps[p]["termination_hist"]) == list(pop.term_hist)
        assert ps[p]["window_sum"] == sum(pop.window) and ps[p]["window_count"] == len(pop.window)


def test_observation_noise_option_vs_oracle():
    """Gaussian noise on the observed relative position / velocity (PKG/observation_utils.py:127-129, the
    manager_node defaults 0.25 m / 0.1 m/s).  The MDP sees the noisy values (states, fly-zone exits, shaping rewards), the
    physical state stays exact; with promotions (w > 0: no exploration draws, the noise still needs the Philox call)."""
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90)
    noise = dict(noise_pos_sd=0.25, noise_vel_sd=0.1)
    n_envs, steps = 70, 260
    eng = _engine(1, n_envs, threads_per_block=64, seeds=[3], tp=kw, dp=noise)
    eng.reset(0)
    tr = eng.train(steps, trace=True)
    eng.check_errors()
    pop = PopulationOracle(n_envs, seed=3, population=0, w0=0, dtype=np.float32, tp=TrainerParams(**kw), sp=StandInParams(**noise))
    clean = PopulationOracle(n_envs, seed=3, population=0, w0=0, dtype=np.float32, tp=TrainerParams(**kw))
    differs = False
    for t in range(steps):
        o = pop.step()
        assert np.array_equal(tr["obs"][t].view(np.uint32), o["obs"].view(np.uint32)), t
        for key in ("action", "code", "done"):
            assert np.array_equal(tr[key][t], o[key]), (t, key)
        assert np.array_equal(tr["next_state"][t].astype(np.uint16), o["next_state"]), t
        assert np.array_equal(tr["reward"][t], o["reward"]), t
        if t < 5:
            differs = differs or not np.array_equal(clean.step()["obs"], o["obs"])
    assert differs, "the noise must actually change the observations"
    qa, _, cnt = eng.get_tables(0, np.float32)
    assert np.array_equal(cnt, pop.agent.count) and np.array_equal(qa.view(np.uint32), pop.agent.qa.view(np.uint32))
    ps = eng.population_state()[0]
    assert int(ps["working_step"]) == pop.w >= 1
    # the production (non-trace) generic instance
    eng2 = _engine(1, n_envs, threads_per_block=64, seeds=[3], tp=kw, dp=noise)
    eng2.reset(0)
    eng2.train(steps)
    assert torch.equal(eng2.tables, eng.tables) and torch.equal(eng2.env_state, eng.env_state)


@pytest.mark.parametrize("accel_mode", ["kalman_reference", "kalman"])
def test_kalman_acceleration_option_vs_oracle(accel_mode):
    """the MDP sees the reference's acceleration estimate (KalmanFilter3D over a finite difference of the true
    relative velocity, PKG/filters.py:4-80 + PKG/observation_utils.py:134-150) instead of the analytic value; sampled at the
    100 Hz sub-step rate (n_sub = 4), combined with the observation noise, through episode resets and promotions (the
    estimator is never reset).  Bit-exact against oracle KalmanAccel, which tests/test_oracle_golden.py pins to the
    unmodified reference filter."""
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90)
    dyn = dict(accel_mode=accel_mode, n_sub=4, kf_measurement_sd=0.1, noise_pos_sd=0.25, noise_vel_sd=0.1)
    n_envs, steps = 70, 260
    eng = _engine(1, n_envs, threads_per_block=64, seeds=[5], tp=kw, dp=dyn)
    eng.reset(0)
    tr = eng.train(steps, trace=True)
    eng.check_errors()
    pop = PopulationOracle(n_envs, seed=5, population=0, w0=0, dtype=np.float32, tp=TrainerParams(**kw), sp=StandInParams(**dyn))
    exact = PopulationOracle(n_envs, seed=5, population=0, w0=0, dtype=np.float32, tp=TrainerParams(**kw),
                             sp=StandInParams(**{**dyn, "accel_mode": "exact"}))
    differs = False
    for t in range(steps):
        o = pop.step()
        assert np.array_equal(tr["obs"][t].view(np.uint32), o["obs"].view(np.uint32)), t
        for key in ("action", "code", "done"):
            assert np.array_equal(tr[key][t], o[key]), (t, key)
        assert np.array_equal(tr["next_state"][t].astype(np.uint16), o["next_state"]), t
        assert np.array_equal(tr["reward"][t], o["reward"]), t
        if t < 5:
            differs = differs or not np.array_equal(exact.step()["obs"][:, 2], o["obs"][:, 2])
    assert differs, "the estimate must differ from the analytic acceleration"
    qa, _, cnt = eng.get_tables(0, np.float32)
    assert np.array_equal(cnt, pop.agent.count) and np.array_equal(qa.view(np.uint32), pop.agent.qa.view(np.uint32))
    assert int(eng.population_state()[0]["working_step"]) == pop.w >= 1
    # the estimator state itself: {x, P, v_ref, n}
    fs = eng.filter_state.cpu().numpy()
    assert np.array_equal(fs[:, 0].view(np.float32).view(np.uint32), pop.dyn.kf.x.view(np.uint32))
    assert np.array_equal(fs[:, 1].view(np.float32).view(np.uint32), pop.dyn.kf.P.view(np.uint32))
    assert np.array_equal(fs[:, 2].view(np.float32).view(np.uint32), pop.dyn.kf.v_ref.view(np.uint32))
    assert np.array_equal(fs[:, 3].view(np.uint32), pop.dyn.kf.n)
    # the production (non-trace) generic instance, and the un-fused env operators under the same actions
    eng2 = _engine(1, n_envs, threads_per_block=64, seeds=[5], tp=kw, dp=dyn)
    eng2.reset(0)
    eng2.train(steps)
    assert torch.equal(eng2.tables, eng.tables) and torch.equal(eng2.env_state, eng.env_state)
    assert torch.equal(eng2.filter_state, eng.filter_state)


def test_second_order_dynamics_option_vs_oracle():
    """second-order attitude (geometric controller torque on the inertia, PKG/attitude_controller.py:124-156) +
    the vertical PID node (PKG/pid.py:62-104, 10 node iterations per 100 Hz sub-step, Butterworth filter with the reference's
    tap shift) + thrust-coupled horizontal acceleration; altitude is state.  Combined with the Kalman acceleration estimate
    and observation noise, through resets and promotions (the PID memory is never reset).  Bit-exact against the oracle,
    whose controllers tests/test_oracle_golden.py pins to the unmodified reference classes."""
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90)
    dyn = dict(dynamics_model="second_order", n_sub=4, accel_mode="kalman", noise_pos_sd=0.1, noise_vel_sd=0.05)
    n_envs, steps = 70, 260
    eng = _engine(1, n_envs, threads_per_block=64, seeds=[9], tp=kw, dp=dyn)
    eng.reset(0)
    tr = eng.train(steps, trace=True)
    eng.check_errors()
    pop = PopulationOracle(n_envs, seed=9, population=0, w0=0, dtype=np.float32, tp=TrainerParams(**kw), sp=StandInParams(**dyn))
    for t in range(steps):
        o = pop.step()
        assert np.array_equal(tr["obs"][t].view(np.uint32), o["obs"].view(np.uint32)), t
        for key in ("action", "code", "done"):
            assert np.array_equal(tr[key][t], o[key]), (t, key)
        assert np.array_equal(tr["next_state"][t].astype(np.uint16), o["next_state"]), t
        assert np.array_equal(tr["reward"][t], o["reward"]), t
    qa, _, cnt = eng.get_tables(0, np.float32)
    assert np.array_equal(cnt, pop.agent.count) and np.array_equal(qa.view(np.uint32), pop.agent.qa.view(np.uint32))
    assert int(eng.population_state()[0]["working_step"]) == pop.w >= 1
    ds = eng.dynamics_state.cpu().numpy()          # [2][n][4]: {omega, z, v_z, integral}, {e1, f1, f2, f3}
    d = pop.dyn
    for k, ref in enumerate((d.omega, d.z, d.v_z, d.pid.integ)):
        assert np.array_equal(ds[0, :, k].view(np.uint32), ref.view(np.uint32)), k
    for k, ref in enumerate((d.pid.e1, d.pid.f1, d.pid.f2, d.pid.f3)):
        assert np.array_equal(ds[1, :, k].view(np.uint32), ref.view(np.uint32)), k
    assert np.median(np.abs(ds[0, :, 2] + 0.1)) < 0.03     # the PID holds the commanded descent rate while the drone pitches
    # the production (non-trace) generic instance
    eng2 = _engine(1, n_envs, threads_per_block=64, seeds=[9], tp=kw, dp=dyn)
    eng2.reset(0)
    eng2.train(steps)
    assert torch.equal(eng2.tables, eng.tables) and torch.equal(eng2.env_state, eng.env_state)
    assert torch.equal(eng2.dynamics_state, eng.dynamics_state) and torch.equal(eng2.filter_state, eng.filter_state)


def test_second_order_greedy_evaluation_vs_oracle():
    """The greedy SimulationMdp path (R15) on the second-order model: v_z set-point -0.4, altitude from the PID loop."""
    from oracle.agent_oracle import state_id
    rng = np.random.default_rng(4)
    lut = rng.integers(0, 3, size=945).astype(np.uint8)
    dyn = dict(dynamics_model="second_order", n_sub=4)
    eng = _engine(1, 32, threads_per_block=32, seeds=[21], dp=dyn)
    n_ep, trace_steps = 6, 470
    tr = eng.eval_greedy(lut, n_ep, trace_steps=trace_steps)["trace"]
    sp = StandInParams(v_z=-0.4, **dyn)
    for ep in range(n_ep):
        rows = eval_episode(lambda s: int(lut[state_id(s)]), 21, 0, ep, sp)
        n = len(rows) - 1
        obs = np.asarray([r["obs"] for r in rows[1:]], np.float32)
        assert np.array_equal(tr["obs"][:n, ep].view(np.uint32), obs.view(np.uint32)), ep
        assert list(tr["action"][:n, ep]) == [r["action"] for r in rows[1:]]
        assert list(tr["code"][:n, ep]) == [r["code"] for r in rows[1:]]
        assert tr["done"][n - 1, ep] == 1


def test_kalman_acceleration_needs_its_state_buffer():
    """accel_mode != 0 without the estimator buffer is refused (no silent fall-back to the analytic acceleration); the host-buffer
    call carries the buffer through dqlb200_train_host_ext (test_train_host_carries_the_extension_state)."""
    from dql_multirotor_landing_b200 import _ffi
    eng = _engine(1, 32, threads_per_block=32, seeds=[1], dp=dict(accel_mode="kalman", n_sub=4))
    _ffi.check(eng.lib.dqlb200_bind_filter_state(eng.handle, None))
    with pytest.raises(RuntimeError, match="dqlb200_bind_filter_state"):
        eng.reset(0)
    with pytest.raises(RuntimeError, match="dqlb200_bind_filter_state"):
        eng.train(1)
    eng2 = _engine(1, 32, threads_per_block=32, seeds=[1], dp=dict(dynamics_model="second_order", n_sub=4))
    _ffi.check(eng2.lib.dqlb200_bind_dynamics_state(eng2.handle, None))
    with pytest.raises(RuntimeError, match="dqlb200_bind_dynamics_state"):
        eng2.reset(0)
    lut = np.zeros(945, np.uint8)
    with pytest.raises(RuntimeError, match="two-axis evaluator"):          # not silently evaluated on the default model
        eng2.eval_greedy_2d(lut, lut, 4)
    with pytest.raises(RuntimeError, match="pid_ticks"):
        _engine(1, 32, threads_per_block=32, dp=dict(dynamics_model="second_order", n_sub=4, pid_ticks=1))


def test_curriculum_bookkeeping_vs_reference_trainer(golden_dir):
    """R14 + the order of R13 against the reference ITSELF: tests/golden/curriculum_ref.npz is the unmodified
    Trainer.curriculum_training() (PKG/trainer.py:169-245) run through all five curriculum steps -- a promotion, a max-episodes
    advance (window kept), a promotion on the carried-over window, two more max-episodes advances, the transfer after every step
    incl. the last (PKG/double_q_learning.py:77-89, quirk Q7).  One env on the device, chunked launches: every observation,
    action, state, check code and float64 reward, the steps at which the curriculum advanced, and the final tables."""
    g = np.load(golden_dir / "curriculum_ref.npz")
    kw = dict(successive_successful_episodes=int(g["successive_successful_episodes"]), success_rate=float(g["success_rate"]),
              max_num_episodes=int(g["max_num_episodes"]))
    eng = _engine(1, 1, threads_per_block=32, seeds=[int(g["seed"])], tp=kw)
    eng.reset(0)
    n = len(g["action"])
    t0 = 0
    for chunk in (1, 460, 1, 2938, 1761, 5391, 5096, 1, 64):      # launch boundaries right at and around the curriculum advances
        tr = eng.train(chunk, trace=True)
        m = min(chunk, n - t0)
        sl = slice(t0, t0 + m)
        assert np.array_equal(tr["obs"][:m, 0].view(np.uint32), g["obs"][sl].view(np.uint32)), t0
        assert np.array_equal(tr["action"][:m, 0], g["action"][sl]), t0
        assert np.array_equal(tr["state"][:m, 0].astype(np.uint16), g["state"][sl]), t0
        assert np.array_equal(tr["next_state"][:m, 0].astype(np.uint16), g["next_state"][sl]), t0
        assert np.array_equal(tr["code"][:m, 0], g["code"][sl]) and np.array_equal(tr["done"][:m, 0], g["done"][sl]), t0
        assert np.array_equal(tr["reward"][:m, 0], g["reward"][sl]), t0          # float64, bit for bit
        assert np.array_equal(tr["episode"][:m, 0], g["episode"][sl]), t0
        t0 += m
    assert t0 == n
    eng.check_errors()
    ps = eng.population_state()[0]
    assert ps["finished"] == 1 and ps["t"] == n and ps["working_step"] == 4
    assert [int(x) for x in ps["promoted_at"]] == [int(x) for x in g["step_end_t"]]
    assert ps["total_steps"] == n and ps["total_episodes"] == int(g["done"].sum())
    assert ps["window_count"] == len(g["window_at_end"]) and ps["window_sum"] == int(g["window_at_end"].sum())
    qa, qb, cnt = eng.get_tables(0, np.float32)
    assert np.array_equal(qa.view(np.uint32), g["qa"].view(np.uint32)) and np.array_equal(qb.view(np.uint32), g["qb"].view(np.uint32))
    assert np.array_equal(cnt, g["count"])


@pytest.mark.parametrize("mode", ["reference", "paper"])
def test_curriculum_promotion_and_transfer(mode):
    """R13/R14: success window, promotion latch, max_num_episodes advance, transfer (quirk Q7 and the
    'paper' variant), fresh-MDP restart -- run until the last curriculum step ends."""
    kw = dict(success_rate=0.2, successive_successful_episodes=5, max_num_episodes=40, transfer_mode=mode)
    n_envs, steps = 48, 700
    eng = _engine(1, n_envs, threads_per_block=32, seeds=[3], tp=kw)
    eng.reset(0)
    pop = PopulationOracle(n_envs, seed=3, population=0, w0=0, dtype=np.float32, tp=TrainerParams(**kw))
    done_steps = 0
    for chunk in (1, 7, 64, 128, 500):
        tr = eng.train(chunk, trace=True)
        for t in range(chunk):
            if pop.finished:
                break
            o = pop.step()
            assert np.array_equal(tr["action"][t], o["action"]), (done_steps, t)
            assert np.array_equal(tr["next_state"][t].astype(np.uint16), o["next_state"]), (done_steps, t)
            assert np.array_equal(tr["reward"][t], o["reward"]), (done_steps, t)
        done_steps += chunk
    ps = eng.population_state()[0]
    assert pop.finished and ps["finished"] == 1
    assert ps["t"] == pop.t and ps["working_step"] == pop.w
    assert [int(x) for x in ps["promoted_at"]] == [t + 1 for (t, _w, _p) in pop.promotions]
    qa, qb, cnt = eng.get_tables(0, np.float32)
    assert np.array_equal(qa.view(np.uint32), pop.agent.qa.view(np.uint32))
    assert np.array_equal(qb.view(np.uint32), pop.agent.qb.view(np.uint32))
    assert np.array_equal(cnt, pop.agent.count)
    assert ps["total_steps"] == pop.total_steps and ps["total_episodes"] == pop.total_episodes


def test_chunking_and_block_size_invariance():
    """K fused steps == K single-step launches, and the result does not depend on threads_per_block."""
    results = []
    for tpb, chunks in ((256, [96]), (256, [1] * 96), (32, [32, 64]), (128, [96])):
        eng = _engine(3, 200, threads_per_block=tpb, seeds=[1, 2, 3], tp=NO_PROMOTION)
        eng.reset(0)
        for c in chunks:
            eng.train(c)
        torch.cuda.synchronize()
        results.append((eng.env_state.cpu().numpy().copy(), eng.tables.cpu().numpy().copy(), eng.pop_state.cpu().numpy().copy()))
    for r in results[1:]:
        assert np.array_equal(r[0], results[0][0]) and np.array_equal(r[1], results[0][1]) and np.array_equal(r[2], results[0][2])


def test_dynamics_vs_float64_equations():
    """R4: the fp32 stand-in vs the float64 'textbook' form of the same equations, full episodes, <= 1e-5."""
    n_envs, steps, seed = 64, 200, 11
    eng = _engine(1, n_envs, threads_per_block=64, seeds=[seed], tp=NO_PROMOTION)
    eng.reset(0)
    rng = np.random.default_rng(0)
    acts = rng.integers(0, 3, size=(steps, n_envs)).astype(np.int8)
    tr = eng.train(steps, trace=True, action_override=acts)
    mp, sp = MdpParams(), StandInParams()
    w0, w1, w2, _ = philox.draws(seed, 0, np.arange(n_envs), 0, philox.PURPOSE_RESET)
    x_init = np.float32(1.5) * det_normal(w0, w1)
    s0, _ = det_sincos_turns(w2)
    x_d0 = np.float32(2.0) * s0 + np.clip(x_init, -4.5, 4.5)
    worst = 0.0
    for i in range(n_envs):
        first_done = np.nonzero(tr["done"][:, i])[0]
        T = int(first_done[0]) + 1 if first_done.size else steps
        th, sps = 0.0, []
        for a in acts[:T, i]:
            th = min(th + mp.delta_theta, mp.theta_max) if a == 0 else (max(th - mp.delta_theta, -mp.theta_max) if a == 1 else th)
            sps.append(float(np.float32(th)))
        # one hover period precedes the first observation: advance the platform phase by one step first
        ref = standin_f64(sp, float(x_d0[i]), int(w2[i]) + int(eng.cfg.n_sub) * K_DPHASE(eng), sps)
        worst = max(worst, float(np.abs(tr["obs"][:T, i, :5] - ref[:, :5]).max()))
    assert worst <= 1e-5, worst


def K_DPHASE(eng):
    from dql_multirotor_landing_b200 import constants as K
    return K.platform_constants(eng.dp.r_mp, eng.dp.v_mp, eng.mp.f_ag, eng.dp.n_sub)[0]


def test_eval_greedy_fixture_and_oracle(golden_dir):
    """R15 / config 2: greedy SimulationMdp episodes of the committed policy; first episodes against the
    reference-generated fixture, a larger batch against the oracle, landing rate as in SURVEY.md A.4."""
    from dql_multirotor_landing_b200.engine import greedy_policy
    assets = golden_dir.parent.parent / "assets"
    qa, qb = np.load(assets / "Q_table_a.npy"), np.load(assets / "Q_table_b.npy")
    policy = greedy_policy(qa, qb)
    g = np.load(golden_dir / "sim_trace.npz")
    eng = _engine(1, 1, threads_per_block=32, seeds=[int(g["seed"])])
    n_ep = int(g["episode"].max()) + 1
    res = eng.eval_greedy(policy, n_ep, trace_steps=460)
    tr = res["trace"]
    for ep in range(n_ep):
        rows = np.nonzero((g["episode"] == ep) & (g["action"] != 255))[0]
        T = len(rows)
        assert np.array_equal(tr["obs"][:T, ep].view(np.uint32), g["obs"][rows].view(np.uint32)), ep
        assert np.array_equal(tr["action"][:T, ep], g["action"][rows])
        assert np.array_equal(tr["next_state"][:T, ep].astype(np.uint16), g["state"][rows])
        assert np.array_equal(tr["code"][:T, ep], g["code"][rows]) and tr["done"][T - 1, ep] == 1
    # oracle on more episodes
    pol = lambda s: int(np.argmax(np.add(qa[s], qb[s]) / 2))
    n2 = 24
    res2 = eng.eval_greedy(policy, n2, first_episode=100, trace_steps=460)
    hist = np.zeros(9, np.int64)
    steps = 0
    for ep in range(n2):
        rows = eval_episode(pol, int(g["seed"]), 0, 100 + ep, StandInParams(v_z=-0.4))[1:]
        hist[rows[-1]["code"]] += 1
        steps += len(rows)
        assert [r["action"] for r in rows] == list(res2["trace"]["action"][: len(rows), ep])
    assert list(hist) == res2["termination_hist"] and steps == res2["steps"] and res2["episodes"] == n2
    # size-independent property at scale: every episode terminates, landing rate in the survey's band
    big = eng.eval_greedy(policy, 1 << 16)
    assert big["episodes"] == 1 << 16 and sum(big["termination_hist"]) == 1 << 16
    assert big["termination_hist"][0] == 0 and big["termination_hist"][1] == 0
    assert big["termination_hist"][3] / big["episodes"] > 0.85


def _rank_ordered_merge(snap, tabs):
    """What shared_apply_kernel must produce (the replica-merge rule with ranks as replicas, float32, RANK ORDER):
    snap, tabs[r]: uint32 [3][CELLS] (Q_a bits, Q_b bits, count).  Returns (Q_a bits, count)."""
    q_s, c_s = snap[0].view(np.float32), snap[2]
    s = np.zeros(q_s.shape, np.float32)
    total = np.zeros(q_s.shape, np.uint64)
    visitors = np.zeros(q_s.shape, np.int32)
    single = np.zeros(q_s.shape, np.uint32)
    for t in tabs:
        dc = (t[2] - c_s).astype(np.uint32)
        vis = dc != 0
        term = ((t[0].view(np.float32) - q_s).astype(np.float32) * dc.astype(np.float32)).astype(np.float32)
        s = np.where(vis, (s + term).astype(np.float32), s)
        total += dc
        visitors += vis
        single = np.where(vis, t[0], single)
    with np.errstate(invalid="ignore", divide="ignore"):
        mean = (q_s + (s / total.astype(np.float32)).astype(np.float32)).astype(np.float32)
    q = np.where(visitors == 1, single, np.where(visitors > 1, mean.view(np.uint32), snap[0]))
    cnt = np.minimum(c_s.astype(np.uint64) + total, np.uint64(0xFFFFFFFF)).astype(np.uint32)
    return q, cnt


def test_shared_table_mode_merge():
    """Shared-table mode kernels: with one rank the sync leaves the tables bit-identical (== no-collective mode); with two
    ranks (the all-gather is emulated by stacking the two packed buffers: the gloo test covers the collective itself) the
    merged Q is the visit-weighted mean of the ranks' deltas summed in RANK ORDER (bit-exact against the NumPy statement of
    the rule), the counts add up exactly, and repeating the sync from the same inputs gives the same bits."""
    from dql_multirotor_landing_b200.parallel import SharedTableSync
    engs, syncs = [], []
    for r in range(2):
        e = _engine(1, 256, threads_per_block=64, seeds=[10 + r], population_ids=[r], tp=NO_PROMOTION)
        e.reset(0)
        engs.append(e)
        syncs.append(SharedTableSync(e))
    for e in engs:
        e.train(40)
    torch.cuda.synchronize()
    before = [e.tables.cpu().numpy().copy() for e in engs]
    # G = 1: pack + apply without any other rank
    syncs[0].sync()
    torch.cuda.synchronize()
    assert np.array_equal(engs[0].tables.cpu().numpy(), before[0])
    assert np.array_equal(syncs[0].snapshot.cpu().numpy(), before[0])

    def two_rank_sync():
        engs, syncs = [], []
        for r in range(2):      # fresh engines so that both start from the same snapshot
            e = _engine(1, 256, threads_per_block=64, seeds=[10 + r], population_ids=[r], tp=NO_PROMOTION)
            e.reset(0)
            engs.append(e)
            syncs.append(SharedTableSync(e))
            e.train(40)
        torch.cuda.synchronize()
        snap = syncs[0].snapshot.cpu().numpy().view(np.uint32)[0].copy()
        tabs = [e.tables.cpu().numpy().view(np.uint32)[0].copy() for e in engs]
        for s in syncs:
            s.pack()
        gathered = torch.stack([syncs[0].packed, syncs[1].packed]).contiguous()
        for s in syncs:
            s.apply(gathered)
        torch.cuda.synchronize()
        return snap, tabs, [e.tables.cpu().numpy().view(np.uint32).copy() for e in engs]

    snap, tabs, out = two_rank_sync()
    assert np.array_equal(out[0], out[1])                         # ranks agree after the merge
    q, cnt = _rank_ordered_merge(snap, tabs)
    assert np.array_equal(out[0][0, 2], cnt) and np.array_equal(out[0][0, 0], q)
    assert ((tabs[0][2] > 0) & (tabs[1][2] > 0)).sum() > 50       # the test does merge cells both ranks visited
    _, _, again = two_rank_sync()
    assert np.array_equal(again[0], out[0])                       # run-to-run identical


def test_shared_table_counts_are_exact_beyond_2_pow_24():
    """The exchange carries raw 32-bit counts and integer trainer counters: a cell with more than 2^24 visits per sync on each
    rank (where an fp32 transport loses the low bits), a saturating 32-bit count, and pooled episode counters above 2^24."""
    from dql_multirotor_landing_b200 import constants as K
    from dql_multirotor_landing_b200.parallel import SharedTableSync
    kw = dict(success_rate=0.5, successive_successful_episodes=100, max_num_episodes=40_000_001)
    engs, syncs = [], []
    for r in range(3):
        e = _engine(1, 32, threads_per_block=32, seeds=[1], population_ids=[r], tp=kw)
        e.reset(0)
        engs.append(e)
        syncs.append(SharedTableSync(e, pooled_promotion=False))
        syncs[-1].pooled_promote = 10 ** 6                     # never reached: only the episode counter can arm the advance
    visits = [20_000_001, 20_000_003, 16_777_217]
    for r, e in enumerate(engs):
        t = e.tables.cpu().numpy().view(np.uint32).copy()
        t[0, 2, 7] = visits[r]
        t[0, 0, 7] = np.float32(1.0 + r).view(np.uint32)
        t[0, 2, 9] = 0xFFFFFFF0 if r == 0 else 0x20             # saturates
        t[0, 0, 9] = np.float32(-2.0 - r).view(np.uint32)
        t[0, 2, 11] = 5 if r == 1 else 0                        # one visitor: its bits survive
        t[0, 0, 11] = np.float32(0.1).view(np.uint32) if r == 1 else 0
        e.tables.copy_(torch.from_numpy(t.view(np.int32)).to(e.device))
        ps = e.population_state()
        ps["episodes_in_step"] = 13_333_334 + r                 # 40,000,005 pooled: exact only as integers
        e.pop_state.copy_(torch.from_numpy(ps.view(np.uint8).reshape(-1)).to(e.device))
    snap = syncs[0].snapshot.cpu().numpy().view(np.uint32)[0].copy()
    tabs = [e.tables.cpu().numpy().view(np.uint32)[0].copy() for e in engs]
    for s in syncs:
        s.pack()
    gathered = torch.stack([s.packed for s in syncs]).contiguous()
    for s in syncs:
        s.apply(gathered)
    torch.cuda.synchronize()
    out = [e.tables.cpu().numpy().view(np.uint32)[0] for e in engs]
    assert np.array_equal(out[0], out[1]) and np.array_equal(out[0], out[2])
    assert int(out[0][2, 7]) == sum(visits) == 56_777_221 and int(np.float32(sum(visits))) != sum(visits)      # not an fp32 number
    assert int(out[0][2, 9]) == 0xFFFFFFFF
    assert out[0][0, 11] == np.float32(0.1).view(np.uint32) and int(out[0][2, 11]) == 5
    q, cnt = _rank_ordered_merge(snap, tabs)
    assert np.array_equal(out[0][0], q) and np.array_equal(out[0][2], cnt)
    # 13,333,334 + 13,333,335 + 13,333,336 = 40,000,005 >= 40,000,001: the advance is armed; one episode less per rank would not
    assert [int(e.population_state()[0]["pending_advance"]) for e in engs] == [2, 2, 2]
    for r, e in enumerate(engs):
        ps = e.population_state()
        ps["pending_advance"] = 0
        ps["episodes_in_step"] = 13_333_332 + r                 # 39,999,999 pooled (fp32 would round it to 40,000,000)
        e.pop_state.copy_(torch.from_numpy(ps.view(np.uint8).reshape(-1)).to(e.device))
    for s in syncs:
        s.pack()
    gathered = torch.stack([s.packed for s in syncs]).contiguous()
    for s in syncs:
        s.apply(gathered)
    torch.cuda.synchronize()
    assert [int(e.population_state()[0]["pending_advance"]) for e in engs] == [0, 0, 0]


def test_discretisation_edge_probes_on_device(golden_dir):
    """R5 on the DEVICE against the reference: all 17 261 probes of tests/golden/discretise.npz (random values and +-3 ulp
    around every threshold of PKG/mdp.py:149-170, 285-323), working steps 0..4, through every instantiation of
    discretise_cuts the production kernels use (run-time / compile-time constants, bounded / unbounded level loop)."""
    from oracle.agent_oracle import state_id
    g = np.load(golden_dir / "discretise.npz")
    eng = _engine(1, 1, threads_per_block=32)
    obs = np.ascontiguousarray(g["obs"][:, :4], np.float32)
    for w in range(5):
        want_train = np.asarray([state_id(tuple(int(x) for x in row)) for row in g["train"][w]], np.uint16)
        want_sim = np.asarray([state_id(tuple(int(x) for x in row)) for row in g["sim"][w]], np.uint16)
        for variant in range(4):
            got = eng.selftest_discretise(obs, w, variant)
            want = want_train if variant < 2 else want_sim
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (w, variant, obs[bad[:4]], got[bad[:4]], want[bad[:4]])


def test_shared_table_mode_with_replicas_and_pooled_promotion():
    """BASELINE config 5 shared-table mode on top of replica-merge mode: two ranks (emulated by two engines on one device, the
    all-gather by stacking their packed buffers; the collective itself is covered by the gloo test), each with R = 3 local
    replicas of ONE agent.  After a sync every copy on both ranks is identical, counts add up, and the promotion is decided
    from the windows of all ranks."""
    import ctypes as C
    from dql_multirotor_landing_b200 import constants as K
    from dql_multirotor_landing_b200.parallel import SharedTableSync
    R, n_r, G = 3, 64, 2
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=10 ** 9)
    engs, syncs = [], []
    for g in range(G):
        e = _engine(R, n_r, threads_per_block=32, seeds=[5] * R, population_ids=[g * R + r for r in range(R)], replicas_per_population=R, tp=kw)
        e.reset(0)
        s = SharedTableSync(e)
        s.pooled_promote = K.promote_threshold(6 * R * G, 0.15)          # as SharedTableSync(pooled_promotion=True) sets it with world = G
        engs.append(e); syncs.append(s)
    lib = engs[0].lib

    def sync_all():
        for e, s in zip(engs, syncs):
            lib.dqlb200_replica_merge(e.handle, e.merge_snapshot.data_ptr(), 0, e._stream())
            s.pack()
        gathered = torch.stack([syncs[0].packed, syncs[1].packed]).contiguous()
        for e, s in zip(engs, syncs):
            s.apply(gathered)
        torch.cuda.synchronize()

    base = [e.tables.cpu().numpy().view(np.uint32).copy() for e in engs]
    promoted_at = None
    for rnd in range(40):
        for e in engs:
            e.train(8)
        torch.cuda.synchronize()
        before = [e.tables.cpu().numpy().view(np.uint32).copy() for e in engs]
        ps_before = [e.population_state() for e in engs]
        sync_all()
        t = [e.tables.cpu().numpy().view(np.uint32) for e in engs]
        for g in range(G):
            assert all(np.array_equal(t[g][0], t[g][r]) for r in range(R))               # local copies agree
            assert np.array_equal(syncs[g].snapshot.cpu().numpy().view(np.uint32)[0], t[g][0])
            assert np.array_equal(engs[g].merge_snapshot.cpu().numpy().view(np.uint32)[0], t[g][0])
        assert np.array_equal(t[0][0], t[1][0])                                          # ranks agree
        # counts: every visit of every replica of every rank since the last sync is in the merged count
        dcount = sum((before[g][r, 2].astype(np.int64) - base[g][r, 2].astype(np.int64)) for g in range(G) for r in range(R))
        assert np.array_equal(t[0][0, 2].astype(np.int64), base[0][0, 2].astype(np.int64) + dcount)
        base = [x.copy() for x in t]
        pend = [int(p["pending_advance"]) for e in engs for p in e.population_state()]
        pooled = sum(int(p["window_sum"]) for ps in ps_before for p in ps)
        alive = all(int(p["pending_advance"]) == 0 and int(p["finished"]) == 0 for ps in ps_before for p in ps)
        want = 1 if (alive and pooled >= syncs[0].pooled_promote) else 0
        if alive:
            assert pend == [want] * (G * R), (rnd, pooled, pend)
        if want and promoted_at is None:
            promoted_at = rnd
    assert promoted_at is not None, "the pooled promotion should fire within the test"
    assert all(int(p["working_step"]) >= 1 for e in engs for p in e.population_state())


def test_full_size_properties():
    """BASELINE configs[4] shard at full size (740 populations x 1434 envs = 1,061,160 envs): properties that do not
    need an oracle run -- every env-step is one counted Q-update, runs are reproducible, and a population's result is
    independent of which other populations share the GPU (bit-identical when run alone)."""
    from dql_multirotor_landing_b200 import parallel
    P, n_p, steps = 740, 1434, 48
    ids = list(range(P))
    seeds, v_mp, aidx = parallel.sweep_axes(ids, 5, [0.4, 0.8, 1.2, 1.6], 2)
    variants = [(0.02949, 0.51), (0.05, 0.6)]

    def run(pop_ids, tpb):
        sel = [ids.index(i) for i in pop_ids]
        e = _engine(len(sel), n_p, threads_per_block=tpb, seeds=[seeds[i] for i in sel], population_ids=[ids[i] for i in sel],
                    v_mp=[v_mp[i] for i in sel], alpha_variants=variants, alpha_index=[aidx[i] for i in sel], tp=NO_PROMOTION)
        e.reset(0)
        e.train(16)
        e.train(steps - 16)
        e.check_errors()
        torch.cuda.synchronize()
        return e.tables.cpu().numpy().view(np.uint32), e.population_state()

    tab, ps = run(ids, 128)
    assert int(ps["total_steps"].sum()) == P * n_p * steps
    counts = tab[:, 2].astype(np.int64).sum(axis=1)
    assert np.array_equal(counts, np.full(P, n_p * steps))                       # one Q-update per env-step
    assert np.array_equal(ps["termination_hist"].sum(axis=1), ps["total_episodes"])
    assert (ps["total_successes"] <= ps["total_episodes"]).all() and ps["total_episodes"].sum() > 0
    assert np.isfinite(tab[:, 0].view(np.float32)).all() and (tab[:, 1] == 0).all()   # table B is never written (quirk Q1)
    tab2, _ = run(ids, 128)
    assert np.array_equal(tab, tab2)                                             # reproducible
    probe = [0, 123, 739]
    tab3, ps3 = run(probe, 64)                                                   # alone, other block size
    for k, p in enumerate(probe):
        assert np.array_equal(tab3[k], tab[p]), p
        assert ps3[k]["total_episodes"] == ps[p]["total_episodes"] and ps3[k]["return_sum"] == ps[p]["return_sum"]


@pytest.mark.parametrize("R,merge_every", [(4, 1), (3, 5)])
def test_replica_merge_mode_vs_oracle(R, merge_every):
    """One agent whose envs are spread over R CTAs (configs 2-3: many envs sharing one Q-table pair): replica tables merged
    every `merge_every` steps, pooled promotion; tables, counts and curriculum progress identical to the oracle."""
    from oracle.loop import ReplicatedPopulationOracle
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90)
    n_r, steps = 40, 360
    eng = _engine(R, n_r, threads_per_block=32, seeds=[5] * R, population_ids=list(range(R)), replicas_per_population=R, tp=kw)
    eng.reset(0)
    ora = ReplicatedPopulationOracle(R, n_r, seed=5, tp=TrainerParams(**kw), merge_every=merge_every)
    eng.train_merged(steps, merge_every)
    eng.check_errors()
    for _ in range(steps):
        ora.step()
    ps = eng.population_state()
    assert max(rep.w for rep in ora.reps) >= 2, "the test should cross at least two promotions"
    for r in range(R):
        qa, qb, cnt = eng.get_tables(r, np.float32)
        rep = ora.reps[r]
        assert np.array_equal(cnt, rep.agent.count), r
        assert np.array_equal(qa.view(np.uint32), rep.agent.qa.view(np.uint32)), r
        assert np.array_equal(qb.view(np.uint32), rep.agent.qb.view(np.uint32)), r
        assert (ps[r]["working_step"], ps[r]["finished"], ps[r]["t"]) == (rep.w, int(rep.finished), rep.t)
        assert ps[r]["total_episodes"] == rep.total_episodes and ps[r]["pending_advance"] == (0 if rep.finished else ora.pending)
    # all replicas agree after the final merge
    t = eng.tables.cpu().numpy()
    assert all(np.array_equal(t[0, 0], t[r, 0]) and np.array_equal(t[0, 2], t[r, 2]) for r in range(R))


@pytest.mark.parametrize("R,n_r,tpb", [(512, 128, 128), (1024, 512, 128)])
def test_replica_merge_at_bench_shapes(R, n_r, tpb):
    """The replica merge at the shapes bench.py runs (config 3: 512 replicas x 128 envs; config 4: 2 agents x 512 replicas x 512
    envs; the 32-warp instance of replica_merge_kernel for R > 128): every replica trains M steps on its own copy (each replica is an
    S1 population, checked against the C oracle at size above), then the device merge must equal the oracle's merge formula
    (oracle/loop.py: ReplicatedPopulationOracle.merge, replica order, float32) evaluated in NumPy on the replicas' tables; twice in a
    row (the second merge starts from the first one's snapshot)."""
    groups = 2 if R == 1024 else 1
    Rg = R // groups
    eng = _engine(R, n_r, threads_per_block=tpb, seeds=[42] * R, population_ids=list(range(R)), replicas_per_population=Rg,
                  axes=(["x"] * Rg + ["y"] * Rg) if groups == 2 else None, tp=dict(success_rate=2.0, max_num_episodes=10 ** 12))
    eng.reset(0)
    eng._ensure_merge_snapshot()
    f32 = np.float32
    for M in (3, 16):
        snap = eng.merge_snapshot.cpu().numpy().view(np.uint32).copy()       # [groups][3][CELLS] as the device holds it
        eng.train(M)
        before = eng.tables.cpu().numpy().view(np.uint32).copy()              # [R][3][CELLS] uint32 words
        eng.replica_merge()
        eng.check_errors()
        after = eng.tables.cpu().numpy().view(np.uint32)
        for g in range(groups):
            t = before[g * Rg:(g + 1) * Rg]
            q = t[:, 0].view(f32)
            c = t[:, 2].astype(np.int64)
            sq, sc = snap[g, 0].view(f32), snap[g, 2].astype(np.int64)
            d = c - sc
            assert (d >= 0).all()
            tot, visitors = d.sum(0), (d > 0).sum(0)
            num = np.zeros_like(sq)
            single = sq.copy()
            for r in range(Rg):                                                # replica order: the float32 sum is order dependent
                hit = d[r] > 0
                contrib = ((q[r] - sq).astype(f32) * d[r].astype(f32)).astype(f32)
                num = np.where(hit, (num + contrib).astype(f32), num)
                single = np.where(hit, q[r], single)
            with np.errstate(invalid="ignore", divide="ignore"):
                mean = (sq + (num / tot.astype(f32)).astype(f32)).astype(f32)
            q_new = np.where(visitors == 1, single, np.where(visitors > 1, mean, sq)).astype(f32)
            c_new = (sc + tot).astype(np.uint32)
            assert (visitors > 1).sum() > 100                                  # the order-dependent path is exercised
            for r in (0, 1, Rg // 2, Rg - 1):
                assert np.array_equal(after[g * Rg + r, 0], q_new.view(np.uint32)), (M, g, r)
                assert np.array_equal(after[g * Rg + r, 2], c_new), (M, g, r)
            assert (after[g * Rg:(g + 1) * Rg, 0] == after[g * Rg, 0]).all() and (after[g * Rg:(g + 1) * Rg, 2] == after[g * Rg, 2]).all()
            assert int(c_new.astype(np.int64).sum() - sc.sum()) == Rg * n_r * M  # every env-step of every replica is in the merged counts
    eng.close()


@pytest.mark.parametrize("P", [3, 21])
def test_train_host_equals_device_resident_training(P):
    """dqlb200_train_host (host buffers, chunk-pipelined copies) leaves exactly the state the device-resident entry point
    leaves: env state, tables and trainer state bit for bit, promotions included (P = 3: fewer populations than chunks)."""
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90)
    a = _engine(P, 70, threads_per_block=64, seeds=list(range(P)), tp=kw)
    b = _engine(P, 70, threads_per_block=64, seeds=list(range(P)), tp=kw)
    a.reset(0)
    b.reset(0)
    env_h = b.env_state.cpu().pin_memory()
    tab_h = b.tables.cpu().pin_memory()
    ps_h = b.pop_state.cpu().pin_memory()
    b.env_state.zero_(); b.tables.zero_(); b.pop_state.zero_()        # the call must not depend on what the staging buffers hold
    for k in (40, 1, 90):
        a.train(k)
        b.train_host(k, env_h, tab_h, ps_h)
    torch.cuda.synchronize()
    assert torch.equal(a.env_state.cpu(), env_h)
    assert torch.equal(a.tables.cpu(), tab_h)
    assert torch.equal(a.pop_state.cpu(), ps_h)
    assert int(a.population_state()["working_step"].max()) >= 1


def test_train_host_carries_the_extension_state():
    """dqlb200_train_host_ext: with the acceleration estimator and the second-order model switched on, their per-env state
    travels beside the env state (one block / one two-plane copy per chunk) and the call leaves exactly what the device-resident
    entry point leaves; the plain entry point refuses such a configuration instead of stepping with a default model."""
    from dql_multirotor_landing_b200 import _ffi
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90)
    dp = dict(accel_mode="kalman", dynamics_model="second_order", n_sub=4, noise_pos_sd=0.02, noise_vel_sd=0.05)
    P = 11
    a = _engine(P, 70, threads_per_block=64, seeds=list(range(P)), tp=kw, dp=dp)
    b = _engine(P, 70, threads_per_block=64, seeds=list(range(P)), tp=kw, dp=dp)
    a.reset(0)
    b.reset(0)
    env_h, tab_h, ps_h = b.env_state.cpu().pin_memory(), b.tables.cpu().pin_memory(), b.pop_state.cpu().pin_memory()
    fs_h, ds_h = b.filter_state.cpu().pin_memory(), b.dynamics_state.cpu().pin_memory()
    with pytest.raises(_ffi.Dqlb200Error, match="must travel too"):
        b.train_host(1, env_h, tab_h, ps_h)
    for t in (b.env_state, b.tables, b.pop_state, b.filter_state, b.dynamics_state):
        t.zero_()                                                      # the call must not depend on what the staging buffers hold
    for k in (40, 1, 90):
        a.train(k)
        b.train_host(k, env_h, tab_h, ps_h, filter_state_host=fs_h, dynamics_state_host=ds_h)
    torch.cuda.synchronize()
    assert torch.equal(a.env_state.cpu(), env_h) and torch.equal(a.tables.cpu(), tab_h) and torch.equal(a.pop_state.cpu(), ps_h)
    assert torch.equal(a.filter_state.cpu(), fs_h) and torch.equal(a.dynamics_state.cpu(), ds_h)
    assert float(fs_h.abs().sum()) > 0 and int(a.population_state()["total_episodes"].sum()) > 0


def test_train_host_with_partial_table_levels():
    """table_levels = L: only levels 0 .. L-1 of the tables travel.  Without promotions the result equals the full transfer
    bit for bit (the rows that stay behind on the host are untouched); a working step beyond L is refused up front, a promotion
    beyond L inside the call is reported."""
    from dql_multirotor_landing_b200 import _ffi
    P = 5
    a = _engine(P, 70, threads_per_block=64, seeds=list(range(P)), tp=NO_PROMOTION)
    b = _engine(P, 70, threads_per_block=64, seeds=list(range(P)), tp=NO_PROMOTION)
    rng = np.random.default_rng(0)
    t0 = rng.standard_normal((P, 3, 2835)).astype(np.float32).view(np.int32)      # non-zero rows at every level
    t0[:, 2] = rng.integers(0, 2000, size=(P, 2835))
    for e in (a, b):
        e.tables.copy_(torch.from_numpy(t0).to(e.device))
        e.reset(0)
    env_h, tab_h, ps_h = b.env_state.cpu().pin_memory(), b.tables.cpu().pin_memory(), b.pop_state.cpu().pin_memory()
    b.env_state.zero_(); b.tables.zero_(); b.pop_state.zero_()
    for k in (30, 1, 50):
        a.train(k)
        b.train_host(k, env_h, tab_h, ps_h, table_levels=2)
    torch.cuda.synchronize()
    assert torch.equal(a.env_state.cpu(), env_h) and torch.equal(a.tables.cpu(), tab_h) and torch.equal(a.pop_state.cpu(), ps_h)
    assert not np.array_equal(tab_h.numpy()[:, 0, :567], t0[:, 0, :567])                # level 0 was trained ...
    assert np.array_equal(tab_h.numpy()[:, :, 2 * 567:], t0[:, :, 2 * 567:])             # ... the levels above L never moved
    # promotions inside the call beyond the transferred levels are reported, a working step beyond them is refused
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=60)
    c = _engine(2, 70, threads_per_block=64, seeds=[1, 2], tp=kw)
    c.reset(0)
    env_h, tab_h, ps_h = c.env_state.cpu().pin_memory(), c.tables.cpu().pin_memory(), c.pop_state.cpu().pin_memory()
    with pytest.raises(_ffi.Dqlb200Error, match="promoted beyond the transferred table levels"):
        for _ in range(20):
            c.train_host(64, env_h, tab_h, ps_h, table_levels=1)
    assert int(np.frombuffer(ps_h.numpy().tobytes(), dtype=c.population_state().dtype)["working_step"].max()) >= 1
    with pytest.raises(_ffi.Dqlb200Error, match="smaller than the live levels"):
        c.train_host(1, env_h, tab_h, ps_h, table_levels=1)


@pytest.mark.parametrize("case", ["reference", "xy", "eight", "ywrong"])
def test_two_axis_greedy_evaluation(golden_dir, case):
    """eval2d_kernel against (a) the fixture the unmodified reference SimulationMdp produced on the two-axis
    stand-in (observations, both states, codes bit-exact) and (b) the oracle for more episodes (termination histogram)."""
    from dql_multirotor_landing_b200 import constants as K
    from oracle.dynamics import sim2d_cases
    from oracle.loop import eval_episode_2d
    g = np.load(golden_dir / "sim2d_trace.npz")
    cases = sim2d_cases()
    p2, ci = cases[case], list(cases).index(case)
    lut_x, lut_y = g["lut_x"], g[f"{case}_lut_y"]
    ta = K.TwoAxisParameters(trajectory=p2.trajectory, r_x=p2.base.r_mp, v_x=p2.base.v_mp, r_y=p2.r_y, v_y=p2.v_y,
                             y_action_enabled=p2.y_action_enabled, y_init_enabled=p2.y_init_enabled)
    eng = _engine(1, 1, threads_per_block=32)
    n_fix = int(g[f"{case}_episode"].max()) + 1
    res = eng.eval_greedy_2d(lut_x, lut_y, n_fix, two_axis=ta, seed=int(g["seed"]), stream_id=ci, trace_steps=460)
    tr = res["trace"]
    for ep in range(n_fix):
        sel = np.nonzero(g[f"{case}_episode"] == ep)[0][1:]           # row 0 of an episode is the reset
        n = len(sel)
        assert np.array_equal(tr["obs"][:n, ep].view(np.uint32), g[f"{case}_obs"][sel].view(np.uint32))
        for k in ("action_x", "action_y", "code", "done", "contact"):
            assert np.array_equal(tr[k][:n, ep], g[f"{case}_{k}"][sel]), k
        assert np.array_equal(tr["state_x"][:n, ep].astype(np.uint16), g[f"{case}_state_x"][sel])
        assert np.array_equal(tr["state_y"][:n, ep].astype(np.uint16), g[f"{case}_state_y"][sel])
        assert tr["done"][n - 1, ep] == 1
    # more episodes, statistics only
    n2 = 48
    res2 = eng.eval_greedy_2d(lut_x, lut_y, n2, two_axis=ta, seed=77, stream_id=3, first_episode=1000)
    hist, steps = np.zeros(9, np.int64), 0
    for ep in range(n2):
        rows = eval_episode_2d(lut_x, lut_y, 77, 3, 1000 + ep, p2)[1:]
        hist[rows[-1]["code"]] += 1
        steps += len(rows)
    assert res2["episodes"] == n2 and res2["steps"] == steps and res2["termination_hist"] == list(hist)


def test_train_merged_graph_replay_equals_python_loop():
    """dqlb200_train_merged (one captured CUDA graph replayed from C, remainder launched directly) leaves exactly the state
    the Python loop over dqlb200_train / dqlb200_replica_merge leaves, promotions included."""
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90)
    out = []
    for graph in (True, False):
        e = _engine(4, 40, threads_per_block=32, seeds=[5] * 4, population_ids=list(range(4)), replicas_per_population=4, tp=kw)
        e.reset(0)
        e.train_merged(203, 5, graph=graph)          # 40 graph replays + a remainder of 3 steps
        e.train_merged(60, 1, graph=graph)           # a different interval re-captures the graph
        e.check_errors()
        out.append((e.tables.cpu(), e.env_state.cpu(), e.pop_state.cpu(), e.merge_snapshot.cpu()))
        assert int(e.population_state()["working_step"].max()) >= 1
    for a, b in zip(*out):
        assert torch.equal(a, b)


@pytest.mark.parametrize("R,n_r,tpb,merge_every,default_cfg", [(8, 128, 128, 1, True), (8, 200, 128, 3, True), (24, 64, 32, 1, False), (5, 300, 128, 2, False)])
def test_train_merged_equals_two_launch_sequence(R, n_r, tpb, merge_every, default_cfg):
    """dqlb200_train_merged (graphs of (train, merge) pairs replayed from C) must leave exactly the state of the two-launch sequence
    train -> replica_merge driven from Python (which the oracle tests pin): tables, env state, trainer state and merge snapshot, for
    TWO agents side by side, through pooled promotions and a max-episodes advance, ragged last slots, 32- and 128-thread blocks, the
    production and the generic instance, and calls whose length is not a multiple of the merge interval.  (Any other way of running
    the merged loop -- round 2 tried one cooperative launch, profiles/r02_coop_merge_rejected.txt -- has to pass this test.)"""
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=60)
    from dql_multirotor_landing_b200 import constants as K
    mp = None if default_cfg else K.MdpParameters(w_v=-12.0)      # any non-default constant selects the generic instance
    out = []
    for graph in (True, False):
        e = _engine(2 * R, n_r, threads_per_block=tpb, seeds=[5] * R + [9] * R, population_ids=list(range(2 * R)), replicas_per_population=R, tp=kw,
                    **({} if mp is None else {"mp": mp}))
        assert e.lib.dqlb200_uses_default_instance(e.handle) == (1 if default_cfg else 0)
        e.reset(0)
        for steps in (101, 64, 7):
            e.train_merged(steps, merge_every, graph=graph)
        e.check_errors()
        out.append((e.tables.cpu(), e.env_state.cpu(), e.pop_state.cpu(), e.merge_snapshot.cpu()))
        ws = e.population_state()["working_step"]
        assert int(ws.max()) >= 1, "the run should cross a pooled curriculum decision"
    for a, b in zip(*out):
        assert torch.equal(a, b)


def test_default_and_generic_instances_agree():
    """The production instance of train_kernel has the reference-default MDP / dynamics constants compiled in (KDef); the generic
    instance reads them at run time.  dqlb200_create picks the production instance only for a bit-identical configuration; the
    trace instance is always generic, and both leave the same tables and env state."""
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90)
    a = _engine(2, 150, threads_per_block=64, seeds=[1, 2], tp=kw)
    assert a.lib.dqlb200_uses_default_instance(a.handle) == 1
    c = _engine(2, 150, threads_per_block=64, seeds=[1, 2], tp=kw)
    for e, trace in ((a, False), (c, True)):
        e.reset(0)
        e.train(300, trace=trace)
    torch.cuda.synchronize()
    assert torch.equal(a.tables, c.tables) and torch.equal(a.env_state, c.env_state) and torch.equal(a.pop_state, c.pop_state)
    assert int(a.population_state()["working_step"].max()) >= 1
    # populations that fill every slot of the block (192 = 3 x 64) run the full-slot production instance (no `valid` predicate)
    for tpb, n_p in ((64, 192), (128, 256), (32, 32)):
        f = _engine(2, n_p, threads_per_block=tpb, seeds=[1, 2], tp=kw)
        g = _engine(2, n_p, threads_per_block=tpb, seeds=[1, 2], tp=kw)
        assert f.lib.dqlb200_uses_default_instance(f.handle) == 1
        for e, trace in ((f, False), (g, True)):
            e.reset(0)
            e.train(300, trace=trace)
        torch.cuda.synchronize()
        assert torch.equal(f.tables, g.tables) and torch.equal(f.env_state, g.env_state) and torch.equal(f.pop_state, g.pop_state), (tpb, n_p)
        assert int(f.population_state()["working_step"].max()) >= 1
    # per-population constants (platform amplitude / speed, axis) are run-time values: the production instance stays
    d = _engine(1, 8, threads_per_block=32, dp=dict(r_mp=3.0, v_mp=0.8))
    assert d.lib.dqlb200_uses_default_instance(d.handle) == 1
    # anything that changes a compiled-in constant selects the generic instance
    for dp in (dict(c_d=0.25), dict(z_init=3.0), dict(noise_pos_sd=0.1), dict(n_sub=2)):
        e = _engine(1, 8, threads_per_block=32, dp=dp)
        assert e.lib.dqlb200_uses_default_instance(e.handle) == 0, dp


@pytest.mark.parametrize("dp", [{}, dict(n_sub=2, accel_mode="kalman", dynamics_model="second_order")])
def test_no_access_outside_the_bound_buffers(dp):
    """Every borrowed buffer sits between two canary regions: ragged populations (70 envs on 64-thread blocks), several
    populations, resets and promotions, the production / extended kernel variants and the un-fused operators leave the
    canaries untouched (compute-sanitizer is not available on the pool)."""
    import ctypes as C
    from dql_multirotor_landing_b200 import _ffi, constants as K
    kw = dict(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90)
    eng = _engine(3, 70, threads_per_block=64, seeds=[1, 2, 3], tp=kw, dp=dp)
    n, pad, dev = eng.n_total, 4096, eng.device
    sizes = dict(env=int(eng.lib.dqlb200_env_state_bytes(3, 70)), tables=3 * 3 * K.MAX_CELLS * 4, pop=3 * C.sizeof(K.PopulationState), filt=n * 16, dyn=2 * n * 16)
    big = {k: torch.full((pad + ((v + 255) // 256) * 256 + pad,), 0xAB, dtype=torch.uint8, device=dev) for k, v in sizes.items()}
    view = {k: big[k][pad:pad + sizes[k]] for k in sizes}
    for k in ("env", "tables", "pop", "filt", "dyn"):
        view[k].zero_()
    view["filt"].view(torch.int32).view(n, 4)[:, 1] = 0x3F800000
    _ffi.check(eng.lib.dqlb200_bind(eng.handle, view["env"].data_ptr(), view["tables"].data_ptr(), view["pop"].data_ptr()))
    if dp:
        _ffi.check(eng.lib.dqlb200_bind_filter_state(eng.handle, view["filt"].data_ptr()))
        _ffi.check(eng.lib.dqlb200_bind_dynamics_state(eng.handle, view["dyn"].data_ptr()))
    eng.env_state, eng.pop_state = view["env"].view(torch.int32), view["pop"]
    eng.tables = view["tables"].view(torch.int32).view(3, 3, K.MAX_CELLS)
    eng.reset(0)
    eng.train(1)
    eng.train(260)
    eng.train(40, trace=True)
    w = int(eng.population_state()["working_step"].max())
    act, _ = eng.agent_select(w, 7)                       # the un-fused operators on the same buffers
    eng.env_step(w, 7, act, auto_reset=True)
    eng.check_errors()
    torch.cuda.synchronize()
    assert int(eng.population_state()["working_step"].max()) >= 1
    for k, b in big.items():
        assert bool((b[:pad] == 0xAB).all()) and bool((b[pad + sizes[k]:] == 0xAB).all()), k
    assert int(view["env"].view(torch.int32).abs().sum()) != 0


def test_production_instance_vs_c_oracle_at_size():
    """The NON-trace production instances (full-slot and ragged) against the C statement of the batched semantics
    (oracle/c/population.c, itself held equal to oracle/loop.py on the CPU): populations of 1 280 / 1 200 envs, 128-thread blocks,
    ten slots per thread, promotions and transfers on the way -- sizes the Python oracle needs minutes for.  Float32 tables,
    counts and trainer counters identical."""
    from oracle.c_loop import run_population_c
    kw = dict(success_rate=0.3, successive_successful_episodes=50, max_num_episodes=2500)
    for n_envs in (1280, 1200):
        seeds = [11, 12]
        eng = _engine(2, n_envs, threads_per_block=128, seeds=seeds, tp=kw)
        assert eng.lib.dqlb200_uses_default_instance(eng.handle) == 1
        eng.reset(0)
        steps = 400
        eng.train(150)
        eng.train(1)
        eng.train(steps - 151)
        eng.check_errors()
        ps = eng.population_state()
        for p in range(2):
            out = run_population_c(n_envs, steps, seed=seeds[p], population=p, w0=0, tp=TrainerParams(**kw))
            qa, qb, cnt = eng.get_tables(p, np.float32)
            r = out["result"]
            assert np.array_equal(cnt, out["count"]), (n_envs, p)
            assert np.array_equal(qa.view(np.uint32), out["qa"].view(np.uint32)) and np.array_equal(qb.view(np.uint32), out["qb"].view(np.uint32))
            assert (int(ps[p]["working_step"]), int(ps[p]["total_episodes"]), int(ps[p]["total_successes"])) == (r.w, r.total_episodes, r.total_successes)
            assert list(ps[p]["termination_hist"]) == list(r.term_hist) and r.w >= 1


def _run_nccl_sync_check(n_ranks):
    import json
    import pathlib
    import subprocess
    import sys
    root = pathlib.Path(__file__).resolve().parent.parent
    if n_ranks == 1:
        cmd = [sys.executable, str(root / "tools" / "nccl_sync_check.py")]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_ranks}", "--master-addr", "127.0.0.1",
               "--master-port", "29631", str(root / "tools" / "nccl_sync_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]
    rows = [json.loads(l) for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(rows) == n_ranks and all(r["identical"] and r["cells_visited"] > 0 for r in rows), rows


def test_shared_sync_nccl_single_rank():
    """dqlb200_shared_sync_nccl (replica merge -> pack -> ncclAllGather on a raw ncclComm_t -> apply, all under the C-ABI) leaves the
    same tables / trainer states as the Python sequence around a torch all-gather; one rank, so it runs in the 1-GPU suite."""
    _run_nccl_sync_check(1)


def test_shared_sync_nccl_two_ranks():
    """The same with two ranks (two GPUs, real NCCL between them); skipped on a one-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    _run_nccl_sync_check(2)
