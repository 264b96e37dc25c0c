"""GPU tests of the drop-in Python classes (TrainingMdp / SimulationMdp / DoubleQLearningAgent / Trainer): written the
way tests of the reference classes would read, checked against fixtures generated from the unmodified reference."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

F_AG, T_MAX, P_MAX = 22.92, 20, 4.5


def _obs(rel_p=0.0, rel_v=0.0, rel_a=0.0, pitch=0.0, z=3.0, contact=False, rel_p_y=0.0):
    from dql_multirotor_landing_b200.mdp import ContinuousObservation
    from dql_multirotor_landing_b200.msg import Observation
    return ContinuousObservation(Observation(rel_p_x=float(rel_p), rel_v_x=float(rel_v), rel_a_x=float(rel_a), rel_p_y=float(rel_p_y),
                                             contact=bool(contact)), float(pitch), 0.0, float(z))


def test_known_answers_goal_state():
    """MDP held at the origin reaches TERMINAL_SUCCESS at step 23 with r = 12.879109319249501."""
    from dql_multirotor_landing_b200.mdp import TrainingMdp
    m = TrainingMdp(0, F_AG, T_MAX, P_MAX)
    m.reset()
    assert m.discrete_state(_obs()) == (0, 1, 1, 1, 3)
    for k in range(1, 24):
        m.continuous_action(2)
        assert m.discrete_state(_obs()) == (0, 1, 1, 1, 3)
        info = m.check()
        assert m.reward() == 12.879109319249501
        assert ("Termination condition" in info) == (k == 23)
    assert info["Termination condition"] == "SUCCESS: Goal state reached" and info["Number of steps"] == 23
    assert info["Cumulative reward"] == 283.3404050234889


def test_known_answers_timeout_and_discretisation():
    from dql_multirotor_landing_b200.mdp import TrainingMdp
    m = TrainingMdp(0, F_AG, T_MAX, P_MAX)
    m.reset()
    m.discrete_state(_obs(rel_p=3.0))
    rewards, info = [], {}
    for k in range(1, 460):
        m.continuous_action(2)
        assert m.discrete_state(_obs(rel_p=3.0)) == (0, 2, 1, 1, 3)
        info = m.check()
        rewards.append(m.reward())
        assert ("Termination condition" in info) == (k == 459)
    assert rewards[0] == -17.765671273874283 and rewards[1] == -13.402669528673584
    assert info["Termination condition"] == "FAILURE: Maximum episode duration"
    for w, want in ((0, (0, 1, 1, 1, 4)), (1, (1, 0, 1, 1, 4)), (3, (3, 0, 1, 1, 4)), (4, (3, 0, 1, 1, 4))):
        mm = TrainingMdp(w, F_AG, T_MAX, P_MAX)
        mm.reset()
        assert mm.discrete_state(_obs(-1.0, 0.5, 0.1, 0.1)) == want
    mm = TrainingMdp(2, F_AG, T_MAX, P_MAX)
    mm.reset()
    assert mm.discrete_state(_obs(-4.6, 5, 2, 1)) == (0, 0, 2, 2, 6)


def test_value_errors_like_the_reference():
    from dql_multirotor_landing_b200.mdp import TrainingMdp
    m = TrainingMdp(0, F_AG, T_MAX, P_MAX)
    m.reset()
    with pytest.raises(ValueError):
        m.check()                                  # PKG/mdp.py:352-356
    m.discrete_state(_obs())
    with pytest.raises(ValueError):
        m.reward()                                 # PKG/mdp.py:442-447
    with pytest.raises(ValueError):
        m.continuous_action(0, 1)                  # PKG/mdp.py:544-545
    with pytest.raises(ValueError):
        m.discrete_state(_obs(rel_p=float("nan")))  # PKG/mdp.py:170


@pytest.mark.parametrize("name", ["w0", "w3", "lowz", "highz"])
def test_training_mdp_against_reference_fixture(golden_dir, name):
    from dql_multirotor_landing_b200.mdp import TrainingMdp, state_id
    g = np.load(golden_dir / f"mdp_trace_{name}.npz")
    m = TrainingMdp(int(g["w"]), F_AG, T_MAX, P_MAX)
    for i in range(min(len(g["action"]), 700)):
        o = g["obs"][i].astype(np.float64)
        if g["action"][i] == 255:
            m.reset()
            assert state_id(m.discrete_state(_obs(o[0], o[1], o[2], o[3], o[4], g["contact"][i]))) == g["state"][i]
            continue
        assert m.continuous_action(int(g["action"][i])).pitch == g["theta_sp"][i]
        s = m.discrete_state(_obs(o[0], o[1], o[2], o[3], o[4], g["contact"][i]))
        info = m.check()
        r = m.reward()
        assert state_id(s) == g["state"][i] and ("Termination condition" in info) == bool(g["done"][i]), i
        assert r == g["reward"][i], i
        if g["done"][i]:
            from dql_multirotor_landing_b200 import constants as K
            assert info["Termination condition"] == K.TERMINATION_STRINGS[int(g["code"][i])]


def test_simulation_mdp_against_reference_fixture(golden_dir):
    from dql_multirotor_landing_b200.mdp import SimulationMdp, state_id
    g = np.load(golden_dir / "sim_trace.npz")
    m = SimulationMdp(4, F_AG, T_MAX)
    for i in range(len(g["action"])):
        o = g["obs"][i].astype(np.float64)
        if g["action"][i] == 255:
            m.reset()
            sx, sy = m.discrete_state(_obs(o[0], o[1], o[2], o[3], o[4], g["contact"][i]))
            assert state_id(sx) == g["state"][i] and sy == (4, 1, 1, 1, 3)
            continue
        m.continuous_action(int(g["action"][i]), 2)
        sx, _ = m.discrete_state(_obs(o[0], o[1], o[2], o[3], o[4], g["contact"][i]))
        info = m.check()
        assert state_id(sx) == g["state"][i] and ("Termination condition" in info) == bool(g["done"][i]), i


def test_agent_float64_replay_exact(golden_dir, monkeypatch):
    """DoubleQLearningAgent.guess/update with the reference's float64 tables: fed the fixture's transitions and the same
    draws, actions and final tables are identical to the reference's (float64 ==)."""
    from dql_multirotor_landing_b200 import constants as K
    from dql_multirotor_landing_b200.double_q_learning import DoubleQLearningAgent
    from dql_multirotor_landing_b200.mdp import state_tuple
    from oracle import philox
    g = np.load(golden_dir / "replay_w0_float64_ep1950.npz")
    agent = DoubleQLearningAgent(5)
    queue = []
    monkeypatch.setattr(np.random, "uniform", lambda lo=0.0, hi=1.0, size=None: float(philox.uniform01(queue.pop(0))))
    monkeypatch.setattr(np.random, "randint", lambda n: int(philox.random_action(queue.pop(0))))
    idx = np.asarray([0])
    n = 1500
    for t in range(n):
        w = philox.draws(int(g["seed"]), 0, idx, t, philox.PURPOSE_STEP)
        queue[:] = [w[0][0], w[1][0], w[2][0]]
        s, s2 = state_tuple(int(g["state"][t])), state_tuple(int(g["next_state"][t]))
        a = agent.guess(s, K.exploration_rate(int(g["episode"][t]), 0))
        assert a == g["action"][t], t
        agent.update(s + (a,), s2, float(g["alpha"][t]), 0.99, float(g["reward"][t]))
        assert not queue
    # replay the same prefix on the oracle side to get the reference tables after n steps
    from oracle.agent_oracle import AgentOracle
    ref = AgentOracle(5, np.float64)
    for t in range(n):
        s, s2 = state_tuple(int(g["state"][t])), state_tuple(int(g["next_state"][t]))
        ref.update(s + (int(g["action"][t]),), s2, float(g["alpha"][t]), float(g["reward"][t]))
    assert np.array_equal(agent.Q_table_a, ref.qa) and np.array_equal(agent.state_action_counter, ref.count)


def test_agent_save_load_transfer(tmp_path, golden_dir):
    from dql_multirotor_landing_b200.double_q_learning import DoubleQLearningAgent
    agent = DoubleQLearningAgent.load()                       # committed assets
    assert agent.Q_table_a.shape == (5, 3, 3, 3, 7, 3) and agent.Q_table_a.dtype == np.float64
    ref_a = np.load(golden_dir.parent.parent / "assets" / "Q_table_a.npy")
    assert agent.predict((4, 1, 1, 1, 3)) == int(np.argmax((ref_a[4, 1, 1, 1, 3] + agent.Q_table_b[4, 1, 1, 1, 3]) / 2))
    agent.transfer_learning(2, 0.8211253690681617)
    assert np.array_equal(agent.Q_table_a[2], ref_a[1] * 0.8211253690681617)
    agent.transfer_learning(0, 1.0)                           # quirk Q7: slot 0 <- slot -1
    assert np.array_equal(agent.Q_table_a[0], ref_a[4])
    agent.save(tmp_path)
    again = DoubleQLearningAgent.load(tmp_path)
    assert np.array_equal(again.Q_table_a, agent.Q_table_a) and np.array_equal(again.state_action_counter, agent.state_action_counter)
    raw = (tmp_path / "Q_table_a.npy").read_bytes()
    assert raw[:6] == b"\x93NUMPY" and raw[6:8] == b"\x01\x00" and b"'<f8'" in raw[:128] and b"(5, 3, 3, 3, 7, 3)" in raw[:128]


def test_agent_three_curriculum_steps_and_stable_mirrors():
    """(a) an agent with curriculum_steps != 5: transfer_learning(0, r) reads slot -1 = 2 of ITS tables (quirk Q7, PKG/
    double_q_learning.py:77-89), not slot 4 of a 5-step layout; (b) the host mirrors are stable objects: a reference taken once
    shows GPU updates, and counter() reads a count without forcing an upload."""
    from dql_multirotor_landing_b200.double_q_learning import DoubleQLearningAgent
    agent = DoubleQLearningAgent(3)
    rng = np.random.default_rng(3)
    qa0, qb0 = rng.normal(size=(3, 3, 3, 3, 7, 3)), rng.normal(size=(3, 3, 3, 3, 7, 3))
    agent.Q_table_a, agent.Q_table_b = qa0.copy(), qb0.copy()
    agent.transfer_learning(0, 0.75)                           # the reference: Q[0] = Q[0 - 1] * r = Q[2] * r
    assert np.array_equal(agent.Q_table_a[0], qa0[2] * 0.75) and np.array_equal(agent.Q_table_b[0], qb0[2] * 0.75)
    agent.transfer_learning(2, 0.5)
    assert np.array_equal(agent.Q_table_a[2], qa0[1] * 0.5) and np.array_equal(agent.Q_table_a[1], qa0[1])
    held = agent.Q_table_a                                     # a reference kept by the caller
    cnt = agent.state_action_counter
    before = held[1, 1, 1, 1, 3, 2]
    agent.update((1, 1, 1, 1, 3, 2), (1, 1, 0, 1, 3), 0.5, 0.99, 10.0)
    assert agent.counter((1, 1, 1, 1, 3, 2)) == 1.0 and not agent._host_dirty          # no upload pending after a counter read
    assert held is agent.Q_table_a and cnt is agent.state_action_counter              # same objects ...
    assert held[1, 1, 1, 1, 3, 2] != before and cnt[1, 1, 1, 1, 3, 2] == 1.0          # ... showing the GPU update
    held[0, 0, 0, 0, 0, 0] = 7.0                                                       # in-place write after re-reading the attribute
    assert agent.predict((0, 0, 0, 0, 0)) == 0 and agent.Q_table_a[0, 0, 0, 0, 0, 0] == 7.0


def test_trainer_schedules_and_short_curriculum(tmp_path):
    from dql_multirotor_landing_b200.trainer import Trainer
    # max_num_episodes is per env (PKG/trainer.py:190): 64 envs x 2 episodes = 128 pooled episodes force the advance of a step
    tr = Trainer(save_path=tmp_path / "run", successive_successful_episodes=5, success_rate=0.2, max_num_episodes=2,
                 num_envs=64, chunk_steps=32, threads_per_block=64, verbose=False, max_global_steps=4000)
    assert tr.alpha((0, 1, 1, 1, 3, 2)) == 0.02949 and tr.exploration_rate(900, 0) == 0.9175
    assert tr.transfer_learning_ratio(0) == 1.0 and tr.transfer_learning_ratio(1) == 0.8172650252856599
    with pytest.raises(ValueError):
        tr.transfer_learning_ratio(5)
    info = tr.curriculum_training()
    assert "Termination condition" in info and (tmp_path / "run" / "trainer.pickle").exists()
    for name in ("Q_table_a.npy", "Q_table_b.npy", "state_action_count.npy"):
        assert (tmp_path / "run" / name).exists() and (tmp_path / name).exists()      # dual save, PKG/trainer.py:148-152
    agent = tr._double_q_learning_agent
    assert agent.state_action_counter.sum() > 0 and agent.state_action_counter.dtype == np.float64
    ps = tr._engine.population_state()[0]
    assert ps["finished"] == 1 and agent.state_action_counter.sum() == ps["total_steps"]
    import pickle
    again = pickle.load(open(tmp_path / "run" / "trainer.pickle", "rb"))
    assert np.array_equal(again._double_q_learning_agent.Q_table_a, agent.Q_table_a)


def test_trainer_episode_budget_is_per_env(tmp_path):
    """with the reference's defaults the forced advance of a curriculum step must not come before the envs have
    walked the epsilon ramp (episodes 800-2000 of EACH env, PKG/trainer.py:112-126): 256 envs, ~1,000 episodes per env."""
    from dql_multirotor_landing_b200.trainer import Trainer
    tr = Trainer(save_path=tmp_path / "run", num_envs=256, chunk_steps=2048, threads_per_block=128, verbose=False,
                 max_global_steps=65536, success_rate=2.0)
    info = tr.curriculum_training()
    ps = tr._engine.population_state()[0]
    per_env = int(ps["episodes_in_step"]) // 256
    assert per_env > 800, per_env
    assert ps["working_step"] == 0 and ps["finished"] == 0            # 50,000 pooled episodes are long past, 50,000 per env are not
    assert info["Curent episode"] == per_env and info["Exploration rate"] < 1.0
    assert info["Remaining episodes"] == 50000 - per_env + 1
    agent = tr._double_q_learning_agent
    greedy_share = 1.0 - info["Exploration rate"]
    assert greedy_share > 0 and agent.state_action_counter.sum() == ps["total_steps"]


@pytest.mark.parametrize("num_envs,merge_every,steps,shape", [(8192, 4, 1024, (64, 128)), (65536, 1, 256, (128, 512))])
def test_trainer_large_population_uses_replica_merge(tmp_path, num_envs, merge_every, steps, shape):
    """Config 3 through the Trainer facade: one agent, 8,192 envs -> 64 replicas x 128 envs merged every 4 steps; 65,536 envs merged
    after every step (the default) -> 128 replicas x 512 envs in 256-thread blocks (trainer.replica_shape)."""
    from dql_multirotor_landing_b200.trainer import Trainer
    tr = Trainer(save_path=tmp_path / "run", success_rate=0.5, max_num_episodes=30000, num_envs=num_envs, chunk_steps=64,
                 merge_every=merge_every, verbose=False, max_global_steps=steps)
    tr.curriculum_training()
    eng = tr._engine
    assert (eng.R, eng.n_p) == shape
    ps = eng.population_state()
    agent = tr._double_q_learning_agent
    assert agent.state_action_counter.sum() == ps["total_steps"].sum() > 0       # every env-step of every replica is in the merged counts
    assert len(set(int(x) for x in ps["working_step"])) == 1                      # replicas move through the curriculum together
    assert np.isfinite(agent.Q_table_a).all() and np.abs(agent.Q_table_a).max() > 0


@pytest.mark.parametrize("name", ["w0", "w2", "w4", "lowz"])
def test_training_env_gym_surface_against_reference_fixture(golden_dir, name):
    """TrainingLandingEnv (num_envs = 1, the reference's calling convention): reset() / step(a) reproduce the fixture the
    unmodified reference TrainingMdp produced on the stand-in -- state tuples, float64 rewards, done flags, info strings."""
    from dql_multirotor_landing_b200 import constants as K
    from dql_multirotor_landing_b200.landing_simulation_env import make
    from dql_multirotor_landing_b200.mdp import state_tuple
    g = np.load(golden_dir / f"mdp_trace_{name}.npz")
    env = make("Landing-Training-v0", initial_curriculum_step=int(g["w"]), z_init=float(g["z_init"]), seed=int(g["seed"]),
               dynamics=K.DynamicsParameters(z_init=float(g["z_init"]), v_z_train=float(g["v_z"]), v_mp=float(g["v_mp"])))
    n = len(g["action"])
    i = 0
    while i < n:
        assert g["action"][i] == 255                      # a reset row
        s = env.reset()
        assert s == state_tuple(int(g["state"][i]))
        i += 1
        done = False
        while not done:
            s, r, done, info = env.step(int(g["action"][i]))
            assert s == state_tuple(int(g["state"][i])) and r == float(g["reward"][i]) and done == bool(g["done"][i]), i
            assert np.array_equal(env.observation[0].view(np.uint32), g["obs"][i].view(np.uint32))
            if done:
                assert info["Termination condition"] == K.TERMINATION_STRINGS[int(g["code"][i])]
                assert info["Cumulative reward"] == float(g["cum"][i - 1])     # check() runs before reward(): quirk Q12
            else:
                assert "Termination condition" not in info
            assert info["Current reward"] == r
            i += 1
    env.close()


@pytest.mark.parametrize("options", [{}, dict(n_sub=4, accel_mode="kalman_reference", dynamics_model="second_order", noise_pos_sd=0.1)])
def test_batched_env_steps_equal_fused_training_trace(options):
    """The un-fused entry points (dqlb200_env_reset / dqlb200_env_step, auto-reset on) walk exactly the trajectory the fused
    train_kernel walks when it is forced to take the same actions: observations, states, rewards, codes, episode ends --
    with the default model and with the realism options of SURVEY 8f-3 / 8f-4 (estimator + second-order model + noise)."""
    from dql_multirotor_landing_b200 import constants as K
    from dql_multirotor_landing_b200.engine import Engine
    from dql_multirotor_landing_b200.landing_simulation_env import TrainingLandingEnv
    n, steps = 300, 150
    rng = np.random.default_rng(3)
    acts = rng.integers(0, 3, size=(steps, n)).astype(np.int8)
    dp = K.DynamicsParameters(**options)
    eng = Engine(1, n, threads_per_block=128, seeds=[11], dp=dp, tp=K.TrainerParameters(success_rate=2.0, max_num_episodes=10 ** 12))
    eng.reset(0)
    tr = eng.train(steps, trace=True, action_override=acts)
    env = TrainingLandingEnv(0, num_envs=n, seed=11, auto_reset=True, dynamics=dp)
    env.reset()
    for t in range(steps):
        s, r, done, info = env.step(acts[t])
        assert np.array_equal(env.observation.view(np.uint32), tr["obs"][t].view(np.uint32)), t
        assert np.array_equal(r, tr["reward"][t]) and np.array_equal(done, tr["done"][t].astype(bool)), t
        assert np.array_equal(info["code"], tr["code"][t]), t
        ids = (((s[:, 0] * 3 + s[:, 1]) * 3 + s[:, 2]) * 3 + s[:, 3]) * 7 + s[:, 4]
        nd = ~done                     # a finished env already shows the first state of its next episode (auto-reset)
        assert np.array_equal(ids[nd], tr["next_state"][t][nd].astype(np.int64)), t
        if t + 1 < steps:
            assert np.array_equal(ids, tr["state"][t + 1].astype(np.int64)), t
    assert tr["done"].sum() > 50
    torch.cuda.synchronize()
    assert torch.equal(env._engine.env_state, eng.env_state)
    if options:
        assert torch.equal(env._engine.filter_state, eng.filter_state) and torch.equal(env._engine.dynamics_state, eng.dynamics_state)
    env.close()


def test_trainer_with_realism_options(tmp_path):
    """The reference-facing Trainer on the extended kernel variant: estimator + second-order model + noise."""
    from dql_multirotor_landing_b200 import constants as K
    from dql_multirotor_landing_b200.trainer import Trainer
    dp = K.DynamicsParameters(n_sub=4, accel_mode="kalman", dynamics_model="second_order", noise_pos_sd=0.25, noise_vel_sd=0.1)
    tr = Trainer(save_path=tmp_path / "run", successive_successful_episodes=5, success_rate=0.2, max_num_episodes=60, num_envs=64,
                 chunk_steps=32, threads_per_block=64, verbose=False, max_global_steps=2000, dynamics=dp)
    info = tr.curriculum_training()
    assert "Termination condition" in info
    ps = tr._engine.population_state()[0]
    agent = tr._double_q_learning_agent
    assert agent.state_action_counter.sum() == ps["total_steps"] > 0 and np.isfinite(agent.Q_table_a).all()
    assert ps["total_episodes"] > 0 and tr._engine.dynamics_state is not None and tr._engine.filter_state is not None


def test_simulation_env_gym_surface(golden_dir):
    """SimulationLandingEnv: reset() -> (state_x, state_y), step(ax, ay) -> (state_x, state_y, done, info) on the fixture of the
    unmodified reference SimulationMdp (greedy episodes of the committed policy)."""
    from dql_multirotor_landing_b200 import constants as K
    from dql_multirotor_landing_b200.landing_simulation_env import make
    from dql_multirotor_landing_b200.mdp import state_tuple
    g = np.load(golden_dir / "sim_trace.npz")
    # episode ep of the fixture used the reset draws of env index ep, birth 0: three envs replay its first three episodes
    env = make("Landing-Simulation-v0", seed=int(g["seed"]), num_envs=3)
    sx, sy = env.reset()
    for ep in range(3):
        rows = np.nonzero(g["episode"] == ep)[0]
        assert tuple(sx[ep]) == state_tuple(int(g["state"][rows[0]])) and tuple(sy[ep]) == (4, 1, 1, 1, 3)
    lens = [int((g["episode"] == ep).sum()) - 1 for ep in range(3)]
    finished = np.zeros(3, bool)
    for k in range(max(lens)):
        a = np.asarray([int(g["action"][np.nonzero(g["episode"] == ep)[0][min(k + 1, lens[ep])]]) for ep in range(3)], np.int8)
        sx, sy, done, info = env.step(a)
        for ep in range(3):
            if k < lens[ep]:
                row = np.nonzero(g["episode"] == ep)[0][k + 1]
                assert tuple(sx[ep]) == state_tuple(int(g["state"][row])) and bool(done[ep]) == bool(g["done"][row]), (ep, k)
                assert int(info["code"][ep]) == int(g["code"][row])
                if done[ep]:
                    assert info["Termination condition"][ep] == K.TERMINATION_STRINGS[int(g["code"][row])]
    env.close()


def test_unfused_select_step_update_loop_equals_fused_kernel():
    """The reference's loop body as separate device operators -- agent_select (guess + exploration_rate), env_step
    (TrainingLandingEnv.step, auto-reset), agent_update (alpha + update) -- leaves tables, counts and env state bit-identical
    to the fused train_kernel (two populations, ragged env count, epsilon phase of curriculum step 0)."""
    from dql_multirotor_landing_b200 import constants as K
    from dql_multirotor_landing_b200.engine import Engine
    P, n, steps = 2, 150, 120
    kw = dict(success_rate=2.0, max_num_episodes=10 ** 12)
    mk = lambda: Engine(P, n, threads_per_block=64, seeds=[42, 7], v_mp=[1.6, 0.8], tp=K.TrainerParameters(**kw))
    fused, loop = mk(), mk()
    fused.reset(0)
    fused.set_episode_index(1400)          # epsilon = 0.505: exploration and greedy actions both occur
    fused.train(steps)
    loop.env_reset(0, birth=0, fresh_mdp=True)
    loop.set_episode_index(1400)
    for t in range(steps):
        act, st = loop.agent_select(0, t)
        out = loop.env_step(0, t, act, auto_reset=True)
        loop.agent_update(st, act, out["next_state"], out["reward"])
    torch.cuda.synchronize()
    assert torch.equal(fused.tables, loop.tables)
    assert torch.equal(fused.env_state, loop.env_state)
    assert int(fused.tables[:, 2].sum()) == P * n * steps


def test_trainer_log_writes_the_reference_tensorboard_tags(tmp_path, capsys):
    """Trainer.log (PKG/trainer.py:247-303): the same scalar / text tags under <save_path>/logs and the same console layout."""
    pytest.importorskip("tensorboard")
    from tensorboard.backend.event_processing.event_accumulator import EventAccumulator
    from dql_multirotor_landing_b200.trainer import Trainer
    tr = Trainer(save_path=tmp_path / "run", successive_successful_episodes=5, success_rate=0.2, max_num_episodes=60,
                 num_envs=64, chunk_steps=32, threads_per_block=64, verbose=False, max_global_steps=256, tensorboard=True)
    info = tr.curriculum_training()
    tr.log(info)
    out = capsys.readouterr().out
    assert "Curiculum step:" in out and "Current episode:" in out and "Press Ctrl-C to exit..." in out
    acc = EventAccumulator(str(tmp_path / "run" / "logs"))
    acc.Reload()
    tags = set(acc.Tags()["scalars"])
    assert {"Episode/Success Rate", "Episode/Cumulative Reward", "Episode/Exploration Rate", "Episode/Learning Rate",
            "Episode/Mean reward"} <= tags
    assert any("Episode/Termination Condition" in t for t in acc.Tags()["tensors"])


def test_reference_shaped_scripts(tmp_path):
    """scripts/training.py and scripts/simulation.py (the counterparts of the reference's scripts of the same names) run."""
    import pathlib
    import subprocess
    import sys
    root = pathlib.Path(__file__).resolve().parent.parent
    run = lambda *a: subprocess.run([sys.executable, *a], capture_output=True, text=True, timeout=600, cwd=root)
    r = run("scripts/simulation.py", "--episodes", "2")
    assert r.returncode == 0, r.stderr[-1500:]
    assert r.stdout.count("Termination condition") == 2 and "current_episode: 2" in r.stdout
    r = run("scripts/simulation.py", "--episodes", "4096", "--two-axis")
    assert r.returncode == 0, r.stderr[-1500:]
    assert "episodes: 4096" in r.stdout and "Touched platform" in r.stdout
    r = run("scripts/training.py", "--num-envs", "64", "--max-global-steps", "300", "--success-rate", "0.2", "--save-path",
            str(tmp_path / "run"), "--kalman", "--second-order", "--noise")
    assert r.returncode == 0, r.stderr[-1500:]
    assert (tmp_path / "run" / "Q_table_a.npy").exists() and "Termination condition" in r.stdout
