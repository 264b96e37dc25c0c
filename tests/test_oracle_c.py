"""The plain-C restatement of the numeric core (oracle/c/standin.c) against the NumPy oracle, bit for bit: Philox4x32-10,
the fp32 building blocks, the stand-in dynamics, the acceleration estimator and the second-order model.  Two independent
statements of the same arithmetic (NumPy float32, C with -ffp-contract=off) agreeing on every bit is the basis of the
claim that the CUDA kernels (explicit __fmul_rn / __fadd_rn, --fmad=false) can be bit-exact against the oracle at all."""
import ctypes as C
import pathlib
import subprocess

import numpy as np
import pytest

from oracle import philox
from oracle.c_loop import LoopParams, Params
from oracle.dynamics import (StandInDet, StandInParams, derive, det_log, det_normal_pair, det_sincos_turns, det_tan)

CDIR = pathlib.Path(__file__).resolve().parent.parent / "oracle" / "c"


class State(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("x_d", "v_d", "theta", "a_d")] + [("phase", C.c_uint32)] + [
        (n, C.c_float) for n in ("kf_x", "kf_P", "kf_vref")] + [("kf_n", C.c_uint32)] + [
        (n, C.c_float) for n in ("omega", "z", "v_z", "integ", "e1", "f1", "f2", "f3")]


@pytest.fixture(scope="module")
def lib():
    so = CDIR / "libstandin_oracle.so"
    if not so.exists() or so.stat().st_mtime < (CDIR / "standin.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(CDIR)], check=True, capture_output=True)
    l = C.CDLL(str(so))
    l.oracle_tan.restype = l.oracle_log.restype = C.c_float
    l.oracle_tan.argtypes = l.oracle_log.argtypes = [C.c_float]
    return l


def bits(x):
    return np.asarray(x, np.float32).view(np.uint32)


def test_philox_c_equals_numpy(lib):
    rng = np.random.default_rng(1)
    ctr = rng.integers(0, 2 ** 32, size=(500, 4), dtype=np.uint64).astype(np.uint32)
    keys = rng.integers(0, 2 ** 32, size=(500, 2), dtype=np.uint64).astype(np.uint32)
    out = (C.c_uint32 * 4)()
    for c, k in zip(ctr, keys):
        lib.oracle_philox4x32_10((C.c_uint32 * 4)(*[int(v) for v in c]), C.c_uint32(int(k[0])), C.c_uint32(int(k[1])), out)
        ref = philox.philox4x32_10(c[0:1], c[1:2], c[2:3], c[3:4], int(k[0]), int(k[1]))
        assert list(out) == [int(r[0]) for r in ref]
    # Random123 known answer: counter = key = 0
    lib.oracle_philox4x32_10((C.c_uint32 * 4)(0, 0, 0, 0), C.c_uint32(0), C.c_uint32(0), out)
    assert list(out) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


def test_fp32_building_blocks_c_equal_numpy(lib):
    rng = np.random.default_rng(2)
    phases = np.concatenate([rng.integers(0, 2 ** 32, size=4000, dtype=np.uint64).astype(np.uint32),
                             np.asarray([0, 1, 0x1FFFFFFF, 0x20000000, 0x3FFFFFFF, 0x40000000, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF], np.uint32)])
    s_np, c_np = det_sincos_turns(phases)
    s, c = C.c_float(), C.c_float()
    for i, ph in enumerate(phases):
        lib.oracle_sincos_turns(C.c_uint32(int(ph)), C.byref(s), C.byref(c))
        assert bits(s.value) == bits(s_np[i]) and bits(c.value) == bits(c_np[i]), hex(int(ph))
    x = rng.uniform(-0.45, 0.45, 3000).astype(np.float32)
    assert np.array_equal(bits([lib.oracle_tan(float(v)) for v in x]), bits(det_tan(x)))
    u = ((rng.integers(0, 2 ** 24, size=3000).astype(np.float32) + np.float32(1)) * np.float32(2.0 ** -24)).astype(np.float32)
    assert np.array_equal(bits([lib.oracle_log(float(v)) for v in u]), bits(det_log(u)))
    w = rng.integers(0, 2 ** 32, size=(2000, 2), dtype=np.uint64).astype(np.uint32)
    n0_np, n1_np = det_normal_pair(w[:, 0], w[:, 1])
    n0, n1 = C.c_float(), C.c_float()
    for i in range(len(w)):
        lib.oracle_normal_pair(C.c_uint32(int(w[i, 0])), C.c_uint32(int(w[i, 1])), C.byref(n0), C.byref(n1))
        assert bits(n0.value) == bits(n0_np[i]) and bits(n1.value) == bits(n1_np[i])


def _params(sp: StandInParams, dyn: StandInDet) -> Params:
    d = derive(sp)
    p = Params(h=d.h, half_h2=d.half_h2, k_theta=d.k_theta, g=d.g, c_d=d.c_d, r=d.r, rw=d.rw, rw2=d.rw2, dphase=d.dphase, n_sub=d.n_sub,
               accel_mode={"exact": 0, "kalman_reference": 1, "kalman": 2}[sp.accel_mode], kf_q=np.float32(sp.kf_process_variance),
               kf_r=np.float32(sp.kf_measurement_sd ** 2), second_order=int(dyn.so), pid_ticks=sp.pid_ticks, z_init=d.z_init)
    if dyn.so:
        pid = dyn.pid
        p.att_kr, p.att_kw, p.inv_m, p.inv_mg, p.g_abs = dyn.att_kr, dyn.att_kw, dyn.inv_m, dyn.inv_mg, dyn.g_abs
        p.pid_kp, p.pid_ki, p.pid_lo, p.pid_hi, p.pid_windup, p.pid_dt = pid.kp, pid.ki, pid.lo, pid.hi, pid.windup, pid.dt
        p.bw_inv_denom, p.bw_k2 = pid.inv_denom, pid.k2
    return p


@pytest.mark.parametrize("opts", [dict(), dict(n_sub=4), dict(n_sub=4, accel_mode="kalman_reference"), dict(n_sub=2, accel_mode="kalman"),
                                  dict(n_sub=4, dynamics_model="second_order"),
                                  dict(n_sub=4, dynamics_model="second_order", accel_mode="kalman", g=-9.81)])
def test_standin_rollouts_c_equal_numpy(lib, opts):
    """Several episodes per env with random set-points and teleport resets in between: observations and the complete state
    (body, estimator, PID memory) agree bit for bit after every agent period."""
    sp = StandInParams(**opts)
    n_env, rng = 6, np.random.default_rng(3)
    dyn = StandInDet(sp, n_env)
    p = _params(sp, dyn)
    vz_sp = np.float32(sp.v_z)
    states = [State(kf_P=1.0, z=float(dyn.d.z_init), integ=float(dyn.pid.integ[0]) if dyn.so else 0.0) for _ in range(n_env)]
    idx = np.arange(n_env)
    obs = (C.c_float * 4)()
    for episode in range(3):
        w = rng.integers(0, 2 ** 32, size=(3, n_env), dtype=np.uint64).astype(np.uint32)
        dyn.reset(idx, w[0], w[1], w[2], normal_init=(episode % 2 == 0))
        for i, s in enumerate(states):          # the teleport: body state from the oracle's reset law, memories untouched
            s.x_d, s.v_d, s.theta, s.phase = float(dyn.x_d[i]), 0.0, 0.0, int(dyn.phase[i])
            if dyn.so:
                s.omega, s.z, s.v_z = 0.0, float(dyn.d.z_init), 0.0
        dyn.advance(np.zeros(n_env, np.float32), idx, hover=True)
        for s in states:
            lib.oracle_advance(C.byref(p), C.byref(s), C.c_float(0.0), C.c_float(0.0))
        for t in range(60):
            sps = rng.choice(np.asarray([-0.3731, -0.1244, 0.0, 0.1244, 0.2487, 0.3731], np.float32), n_env)
            dyn.advance(sps, idx)
            rel_p, rel_v, rel_a, pitch, z, _ = dyn.observe(np.full(n_env, t + 1), idx)
            for i, s in enumerate(states):
                lib.oracle_advance(C.byref(p), C.byref(s), C.c_float(float(sps[i])), C.c_float(float(vz_sp)))
                lib.oracle_observe(C.byref(p), C.byref(s), obs)
                assert list(bits(list(obs))) == list(bits([rel_p[i], rel_v[i], rel_a[i], pitch[i]])), (episode, t, i)
                assert bits(s.x_d) == bits(dyn.x_d[i]) and bits(s.v_d) == bits(dyn.v_d[i]) and s.phase == int(dyn.phase[i])
                if dyn.kf is not None:
                    assert bits(s.kf_x) == bits(dyn.kf.x[i]) and bits(s.kf_P) == bits(dyn.kf.P[i]) and s.kf_n == int(dyn.kf.n[i])
                if dyn.so:
                    assert bits(s.z) == bits(dyn.z[i]) == bits(z[i]) and bits(s.v_z) == bits(dyn.v_z[i]) and bits(s.omega) == bits(dyn.omega[i])
                    assert bits(s.integ) == bits(dyn.pid.integ[i]) and bits(s.f3) == bits(dyn.pid.f3[i]) and bits(s.e1) == bits(dyn.pid.e1[i])


@pytest.mark.parametrize("name", ["w0", "w1", "w2", "w3", "w4", "lowz", "highz"])
def test_mdp_c_restatement_replays_the_reference_fixtures(lib, name):
    """oracle/c/mdp.c (discrete_state, check, reward, continuous_action, reset of PKG/mdp.py) on the forced-action traces the
    UNMODIFIED reference TrainingMdp produced (tests/golden/mdp_trace_*.npz): set-points, state ids, result codes, done flags,
    float64 rewards and cumulative rewards identical (==)."""
    golden = pathlib.Path(__file__).resolve().parent / "golden"
    g = np.load(golden / f"mdp_trace_{name}.npz")
    lib.mdp_sizeof.restype = C.c_size_t
    lib.mdp_act.restype = lib.mdp_reward.restype = lib.mdp_cumulative.restype = C.c_double
    lib.mdp_act.argtypes = [C.c_void_p, C.c_int]
    lib.mdp_observe.argtypes = [C.c_void_p] + [C.c_double] * 5 + [C.c_int]
    lib.mdp_init.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
    for f in (lib.mdp_reset, lib.mdp_check, lib.mdp_reward, lib.mdp_cumulative, lib.mdp_done):
        f.argtypes = [C.c_void_p]
    buf = C.create_string_buffer(lib.mdp_sizeof())
    m = C.cast(buf, C.c_void_p)
    lib.mdp_init(m, int(g["w"]), 22.92, 20.0, 4.5)
    n_done = 0
    for i in range(len(g["action"])):
        o = [float(x) for x in g["obs"][i].astype(np.float64)]
        if g["action"][i] == 255:
            lib.mdp_reset(m)
            assert lib.mdp_observe(m, *o, int(g["contact"][i])) == g["state"][i]
            continue
        assert lib.mdp_act(m, int(g["action"][i])) == g["theta_sp"][i]
        sid = lib.mdp_observe(m, *o, int(g["contact"][i]))
        code = lib.mdp_check(m)
        r = lib.mdp_reward(m)
        assert (sid, code, lib.mdp_done(m)) == (g["state"][i], g["code"][i], g["done"][i]), i
        assert r == g["reward"][i] and lib.mdp_cumulative(m) == g["cum"][i], i
        n_done += int(g["done"][i])
    assert n_done >= 1


@pytest.mark.parametrize("name", ["replay_w0_float32", "replay_w0_float32_ep1950", "replay_w2_float32", "replay_w4_float32"])
def test_single_env_loop_c_replays_the_reference_trainer_fixtures(lib, name):
    """oracle/c/loop.c (guess / exploration_rate / alpha / update on float32 tables + MDP + stand-in + Philox contract, one env)
    against the replay fixtures produced by the UNMODIFIED reference trainer loop (tests/golden/replay_*_float32.npz):
    observations (bits), actions, states, codes, done flags, float64 rewards, episode indices, and the final float32 Q table and
    counts are identical."""
    golden = pathlib.Path(__file__).resolve().parent / "golden"
    g = np.load(golden / f"{name}.npz")
    sp = StandInParams()
    dyn = StandInDet(sp, 1)
    p, d = _params(sp, dyn), derive(sp)
    lp = LoopParams(dz=d.dz, z_touch=d.z_touch, half_platform=d.half_platform, p_max_f=d.p_max, two_p_max_f=d.two_p_max, sigma_x=d.sigma_x,
                    f_ag=22.92, t_max=20.0, p_max=4.5, alpha_min=0.02949, omega=0.51, gamma=0.99)
    n = len(g["action"])
    qa = np.ascontiguousarray(g["qa0"], np.float32).reshape(-1).copy()
    qb = np.ascontiguousarray(g["qb0"], np.float32).reshape(-1).copy()
    count = np.zeros(2835, np.float64)
    obs = np.zeros((n, 5), np.float32)
    i32 = lambda: np.zeros(n, np.int32)
    action, state, nstate, code, done, episode = i32(), i32(), i32(), i32(), i32(), i32()
    reward = np.zeros(n, np.float64)
    lib.mdp_sizeof.restype = C.c_size_t
    buf = C.create_string_buffer(lib.mdp_sizeof())
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.oracle_single_env_loop.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 11
    episodes = lib.oracle_single_env_loop(C.byref(p), C.byref(lp), C.cast(buf, C.c_void_p), int(g["seed"]), 0, int(g["w"]), int(g["ep0"]), n,
                                          ptr(qa), ptr(qb), ptr(count), ptr(obs), ptr(action), ptr(state), ptr(nstate), ptr(code),
                                          ptr(done), ptr(reward), ptr(episode))
    assert np.array_equal(obs.view(np.uint32), g["obs"].view(np.uint32))
    for mine, key in ((action, "action"), (state, "state"), (nstate, "next_state"), (code, "code"), (done, "done"), (episode, "episode")):
        assert np.array_equal(mine, g[key].astype(np.int32)), key
    assert np.array_equal(reward, g["reward"])
    assert np.array_equal(qa.view(np.uint32), g["qa"].reshape(-1).view(np.uint32))
    assert np.array_equal(count, g["count"].reshape(-1)) and episodes == int(g["done"].sum()) >= 1


@pytest.mark.parametrize("mode,opts", [("reference", {}), ("paper", {}), ("reference", dict(n_sub=2, accel_mode="kalman"))])
def test_population_c_equals_python_population_oracle(lib, mode, opts):
    """oracle/c/population.c (batched semantics S1 incl. success window, promotion, transfer, fresh-MDP restart) against
    oracle/loop.py: PopulationOracle on 40 envs through several curriculum steps: traces, float32 tables, counts, counters."""
    from oracle.c_loop import run_population_c
    from oracle.loop import PopulationOracle, TrainerParams
    tp = TrainerParams(success_rate=0.15, successive_successful_episodes=6, max_num_episodes=90, transfer_mode=mode)
    sp = StandInParams(**opts)
    n_envs, steps = 40, 420
    out = run_population_c(n_envs, steps, seed=3, population=2, w0=0, tp=tp, sp=sp, trace=True)
    pop = PopulationOracle(n_envs, seed=3, population=2, w0=0, dtype=np.float32, tp=tp, sp=sp)
    for t in range(steps):
        if pop.finished:
            break
        o = pop.step()
        assert np.array_equal(out["action"][t], o["action"]) and np.array_equal(out["next_state"][t], o["next_state"]), t
        assert np.array_equal(out["code"][t], o["code"]) and np.array_equal(out["reward"][t], o["reward"]), t
    r = out["result"]
    assert np.array_equal(out["qa"].view(np.uint32), pop.agent.qa.view(np.uint32)) and np.array_equal(out["qb"].view(np.uint32), pop.agent.qb.view(np.uint32))
    assert np.array_equal(out["count"], pop.agent.count)
    assert (r.w, bool(r.finished), r.total_episodes, r.total_successes, r.total_steps) == (pop.w, pop.finished, pop.total_episodes, pop.total_successes, pop.total_steps)
    assert list(r.term_hist) == list(pop.term_hist) and r.window_sum == sum(pop.window) and r.window_count == len(pop.window)
    assert r.n_promotions == len(pop.promotions) >= 2


def test_population_c_replays_the_reference_curriculum_run(lib, golden_dir):
    """oracle/c/population.c with one env against tests/golden/curriculum_ref.npz = the UNMODIFIED Trainer.curriculum_training()
    (PKG/trainer.py:169-245): every action, state, check code and float64 reward of all five curriculum steps, the step at which
    each curriculum step ended (promotion or max-episodes advance) and the final tables after the last transfer (quirk Q7)."""
    from oracle.c_loop import run_population_c
    from oracle.loop import TrainerParams
    g = np.load(golden_dir / "curriculum_ref.npz")
    tp = TrainerParams(successive_successful_episodes=int(g["successive_successful_episodes"]), success_rate=float(g["success_rate"]),
                       max_num_episodes=int(g["max_num_episodes"]))
    n = len(g["action"])
    out = run_population_c(1, n, seed=int(g["seed"]), population=0, w0=0, tp=tp, trace=True)
    assert np.array_equal(out["action"][:, 0], g["action"])
    assert np.array_equal(out["next_state"][:, 0], g["next_state"])
    assert np.array_equal(out["code"][:, 0], g["code"])
    assert np.array_equal(out["reward"][:, 0], g["reward"])
    res = out["result"]
    assert res.finished == 1 and res.t == n and res.n_promotions == 5
    assert res.total_episodes == int(g["done"].sum()) and res.total_steps == n
    assert np.array_equal(out["qa"].view(np.uint32), g["qa"].view(np.uint32)) and np.array_equal(out["qb"].view(np.uint32), g["qb"].view(np.uint32))
    assert np.array_equal(out["count"], g["count"])
    # one step fewer: the run is not finished and the last transfer has not happened
    out2 = run_population_c(1, n - 1, seed=int(g["seed"]), population=0, w0=0, tp=tp)
    assert out2["result"].finished == 0 and out2["result"].w == 4 and out2["result"].n_promotions == 4
