"""Host logic: the fp32 cut tables / LUTs of constants.py against the reference-generated fixtures
(no GPU; a NumPy emulation of the kernel's compare chain is used to evaluate the cut tables)."""
import ctypes as C

import numpy as np
import pytest

from dql_multirotor_landing_b200 import constants as K


def cut_discretise(cfg, w, obs):
    """What the kernel does with the cut tables (csrc/dqlb200_device.cuh: discretise_cuts)."""
    c = cfg.cuts[w]
    x = [obs[:, 0], obs[:, 1], obs[:, 2]]
    with np.errstate(invalid="ignore"):
        lvl = []
        for q in range(2):
            n = np.zeros(len(obs), np.int32)
            for i in range(4):
                n += (x[q] >= np.float32(c.lvl_lo[q][i])) & ~(x[q] >= np.float32(c.lvl_hi[q][i]))
            lvl.append(n)
        level = np.minimum(lvl[0], lvl[1])
        bins = []
        for q in range(3):
            b1 = np.asarray([c.bin1[q][l] for l in range(5)], np.float32)[level]
            b2 = np.asarray([c.bin2[q][l] for l in range(5)], np.float32)[level]
            bins.append((x[q] >= b1).astype(np.int32) + (x[q] >= b2))
        ang = np.zeros(len(obs), np.int32)
        for i in range(6):
            ang += obs[:, 3] >= np.float32(cfg.angle_cut[i])
    return np.stack([level, bins[0], bins[1], bins[2], ang], axis=1).astype(np.int8)


@pytest.fixture(scope="module")
def cfg():
    return K.build_config(1, 1)


def test_struct_sizes(cfg):
    assert cfg.struct_bytes == C.sizeof(K.Config)
    assert C.sizeof(K.PopulationState) == K.POPULATION_STATE_DTYPE.itemsize == 320
    assert C.sizeof(K.PopulationParams) == 40


def test_cut_tables_vs_reference_fixture(cfg, golden_dir):
    g = np.load(golden_dir / "discretise.npz")
    for w in range(5):
        got = cut_discretise(cfg, w, g["obs"])
        bad = np.nonzero((got != g["train"][w]).any(axis=1))[0]
        assert bad.size == 0, (w, g["obs"][bad[:5]], got[bad[:5]], g["train"][w][bad[:5]])


def test_cut_tables_vs_reference_live(cfg, reference_ns):
    """Dense probe around every cut (+-4 ulp) against the imported reference."""
    ns = reference_ns
    for w in (0, 2, 4):
        mdp = ns.mdp.TrainingMdp(w, 22.92, 20, 4.5)
        mdp.reset()
        c = cfg.cuts[w]
        probes = []
        for q in range(3):
            vals = [c.bin1[q][l] for l in range(w + 1)] + [c.bin2[q][l] for l in range(w + 1)]
            if q < 2:
                vals += [c.lvl_lo[q][i] for i in range(w)] + [c.lvl_hi[q][i] for i in range(w)]
            for v in vals:
                x = np.float32(v)
                for _ in range(4):
                    x = np.nextafter(x, np.float32(-np.inf))
                for _ in range(9):
                    row = [np.float32(0.01), np.float32(0.01), np.float32(0.01), np.float32(0.0)]
                    row[q] = x
                    probes.append(row)
                    x = np.nextafter(x, np.float32(np.inf))
        for v in cfg.angle_cut:
            x = np.float32(v)
            for _ in range(4):
                x = np.nextafter(x, np.float32(-np.inf))
            for _ in range(9):
                probes.append([np.float32(0.3), np.float32(-0.2), np.float32(0.1), x])
                x = np.nextafter(x, np.float32(np.inf))
        obs = np.asarray(probes, np.float32)
        got = cut_discretise(cfg, w, obs)
        for k, row in enumerate(obs):
            o = ns.Observation(rel_p_x=float(row[0]), rel_v_x=float(row[1]), rel_a_x=float(row[2]))
            ref = mdp.discrete_state(ns.mdp.ContinuousObservation(o, float(row[3]), 0.0, 3.0))
            assert tuple(got[k]) == ref, (w, row, got[k], ref)


def test_schedule_luts(cfg, golden_dir):
    g = np.load(golden_dir / "schedules.npz")
    lut = K.alpha_lut()
    assert np.array_equal(lut, g["alpha"][: K.ALPHA_LUT].astype(np.float32))
    assert lut[-1] == np.float32(0.02949) and np.all(g["alpha"][K.ALPHA_LUT - 1:] == 0.02949)
    thr = np.asarray(list(cfg.eps_threshold), np.uint64)
    eps = g["eps"]
    # exhaustive equivalence on a sample of 24-bit draws: (u24 * 2^-24 < eps)  <=>  (u24 < thr)
    rng = np.random.default_rng(0)
    u = np.concatenate([rng.integers(0, 2 ** 24, 4096), [0, 1, 2 ** 24 - 1, 167772, 167773]]).astype(np.uint64)
    for e in list(range(0, 2001, 37)) + [800, 801, 1999, 2000, 2001]:
        ref = (u.astype(np.float64) * 2.0 ** -24) < eps[e]
        assert np.array_equal(ref, u < thr[min(e, K.EPS_LUT - 1)]), e
    assert np.all(eps[2001:] == eps[2001]) and thr[2001] == K.explore_threshold(0.01)
    assert [cfg.transfer_ratio[k] for k in range(5)] == [np.float32(r) for r in g["ratios"]]


def test_scalar_thresholds(cfg):
    assert cfg.timeout_steps == 459 and cfg.success_steps == 23
    assert cfg.promote_successes == 97                                 # > 0.96 of 100
    assert cfg.fz_hi == np.nextafter(np.float32(4.5), np.float32(np.inf)) and cfg.fz_lo == np.float32(-4.5)
    assert np.float64(cfg.z_min_cut) >= 0.2 > np.float64(np.nextafter(np.float32(cfg.z_min_cut), np.float32(-1)))
    r0 = cfg.reward[0]
    assert abs(r0.r_term_fail / -2.6 - 5.054188) < 1e-6                # r_max level 0, SURVEY.md A.5
    with pytest.raises(ValueError):
        K.alpha_lut(alpha_min=0.001)                                   # does not saturate within the LUT


def test_setpoint_tables_closure(cfg):
    """The pitch set-point is stored as an index: the tables must be closed under the three actions and reproduce the float64
    arithmetic of continuous_action (PKG/mdp.py:543-560) from every reachable value."""
    mp = K.MdpParameters()
    n, zero = cfg.n_setpoints, cfg.setpoint_zero
    vals = [cfg.setpoint_value[i] for i in range(n)]
    assert n == 33 and vals[zero] == 0.0 and vals == sorted(vals) and len(set(vals)) == n       # 3 interleaved lattices, SURVEY.md R3
    assert max(vals) == mp.theta_max and min(vals) == -mp.theta_max
    for i, v in enumerate(vals):
        want = [min((v + mp.delta_theta, mp.theta_max)), max((v - mp.delta_theta, -mp.theta_max)), v]
        for a in range(3):
            nx = cfg.setpoint_next[i][a]
            assert vals[nx.next] == want[a] and nx.value_f32 == np.float32(want[a])
    # every value is reachable from 0
    seen, todo = {zero}, [zero]
    while todo:
        i = todo.pop()
        for a in range(3):
            j = cfg.setpoint_next[i][a].next
            if j not in seen:
                seen.add(j); todo.append(j)
    assert len(seen) == n


def test_setpoint_tables_vs_reference_live(cfg, reference_ns):
    """Against the unmodified TrainingMdp: continuous_action from every reachable set-point, and the set-point term of reward()
    (everything else held at zero) for every (previous index, action), in a running and in a fresh episode (quirk Q11)."""
    ns = reference_ns
    n, zero = cfg.n_setpoints, cfg.setpoint_zero
    vals = [cfg.setpoint_value[i] for i in range(n)]
    lim_v = 1.0                                                  # Limits.velocity[0]
    obs0 = ns.mdp.ContinuousObservation(ns.Observation(), 0.0, 0.0, 3.0)
    for p in range(n):
        for a in range(3):
            for fresh in (0, 1):
                mdp = ns.mdp.TrainingMdp(0, 22.92, 20, 4.5)
                mdp.reset()
                mdp._current_continuous_action.pitch = vals[p]
                mdp.discrete_state(obs0); mdp.continuous_action(2); mdp.discrete_state(obs0); mdp.check(); mdp.reward()   # phi_theta(prev) in place
                base = mdp.reward()                               # a step that changes nothing: r_p = r_v = r_theta = 0
                if fresh:
                    mdp.reset()                                   # keeps the shaping values (quirk Q11), zeroes the set-point
                    mdp.discrete_state(obs0)
                act = mdp.continuous_action(a)
                assert act.pitch == vals[cfg.setpoint_next[zero if fresh else p][a].next]
                mdp.discrete_state(obs0); mdp.check()
                r = mdp.reward()
                want = cfg.setpoint_rtheta[fresh][p][a] * lim_v
                assert r - base == pytest.approx(want, abs=1e-12), (p, a, fresh, r - base, want)
                # exactly: the reference's own expression (PKG/mdp.py:506-514) on the shaping values reward() just used
                cur, prev = mdp.current_shaping_value.angle, mdp.previous_shaping_value.angle
                assert mdp._w_theta * (np.abs(cur) - np.abs(prev)) / mdp._theta_max * lim_v == want, (p, a, fresh)


def test_trainer_replica_shape():
    """Trainer splits ONE agent's envs into replicas of trainer.replica_shape() envs: about 128 replicas when merging after every
    step (multiples of 128 envs, at most 1,024 per replica), 128-env replicas for rarer merges; every env is covered."""
    from dql_multirotor_landing_b200.trainer import replica_shape
    assert [replica_shape(n) for n in (4096, 16384, 65536, 262144, 1 << 20)] == [128, 128, 512, 1024, 1024]
    assert replica_shape(65536, merge_every=16) == 128
    for n in (2049, 5000, 70000, 300000):
        n_r = replica_shape(n)
        assert n_r % 128 == 0 and 128 <= n_r <= 1024 and -(-n // n_r) * n_r >= n
