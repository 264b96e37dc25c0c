"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Committed NumPy version of the analytic relative-dynamics stand-in (SURVEY.md
A.3).  There is NO reference function for this row (R4): the reference gets
its observations from Gazebo + RotorS + PID/attitude nodes.  The parameters
are traced to the reference:

  * pitch first-order lag  tau = k_omega/k_R = 0.1/0.7   (PKG/attitude_controller.py:86-87)
  * g = 9.81                                             (PKG/attitude_controller.py:59)
  * rotor-drag c_d ~= 0.2 1/s        (rotors_gazebo_plugins/src/gazebo_motor_model.cpp:464-466,
                                      rotors_description/urdf/hummingbird.xacro:29-42)
  * platform x = r sin(w t), u = r w cos(w t), w = v/r    (PKG/moving_platform.py:116-125)
  * rel = platform - drone                                (PKG/observation_utils.py:225,249)
  * contact height 0.515 m, platform 1x1 m                (urdf/moving_platform.urdf:16,38,51,58)

Two versions of the SAME equations:

``StandInDet``  fp32, every operation a separately rounded IEEE add/mul/div/sqrt
                (no FMA, own polynomial sin/cos/tan/log), so the CUDA kernel, the
                C oracle and NumPy agree BIT FOR BIT.  Used for replay parity.
``standin_f64`` float64 with np.sin/np.tan -- the "textbook" form; the fp32
                path must stay within 1e-5 of it (north_star tolerance).

The platform phase is a uint32 in "turns" (2**32 == one period): phase
advance is exact integer arithmetic, range reduction is exact, and a random
start phase is just a Philox word.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass

import numpy as np

f32 = np.float32
TWO_PI = 2.0 * math.pi


@dataclass(frozen=True)
class StandInParams:
    f_ag: float = 22.92          # PKG/trainer.py:42
    tau_theta: float = 1.0 / 7.0
    c_d: float = 0.2
    g: float = 9.81
    r_mp: float = 2.0            # PKG/moving_platform.py r_x default
    v_mp: float = 1.6            # launch/environment.launch:62-64 (intended)
    z_init: float = 4.0          # PKG/trainer.py:41
    v_z: float = -0.1            # PKG/mdp.py:212 (training); -0.4 PKG/mdp.py:580 (simulation)
    z_touch: float = 0.515
    half_platform: float = 0.5
    p_max: float = 4.5
    n_sub: int = 1
    noise_pos_sd: float = 0.0    # PKG/observation_utils.py:127-129 (launch/environment.launch:56-57 sets both to 0)
    noise_vel_sd: float = 0.0
    # relative acceleration the MDP sees: "exact" | "kalman_reference" | "kalman" -- see KalmanAccel
    accel_mode: str = "exact"
    kf_process_variance: float = 1e-4     # scripts/manager_node.py:96-98
    kf_measurement_sd: float = 0.1        # manager_node hands noise_vel_sd (default 0.1) to KalmanFilter3D; R = sd ** 2 (PKG/filters.py:50-52)
    # "first_order" | "second_order" (SURVEY 8f-4): see SecondOrder
    dynamics_model: str = "first_order"
    mass: float = 0.68                    # PKG/attitude_controller.py:57
    inertia: float = 0.007                # PKG/attitude_controller.py:59
    k_R: float = 0.7                      # PKG/attitude_controller.py:86
    k_omega: float = 0.1                  # PKG/attitude_controller.py:87
    pid_kp: float = 5.0                   # launch/drone.launch:35-40
    pid_ki: float = 10.0
    pid_lower: float = 0.0
    pid_upper: float = 10.0
    pid_windup: float = 10.0
    pid_ticks: int = 10                   # PKG/pid.py:14 rate_hz = 1000 against the 100 Hz state topic


@dataclass(frozen=True)
class DerivedF32:
    """fp32 constants derived on the host in float64, then rounded once."""
    h: np.float32
    half_h2: np.float32
    k_theta: np.float32
    g: np.float32
    c_d: np.float32
    r: np.float32
    rw: np.float32
    rw2: np.float32
    dz: np.float32
    z_init: np.float32
    z_touch: np.float32
    half_platform: np.float32
    p_max: np.float32
    two_p_max: np.float32
    sigma_x: np.float32
    dphase: int                  # uint32 turns per sub-step
    n_sub: int


def derive(p: StandInParams) -> DerivedF32:
    h = (1.0 / p.f_ag) / p.n_sub
    w = p.v_mp / p.r_mp
    return DerivedF32(
        h=f32(h), half_h2=f32(0.5 * h * h), k_theta=f32(-math.expm1(-h / p.tau_theta)),
        g=f32(p.g), c_d=f32(p.c_d), r=f32(p.r_mp), rw=f32(p.r_mp * w), rw2=f32(p.r_mp * w * w),
        dz=f32(p.v_z * (1.0 / p.f_ag)), z_init=f32(p.z_init), z_touch=f32(p.z_touch),
        half_platform=f32(p.half_platform), p_max=f32(p.p_max), two_p_max=f32(2.0 * p.p_max),
        sigma_x=f32(p.p_max / 3.0),
        dphase=int(round(w * h / TWO_PI * 2.0 ** 32)) & 0xFFFFFFFF, n_sub=p.n_sub,
    )


# ----------------------------------------------------------------------------------------
# deterministic fp32 math: every op is one correctly rounded IEEE operation on np.float32
# ----------------------------------------------------------------------------------------
_TURN_TO_RAD = f32(TWO_PI / 2.0 ** 32)
_S = [f32(-1.0 / 6.0), f32(1.0 / 120.0), f32(-1.0 / 5040.0), f32(1.0 / 362880.0)]
_C = [f32(-0.5), f32(1.0 / 24.0), f32(-1.0 / 720.0), f32(1.0 / 40320.0), f32(-1.0 / 3628800.0)]
_T = [f32(1.0 / 3.0), f32(2.0 / 15.0), f32(17.0 / 315.0), f32(62.0 / 2835.0),
      f32(1382.0 / 155925.0), f32(21844.0 / 6081075.0)]
_L = [f32(1.0 / 3.0), f32(1.0 / 5.0), f32(1.0 / 7.0), f32(1.0 / 9.0)]
_LN2 = f32(math.log(2.0))
_SQRT2 = f32(math.sqrt(2.0))
ONE = f32(1.0)


def fma32(a, b, c):
    """Correctly rounded float32 fused multiply-add RN(a * b + c) on arrays (what FFMA / fmaf compute), in NumPy.  The product of
    two float32 is exact in float64.  Rounding the float64 sum s to float32 can only go wrong (double rounding) when s sits exactly
    on a float32 tie AND the float64 addition was inexact: no other float32 tie can lie between the exact sum and s, because a tie
    is itself a float64 number and s is the float64 nearest to the exact sum.  Those entries (one in 2**29) are redone with the
    exact error of the addition (TwoSum): the sum is moved off the tie toward the exact value before the final rounding."""
    a, b, c = np.asarray(a), np.asarray(b), np.asarray(c)
    if a.size == 1 and b.size == 1 and c.size == 1:      # single env: Python floats (IEEE double) instead of NumPy call overhead
        sf = float(a.reshape(-1)[0]) * float(b.reshape(-1)[0]) + float(c.reshape(-1)[0])
        if (struct.unpack("<Q", struct.pack("<d", sf))[0] & 0x1FFFFFFF) != 0x10000000:
            return np.full(np.broadcast(a, b, c).shape, sf, dtype=np.float32)
    p = np.multiply(a, b, dtype=np.float64)
    c64 = np.asarray(c, dtype=np.float64)
    s = np.asarray(p + c64)
    bits = s.view(np.uint64)
    tie = (bits & np.uint64(0x1FFFFFFF)) == np.uint64(0x10000000)
    if tie.any():
        p, c64 = np.broadcast_arrays(p, c64)
        bb = s - p
        err = (p - (s - bb)) + (c64 - bb)
        grows = (err > 0) == (s > 0)                      # the exact sum is further from zero than s
        fix = tie & (err != 0)
        bits = np.where(fix, np.where(grows, bits + np.uint64(1), bits - np.uint64(1)), bits)
        s = bits.view(np.float64)
    return s.astype(np.float32)


def det_sincos_turns(phase):
    """sin, cos of 2*pi*phase/2**32 for uint32 phase (array).  Horner steps are fused multiply-adds (fma32)."""
    phase = np.atleast_1d(np.asarray(phase, dtype=np.uint32))
    q = ((phase + np.uint32(0x20000000)) >> np.uint32(30)).astype(np.uint32)
    rem = (phase - (q << np.uint32(30))).astype(np.uint32).view(np.int32)
    x = rem.astype(np.float32) * _TURN_TO_RAD
    z = x * x
    ps = _S[3]
    for c in (_S[2], _S[1], _S[0]):
        ps = fma32(ps, z, c)
    s = fma32(x, z * ps, x)
    pc = _C[4]
    for c in (_C[3], _C[2], _C[1], _C[0]):
        pc = fma32(pc, z, c)
    c_ = fma32(z, pc, ONE)
    qq = q & np.uint32(3)
    sin = np.where(qq == 0, s, np.where(qq == 1, c_, np.where(qq == 2, -s, -c_)))
    cos = np.where(qq == 0, c_, np.where(qq == 1, -s, np.where(qq == 2, -c_, s)))
    return sin.astype(np.float32), cos.astype(np.float32)


def det_tan(x):
    """tan(x) for |x| <= ~0.45 rad (odd Taylor polynomial to x**13, Horner in fused multiply-adds)."""
    x = np.atleast_1d(np.asarray(x, dtype=np.float32))
    z = x * x
    p = _T[5]
    for c in (_T[4], _T[3], _T[2], _T[1], _T[0]):
        p = fma32(p, z, c)
    return fma32(x, z * p, x)


def det_log(u):
    """ln(u) for normal positive fp32 u (atanh series after exact exponent split)."""
    u = np.ascontiguousarray(np.atleast_1d(np.asarray(u, dtype=np.float32)))
    bits = u.reshape(-1).view(np.uint32).reshape(u.shape)
    e = ((bits >> np.uint32(23)) & np.uint32(0xFF)).astype(np.int32) - 127
    m = ((bits & np.uint32(0x007FFFFF)) | np.uint32(0x3F800000)).reshape(-1).view(np.float32).reshape(u.shape)
    big = m > _SQRT2
    m = np.where(big, m * f32(0.5), m).astype(np.float32)
    e = np.where(big, e + 1, e)
    s = (m - ONE) / (m + ONE)
    z = s * s
    p = _L[3]
    for c in (_L[2], _L[1], _L[0]):
        p = p * z + c
    lm = (s + s) * (ONE + z * p)
    return (e.astype(np.float32) * _LN2 + lm).astype(np.float32)


def det_normal(x0, x1):
    """Standard normal from two Philox words (Box-Muller, deterministic)."""
    u1 = ((np.atleast_1d(np.asarray(x0, dtype=np.uint32)) >> np.uint32(8)).astype(np.float32) + ONE) * f32(2.0 ** -24)
    rad = np.sqrt(f32(-2.0) * det_log(u1)).astype(np.float32)
    _, c = det_sincos_turns(x1)
    return (rad * c).astype(np.float32)


def det_normal_pair(x0, x1):
    """Both Box-Muller normals of one (u1, angle) pair (rad * cos, rad * sin)."""
    u1 = ((np.atleast_1d(np.asarray(x0, dtype=np.uint32)) >> np.uint32(8)).astype(np.float32) + ONE) * f32(2.0 ** -24)
    rad = np.sqrt(f32(-2.0) * det_log(u1)).astype(np.float32)
    s, c = det_sincos_turns(x1)
    return (rad * c).astype(np.float32), (rad * s).astype(np.float32)


def add_observation_noise(p: "StandInParams", rel_p, rel_v, w0, w1):
    """What the MDP sees: true relative position / velocity plus independent Gaussians (PKG/observation_utils.py:127-129)."""
    n0, n1 = det_normal_pair(w0, w1)
    return ((np.asarray(rel_p, np.float32) + f32(p.noise_pos_sd) * n0).astype(np.float32),
            (np.asarray(rel_v, np.float32) + f32(p.noise_vel_sd) * n1).astype(np.float32))


_SS = [f32(-1.0 / 6.0), f32(1.0 / 120.0), f32(-1.0 / 5040.0), f32(1.0 / 362880.0)]


def det_sin_small(x):
    """sin(x), |x| <= ~0.8 rad (Taylor to x**9, Horner, no FMA) -- the kernels' det_sin_small."""
    x = np.asarray(x, np.float32)
    z = x * x
    p = _SS[3]
    for c in (_SS[2], _SS[1], _SS[0]):
        p = p * z + c
    return (x + x * (z * p)).astype(np.float32)


def det_cos_small(x):
    """cos(x), |x| <= ~0.8 rad (Taylor to x**10) -- the kernels' det_cos_small."""
    x = np.asarray(x, np.float32)
    z = x * x
    p = _C[4]
    for c in (_C[3], _C[2], _C[1], _C[0]):
        p = p * z + c
    return (ONE + z * p).astype(np.float32)


class VerticalPid:
    """The pid_v_z node (PKG/pid.py:62-104; gains launch/drone.launch:33-46, Kd = 0) in fp32, vectorised over envs: integral
    with wind-up clip (:85-86), ButterworthFilter on the error (PKG/filters.py:83-108, c = 1), effort = clip(Kp f + Ki I, lower,
    upper) (:97-103).  The filter as written pushes the new input BEFORE the sum and the new output after it, so its output
    taps are one sample older than its input taps (quirk Q14): y_k = (x_k-2 + 2 x_k-1 + x_k - 0.586 y_k-3 - 0 * y_k-2) / 3.414.  `ticks` node iterations per call on a
    held error (the node spins at 1 kHz, its state topic arrives at 100 Hz); dt = h / ticks.  The memory is the node's: it
    survives episode resets.  A simulator that has been hovering starts with I = m g / Ki."""

    def __init__(self, n: int, p: "StandInParams", h: float):
        self.kp, self.ki = f32(p.pid_kp), f32(p.pid_ki)
        self.lo, self.hi, self.windup = f32(p.pid_lower), f32(p.pid_upper), f32(p.pid_windup)
        self.ticks, self.dt = p.pid_ticks, f32(h / p.pid_ticks)
        self.inv_denom, self.k2 = f32(1.0 / (1 + 1.0 ** 2 + 1.414 * 1.0)), f32(1.0 ** 2 - 1.414 * 1.0 + 1)
        self.integ = np.full(n, f32(p.mass * abs(p.g) / p.pid_ki), f32)
        assert self.ticks >= 2        # the held error makes both previous filter inputs equal at a sub-step boundary
        self.e1, self.f1, self.f2, self.f3 = (np.zeros(n, f32) for _ in range(4))

    def thrust(self, idx, error):
        e = np.asarray(error, f32)
        I, e1, f1, f2, f3 = self.integ[idx], self.e1[idx], self.f1[idx], self.f2[idx], self.f3[idx]
        e2 = e1
        T = np.zeros_like(e)
        for _ in range(self.ticks):
            I = np.clip(I + e * self.dt, -self.windup, self.windup).astype(np.float32)
            f = (self.inv_denom * (((e2 + f32(2.0) * e1) + e) - self.k2 * f3)).astype(np.float32)
            e2, e1 = e1, e
            f3, f2, f1 = f2, f1, f
            T = np.clip(self.kp * f + self.ki * I, self.lo, self.hi).astype(np.float32)
        self.integ[idx], self.e1[idx], self.f1[idx], self.f2[idx], self.f3[idx] = I, e + np.zeros_like(I), f1, f2, f3
        return T


class KalmanAccel:
    """The acceleration estimator of the reference's observation node, vectorised over envs, fp32 with one rounding per
    operation (the CUDA kernels' kf_sample is the same arithmetic): KalmanFilter1D (PKG/filters.py:4-37; x = 0, P = 1, Q = process
    variance, R = measurement_sd ** 2) fed by KalmanFilter3D.filter (PKG/filters.py:54-80) with the finite difference of the TRUE
    relative velocity (PKG/observation_utils.py:134-150 passes rel_vel, not the noisy copy).

    mode "kalman_reference" reproduces the node as written: last_velocity / last_timestep are set by the first observation and
    never refreshed (PKG/observation_utils.py:137-139 is their only assignment), so raw = (v_now - v_first) / (t_now - t_first)
    with an ever-growing time base (quirk Q13); the first observation reports 0 and does not touch the filter (:140-143).
    mode "kalman" is the evident intention: raw = (v_now - v_prev) / h.
    The filter belongs to the node, not to the episode: nothing resets it (no reset path in scripts/manager_node.py touches
    self.utils), so it runs through teleports, hover periods and curriculum steps."""

    def __init__(self, n: int, mode: str, h: np.float32, q: float, r: float):
        assert mode in ("kalman_reference", "kalman")
        self.mode, self.h, self.q, self.r = mode, f32(h), f32(q), f32(r)
        self.x = np.zeros(n, f32)
        self.P = np.ones(n, f32)
        self.v_ref = np.zeros(n, f32)
        self.n = np.zeros(n, np.uint32)

    def sample(self, idx, rel_v):
        idx = np.asarray(idx)
        rel_v = np.asarray(rel_v, f32)
        first = self.n[idx] == 0
        x, P, v_ref, n = self.x[idx], self.P[idx], self.v_ref[idx], self.n[idx]
        with np.errstate(divide="ignore", invalid="ignore"):
            dt = (n.astype(np.float32) * self.h) if self.mode == "kalman_reference" else np.full(n.shape, self.h, f32)
            raw = ((rel_v - v_ref) / dt).astype(np.float32)
            P1 = (P + self.q).astype(np.float32)
            K = (P1 / (P1 + self.r)).astype(np.float32)
            x1 = (x + K * (raw - x)).astype(np.float32)
            P2 = (P1 * (ONE - K)).astype(np.float32)
        self.x[idx] = np.where(first, x, x1)
        self.P[idx] = np.where(first, P, P2)
        if self.mode == "kalman_reference":
            self.v_ref[idx] = np.where(first, rel_v, v_ref)
            self.n[idx] = n + np.uint32(1)
        else:
            self.v_ref[idx] = rel_v
            self.n[idx] = np.uint32(1)


class StandInDet:
    """Vectorised (over envs) deterministic fp32 stand-in.  State arrays are np.float32/uint32."""

    def __init__(self, params: StandInParams, n: int):
        self.p = params
        self.d = derive(params)
        self.n = n
        self.x_d = np.zeros(n, f32)
        self.v_d = np.zeros(n, f32)
        self.theta = np.zeros(n, f32)
        self.phase = np.zeros(n, np.uint32)
        self.a_d = np.zeros(n, f32)
        self.so = params.dynamics_model == "second_order"
        if self.so:       # pitch rate, altitude, vertical velocity are state; the PID node has memory
            self.omega = np.zeros(n, f32)
            self.z = np.full(n, self.d.z_init, f32)
            self.v_z = np.zeros(n, f32)
            self.pid = VerticalPid(n, params, float(self.d.h))
            self.att_kr, self.att_kw = f32(params.k_R / params.inertia), f32(params.k_omega / params.inertia)
            self.inv_m, self.inv_mg, self.g_abs = f32(1.0 / params.mass), f32(1.0 / (params.mass * abs(params.g))), f32(abs(params.g))
            self.vz_sp = f32(params.v_z)
        self.kf = None
        if params.accel_mode != "exact":
            self.kf = KalmanAccel(n, params.accel_mode, self.d.h, params.kf_process_variance, params.kf_measurement_sd ** 2)

    # R1 (PKG/landing_simulation_env.py:181-216) / R15 (:327-340)
    def reset(self, idx, w0, w1, w2, *, normal_init: bool, simulation: bool = False):
        d = self.d
        idx = np.asarray(idx)
        w0 = np.asarray(w0, dtype=np.uint32)
        if normal_init:
            x_init = d.sigma_x * det_normal(w0, w1)
        else:
            u = (w0 >> np.uint32(8)).astype(np.float32) * f32(2.0 ** -24)
            x_init = -d.p_max + d.two_p_max * u
        phase = np.asarray(w2, dtype=np.uint32)
        s, _ = det_sincos_turns(phase)
        x_mp = d.r * s
        if simulation:
            # absolute clip, platform minus offset (PKG/landing_simulation_env.py:331-335)
            x_d = np.clip(x_mp - x_init, -d.p_max, d.p_max)
        else:
            # clip(x_init + x_mp, x_mp +- p_max) == x_mp + clip(x_init, +-p_max) up to rounding;
            # the stand-in defines it as the latter (PKG/landing_simulation_env.py:197-201)
            x_d = x_mp + np.clip(x_init, -d.p_max, d.p_max)
        self.x_d[idx] = x_d.astype(np.float32)
        self.v_d[idx] = 0
        self.theta[idx] = 0
        self.phase[idx] = phase
        if self.so:       # teleport with zero twist (PKG/landing_simulation_env.py:203-216); the PID memory stays
            self.omega[idx] = 0
            self.z[idx] = self.d.z_init
            self.v_z[idx] = 0

    def advance(self, theta_sp, idx=None, hover: bool = False):
        """One agent period (n_sub sub-steps) toward set-point theta_sp (fp32 array).  hover = True: the period after a reset,
        every set-point zero (scripts/manager_node.py:328) -- only the second-order model has a vertical set-point."""
        d = self.d
        sl = slice(None) if idx is None else idx
        x, v, th, ph = self.x_d[sl], self.v_d[sl], self.theta[sl], self.phase[sl]
        sp = np.asarray(theta_sp, dtype=np.float32)
        a = self.a_d[sl]
        for _ in range(d.n_sub):
            if self.so:
                ii = np.arange(self.n) if idx is None else np.asarray(idx)
                om, z, vz = self.omega[ii], self.z[ii], self.v_z[ii]
                T = self.pid.thrust(ii, (f32(0.0) if hover else self.vz_sp) - vz)
                # PKG/attitude_controller.py:124-156 on one axis: torque -k_R sin(theta - theta_sp) - k_w omega on inertia J
                alpha = -(self.att_kr * det_sin_small(th - sp)) - self.att_kw * om
                om = (om + alpha * d.h).astype(np.float32)
                th = (th + om * d.h).astype(np.float32)
                a = (d.g * (T * self.inv_mg)) * det_sin_small(th) - d.c_d * v
                a_z = (T * det_cos_small(th)) * self.inv_m - self.g_abs
                z = ((z + vz * d.h) + a_z * d.half_h2).astype(np.float32)
                vz = (vz + a_z * d.h).astype(np.float32)
                self.omega[ii], self.z[ii], self.v_z[ii] = om, z, vz
            else:       # first-order lag and a = g tan(theta) - c_d v, as fused multiply-adds (fma32)
                th = fma32(sp - th, d.k_theta, th)
                a = fma32(-d.c_d, v, d.g * det_tan(th))
            x = fma32(a, d.half_h2, fma32(v, d.h, x))       # x + v h + a h^2 / 2 and v + a h, fused
            v = fma32(a, d.h, v)
            ph = ph + np.uint32(d.dphase)
            if self.kf is not None:       # one estimator sample per sub-step
                _, c = det_sincos_turns(ph)
                self.kf.sample(np.arange(self.n) if idx is None else idx, d.rw * c - v)
        self.x_d[sl], self.v_d[sl], self.theta[sl], self.phase[sl], self.a_d[sl] = x, v, th, ph, a

    def observe(self, step_count, idx=None):
        """(rel_p, rel_v, rel_a, pitch, z, contact) after `step_count` agent steps of the episode."""
        d = self.d
        sl = slice(None) if idx is None else idx
        s, c = det_sincos_turns(self.phase[sl])
        rel_p = fma32(d.r, s, -self.x_d[sl])          # platform minus drone, one rounding each
        rel_v = fma32(d.rw, c, -self.v_d[sl])
        rel_a = fma32(-d.rw2, s, -self.a_d[sl])
        if self.kf is not None:
            rel_a = self.kf.x[sl].copy()
        z = d.z_init + np.asarray(step_count).astype(np.float32) * d.dz
        if self.so:
            z = self.z[sl].copy()
        contact = (z <= d.z_touch) & (np.abs(rel_p) <= d.half_platform)
        return (rel_p.astype(np.float32), rel_v.astype(np.float32), rel_a.astype(np.float32),
                self.theta[sl].copy(), z.astype(np.float32), contact)


def standin_f64(params: StandInParams, x_d0, phase0, theta_sp_seq):
    """Float64 'textbook' trajectory for one env: returns arrays (rel_p, rel_v, rel_a, pitch, z)
    after each agent step, given the set-point applied at each step.  Same equations, np.sin/np.tan."""
    h = (1.0 / params.f_ag) / params.n_sub
    w = params.v_mp / params.r_mp
    k = -math.expm1(-h / params.tau_theta)
    dph = float(int(round(w * h / TWO_PI * 2.0 ** 32)) & 0xFFFFFFFF)
    x, v, th, ph = float(x_d0), 0.0, 0.0, float(phase0)
    out = []
    for i, sp in enumerate(theta_sp_seq):
        for _ in range(params.n_sub):
            th = th + (float(sp) - th) * k
            a = params.g * math.tan(th) - params.c_d * v
            x = x + v * h + 0.5 * a * h * h
            v = v + a * h
            ph = (ph + dph) % 2.0 ** 32
        ang = TWO_PI * ph / 2.0 ** 32
        xm, vm, am = params.r_mp * math.sin(ang), params.r_mp * w * math.cos(ang), -params.r_mp * w * w * math.sin(ang)
        z = params.z_init + (i + 1) * params.v_z * (1.0 / params.f_ag)
        out.append((xm - x, vm - v, am - a, th, z))
    return np.asarray(out, dtype=np.float64)


# ----------------------------------------------------------------------------------------
# two-axis stand-in: pitch drives x, roll drives y, one platform for both
# ----------------------------------------------------------------------------------------
TRAJ_RECTILINEAR_X, TRAJ_RECTILINEAR_XY, TRAJ_EIGHT = 0, 1, 2


@dataclass(frozen=True)
class StandIn2DParams:
    """Platform + axis conventions of the two-axis evaluator.
    trajectory 0: x = r_x sin(w_x t), y = 0                       (PKG/moving_platform.py:113-125, omega_y = 0)
    trajectory 1: x = r_x sin(w_x t), y = r_y sin(w_y t)          (the "possible future extension" of :113-125)
    trajectory 2: x = r_x cos(w t),  y = r_y sin(w t) cos(w t)    ("eight", PKG/moving_platform.py:92-111), w = v_x / r_x
    g_y is SIGNED: a_y = g_y tan(roll) - c_d v_y; in the reference's ENU frame a positive roll accelerates toward -y,
    so g_y = -g; the same holds for x with g_x = +g (sign verified in SURVEY.md A.3)."""
    base: StandInParams = StandInParams(v_z=-0.4)
    trajectory: int = TRAJ_RECTILINEAR_X
    r_y: float = 2.0
    v_y: float = 1.0             # PKG/moving_platform.py t_y default
    g_y: float = -9.81
    y_action_enabled: bool = False   # False = the reference: the roll branch of continuous_action is dead code (PKG/mdp.py:863-876)
    y_init_enabled: bool = False     # False = the reference: drone y = 0 * clip(...) (PKG/landing_simulation_env.py:336-340)


def derive_2d(p: StandIn2DParams):
    """fp32 platform constants of both axes: (dphase_x, dphase_y, r_x, rw_x, rw2_x, r_y, rw_y, rw2_y)."""
    b = p.base
    h = (1.0 / b.f_ag) / b.n_sub
    wx = b.v_mp / b.r_mp
    to_turns = lambda w: int(round(w * h / TWO_PI * 2.0 ** 32)) & 0xFFFFFFFF
    if p.trajectory == TRAJ_EIGHT:
        # d/dt [r_y sin cos] = r_y w (cos^2 - sin^2);  d2/dt2 = -4 r_y w^2 sin cos
        return (to_turns(wx), to_turns(wx), f32(b.r_mp), f32(b.r_mp * wx), f32(b.r_mp * wx * wx),
                f32(p.r_y), f32(p.r_y * wx), f32(4.0 * p.r_y * wx * wx))
    wy = p.v_y / p.r_y if p.trajectory == TRAJ_RECTILINEAR_XY else 0.0
    ry = p.r_y if p.trajectory == TRAJ_RECTILINEAR_XY else 0.0
    return (to_turns(wx), to_turns(wy), f32(b.r_mp), f32(b.r_mp * wx), f32(b.r_mp * wx * wx),
            f32(ry), f32(ry * wy), f32(ry * wy * wy))


class StandIn2D:
    """Deterministic fp32 two-axis stand-in, vectorised over n episodes; same operation order as eval2d_kernel."""

    def __init__(self, params: StandIn2DParams, n: int):
        self.p = params
        self.d = derive(params.base)
        (self.dphase_x, self.dphase_y, self.r_x, self.rw_x, self.rw2_x, self.r_y, self.rw_y, self.rw2_y) = derive_2d(params)
        self.g = (self.d.g, f32(params.g_y))
        z = lambda dt=f32: np.zeros(n, dt)
        self.pos, self.vel, self.ang, self.acc = [z(), z()], [z(), z()], [z(), z()], [z(), z()]
        self.phase = [z(np.uint32), z(np.uint32)]

    def platform(self):
        """((x_mp, u_mp, ax_mp), (y_mp, v_mp, ay_mp))"""
        sx, cx = det_sincos_turns(self.phase[0])
        if self.p.trajectory == TRAJ_EIGHT:
            sc = sx * cx
            return ((self.r_x * cx, -(self.rw_x * sx), -(self.rw2_x * cx)),
                    (self.r_y * sc, self.rw_y * (cx * cx - sx * sx), -(self.rw2_y * sc)))
        sy, cy = det_sincos_turns(self.phase[1])
        return ((self.r_x * sx, self.rw_x * cx, -(self.rw2_x * sx)), (self.r_y * sy, self.rw_y * cy, -(self.rw2_y * sy)))

    def reset(self, w0, w1, w2, w3):
        """PKG/landing_simulation_env.py:327-340: uniform offsets inside the fly zone, absolute clip; random platform phase."""
        d = self.d
        u = lambda w: (np.asarray(w, np.uint32) >> np.uint32(8)).astype(np.float32) * f32(2.0 ** -24)
        x_init = -d.p_max + d.two_p_max * u(w0)
        y_init = -d.p_max + d.two_p_max * u(w1)
        self.phase[0][:] = np.asarray(w2, np.uint32)
        self.phase[1][:] = np.asarray(w2 if self.p.trajectory == TRAJ_EIGHT else w3, np.uint32)
        (xm, _, _), (ym, _, _) = self.platform()
        self.pos[0][:] = np.clip(xm - x_init, -d.p_max, d.p_max)
        self.pos[1][:] = np.clip(ym - y_init, -d.p_max, d.p_max) if self.p.y_init_enabled else f32(0.0)
        for a in (0, 1):
            self.vel[a][:] = 0
            self.ang[a][:] = 0
            self.acc[a][:] = 0

    def advance(self, sp_x, sp_y):
        d = self.d
        sps = (np.asarray(sp_x, np.float32), np.asarray(sp_y, np.float32))
        for _ in range(d.n_sub):
            for a in (0, 1):
                th = self.ang[a] + (sps[a] - self.ang[a]) * d.k_theta
                acc = self.g[a] * det_tan(th) - d.c_d * self.vel[a]
                self.pos[a] = ((self.pos[a] + self.vel[a] * d.h) + acc * d.half_h2).astype(np.float32)
                self.vel[a] = (self.vel[a] + acc * d.h).astype(np.float32)
                self.ang[a], self.acc[a] = th.astype(np.float32), acc.astype(np.float32)
            self.phase[0] = self.phase[0] + np.uint32(self.dphase_x)
            self.phase[1] = self.phase[1] + np.uint32(self.dphase_y)

    def observe(self, step_count):
        """dict of fp32 arrays: rel_p/v/a for x and y, pitch, roll, z, contact."""
        d = self.d
        (xm, um, axm), (ym, vm, aym) = self.platform()
        o = dict(rel_p_x=xm - self.pos[0], rel_v_x=um - self.vel[0], rel_a_x=axm - self.acc[0], pitch=self.ang[0].copy(),
                 rel_p_y=ym - self.pos[1], rel_v_y=vm - self.vel[1], rel_a_y=aym - self.acc[1], roll=self.ang[1].copy())
        o = {k: np.asarray(v, np.float32) for k, v in o.items()}
        o["z"] = (d.z_init + np.asarray(step_count).astype(np.float32) * d.dz).astype(np.float32)
        o["contact"] = (o["z"] <= d.z_touch) & (np.abs(o["rel_p_x"]) <= d.half_platform) & (np.abs(o["rel_p_y"]) <= d.half_platform)
        return o


def sim2d_cases():
    """Named two-axis cases shared by the fixture generator (oracle/gen_golden.py) and the parity tests."""
    return {
        "reference": StandIn2DParams(trajectory=TRAJ_RECTILINEAR_X),
        "xy": StandIn2DParams(trajectory=TRAJ_RECTILINEAR_XY, y_action_enabled=True, y_init_enabled=True, v_y=1.0),
        "eight": StandIn2DParams(base=StandInParams(v_z=-0.4, r_mp=3.0, v_mp=0.8), trajectory=TRAJ_EIGHT, r_y=3.0,
                                 y_action_enabled=True, y_init_enabled=True),
        # the x policy on the y axis has the wrong sign for a_y = -g tan(roll): the drone runs away along y (FLYZONE_Y)
        "ywrong": StandIn2DParams(trajectory=TRAJ_RECTILINEAR_XY, y_action_enabled=True, y_init_enabled=True, v_y=1.0),
    }
