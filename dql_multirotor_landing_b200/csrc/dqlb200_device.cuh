// Device-side building blocks of libdqlb200: Philox4x32-10, the deterministic fp32 math used by the
// stand-in dynamics, the cut-table discretisation, terminal checks and the float64 reward.
//
// Reference rows: R3 continuous_action PKG/mdp.py:543-560, R4 analytic stand-in
// (no reference function), R5 discrete_state PKG/mdp.py:257-333, R6 check PKG/mdp.py:335-439,
// R7 reward PKG/mdp.py:441-541, R9 guess/predict PKG/double_q_learning.py:110-124.
//
// Bit-exactness rules used throughout:
//   * fp32 dynamics: every operation is an explicit __fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn/__fmaf_rn (the
//     compiler never contracts on its own: --fmad=false); the Horner steps of sin/cos/tan are explicit fused
//     multiply-adds -> identical to NumPy float32 (with an exact FMA emulation) and to C with -ffp-contract=off;
//   * float64 reward: explicit __dmul_rn/__dadd_rn/__ddiv_rn in the reference's operation order;
//   * comparisons on fp32 observations use host-computed cut points (constants.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dqlb200.h"

namespace dql {

// Compact copy of dqlb200_config passed to kernels by value (constant bank).
struct KC {
  dqlb200_cuts cuts[DQLB200_MAX_CURRICULUM];
  dqlb200_reward_level reward[DQLB200_MAX_CURRICULUM];
  double p_max, v_max, theta_max, delta_theta, w_p, w_v, w_theta;
  double rcp_p_max, rcp_v_max, rcp_theta_max;   // RN(1/x) for div_f32_by_const / div_f64_by_const
  float angle_cut[6];
  float fz_lo, fz_hi, z_min_cut, z_max_cut;
  float h, half_h2, k_theta, g, c_d, dz_train, dz_sim, z_init, z_touch, half_platform;
  float p_max_f, two_p_max_f, sigma_x;
  float clip_p_f, clip_v_f;     // smallest fp32 |x| whose quotient x / p_max (x / v_max) rounds to >= 1
  float noise_pos_sd, noise_vel_sd;
  int32_t noise_enabled;        // either standard deviation is non-zero
  int32_t accel_mode;           // 0 exact, 1 / 2 Kalman-filtered finite difference (dqlb200_config.accel_mode)
  float kf_q, kf_r;
  // second-order attitude + vertical PID model (dqlb200_config.dynamics_model == 1)
  int32_t dynamics_model, pid_ticks;
  float att_kr, att_kw;         // k_R / J, k_omega / J
  float inv_m, inv_mg, g_abs;   // 1 / m, 1 / (m g), |g|
  float pid_kp, pid_ki, pid_lo, pid_hi, pid_windup, pid_dt, pid_i0;
  float bw_inv_denom, bw_k2;    // ButterworthFilter (c = 1): 1 / (1 + c^2 + 1.414 c), c^2 - 1.414 c + 1
  float vz_train, vz_sim;       // vertical-velocity set-points (PKG/mdp.py:212, 580)
  float gamma;
  float transfer_ratio[DQLB200_MAX_CURRICULUM];
  int32_t timeout_steps, success_steps, n_sub;
  int32_t transfer_mode, window_len, promote_successes;
  int32_t curriculum_steps, envs_per_population, n_populations;
  int32_t replicas;             // > 1: replica-merge mode, promotion is decided by replica_merge_kernel
  int32_t div_two_steps;        // 0 only for the exhaustively verified default divisors
  long long max_num_episodes;
};

// The reference-default configuration as COMPILE-TIME constants (MdpParameters / DynamicsParameters defaults of constants.py;
// PKG/mdp.py:214-235, PKG/trainer.py:41-44).  sm_100a has no constant-bank operands: a run-time constant costs
// a uniform-register load next to its use (about 70 of them per env-step); literals are encoded in the instructions.  The
// production instance of train_kernel uses this type; dqlb200_create checks every value below against the run-time
// configuration bit for bit and falls back to the generic instance (run-time KC) when anything differs.
struct KDef {
  static constexpr float h = 0x1.656ac8p-5f, half_h2 = 0x1.f302fcp-11f, k_theta = 0x1.0d7ec4p-2f, c_d = 0x1.99999ap-3f;
  static constexpr float dz_train = -0x1.1def06p-8f, dz_sim = -0x1.1def06p-6f, z_init = 4.0f, z_touch = 0x1.07ae14p-1f;
  static constexpr float half_platform = 0.5f, p_max_f = 4.5f, two_p_max_f = 9.0f, sigma_x = 1.5f, gamma = 0x1.fae148p-1f;
  static constexpr float fz_lo = -4.5f, fz_hi = 0x1.200002p+2f, z_min_cut = 0x1.99999ap-3f, z_max_cut = 0x1.200002p+2f;
  static constexpr float clip_p_f = 4.5f, clip_v_f = 0x1.b27234p+1f;
  static constexpr double p_max = 4.5, v_max = 0x1.b272324c83665p+1, theta_max = 0x1.7e0eb9bca0a66p-2, delta_theta = 0x1.fd68e8083ba2ep-4;
  static constexpr double w_p = -100.0, w_v = -10.0, w_theta = -0x1.8cccccccccccdp+0;
  static constexpr double rcp_p_max = 1.0 / p_max, rcp_v_max = 1.0 / v_max, rcp_theta_max = 1.0 / theta_max;
  static constexpr int32_t timeout_steps = 459, success_steps = 23, n_sub = 1;
  static constexpr int32_t noise_enabled = 0, accel_mode = 0;
  static constexpr float kf_q = 0.0f, kf_r = 0.0f;
  static constexpr int32_t dynamics_model = 0, pid_ticks = 0;
  static constexpr float att_kr = 0.0f, att_kw = 0.0f, inv_m = 0.0f, inv_mg = 0.0f, g_abs = 0.0f, pid_kp = 0.0f, pid_ki = 0.0f, pid_lo = 0.0f,
                         pid_hi = 0.0f, pid_windup = 0.0f, pid_dt = 0.0f, pid_i0 = 0.0f, bw_inv_denom = 0.0f, bw_k2 = 0.0f, vz_train = 0.0f, vz_sim = 0.0f;
  static constexpr float noise_pos_sd = 0.0f, noise_vel_sd = 0.0f;
  struct AngleCut {          // constant-index reads fold into literals
    __host__ __device__ constexpr float operator[](int i) const {
      return i == 0 ? -0x1.3e619ap-2f : i == 1 ? -0x1.7e0eb8p-3f : i == 2 ? -0x1.fd68f6p-5f : i == 3 ? 0x1.fd68f8p-5f : i == 4 ? 0x1.7e0ebap-3f : 0x1.3e619cp-2f;
    }
  };
  AngleCut angle_cut;       // an (empty) member: passed by reference to discretise_cuts
};
// true when every compile-time value of KDef equals the run-time configuration (host side, dqlb200_create)
inline bool kdef_matches(const KC& k) {
  bool ok = k.h == KDef::h && k.half_h2 == KDef::half_h2 && k.k_theta == KDef::k_theta && k.c_d == KDef::c_d && k.dz_train == KDef::dz_train &&
            k.dz_sim == KDef::dz_sim && k.z_init == KDef::z_init && k.z_touch == KDef::z_touch && k.half_platform == KDef::half_platform &&
            k.p_max_f == KDef::p_max_f && k.two_p_max_f == KDef::two_p_max_f && k.sigma_x == KDef::sigma_x && k.gamma == KDef::gamma &&
            k.fz_lo == KDef::fz_lo && k.fz_hi == KDef::fz_hi && k.z_min_cut == KDef::z_min_cut && k.z_max_cut == KDef::z_max_cut &&
            k.clip_p_f == KDef::clip_p_f && k.clip_v_f == KDef::clip_v_f && k.p_max == KDef::p_max && k.v_max == KDef::v_max &&
            k.theta_max == KDef::theta_max && k.delta_theta == KDef::delta_theta && k.w_p == KDef::w_p && k.w_v == KDef::w_v &&
            k.w_theta == KDef::w_theta && k.rcp_p_max == KDef::rcp_p_max && k.rcp_v_max == KDef::rcp_v_max && k.rcp_theta_max == KDef::rcp_theta_max &&
            k.timeout_steps == KDef::timeout_steps && k.success_steps == KDef::success_steps && k.n_sub == KDef::n_sub && k.noise_enabled == 0 &&
            k.accel_mode == 0 && k.dynamics_model == 0 && k.div_two_steps == 0;
  for (int i = 0; i < 6; ++i) ok = ok && k.angle_cut[i] == KDef::AngleCut{}[i];
  return ok;
}

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011).  Counter = (env, step, purpose, population), key = seed.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c.x;   // one IMAD.WIDE each
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c.z;
    c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k0, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k1, (uint32_t)p0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}
// The same with the round keys precomputed (k0 + i * 0x9E3779B9, k1 + i * 0xBB67AE85 for i = 0..9, as five uint4 {k0_2j, k1_2j,
// k0_2j+1, k1_2j+1}): the key is per population, so a CTA of the training kernel computes the schedule once and keeps it in
// shared memory -- 5 vector loads per call instead of 18 additions.
__device__ __forceinline__ void philox_round_keys(uint32_t k0, uint32_t k1, uint4* keys) {
#pragma unroll
  for (int j = 0; j < 5; ++j)
    keys[j] = make_uint4(k0 + (uint32_t)(2 * j) * 0x9E3779B9u, k1 + (uint32_t)(2 * j) * 0xBB67AE85u,
                         k0 + (uint32_t)(2 * j + 1) * 0x9E3779B9u, k1 + (uint32_t)(2 * j + 1) * 0xBB67AE85u);
}
__device__ __forceinline__ uint4 philox4x32_10_keyed(uint4 c, const uint4* __restrict__ keys) {
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const uint4 k = keys[j];
    unsigned long long p0 = (unsigned long long)0xD2511F53u * c.x, p1 = (unsigned long long)0xCD9E8D57u * c.z;
    c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k.x, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.y, (uint32_t)p0);
    p0 = (unsigned long long)0xD2511F53u * c.x; p1 = (unsigned long long)0xCD9E8D57u * c.z;
    c = make_uint4((uint32_t)(p1 >> 32) ^ c.y ^ k.z, (uint32_t)p1, (uint32_t)(p0 >> 32) ^ c.w ^ k.w, (uint32_t)p0);
  }
  return c;
}
constexpr uint32_t PURPOSE_STEP = 0u, PURPOSE_RESET = 1u, PURPOSE_RESET_NOISE = 2u;

// ---------------------------------------------------------------------------------------------
// deterministic fp32 math
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void det_sincos_turns(uint32_t phase, float& s_out, float& c_out) {
  const uint32_t q = (phase + 0x20000000u) >> 30;
  const int32_t rem = (int32_t)(phase - (q << 30));
  const float x = fmul(__int2float_rn(rem), (float)(6.283185307179586 / 4294967296.0));
  const float z = fmul(x, x);
  // Horner steps as fused multiply-adds (FFMA, one rounding each): the stand-in is DEFINED that way (oracle/dynamics.py: fma32,
  // oracle/c/standin.c: fmaf) -- a separately rounded product would cost one more instruction per step
  float ps = (float)(1.0 / 362880.0);
  ps = __fmaf_rn(ps, z, (float)(-1.0 / 5040.0));
  ps = __fmaf_rn(ps, z, (float)(1.0 / 120.0));
  ps = __fmaf_rn(ps, z, (float)(-1.0 / 6.0));
  const float s = __fmaf_rn(x, fmul(z, ps), x);
  float pc = (float)(-1.0 / 3628800.0);
  pc = __fmaf_rn(pc, z, (float)(1.0 / 40320.0));
  pc = __fmaf_rn(pc, z, (float)(-1.0 / 720.0));
  pc = __fmaf_rn(pc, z, (float)(1.0 / 24.0));
  pc = __fmaf_rn(pc, z, -0.5f);
  const float c = __fmaf_rn(z, pc, 1.0f);
  // quadrant q: (sin, cos) = (s, c), (c, -s), (-s, -c), (-c, s) -- branch-free: swap on odd q, then sign bits
  const bool odd = (q & 1u) != 0u;
  const uint32_t sb = __float_as_uint(odd ? c : s), cb = __float_as_uint(odd ? s : c);
  s_out = __uint_as_float(sb ^ ((q & 2u) << 30));
  c_out = __uint_as_float(cb ^ (((q + 1u) & 2u) << 30));
}

__device__ __forceinline__ float det_tan(float x) {
  const float z = fmul(x, x);
  float p = (float)(21844.0 / 6081075.0);
  p = __fmaf_rn(p, z, (float)(1382.0 / 155925.0));
  p = __fmaf_rn(p, z, (float)(62.0 / 2835.0));
  p = __fmaf_rn(p, z, (float)(17.0 / 315.0));
  p = __fmaf_rn(p, z, (float)(2.0 / 15.0));
  p = __fmaf_rn(p, z, (float)(1.0 / 3.0));
  return __fmaf_rn(x, fmul(z, p), x);
}

// sin / cos of a small angle in radians (|x| <= ~0.8: attitude angles and attitude errors), Taylor polynomials in Horner form
__device__ __forceinline__ float det_sin_small(float x) {
  const float z = fmul(x, x);
  float p = (float)(1.0 / 362880.0);
  p = fadd(fmul(p, z), (float)(-1.0 / 5040.0));
  p = fadd(fmul(p, z), (float)(1.0 / 120.0));
  p = fadd(fmul(p, z), (float)(-1.0 / 6.0));
  return fadd(x, fmul(x, fmul(z, p)));
}
__device__ __forceinline__ float det_cos_small(float x) {
  const float z = fmul(x, x);
  float p = (float)(-1.0 / 3628800.0);
  p = fadd(fmul(p, z), (float)(1.0 / 40320.0));
  p = fadd(fmul(p, z), (float)(-1.0 / 720.0));
  p = fadd(fmul(p, z), (float)(1.0 / 24.0));
  p = fadd(fmul(p, z), -0.5f);
  return fadd(1.0f, fmul(z, p));
}

__device__ __forceinline__ float det_log(float u) {
  const uint32_t bits = __float_as_uint(u);
  int e = (int)((bits >> 23) & 0xFFu) - 127;
  float m = __uint_as_float((bits & 0x007FFFFFu) | 0x3F800000u);
  if (m > (float)1.4142135623730951) {
    m = fmul(m, 0.5f);
    e += 1;
  }
  const float s = __fdiv_rn(fsub(m, 1.0f), fadd(m, 1.0f));
  const float z = fmul(s, s);
  float p = (float)(1.0 / 9.0);
  p = fadd(fmul(p, z), (float)(1.0 / 7.0));
  p = fadd(fmul(p, z), (float)(1.0 / 5.0));
  p = fadd(fmul(p, z), (float)(1.0 / 3.0));
  const float lm = fmul(fadd(s, s), fadd(1.0f, fmul(z, p)));
  return fadd(fmul(__int2float_rn(e), (float)0.6931471805599453), lm);
}

__device__ __forceinline__ float det_normal(uint32_t x0, uint32_t x1) {
  const float u1 = fmul(fadd(__uint2float_rn(x0 >> 8), 1.0f), (float)(1.0 / 16777216.0));
  const float rad = __fsqrt_rn(fmul(-2.0f, det_log(u1)));
  float s, c;
  det_sincos_turns(x1, s, c);
  return fmul(rad, c);
}

// Both Box-Muller normals of one (u1, angle) pair.
__device__ __forceinline__ void det_normal_pair(uint32_t x0, uint32_t x1, float& n0, float& n1) {
  const float u1 = fmul(fadd(__uint2float_rn(x0 >> 8), 1.0f), (float)(1.0 / 16777216.0));
  const float rad = __fsqrt_rn(fmul(-2.0f, det_log(u1)));
  float s, c;
  det_sincos_turns(x1, s, c);
  n0 = fmul(rad, c);
  n1 = fmul(rad, s);
}

__device__ __forceinline__ float clipf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ double clipd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
// np.clip for a value that is never NaN (fmin/fmax cost ~6 instructions each for their NaN rules; this is 2 compares + selects)
__device__ __forceinline__ double clipd_finite(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ---------------------------------------------------------------------------------------------
// stand-in dynamics (R4).  Body state: drone position/velocity, pitch, platform phase.
// ---------------------------------------------------------------------------------------------
struct Body {
  float x_d, v_d, theta, a_d;
  uint32_t phase;
};

struct Obs {
  float rel_p, rel_v, rel_a, pitch, z;
  bool contact;
};

// Acceleration estimator of the reference's observation node: KalmanFilter1D (PKG/filters.py:4-37) fed with a
// finite difference of the TRUE relative velocity (PKG/filters.py:54-80, PKG/observation_utils.py:134-150).  fp32, every operation
// rounded separately (oracle/dynamics.py: KalmanAccel is the same arithmetic in NumPy float32).
struct Kf {
  float x, P, v_ref;      // estimate, variance, anchor (mode 1: first sample ever) or previous (mode 2) relative velocity
  uint32_t n;             // samples seen (mode 2: 0 or 1)
};
template <class KT>
__device__ __forceinline__ void kf_sample(const KT& kc, Kf& f, float rel_v) {
  if (f.n == 0u) {        // PKG/observation_utils.py:137-143: the first sample only sets the anchor, the filter is not touched
    f.v_ref = rel_v;
    f.n = 1u;
    return;
  }
  const float dt = (kc.accel_mode == 1) ? fmul(__uint2float_rn(f.n), kc.h) : kc.h;
  const float raw = __fdiv_rn(fsub(rel_v, f.v_ref), dt);
  f.P = fadd(f.P, kc.kf_q);                                   // PKG/filters.py:31-35
  const float K = __fdiv_rn(f.P, fadd(f.P, kc.kf_r));
  f.x = fadd(f.x, fmul(K, fsub(raw, f.x)));
  f.P = fmul(f.P, fsub(1.0f, K));
  if (kc.accel_mode == 1) f.n += 1u;      // quirk Q13: the anchor is never refreshed, the time base keeps growing
  else f.v_ref = rel_v;
}

// Second-order attitude + vertical PID model, the state the first-order stand-in does not have:
// pitch (roll) rate, altitude, vertical velocity, and the PID node's memory (integral, the Butterworth filter's previous input
// and three previous outputs).  The PID memory belongs to the node: like the acceleration estimator it survives episode resets.
// The filter's two previous inputs are equal at every sub-step boundary (the error is held for pid_ticks >= 2 node
// iterations), so one word holds both.
struct Ext {
  float omega, z, v_z, integ;
  float e1, f1, f2, f3;
};
// One sub-step of the vertical loop (PKG/pid.py:62-104 with the gains of launch/drone.launch:33-46, Kd = 0): the node runs
// `pid_ticks` times per sub-step on the held error (1 kHz against the 100 Hz state topic); returns the thrust [N].
template <class KT>
__device__ __forceinline__ float pid_thrust(const KT& kc, Ext& x, float vz_sp) {
  const float e = fsub(vz_sp, x.v_z);
  float thrust = 0.0f, e1 = x.e1, e2 = x.e1;
  for (int k = 0; k < kc.pid_ticks; ++k) {
    x.integ = clipf(fadd(x.integ, fmul(e, kc.pid_dt)), -kc.pid_windup, kc.pid_windup);            // PKG/pid.py:85-86
    // ButterworthFilter.update (PKG/filters.py:96-108) with c = 1, as written: the new input is pushed BEFORE the sum, the new
    // output after it, so the output taps are one sample older than the input taps (quirk Q14): y_k = (x_k-2 + 2 x_k-1 + x_k
    // - (c^2 - 1.414 c + 1) y_k-3 - (-2 c^2 + 2) y_k-2) / denom, and the last coefficient is zero
    const float f = fmul(kc.bw_inv_denom, fsub(fadd(fadd(e2, fmul(2.0f, e1)), e), fmul(kc.bw_k2, x.f3)));
    e2 = e1; e1 = e;
    x.f3 = x.f2; x.f2 = x.f1; x.f1 = f;
    thrust = clipf(fadd(fmul(kc.pid_kp, f), fmul(kc.pid_ki, x.integ)), kc.pid_lo, kc.pid_hi);    // PKG/pid.py:97-103
  }
  x.e1 = e;
  return thrust;
}

template <class KT>
__device__ __forceinline__ void dyn_advance(const KT& kc, const dqlb200_population_params& pp, Body& b, float sp, Kf* kf = nullptr,
                                            Ext* ext = nullptr, float vz_sp = 0.0f) {
  for (int i = 0; i < kc.n_sub; ++i) {
    if (kc.dynamics_model != 0 && ext) {
      Ext& x = *ext;
      const float thrust = pid_thrust(kc, x, vz_sp);
      // geometric attitude controller reduced to one axis (PKG/attitude_controller.py:124-156): M = -k_R sin(theta - theta_sp)
      // - k_w omega is applied as a torque (:107-113), theta'' = M / J; semi-implicit Euler
      const float alpha = fsub(-fmul(kc.att_kr, det_sin_small(fsub(b.theta, sp))), fmul(kc.att_kw, x.omega));
      x.omega = fadd(x.omega, fmul(alpha, kc.h));
      b.theta = fadd(b.theta, fmul(x.omega, kc.h));
      // thrust along the body z axis: horizontal and vertical components (pp.g carries the sign of the axis)
      b.a_d = fsub(fmul(fmul(pp.g, fmul(thrust, kc.inv_mg)), det_sin_small(b.theta)), fmul(kc.c_d, b.v_d));
      const float a_z = fsub(fmul(fmul(thrust, det_cos_small(b.theta)), kc.inv_m), kc.g_abs);
      x.z = fadd(fadd(x.z, fmul(x.v_z, kc.h)), fmul(a_z, kc.half_h2));
      x.v_z = fadd(x.v_z, fmul(a_z, kc.h));
    } else {
      // first-order lag, a = g tan(theta) - c_d v, x + v h + a h^2 / 2, v + a h: fused multiply-adds by definition (the oracles
      // say fma32 / fmaf), one FFMA where a separately rounded product would be two instructions
      b.theta = __fmaf_rn(fsub(sp, b.theta), kc.k_theta, b.theta);
      b.a_d = __fmaf_rn(-kc.c_d, b.v_d, fmul(pp.g, det_tan(b.theta)));
    }
    b.x_d = __fmaf_rn(b.a_d, kc.half_h2, __fmaf_rn(b.v_d, kc.h, b.x_d));
    b.v_d = __fmaf_rn(b.a_d, kc.h, b.v_d);
    b.phase += pp.dphase;
    if (kc.accel_mode != 0 && kf) {       // one estimator sample per sub-step (the node publishes at the sub-step rate)
      float s, c;
      det_sincos_turns(b.phase, s, c);
      kf_sample(kc, *kf, fsub(fmul(pp.rw, c), b.v_d));
    }
  }
}

template <class KT>
__device__ __forceinline__ Obs dyn_observe(const KT& kc, const dqlb200_population_params& pp, const Body& b,
                                           int step_count, float dz, const Kf* kf = nullptr, const Ext* ext = nullptr) {
  float s, c;
  det_sincos_turns(b.phase, s, c);
  Obs o;
  o.rel_p = __fmaf_rn(pp.r, s, -b.x_d);
  o.rel_v = __fmaf_rn(pp.rw, c, -b.v_d);
  o.rel_a = __fmaf_rn(-pp.rw2, s, -b.a_d);
  if (kc.accel_mode != 0 && kf) o.rel_a = kf->x;
  o.pitch = b.theta;
  o.z = fadd(kc.z_init, fmul(__int2float_rn(step_count), dz));
  if (kc.dynamics_model != 0 && ext) o.z = ext->z;
  o.contact = (o.z <= kc.z_touch) && (fabsf(o.rel_p) <= kc.half_platform);
  return o;
}

// Observation noise (PKG/observation_utils.py:127-129): what the MDP sees is the true relative position / velocity plus
// independent Gaussians; contact (a bumper in the reference) and the physical state are untouched.
template <class KT>
__device__ __forceinline__ void add_observation_noise(const KT& kc, Obs& o, uint32_t w0, uint32_t w1) {
  float n0, n1;
  det_normal_pair(w0, w1, n0, n1);
  o.rel_p = fadd(o.rel_p, fmul(kc.noise_pos_sd, n0));
  o.rel_v = fadd(o.rel_v, fmul(kc.noise_vel_sd, n1));
}

// R1 (PKG/landing_simulation_env.py:181-216) and R15 (:327-340), then one hover period (:222-224).
template <class KT>
__device__ __forceinline__ Obs dyn_reset(const KT& kc, const dqlb200_population_params& pp, Body& b, uint4 w,
                                         bool normal_init, bool simulation, float dz, Kf* kf = nullptr, Ext* ext = nullptr) {
  float x_init;
  if (normal_init) {
    x_init = fmul(kc.sigma_x, det_normal(w.x, w.y));
  } else {
    const float u = fmul(__uint2float_rn(w.x >> 8), (float)(1.0 / 16777216.0));
    x_init = fadd(-kc.p_max_f, fmul(kc.two_p_max_f, u));
  }
  b.phase = w.z;
  float s, c;
  det_sincos_turns(b.phase, s, c);
  const float x_mp = fmul(pp.r, s);
  b.x_d = simulation ? clipf(fsub(x_mp, x_init), -kc.p_max_f, kc.p_max_f)
                     : fadd(x_mp, clipf(x_init, -kc.p_max_f, kc.p_max_f));
  b.v_d = 0.0f;
  b.theta = 0.0f;
  b.a_d = 0.0f;
  if (kc.dynamics_model != 0 && ext) {    // teleport with zero twist (PKG/landing_simulation_env.py:203-216); the PID memory stays
    ext->omega = 0.0f;
    ext->z = kc.z_init;
    ext->v_z = 0.0f;
  }
  dyn_advance(kc, pp, b, 0.0f, kf, ext, 0.0f);       // hover period: all set-points zero (scripts/manager_node.py:328)
  return dyn_observe(kc, pp, b, 0, dz, kf, ext);
}

// ---------------------------------------------------------------------------------------------
// R5: discretisation through the fp32 cut tables of the current working step
// ---------------------------------------------------------------------------------------------
struct DState {
  int level, bp, bv, ba, bt;
  __device__ __forceinline__ int id() const { return (((level * 3 + bp) * 3 + bv) * 3 + ba) * 7 + bt; }
};

template <class AC>
__device__ __forceinline__ DState discretise_cuts(const dqlb200_cuts& c, const AC& angle_cut,
                                                  const Obs& o, int w = 4) {
  int lp = 0, lv = 0;
  if (w > 0) {         // one uniform branch at working step 0 (the only level there); w is CTA-uniform
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i < w) {     // levels above the working step do not exist (their cuts are NaN): skip
        lp += (o.rel_p >= c.lvl_lo[0][i]) && !(o.rel_p >= c.lvl_hi[0][i]);
        lv += (o.rel_v >= c.lvl_lo[1][i]) && !(o.rel_v >= c.lvl_hi[1][i]);
      }
    }
  }
  DState d;
  d.level = min(lp, lv);   // the acceleration limit is 1.0 at every level and never restricts
  d.bp = (o.rel_p >= c.bin1[0][d.level]) + (o.rel_p >= c.bin2[0][d.level]);
  d.bv = (o.rel_v >= c.bin1[1][d.level]) + (o.rel_v >= c.bin2[1][d.level]);
  d.ba = (o.rel_a >= c.bin1[2][d.level]) + (o.rel_a >= c.bin2[2][d.level]);
  int t = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i) t += (o.pitch >= angle_cut[i]);
  d.bt = t;
  return d;
}

// R3: pitch set-point integrator, float64 like the reference
template <class KT>
__device__ __forceinline__ double apply_action(const KT& kc, double theta_sp, int a) {
  // branch-free (the three actions diverge inside a warp): a = 2 adds 0.0 and the clamps are no-ops for |theta_sp| <= theta_max
  const double s = __dadd_rn(theta_sp, a == 0 ? kc.delta_theta : (a == 1 ? -kc.delta_theta : 0.0));
  // selects on the 32-bit halves (a select on a double may be compiled into a branch)
  const bool hi = s > kc.theta_max, lo = s < -kc.theta_max;
  const int t_hi = __double2hiint(kc.theta_max), t_lo = __double2loint(kc.theta_max);
  const int r_hi = hi ? t_hi : (lo ? (int)((uint32_t)t_hi ^ 0x80000000u) : __double2hiint(s));
  const int r_lo = (hi || lo) ? t_lo : __double2loint(s);
  return __hiloint2double(r_hi, r_lo);
}

// Correctly rounded (double)x / d for a finite fp32 numerator x and a constant divisor d, rcp = RN(1/d):
// q0 = RN(x*rcp);  r = x - q0*d (exact in one FMA);  q = RN(q0 + r*rcp)   (Markstein's correction step).
// Three instructions instead of the ~25 + slow path of the generic __ddiv_rn (which is taken for zero
// numerators).  For the reference's divisors (p_max = 4.5, v_max = 3.39411) one correction step is verified
// EXHAUSTIVELY on the device against __ddiv_rn for all finite fp32 numerators (dqlb200_selftest_division,
// tests/test_gpu_parity.py); for any other divisor a second step is added (`two_steps`), which is exact
// whenever the first step is faithful.  |q| is only ever used through fabs/clip, so the sign of a zero
// quotient is irrelevant.
__device__ __forceinline__ double div_f32_by_const(float x, double d, double rcp, bool two_steps) {
  const double xd = (double)x;
  double q = __dmul_rn(xd, rcp);
  q = __fma_rn(__fma_rn(-q, d, xd), rcp, q);
  if (two_steps) q = __fma_rn(__fma_rn(-q, d, xd), rcp, q);
  return q;
}

// Correctly rounded x / d for a float64 numerator of ordinary magnitude (|x| in [2^-900, 2^900] or zero) and a
// constant divisor d with rcp = RN(1/d): q0 = RN(x*rcp) has relative error < 2^-52; the first correction step
// (exact residual in one FMA) makes it faithful, the second one correctly rounded (Markstein 1990, Cornea et al.
// 1999: q' = RN(q + r*y) = RN(a/b) when y = RN(1/b) and q is faithful).  Five instructions, no slow path for
// zero numerators (the generic __ddiv_rn takes a ~90-instruction subroutine for them, and set-points and
// set-point differences are zero most of the time).  Cross-checked against __ddiv_rn on 2^32 random numerators by
// dqlb200_selftest_division.  Callers use |q| or add q to a non-zero sum: the sign of a zero quotient is irrelevant.
__device__ __forceinline__ double div_f64_by_const(double x, double d, double rcp) {
  double q = __dmul_rn(x, rcp);
  q = __fma_rn(__fma_rn(-q, d, x), rcp, q);
  q = __fma_rn(__fma_rn(-q, d, x), rcp, q);
  return q;
}

// Shaping potential of one fp32 observation (PKG/mdp.py:457-474): w * |clip(x / x_max, -1, 1)|
__device__ __forceinline__ double shaping(double w, float x, double x_max, double rcp, float clip_cut, bool two_steps) {
  // |clip(q, -1, 1)| = min(|q|, 1), and the quotient reaches 1 exactly when |x| >= clip_cut (host-computed, monotone
  // rounding): the clip is one fp32 compare and a select of the halves.  NaN gives 1 like fmin/fmax would.
  const double aq = fabs(div_f32_by_const(x, x_max, rcp, two_steps));
  const bool sat = !(fabsf(x) < clip_cut);
  return __dmul_rn(w, __hiloint2double(sat ? 0x3FF00000 : __double2hiint(aq), sat ? 0 : __double2loint(aq)));
}

// R7 with the level-dependent constants pre-evaluated on the host.
template <class KT>
__device__ __forceinline__ double reward_f64(const KT& kc, const dqlb200_reward_level& rl, double phi_p,
                                             double phi_v, double phi_t, double prev_p, double prev_v,
                                             double prev_t, bool success) {
  const double r_p = clipd_finite(__dsub_rn(phi_p, prev_p), -rl.r_p_max, rl.r_p_max);
  const double r_v = clipd_finite(__dsub_rn(phi_v, prev_v), -rl.r_v_max, rl.r_v_max);
  const double r_t =
      __dmul_rn(div_f64_by_const(__dmul_rn(kc.w_theta, __dsub_rn(fabs(phi_t), fabs(prev_t))), kc.theta_max, kc.rcp_theta_max), rl.lim_v);
  const double r_term = success ? rl.r_term_succ : rl.r_term_fail;
  return __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(r_p, r_v), r_t), rl.r_dur), r_term);
}

// R7 with the set-point term taken from the configuration's table (dqlb200_config.setpoint_rtheta: w_theta * (|phi'| - |phi|) /
// theta_max evaluated on the host in the reference's operation order); same sum order as reward_f64.
__device__ __forceinline__ double reward_sp(const dqlb200_reward_level& rl, double phi_p, double phi_v, double prev_p, double prev_v,
                                            double r_theta0, bool success) {
  const double r_p = clipd_finite(__dsub_rn(phi_p, prev_p), -rl.r_p_max, rl.r_p_max);
  const double r_v = clipd_finite(__dsub_rn(phi_v, prev_v), -rl.r_v_max, rl.r_v_max);
  const double r_t = __dmul_rn(r_theta0, rl.lim_v);
  const double r_term = success ? rl.r_term_succ : rl.r_term_fail;
  return __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(r_p, r_v), r_t), rl.r_dur), r_term);
}

}  // namespace dql
