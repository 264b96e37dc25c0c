/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
 *
 * Plain-C restatement of the deterministic numeric core of the hot path, the third statement of the same arithmetic
 * next to oracle/dynamics.py + oracle/philox.py (NumPy float32) and the CUDA kernels (csrc/dqlb200_device.cuh):
 *   - Philox4x32-10 (Salmon et al. 2011) and the draw contract counter = (env, step, purpose, population), key = seed
 *   - the fp32 building blocks: sin/cos of a uint32 "turns" phase, tan (Horner in fused multiply-adds), log, Box-Muller normal (Horner, no FMA)
 *   - the analytic stand-in (there is no reference function for it): first-order pitch lag,
 *     a = g tan(theta) - c_d v, platform x = r sin(w t) (PKG/moving_platform.py:116-125), rel = platform - drone
 *     (PKG/observation_utils.py:225,249)
 *   - the acceleration estimator (PKG/filters.py:4-80 as PKG/observation_utils.py:134-150 drives it)
 *   - the second-order model: geometric-controller torque on the inertia (PKG/attitude_controller.py:124-156), the v_z PID
 *     node with its Butterworth filter (PKG/pid.py:62-104, PKG/filters.py:83-108)
 * Compiled with -ffp-contract=off -fno-fast-math: every operation is one IEEE-754 rounding, so the three statements
 * agree BIT FOR BIT (tests/test_oracle_c.py).  Built by oracle/c/Makefile into oracle/c/libstandin_oracle.so.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  float h, half_h2, k_theta, g, c_d, r, rw, rw2;
  uint32_t dphase;
  int32_t n_sub;
  /* estimator: 0 exact, 1 anchor quirk ("kalman_reference"), 2 consecutive samples ("kalman") */
  int32_t accel_mode;
  float kf_q, kf_r;
  /* second-order model */
  int32_t second_order, pid_ticks;
  float att_kr, att_kw, inv_m, inv_mg, g_abs, pid_kp, pid_ki, pid_lo, pid_hi, pid_windup, pid_dt, bw_inv_denom, bw_k2, z_init;
} standin_params;

typedef struct {
  float x_d, v_d, theta, a_d;
  uint32_t phase;
  float kf_x, kf_P, kf_vref;
  uint32_t kf_n;
  float omega, z, v_z, integ, e1, f1, f2, f3;
} standin_state;

void oracle_philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  for (int i = 0; i < 10; ++i) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void oracle_sincos_turns(uint32_t phase, float* s_out, float* c_out) {
  const uint32_t q = (phase + 0x20000000u) >> 30;
  const int32_t rem = (int32_t)(phase - (q << 30));
  const float x = (float)rem * (float)(6.283185307179586 / 4294967296.0);
  const float z = x * x;
  /* Horner steps as FUSED multiply-adds (one rounding each, fmaf is exact): the statement the kernels' FFMA and the NumPy
   * emulation (oracle/dynamics.py: fma32) agree with bit for bit */
  float ps = (float)(1.0 / 362880.0);
  ps = fmaf(ps, z, (float)(-1.0 / 5040.0));
  ps = fmaf(ps, z, (float)(1.0 / 120.0));
  ps = fmaf(ps, z, (float)(-1.0 / 6.0));
  const float s = fmaf(x, z * ps, x);
  float pc = (float)(-1.0 / 3628800.0);
  pc = fmaf(pc, z, (float)(1.0 / 40320.0));
  pc = fmaf(pc, z, (float)(-1.0 / 720.0));
  pc = fmaf(pc, z, (float)(1.0 / 24.0));
  pc = fmaf(pc, z, -0.5f);
  const float c = fmaf(z, pc, 1.0f);
  switch (q & 3u) {
    case 0: *s_out = s; *c_out = c; break;
    case 1: *s_out = c; *c_out = -s; break;
    case 2: *s_out = -s; *c_out = -c; break;
    default: *s_out = -c; *c_out = s; break;
  }
}

float oracle_tan(float x) {
  const float z = x * x;
  float p = (float)(21844.0 / 6081075.0);
  p = fmaf(p, z, (float)(1382.0 / 155925.0));
  p = fmaf(p, z, (float)(62.0 / 2835.0));
  p = fmaf(p, z, (float)(17.0 / 315.0));
  p = fmaf(p, z, (float)(2.0 / 15.0));
  p = fmaf(p, z, (float)(1.0 / 3.0));
  return fmaf(x, z * p, x);
}

static float sin_small(float x) {
  const float z = x * x;
  float p = (float)(1.0 / 362880.0);
  p = p * z + (float)(-1.0 / 5040.0);
  p = p * z + (float)(1.0 / 120.0);
  p = p * z + (float)(-1.0 / 6.0);
  return x + x * (z * p);
}
static float cos_small(float x) {
  const float z = x * x;
  float p = (float)(-1.0 / 3628800.0);
  p = p * z + (float)(1.0 / 40320.0);
  p = p * z + (float)(-1.0 / 720.0);
  p = p * z + (float)(1.0 / 24.0);
  p = p * z + -0.5f;
  return 1.0f + z * p;
}

float oracle_log(float u) {
  uint32_t bits;
  memcpy(&bits, &u, 4);
  int e = (int)((bits >> 23) & 0xFFu) - 127;
  uint32_t mb = (bits & 0x007FFFFFu) | 0x3F800000u;
  float m;
  memcpy(&m, &mb, 4);
  if (m > (float)1.4142135623730951) { m = m * 0.5f; e += 1; }
  const float s = (m - 1.0f) / (m + 1.0f);
  const float z = s * s;
  float p = (float)(1.0 / 9.0);
  p = p * z + (float)(1.0 / 7.0);
  p = p * z + (float)(1.0 / 5.0);
  p = p * z + (float)(1.0 / 3.0);
  const float lm = (s + s) * (1.0f + z * p);
  return (float)e * (float)0.6931471805599453 + lm;
}

void oracle_normal_pair(uint32_t x0, uint32_t x1, float* n0, float* n1) {
  const float u1 = ((float)(x0 >> 8) + 1.0f) * (float)(1.0 / 16777216.0);
  const float rad = sqrtf(-2.0f * oracle_log(u1));
  float s, c;
  oracle_sincos_turns(x1, &s, &c);
  *n0 = rad * c;
  *n1 = rad * s;
}

static float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

static void kf_sample(const standin_params* p, standin_state* s, float rel_v) {
  if (s->kf_n == 0u) { s->kf_vref = rel_v; s->kf_n = 1u; return; }
  const float dt = (p->accel_mode == 1) ? (float)s->kf_n * p->h : p->h;
  const float raw = (rel_v - s->kf_vref) / dt;
  s->kf_P = s->kf_P + p->kf_q;
  const float K = s->kf_P / (s->kf_P + p->kf_r);
  s->kf_x = s->kf_x + K * (raw - s->kf_x);
  s->kf_P = s->kf_P * (1.0f - K);
  if (p->accel_mode == 1) s->kf_n += 1u; else s->kf_vref = rel_v;
}

static float pid_thrust(const standin_params* p, standin_state* s, float vz_sp) {
  const float e = vz_sp - s->v_z;
  float thrust = 0.0f, e1 = s->e1, e2 = s->e1;
  for (int k = 0; k < p->pid_ticks; ++k) {
    s->integ = clipf(s->integ + e * p->pid_dt, -p->pid_windup, p->pid_windup);
    const float f = p->bw_inv_denom * (((e2 + 2.0f * e1) + e) - p->bw_k2 * s->f3);
    e2 = e1; e1 = e;
    s->f3 = s->f2; s->f2 = s->f1; s->f1 = f;
    thrust = clipf(p->pid_kp * f + p->pid_ki * s->integ, p->pid_lo, p->pid_hi);
  }
  s->e1 = e;
  return thrust;
}

/* one agent period (n_sub sub-steps) toward set-point sp; vz_sp only matters for the second-order model */
void oracle_advance(const standin_params* p, standin_state* s, float sp, float vz_sp) {
  for (int i = 0; i < p->n_sub; ++i) {
    if (p->second_order) {
      const float thrust = pid_thrust(p, s, vz_sp);
      const float alpha = -(p->att_kr * sin_small(s->theta - sp)) - p->att_kw * s->omega;
      s->omega = s->omega + alpha * p->h;
      s->theta = s->theta + s->omega * p->h;
      s->a_d = (p->g * (thrust * p->inv_mg)) * sin_small(s->theta) - p->c_d * s->v_d;
      const float a_z = (thrust * cos_small(s->theta)) * p->inv_m - p->g_abs;
      s->z = (s->z + s->v_z * p->h) + a_z * p->half_h2;
      s->v_z = s->v_z + a_z * p->h;
    } else {
      s->theta = fmaf(sp - s->theta, p->k_theta, s->theta);      /* fused multiply-adds, like the kernels' FFMA */
      s->a_d = fmaf(-p->c_d, s->v_d, p->g * oracle_tan(s->theta));
    }
    s->x_d = fmaf(s->a_d, p->half_h2, fmaf(s->v_d, p->h, s->x_d));
    s->v_d = fmaf(s->a_d, p->h, s->v_d);
    s->phase += p->dphase;
    if (p->accel_mode != 0) {
      float sn, cs;
      oracle_sincos_turns(s->phase, &sn, &cs);
      kf_sample(p, s, p->rw * cs - s->v_d);
    }
  }
}

/* out = rel_p, rel_v, rel_a, pitch */
void oracle_observe(const standin_params* p, const standin_state* s, float out[4]) {
  float sn, cs;
  oracle_sincos_turns(s->phase, &sn, &cs);
  out[0] = fmaf(p->r, sn, -s->x_d);
  out[1] = fmaf(p->rw, cs, -s->v_d);
  out[2] = fmaf(-p->rw2, sn, -s->a_d);
  if (p->accel_mode != 0) out[2] = s->kf_x;
  out[3] = s->theta;
}

/* A whole trajectory for the cross-check: n_steps agent periods with the given set-points; obs_out[n_steps][4], state in/out */
void oracle_rollout(const standin_params* p, standin_state* s, const float* sps, int n_steps, float vz_sp, float* obs_out) {
  for (int t = 0; t < n_steps; ++t) {
    oracle_advance(p, s, sps[t], vz_sp);
    oracle_observe(p, s, obs_out + 4 * t);
  }
}
