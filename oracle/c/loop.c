/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
 *
 * The reference trainer loop for ONE environment (PKG/trainer.py:191-212 with the ordering of
 * PKG/landing_simulation_env.py:245-282) in plain C: guess (PKG/double_q_learning.py:110-124), exploration_rate and alpha
 * (PKG/trainer.py:88-126), update (PKG/double_q_learning.py:91-108,126-146) on float32 tables -- the arithmetic the unmodified
 * reference performs under NumPy >= 2 when float32 tables are assigned to it (NEP 50) -- driven by the stand-in
 * (standin.c), the MDP (mdp.c) and the Philox draw contract.  It is oracle/loop.py: PopulationOracle(n_envs = 1) restated;
 * tests/test_oracle_c.py replays the replay_*_float32 fixtures of the unmodified reference through it.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* standin.c */
typedef struct {
  float h, half_h2, k_theta, g, c_d, r, rw, rw2;
  uint32_t dphase;
  int32_t n_sub, accel_mode;
  float kf_q, kf_r;
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
void oracle_philox4x32_10(const uint32_t ctr[4], uint32_t k0, uint32_t k1, uint32_t out[4]);
void oracle_sincos_turns(uint32_t phase, float* s_out, float* c_out);
void oracle_normal_pair(uint32_t x0, uint32_t x1, float* n0, float* n1);
void oracle_advance(const standin_params* p, standin_state* s, float sp, float vz_sp);
void oracle_observe(const standin_params* p, const standin_state* s, float out[4]);
/* mdp.c (opaque here) */
size_t mdp_sizeof(void);
void mdp_init(void* m, int w, double f_ag, double t_max, double p_max);
void mdp_reset(void* m);
double mdp_act(void* m, int a);
int mdp_observe(void* m, double rel_p, double rel_v, double rel_a, double pitch, double z, int contact);
int mdp_check(void* m);
double mdp_reward(void* m);
int mdp_done(const void* m);

typedef struct {
  float dz, z_touch, half_platform, p_max_f, two_p_max_f, sigma_x;
  double f_ag, t_max, p_max, alpha_min, omega, gamma;
} loop_params;

static double exploration_rate(int episode, int w) {                 /* PKG/trainer.py:112-126 */
  if (w > 0) return 0.0;
  if (0 <= episode && episode <= 800) return 1.0;
  const double e = 1 + (0.01 - 1) * (episode - 800) / (2000 - 800);
  return e > 0.01 ? e : 0.01;
}

static float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

static int reset_env(const standin_params* sp, const loop_params* lp, standin_state* st, void* mdp, uint64_t seed, uint32_t population,
                     uint32_t env, uint32_t birth, int w, float obs5[5]) {
  /* R1 (PKG/landing_simulation_env.py:181-216) + one hover period (:222-224) + first discrete_state */
  const uint32_t ctr[4] = {env, birth, 1u /* PURPOSE_RESET */, population};
  uint32_t d[4];
  oracle_philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), d);
  float x_init;
  if (w == 0) {
    float n0, n1;
    oracle_normal_pair(d[0], d[1], &n0, &n1);
    x_init = lp->sigma_x * n0;
  } else {
    const float u = (float)(d[0] >> 8) * (float)(1.0 / 16777216.0);
    x_init = -lp->p_max_f + lp->two_p_max_f * u;
  }
  st->phase = d[2];
  float s, c;
  oracle_sincos_turns(st->phase, &s, &c);
  st->x_d = sp->r * s + clipf(x_init, -lp->p_max_f, lp->p_max_f);
  st->v_d = 0.0f;
  st->theta = 0.0f;
  oracle_advance(sp, st, 0.0f, 0.0f);
  float o[4];
  oracle_observe(sp, st, o);
  const float z = sp->z_init + 0.0f * lp->dz;
  const int contact = (z <= lp->z_touch) && (fabsf(o[0]) <= lp->half_platform);
  obs5[0] = o[0]; obs5[1] = o[1]; obs5[2] = o[2]; obs5[3] = o[3]; obs5[4] = z;
  mdp_reset(mdp);
  return mdp_observe(mdp, o[0], o[1], o[2], o[3], z, contact);
}

/* n_steps agent steps of one env (env index 0 of `population`), float32 tables qa/qb [2835], float64 counts [2835] in/out.
 * Outputs per step: obs[5] (f32), action, state, next_state, code, done (u8/u16 as int32), reward (f64), episode.
 * mdp_buf: mdp_sizeof() bytes of scratch.  Returns the number of finished episodes. */
int oracle_single_env_loop(const standin_params* sp, const loop_params* lp, void* mdp_buf, uint64_t seed, uint32_t population, int w, int ep0,
                           int n_steps, float* qa, const float* qb, double* count, float* out_obs, int32_t* out_action, int32_t* out_state,
                           int32_t* out_next_state, int32_t* out_code, int32_t* out_done, double* out_reward, int32_t* out_episode) {
  standin_state st;
  memset(&st, 0, sizeof(st));
  st.kf_P = 1.0f;
  mdp_init(mdp_buf, w, lp->f_ag, lp->t_max, lp->p_max);
  float o5[5];
  int sid = reset_env(sp, lp, &st, mdp_buf, seed, population, 0u, 0u, w, o5);
  int ep = ep0, episodes = 0, step_count = 0;
  for (uint32_t t = 0; t < (uint32_t)n_steps; ++t) {
    const uint32_t ctr[4] = {0u, t, 0u /* PURPOSE_STEP */, population};
    uint32_t d[4];
    oracle_philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), d);
    /* guess: both draws always consumed (quirk Q4); greedy = first max of (Q_a + Q_b) / 2 in float32 */
    const double eps = exploration_rate(ep, w);
    const int explore = (double)(d[0] >> 8) < ceil(eps * 16777216.0);
    int greedy = 0;
    float best = (qa[sid * 3] + qb[sid * 3]) / 2.0f;
    for (int a = 1; a < 3; ++a) {
      const float v = (qa[sid * 3 + a] + qb[sid * 3 + a]) / 2.0f;
      if (v > best) { best = v; greedy = a; }
    }
    const int a = explore ? (int)(((uint64_t)d[1] * 3u) >> 32) : greedy;
    const double th_sp = mdp_act(mdp_buf, a);
    oracle_advance(sp, &st, (float)th_sp, 0.0f);
    step_count += 1;
    float o[4];
    oracle_observe(sp, &st, o);
    const float z = sp->z_init + (float)step_count * lp->dz;
    const int contact = (z <= lp->z_touch) && (fabsf(o[0]) <= lp->half_platform);
    const int sid2 = mdp_observe(mdp_buf, o[0], o[1], o[2], o[3], z, contact);
    const int code = mdp_check(mdp_buf);
    const int done = mdp_done(mdp_buf);
    const double r = mdp_reward(mdp_buf);
    /* alpha from the count BEFORE the increment (PKG/trainer.py:88-110, argument order at :203-209) */
    const int cell = sid * 3 + a;
    const double c0 = count[cell];
    double alpha = lp->alpha_min;
    if (c0 != 0) {
      const double pw = pow(1 / c0, lp->omega);
      alpha = pw > lp->alpha_min ? pw : lp->alpha_min;
    }
    count[cell] += 1;
    /* update of table A either way (quirks Q1-Q3), float32: T = r + (gamma * max_a Q_a[s']) * [p-bin changed] */
    float qn = qa[sid2 * 3];
    for (int k = 1; k < 3; ++k) if (qa[sid2 * 3 + k] > qn) qn = qa[sid2 * 3 + k];
    const int bp = (sid / 63) % 3, bp2 = (sid2 / 63) % 3;
    const float tgt = (float)r + ((float)lp->gamma * qn) * (float)(bp != bp2);
    const float loss = (float)alpha * (tgt - qa[cell]);
    qa[cell] = qa[cell] + loss;
    out_obs[5 * t + 0] = o[0]; out_obs[5 * t + 1] = o[1]; out_obs[5 * t + 2] = o[2]; out_obs[5 * t + 3] = o[3]; out_obs[5 * t + 4] = z;
    out_action[t] = a; out_state[t] = sid; out_next_state[t] = sid2; out_code[t] = code; out_done[t] = done; out_reward[t] = r;
    out_episode[t] = ep;
    if (done) {
      ep += 1;
      episodes += 1;
      step_count = 0;
      sid = reset_env(sp, lp, &st, mdp_buf, seed, population, 0u, t + 1u, w, o5);
    } else {
      sid = sid2;
    }
  }
  return episodes;
}
