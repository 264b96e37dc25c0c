/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
 *
 * The batched semantics "S1" (oracle/loop.py: PopulationOracle) in plain C: one agent, n_envs
 * environments.  In a global step every env selects its action and reads its bootstrap value from the tables as they were at
 * the START of the step; the updates are applied one after the other in env order on the live table, each with the learning
 * rate of the live pre-increment count; finished episodes enter the success window in env order, the promotion test
 * (PKG/trainer.py:219-236) runs after every append and takes effect at the end of the step: transfer
 * (PKG/double_q_learning.py:77-89 / the "paper" variant), next working step, fresh MDPs.  With n_envs = 1 this is the reference
 * loop (loop.c).  Fast enough for populations of thousands of envs, where the Python statement takes minutes.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

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

typedef struct {
  int32_t curriculum_steps, window_len, transfer_mode /* 0 reference, 1 paper */;
  double success_rate;
  int64_t max_num_episodes;
  float transfer_ratio[5];                 /* float32(transfer_learning_ratio(k)), PKG/trainer.py:128-138 */
} trainer_params;

typedef struct {
  int32_t w, finished, window_count, window_sum;
  int64_t t, episodes_done, total_steps, total_episodes, total_successes, term_hist[9];
  int32_t n_promotions;
} population_result;

enum { CELLS_PER_LEVEL = 567, CELLS = 2835, TERMINAL_SUCCESS = 2 };

static double exploration_rate(int episode, int w) {
  if (w > 0) return 0.0;
  if (0 <= episode && episode <= 800) return 1.0;
  const double e = 1 + (0.01 - 1) * (episode - 800) / (2000 - 800);
  return e > 0.01 ? e : 0.01;
}
static float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

static int reset_env(const standin_params* sp, const loop_params* lp, standin_state* st, void* mdp, uint64_t seed, uint32_t population,
                     uint32_t env, uint32_t birth, int w) {
  const uint32_t ctr[4] = {env, birth, 1u, population};
  uint32_t d[4];
  oracle_philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), d);
  float x_init;
  if (w == 0) {
    float n0, n1;
    oracle_normal_pair(d[0], d[1], &n0, &n1);
    x_init = lp->sigma_x * n0;
  } else {
    x_init = -lp->p_max_f + lp->two_p_max_f * ((float)(d[0] >> 8) * (float)(1.0 / 16777216.0));
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
  mdp_reset(mdp);
  return mdp_observe(mdp, o[0], o[1], o[2], o[3], z, (z <= lp->z_touch) && (fabsf(o[0]) <= lp->half_platform));
}

/* tables: qa, qb float32 [2835], count float64 [2835] in/out.  Optional traces [n_steps][n_envs] (may be NULL): action (u8),
 * next_state (u16), code (u8), reward (f64).  Returns 0, or -1 when out of memory. */
int oracle_population_run(const standin_params* sp, const loop_params* lp, const trainer_params* tp, int n_envs, uint64_t seed,
                          uint32_t population, int w0, int n_steps, float* qa, float* qb, double* count, population_result* res,
                          uint8_t* tr_action, uint16_t* tr_next_state, uint8_t* tr_code, double* tr_reward) {
  const size_t msz = (mdp_sizeof() + 15) & ~(size_t)15;
  unsigned char* mdps = (unsigned char*)malloc(msz * (size_t)n_envs);
  standin_state* st = (standin_state*)calloc((size_t)n_envs, sizeof(standin_state));
  int* sid = (int*)malloc(sizeof(int) * (size_t)n_envs);
  int* ep = (int*)calloc((size_t)n_envs, sizeof(int));
  int* steps_in_ep = (int*)calloc((size_t)n_envs, sizeof(int));
  uint8_t* finished_env = (uint8_t*)malloc((size_t)n_envs);
  float* snap_a = (float*)malloc(sizeof(float) * CELLS);
  float* snap_b = (float*)malloc(sizeof(float) * CELLS);
  uint8_t window[128];
  if (!mdps || !st || !sid || !ep || !steps_in_ep || !finished_env || !snap_a || !snap_b || tp->window_len > 128) return -1;
  if (sp->second_order) return -2;          /* altitude is derived from the step count here: first-order model (+ estimator) only */
  memset(res, 0, sizeof(*res));
  int w = w0, head = 0, wcount = 0, wsum = 0;
  int64_t t = 0;
  for (int i = 0; i < n_envs; ++i) {
    st[i].kf_P = 1.0f;
    mdp_init(mdps + msz * i, w, lp->f_ag, lp->t_max, lp->p_max);
    sid[i] = reset_env(sp, lp, &st[i], mdps + msz * i, seed, population, (uint32_t)i, (uint32_t)t, w);
  }
  for (int step = 0; step < n_steps && !res->finished; ++step) {
    memcpy(snap_a, qa, sizeof(float) * CELLS);
    memcpy(snap_b, qb, sizeof(float) * CELLS);
    int promote = 0, advance = 0;
    for (int i = 0; i < n_envs; ++i) {
      void* m = mdps + msz * i;
      const uint32_t ctr[4] = {(uint32_t)i, (uint32_t)t, 0u, population};
      uint32_t d[4];
      oracle_philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32), d);
      const int s = sid[i];
      const int explore = (double)(d[0] >> 8) < ceil(exploration_rate(ep[i], w) * 16777216.0);
      int greedy = 0;
      float best = (snap_a[s * 3] + snap_b[s * 3]) / 2.0f;
      for (int a = 1; a < 3; ++a) {
        const float v = (snap_a[s * 3 + a] + snap_b[s * 3 + a]) / 2.0f;
        if (v > best) { best = v; greedy = a; }
      }
      const int a = explore ? (int)(((uint64_t)d[1] * 3u) >> 32) : greedy;
      const double th_sp = mdp_act(m, a);
      oracle_advance(sp, &st[i], (float)th_sp, 0.0f);
      steps_in_ep[i] += 1;
      float o[4];
      oracle_observe(sp, &st[i], o);
      const float z = sp->z_init + (float)steps_in_ep[i] * lp->dz;
      const int s2 = mdp_observe(m, o[0], o[1], o[2], o[3], z, (z <= lp->z_touch) && (fabsf(o[0]) <= lp->half_platform));
      const int code = mdp_check(m);
      const int done = mdp_done(m);
      const double r = mdp_reward(m);
      const int cell = s * 3 + a;
      const double c0 = count[cell];
      double alpha = lp->alpha_min;
      if (c0 != 0) {
        const double pw = pow(1 / c0, lp->omega);
        alpha = pw > lp->alpha_min ? pw : lp->alpha_min;
      }
      count[cell] += 1;
      float qn = snap_a[s2 * 3];
      for (int k = 1; k < 3; ++k) if (snap_a[s2 * 3 + k] > qn) qn = snap_a[s2 * 3 + k];
      const float tgt = (float)r + ((float)lp->gamma * qn) * (float)(((s / 63) % 3) != ((s2 / 63) % 3));
      qa[cell] = qa[cell] + (float)alpha * (tgt - qa[cell]);
      if (tr_action) tr_action[(size_t)step * n_envs + i] = (uint8_t)a;
      if (tr_next_state) tr_next_state[(size_t)step * n_envs + i] = (uint16_t)s2;
      if (tr_code) tr_code[(size_t)step * n_envs + i] = (uint8_t)code;
      if (tr_reward) tr_reward[(size_t)step * n_envs + i] = r;
      res->total_steps += 1;
      finished_env[i] = (uint8_t)done;
      if (done) {
        const int ok = code == TERMINAL_SUCCESS;                     /* PKG/trainer.py:219-221 */
        if (wcount == tp->window_len) wsum -= window[head]; else wcount += 1;      /* deque(maxlen) */
        window[head] = (uint8_t)ok;
        wsum += ok;
        head = (head + 1 == tp->window_len) ? 0 : head + 1;
        res->total_episodes += 1;
        res->total_successes += ok;
        res->term_hist[code] += 1;
        res->episodes_done += 1;
        if ((double)wsum / tp->window_len > tp->success_rate) promote = 1;
        if (res->episodes_done >= tp->max_num_episodes) advance = 1;
        ep[i] += 1;
      } else {
        sid[i] = s2;
      }
    }
    for (int i = 0; i < n_envs; ++i)
      if (finished_env[i]) {
        steps_in_ep[i] = 0;
        sid[i] = reset_env(sp, lp, &st[i], mdps + msz * i, seed, population, (uint32_t)i, (uint32_t)(t + 1), w);
      }
    if (promote || advance) {                                        /* PKG/trainer.py:232-245 */
      if (promote) { head = wcount = wsum = 0; }
      res->n_promotions += 1;
      const int cs = tp->curriculum_steps;
      int dst = -1, src = 0;
      if (tp->transfer_mode == 0) { dst = w; src = (w - 1 + cs) % cs; }
      else if (w + 1 < cs) { dst = w + 1; src = w; }
      if (dst >= 0) {
        const float ratio = tp->transfer_ratio[dst];
        for (int k = 0; k < CELLS_PER_LEVEL; ++k) {
          qa[dst * CELLS_PER_LEVEL + k] = qa[src * CELLS_PER_LEVEL + k] * ratio;
          qb[dst * CELLS_PER_LEVEL + k] = qb[src * CELLS_PER_LEVEL + k] * ratio;
        }
      }
      w += 1;
      res->episodes_done = 0;
      if (w >= cs) {
        res->finished = 1;
        w = cs - 1;
      } else {
        for (int i = 0; i < n_envs; ++i) {
          mdp_init(mdps + msz * i, w, lp->f_ag, lp->t_max, lp->p_max);
          ep[i] = 0;
          steps_in_ep[i] = 0;
          sid[i] = reset_env(sp, lp, &st[i], mdps + msz * i, seed, population, (uint32_t)i, (uint32_t)(t + 1), w);
        }
      }
    }
    t += 1;
  }
  res->w = w; res->t = t; res->window_count = wcount; res->window_sum = wsum;
  free(mdps); free(st); free(sid); free(ep); free(steps_in_ep); free(finished_env); free(snap_a); free(snap_b);
  return 0;
}
