/* TEST INFRASTRUCTURE ONLY -- never linked into or called by the product library.
 *
 * Plain-C restatement (IEEE double, one rounding per operation, -ffp-contract=off) of the reference's TrainingMdp numerics:
 *   discrete_state   PKG/mdp.py:257-333 with _latest_valid_curriculum_step_for_state :149-158, _discretiazion_function :160-170
 *   check            PKG/mdp.py:335-439 (sticky result, goal counter)
 *   reward           PKG/mdp.py:441-541 (operation order preserved; shaping potentials survive reset(), quirk Q11)
 *   continuous_action PKG/mdp.py:543-560
 *   reset            PKG/mdp.py:562-569 + :194-200
 * A second, independent statement next to oracle/mdp_oracle.py; tests/test_oracle_c.py replays the fixtures generated from the
 * UNMODIFIED reference (tests/golden/mdp_trace_*.npz) through it: states, result codes and float64 rewards must be identical.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

enum { NON_TERMINAL = 0, NON_TERMINAL_SUCCESS = 1, TERMINAL_SUCCESS = 2, TERMINAL_CONTACT = 3, TERMINAL_FLYZONE_X = 4,
       TERMINAL_FLYZONE_Y = 5, TERMINAL_FLYZONE_Z = 6, TERMINAL_MINIMUM_ALTITUDE = 7, TERMINAL_TIMEOUT = 8 };

static const double LIMITS_P[5] = {1.0, 0.64, 0.4096, 0.262144, 0.16777216};   /* PKG/mdp.py:45-47 */
static const double LIMITS_V[5] = {1.0, 0.8, 0.64, 0.512, 0.4096};              /* PKG/mdp.py:48-50 */
static const double LIMITS_A[5] = {1.0, 1.0, 1.0, 1.0, 1.0};                    /* PKG/mdp.py:51-53 */

typedef struct {
  /* parameters (TrainingMdp.__init__, PKG/mdp.py:214-235) */
  int32_t w;
  double f_ag, t_max, p_max, w_p, w_v, w_theta, w_dur, w_fail, w_succ, v_max, a_max, theta_max, delta_theta, beta, sigma_a,
      minimum_altitude, delta_t;
  double angles[7];
  /* state */
  int32_t step_count, result, curriculum_check, has_prev, has_cur, done;
  int32_t cur[5], prev[5];
  double phi[3], cumulative_reward, theta_sp, info_cumulative;
  double rel_p, rel_v, rel_a, pitch, z, rel_p_y;
  int32_t contact;
} mdp_t;

static double clipd(double x, double lo, double hi) {
  if (x != x) return x;
  return x < lo ? lo : (x > hi ? hi : x);
}

static int level_of(const double* limits, int n, double value) {     /* PKG/mdp.py:149-158 */
  for (int i = 1; i < n; ++i)
    if (value < -limits[i] || value > limits[i]) return i - 1;
  return n - 1;
}

static int bin_of(double value, double goal, double limit) {         /* PKG/mdp.py:160-170; -1 = the reference raises */
  if (-limit <= value && value < -goal) return 0;
  if (-goal <= value && value <= goal) return 1;
  if (value <= limit) return 2;
  return -1;
}

void mdp_reset(mdp_t* m) {
  m->step_count = 0;
  m->result = NON_TERMINAL;
  m->has_cur = m->has_prev = 0;
  m->curriculum_check = 0;
  m->cumulative_reward = 0.0;
  m->theta_sp = 0.0;
  m->rel_p = m->rel_v = m->rel_a = m->pitch = m->z = m->rel_p_y = 0.0;
  m->contact = 0;
  m->done = 0;
  m->info_cumulative = 0.0;
}

void mdp_init(mdp_t* m, int w, double f_ag, double t_max, double p_max) {
  m->w = w; m->f_ag = f_ag; m->t_max = t_max; m->p_max = p_max;
  m->w_p = -100.0; m->w_v = -10.0; m->w_theta = -1.55; m->w_dur = -6.0; m->w_fail = -2.6; m->w_succ = 2.6;
  m->v_max = 3.39411; m->a_max = 1.28;
  m->theta_max = 21.37723 * (3.141592653589793 / 180.0);        /* np.deg2rad(x) = x * (pi / 180) */
  m->delta_theta = 7.12574 * (3.141592653589793 / 180.0);
  m->beta = 1.0 / 3.0; m->sigma_a = 0.416; m->minimum_altitude = 0.2;
  m->delta_t = 1 / f_ag;                                          /* PKG/mdp.py:147 */
  /* np.linspace(-theta_max, theta_max, 7): start + i * step, last sample overwritten (PKG/mdp.py:145) */
  const double step = (m->theta_max - (-m->theta_max)) / 6;
  for (int i = 0; i < 7; ++i) m->angles[i] = i * step + (-m->theta_max);
  m->angles[6] = m->theta_max;
  m->phi[0] = m->phi[1] = m->phi[2] = 0.0;                        /* outside reset(): quirk Q11 */
  mdp_reset(m);
}

double mdp_act(mdp_t* m, int a) {                                   /* PKG/mdp.py:543-560 */
  if (a == 0) m->theta_sp = fmin(m->theta_sp + m->delta_theta, m->theta_max);
  else if (a == 1) m->theta_sp = fmax(m->theta_sp - m->delta_theta, -m->theta_max);
  return m->theta_sp;
}

/* returns the state id (((l*3+p)*3+v)*3+a)*7+theta, or -1 where the reference raises */
int mdp_observe(mdp_t* m, double rel_p, double rel_v, double rel_a, double pitch, double z, int contact) {
  for (int i = 0; i < 5; ++i) m->prev[i] = m->cur[i];
  m->has_prev = m->has_cur;
  m->rel_p = rel_p; m->rel_v = rel_v; m->rel_a = rel_a; m->pitch = pitch; m->z = z; m->contact = contact; m->rel_p_y = 0.0;
  const int w = m->w, n = w + 1;
  const double p = clipd(rel_p / m->p_max, -1, 1), v = clipd(rel_v / m->v_max, -1, 1), a = clipd(rel_a / m->a_max, -1, 1);
  int lvl = level_of(LIMITS_P, n, p);
  const int lv = level_of(LIMITS_V, n, v), la = level_of(LIMITS_A, n, a);
  if (lv < lvl) lvl = lv;
  if (la < lvl) lvl = la;
  const double cp = lvl >= w ? m->beta : LIMITS_P[lvl + 1] / LIMITS_P[lvl];
  const double cv = lvl >= w ? m->beta : LIMITS_V[lvl + 1] / LIMITS_V[lvl];
  const double ca = lvl == w ? m->sigma_a * m->beta : m->sigma_a;
  const int bp = bin_of(p, LIMITS_P[lvl] * cp, LIMITS_P[lvl]);
  const int bv = bin_of(v, LIMITS_V[lvl] * cv, LIMITS_V[lvl]);
  const int ba = bin_of(a, LIMITS_A[lvl] * ca, LIMITS_A[lvl]);
  if (bp < 0 || bv < 0 || ba < 0) return -1;
  const double cl = clipd(pitch, -m->theta_max, m->theta_max);
  int bi = 0;
  double best = fabs(m->angles[0] - cl);
  for (int i = 1; i < 7; ++i) {                                     /* first-min argmin (np.argmin, PKG/mdp.py:323) */
    const double d = fabs(m->angles[i] - cl);
    if (d < best) { best = d; bi = i; }
  }
  m->cur[0] = lvl; m->cur[1] = bp; m->cur[2] = bv; m->cur[3] = ba; m->cur[4] = bi;
  m->has_cur = 1;
  return (((lvl * 3 + bp) * 3 + bv) * 3 + ba) * 7 + bi;
}

int mdp_check(mdp_t* m) {                                           /* PKG/mdp.py:335-439; returns the sticky result code */
  m->step_count += 1;
  if (m->contact) m->result = TERMINAL_CONTACT;
  else if (m->rel_p < -m->p_max || m->rel_p > m->p_max) m->result = TERMINAL_FLYZONE_X;
  else if (m->rel_p_y < -m->p_max || m->rel_p_y > m->p_max) m->result = TERMINAL_FLYZONE_Y;
  else if (m->z < m->minimum_altitude) m->result = TERMINAL_MINIMUM_ALTITUDE;
  else if (m->z > m->p_max) m->result = TERMINAL_FLYZONE_Z;
  else if (m->step_count >= m->t_max * m->f_ag) m->result = TERMINAL_TIMEOUT;
  else if (m->has_prev && m->cur[1] == 1 && m->cur[2] == 1) {
    if (m->prev[0] == m->w && m->cur[0] == m->w) {
      m->curriculum_check += 1;
      m->result = m->curriculum_check >= m->f_ag ? TERMINAL_SUCCESS : NON_TERMINAL_SUCCESS;
    } else {
      m->curriculum_check = 0;
    }
  }
  m->done = m->result >= TERMINAL_SUCCESS;
  if (m->done) m->info_cumulative = m->cumulative_reward;            /* quirk Q12 (PKG/mdp.py:437) */
  return m->result;
}

double mdp_reward(mdp_t* m) {                                        /* PKG/mdp.py:441-541, operation order preserved */
  const double pn = clipd(m->rel_p / m->p_max, -1, 1), vn = clipd(m->rel_v / m->v_max, -1, 1), tn = m->theta_sp / m->theta_max;
  const int lvl = m->cur[0];
  const double prev0 = m->phi[0], prev1 = m->phi[1], prev2 = m->phi[2];
  const double cur0 = m->w_p * fabs(pn), cur1 = m->w_v * fabs(vn), cur2 = m->w_theta * fabs(tn);
  m->phi[0] = cur0; m->phi[1] = cur1; m->phi[2] = cur2;
  const double lv = LIMITS_V[lvl], la = LIMITS_A[lvl];
  const double r_p_max = fabs(m->w_p) * lv * m->delta_t;
  const double r_v_max = fabs(m->w_v) * la * m->delta_t;
  const double r_th_max = fabs(m->w_theta) * (m->delta_theta / m->theta_max) * lv;
  const double r_dur_max = m->w_dur * lv * m->delta_t;
  const double r_max = r_p_max + r_v_max + r_th_max + r_dur_max;
  const double r_p = clipd(cur0 - prev0, -r_p_max, r_p_max);
  const double r_v = clipd(cur1 - prev1, -r_v_max, r_v_max);
  const double r_th = m->w_theta * (fabs(cur2) - fabs(prev2)) / m->theta_max * lv;
  const double r_dur = m->w_dur * lv * m->delta_t;
  const double r_term = (m->result == NON_TERMINAL_SUCCESS || m->result == TERMINAL_SUCCESS) ? m->w_succ * r_max : m->w_fail * r_max;
  const double r = r_p + r_v + r_th + r_dur + r_term;
  m->cumulative_reward += r;
  return r;
}

size_t mdp_sizeof(void) { return sizeof(mdp_t); }
double mdp_cumulative(const mdp_t* m) { return m->cumulative_reward; }
int mdp_done(const mdp_t* m) { return m->done; }
