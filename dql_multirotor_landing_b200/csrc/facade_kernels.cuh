// facade_kernels.cuh -- float64 single-object kernels behind the Python facade classes, and the division self-test
// Part of libdqlb200 (see dqlb200.cu for the kernel inventory and the C-ABI).
#pragma once
#include "env_state.cuh"

namespace dql {

// Exhaustive self-test of div_f32_by_const against __ddiv_rn: every FINITE fp32 bit pattern (non-finite
// observations raise the population's error flag instead).  out[0]: mismatches of the production routine,
// out[1]: of the variant with a single correction step (diagnostic).
__global__ void selftest_division_kernel(const __grid_constant__ KC kc, unsigned long long* mismatches) {
  unsigned long long bad = 0, bad1 = 0;
  for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32);
       b += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((uint32_t)b);
    if (!(fabsf(x) <= 3.4028234664e38f)) continue;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
      const double d = v ? kc.v_max : kc.p_max, rcp = v ? kc.rcp_v_max : kc.rcp_p_max;
      const double exact = __ddiv_rn((double)x, d);
      // compare magnitudes bit for bit (the sign of a zero quotient is irrelevant to the callers)
      bad += __double_as_longlong(fabs(div_f32_by_const(x, d, rcp, kc.div_two_steps != 0))) != __double_as_longlong(fabs(exact));
      const double q0 = __dmul_rn((double)x, rcp);
      const double q1 = __fma_rn(__fma_rn(-q0, d, (double)x), rcp, q0);
      bad1 += __double_as_longlong(fabs(q1)) != __double_as_longlong(fabs(exact));
    }
  }
  // float64 numerators (set-points, set-point differences): 2^32 pseudo-random values in [-1, 1] with all
  // 52 mantissa bits random, exponents spread over 2^-40 .. 2^0, plus the multiples of delta_theta
  unsigned long long bad2 = 0;
  for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32);
       b += (unsigned long long)gridDim.x * blockDim.x) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)b, 0u, 7u, 0u), 0x5EEDu, 0u);
    const unsigned long long mant = (((unsigned long long)r.x << 32) | r.y) & 0x000FFFFFFFFFFFFFull;
    const unsigned long long expo = 1023ull - (unsigned long long)(r.z % 41u);
    const unsigned long long sign = (unsigned long long)(r.w & 1u) << 63;
    double x = __longlong_as_double((long long)(sign | (expo << 52) | mant));
    if (b < 16) x = (double)((long long)b - 8) * kc.delta_theta;
    const double q = div_f64_by_const(x, kc.theta_max, kc.rcp_theta_max);
    bad2 += __double_as_longlong(fabs(q)) != __double_as_longlong(fabs(__ddiv_rn(x, kc.theta_max)));
  }
  if (bad) atomicAdd(mismatches, bad);
  if (bad1) atomicAdd(mismatches + 1, bad1);
  if (bad2) atomicAdd(mismatches + 2, bad2);
}


// Device self-test of the fp32 cut-table discretisation (R5) on caller-supplied observations: obs [n][4] = rel_p, rel_v, rel_a,
// pitch -> state id.  The three ways the production kernels call discretise_cuts:
//   KDEF = false, WITH_W = true   train_kernel generic instances / env_step_kernel: run-time angle cuts, level loop bounded by w
//   KDEF = true,  WITH_W = true   train_kernel production instances: compile-time angle cuts (KDef), cut table staged in shared memory
//   *,            WITH_W = false  eval kernels, env_reset: every level index probed (cuts above the working step are NaN)
// One population-state error word: any observation with a NaN component sets bit 0 (the reference raises, PKG/mdp.py:170).
template <bool KDEF, bool WITH_W>
__global__ void __launch_bounds__(128) selftest_discretise_kernel(const __grid_constant__ KC kc, int w, long long n, const float* __restrict__ obs,
                                                                  uint16_t* __restrict__ out_state) {
  __shared__ dqlb200_cuts cuts;
  if (threadIdx.x == 0) cuts = kc.cuts[w];
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Obs o;
  o.rel_p = obs[4 * i + 0]; o.rel_v = obs[4 * i + 1]; o.rel_a = obs[4 * i + 2]; o.pitch = obs[4 * i + 3];
  o.z = 0.0f; o.contact = false;
  DState d;
  if (KDEF) {
    const KDef kk{};
    d = WITH_W ? discretise_cuts(cuts, kk.angle_cut, o, w) : discretise_cuts(cuts, kk.angle_cut, o);
  } else {
    d = WITH_W ? discretise_cuts(kc.cuts[w], kc.angle_cut, o, w) : discretise_cuts(kc.cuts[w], kc.angle_cut, o);
  }
  out_state[i] = (uint16_t)d.id();
}

// OR of the populations' error flags + the smallest offending population (dqlb200_check_errors): out[0] = flags, out[1] = index
__global__ void error_reduce_kernel(const dqlb200_population_state* __restrict__ ps, int n_pop, uint32_t* out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t f = p < n_pop ? ps[p].error_flags : 0u;
  const uint32_t any = __ballot_sync(FULL, f != 0u);
  if (f != 0u) {
    atomicOr(out, f);
    if ((threadIdx.x & 31) == __ffs(any) - 1) atomicMin(out + 1, (uint32_t)p);
  }
}

// -------------------------------------------------------------------------------------------------
// Facade kernel: float64 observations, the reference's comparisons in float64 (PKG/mdp.py:149-170,
// 257-333, 335-439, 441-541, 784-845).  One thread per MDP object.
// -------------------------------------------------------------------------------------------------
__device__ int level_f64(const double* lim, int w, double v) {
  for (int idx = 1; idx <= w; ++idx)
    if (v < -lim[idx] || v > lim[idx]) return idx - 1;
  return w;
}
__device__ int bin_f64(double v, double goal, double limit) {
  if (-limit <= v && v < -goal) return 0;
  if (-goal <= v && v <= goal) return 1;
  if (v <= limit) return 2;
  return -1;   // NaN: the reference raises ValueError (PKG/mdp.py:170)
}
__device__ int discretise_f64(const dqlb200_config* cfg, int w, double rel_p, double rel_v, double rel_a, double pitch) {
  if (rel_p != rel_p || rel_v != rel_v || rel_a != rel_a) return -1;   // fmin/fmax would swallow the NaN np.clip keeps
  const double p = clipd(__ddiv_rn(rel_p, cfg->p_max), -1.0, 1.0);
  const double v = clipd(__ddiv_rn(rel_v, cfg->v_max), -1.0, 1.0);
  const double a = clipd(__ddiv_rn(rel_a, cfg->a_max), -1.0, 1.0);
  const int lvl = min(min(level_f64(cfg->limits[0], w, p), level_f64(cfg->limits[1], w, v)), level_f64(cfg->limits[2], w, a));
  const int bp = bin_f64(p, cfg->goal_width[w][0][lvl], cfg->limits[0][lvl]);
  const int bv = bin_f64(v, cfg->goal_width[w][1][lvl], cfg->limits[1][lvl]);
  const int ba = bin_f64(a, cfg->goal_width[w][2][lvl], cfg->limits[2][lvl]);
  if (bp < 0 || bv < 0 || ba < 0 || pitch != pitch) return -1;
  const double cl = clipd(pitch, -cfg->theta_max, cfg->theta_max);
  int bi = 0;
  double best = fabs(__dsub_rn(cfg->angles[0], cl));
  for (int i = 1; i < 7; ++i) {
    const double d = fabs(__dsub_rn(cfg->angles[i], cl));
    if (d < best) { best = d; bi = i; }
  }
  return (((lvl * 3 + bp) * 3 + bv) * 3 + ba) * 7 + bi;
}

__global__ void facade_kernel(const dqlb200_config* __restrict__ cfg, int w, int ops, long long n,
                              const double* __restrict__ obs, const uint8_t* __restrict__ contact,
                              const int8_t* __restrict__ action, double* __restrict__ st,
                              uint16_t* out_state, uint8_t* out_code, double* out_reward, uint32_t* error_flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double* s = st + i * 12;
  const bool sim = (ops & DQLB200_OP_SIMULATION) != 0;
  if (ops & DQLB200_OP_RESET) {        // phi (s[1..3]) survives: quirk Q11
    s[0] = 0.0; s[4] = 0.0; s[5] = 0.0; s[6] = 0.0; s[7] = 0.0; s[8] = -1.0; s[9] = -1.0;
  }
  if (ops & DQLB200_OP_ACTION) {
    const int a = action[i];
    if (a == 0) s[0] = fmin(__dadd_rn(s[0], cfg->delta_theta), cfg->theta_max);
    else if (a == 1) s[0] = fmax(__dsub_rn(s[0], cfg->delta_theta), -cfg->theta_max);
  }
  if (ops & DQLB200_OP_OBSERVE) {
    const double* o = obs + i * 6;
    const int sid = discretise_f64(cfg, w, o[0], o[1], o[2], o[3]);
    if (sid < 0) { atomicOr(error_flag, 1u); return; }
    s[9] = s[8];
    s[8] = (double)sid;
    s[10] = o[0];
    s[11] = o[1];
    if (out_state) out_state[i] = (uint16_t)sid;
  }
  if (ops & DQLB200_OP_CHECK) {
    const double* o = obs + i * 6;
    if (s[8] < 0.0) { atomicOr(error_flag, 2u); return; }
    const int cur = (int)s[8];
    const int lvl = cur / DQLB200_STATES_PER_LEVEL, bp = (cur / 63) % 3, bv = (cur / 21) % 3;
    int code = (int)s[7];
    s[5] += 1.0;
    if (contact[i]) code = DQLB200_TERMINAL_CONTACT;
    else if (o[0] < -cfg->p_max || o[0] > cfg->p_max) code = DQLB200_TERMINAL_FLYZONE_X;
    else if (o[5] < -cfg->p_max || o[5] > cfg->p_max) code = DQLB200_TERMINAL_FLYZONE_Y;
    else if (o[4] < cfg->minimum_altitude) code = DQLB200_TERMINAL_MINIMUM_ALTITUDE;
    else if (o[4] > cfg->p_max) code = DQLB200_TERMINAL_FLYZONE_Z;
    else if (s[5] >= cfg->timeout_threshold) code = DQLB200_TERMINAL_TIMEOUT;
    else if (!sim && s[9] >= 0.0 && bp == 1 && bv == 1) {
      const int prev_lvl = (int)s[9] / DQLB200_STATES_PER_LEVEL;
      if (prev_lvl == w && lvl == w) {
        s[6] += 1.0;
        code = (s[6] >= cfg->f_ag) ? DQLB200_TERMINAL_SUCCESS : DQLB200_NON_TERMINAL_SUCCESS;
      } else {
        s[6] = 0.0;
      }
    }
    s[7] = (double)code;
    if (out_code) out_code[i] = (uint8_t)code;
  }
  if (ops & DQLB200_OP_REWARD) {
    if (s[8] < 0.0 || s[9] < 0.0) { atomicOr(error_flag, 4u); return; }
    const int lvl = (int)s[8] / DQLB200_STATES_PER_LEVEL;
    const dqlb200_reward_level rl = cfg->reward[lvl];
    const double phi_p = __dmul_rn(cfg->w_p, fabs(clipd(__ddiv_rn(s[10], cfg->p_max), -1.0, 1.0)));
    const double phi_v = __dmul_rn(cfg->w_v, fabs(clipd(__ddiv_rn(s[11], cfg->v_max), -1.0, 1.0)));
    const double phi_t = __dmul_rn(cfg->w_theta, fabs(__ddiv_rn(s[0], cfg->theta_max)));
    const int code = (int)s[7];
    const double r_p = clipd(__dsub_rn(phi_p, s[1]), -rl.r_p_max, rl.r_p_max);
    const double r_v = clipd(__dsub_rn(phi_v, s[2]), -rl.r_v_max, rl.r_v_max);
    const double r_t = __dmul_rn(__ddiv_rn(__dmul_rn(cfg->w_theta, __dsub_rn(fabs(phi_t), fabs(s[3]))), cfg->theta_max), rl.lim_v);
    const double r_term = (code == DQLB200_NON_TERMINAL_SUCCESS || code == DQLB200_TERMINAL_SUCCESS) ? rl.r_term_succ : rl.r_term_fail;
    const double r = __dadd_rn(__dadd_rn(__dadd_rn(__dadd_rn(r_p, r_v), r_t), rl.r_dur), r_term);
    s[1] = phi_p; s[2] = phi_v; s[3] = phi_t;
    s[4] = __dadd_rn(s[4], r);
    if (out_reward) out_reward[i] = r;
  }
}

// Single-object DoubleQLearningAgent calls in float64 (the reference's table dtype).  One thread: the
// facade is an API mirror, not a throughput path.
__global__ void agent_facade_kernel(int op, long long n, double* t, int cs, const int32_t* state, const int32_t* action,
                                    const int32_t* next_state, const double* alpha, const double* reward, double gamma,
                                    int32_t* out_action) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double* qa = t;
  double* qb = t + CELLS;
  double* cnt = t + 2 * CELLS;
  if (op == DQLB200_AGENT_PREDICT) {
    for (long long i = 0; i < n; ++i) {
      const int s = state[i] * 3;
      int a = 0;
      double best = __ddiv_rn(__dadd_rn(qa[s], qb[s]), 2.0);
      for (int k = 1; k < 3; ++k) {
        const double v = __ddiv_rn(__dadd_rn(qa[s + k], qb[s + k]), 2.0);
        if (v > best) { best = v; a = k; }
      }
      out_action[i] = a;
    }
  } else if (op == DQLB200_AGENT_UPDATE) {
    for (long long i = 0; i < n; ++i) {
      const int sa = state[i] * 3 + action[i];
      const int s2 = next_state[i] * 3;
      cnt[sa] = __dadd_rn(cnt[sa], 1.0);
      int b = 0;
      for (int k = 1; k < 3; ++k)
        if (qa[s2 + k] > qa[s2 + b]) b = k;
      const double changed = (((state[i] / 63) % 3) != ((next_state[i] / 63) % 3)) ? 1.0 : 0.0;
      const double tgt = __dadd_rn(reward[i], __dmul_rn(__dmul_rn(gamma, qa[s2 + b]), changed));
      qa[sa] = __dadd_rn(qa[sa], __dmul_rn(alpha[i], __dsub_rn(tgt, qa[sa])));
    }
  } else if (op == DQLB200_AGENT_TRANSFER) {
    const int step = state[0];
    // the source slot (step - 1) mod the AGENT's curriculum_steps comes from the caller (next_state[0]); cs is the handle's
    const int src = next_state ? next_state[0] : (step - 1 + cs) % cs;
    const double ratio = alpha[0];
    for (int i = 0; i < DQLB200_CELLS_PER_LEVEL; ++i) {
      qa[step * DQLB200_CELLS_PER_LEVEL + i] = __dmul_rn(qa[src * DQLB200_CELLS_PER_LEVEL + i], ratio);
      qb[step * DQLB200_CELLS_PER_LEVEL + i] = __dmul_rn(qb[src * DQLB200_CELLS_PER_LEVEL + i], ratio);
    }
  }
}


}  // namespace dql
