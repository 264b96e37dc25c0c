// libdqlb200: kernels + C-ABI (include/dqlb200.h).  Compile for sm_100a only:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -shared
//
// Kernel inventory
//   train_kernel<WARPS>   one CTA per population (agent); K fused global steps per launch.  Per step and
//                         env: epsilon-greedy select (R9/R10), set-point (R3), stand-in dynamics (R4),
//                         discretise (R5), check (R6), reward (R7), learning rate (R11), table update
//                         (R12), auto-reset (R1/R8), success window / promotion / transfer (R13/R14).
//                         Q_a/Q_b/count live in shared memory for the whole launch.
//   reset_kernel          R1 + R8 for every env of every population.
//   eval_kernel           R15: greedy SimulationMdp episodes, one thread per episode.
//   facade_kernel         float64 single-object TrainingMdp/SimulationMdp calls for the Python facade.
//   transfer/shared_*     R13 on bound tables; shared-table mode pack/apply.
//
// Same-cell update semantics ("S1"): all envs of a population select and bootstrap from the
// tables as of the START of the global step; the updates are then applied one by one in env-index order
// to the live table, each with the learning rate of the live pre-increment count.  Implementation: env
// index = slot * blockDim + thread; the commit of (slot, warp) chunks is serialised by a baton passed
// between warps with named barriers; inside a chunk, lanes hitting the same cell are found with
// __match_any_sync and applied sequentially in lane order by shuffles.  Bit-exact vs the sequential
// oracle for any number of envs.
#include <cuda_runtime.h>

#include <cstdio>
#include <cmath>
#include <cstring>
#include <new>
#include <string>

#include "dqlb200_device.cuh"

namespace dql {

constexpr int CELLS = DQLB200_MAX_CELLS;
constexpr uint32_t FULL = 0xFFFFFFFFu;

// env-state word C.x layout
constexpr uint32_t SID_BITS = 10, STEP_SHIFT = 10, STEP_BITS = 9, CC_SHIFT = 19, CC_BITS = 5;
constexpr uint32_t STICKY_BIT = 1u << 24, FRESH_BIT = 1u << 25, BP_SHIFT = 26;   // bits 26-27: position bin of `sid`

struct Env {
  Body b;
  double theta_sp;     // NOT cleared by an episode reset while `fresh` (keeps the shaping potential, quirk Q11)
  float prev_rel_p, prev_rel_v;
  uint32_t sid, bp, step_count, curriculum_check;     // bp = position bin of sid ((sid / 63) % 3, kept to avoid the division)
  bool sticky_success, fresh;
  uint32_t episode;
  double cum_reward;
};

struct EnvPtrs {
  float4* a;
  uint4* b;
  uint4* c;
};

struct EnvRaw {
  float4 A;
  uint4 B, C;
};
__device__ __forceinline__ EnvRaw env_fetch(const EnvPtrs& p, size_t i) {
  EnvRaw r;
  r.A = p.a[i];
  r.B = p.b[i];
  r.C = p.c[i];
  return r;
}
// Asynchronous prefetch of one env's 48 bytes into the thread's private staging slots in shared memory (cp.async, L2 only):
// unlike a register prefetch it holds no registers while in flight and cannot be consumed early by the scheduler's copies.
__device__ __forceinline__ void env_prefetch_async(const EnvPtrs& p, size_t i, uint4* stage, int nt, int tid) {
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(stage + tid);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0), "l"(p.a + i) : "memory");
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + 16u * nt), "l"(p.b + i) : "memory");
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + 32u * nt), "l"(p.c + i) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ EnvRaw env_prefetch_take(const uint4* stage, int nt, int tid) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  EnvRaw r;
  const uint4 a = stage[tid];
  r.A = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
  r.B = stage[nt + tid];
  r.C = stage[2 * nt + tid];
  return r;
}

__device__ __forceinline__ void env_unpack(const EnvRaw& r, Env& e) {
  const float4 A = r.A;
  const uint4 B = r.B;
  const uint4 Cw = r.C;
  e.b.x_d = A.x; e.b.v_d = A.y; e.b.theta = A.z; e.b.phase = __float_as_uint(A.w); e.b.a_d = 0.0f;
  e.theta_sp = __hiloint2double((int)B.y, (int)B.x);
  e.prev_rel_p = __uint_as_float(B.z);
  e.prev_rel_v = __uint_as_float(B.w);
  e.sid = Cw.x & ((1u << SID_BITS) - 1u);
  e.step_count = (Cw.x >> STEP_SHIFT) & ((1u << STEP_BITS) - 1u);
  e.curriculum_check = (Cw.x >> CC_SHIFT) & ((1u << CC_BITS) - 1u);
  e.sticky_success = (Cw.x & STICKY_BIT) != 0u;
  e.bp = (Cw.x >> BP_SHIFT) & 3u;
  e.fresh = (Cw.x & FRESH_BIT) != 0u;
  e.episode = Cw.y;
  e.cum_reward = __hiloint2double((int)Cw.w, (int)Cw.z);
}
__device__ __forceinline__ void env_load(const EnvPtrs& p, size_t i, Env& e) { env_unpack(env_fetch(p, i), e); }

__device__ __forceinline__ void env_store(const EnvPtrs& p, size_t i, const Env& e) {
  p.a[i] = make_float4(e.b.x_d, e.b.v_d, e.b.theta, __uint_as_float(e.b.phase));
  p.b[i] = make_uint4((uint32_t)__double2loint(e.theta_sp), (uint32_t)__double2hiint(e.theta_sp),
                      __float_as_uint(e.prev_rel_p), __float_as_uint(e.prev_rel_v));
  const uint32_t packed = e.sid | (e.step_count << STEP_SHIFT) | (e.curriculum_check << CC_SHIFT) |
                          (e.sticky_success ? STICKY_BIT : 0u) | (e.fresh ? FRESH_BIT : 0u) | (e.bp << BP_SHIFT);
  p.c[i] = make_uint4(packed, e.episode, (uint32_t)__double2loint(e.cum_reward),
                      (uint32_t)__double2hiint(e.cum_reward));
}

// R1 + R8: new episode.  `fresh_mdp` additionally clears what only a NEW TrainingMdp clears
// (shaping potentials, PKG/trainer.py:176 + quirk Q11) and the per-step episode index.
__device__ __forceinline__ void env_reset(const KC& kc, const dqlb200_population_params& pp,
                                          const dqlb200_cuts& cuts, const float* angle_cut, Env& e,
                                          uint32_t env_index, uint32_t birth, int w, bool fresh_mdp) {
  const uint4 d = philox4x32_10(make_uint4(env_index, birth, PURPOSE_RESET, pp.population_id), pp.seed_lo, pp.seed_hi);
  const Obs o = dyn_reset(kc, pp, e.b, d, /*normal_init=*/w == 0, /*simulation=*/false, kc.dz_train);
  const DState ds0 = discretise_cuts(cuts, angle_cut, o);
  e.sid = (uint32_t)ds0.id();
  e.bp = (uint32_t)ds0.bp;
  e.step_count = 0;
  e.curriculum_check = 0;
  e.sticky_success = false;
  e.fresh = true;
  e.cum_reward = 0.0;
  if (fresh_mdp) {
    e.theta_sp = 0.0;
    e.prev_rel_p = 0.0f;
    e.prev_rel_v = 0.0f;
    e.episode = 0;
  }
}

// Baton between the warps of a CTA: warp w waits on named barrier 1+w (its 32 threads + the 32 arriving
// threads of the previous warp).  The ids are IMMEDIATES so that ptxas allocates WARPS+1 barriers per CTA;
// with a register id it reserves all 16 and the 64-barriers-per-SM limit caps occupancy at 4 CTAs
// (ncu launch__occupancy_limit_barriers).
#define DQL_BAR_CASE(OP, ID) case (ID - 1): if (WARPS >= ID) asm volatile("barrier." OP " " #ID ", 64;" ::: "memory"); break;
template <int WARPS>
__device__ __forceinline__ void baton_wait(int warp) {
  switch (warp) {
    DQL_BAR_CASE("sync", 1) DQL_BAR_CASE("sync", 2) DQL_BAR_CASE("sync", 3) DQL_BAR_CASE("sync", 4)
    DQL_BAR_CASE("sync", 5) DQL_BAR_CASE("sync", 6) DQL_BAR_CASE("sync", 7) DQL_BAR_CASE("sync", 8)
    default: break;
  }
}
template <int WARPS>
__device__ __forceinline__ void baton_pass(int next_warp) {
  switch (next_warp) {
    DQL_BAR_CASE("arrive", 1) DQL_BAR_CASE("arrive", 2) DQL_BAR_CASE("arrive", 3) DQL_BAR_CASE("arrive", 4)
    DQL_BAR_CASE("arrive", 5) DQL_BAR_CASE("arrive", 6) DQL_BAR_CASE("arrive", 7) DQL_BAR_CASE("arrive", 8)
    default: break;
  }
}
#undef DQL_BAR_CASE

struct TrainArgs {
  EnvPtrs env;
  uint32_t* tables;                        // [P][3][CELLS]
  dqlb200_population_state* pop_state;     // [P]
  const dqlb200_population_params* pop_params;
  const float* alpha_luts;                 // [n_luts][ALPHA_LUT]
  const uint32_t* eps_threshold;           // [EPS_LUT]
  dqlb200_trace trace;
  uint32_t* merge_snapshot;                // replica-merge mode: [n_groups][3][CELLS] merged tables (may be null)
  int k_steps;
  int pop_offset;                          // first population of this launch (chunked host-buffer calls)
  long long n_total;
};

// Shared memory of one population (CTA).  Q_b and the alpha LUT stay in global memory (read-only in the step
// loop, L1-resident): that keeps the footprint at ~38 KB so that five CTAs fit on one SM.
constexpr int RESET_QUEUE = 128;   // finished envs a warp collects before it runs the batched reset pass

constexpr int STATES = DQLB200_MAX_CURRICULUM * DQLB200_STATES_PER_LEVEL;   // 945

struct Shared {
  float qa[CELLS];        // live table A
  uint32_t cnt[CELLS];    // state_action_counter
  // Snapshot of the start of the global step.  Phase A reads the tables only through two per-STATE quantities, so the
  // snapshot is those two instead of a copy of Q_a: the greedy action argmax_a (Q_a+Q_b)/2 (R9) and max_a Q_a (R12).
  float qmax[STATES];
  uint8_t greedy[STATES + 3];
  dqlb200_cuts cuts;
  dqlb200_reward_level reward[DQLB200_MAX_CURRICULUM];
  dqlb200_population_state ps;
  unsigned long long n_episodes, n_success, ep_steps, hist[9];     // totals of this launch
  uint32_t step_episodes, step_success, step_ep_steps, step_hist[9];   // counters of the current global step (native 32-bit shared atomics)
  int promote, advance, do_advance;
  // followed by: uint16_t reset_queue[WARPS][RESET_QUEUE]   (dynamic)
};

#ifndef DQL_WARPS_PER_SM
#define DQL_WARPS_PER_SM 24     // resident warps per SM the register allocation is tuned for (launch bounds)
#endif
// DIV2: second Markstein correction step of x / p_max, x / v_max (needed unless the divisors are the exhaustively
// verified defaults; the trace instances always take it: both variants are correctly rounded, hence identical)
template <int WARPS, bool TRACE, bool DIV2>
__global__ void __launch_bounds__(WARPS * 32, (DQL_WARPS_PER_SM / WARPS) > 0 ? (DQL_WARPS_PER_SM / WARPS) : 1) train_kernel(const __grid_constant__ KC kc, const TrainArgs args) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NT = WARPS * 32;
  const int pop = blockIdx.x + args.pop_offset;
  const int n_p = kc.envs_per_population;
  const int n_slots = (n_p + NT - 1) / NT;
  uint16_t* reset_queue = reinterpret_cast<uint16_t*>(smem_raw + ((sizeof(Shared) + 15) & ~size_t(15))) + (size_t)warp * RESET_QUEUE;
  uint4* stage = reinterpret_cast<uint4*>(smem_raw + ((sizeof(Shared) + 15) & ~size_t(15)) + (size_t)WARPS * RESET_QUEUE * sizeof(uint16_t));   // [3][NT]
  const size_t env_base = (size_t)pop * n_p;
  uint32_t* gt = args.tables + (size_t)pop * 3 * CELLS;
  float* gqb = reinterpret_cast<float*>(gt + CELLS);      // table B, written only by the transfer below

  // ---- stage population state and the LIVE rows of the tables in shared memory --------------------
  // At working step w only levels 0..w can be visited (a state's level never exceeds w), so only rows
  // [0, (w+1)*567) of Q_a / count are staged, snapshotted and written back; a promotion loads the next level.
  // The launch prologue is ONE round trip to memory: every load below is independent of the others (the working step
  // and the population constants are broadcast loads by every thread instead of a hop through shared memory), and the
  // env state of slot 0 -- a cold HBM read when one global step is run per launch -- is in flight during all of it.
  static_assert(sizeof(dqlb200_population_state) % 4 == 0, "word copies");
  constexpr int PS_WORDS = sizeof(dqlb200_population_state) / 4;
  const dqlb200_population_params pp = args.pop_params[pop];
  env_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), stage, NT, tid);
  {
    const int w_start = args.pop_state[pop].working_step;
    const uint32_t* gps = reinterpret_cast<const uint32_t*>(args.pop_state + pop);
    for (int i = tid; i < PS_WORDS; i += NT) reinterpret_cast<uint32_t*>(&sh.ps)[i] = gps[i];
    if (tid == 0) {
      sh.n_episodes = sh.n_success = sh.ep_steps = 0ull;
      sh.step_episodes = sh.step_success = sh.step_ep_steps = 0u;
      for (int i = 0; i < 9; ++i) { sh.hist[i] = 0ull; sh.step_hist[i] = 0u; }
      sh.promote = sh.advance = sh.do_advance = 0;
      sh.cuts = kc.cuts[w_start];
    }
    if (tid < 5) sh.reward[tid] = kc.reward[tid];
    const int live = (w_start + 1) * DQLB200_CELLS_PER_LEVEL;
    for (int i = tid; i < live; i += NT) {
      sh.qa[i] = __uint_as_float(gt[i]);
      sh.cnt[i] = gt[2 * CELLS + i];
      if ((i & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(gqb + i));     // table B rows for the first snapshot
    }
  }
  __syncthreads();

  const float* __restrict__ alpha_lut = args.alpha_luts + (size_t)pp.alpha_lut * DQLB200_ALPHA_LUT;
  const float alpha_min = __ldg(alpha_lut + DQLB200_ALPHA_LUT - 1);      // alpha(count >= 1002), PKG/trainer.py:95-105
  uint64_t steps_done = 0;

  // R13/R14 end of a curriculum step: transfer (PKG/double_q_learning.py:77-89), window handling, next working step,
  // fresh env + TrainingMdp for every env (PKG/trainer.py:176-189, 232-245).  All threads call it (uniform).
  auto advance_curriculum = [&](int w, uint32_t birth) {
    const int cs = kc.curriculum_steps;
    int dst = -1, src = 0;
    float ratio = 1.0f;
    if (kc.transfer_mode == 0) { dst = w; src = (w - 1 + cs) % cs; ratio = kc.transfer_ratio[w]; }
    else if (w + 1 < cs) { dst = w + 1; src = w; ratio = kc.transfer_ratio[w + 1]; }
    if (dst >= 0) {
      // replica-merge mode: the transfer acts on the MERGED table (every replica applies it identically right after a
      // merge), so the first replica of a group also refreshes the group's merge snapshot
      uint32_t* sg = (args.merge_snapshot && pop % kc.replicas == 0) ? args.merge_snapshot + (size_t)(pop / kc.replicas) * 3 * CELLS : nullptr;
      for (int i = tid; i < DQLB200_CELLS_PER_LEVEL; i += NT) {
        // a source row above the working step is not staged: it is unmodified in global memory
        const float q_src = (src <= w) ? sh.qa[src * DQLB200_CELLS_PER_LEVEL + i] : __uint_as_float(gt[src * DQLB200_CELLS_PER_LEVEL + i]);
        const float qa_new = fmul(q_src, ratio), qb_new = fmul(gqb[src * DQLB200_CELLS_PER_LEVEL + i], ratio);
        sh.qa[dst * DQLB200_CELLS_PER_LEVEL + i] = qa_new;
        gqb[dst * DQLB200_CELLS_PER_LEVEL + i] = qb_new;
        if (sg) {
          sg[dst * DQLB200_CELLS_PER_LEVEL + i] = __float_as_uint(qa_new);
          sg[CELLS + dst * DQLB200_CELLS_PER_LEVEL + i] = __float_as_uint(qb_new);
        }
      }
    }
    if (w + 1 < cs) {      // level w+1 becomes live: stage its rows (Q_a unless the transfer just wrote it)
      for (int i = tid; i < DQLB200_CELLS_PER_LEVEL; i += NT) {
        const int c = (w + 1) * DQLB200_CELLS_PER_LEVEL + i;
        if (dst != w + 1) sh.qa[c] = __uint_as_float(gt[c]);
        sh.cnt[c] = gt[2 * CELLS + c];
      }
    }
    __syncthreads();
    if (tid == 0) {
      dqlb200_population_state& ps = sh.ps;
      if (sh.promote) { ps.window_head = ps.window_count = ps.window_sum = 0; }
      ps.promoted_at[w] = birth;
      ps.episodes_in_step = 0;
      ps.pending_advance = 0;
      sh.promote = sh.advance = sh.do_advance = 0;
      if (w + 1 >= cs) ps.finished = 1;
      else {
        ps.working_step = w + 1;
        sh.cuts = kc.cuts[w + 1];
      }
    }
    __syncthreads();
    if (!sh.ps.finished) {
      for (int slot = 0; slot < n_slots; ++slot) {
        const int env_i = slot * NT + tid;
        if (env_i < n_p) {
          Env e;
          env_reset(kc, pp, sh.cuts, kc.angle_cut, e, (uint32_t)env_i, birth, w + 1, /*fresh_mdp=*/true);
          env_store(args.env, env_base + env_i, e);
        }
      }
    }
    __syncthreads();
  };
  // replica-merge mode: a promotion decided by replica_merge_kernel takes effect before the first step of this launch
  if (sh.ps.pending_advance && !sh.ps.finished) {
    if (tid == 0) sh.promote = (sh.ps.pending_advance == 1) ? 1 : 0;
    __syncthreads();
    advance_curriculum(sh.ps.working_step, sh.ps.t);
    (void)env_prefetch_take(stage, NT, tid);
    env_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), stage, NT, tid);     // every env was just restarted
  }

  for (int k = 0; k < args.k_steps; ++k) {
    if (sh.ps.finished) break;     // uniform: written only between barriers
    const int w = sh.ps.working_step;
    const uint32_t t = sh.ps.t;
    // snapshot of the step: greedy action (first max of (Q_a+Q_b)/2, PKG/double_q_learning.py:119-124) and bootstrap
    // value max_a Q_a (:136-141) of every live state
    for (int st = tid; st < (w + 1) * DQLB200_STATES_PER_LEVEL; st += NT) {
      const float q0 = sh.qa[st * 3 + 0], q1 = sh.qa[st * 3 + 1], q2 = sh.qa[st * 3 + 2];
      const float p0 = fmul(fadd(q0, gqb[st * 3 + 0]), 0.5f);
      const float p1 = fmul(fadd(q1, gqb[st * 3 + 1]), 0.5f);
      const float p2 = fmul(fadd(q2, gqb[st * 3 + 2]), 0.5f);
      int a = 0;
      float best = p0;
      if (p1 > best) { best = p1; a = 1; }
      if (p2 > best) { a = 2; }
      sh.greedy[st] = (uint8_t)a;
      sh.qmax[st] = fmaxf(fmaxf(q0, q1), q2);
    }
    __syncthreads();
    int n_queued = 0;              // warp-uniform: envs of this warp that finished an episode in this step

    // batched R1/R8: new episodes for the queued envs of this warp, all lanes busy (a warp would otherwise run
    // the whole reset path for one or two lanes in half of its slots)
    auto flush_resets = [&]() {
      __syncwarp();
      for (int base = 0; base < n_queued; base += 32) {
        if (base + lane < n_queued) {
          const int qv = reset_queue[base + lane];
          const int env_r = (qv >> 5) * NT + warp * 32 + (qv & 31);
          const size_t gr = env_base + (size_t)env_r;
          Env e;
          env_load(args.env, gr, e);
          env_reset(kc, pp, sh.cuts, kc.angle_cut, e, (uint32_t)env_r, t + 1u, w, /*fresh_mdp=*/false);
          env_store(args.env, gr, e);
        }
      }
      __syncwarp();
      n_queued = 0;
    };

    for (int slot = 0; slot < n_slots; ++slot) {
      const int env_i = slot * NT + tid;
      const bool valid = env_i < n_p;
      const size_t gi = env_base + (size_t)(valid ? env_i : 0);
      const EnvRaw cur_raw = env_prefetch_take(stage, NT, tid);
      if (env_i + NT < n_p) env_prefetch_async(args.env, gi + NT, stage, NT, tid);      // in flight during this slot
      // ---------------- phase A: everything that only reads the snapshot ----------------------
      uint32_t cell = 0;
      float target = 0.0f;
      bool done = false, success = false;
      int code = 0;
      uint32_t ep_steps = 0;
      double ep_return = 0.0;
      Env e;
      uint32_t c_hint = 0;
      float a_hint = 0.0f;
      if (valid) {
        env_unpack(cur_raw, e);
        const uint32_t sid = e.sid;
        // R9/R10: epsilon-greedy on the snapshot; both draws are always consumed (quirk Q4).  With eps = 0
        // (every working step > 0) no draw can change the outcome and the Philox call is skipped.
        int a = sh.greedy[sid];
        if (w == 0) {
          const uint32_t thr = __ldg(args.eps_threshold + min(e.episode, (uint32_t)(DQLB200_EPS_LUT - 1)));
          const uint4 d = philox4x32_10(make_uint4((uint32_t)env_i, t, PURPOSE_STEP, pp.population_id), pp.seed_lo, pp.seed_hi);
          if ((d.x >> 8) < thr) a = (int)__umulhi(d.y, 3u);
        }
        size_t trace_i = 0;
        if (TRACE) {
          trace_i = (size_t)k * (size_t)args.n_total + gi;
          if (args.trace.action_override) {
            const int o = args.trace.action_override[trace_i];
            if (o >= 0) a = o;
          }
        }
        // learning-rate hint for phase B: the LUT entry of the cell's count as it is NOW (an unsynchronised peek at the
        // live table; phase B uses it only if the count is still the same, so the result does not depend on it).  Issued
        // here, a whole phase A before the baton: a barrier waits for the thread's outstanding global loads too.
        cell = sid * 3u + (uint32_t)a;
        c_hint = min(sh.cnt[cell], (uint32_t)(DQLB200_ALPHA_LUT - 1));
        a_hint = __ldg(alpha_lut + c_hint);
        // R3: set-point (float64).  A fresh episode starts from 0 but keeps the old value for shaping.
        const double prev_sp = e.theta_sp;
        const double sp = apply_action(kc, e.fresh ? 0.0 : e.theta_sp, a);
        // R4
        dyn_advance(kc, pp, e.b, (float)sp);
        const uint32_t step_count = e.step_count + 1u;
        const Obs o = dyn_observe(kc, pp, e.b, (int)step_count, kc.dz_train);
        // R5
        const DState ds = discretise_cuts(sh.cuts, kc.angle_cut, o, w);
        const uint32_t sid2 = (uint32_t)ds.id();
        // R6 (sticky result: only ever set, quirk Q9)
        // The priority chain of PKG/mdp.py:359-425 as selects (the ladder of branches diverges inside a warp).
        const bool t_fx = !(o.rel_p >= kc.fz_lo) || (o.rel_p >= kc.fz_hi);
        const bool t_zmin = !(o.z >= kc.z_min_cut), t_zmax = o.z >= kc.z_max_cut;
        const bool t_time = (int)step_count >= kc.timeout_steps;
        const bool goal_bins = !(o.contact || t_fx || t_zmin || t_zmax || t_time) && ds.bp == 1 && ds.bv == 1;
        const bool at_level = sid >= (uint32_t)(w * DQLB200_STATES_PER_LEVEL) && ds.level == w;   // previous level == w (it never exceeds w)
        const uint32_t cc = goal_bins ? (at_level ? e.curriculum_check + 1u : 0u) : e.curriculum_check;
        code = e.sticky_success ? DQLB200_NON_TERMINAL_SUCCESS : DQLB200_NON_TERMINAL;
        if (goal_bins && at_level) code = ((int)cc >= kc.success_steps) ? DQLB200_TERMINAL_SUCCESS : DQLB200_NON_TERMINAL_SUCCESS;
        code = t_time ? DQLB200_TERMINAL_TIMEOUT : code;
        code = t_zmax ? DQLB200_TERMINAL_FLYZONE_Z : code;
        code = t_zmin ? DQLB200_TERMINAL_MINIMUM_ALTITUDE : code;
        code = t_fx ? DQLB200_TERMINAL_FLYZONE_X : code;
        code = o.contact ? DQLB200_TERMINAL_CONTACT : code;
        done = code >= DQLB200_TERMINAL_SUCCESS;
        success = code == DQLB200_TERMINAL_SUCCESS;
        if (!(fabsf(o.rel_p) <= 3.4028234664e38f) || !(fabsf(o.rel_v) <= 3.4028234664e38f) || !(fabsf(o.rel_a) <= 3.4028234664e38f))
          atomicOr(&sh.ps.error_flags, 1u);      // NaN/inf observation (PKG/mdp.py:170 raises)
        // R7 (float64, reference operation order; level-dependent constants from the host)
        const double phi_p = shaping(kc.w_p, o.rel_p, kc.p_max, kc.rcp_p_max, kc.clip_p_f, DIV2);
        const double phi_v = shaping(kc.w_v, o.rel_v, kc.v_max, kc.rcp_v_max, kc.clip_v_f, DIV2);
        const double phi_t = __dmul_rn(kc.w_theta, fabs(div_f64_by_const(sp, kc.theta_max, kc.rcp_theta_max)));
        const double prev_p = shaping(kc.w_p, e.prev_rel_p, kc.p_max, kc.rcp_p_max, kc.clip_p_f, DIV2);
        const double prev_v = shaping(kc.w_v, e.prev_rel_v, kc.v_max, kc.rcp_v_max, kc.clip_v_f, DIV2);
        const double prev_t = __dmul_rn(kc.w_theta, fabs(div_f64_by_const(prev_sp, kc.theta_max, kc.rcp_theta_max)));
        const bool succ_reward = code == DQLB200_NON_TERMINAL_SUCCESS || code == DQLB200_TERMINAL_SUCCESS;
        const double r = reward_f64(kc, sh.reward[ds.level], phi_p, phi_v, phi_t, prev_p, prev_v, prev_t, succ_reward);
        // R12 target: r + (gamma * max_a Q_a[s'][a]) * [p-bin changed]   (quirks Q2, Q3), float32 like NEP 50
        const float qn = sh.qmax[sid2];
        const float changed = (e.bp != (uint32_t)ds.bp) ? 1.0f : 0.0f;
        target = fadd((float)r, fmul(fmul(kc.gamma, qn), changed));
        if (TRACE) {
          if (args.trace.obs) {
            float* po = args.trace.obs + trace_i * 5;
            po[0] = o.rel_p; po[1] = o.rel_v; po[2] = o.rel_a; po[3] = o.pitch; po[4] = o.z;
          }
          if (args.trace.reward) args.trace.reward[trace_i] = r;
          if (args.trace.action) args.trace.action[trace_i] = (uint8_t)a;
          if (args.trace.code) args.trace.code[trace_i] = (uint8_t)code;
          if (args.trace.done) args.trace.done[trace_i] = (uint8_t)done;
          if (args.trace.contact) args.trace.contact[trace_i] = (uint8_t)o.contact;
          if (args.trace.state) args.trace.state[trace_i] = (uint16_t)sid;
          if (args.trace.next_state) args.trace.next_state[trace_i] = (uint16_t)sid2;
          if (args.trace.episode) args.trace.episode[trace_i] = (int32_t)e.episode;
        }
        // carry, without a branch on `done`: of a finished env only the shaping memory and the episode index survive -- the
        // batched reset pass after the slot loop (R1/R8) overwrites every other field -- so all fields are written alike.
        ep_steps = step_count;
        ep_return = e.cum_reward;                 // quirk Q12: the last reward is not in the logged sum
        e.theta_sp = sp;
        e.prev_rel_p = o.rel_p;
        e.prev_rel_v = o.rel_v;
        e.episode += done ? 1u : 0u;
        e.sid = sid2;
        e.bp = (uint32_t)ds.bp;
        e.step_count = step_count;
        e.curriculum_check = cc;
        e.sticky_success = (code == DQLB200_NON_TERMINAL_SUCCESS);
        e.fresh = false;
        e.cum_reward = __dadd_rn(e.cum_reward, r);
      }
      // ---------------- phase B: ordered commit (baton between warps) --------------------------
      // The serialised section is the critical path of a global step (n_p / 32 links per population), so everything that
      // does not read the live table happens BEFORE the baton arrives: same-cell groups, ranks, and the reductions over
      // the finished episodes of this warp-slot.
      const uint32_t dmask = __ballot_sync(FULL, valid && done);
      const uint32_t key = valid ? cell : (0x80000000u | (uint32_t)lane);
      const uint32_t peers = __match_any_sync(FULL, key);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      uint32_t smask = 0u;
      double ret = 0.0, last_cum = 0.0;
      int last_steps = 0, last_code = 0;
      if (dmask) {
        smask = __ballot_sync(FULL, valid && success);
        // deterministic (fixed-tree) sum of the finished episodes' returns
        ret = (valid && done) ? ep_return : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ret = __dadd_rn(ret, __shfl_xor_sync(FULL, ret, off));
        const int last = 31 - __clz(dmask);
        last_steps = __shfl_sync(FULL, (int)ep_steps, last);
        last_code = __shfl_sync(FULL, code, last);
        last_cum = __shfl_sync(FULL, ep_return, last);
      }
      if (WARPS > 1 && !(slot == 0 && warp == 0)) baton_wait<WARPS>(warp);
      {
        float q = valid ? sh.qa[cell] : 0.0f;
        const uint32_t c0 = valid ? sh.cnt[cell] : 0u;
        const uint32_t c_pre = c0 + (uint32_t)rank;                                  // R11: pre-increment count
        float alpha = (c_pre == c_hint) ? a_hint : alpha_min;
        if (c_pre != c_hint && c_pre < (uint32_t)(DQLB200_ALPHA_LUT - 1)) alpha = __ldg(alpha_lut + c_pre);
        // the group's updates in lane order, two members per round (their four shuffles are issued together)
        uint32_t rem = valid ? peers : 0u;
        while (__any_sync(FULL, rem != 0u)) {
          const uint32_t rem1 = rem & (rem - 1u);
          const int src0 = rem ? (__ffs(rem) - 1) : lane, src1 = rem1 ? (__ffs(rem1) - 1) : lane;
          const float a_0 = __shfl_sync(FULL, alpha, src0), t_0 = __shfl_sync(FULL, target, src0);
          const float a_1 = __shfl_sync(FULL, alpha, src1), t_1 = __shfl_sync(FULL, target, src1);
          if (rem) q = fadd(q, fmul(a_0, fsub(t_0, q)));       // q += alpha * (target - q)
          if (rem1) q = fadd(q, fmul(a_1, fsub(t_1, q)));
          rem = rem1 & (rem1 - 1u);
        }
        if (valid && rank == 0) {
          sh.qa[cell] = q;
          sh.cnt[cell] = c0 + (uint32_t)__popc(peers);
        }
        // finished episodes, in env order: success window + promotion test after every append (R14)
        if (dmask && lane == 0) {
          dqlb200_population_state& ps = sh.ps;
          int head = ps.window_head, count = ps.window_count, sum = ps.window_sum;
          long long eps = ps.episodes_in_step;
          bool promote = false, advance = false;
          uint32_t m = dmask;
          while (m) {
            const int b = __ffs(m) - 1;
            m &= m - 1u;
            const int ok = (smask >> b) & 1u;
            if (count == kc.window_len) sum -= ps.window[head];
            else count += 1;
            ps.window[head] = (uint8_t)ok;
            sum += ok;
            head = (head + 1 == kc.window_len) ? 0 : head + 1;
            eps += 1;
            promote = promote || (sum >= kc.promote_successes);
            advance = advance || (eps >= kc.max_num_episodes);
          }
          ps.window_head = head; ps.window_count = count; ps.window_sum = sum;
          ps.episodes_in_step = eps;
          if (kc.replicas == 1) {      // replicas are promoted together by replica_merge_kernel
            if (promote) sh.promote = 1;
            if (advance) sh.advance = 1;
          }
          ps.return_sum = __dadd_rn(ps.return_sum, ret);
          ps.last_code = last_code;
          ps.last_steps = last_steps;
          ps.last_cumulative = last_cum;
        }
      }
      if (WARPS > 1 && !(slot == n_slots - 1 && warp == WARPS - 1)) {
        __threadfence_block();
        baton_pass<WARPS>((warp + 1) % WARPS);
      }
      // order-independent episode counters: after the baton
      if (dmask) {
        if (valid && done) {
          atomicAdd(&sh.step_hist[code], 1u);
          atomicAdd(&sh.step_ep_steps, ep_steps);
        }
        if (lane == 0) {
          atomicAdd(&sh.step_episodes, (uint32_t)__popc(dmask));
          atomicAdd(&sh.step_success, (uint32_t)__popc(smask));
        }
      }
      // The env state is written AFTER the baton: a barrier waits for the thread's outstanding global stores, which
      // would put an L2 round trip into the serialised section (ncu: stall_lg on the named barrier).
      if (valid) env_store(args.env, gi, e);
      // queue the finished envs of this warp for the batched reset (outside the baton)
      if (dmask) {
        if (valid && done) reset_queue[n_queued + __popc(dmask & ((1u << lane) - 1u))] = (uint16_t)(slot * 32 + lane);
        n_queued += __popc(dmask);
        if (n_queued > RESET_QUEUE - 32) flush_resets();       // never overflows, whatever fraction of envs finishes at once
      }
    }
    flush_resets();
    // software prefetch of slot 0 of the next global step (this warp's envs are final: resets only touch the warp's own)
    if (k + 1 < args.k_steps) env_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), stage, NT, tid);
    __syncthreads();
    // ---------------- end of the global step: promotion / next curriculum step (R13, R14) -----
    steps_done += (uint64_t)n_p;
    if (tid == 0) {
      sh.ps.t = t + 1u;
      sh.do_advance = (sh.promote || sh.advance) ? 1 : 0;
      sh.n_episodes += sh.step_episodes; sh.n_success += sh.step_success; sh.ep_steps += sh.step_ep_steps;
      sh.step_episodes = sh.step_success = sh.step_ep_steps = 0u;
    }
    if (tid < 9) { sh.hist[tid] += sh.step_hist[tid]; sh.step_hist[tid] = 0u; }
    __syncthreads();
    if (sh.do_advance) {
      advance_curriculum(w, t + 1u);
      (void)env_prefetch_take(stage, NT, tid);
      env_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), stage, NT, tid);
    }
  }

  // ---- write back (live rows only) ----------------------------------------------------------------
  asm volatile("cp.async.wait_all;" ::: "memory");      // a prefetch issued for a step that did not run
  __syncthreads();
  for (int i = tid; i < (sh.ps.working_step + 1) * DQLB200_CELLS_PER_LEVEL; i += NT) {
    gt[i] = __float_as_uint(sh.qa[i]);
    gt[2 * CELLS + i] = sh.cnt[i];
  }
  if (tid == 0) {
    dqlb200_population_state& ps = sh.ps;
    ps.total_steps += steps_done;
    ps.total_episodes += sh.n_episodes;
    ps.total_successes += sh.n_success;
    ps.episode_steps_sum += sh.ep_steps;
    for (int i = 0; i < 9; ++i) ps.termination_hist[i] += sh.hist[i];
  }
  __syncthreads();
  {
    uint32_t* gps = reinterpret_cast<uint32_t*>(args.pop_state + pop);
    for (int i = tid; i < PS_WORDS; i += NT) gps[i] = reinterpret_cast<const uint32_t*>(&sh.ps)[i];
  }
}

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

// -------------------------------------------------------------------------------------------------
__global__ void reset_kernel(const __grid_constant__ KC kc, EnvPtrs env, dqlb200_population_state* pop_state,
                             const dqlb200_population_params* pop_params, int initial_step) {
  const int pop = blockIdx.y;
  const int env_i = blockIdx.x * blockDim.x + threadIdx.x;
  __shared__ dqlb200_cuts cuts;
  if (threadIdx.x == 0) cuts = kc.cuts[initial_step];
  __syncthreads();
  if (env_i < kc.envs_per_population) {
    const dqlb200_population_params pp = pop_params[pop];
    Env e;
    env_reset(kc, pp, cuts, kc.angle_cut, e, (uint32_t)env_i, 0u, initial_step, /*fresh_mdp=*/true);
    env_store(env, (size_t)pop * kc.envs_per_population + env_i, e);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    dqlb200_population_state ps;
    memset(&ps, 0, sizeof(ps));
    ps.working_step = initial_step;
    pop_state[pop] = ps;
  }
}

// -------------------------------------------------------------------------------------------------
// R15: greedy evaluation, SimulationMdp semantics (PKG/mdp.py:784-886, scripts/simulation.py:48-63)
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) eval_kernel(const __grid_constant__ KC kc, const dqlb200_population_params* pop_params,
                                                   int population, const uint8_t* __restrict__ policy,
                                                   long long first_episode, long long n_episodes, int w,
                                                   dqlb200_eval_stats* stats, dqlb200_trace trace, int trace_steps) {
  __shared__ uint8_t s_policy[DQLB200_MAX_CURRICULUM * DQLB200_STATES_PER_LEVEL];
  __shared__ dqlb200_cuts cuts;
  __shared__ unsigned long long s_hist[9], s_steps, s_eps;
  for (int i = threadIdx.x; i < DQLB200_MAX_CURRICULUM * DQLB200_STATES_PER_LEVEL; i += blockDim.x) s_policy[i] = policy[i];
  if (threadIdx.x == 0) { cuts = kc.cuts[w]; s_steps = s_eps = 0ull; }
  if (threadIdx.x < 9) s_hist[threadIdx.x] = 0ull;
  __syncthreads();
  const dqlb200_population_params pp = pop_params[population];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_episodes) {
    const unsigned long long ep = (unsigned long long)(first_episode + i);
    const uint4 d = philox4x32_10(make_uint4((uint32_t)ep, 0u, PURPOSE_RESET, pp.population_id), pp.seed_lo, pp.seed_hi);
    Body b;
    Obs o = dyn_reset(kc, pp, b, d, /*normal_init=*/false, /*simulation=*/true, kc.dz_sim);
    uint32_t sid = (uint32_t)discretise_cuts(cuts, kc.angle_cut, o).id();
    double sp = 0.0;
    int code = DQLB200_NON_TERMINAL;
    int step = 0;
    while (code < DQLB200_TERMINAL_SUCCESS) {
      const int a = s_policy[sid];
      sp = apply_action(kc, sp, a);
      dyn_advance(kc, pp, b, (float)sp);
      step += 1;
      o = dyn_observe(kc, pp, b, step, kc.dz_sim);
      const uint32_t sid2 = (uint32_t)discretise_cuts(cuts, kc.angle_cut, o).id();
      if (o.contact) code = DQLB200_TERMINAL_CONTACT;
      else if (!(o.rel_p >= kc.fz_lo) || (o.rel_p >= kc.fz_hi)) code = DQLB200_TERMINAL_FLYZONE_X;
      else if (!(o.z >= kc.z_min_cut)) code = DQLB200_TERMINAL_MINIMUM_ALTITUDE;
      else if (o.z >= kc.z_max_cut) code = DQLB200_TERMINAL_FLYZONE_Z;
      else if (step >= kc.timeout_steps) code = DQLB200_TERMINAL_TIMEOUT;
      if (step <= trace_steps) {
        const size_t ti = (size_t)(step - 1) * (size_t)n_episodes + (size_t)i;
        if (trace.obs) {
          float* po = trace.obs + ti * 5;
          po[0] = o.rel_p; po[1] = o.rel_v; po[2] = o.rel_a; po[3] = o.pitch; po[4] = o.z;
        }
        if (trace.action) trace.action[ti] = (uint8_t)a;
        if (trace.code) trace.code[ti] = (uint8_t)code;
        if (trace.done) trace.done[ti] = (uint8_t)(code >= DQLB200_TERMINAL_SUCCESS);
        if (trace.contact) trace.contact[ti] = (uint8_t)o.contact;
        if (trace.state) trace.state[ti] = (uint16_t)sid;
        if (trace.next_state) trace.next_state[ti] = (uint16_t)sid2;
      }
      sid = sid2;
    }
    atomicAdd(&s_hist[code], 1ull);
    atomicAdd(&s_steps, (unsigned long long)step);
    atomicAdd(&s_eps, 1ull);
  }
  __syncthreads();
  if (threadIdx.x < 9 && s_hist[threadIdx.x]) atomicAdd((unsigned long long*)&stats->termination_hist[threadIdx.x], s_hist[threadIdx.x]);
  if (threadIdx.x == 0) {
    atomicAdd((unsigned long long*)&stats->steps, s_steps);
    atomicAdd((unsigned long long*)&stats->episodes, s_eps);
  }
}

// -------------------------------------------------------------------------------------------------
// Un-fused environment entry points (the gym surface of the reference: TrainingLandingEnv / SimulationLandingEnv reset()
// and step(), PKG/landing_simulation_env.py:167-282, 327-400): the caller supplies the actions, no agent, no table.  Same
// device functions and the same operation order as phase A of train_kernel; tests/test_gpu_facade.py holds the two
// bit-identical (a traced train launch with forced actions == a sequence of env steps).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) env_reset_kernel(const __grid_constant__ KC kc, EnvPtrs env, const dqlb200_population_params* pop_params,
                                                        int w, uint32_t birth, const uint8_t* __restrict__ mask, int fresh_mdp, int simulation,
                                                        uint16_t* out_state) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_total = (long long)kc.n_populations * kc.envs_per_population;
  if (i >= n_total) return;
  Env e;
  env_load(env, (size_t)i, e);
  if (!mask || mask[i]) {
    const int pop = (int)(i / kc.envs_per_population);
    const uint32_t env_i = (uint32_t)(i % kc.envs_per_population);
    const dqlb200_population_params pp = pop_params[pop];
    if (simulation) {      // SimulationLandingEnv.reset (PKG/landing_simulation_env.py:327-340) + SimulationMdp.reset (PKG/mdp.py:879-886)
      const uint4 d = philox4x32_10(make_uint4(env_i, birth, PURPOSE_RESET, pp.population_id), pp.seed_lo, pp.seed_hi);
      const Obs o = dyn_reset(kc, pp, e.b, d, /*normal_init=*/false, /*simulation=*/true, kc.dz_sim);
      const DState ds = discretise_cuts(kc.cuts[w], kc.angle_cut, o);
      e.sid = (uint32_t)ds.id(); e.bp = (uint32_t)ds.bp;
      e.step_count = 0; e.curriculum_check = 0; e.sticky_success = false; e.fresh = true; e.cum_reward = 0.0;
      e.theta_sp = 0.0; e.prev_rel_p = 0.0f; e.prev_rel_v = 0.0f;
      if (fresh_mdp) e.episode = 0;
    } else {
      env_reset(kc, pp, kc.cuts[w], kc.angle_cut, e, env_i, birth, w, fresh_mdp != 0);
    }
    env_store(env, (size_t)i, e);
  }
  if (out_state) out_state[i] = (uint16_t)e.sid;
}

template <bool DIV2>
__global__ void __launch_bounds__(128) env_step_kernel(const __grid_constant__ KC kc, EnvPtrs env, const dqlb200_population_params* pop_params,
                                                       int w, uint32_t t, const int8_t* __restrict__ actions, int auto_reset, int simulation,
                                                       uint16_t* out_state, double* out_reward, uint8_t* out_code, uint8_t* out_done,
                                                       float* out_obs, uint32_t* out_steps, double* out_cumulative, uint32_t* error_flag) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_total = (long long)kc.n_populations * kc.envs_per_population;
  if (i >= n_total) return;
  const int pop = (int)(i / kc.envs_per_population);
  const uint32_t env_i = (uint32_t)(i % kc.envs_per_population);
  const dqlb200_population_params pp = pop_params[pop];
  const dqlb200_cuts& cuts = kc.cuts[w];
  Env e;
  env_load(env, (size_t)i, e);
  const int a = actions[i];
  // R3 .. R8 in the order of TrainingLandingEnv.step (PKG/landing_simulation_env.py:245-282)
  const double prev_sp = e.theta_sp;
  const double sp = apply_action(kc, e.fresh ? 0.0 : e.theta_sp, a);
  dyn_advance(kc, pp, e.b, (float)sp);
  const uint32_t step_count = e.step_count + 1u;
  const Obs o = dyn_observe(kc, pp, e.b, (int)step_count, simulation ? kc.dz_sim : kc.dz_train);
  const DState ds = discretise_cuts(cuts, kc.angle_cut, o, w);
  const uint32_t sid2 = (uint32_t)ds.id();
  const bool t_fx = !(o.rel_p >= kc.fz_lo) || (o.rel_p >= kc.fz_hi);
  const bool t_zmin = !(o.z >= kc.z_min_cut), t_zmax = o.z >= kc.z_max_cut;
  const bool t_time = (int)step_count >= kc.timeout_steps;
  const bool goal_bins = !simulation && !(o.contact || t_fx || t_zmin || t_zmax || t_time) && ds.bp == 1 && ds.bv == 1;
  const bool at_level = e.sid >= (uint32_t)(w * DQLB200_STATES_PER_LEVEL) && ds.level == w;
  const uint32_t cc = goal_bins ? (at_level ? e.curriculum_check + 1u : 0u) : e.curriculum_check;
  int code = e.sticky_success ? DQLB200_NON_TERMINAL_SUCCESS : DQLB200_NON_TERMINAL;
  if (goal_bins && at_level) code = ((int)cc >= kc.success_steps) ? DQLB200_TERMINAL_SUCCESS : DQLB200_NON_TERMINAL_SUCCESS;
  code = t_time ? DQLB200_TERMINAL_TIMEOUT : code;
  code = t_zmax ? DQLB200_TERMINAL_FLYZONE_Z : code;
  code = t_zmin ? DQLB200_TERMINAL_MINIMUM_ALTITUDE : code;
  code = t_fx ? DQLB200_TERMINAL_FLYZONE_X : code;
  code = o.contact ? DQLB200_TERMINAL_CONTACT : code;
  const bool done = code >= DQLB200_TERMINAL_SUCCESS;
  if (!(fabsf(o.rel_p) <= 3.4028234664e38f) || !(fabsf(o.rel_v) <= 3.4028234664e38f) || !(fabsf(o.rel_a) <= 3.4028234664e38f))
    atomicOr(error_flag, 1u);
  double r = 0.0;
  if (!simulation) {
    const double phi_p = shaping(kc.w_p, o.rel_p, kc.p_max, kc.rcp_p_max, kc.clip_p_f, DIV2);
    const double phi_v = shaping(kc.w_v, o.rel_v, kc.v_max, kc.rcp_v_max, kc.clip_v_f, DIV2);
    const double phi_t = __dmul_rn(kc.w_theta, fabs(div_f64_by_const(sp, kc.theta_max, kc.rcp_theta_max)));
    const double prev_p = shaping(kc.w_p, e.prev_rel_p, kc.p_max, kc.rcp_p_max, kc.clip_p_f, DIV2);
    const double prev_v = shaping(kc.w_v, e.prev_rel_v, kc.v_max, kc.rcp_v_max, kc.clip_v_f, DIV2);
    const double prev_t = __dmul_rn(kc.w_theta, fabs(div_f64_by_const(prev_sp, kc.theta_max, kc.rcp_theta_max)));
    const bool succ_reward = code == DQLB200_NON_TERMINAL_SUCCESS || code == DQLB200_TERMINAL_SUCCESS;
    r = reward_f64(kc, kc.reward[ds.level], phi_p, phi_v, phi_t, prev_p, prev_v, prev_t, succ_reward);
  }
  if (out_reward) out_reward[i] = r;
  if (out_code) out_code[i] = (uint8_t)code;
  if (out_done) out_done[i] = (uint8_t)done;
  if (out_obs) { float* po = out_obs + i * 5; po[0] = o.rel_p; po[1] = o.rel_v; po[2] = o.rel_a; po[3] = o.pitch; po[4] = o.z; }
  if (out_steps) out_steps[i] = step_count;
  if (out_cumulative) out_cumulative[i] = e.cum_reward;          // quirk Q12: without this step's reward
  e.theta_sp = sp;
  e.prev_rel_p = o.rel_p;
  e.prev_rel_v = o.rel_v;
  e.episode += done ? 1u : 0u;
  e.sid = sid2;
  e.bp = (uint32_t)ds.bp;
  e.step_count = step_count;
  e.curriculum_check = cc;
  e.sticky_success = (code == DQLB200_NON_TERMINAL_SUCCESS);
  e.fresh = false;
  e.cum_reward = __dadd_rn(e.cum_reward, r);
  if (done && auto_reset && !simulation) env_reset(kc, pp, cuts, kc.angle_cut, e, env_i, t + 1u, w, /*fresh_mdp=*/false);
  if (out_state) out_state[i] = (uint16_t)e.sid;       // of a finished env with auto_reset: the first state of its next episode
  env_store(env, (size_t)i, e);
}

// -------------------------------------------------------------------------------------------------
// two-axis greedy evaluation.  One thread per episode; pitch drives x, roll drives y (signed gravity per
// axis), one platform under both (three trajectories).  Same operation order as oracle/dynamics.py: StandIn2D.
// -------------------------------------------------------------------------------------------------
struct Axis {
  float pos, vel, ang, acc;
};
__device__ __forceinline__ void axis_advance(const KC& kc, Axis& b, float sp, float g) {
  b.ang = fadd(b.ang, fmul(fsub(sp, b.ang), kc.k_theta));
  b.acc = fsub(fmul(g, det_tan(b.ang)), fmul(kc.c_d, b.vel));
  b.pos = fadd(fadd(b.pos, fmul(b.vel, kc.h)), fmul(b.acc, kc.half_h2));
  b.vel = fadd(b.vel, fmul(b.acc, kc.h));
}
struct Platform2D {
  float xm, um, axm, ym, vm, aym;
};
__device__ __forceinline__ Platform2D platform_2d(const dqlb200_eval2d_params& p, uint32_t phase_x, uint32_t phase_y) {
  Platform2D m;
  float sx, cx;
  det_sincos_turns(phase_x, sx, cx);
  if (p.trajectory == 2) {
    const float sc = fmul(sx, cx);
    m.xm = fmul(p.r_x, cx); m.um = -fmul(p.rw_x, sx); m.axm = -fmul(p.rw2_x, cx);
    m.ym = fmul(p.r_y, sc); m.vm = fmul(p.rw_y, fsub(fmul(cx, cx), fmul(sx, sx))); m.aym = -fmul(p.rw2_y, sc);
  } else {
    float sy, cy;
    det_sincos_turns(phase_y, sy, cy);
    m.xm = fmul(p.r_x, sx); m.um = fmul(p.rw_x, cx); m.axm = -fmul(p.rw2_x, sx);
    m.ym = fmul(p.r_y, sy); m.vm = fmul(p.rw_y, cy); m.aym = -fmul(p.rw2_y, sy);
  }
  return m;
}

__global__ void __launch_bounds__(256) eval2d_kernel(const __grid_constant__ KC kc, const __grid_constant__ dqlb200_eval2d_params p,
                                                     const uint8_t* __restrict__ policy_x, const uint8_t* __restrict__ policy_y,
                                                     long long first_episode, long long n_episodes, dqlb200_eval_stats* stats,
                                                     dqlb200_trace2d trace, int trace_steps) {
  __shared__ uint8_t s_pol_x[DQLB200_MAX_CURRICULUM * DQLB200_STATES_PER_LEVEL], s_pol_y[DQLB200_MAX_CURRICULUM * DQLB200_STATES_PER_LEVEL];
  __shared__ dqlb200_cuts cuts;
  __shared__ unsigned long long s_hist[9], s_steps, s_eps;
  for (int i = threadIdx.x; i < DQLB200_MAX_CURRICULUM * DQLB200_STATES_PER_LEVEL; i += blockDim.x) {
    s_pol_x[i] = policy_x[i];
    s_pol_y[i] = policy_y[i];
  }
  if (threadIdx.x == 0) { cuts = kc.cuts[p.working_step]; s_steps = s_eps = 0ull; }
  if (threadIdx.x < 9) s_hist[threadIdx.x] = 0ull;
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_episodes) {
    const unsigned long long ep = (unsigned long long)(first_episode + i);
    const uint4 d = philox4x32_10(make_uint4((uint32_t)ep, 0u, PURPOSE_RESET, p.stream_id), p.seed_lo, p.seed_hi);
    // PKG/landing_simulation_env.py:327-340: uniform offsets inside the fly zone, absolute clip, random platform phase
    const float x_init = fadd(-kc.p_max_f, fmul(kc.two_p_max_f, fmul(__uint2float_rn(d.x >> 8), (float)(1.0 / 16777216.0))));
    const float y_init = fadd(-kc.p_max_f, fmul(kc.two_p_max_f, fmul(__uint2float_rn(d.y >> 8), (float)(1.0 / 16777216.0))));
    uint32_t phase_x = d.z, phase_y = (p.trajectory == 2) ? d.z : d.w;
    Platform2D m = platform_2d(p, phase_x, phase_y);
    Axis bx, by;
    bx.pos = clipf(fsub(m.xm, x_init), -kc.p_max_f, kc.p_max_f);
    by.pos = p.y_init_enabled ? clipf(fsub(m.ym, y_init), -kc.p_max_f, kc.p_max_f) : 0.0f;
    bx.vel = bx.ang = bx.acc = by.vel = by.ang = by.acc = 0.0f;
    double sp_x = 0.0, sp_y = 0.0;
    int code = DQLB200_NON_TERMINAL, step = -1;
    uint32_t sid_x = 0, sid_y = 0;
    while (code < DQLB200_TERMINAL_SUCCESS) {
      int ax = 255, ay = 255;
      if (step >= 0) {          // step == -1: the hover period after the reset (PKG/landing_simulation_env.py:222-224)
        ax = s_pol_x[sid_x];
        ay = s_pol_y[sid_y];
        sp_x = apply_action(kc, sp_x, ax);
        if (p.y_action_enabled) sp_y = apply_action(kc, sp_y, ay);
      }
      for (int k = 0; k < kc.n_sub; ++k) {
        axis_advance(kc, bx, (float)sp_x, p.g_x);
        axis_advance(kc, by, (float)sp_y, p.g_y);
        phase_x += p.dphase_x;
        phase_y += p.dphase_y;
      }
      step += 1;
      m = platform_2d(p, phase_x, phase_y);
      Obs ox, oy;
      ox.rel_p = fsub(m.xm, bx.pos); ox.rel_v = fsub(m.um, bx.vel); ox.rel_a = fsub(m.axm, bx.acc); ox.pitch = bx.ang;
      oy.rel_p = fsub(m.ym, by.pos); oy.rel_v = fsub(m.vm, by.vel); oy.rel_a = fsub(m.aym, by.acc); oy.pitch = by.ang;
      const float z = fadd(kc.z_init, fmul(__int2float_rn(step), kc.dz_sim));
      const bool contact = (z <= kc.z_touch) && (fabsf(ox.rel_p) <= kc.half_platform) && (fabsf(oy.rel_p) <= kc.half_platform);
      sid_x = (uint32_t)discretise_cuts(cuts, kc.angle_cut, ox).id();
      sid_y = (uint32_t)discretise_cuts(cuts, kc.angle_cut, oy).id();
      if (step == 0) continue;          // the reset only observes (no check, PKG/landing_simulation_env.py:236-243)
      if (contact) code = DQLB200_TERMINAL_CONTACT;
      else if (!(ox.rel_p >= kc.fz_lo) || (ox.rel_p >= kc.fz_hi)) code = DQLB200_TERMINAL_FLYZONE_X;
      else if (!(oy.rel_p >= kc.fz_lo) || (oy.rel_p >= kc.fz_hi)) code = DQLB200_TERMINAL_FLYZONE_Y;
      else if (!(z >= kc.z_min_cut)) code = DQLB200_TERMINAL_MINIMUM_ALTITUDE;
      else if (z >= kc.z_max_cut) code = DQLB200_TERMINAL_FLYZONE_Z;
      else if (step >= kc.timeout_steps) code = DQLB200_TERMINAL_TIMEOUT;
      if (step <= trace_steps) {
        const size_t ti = (size_t)(step - 1) * (size_t)n_episodes + (size_t)i;
        if (trace.obs) {
          float* po = trace.obs + ti * 9;
          po[0] = ox.rel_p; po[1] = ox.rel_v; po[2] = ox.rel_a; po[3] = ox.pitch; po[4] = z;
          po[5] = oy.rel_p; po[6] = oy.rel_v; po[7] = oy.rel_a; po[8] = oy.pitch;
        }
        if (trace.action_x) trace.action_x[ti] = (uint8_t)ax;
        if (trace.action_y) trace.action_y[ti] = (uint8_t)ay;
        if (trace.code) trace.code[ti] = (uint8_t)code;
        if (trace.done) trace.done[ti] = (uint8_t)(code >= DQLB200_TERMINAL_SUCCESS);
        if (trace.contact) trace.contact[ti] = (uint8_t)contact;
        if (trace.state_x) trace.state_x[ti] = (uint16_t)sid_x;
        if (trace.state_y) trace.state_y[ti] = (uint16_t)sid_y;
      }
    }
    atomicAdd(&s_hist[code], 1ull);
    atomicAdd(&s_steps, (unsigned long long)step);
    atomicAdd(&s_eps, 1ull);
  }
  __syncthreads();
  if (threadIdx.x < 9 && s_hist[threadIdx.x]) atomicAdd((unsigned long long*)&stats->termination_hist[threadIdx.x], s_hist[threadIdx.x]);
  if (threadIdx.x == 0) {
    atomicAdd((unsigned long long*)&stats->steps, s_steps);
    atomicAdd((unsigned long long*)&stats->episodes, s_eps);
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
    const int src = (step - 1 + cs) % cs;
    const double ratio = alpha[0];
    for (int i = 0; i < DQLB200_CELLS_PER_LEVEL; ++i) {
      qa[step * DQLB200_CELLS_PER_LEVEL + i] = __dmul_rn(qa[src * DQLB200_CELLS_PER_LEVEL + i], ratio);
      qb[step * DQLB200_CELLS_PER_LEVEL + i] = __dmul_rn(qb[src * DQLB200_CELLS_PER_LEVEL + i], ratio);
    }
  }
}

// -------------------------------------------------------------------------------------------------
__global__ void transfer_kernel(uint32_t* tables, int n_pop, int cs, int step, float ratio) {
  const int pop = blockIdx.x;
  float* qa = reinterpret_cast<float*>(tables + (size_t)pop * 3 * CELLS);
  float* qb = qa + CELLS;
  const int src = (step - 1 + cs) % cs;
  for (int i = threadIdx.x; i < DQLB200_CELLS_PER_LEVEL; i += blockDim.x) {
    qa[step * DQLB200_CELLS_PER_LEVEL + i] = fmul(qa[src * DQLB200_CELLS_PER_LEVEL + i], ratio);
    qb[step * DQLB200_CELLS_PER_LEVEL + i] = fmul(qb[src * DQLB200_CELLS_PER_LEVEL + i], ratio);
  }
}

// Shared-table mode (one agent replicated on G devices).  Each replica trains on its own envs for a few
// steps; the replicas are then merged with a visit-weighted mean of their Q deltas and the sum of
// their visit counts:  Q <- Q_snap + sum_g(dcount_g * dQ_g) / sum_g(dcount_g),  count <- count_snap + sum_g dcount_g.
// One "agent" below = a group of R = replicas_per_population consecutive populations whose tables are identical (after
// replica_merge_kernel; R = 1: a plain population).  snap holds ONE [3][CELLS] entry per agent, delta ONE entry of
// DQLB200_SHARED_DELTA_WORDS floats per agent: [0] sum dQ*dcount, [1] sum dcount, [2] number of ranks that visited the cell,
// [3] sum of the visiting ranks' Q (exact when one rank visited: the value that rank keeps), then the agent's pooled trainer
// counters: successes in the windows, finished episodes of the curriculum step, number of ranks, ranks that are alive.
constexpr int DELTA_WORDS = DQLB200_SHARED_DELTA_WORDS;
__global__ void shared_pack_kernel(const uint32_t* tables, const uint32_t* snap, float* delta, const dqlb200_population_state* ps,
                                   int n_agents, int R) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_agents * CELLS) return;
  const long long g = i / CELLS, c = i % CELLS;
  const size_t tb = (size_t)g * R * 3 * CELLS, sb = (size_t)g * 3 * CELLS;
  float* d = delta + (size_t)g * DELTA_WORDS;
  const uint32_t dcu = tables[tb + 2 * CELLS + c] - snap[sb + 2 * CELLS + c];
  const float dc = (float)dcu, q = __uint_as_float(tables[tb + c]);
  d[c] = fmul(fsub(q, __uint_as_float(snap[sb + c])), dc);
  d[CELLS + c] = dc;
  d[2 * CELLS + c] = dcu ? 1.0f : 0.0f;
  d[3 * CELLS + c] = dcu ? q : 0.0f;
  if (c < 4) {
    long long successes = 0, episodes = 0;
    bool alive = true;
    for (int r = 0; r < R; ++r) {
      const dqlb200_population_state& p = ps[g * R + r];
      successes += p.window_sum;
      episodes += p.episodes_in_step;
      alive = alive && !p.finished && !p.pending_advance;
    }
    d[4 * CELLS + c] = c == 0 ? (float)successes : c == 1 ? (float)episodes : c == 2 ? 1.0f : (alive ? 1.0f : 0.0f);
  }
}
__global__ void shared_apply_kernel(uint32_t* tables, uint32_t* snap, uint32_t* merge_snap, const float* delta,
                                    dqlb200_population_state* ps, int n_agents, int R, int pooled_promote, long long max_episodes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_agents * CELLS) return;
  const long long g = i / CELLS, c = i % CELLS;
  const size_t tb = (size_t)g * R * 3 * CELLS, sb = (size_t)g * 3 * CELLS;
  const float* d = delta + (size_t)g * DELTA_WORDS;
  const float visitors = d[2 * CELLS + c];
  uint32_t q_bits = snap[sb + c], cnt = snap[sb + 2 * CELLS + c];
  if (visitors == 1.0f) q_bits = __float_as_uint(d[3 * CELLS + c]);       // one rank visited: its value, bit for bit, on every rank
  else if (visitors > 1.0f) q_bits = __float_as_uint(fadd(__uint_as_float(q_bits), __fdiv_rn(d[c], d[CELLS + c])));
  if (visitors > 0.0f) {
    cnt += (uint32_t)__float2uint_rn(d[CELLS + c]);
    for (int r = 0; r < R; ++r) {
      tables[tb + (size_t)r * 3 * CELLS + c] = q_bits;
      tables[tb + (size_t)r * 3 * CELLS + 2 * CELLS + c] = cnt;
    }
  } else {          // nobody visited: the cell can still have changed by the (identical) curriculum transfers of every copy
    q_bits = tables[tb + c];
  }
  const uint32_t qb_bits = tables[tb + CELLS + c];
  snap[sb + c] = q_bits; snap[sb + CELLS + c] = qb_bits; snap[sb + 2 * CELLS + c] = cnt;
  if (merge_snap) { merge_snap[sb + c] = q_bits; merge_snap[sb + CELLS + c] = qb_bits; merge_snap[sb + 2 * CELLS + c] = cnt; }
  if (c == 0 && pooled_promote > 0) {      // promotion pooled over every rank's windows (same decision on every rank)
    const float successes = d[4 * CELLS + 0], episodes = d[4 * CELLS + 1];
    const bool alive = d[4 * CELLS + 3] == d[4 * CELLS + 2];
    const int pending = !alive ? 0 : (successes >= (float)pooled_promote ? 1 : (episodes >= (float)max_episodes ? 2 : 0));
    if (pending)
      for (int r = 0; r < R; ++r) ps[g * R + r].pending_advance = pending;
  }
}

// Replica-merge mode: R consecutive populations are replicas of ONE agent.  One CTA (8 warps) per tile of 32 live cells:
//   load   : all warps stream the replicas' (count, Q_a) of the tile, lane <-> cell (128-byte coalesced rows), into shared memory,
//            MERGE_CHUNK replicas at a time -- every load is independent of every other;
//   reduce : warp 0 (lane <-> cell) accumulates the visitors of the chunk STRICTLY in replica order (the summation order
//            is part of the semantics: bit-exact vs oracle/loop.py) -- the only serial part, one dependent fadd per visitor;
//   write  : the merged value goes to every replica (all warps, coalesced) and to the snapshot.
// Only the live rows (levels 0..working step) can differ from the snapshot.  Thread 0 of block (0, g) pools the success
// windows and arms the promotion.
constexpr int MERGE_CHUNK = 128;
__global__ void __launch_bounds__(256) replica_merge_kernel(uint32_t* tables, uint32_t* snap, dqlb200_population_state* ps,
                                                            int R, int pooled_promote, long long max_episodes) {
  __shared__ uint32_t s_q[MERGE_CHUNK][32], s_dc[MERGE_CHUNK][32];
  __shared__ uint32_t s_qnew[32], s_cnew[32], s_vis[32];
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  uint32_t* sg = snap + (size_t)g * 3 * CELLS;
  const uint32_t* tg = tables + (size_t)g * R * 3 * CELLS;
  const int live = (ps[g * R].working_step + 1) * DQLB200_CELLS_PER_LEVEL;
  if (blockIdx.x * 32 < live) {                       // block-uniform
    const bool in = c < live;
    const float q_snap = in ? __uint_as_float(sg[c]) : 0.0f;
    const uint32_t cnt_snap = in ? sg[2 * CELLS + c] : 0u;
    float num = 0.0f, q_single = q_snap;
    uint32_t tot = 0;
    int visitors = 0;
    for (int r0 = 0; r0 < R; r0 += MERGE_CHUNK) {
      const int n = min(MERGE_CHUNK, R - r0);
      {   // MERGE_CHUNK / 8 replicas per warp: all their loads are issued before the first one is consumed
        uint32_t cv[MERGE_CHUNK / 8], qv[MERGE_CHUNK / 8];
#pragma unroll
        for (int i = 0; i < MERGE_CHUNK / 8; ++i) {
          const int j = warp + 8 * i;
          const uint32_t* tr = tg + (size_t)(r0 + min(j, n - 1)) * 3 * CELLS;
          cv[i] = in ? __ldcg(tr + 2 * CELLS + c) : cnt_snap;
          qv[i] = in ? __ldcg(tr + c) : 0u;
        }
#pragma unroll
        for (int i = 0; i < MERGE_CHUNK / 8; ++i) {
          const int j = warp + 8 * i;
          if (j < n) {
            s_dc[j][lane] = cv[i] - cnt_snap;
            s_q[j][lane] = qv[i];
          }
        }
      }
      __syncthreads();
      if (warp == 0) {
#pragma unroll 8
        for (int j = 0; j < n; ++j) {
          const uint32_t dc = s_dc[j][lane];
          const float q_r = __uint_as_float(s_q[j][lane]);
          const float term = fmul(fsub(q_r, q_snap), __uint2float_rn(dc));      // off the dependent chain
          if (dc) {
            visitors += 1;
            q_single = q_r;
            num = fadd(num, term);
            tot += dc;
          }
        }
      }
      __syncthreads();
    }
    if (warp == 0) {
      float q_new = q_snap;
      if (visitors == 1) q_new = q_single;
      else if (visitors > 1) q_new = fadd(q_snap, __fdiv_rn(num, __uint2float_rn(tot)));
      s_qnew[lane] = __float_as_uint(q_new);
      s_cnew[lane] = cnt_snap + tot;
      s_vis[lane] = (uint32_t)visitors;
      if (in && visitors) {
        sg[c] = __float_as_uint(q_new);
        sg[2 * CELLS + c] = cnt_snap + tot;
      }
    }
    __syncthreads();
    if (in && s_vis[lane]) {
      const uint32_t qn = s_qnew[lane], cn = s_cnew[lane];
      for (int r = warp; r < R; r += 8) {
        uint32_t* tr = tables + (size_t)(g * R + r) * 3 * CELLS;
        tr[c] = qn;
        tr[2 * CELLS + c] = cn;
      }
    }
  }
  if (blockIdx.x == 0) {          // pooled trainer counters of the group: block-wide reduction over the R replicas
    __shared__ unsigned long long s_succ, s_eps;
    __shared__ int s_dead, s_pending;
    if (threadIdx.x == 0) { s_succ = s_eps = 0ull; s_dead = 0; s_pending = 0; }
    __syncthreads();
    unsigned long long successes = 0, episodes = 0;
    int dead = 0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
      const dqlb200_population_state& p = ps[g * R + r];
      successes += (unsigned long long)p.window_sum;
      episodes += (unsigned long long)p.episodes_in_step;
      dead |= (p.finished || p.pending_advance) ? 1 : 0;
    }
    if (successes) atomicAdd(&s_succ, successes);
    if (episodes) atomicAdd(&s_eps, episodes);
    if (dead) atomicOr(&s_dead, 1);
    __syncthreads();
    if (threadIdx.x == 0)
      s_pending = (s_dead || pooled_promote <= 0) ? 0 : ((long long)s_succ >= pooled_promote ? 1 : ((long long)s_eps >= max_episodes ? 2 : 0));
    __syncthreads();
    const int pending = s_pending;
    if (pending)
      for (int r = threadIdx.x; r < R; r += blockDim.x) ps[g * R + r].pending_advance = pending;
  }
}

// Measurement aid: the table update as UNORDERED shared-memory atomics on a recorded cell sequence (the "atomic roof").
__global__ void table_rmw_roof_kernel(const uint16_t* __restrict__ cells, long long n_cells, int visits_per_thread,
                                      unsigned long long* checksum) {
  __shared__ float qa[CELLS];
  __shared__ uint32_t cnt[CELLS];
  for (int i = threadIdx.x; i < CELLS; i += blockDim.x) { qa[i] = 0.0f; cnt[i] = 0u; }
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long idx = ((long long)blockIdx.x * blockDim.x + threadIdx.x) % n_cells;
  for (int k = 0; k < visits_per_thread; ++k) {
    const uint32_t c = cells[idx];
    atomicAdd(&qa[c], 0.015625f);
    atomicAdd(&cnt[c], 1u);
    idx += stride;
    if (idx >= n_cells) idx %= n_cells;
  }
  __syncthreads();
  unsigned long long sum = 0;
  for (int i = threadIdx.x; i < CELLS; i += blockDim.x) sum += cnt[i] + (unsigned long long)qa[i];
  if (sum) atomicAdd(checksum, sum);
}

}  // namespace dql

// =================================================================================================
// C-ABI
// =================================================================================================
struct dqlb200_handle {
  dqlb200_config cfg;
  dql::KC kc;
  int device;
  dqlb200_config* d_cfg;
  float* d_alpha;
  dqlb200_population_params* d_pop_params;
  uint32_t* d_error;
  void* env_state;
  void* tables;
  void* pop_state;
  void* merge_snapshot;
  size_t smem_bytes;
  // dqlb200_train_host pipelines the populations in chunks over these streams (copy-in / train / copy-out overlap)
  static constexpr int MAX_HOST_CHUNKS = 8;
  cudaStream_t chunk_stream[MAX_HOST_CHUNKS];
  cudaEvent_t chunk_done[MAX_HOST_CHUNKS];
  cudaEvent_t host_start;
  bool chunk_ready;
  // dqlb200_train_merged replays ONE captured CUDA graph (train launch of `merged_k` steps + replica merge) on its own stream
  cudaStream_t merged_stream;
  cudaEvent_t merged_in, merged_out;
  cudaGraphExec_t merged_exec;
  int merged_k, merged_promote;
  bool merged_ready;
};

static thread_local std::string g_last_error;
static int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
#define CUDA_TRY(expr)                                                                                      \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) return fail(DQLB200_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

extern "C" {

int dqlb200_abi_version(void) { return DQLB200_ABI_VERSION; }
size_t dqlb200_config_bytes(void) { return sizeof(dqlb200_config); }
size_t dqlb200_population_state_bytes(void) { return sizeof(dqlb200_population_state); }
size_t dqlb200_eval2d_params_bytes(void) { return sizeof(dqlb200_eval2d_params); }
const char* dqlb200_last_error(void) { return g_last_error.c_str(); }

const char* dqlb200_termination_string(int code) {
  switch (code) {
    case DQLB200_TERMINAL_SUCCESS: return "SUCCESS: Goal state reached";
    case DQLB200_TERMINAL_CONTACT: return "SUCCESS: Touched platform";
    case DQLB200_TERMINAL_FLYZONE_X: return "FAILURE: Drone moved too far from platform in x direction";
    case DQLB200_TERMINAL_FLYZONE_Y: return "FAILURE: Drone moved too far from platform in y direction";
    case DQLB200_TERMINAL_FLYZONE_Z: return "FAILURE: Drone moved too far from platform in z direction";
    case DQLB200_TERMINAL_MINIMUM_ALTITUDE: return "FAILURE: Reached minimum altitude";
    case DQLB200_TERMINAL_TIMEOUT: return "FAILURE: Maximum episode duration";
    default: return nullptr;
  }
}

// Smallest non-negative fp32 x with RN((double)x / d) >= 1.0: np.clip(x / d, -1, 1) saturates exactly for |x| >= this cut
// (the correctly rounded quotient is monotone in x), so the shaping potentials clip on the fp32 observation.
static float first_f32_with_unit_quotient(double d) {
  float x = (float)d;
  while ((double)x / d >= 1.0) x = nextafterf(x, 0.0f);
  while (!((double)x / d >= 1.0)) x = nextafterf(x, INFINITY);
  return x;
}

static void fill_kc(const dqlb200_config& c, dql::KC& k) {
  memcpy(k.cuts, c.cuts, sizeof(k.cuts));
  memcpy(k.reward, c.reward, sizeof(k.reward));
  k.p_max = c.p_max; k.v_max = c.v_max; k.theta_max = c.theta_max; k.delta_theta = c.delta_theta;
  k.w_p = c.w_p; k.w_v = c.w_v; k.w_theta = c.w_theta;
  k.rcp_p_max = 1.0 / c.p_max; k.rcp_v_max = 1.0 / c.v_max; k.rcp_theta_max = 1.0 / c.theta_max;
  k.div_two_steps = (c.p_max == 4.5 && c.v_max == 3.39411) ? 0 : 1;
  k.clip_p_f = first_f32_with_unit_quotient(c.p_max);
  k.clip_v_f = first_f32_with_unit_quotient(c.v_max);
  memcpy(k.angle_cut, c.angle_cut, sizeof(k.angle_cut));
  k.fz_lo = c.fz_lo; k.fz_hi = c.fz_hi; k.z_min_cut = c.z_min_cut; k.z_max_cut = c.z_max_cut;
  k.h = c.h; k.half_h2 = c.half_h2; k.k_theta = c.k_theta; k.g = c.g; k.c_d = c.c_d;
  k.dz_train = c.dz_train; k.dz_sim = c.dz_sim; k.z_init = c.z_init; k.z_touch = c.z_touch;
  k.half_platform = c.half_platform; k.p_max_f = c.p_max_f; k.two_p_max_f = c.two_p_max_f; k.sigma_x = c.sigma_x;
  k.gamma = c.gamma;
  memcpy(k.transfer_ratio, c.transfer_ratio, sizeof(k.transfer_ratio));
  k.timeout_steps = c.timeout_steps; k.success_steps = c.success_steps; k.n_sub = c.n_sub;
  k.transfer_mode = c.transfer_mode; k.window_len = c.window_len; k.promote_successes = c.promote_successes;
  k.curriculum_steps = c.curriculum_steps; k.envs_per_population = c.envs_per_population;
  k.n_populations = c.n_populations; k.max_num_episodes = c.max_num_episodes;
  k.replicas = c.replicas_per_population;
}

int dqlb200_create(const dqlb200_config* cfg, const float* alpha_luts, const dqlb200_population_params* pop_params,
                   int device, dqlb200_handle** out) {
  if (!cfg || !alpha_luts || !pop_params || !out) return fail(DQLB200_ERR_ARG, "null argument");
  if (cfg->struct_bytes != sizeof(dqlb200_config) || cfg->abi_version != DQLB200_ABI_VERSION)
    return fail(DQLB200_ERR_ARG, "dqlb200_config size/ABI mismatch: host " + std::to_string(cfg->struct_bytes) +
                                     " vs library " + std::to_string(sizeof(dqlb200_config)));
  if (cfg->n_populations < 1 || cfg->envs_per_population < 1) return fail(DQLB200_ERR_ARG, "empty layout");
  if (cfg->curriculum_steps < 1 || cfg->curriculum_steps > DQLB200_MAX_CURRICULUM) return fail(DQLB200_ERR_ARG, "curriculum_steps out of range");
  const int tpb = cfg->threads_per_block;
  if (tpb != 32 && tpb != 64 && tpb != 128 && tpb != 256) return fail(DQLB200_ERR_ARG, "threads_per_block must be 32, 64, 128 or 256");
  if (cfg->window_len < 1 || cfg->window_len > DQLB200_MAX_WINDOW) return fail(DQLB200_ERR_ARG, "window_len out of range");
  if (cfg->n_alpha_luts < 1 || cfg->n_sub < 1) return fail(DQLB200_ERR_ARG, "n_alpha_luts / n_sub must be >= 1");
  if (cfg->replicas_per_population < 1 || cfg->n_populations % cfg->replicas_per_population)
    return fail(DQLB200_ERR_ARG, "n_populations must be a multiple of replicas_per_population (>= 1)");
  for (int p = 0; p < cfg->n_populations; ++p)
    if (pop_params[p].alpha_lut < 0 || pop_params[p].alpha_lut >= cfg->n_alpha_luts) return fail(DQLB200_ERR_ARG, "population alpha_lut index out of range");
  int count = 0;
  CUDA_TRY(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(DQLB200_ERR_ARG, "no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));
  dqlb200_handle* h = new (std::nothrow) dqlb200_handle();
  if (!h) return fail(DQLB200_ERR_STATE, "out of host memory");
  h->cfg = *cfg;
  fill_kc(*cfg, h->kc);
  h->device = device;
  h->env_state = h->tables = h->pop_state = h->merge_snapshot = nullptr;
  h->chunk_ready = false;
  h->merged_ready = false;
  h->merged_exec = nullptr;
  h->merged_k = 0;
  h->merged_promote = 0;
  CUDA_TRY(cudaMalloc(&h->d_cfg, sizeof(dqlb200_config)));
  CUDA_TRY(cudaMemcpy(h->d_cfg, cfg, sizeof(dqlb200_config), cudaMemcpyHostToDevice));
  const size_t lut_bytes = (size_t)cfg->n_alpha_luts * DQLB200_ALPHA_LUT * sizeof(float);
  CUDA_TRY(cudaMalloc(&h->d_alpha, lut_bytes));
  CUDA_TRY(cudaMemcpy(h->d_alpha, alpha_luts, lut_bytes, cudaMemcpyHostToDevice));
  const size_t pp_bytes = (size_t)cfg->n_populations * sizeof(dqlb200_population_params);
  CUDA_TRY(cudaMalloc(&h->d_pop_params, pp_bytes));
  CUDA_TRY(cudaMemcpy(h->d_pop_params, pop_params, pp_bytes, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&h->d_error, sizeof(uint32_t)));
  CUDA_TRY(cudaMemset(h->d_error, 0, sizeof(uint32_t)));
  {
    const int tpb_ = cfg->threads_per_block;
    const int n_slots = (cfg->envs_per_population + tpb_ - 1) / tpb_;
    if (n_slots > 2047) return fail(DQLB200_ERR_ARG, "envs_per_population too large for threads_per_block (max 2047 slots per thread)");
    h->smem_bytes = ((sizeof(dql::Shared) + 15) & ~size_t(15)) + (size_t)(tpb_ / 32) * dql::RESET_QUEUE * sizeof(uint16_t) +
                    (size_t)3 * tpb_ * 16;          // + the cp.async staging slots of the env prefetch
    if (h->smem_bytes > 227 * 1024) return fail(DQLB200_ERR_ARG, "population does not fit in shared memory: lower envs_per_population");
  }
#define DQL_SET_SMEM1(W, T, D)                                                                                            \
  CUDA_TRY(cudaFuncSetAttribute(dql::train_kernel<W, T, D>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
  CUDA_TRY(cudaFuncSetAttribute(dql::train_kernel<W, T, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bytes));
#define DQL_SET_SMEM(W) DQL_SET_SMEM1(W, false, false) DQL_SET_SMEM1(W, false, true) DQL_SET_SMEM1(W, true, true)
  DQL_SET_SMEM(1) DQL_SET_SMEM(2) DQL_SET_SMEM(4) DQL_SET_SMEM(8)
#undef DQL_SET_SMEM
#undef DQL_SET_SMEM1
  *out = h;
  return DQLB200_OK;
}

int dqlb200_destroy(dqlb200_handle* h) {
  if (!h) return DQLB200_OK;
  cudaSetDevice(h->device);
  cudaFree(h->d_cfg);
  cudaFree(h->d_alpha);
  cudaFree(h->d_pop_params);
  cudaFree(h->d_error);
  if (h->chunk_ready) {
    for (int c = 0; c < dqlb200_handle::MAX_HOST_CHUNKS; ++c) {
      cudaStreamDestroy(h->chunk_stream[c]);
      cudaEventDestroy(h->chunk_done[c]);
    }
    cudaEventDestroy(h->host_start);
  }
  if (h->merged_exec) cudaGraphExecDestroy(h->merged_exec);
  if (h->merged_ready) {
    cudaStreamDestroy(h->merged_stream);
    cudaEventDestroy(h->merged_in);
    cudaEventDestroy(h->merged_out);
  }
  delete h;
  return DQLB200_OK;
}

int dqlb200_bind(dqlb200_handle* h, void* env_state, void* tables, void* pop_state) {
  if (!h || !env_state || !tables || !pop_state) return fail(DQLB200_ERR_ARG, "null argument");
  if (((uintptr_t)env_state & 15u) || ((uintptr_t)tables & 3u) || ((uintptr_t)pop_state & 7u))
    return fail(DQLB200_ERR_ARG, "misaligned buffer (env_state needs 16 B, pop_state 8 B)");
  h->env_state = env_state;
  h->tables = tables;
  h->pop_state = pop_state;
  h->merged_k = 0;                 // a captured graph holds the old pointers
  return DQLB200_OK;
}

static dql::EnvPtrs env_ptrs(const dqlb200_handle* h, void* base) {
  const size_t n = (size_t)h->cfg.n_populations * h->cfg.envs_per_population;
  dql::EnvPtrs p;
  p.a = reinterpret_cast<float4*>(base);
  p.b = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(base) + 16 * n);
  p.c = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(base) + 32 * n);
  return p;
}

int dqlb200_reset(dqlb200_handle* h, int initial_step, void* stream) {
  if (!h || !h->env_state) return fail(DQLB200_ERR_STATE, "buffers not bound");
  if (initial_step < 0 || initial_step >= h->cfg.curriculum_steps) return fail(DQLB200_ERR_ARG, "initial_step out of range");
  CUDA_TRY(cudaSetDevice(h->device));
  const dim3 grid((h->cfg.envs_per_population + 255) / 256, h->cfg.n_populations);
  dql::reset_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h->kc, env_ptrs(h, h->env_state),
                                                           (dqlb200_population_state*)h->pop_state, h->d_pop_params, initial_step);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

static int launch_train(dqlb200_handle* h, int k_steps, const dqlb200_trace* trace, void* env_state, void* tables,
                        void* pop_state, cudaStream_t stream, int pop_offset = 0, int pop_count = -1) {
  dql::TrainArgs a;
  a.env = env_ptrs(h, env_state);
  a.tables = (uint32_t*)tables;
  a.pop_state = (dqlb200_population_state*)pop_state;
  a.pop_params = h->d_pop_params;
  a.alpha_luts = h->d_alpha;
  a.eps_threshold = h->d_cfg->eps_threshold;
  if (trace) a.trace = *trace; else memset(&a.trace, 0, sizeof(a.trace));
  a.merge_snapshot = (h->cfg.replicas_per_population > 1) ? (uint32_t*)h->merge_snapshot : nullptr;
  a.k_steps = k_steps;
  a.pop_offset = pop_offset;
  a.n_total = (long long)h->cfg.n_populations * h->cfg.envs_per_population;
  const int grid = pop_count < 0 ? h->cfg.n_populations : pop_count;
  const size_t smem = h->smem_bytes;
  const bool tracing = trace != nullptr;
#define DQL_LAUNCH(W)                                                                        \
  if (tracing) dql::train_kernel<W, true, true><<<grid, W * 32, smem, stream>>>(h->kc, a);                     \
  else if (h->kc.div_two_steps) dql::train_kernel<W, false, true><<<grid, W * 32, smem, stream>>>(h->kc, a);    \
  else dql::train_kernel<W, false, false><<<grid, W * 32, smem, stream>>>(h->kc, a);
  switch (h->cfg.threads_per_block) {
    case 32: DQL_LAUNCH(1) break;
    case 64: DQL_LAUNCH(2) break;
    case 128: DQL_LAUNCH(4) break;
    default: DQL_LAUNCH(8) break;
  }
#undef DQL_LAUNCH
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_train(dqlb200_handle* h, int k_steps, const dqlb200_trace* trace, void* stream) {
  if (!h || !h->env_state) return fail(DQLB200_ERR_STATE, "buffers not bound");
  if (k_steps < 0) return fail(DQLB200_ERR_ARG, "k_steps < 0");
  if (k_steps == 0) return DQLB200_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  return launch_train(h, k_steps, trace, h->env_state, h->tables, h->pop_state, (cudaStream_t)stream);
}

int dqlb200_train_host(dqlb200_handle* h, int k_steps, void* env_state_host, void* tables_host, void* pop_state_host,
                       void* stream) {
  if (!h || !h->env_state) return fail(DQLB200_ERR_STATE, "buffers not bound (device staging buffers are the bound ones)");
  if (!env_state_host || !tables_host || !pop_state_host) return fail(DQLB200_ERR_ARG, "null host buffer");
  if (k_steps < 0) return fail(DQLB200_ERR_ARG, "k_steps < 0");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->chunk_ready) {
    for (int c = 0; c < dqlb200_handle::MAX_HOST_CHUNKS; ++c) {
      CUDA_TRY(cudaStreamCreateWithFlags(&h->chunk_stream[c], cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreateWithFlags(&h->chunk_done[c], cudaEventDisableTiming));
    }
    CUDA_TRY(cudaEventCreateWithFlags(&h->host_start, cudaEventDisableTiming));
    h->chunk_ready = true;
  }
  // Populations are independent within a launch, so the call is pipelined over chunks of populations: while chunk c
  // trains, chunk c+1 is copied in and chunk c-1 is copied out (PCIe is full duplex, the copy engines run beside the SMs).
  const int P = h->cfg.n_populations, n_p = h->cfg.envs_per_population;
  const int n_chunks = P < dqlb200_handle::MAX_HOST_CHUNKS ? P : dqlb200_handle::MAX_HOST_CHUNKS;
  const size_t n = (size_t)P * n_p;
  const size_t tab_stride = (size_t)3 * DQLB200_MAX_CELLS * 4, ps_stride = sizeof(dqlb200_population_state);
  CUDA_TRY(cudaEventRecord(h->host_start, s));
  for (int c = 0; c < n_chunks; ++c) {
    const int p0 = (int)((long long)P * c / n_chunks), p1 = (int)((long long)P * (c + 1) / n_chunks);
    if (p1 == p0) continue;
    cudaStream_t cs = h->chunk_stream[c];
    CUDA_TRY(cudaStreamWaitEvent(cs, h->host_start, 0));
    const size_t e0 = (size_t)p0 * n_p * 16, eb = (size_t)(p1 - p0) * n_p * 16;
    for (int a = 0; a < 3; ++a)       // the three 16-byte vectors of the env-state SoA
      CUDA_TRY(cudaMemcpyAsync((char*)h->env_state + a * 16 * n + e0, (const char*)env_state_host + a * 16 * n + e0, eb, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(cudaMemcpyAsync((char*)h->tables + p0 * tab_stride, (const char*)tables_host + p0 * tab_stride, (p1 - p0) * tab_stride, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(cudaMemcpyAsync((char*)h->pop_state + p0 * ps_stride, (const char*)pop_state_host + p0 * ps_stride, (p1 - p0) * ps_stride, cudaMemcpyHostToDevice, cs));
    if (k_steps > 0) {
      const int rc = launch_train(h, k_steps, nullptr, h->env_state, h->tables, h->pop_state, cs, p0, p1 - p0);
      if (rc) return rc;
    }
    for (int a = 0; a < 3; ++a)
      CUDA_TRY(cudaMemcpyAsync((char*)env_state_host + a * 16 * n + e0, (const char*)h->env_state + a * 16 * n + e0, eb, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(cudaMemcpyAsync((char*)tables_host + p0 * tab_stride, (const char*)h->tables + p0 * tab_stride, (p1 - p0) * tab_stride, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(cudaMemcpyAsync((char*)pop_state_host + p0 * ps_stride, (const char*)h->pop_state + p0 * ps_stride, (p1 - p0) * ps_stride, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(cudaEventRecord(h->chunk_done[c], cs));
    CUDA_TRY(cudaStreamWaitEvent(s, h->chunk_done[c], 0));
  }
  CUDA_TRY(cudaStreamSynchronize(s));
  return DQLB200_OK;
}

int dqlb200_eval_greedy(dqlb200_handle* h, int population, const uint8_t* policy, int64_t first_episode, int64_t n_episodes,
                        int working_step, void* stats_out, const dqlb200_trace* trace, int trace_steps, void* stream) {
  if (!h || !policy || !stats_out) return fail(DQLB200_ERR_ARG, "null argument");
  if (population < 0 || population >= h->cfg.n_populations) return fail(DQLB200_ERR_ARG, "population out of range");
  if (working_step < 0 || working_step >= DQLB200_MAX_CURRICULUM) return fail(DQLB200_ERR_ARG, "working_step out of range");
  if (n_episodes <= 0) return DQLB200_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  dqlb200_trace tr;
  if (trace) tr = *trace; else memset(&tr, 0, sizeof(tr));
  const long long blocks = (n_episodes + 255) / 256;
  dql::eval_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(h->kc, h->d_pop_params, population, policy, first_episode,
                                                                     n_episodes, working_step, (dqlb200_eval_stats*)stats_out,
                                                                     tr, trace ? trace_steps : 0);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_env_reset(dqlb200_handle* h, int working_step, uint32_t birth, const uint8_t* mask, int fresh_mdp, int simulation,
                      uint16_t* out_state, void* stream) {
  if (!h || !h->env_state) return fail(DQLB200_ERR_STATE, "buffers not bound");
  if (working_step < 0 || working_step >= DQLB200_MAX_CURRICULUM) return fail(DQLB200_ERR_ARG, "working_step out of range");
  CUDA_TRY(cudaSetDevice(h->device));
  const long long n = (long long)h->cfg.n_populations * h->cfg.envs_per_population;
  dql::env_reset_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->kc, env_ptrs(h, h->env_state), h->d_pop_params, working_step,
                                                                                     birth, mask, fresh_mdp, simulation, out_state);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_env_step(dqlb200_handle* h, int working_step, uint32_t t, const int8_t* actions, int auto_reset, int simulation,
                     uint16_t* out_state, double* out_reward, uint8_t* out_code, uint8_t* out_done, float* out_obs, uint32_t* out_steps,
                     double* out_cumulative, void* stream) {
  if (!h || !h->env_state) return fail(DQLB200_ERR_STATE, "buffers not bound");
  if (!actions) return fail(DQLB200_ERR_ARG, "actions required");
  if (working_step < 0 || working_step >= DQLB200_MAX_CURRICULUM) return fail(DQLB200_ERR_ARG, "working_step out of range");
  CUDA_TRY(cudaSetDevice(h->device));
  const long long n = (long long)h->cfg.n_populations * h->cfg.envs_per_population;
  const unsigned blocks = (unsigned)((n + 127) / 128);
  if (h->kc.div_two_steps)
    dql::env_step_kernel<true><<<blocks, 128, 0, (cudaStream_t)stream>>>(h->kc, env_ptrs(h, h->env_state), h->d_pop_params, working_step, t, actions, auto_reset,
                                                                       simulation, out_state, out_reward, out_code, out_done, out_obs, out_steps, out_cumulative, h->d_error);
  else
    dql::env_step_kernel<false><<<blocks, 128, 0, (cudaStream_t)stream>>>(h->kc, env_ptrs(h, h->env_state), h->d_pop_params, working_step, t, actions, auto_reset,
                                                                        simulation, out_state, out_reward, out_code, out_done, out_obs, out_steps, out_cumulative, h->d_error);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_eval_greedy_2d(dqlb200_handle* h, const dqlb200_eval2d_params* p, const uint8_t* policy_x, const uint8_t* policy_y,
                           int64_t first_episode, int64_t n_episodes, void* stats_out, const dqlb200_trace2d* trace, int trace_steps,
                           void* stream) {
  if (!h || !p || !policy_x || !policy_y || !stats_out) return fail(DQLB200_ERR_ARG, "null argument");
  if (p->trajectory < 0 || p->trajectory > 2) return fail(DQLB200_ERR_ARG, "trajectory must be 0 (rectilinear x), 1 (rectilinear x and y) or 2 (eight)");
  if (p->working_step < 0 || p->working_step >= DQLB200_MAX_CURRICULUM) return fail(DQLB200_ERR_ARG, "working_step out of range");
  if (n_episodes <= 0) return DQLB200_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  dqlb200_trace2d tr;
  if (trace) tr = *trace; else memset(&tr, 0, sizeof(tr));
  const long long blocks = (n_episodes + 255) / 256;
  dql::eval2d_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(h->kc, *p, policy_x, policy_y, first_episode, n_episodes,
                                                                       (dqlb200_eval_stats*)stats_out, tr, trace ? trace_steps : 0);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_transfer(dqlb200_handle* h, int step, float ratio, void* stream) {
  if (!h || !h->tables) return fail(DQLB200_ERR_STATE, "buffers not bound");
  if (step < 0 || step >= h->cfg.curriculum_steps) return fail(DQLB200_ERR_ARG, "step out of range");
  CUDA_TRY(cudaSetDevice(h->device));
  dql::transfer_kernel<<<h->cfg.n_populations, 256, 0, (cudaStream_t)stream>>>((uint32_t*)h->tables, h->cfg.n_populations,
                                                                             h->cfg.curriculum_steps, step, ratio);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_check_errors(dqlb200_handle* h, void* stream) {
  if (!h) return fail(DQLB200_ERR_ARG, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  uint32_t facade = 0;
  CUDA_TRY(cudaMemcpy(&facade, h->d_error, sizeof(facade), cudaMemcpyDeviceToHost));
  if (facade) {
    CUDA_TRY(cudaMemset(h->d_error, 0, sizeof(uint32_t)));
    return fail(DQLB200_ERR_DEVICE_FLAG, facade & 1u ? "Unexpected discretization case: NaN observation"
                                                     : (facade & 2u ? "Cannot check an empty state" : "Previous state missing"));
  }
  for (int p = 0; h->pop_state && p < h->cfg.n_populations; ++p) {
    dqlb200_population_state ps;
    CUDA_TRY(cudaMemcpy(&ps, (dqlb200_population_state*)h->pop_state + p, sizeof(ps), cudaMemcpyDeviceToHost));
    if (ps.error_flags) return fail(DQLB200_ERR_DEVICE_FLAG, "population " + std::to_string(p) + ": NaN observation (error_flags=" + std::to_string(ps.error_flags) + ")");
  }
  return DQLB200_OK;
}

int dqlb200_shared_pack(dqlb200_handle* h, const void* snapshot, void* delta, void* stream) {
  if (!h || !h->tables || !h->pop_state || !snapshot || !delta) return fail(DQLB200_ERR_ARG, "null argument / not bound");
  CUDA_TRY(cudaSetDevice(h->device));
  const int R = h->cfg.replicas_per_population, n_agents = h->cfg.n_populations / R;
  const long long n = (long long)n_agents * DQLB200_MAX_CELLS;
  dql::shared_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)h->tables, (const uint32_t*)snapshot,
                                                                                       (float*)delta, (const dqlb200_population_state*)h->pop_state,
                                                                                       n_agents, R);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_shared_apply(dqlb200_handle* h, void* snapshot, const void* delta_reduced, int pooled_promote_successes, void* stream) {
  if (!h || !h->tables || !h->pop_state || !snapshot || !delta_reduced) return fail(DQLB200_ERR_ARG, "null argument / not bound");
  CUDA_TRY(cudaSetDevice(h->device));
  const int R = h->cfg.replicas_per_population, n_agents = h->cfg.n_populations / R;
  if (R > 1 && !h->merge_snapshot) return fail(DQLB200_ERR_STATE, "replicated layout: bind the merge snapshot first (dqlb200_bind_merge_snapshot)");
  const long long n = (long long)n_agents * DQLB200_MAX_CELLS;
  dql::shared_apply_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((uint32_t*)h->tables, (uint32_t*)snapshot,
                                                                                        R > 1 ? (uint32_t*)h->merge_snapshot : nullptr,
                                                                                        (const float*)delta_reduced, (dqlb200_population_state*)h->pop_state,
                                                                                        n_agents, R, pooled_promote_successes, h->cfg.max_num_episodes);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_mdp_facade_step(dqlb200_handle* h, int working_step, int ops, int64_t n, const double* obs, const uint8_t* contact,
                            const int8_t* action, double* mdp_state, uint16_t* out_state, uint8_t* out_code, double* out_reward,
                            void* stream) {
  if (!h || !mdp_state) return fail(DQLB200_ERR_ARG, "null argument");
  if (working_step < 0 || working_step >= DQLB200_MAX_CURRICULUM) return fail(DQLB200_ERR_ARG, "working_step out of range");
  if ((ops & (DQLB200_OP_OBSERVE | DQLB200_OP_CHECK)) && !obs) return fail(DQLB200_ERR_ARG, "obs required");
  if ((ops & DQLB200_OP_CHECK) && !contact) return fail(DQLB200_ERR_ARG, "contact required");
  if ((ops & DQLB200_OP_ACTION) && !action) return fail(DQLB200_ERR_ARG, "action required");
  if (n <= 0) return DQLB200_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  dql::facade_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(h->d_cfg, working_step, ops, n, obs, contact, action,
                                                                                   mdp_state, out_state, out_code, out_reward, h->d_error);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_bind_merge_snapshot(dqlb200_handle* h, void* snapshot) {
  if (!h) return fail(DQLB200_ERR_ARG, "null handle");
  if ((uintptr_t)snapshot & 3u) return fail(DQLB200_ERR_ARG, "misaligned snapshot");
  h->merge_snapshot = snapshot;
  h->merged_k = 0;
  return DQLB200_OK;
}

static int launch_merge(dqlb200_handle* h, void* snapshot, int pooled_promote_successes, cudaStream_t stream) {
  const int R = h->cfg.replicas_per_population;
  const dim3 grid((DQLB200_MAX_CELLS + 31) / 32, h->cfg.n_populations / R);     // one CTA per tile of 32 cells
  dql::replica_merge_kernel<<<grid, 256, 0, stream>>>((uint32_t*)h->tables, (uint32_t*)snapshot, (dqlb200_population_state*)h->pop_state, R,
                                                     pooled_promote_successes, h->cfg.max_num_episodes);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_replica_merge(dqlb200_handle* h, void* snapshot, int pooled_promote_successes, void* stream) {
  if (!h || !h->tables || !h->pop_state || !snapshot) return fail(DQLB200_ERR_ARG, "null argument / not bound");
  if (h->cfg.replicas_per_population > 1 && snapshot != h->merge_snapshot)
    return fail(DQLB200_ERR_STATE, "snapshot is not the one bound with dqlb200_bind_merge_snapshot (train launches keep it current across transfers)");
  const int R = h->cfg.replicas_per_population;
  if (R < 1 || h->cfg.n_populations % R) return fail(DQLB200_ERR_STATE, "n_populations is not a multiple of replicas_per_population");
  CUDA_TRY(cudaSetDevice(h->device));
  return launch_merge(h, snapshot, pooled_promote_successes, (cudaStream_t)stream);
}

int dqlb200_train_merged(dqlb200_handle* h, int total_steps, int merge_every, int pooled_promote_successes, void* stream) {
  if (!h || !h->env_state || !h->tables || !h->pop_state) return fail(DQLB200_ERR_STATE, "buffers not bound");
  if (!h->merge_snapshot) return fail(DQLB200_ERR_STATE, "bind the merge snapshot first (dqlb200_bind_merge_snapshot)");
  if (total_steps < 0 || merge_every < 1) return fail(DQLB200_ERR_ARG, "total_steps >= 0 and merge_every >= 1 required");
  if (total_steps == 0) return DQLB200_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  if (!h->merged_ready) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->merged_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&h->merged_in, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&h->merged_out, cudaEventDisableTiming));
    h->merged_ready = true;
  }
  cudaStream_t ms = h->merged_stream, cs = (cudaStream_t)stream;
  const int n_full = total_steps / merge_every, rem = total_steps % merge_every;
  if (n_full > 0 && (h->merged_k != merge_every || h->merged_promote != pooled_promote_successes || !h->merged_exec)) {
    // (train launch of merge_every steps, replica merge) captured once and replayed: the pair is launch-bound for
    // populations of a few thousand CTAs (two ~5 us launches around ~10 us of work), a graph launch is one submission
    if (h->merged_exec) { cudaGraphExecDestroy(h->merged_exec); h->merged_exec = nullptr; }
    cudaGraph_t graph = nullptr;
    CUDA_TRY(cudaStreamBeginCapture(ms, cudaStreamCaptureModeThreadLocal));
    int rc = launch_train(h, merge_every, nullptr, h->env_state, h->tables, h->pop_state, ms);
    if (!rc) rc = launch_merge(h, h->merge_snapshot, pooled_promote_successes, ms);
    const cudaError_t ce = cudaStreamEndCapture(ms, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    CUDA_TRY(ce);
    CUDA_TRY(cudaGraphInstantiate(&h->merged_exec, graph, 0));
    CUDA_TRY(cudaGraphDestroy(graph));
    h->merged_k = merge_every;
    h->merged_promote = pooled_promote_successes;
  }
  CUDA_TRY(cudaEventRecord(h->merged_in, cs));
  CUDA_TRY(cudaStreamWaitEvent(ms, h->merged_in, 0));
  for (int i = 0; i < n_full; ++i) CUDA_TRY(cudaGraphLaunch(h->merged_exec, ms));
  if (rem) {
    int rc = launch_train(h, rem, nullptr, h->env_state, h->tables, h->pop_state, ms);
    if (!rc) rc = launch_merge(h, h->merge_snapshot, pooled_promote_successes, ms);
    if (rc) return rc;
  }
  CUDA_TRY(cudaEventRecord(h->merged_out, ms));
  CUDA_TRY(cudaStreamWaitEvent(cs, h->merged_out, 0));
  return DQLB200_OK;
}

int dqlb200_bench_table_rmw(dqlb200_handle* h, const uint16_t* cells, int64_t n_cells, int visits_per_thread, int threads, int blocks,
                            void* checksum_out, void* stream) {
  if (!h || !cells || !checksum_out) return fail(DQLB200_ERR_ARG, "null argument");
  if (n_cells < 1 || visits_per_thread < 1 || threads < 32 || threads > 1024 || blocks < 1) return fail(DQLB200_ERR_ARG, "bad launch shape");
  CUDA_TRY(cudaSetDevice(h->device));
  dql::table_rmw_roof_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(cells, n_cells, visits_per_thread, (unsigned long long*)checksum_out);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

int dqlb200_selftest_division(dqlb200_handle* h, uint64_t* mismatches_out, void* stream) {
  if (!h || !mismatches_out) return fail(DQLB200_ERR_ARG, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  unsigned long long* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, 3 * sizeof(unsigned long long)));
  CUDA_TRY(cudaMemsetAsync(d, 0, 3 * sizeof(unsigned long long), (cudaStream_t)stream));
  dql::selftest_division_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(h->kc, d);
  CUDA_TRY(cudaGetLastError());
  unsigned long long v[3] = {0, 0, 0};
  CUDA_TRY(cudaMemcpyAsync(v, d, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  CUDA_TRY(cudaFree(d));
  mismatches_out[0] = v[0];
  mismatches_out[1] = v[1];
  mismatches_out[2] = v[2];
  return DQLB200_OK;
}

int dqlb200_agent_facade(dqlb200_handle* h, int op, int64_t n, double* tables_f64, const int32_t* state, const int32_t* action,
                         const int32_t* next_state, const double* alpha, const double* reward, double gamma, int32_t* out_action,
                         void* stream) {
  if (!h || !tables_f64 || !state) return fail(DQLB200_ERR_ARG, "null argument");
  if (op == DQLB200_AGENT_PREDICT && !out_action) return fail(DQLB200_ERR_ARG, "out_action required");
  if (op == DQLB200_AGENT_UPDATE && (!action || !next_state || !alpha || !reward)) return fail(DQLB200_ERR_ARG, "update operands required");
  if (op == DQLB200_AGENT_TRANSFER && !alpha) return fail(DQLB200_ERR_ARG, "ratio required");
  if (op != DQLB200_AGENT_PREDICT && op != DQLB200_AGENT_UPDATE && op != DQLB200_AGENT_TRANSFER) return fail(DQLB200_ERR_ARG, "unknown op");
  if (n <= 0) return DQLB200_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  dql::agent_facade_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(op, n, tables_f64, h->cfg.curriculum_steps, state, action, next_state,
                                                            alpha, reward, gamma, out_action);
  CUDA_TRY(cudaGetLastError());
  return DQLB200_OK;
}

}  // extern "C"
