// train_kernel.cuh -- the fused training kernel (one CTA per population): phase A per env, ordered commit, auto-reset, curriculum
// Part of libdqlb200 (see dqlb200.cu for the kernel inventory and the C-ABI).
#pragma once
#include <type_traits>

#include "env_state.cuh"

namespace dql {

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

// Probe builds only (-DDQL_TIMING, tools/perf_probe_timeline.py): %globaltimer stamps of every CTA at a few points of a launch and
// the SM / hardware warp slot it runs in, fetched with dqlb200_debug_timing.  Compiled out of the product library.
#ifdef DQL_TIMING
__device__ unsigned long long dql_timing[4096 * 8];
#define DQL_STAMP(i)                                                                                                      \
  do {                                                                                                                    \
    if (threadIdx.x == 0 && blockIdx.x < 2048) {                                                                          \
      unsigned long long t_;                                                                                              \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)::"memory");                                                    \
      dql_timing[blockIdx.x * 8 + (i)] = t_;                                                                              \
    }                                                                                                                     \
  } while (0)
#else
#define DQL_STAMP(i) do { } while (0)
#endif

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

constexpr int RESET_QUEUE = 64;    // finished envs a warp collects before it runs the batched reset pass (flushed early beyond 32)

constexpr int STATES = DQLB200_MAX_CURRICULUM * DQLB200_STATES_PER_LEVEL;   // 945

// Finished episodes of a global step, logged by the ordered commit and folded into the trainer state at the end of the step:
// the success flags of the finished episodes in commit (= env) order, one byte each, and per warp-slot the fixed-tree float64 sum
// of its finished episodes' returns (added to the running sum one after the other, in commit order).
constexpr int EP_OK_CAP = 128;     // success flags buffered before a drain (a warp-slot appends at most 32)
constexpr int EP_RET_CAP = 32;     // warp-slots with finished episodes buffered before a drain

// Shared memory of one population (CTA).  Q_b and the alpha LUT stay in global memory (read-only in the step
// loop, L1-resident): that keeps the footprint at ~37 KB so that six CTAs fit on one SM.
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
  uint32_t step_episodes, step_success, step_ep_steps, step_hist[9];   // counters of the current global step (native 32-bit shared atomics), folded into the 64-bit totals of `ps` once per step
  int promote, advance, do_advance;
  uint4 philox_keys[5];   // round keys of the population's Philox key (philox_round_keys)
  // Finished episodes of the warp-slots committed so far (see EP_OK_CAP): the serialised section only appends; success window,
  // promotion test and logged sums are brought up to date at the end of the global step (or when a buffer is full), off the
  // critical path of the baton.
  double ep_ret[EP_RET_CAP];
  uint8_t ep_ok[EP_OK_CAP];
  int n_ep_ok, n_ep_ret;
  // followed by (dynamic): uint16_t reset_queue[WARPS][RESET_QUEUE]; uint4 stage[WARPS][3][32] (+ [3][NT]); uint64_t tile_mbar[WARPS] (one mbarrier per warp: completion of the bulk copy of its next env tile); SpEntry sp_tab[n_setpoints][3]
};

// dynamic shared memory of a launch: Shared + the set-point table + the reset queues + the cp.async staging slots of the env
// (and, for the extended / trace instances, extension-state) prefetch
// The set-point tables of the configuration in shared memory, one entry per (set-point index, action): the integrator's next index
// and its float32 value (R3), and the set-point term of the reward for an ordinary step and for the first step of an episode
// (dqlb200_config.setpoint_rtheta[0 / 1]).  In global memory the reward term was an L2 round trip per env-step: the L1 next to
// 228 KB of shared memory does not hold the tables (ncu: 26 % hit rate).
struct SpEntry {
  uint32_t next;
  float value;
  double r_step, r_first;
};
__host__ __device__ constexpr size_t train_sp_bytes(int n_setpoints) { return ((size_t)n_setpoints * 3 * sizeof(SpEntry) + 15) & ~size_t(15); }
__host__ __device__ constexpr size_t train_smem_bytes(int threads, bool extended, int n_setpoints) {
  return ((sizeof(Shared) + 15) & ~size_t(15)) + train_sp_bytes(n_setpoints) + (size_t)(threads / 32) * RESET_QUEUE * sizeof(uint16_t) +
         (size_t)(extended ? 6 : 3) * threads * 16 + (((size_t)(threads / 32) * 8 + 15) & ~size_t(15));
}

// the production launch shape (128 threads, 33 reachable set-points of the reference defaults) must keep six populations resident
// per SM: 228 KB of shared memory per SM, 1 KB reserved per CTA (a 96-byte overshoot once halved the occupancy unnoticed)
static_assert((train_smem_bytes(128, false, 33) + 1024) * 6 <= 228 * 1024, "train_kernel<4>: six CTAs per SM no longer fit in shared memory");
#ifndef DQL_WARPS_PER_SM
#define DQL_WARPS_PER_SM 24     // resident warps per SM the register allocation is tuned for (launch bounds)
#endif
#ifndef DQL_WARPS_PER_SM_GENERIC
#define DQL_WARPS_PER_SM_GENERIC 20     // the extended instance carries the estimator and the second-order model: 5 CTAs/SM, 102 registers
#endif
template <bool GENERIC_> struct ConstsOf;
template <> struct ConstsOf<true> {
  __device__ __forceinline__ static const KC& get(const KC& kc) { return kc; }
};
template <> struct ConstsOf<false> {
  __device__ __forceinline__ static KDef get(const KC&) { return KDef{}; }
};

// GENERIC = false is the production instance of the reference's default configuration; GENERIC = true adds what only
// non-default configurations need: run-time constants instead of the compile-time defaults (KDef), the second Markstein
// correction step of x / p_max, x / v_max (required unless the divisors are the exhaustively verified defaults) and the
// observation-noise option.  The trace instances are generic
// (both division variants are correctly rounded, hence identical).
// VARIANT 0 = production, 1 = generic, 2 = extended: generic plus the options that carry extra per-env state (acceleration
// estimator, second-order model) -- a separate instance so that the generic one does not pay for their branches and registers.
template <int WARPS, bool TRACE, int VARIANT>
__global__ void __launch_bounds__(WARPS * 32, ((VARIANT == 2 ? DQL_WARPS_PER_SM_GENERIC : DQL_WARPS_PER_SM) / WARPS) > 0 ? ((VARIANT == 2 ? DQL_WARPS_PER_SM_GENERIC : DQL_WARPS_PER_SM) / WARPS) : 1) train_kernel(const __grid_constant__ KC kc, const TrainArgs args) {
  // VARIANT 3 = production for populations that fill every slot (envs_per_population a multiple of the block size): `valid`
  // is a compile-time constant, which removes the predicate, the defaults of the invalid lanes and their reconvergence points
  constexpr bool GENERIC = VARIANT == 1 || VARIANT == 2, EXT = VARIANT == 2, FULL_SLOTS = VARIANT == 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
  DQL_STAMP(0);      // entry
#ifdef DQL_TIMING
  if (threadIdx.x == 0 && blockIdx.x < 2048) {      // where the CTA runs (written from the tail of the buffer)
    unsigned smid_, warpid_;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid_));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid_));
    dql_timing[4096 * 8 - 1 - blockIdx.x] = ((unsigned long long)smid_ << 32) | warpid_;
  }
#endif
  const auto& kk = ConstsOf<GENERIC>::get(kc);       // run-time KC (generic) or the compile-time defaults KDef (production)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NT = WARPS * 32;
  const int pop = blockIdx.x + args.pop_offset;
  const int n_p = kc.envs_per_population;
  const int n_slots = (n_p + NT - 1) / NT;
  // dynamic part: everything of run-time size (the set-point table) comes LAST, so that every base is a compile-time offset
  unsigned char* dyn = smem_raw + ((sizeof(Shared) + 15) & ~size_t(15));
  uint16_t* reset_queue = reinterpret_cast<uint16_t*>(dyn) + (size_t)warp * RESET_QUEUE;
  dyn += (size_t)WARPS * RESET_QUEUE * sizeof(uint16_t);
  uint4* stage = reinterpret_cast<uint4*>(dyn);      // [WARPS][3][32]: one staging tile per warp (+ [3][NT] extension-state slots in the extended / trace instances)
  dyn += (size_t)((TRACE || EXT) ? 6 : 3) * NT * 16;
  unsigned long long* tile_mbar = reinterpret_cast<unsigned long long*>(dyn);
  dyn += ((size_t)WARPS * 8 + 15) & ~size_t(15);
  SpEntry* sp_tab = reinterpret_cast<SpEntry*>(dyn);     // [n_setpoints][3]
  const unsigned stage_tile = (unsigned)__cvta_generic_to_shared(stage) + (unsigned)warp * (unsigned)ENV_TILE_BYTES;
  const unsigned stage_lane = stage_tile + (unsigned)lane * 16u;                                   // this lane's A vector in the staged tile
  const unsigned ext_stage = (unsigned)__cvta_generic_to_shared(stage + WARPS * 96 + tid);       // extension-state slots of this thread
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(tile_mbar + warp);
  const int tiles_pp = args.env.tiles_per_pop;
  uint32_t tile_phase = 0u;       // parity of the mbarrier phase the next wait is for
  bool tile_pending = false;      // a bulk copy is in flight (warp-uniform)
  // issue the bulk copy of `tile` into the warp's staging tile (every lane has read the staged tile: the warp converges first).
  // The copy engine reads through the async proxy what the lanes of this warp wrote through the generic proxy (env_store, resets):
  // every thread orders its writes of a global step with ONE proxy fence behind its last store of the step (env_writes_done),
  // and a warp / CTA barrier lies between that fence and every later prefetch (a fence per prefetch would wait for the stores of
  // the slot before it: an L2 round trip in every slot).
  auto prefetch_tile = [&](const unsigned char* tile) {
    __syncwarp();
    if (lane == 0) tile_prefetch_bulk(tile, stage_tile, mbar);
    tile_pending = true;
  };
  auto env_writes_done = [&]() { asm volatile("fence.proxy.async.global;" ::: "memory"); };
  auto wait_tile = [&]() {
    tile_wait(mbar, tile_phase);
    tile_phase ^= 1u;
    tile_pending = false;
  };
  const size_t env_base = (size_t)pop * n_p;
  uint32_t* gt = args.tables + (size_t)pop * 3 * CELLS;
  float* gqb = reinterpret_cast<float*>(gt + CELLS);      // table B, written only by the transfer below
  const uint32_t sp_zero = (uint32_t)args.env.sp_zero;

  // ---- stage population state and the LIVE rows of the tables in shared memory --------------------
  // At working step w only levels 0..w can be visited (a state's level never exceeds w), so only rows
  // [0, (w+1)*567) of Q_a / count are staged, snapshotted and written back; a promotion loads the next level.
  // The launch prologue is ONE round trip to memory: every load below is independent of the others (the working step
  // and the population constants are broadcast loads by every thread instead of a hop through shared memory), and the
  // env state of slot 0 -- a cold HBM read when one global step is run per launch -- is in flight during all of it.
  static_assert(sizeof(dqlb200_population_state) % 4 == 0, "word copies");
  constexpr int PS_WORDS = sizeof(dqlb200_population_state) / 4;
  const dqlb200_population_params pp = args.pop_params[pop];
  unsigned char* const p_tile0 = env_addr(args.env, pop, warp * 32);      // tile `warp` of the population (exists if warp < tiles_pp)
  if (lane == 0) tile_mbar_init(mbar);
  if (warp < tiles_pp) prefetch_tile(p_tile0);
  if (EXT) ext_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), ext_stage, NT, kc.accel_mode != 0, kc.dynamics_model != 0);
  {
    const int w_start = args.pop_state[pop].working_step;
    const uint32_t* gps = reinterpret_cast<const uint32_t*>(args.pop_state + pop);
    for (int i = tid; i < PS_WORDS; i += NT) reinterpret_cast<uint32_t*>(&sh.ps)[i] = gps[i];
    if (tid == 0) {
      sh.step_episodes = sh.step_success = sh.step_ep_steps = 0u;
      for (int i = 0; i < 9; ++i) sh.step_hist[i] = 0u;
      sh.promote = sh.advance = sh.do_advance = 0;
      sh.n_ep_ok = sh.n_ep_ret = 0;
      sh.cuts = kc.cuts[w_start];
    }
    if (tid < 5) sh.reward[tid] = kc.reward[tid];
    if (tid == 31) philox_round_keys(pp.seed_lo, pp.seed_hi, sh.philox_keys);
    for (int i = tid; i < args.env.n_sp * 3; i += NT) {
      const uint2 nx = args.env.sp_next[i];
      sp_tab[i] = SpEntry{nx.x, __uint_as_float(nx.y), args.env.sp_rtheta[i], args.env.sp_rtheta[DQLB200_MAX_SETPOINTS * 3 + i]};
    }
    // level 0 is live at every working step: its rows do not wait for the working step to arrive (a second, dependent round trip
    // to HBM when one global step is run per launch); the rows of levels 1..w follow
    const int live = (w_start + 1) * DQLB200_CELLS_PER_LEVEL;
    for (int i = tid; i < DQLB200_CELLS_PER_LEVEL; i += NT) {
      sh.qa[i] = __uint_as_float(gt[i]);
      sh.cnt[i] = gt[2 * CELLS + i];
      if ((i & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(gqb + i));     // table B rows for the first snapshot
    }
    for (int i = DQLB200_CELLS_PER_LEVEL + tid; i < live; i += NT) {
      sh.qa[i] = __uint_as_float(gt[i]);
      sh.cnt[i] = gt[2 * CELLS + i];
      if ((i & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(gqb + i));
    }
  }
  __syncthreads();

  DQL_STAMP(1);      // tables staged (first barrier passed)
  const float* __restrict__ alpha_lut = args.alpha_luts + (size_t)pp.alpha_lut * DQLB200_ALPHA_LUT;
  const float alpha_min = __ldg(alpha_lut + DQLB200_ALPHA_LUT - 1);      // alpha(count >= 1002), PKG/trainer.py:95-105
  const bool filt = EXT && kc.accel_mode != 0;      // acceleration estimator: 16 more bytes per env, extended instance only
  const bool so = EXT && kc.dynamics_model != 0;    // second-order attitude + vertical PID: 32 more bytes per env
  uint64_t steps_done = 0;

  // snapshot of a global step: greedy action (first max of (Q_a+Q_b)/2, PKG/double_q_learning.py:119-124) and bootstrap
  // value max_a Q_a (:136-141) of every live state, from the live table as the previous step left it
  auto build_snapshot = [&](int w) {
    for (int st = tid; st < (w + 1) * DQLB200_STATES_PER_LEVEL; st += NT) {
      const float q0 = sh.qa[st * 3 + 0], q1 = sh.qa[st * 3 + 1], q2 = sh.qa[st * 3 + 2];
      const float p0 = fmul(fadd(q0, __ldcg(gqb + st * 3 + 0)), 0.5f);
      const float p1 = fmul(fadd(q1, __ldcg(gqb + st * 3 + 1)), 0.5f);
      const float p2 = fmul(fadd(q2, __ldcg(gqb + st * 3 + 2)), 0.5f);
      int a = 0;
      float best = p0;
      if (p1 > best) { best = p1; a = 1; }
      if (p2 > best) { a = 2; }
      sh.greedy[st] = (uint8_t)a;
      sh.qmax[st] = fmaxf(fmaxf(q0, q1), q2);
    }
  };

  // R14 from the episode log: the finished episodes enter the success window in env order, the promotion test runs after every
  // append (PKG/trainer.py:219-236).  Called by ONE converged warp that has exclusive access to the trainer state: the baton
  // holder when the log is full, warp 0 between the two barriers at the end of a global step.
  auto drain_episode_log = [&]() {
    dqlb200_population_state& ps = sh.ps;
    const int n_fin = sh.n_ep_ok, n_ret = sh.n_ep_ret;
    if (n_fin == 0) return;      // warp-uniform
    const int L = kc.window_len;
    bool promote = false;
    if (lane == 31) {      // the float64 sum of the returns, one warp-slot after the other (a lane the window code below does not single out)
      double s = ps.return_sum;
      for (int i = 0; i < n_ret; ++i) s = __dadd_rn(s, sh.ep_ret[i]);
      ps.return_sum = s;
    }
    for (int base = 0; base < n_fin; base += 32) {
      const int n = min(32, n_fin - base);
      const bool mine = lane < n;
      const int ok = mine ? (int)sh.ep_ok[base + lane] : 0;
      const uint32_t smask = __ballot_sync(FULL, ok != 0);
      const int head = ps.window_head, count = ps.window_count, sum = ps.window_sum;      // broadcast reads
      if (n <= L) {
        // 32 appends at once: the j-th finished episode writes ring position head + j, evicts what was there once the window is
        // full, and sees the running sum of the appends up to and including its own
        int pos = head + lane;
        pos -= (pos >= L) ? L : 0;
        const bool evicts = mine && (count + lane >= L);
        const int old = evicts ? (int)ps.window[pos] : 0;
        const uint32_t emask = __ballot_sync(FULL, evicts && old != 0);
        const uint32_t upto = (2u << lane) - 1u;      // lanes 0..lane
        const int s_here = sum + __popc(smask & upto) - __popc(emask & upto);
        promote = promote || __any_sync(FULL, mine && s_here >= kc.promote_successes);
        __syncwarp();                                  // every eviction read precedes every write
        if (mine) ps.window[pos] = (uint8_t)ok;
        if (lane == 0) {
          int h2 = head + n;
          h2 -= (h2 >= L) ? L : 0;
          ps.window_head = h2;
          ps.window_count = min(count + n, L);
          ps.window_sum = sum + __popc(smask) - __popc(emask);
        }
      } else {      // a window shorter than the chunk: one by one
        bool pr = false;
        if (lane == 0) {
          int h2 = head, c2 = count, s2 = sum;
          for (int b = 0; b < n; ++b) {
            const int okb = (smask >> b) & 1u;
            if (c2 == L) s2 -= ps.window[h2];
            else c2 += 1;
            ps.window[h2] = (uint8_t)okb;
            s2 += okb;
            h2 = (h2 + 1 == L) ? 0 : h2 + 1;
            pr = pr || (s2 >= kc.promote_successes);
          }
          ps.window_head = h2; ps.window_count = c2; ps.window_sum = s2;
        }
        promote = promote || (__shfl_sync(FULL, (int)pr, 0) != 0);
      }
      __syncwarp();
    }
    if (lane == 0) {
      ps.episodes_in_step += n_fin;
      if (kc.replicas == 1) {      // replicas are promoted together by replica_merge_kernel
        if (promote) sh.promote = 1;
        if (ps.episodes_in_step >= kc.max_num_episodes) sh.advance = 1;
      }
      sh.n_ep_ok = sh.n_ep_ret = 0;
    }
    __syncwarp();
  };

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
          Kf kf;
          Ext ex;
          if (filt) kf = kf_load(args.env, env_base + env_i);       // the estimator outlives the curriculum step
          if (so) ex = ext_load(args.env, env_base + env_i);
          env_reset(kk, pp, sh.cuts, kk.angle_cut, e, (uint32_t)env_i, birth, w + 1, /*fresh_mdp=*/true, sp_zero, filt ? &kf : nullptr, so ? &ex : nullptr);
          env_store(env_addr(args.env, pop, env_i), e);
          if (filt) kf_store(args.env, env_base + env_i, kf);
          if (so) ext_store(args.env, env_base + env_i, ex);
        }
      }
      // every env was just restarted: the slot-0 prefetch in flight is stale
      env_writes_done();
      if (tile_pending) wait_tile();
      if (warp < tiles_pp) prefetch_tile(p_tile0);
      if (EXT) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        ext_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), ext_stage, NT, filt, so);
      }
      build_snapshot(w + 1);
    }
    __syncthreads();
  };
  // replica-merge mode: a promotion decided by replica_merge_kernel takes effect before the first step of this launch
  if (sh.ps.pending_advance && !sh.ps.finished) {
    if (tid == 0) sh.promote = (sh.ps.pending_advance == 1) ? 1 : 0;
    __syncthreads();
    advance_curriculum(sh.ps.working_step, sh.ps.t);
  } else {
    build_snapshot(sh.ps.working_step);
    __syncthreads();
  }

  DQL_STAMP(2);      // first snapshot built: the step loop starts
  for (int k = 0; k < args.k_steps; ++k) {
    if (sh.ps.finished) break;     // uniform: written only between barriers
    const int w = sh.ps.working_step;
    const uint32_t t = sh.ps.t;
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
          unsigned char* const pr = env_addr(args.env, pop, env_r);
          Env e;
          env_unpack(env_fetch(pr), e);
          Kf kf;
          Ext ex;
          if (filt) kf = kf_load(args.env, gr);
          if (so) ex = ext_load(args.env, gr);
          env_reset(kk, pp, sh.cuts, kk.angle_cut, e, (uint32_t)env_r, t + 1u, w, /*fresh_mdp=*/false, sp_zero, filt ? &kf : nullptr, so ? &ex : nullptr);
          env_store(pr, e);
          if (filt) kf_store(args.env, gr, kf);
          if (so) ext_store(args.env, gr, ex);
        }
      }
      __syncwarp();
      n_queued = 0;
    };

    // The slot loop, specialised at compile time for working step 0 (W0): there the only level is 0, so the level search and the
    // level tests vanish and the bin cuts / reward constants of level 0 are read from the kernel parameters (constant bank ->
    // uniform registers, loop-invariant) instead of shared memory; the Philox draw is unconditional.  Curriculum step 0 is
    // where training spends most of its time (and what the benchmark measures); only the production instances pay the code size.
    auto slot_loop = [&](auto w0_tag) {
    constexpr bool W0 = decltype(w0_tag)::value;
    bool bad_obs = false;
    unsigned char* p_env = env_addr(args.env, pop, tid);      // running pointer to the A vector of this thread's env: tile = slot * WARPS + warp
    for (int slot = 0; slot < n_slots; ++slot) {
      const int env_i = slot * NT + tid;
      const bool valid = FULL_SLOTS || env_i < n_p;
      const size_t gi = env_base + (size_t)env_i;          // only dereferenced under `valid`
      if (FULL_SLOTS || slot * WARPS + warp < tiles_pp) wait_tile();      // warp-uniform: this warp has a tile in this slot
      if (slot == 0) DQL_STAMP(7);      // the first tile has landed
      const EnvRaw cur_raw = tile_take(stage_lane);
      unsigned char* const p_cur = p_env;
      p_env += WARPS * ENV_TILE_BYTES;
      Kf kf;
      Ext ex;
      if (filt) kf = kf_take(ext_stage, NT);
      if (so) ex = ext_take(ext_stage, NT);
      // the next tile of this warp: in flight during this slot (every lane has read the staged tile: prefetch_tile converges first)
      if (FULL_SLOTS ? (slot + 1 < n_slots) : ((slot + 1) * WARPS + warp < tiles_pp)) prefetch_tile(p_env - (size_t)lane * 16);
      if (EXT && env_i + NT < n_p) ext_prefetch_async(args.env, gi + NT, ext_stage, NT, filt, so);
      // ---------------- phase A: everything that only reads the snapshot ----------------------
      uint32_t cell = 0;
      float target = 0.0f;
      bool done = false, success = false;
      int code = 0;
      uint32_t ep_steps = 0;
      double ep_return = 0.0;
      Env e;
      int a = 0;
      uint32_t sid = 0u, noise_w0 = 0u, noise_w1 = 0u;      // words z, w of the step draw feed the observation noise (off by default)
      size_t trace_i = 0;
      if (valid) {
        env_unpack(cur_raw, e);
        sid = e.sid;
        // R9/R10: epsilon-greedy on the snapshot; both draws are always consumed (quirk Q4).  With eps = 0
        // (every working step > 0) no draw can change the outcome and the Philox call is skipped.
        a = sh.greedy[sid];
        if (W0 || w == 0 || (GENERIC && kk.noise_enabled)) {
          const uint4 d = philox4x32_10_keyed(make_uint4((uint32_t)env_i, t, PURPOSE_STEP, pp.population_id), sh.philox_keys);
          if (W0 || w == 0) {
            const uint32_t thr = __ldg(args.eps_threshold + min(e.episode, (uint32_t)(DQLB200_EPS_LUT - 1)));
            if ((d.x >> 8) < thr) a = (int)__umulhi(d.y, 3u);
          }
          if (GENERIC) { noise_w0 = d.z; noise_w1 = d.w; }
        }
        if (TRACE) {
          trace_i = (size_t)k * (size_t)args.n_total + gi;
          if (args.trace.action_override) {
            const int o = args.trace.action_override[trace_i];
            if (o >= 0) a = o;
          }
        }
        cell = sid * 3u + (uint32_t)a;
      }
      // Same-cell groups of the warp-slot (lanes that update the same cell), formed as soon as the actions are known: the
      // ordered commit applies a group's updates in lane order by handing the running value from member to member.
      // (volatile asm: the compiler otherwise sinks the match down to its first consumer, where its latency -- a few hundred
      // cycles -- is exposed; issued here it is covered by the dynamics)
      uint32_t peers;
      asm volatile("match.any.sync.b32 %0, %1, 0xffffffff;" : "=r"(peers) : "r"(valid ? cell : (0x80000000u | (uint32_t)lane)) : "memory");
      uint2 spn = make_uint2(0u, 0u);
      double r_theta0 = 0.0;
      Obs o = {};
      DState ds = {};
      uint32_t sid2 = 0u, step_count = 0u, cc = 0u;
      if (valid) {
        // R3: the set-point through the tables of the configuration (memoised float64 arithmetic of continuous_action, see
        // dqlb200_config.setpoint_*).  A fresh episode starts from 0 but keeps the old value for shaping (quirk Q11).
        const uint32_t sp_prev = e.sp_idx;
        {
          const SpEntry* en = sp_tab + (e.fresh ? sp_zero : sp_prev) * 3u + (uint32_t)a;
          spn = make_uint2(en->next, __float_as_uint(en->value));
        }
        const float sp = __uint_as_float(spn.y);
        // set-point part of the reward (without the level factor); the first step of an episode shapes against the old set-point
        r_theta0 = *(&sp_tab[sp_prev * 3u + (uint32_t)a].r_step + (e.fresh ? 1 : 0));
        // R4
        dyn_advance(kk, pp, e.b, sp, filt ? &kf : nullptr, so ? &ex : nullptr, kk.vz_train);
        step_count = e.step_count + 1u;
        o = dyn_observe(kk, pp, e.b, (int)step_count, kk.dz_train, filt ? &kf : nullptr, so ? &ex : nullptr);
        if (GENERIC && kk.noise_enabled) add_observation_noise(kk, o, noise_w0, noise_w1);
        // R5
        ds = W0 ? discretise_cuts(kc.cuts[0], kk.angle_cut, o, 0) : discretise_cuts(sh.cuts, kk.angle_cut, o, w);
        sid2 = (uint32_t)ds.id();
        // R6 (sticky result: only ever set, quirk Q9)
        // The priority chain of PKG/mdp.py:359-425 as selects (the ladder of branches diverges inside a warp).
        const bool t_fx = !(o.rel_p >= kk.fz_lo) || (o.rel_p >= kk.fz_hi);
        const bool t_zmin = !(o.z >= kk.z_min_cut), t_zmax = o.z >= kk.z_max_cut;
        const bool t_time = (int)step_count >= kk.timeout_steps;
        const bool goal_bins = !(o.contact || t_fx || t_zmin || t_zmax || t_time) && ds.bp == 1 && ds.bv == 1;
        const bool at_level = W0 || (sid >= (uint32_t)(w * DQLB200_STATES_PER_LEVEL) && ds.level == w);   // previous level == w (it never exceeds w)
        cc = goal_bins ? (at_level ? e.curriculum_check + 1u : 0u) : e.curriculum_check;
        code = e.sticky_success ? DQLB200_NON_TERMINAL_SUCCESS : DQLB200_NON_TERMINAL;
        if (goal_bins && at_level) code = ((int)cc >= kk.success_steps) ? DQLB200_TERMINAL_SUCCESS : DQLB200_NON_TERMINAL_SUCCESS;
        code = t_time ? DQLB200_TERMINAL_TIMEOUT : code;
        code = t_zmax ? DQLB200_TERMINAL_FLYZONE_Z : code;
        code = t_zmin ? DQLB200_TERMINAL_MINIMUM_ALTITUDE : code;
        code = t_fx ? DQLB200_TERMINAL_FLYZONE_X : code;
        code = o.contact ? DQLB200_TERMINAL_CONTACT : code;
        done = code >= DQLB200_TERMINAL_SUCCESS;
        success = code == DQLB200_TERMINAL_SUCCESS;
        // NaN/inf observation (PKG/mdp.py:170 raises): remembered in a register and reported after the slot loop -- a branch here
        // would end the basic block, and ptxas waits at the end of a block for the match issued above
        bad_obs = bad_obs || !(fabsf(o.rel_p) <= 3.4028234664e38f) || !(fabsf(o.rel_v) <= 3.4028234664e38f) || !(fabsf(o.rel_a) <= 3.4028234664e38f);
      }
      const uint32_t lower = peers & ((1u << lane) - 1u);
      const int rank = __popc(lower);                                   // position in the group
      const int n_group = valid ? __popc(peers) : 0;
      const int pred = lower ? (31 - __clz(lower)) : lane;              // the member before this lane
      const bool is_last = valid && (peers >> lane) == 1u;              // no member above: stores the group's result
      const int n_max = __reduce_max_sync(FULL, n_group);               // rounds of the commit (warp-uniform)
      // (the match result is consumed here, a dynamics step after it was issued: its latency is a few hundred cycles)
      // learning-rate hint: the LUT entry for the count this update will PROBABLY see (the cell's count now + the lane's rank;
      // an unsynchronised peek at the live table -- the commit uses it only if the count is still that, so the result does not
      // depend on it).  Issued a reward evaluation before the baton: the load is an L1/L2 round trip.
      const uint32_t c_hint = min(sh.cnt[cell] + (uint32_t)rank, (uint32_t)(DQLB200_ALPHA_LUT - 1));
      const float a_hint = __ldg(alpha_lut + c_hint);
      if (valid) {
        // R7 (float64, reference operation order; level-dependent constants from the host)
        const double phi_p = shaping(kk.w_p, o.rel_p, kk.p_max, kk.rcp_p_max, kk.clip_p_f, GENERIC);
        const double phi_v = shaping(kk.w_v, o.rel_v, kk.v_max, kk.rcp_v_max, kk.clip_v_f, GENERIC);
        const double prev_p = shaping(kk.w_p, e.prev_rel_p, kk.p_max, kk.rcp_p_max, kk.clip_p_f, GENERIC);
        const double prev_v = shaping(kk.w_v, e.prev_rel_v, kk.v_max, kk.rcp_v_max, kk.clip_v_f, GENERIC);
        const bool succ_reward = code == DQLB200_NON_TERMINAL_SUCCESS || code == DQLB200_TERMINAL_SUCCESS;
        const double r = reward_sp(W0 ? kc.reward[0] : sh.reward[ds.level], phi_p, phi_v, prev_p, prev_v, r_theta0, succ_reward);
        // R12 target: r + (gamma * max_a Q_a[s'][a]) * [p-bin changed]   (quirks Q2, Q3), float32 like NEP 50
        const float qn = sh.qmax[sid2];
        const float changed = (e.bp != (uint32_t)ds.bp) ? 1.0f : 0.0f;
        target = fadd((float)r, fmul(fmul(kk.gamma, qn), changed));
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
        e.sp_idx = spn.x;
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
      // The serialised section is the critical path of a global step (n_p / 32 links per population, and every warp waits for
      // its turn), so it holds nothing that can be done outside: groups, ranks and learning rates are known before the baton
      // arrives, the finished episodes are only logged (the window is updated from the log at the end of the step).
      const uint32_t dmask = __ballot_sync(FULL, valid && done);
      uint32_t smask = 0u;
      double ret = 0.0;
      if (dmask) {
        smask = __ballot_sync(FULL, valid && success);
        // deterministic (fixed-tree) sum of the finished episodes' returns
        ret = (valid && done) ? ep_return : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) ret = __dadd_rn(ret, __shfl_xor_sync(FULL, ret, off));
      }
      if (WARPS > 1 && !(slot == 0 && warp == 0)) baton_wait<WARPS>(warp);
      {
        float q = sh.qa[cell];
        const uint32_t c0 = sh.cnt[cell];
        const uint32_t c_pre = min(c0 + (uint32_t)rank, (uint32_t)(DQLB200_ALPHA_LUT - 1));      // R11: pre-increment count
        float alpha = a_hint;
        if (c_pre != c_hint) alpha = __ldg(alpha_lut + c_pre);      // the cell was visited between the peek and the baton (rare below the LUT end)
        // q += alpha * (target - q), members of a group in lane order: round r is member r's, which takes the running value from
        // member r - 1 (one shuffle per round; singleton groups -- most lanes -- are done after round 0)
        if (rank == 0) q = fadd(q, fmul(alpha, fsub(target, q)));
#pragma unroll 1
        for (int r = 1; r < n_max; ++r) {
          const float qp = __shfl_sync(FULL, q, pred);
          if (rank == r) q = fadd(qp, fmul(alpha, fsub(target, qp)));
        }
        if (is_last) {
          sh.qa[cell] = q;
          sh.cnt[cell] = c0 + (uint32_t)n_group;
        }
        if (dmask) {      // finished episodes: append to the log, in commit order
          if (sh.n_ep_ret == EP_RET_CAP || sh.n_ep_ok > EP_OK_CAP - 32) drain_episode_log();      // warp-uniform (every lane reads the same words)
          const int idx = sh.n_ep_ret, base = sh.n_ep_ok;
          const int last = 31 - __clz(dmask);
          __syncwarp();
          if (valid && done) sh.ep_ok[base + __popc(dmask & ((1u << lane) - 1u))] = (uint8_t)success;
          if (lane == 0) {
            sh.ep_ret[idx] = ret;
            sh.n_ep_ret = idx + 1;
            sh.n_ep_ok = base + __popc(dmask);
          }
          if (lane == last) {          // the last finished episode in env order is what the trainer logs
            sh.ps.last_cumulative = ep_return;
            sh.ps.last_steps = (int)ep_steps;
            sh.ps.last_code = code;
          }
        }
      }
      if (WARPS > 1 && !(slot == n_slots - 1 && warp == WARPS - 1)) {
        // no fence: barrier.arrive / barrier.sync on the same barrier order the producer's shared-memory stores before the
        // consumer's loads (the producer / consumer pattern of the PTX ISA, "bar.arrive"); the asm memory clobber keeps the
        // compiler from moving the stores below the arrive
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
      if (valid) env_store(p_cur, e);
      if (filt && valid) kf_store(args.env, gi, kf);
      if (so && valid) ext_store(args.env, gi, ex);
      // queue the finished envs of this warp for the batched reset (outside the baton)
      if (dmask) {
        if (valid && done) reset_queue[n_queued + __popc(dmask & ((1u << lane) - 1u))] = (uint16_t)(slot * 32 + lane);
        n_queued += __popc(dmask);
        if (n_queued > RESET_QUEUE - 32) flush_resets();       // never overflows, whatever fraction of envs finishes at once
      }
      if (slot == 0) DQL_STAMP(6);      // slot 0 done
    }
    if (bad_obs) atomicOr(&sh.ps.error_flags, 1u);
    };
    if constexpr (!GENERIC && !TRACE) {
      if (w == 0) slot_loop(std::true_type{});
      else slot_loop(std::false_type{});
    } else {
      slot_loop(std::false_type{});
    }
    flush_resets();
    env_writes_done();
    // software prefetch of slot 0 of the next global step (this warp's envs are final: resets only touch the warp's own)
    if (k + 1 < args.k_steps) {
      if (warp < tiles_pp) prefetch_tile(p_tile0);
      if (EXT) ext_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), ext_stage, NT, filt, so);
    }
    __syncthreads();
    // ---------------- end of the global step: promotion / next curriculum step (R13, R14) -----
    DQL_STAMP(3);      // first barrier of the end of the global step passed
    steps_done += (uint64_t)n_p;
    if (warp == 0) drain_episode_log();      // the other warps build the next snapshot meanwhile
    if (tid == 0) {
      sh.ps.t = t + 1u;
      sh.do_advance = (sh.promote || sh.advance) ? 1 : 0;
      sh.ps.total_episodes += sh.step_episodes; sh.ps.total_successes += sh.step_success; sh.ps.episode_steps_sum += sh.step_ep_steps;
      sh.step_episodes = sh.step_success = sh.step_ep_steps = 0u;
    }
    if (tid < 9) { sh.ps.termination_hist[tid] += sh.step_hist[tid]; sh.step_hist[tid] = 0u; }
    if (k + 1 < args.k_steps) build_snapshot(w);      // for the next step, unless the curriculum advances (rebuilt there)
    __syncthreads();
    if (sh.do_advance) advance_curriculum(w, t + 1u);
  }

  DQL_STAMP(4);      // step loop left
  // ---- write back (live rows only) ----------------------------------------------------------------
  if (tile_pending) wait_tile();                         // a prefetch issued for a step that did not run: the CTA must not exit under it
  if (EXT) asm volatile("cp.async.wait_all;" ::: "memory");
  // No barrier here: every path into this point ends with a CTA barrier that lies behind the last write of another warp to the
  // tables and to the trainer state (the second barrier of a global step, the one that ends advance_curriculum, the prologue's).
  for (int i = tid; i < (sh.ps.working_step + 1) * DQLB200_CELLS_PER_LEVEL; i += NT) {
    gt[i] = __float_as_uint(sh.qa[i]);
    gt[2 * CELLS + i] = sh.cnt[i];
  }
  if (warp == 0) {      // the trainer state is finished and written back by one warp (no CTA barrier between the two)
    if (lane == 0) {
      dqlb200_population_state& ps = sh.ps;
      ps.total_steps += steps_done;
    }
    __syncwarp();
    uint32_t* gps = reinterpret_cast<uint32_t*>(args.pop_state + pop);
    for (int i = lane; i < PS_WORDS; i += 32) gps[i] = reinterpret_cast<const uint32_t*>(&sh.ps)[i];
  }
  DQL_STAMP(5);      // exit
}


}  // namespace dql
