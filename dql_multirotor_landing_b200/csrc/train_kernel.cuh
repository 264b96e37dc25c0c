// train_kernel.cuh -- the fused training kernel: one CTA per population = WARPS worker warps + one service warp.
// Workers run phase A of their envs (select, set-point, dynamics, discretise, check, reward, target) and stream the env state
// (bulk async copies into shared memory, mbarrier completion); the service warp applies the Q / count updates in env order
// (the ordered commit of semantics S1), keeps the success window and decides promotions.  Records travel from the workers to
// the service warp through a ring in shared memory guarded by mbarriers, so a worker never waits for another worker.
// Part of libdqlb200 (see dqlb200.cu for the kernel inventory and the C-ABI).
#pragma once
#include "env_state.cuh"

namespace dql {

// ---- mbarrier + bulk-copy primitives (sm_90+; SASS: SYNCS.*, UBLKCP) ------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
// waits until the phase with the given parity has completed (a fresh barrier has "completed" the phase of parity 1)
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "DQL_MBAR_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra DQL_MBAR_DONE_%=;\n\t"
      "bra DQL_MBAR_WAIT_%=;\n"
      "DQL_MBAR_DONE_%=:\n\t}" ::"r"(addr), "r"(parity), "r"(100000u) : "memory");      // suspend-time hint [ns]: sleep, do not poll
}
__device__ __forceinline__ void mbar_arrive(unsigned addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned addr, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16); completion is signalled on the mbarrier as transaction bytes
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(mbar)
               : "memory");
}
// orders this thread's earlier generic-proxy accesses (plain loads / stores) before its later async-proxy operations (bulk copies)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

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
  size_t env_stride, env_stride2;          // 16 n_total, 32 n_total: byte offsets of the B and C vectors behind an env's A vector
};

constexpr int RESET_QUEUE = 64;    // finished envs a warp collects before it runs the batched reset pass
constexpr int RING = 2;            // records a worker warp may be ahead of the service warp
constexpr int MAX_WORKER_WARPS = 8;

constexpr int STATES = DQLB200_MAX_CURRICULUM * DQLB200_STATES_PER_LEVEL;   // 945

// What a worker warp hands to the service warp for one slot: per lane the visited cell (+ flags) and the update target; per
// warp-slot the reductions over its finished episodes (only written when there is one).
constexpr uint32_t REC_CELL_MASK = 0xFFFu, REC_DONE = 1u << 16, REC_SUCCESS = 1u << 17, REC_INVALID = 1u << 31;
struct RecordTail {
  double ret, last_cum;            // fixed-tree sum of the finished episodes' returns; return of the last one (env order)
  int last_steps, last_code;
};

// Shared memory of one population (CTA).  Q_b and the alpha LUT stay in global memory (read-only in the step loop).
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
  unsigned long long bar_state[MAX_WORKER_WARPS];          // env tile of a worker warp has landed (transaction bytes)
  unsigned long long bar_full[MAX_WORKER_WARPS][RING];     // record posted by the worker warp
  unsigned long long bar_empty[MAX_WORKER_WARPS][RING];    // record consumed by the service warp
  RecordTail tail[MAX_WORKER_WARPS][RING];
  // followed by (dynamic): uint16_t reset_queue[WARPS][RESET_QUEUE]; uint4 stage[3 or 6][NT]; uint2 rec[WARPS][RING][32]
};

__host__ __device__ constexpr size_t train_smem_bytes(int worker_threads, bool extended) {
  return ((sizeof(Shared) + 15) & ~size_t(15)) + (size_t)(worker_threads / 32) * RESET_QUEUE * sizeof(uint16_t) +
         (size_t)(extended ? 6 : 3) * worker_threads * 16 + (size_t)(worker_threads / 32) * RING * 32 * sizeof(uint2);
}

#ifndef DQL_REGS
#define DQL_REGS 64                  // register budget per thread of the first-order instances (6 CTAs of 4 + 1 warps per SM)
#endif
#ifndef DQL_REGS_EXT
#define DQL_REGS_EXT 96              // the extended instance carries the estimator and the second-order model
#endif
template <bool GENERIC_> struct ConstsOf;
template <> struct ConstsOf<true> {
  __device__ __forceinline__ static const KC& get(const KC& kc) { return kc; }
};
template <> struct ConstsOf<false> {
  __device__ __forceinline__ static KDef get(const KC&) { return KDef{}; }
};
__host__ __device__ constexpr int train_min_blocks(int warps, int variant) {      // resident CTAs per SM the register budget allows
  const int c = 65536 / (variant == 2 ? DQL_REGS_EXT : DQL_REGS) / ((warps + 1) * 32);
  return c > 0 ? c : 1;
}

// GENERIC = false is the production instance of the reference's default configuration; GENERIC = true adds what only
// non-default configurations need: run-time constants instead of the compile-time defaults (KDef), the second Markstein
// correction step of x / p_max, x / v_max (required unless the divisors are the exhaustively verified defaults) and the
// observation-noise option.  The trace instances are generic (both division variants are correctly rounded, hence identical).
// VARIANT 0 = production, 1 = generic, 2 = extended: generic plus the options that carry extra per-env state (acceleration
// estimator, second-order model) -- a separate instance so that the generic one does not pay for their branches and registers;
// 3 = production for populations that fill every slot (envs_per_population a multiple of the worker threads): `valid` is a
// compile-time constant, which removes the predicate, the defaults of the invalid lanes and their reconvergence points.
template <int WARPS, bool TRACE, int VARIANT>
__global__ void __launch_bounds__((WARPS + 1) * 32, train_min_blocks(WARPS, VARIANT)) train_kernel(const __grid_constant__ KC kc, const TrainArgs args) {
  constexpr bool GENERIC = VARIANT == 1 || VARIANT == 2, EXT = VARIANT == 2, FULL_SLOTS = VARIANT == 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
  const auto& kk = ConstsOf<GENERIC>::get(kc);       // run-time KC (generic) or the compile-time defaults KDef (production)
  constexpr int NT = WARPS * 32;                     // worker threads: env index = slot * NT + tid, the order of semantics S1
  constexpr int NTT = NT + 32;                       // + the service warp
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(FULL, tid >> 5, 0);   // broadcast: the compiler may keep warp-dependent addresses in uniform registers
  const bool worker = warp < WARPS;                  // warp-uniform role
  const int pop = blockIdx.x + args.pop_offset;
  const int n_p = kc.envs_per_population;
  const int n_slots = (n_p + NT - 1) / NT;
  unsigned char* dyn = smem_raw + ((sizeof(Shared) + 15) & ~size_t(15));
  uint16_t* reset_queue = reinterpret_cast<uint16_t*>(dyn) + (size_t)(worker ? warp : 0) * RESET_QUEUE;
  uint4* stage = reinterpret_cast<uint4*>(dyn + (size_t)WARPS * RESET_QUEUE * sizeof(uint16_t));   // [3][NT] (+ [3][NT] extension-state slots, extended variant)
  uint2* rec = reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(stage) + (size_t)(EXT ? 6 : 3) * NT * 16);   // [WARPS][RING][32]
  const unsigned stage_addr = (unsigned)__cvta_generic_to_shared(stage + (worker ? tid : 0));
  const size_t env_base = (size_t)pop * n_p;
  uint32_t* gt = args.tables + (size_t)pop * 3 * CELLS;
  float* gqb = reinterpret_cast<float*>(gt + CELLS);      // table B, written only by the transfer below
  const unsigned bar_state = (unsigned)__cvta_generic_to_shared(&sh.bar_state[worker ? warp : 0]);
  const unsigned bar_full0 = (unsigned)__cvta_generic_to_shared(&sh.bar_full[0][0]);
  const unsigned bar_empty0 = (unsigned)__cvta_generic_to_shared(&sh.bar_empty[0][0]);

  // ---- env-state streaming (workers): the 48 bytes of the 32 envs of a warp-slot are three contiguous 512-byte runs of the
  // SoA; lane 0 fetches them with three bulk copies into the warp's part of `stage`, completion on the warp's mbarrier.
  uint32_t n_landed = 0;            // completed waits on bar_state (phase parity)
  bool st_pending = false;          // a tile is in flight or landed and not yet taken
  const unsigned tile_addr = (unsigned)__cvta_generic_to_shared(stage + (worker ? warp * 32 : 0));
  // `rewritten`: the envs of the tile may have been written with plain stores since the warp's last fence (the previous global
  // step's write-back, resets) -- the generic-proxy stores of all lanes (ordered before lane 0 by the __syncwarp in front of
  // every call) must be ordered before the async-proxy read.  One fence per global step: within a step a tile is read before
  // it is written.
  auto issue_tile = [&](int slot, bool rewritten) {       // called by all lanes of a worker warp (converged)
    const int first = slot * NT + warp * 32;
    if (!FULL_SLOTS && first >= n_p) return;
    const unsigned nv = FULL_SLOTS ? 32u : (unsigned)min(32, n_p - first);
    if (lane == 0) {
      if (rewritten) fence_proxy_async();
      mbar_arrive_expect_tx(bar_state, 48u * nv);
      const char* src = reinterpret_cast<const char*>(args.env.a + env_base + first);
      bulk_g2s(tile_addr, src, 16u * nv, bar_state);
      bulk_g2s(tile_addr + 16u * NT, src + args.env_stride, 16u * nv, bar_state);
      bulk_g2s(tile_addr + 32u * NT, src + args.env_stride2, 16u * nv, bar_state);
    }
    st_pending = true;
  };
  auto take_tile = [&]() {                // waits for the tile issued last (no-op for a warp without envs in that slot)
    if (st_pending) {
      mbar_wait(bar_state, n_landed & 1u);
      ++n_landed;
      st_pending = false;
    }
  };

  // ---- stage population state and the LIVE rows of the tables in shared memory --------------------
  // At working step w only levels 0..w can be visited (a state's level never exceeds w), so only rows
  // [0, (w+1)*567) of Q_a / count are staged, snapshotted and written back; a promotion loads the next level.
  static_assert(sizeof(dqlb200_population_state) % 4 == 0, "word copies");
  constexpr int PS_WORDS = sizeof(dqlb200_population_state) / 4;
  const dqlb200_population_params pp = args.pop_params[pop];
  if (tid == 0) {
    for (int w = 0; w < WARPS; ++w) {
      mbar_init((unsigned)__cvta_generic_to_shared(&sh.bar_state[w]), 1u);
      for (int d = 0; d < RING; ++d) {
        mbar_init(bar_full0 + 8u * (w * RING + d), 1u);
        mbar_init(bar_empty0 + 8u * (w * RING + d), 1u);
      }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    const int w_start = args.pop_state[pop].working_step;
    const uint32_t* gps = reinterpret_cast<const uint32_t*>(args.pop_state + pop);
    for (int i = tid; i < PS_WORDS; i += NTT) reinterpret_cast<uint32_t*>(&sh.ps)[i] = gps[i];
    if (tid == NT) {
      sh.n_episodes = sh.n_success = sh.ep_steps = 0ull;
      sh.step_episodes = sh.step_success = sh.step_ep_steps = 0u;
      for (int i = 0; i < 9; ++i) { sh.hist[i] = 0ull; sh.step_hist[i] = 0u; }
      sh.promote = sh.advance = sh.do_advance = 0;
      sh.cuts = kc.cuts[w_start];
    }
    if (tid >= NT && tid < NT + 5) sh.reward[tid - NT] = kc.reward[tid - NT];
    const int live = (w_start + 1) * DQLB200_CELLS_PER_LEVEL;
    for (int i = tid; i < live; i += NTT) {
      sh.qa[i] = __uint_as_float(gt[i]);
      sh.cnt[i] = gt[2 * CELLS + i];
      if ((i & 31) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(gqb + i));     // table B rows for the first snapshot
    }
  }
  __syncthreads();      // barriers initialised, population state staged
  if (worker) {         // the env state of slot 0 -- a cold HBM read when one global step is run per launch -- flies during the snapshot
    issue_tile(0, false);
    if (EXT) ext_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), stage_addr, NT, kc.accel_mode != 0, kc.dynamics_model != 0);
  }

  const float* __restrict__ alpha_lut = args.alpha_luts + (size_t)pp.alpha_lut * DQLB200_ALPHA_LUT;
  const float alpha_min = __ldg(alpha_lut + DQLB200_ALPHA_LUT - 1);      // alpha(count >= 1002), PKG/trainer.py:95-105
  const bool filt = EXT && kc.accel_mode != 0;      // acceleration estimator: 16 more bytes per env, extended instance only
  const bool so = EXT && kc.dynamics_model != 0;    // second-order attitude + vertical PID: 32 more bytes per env
  uint64_t steps_done = 0;
  uint32_t ring_it = 0;             // records posted (worker) / consumed per worker warp (service): entry = it % RING, phase = it / RING

  // snapshot of a global step: greedy action (first max of (Q_a+Q_b)/2, PKG/double_q_learning.py:119-124) and bootstrap
  // value max_a Q_a (:136-141) of every live state, from the live table as the previous step left it
  auto build_snapshot = [&](int w) {
    for (int st = tid; st < (w + 1) * DQLB200_STATES_PER_LEVEL; st += NTT) {
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
      for (int i = tid; i < DQLB200_CELLS_PER_LEVEL; i += NTT) {
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
      for (int i = tid; i < DQLB200_CELLS_PER_LEVEL; i += NTT) {
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
    if (!sh.ps.finished && worker) {
      take_tile();        // a tile fetched before the restart is stale
      for (int slot = 0; slot < n_slots; ++slot) {
        const int env_i = slot * NT + tid;
        if (env_i < n_p) {
          Env e;
          Kf kf;
          Ext ex;
          if (filt) kf = kf_load(args.env, env_base + env_i);       // the estimator outlives the curriculum step
          if (so) ex = ext_load(args.env, env_base + env_i);
          env_reset(kk, pp, sh.cuts, kk.angle_cut, e, (uint32_t)env_i, birth, w + 1, /*fresh_mdp=*/true, filt ? &kf : nullptr, so ? &ex : nullptr);
          env_store(args.env, env_base + env_i, e);
          if (filt) kf_store(args.env, env_base + env_i, kf);
          if (so) ext_store(args.env, env_base + env_i, ex);
        }
      }
      __syncwarp();
      issue_tile(0, true);      // every env of this warp was just rewritten by this warp
      if (EXT) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        ext_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), stage_addr, NT, filt, so);
      }
    }
    if (!sh.ps.finished) build_snapshot(w + 1);
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

  for (int k = 0; k < args.k_steps; ++k) {
    if (sh.ps.finished) break;     // uniform: written only between barriers
    const int w = sh.ps.working_step;
    const uint32_t t = sh.ps.t;

    if (worker) {
      // =============================== worker warps ===============================
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
            Kf kf;
            Ext ex;
            if (filt) kf = kf_load(args.env, gr);
            if (so) ex = ext_load(args.env, gr);
            env_reset(kk, pp, sh.cuts, kk.angle_cut, e, (uint32_t)env_r, t + 1u, w, /*fresh_mdp=*/false, filt ? &kf : nullptr, so ? &ex : nullptr);
            env_store(args.env, gr, e);
            if (filt) kf_store(args.env, gr, kf);
            if (so) ext_store(args.env, gr, ex);
          }
        }
        __syncwarp();
        n_queued = 0;
      };

      char* p_env = reinterpret_cast<char*>(args.env.a + env_base + tid);      // running pointer to the A vector of this thread's env
      for (int slot = 0; slot < n_slots; ++slot) {
        const int env_i = slot * NT + tid;
        const bool valid = FULL_SLOTS || env_i < n_p;
        const size_t gi = env_base + (size_t)env_i;          // only dereferenced under `valid`
        take_tile();
        EnvRaw cur_raw;
        {
          const uint4 a = stage[tid];
          cur_raw.A = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
          cur_raw.B = stage[NT + tid];
          cur_raw.C = stage[2 * NT + tid];
        }
        char* const p_cur = p_env;
        p_env += NT * 16;
        Kf kf;
        Ext ex;
        if (EXT) asm volatile("cp.async.wait_group 0;" ::: "memory");
        if (filt) kf = kf_take(stage, NT, tid);
        if (so) ex = ext_take(stage, NT, tid);
        __syncwarp();            // every lane has read its part of the tile: it may be overwritten
        if (slot + 1 < n_slots) {      // in flight during this slot
          issue_tile(slot + 1, false);
          if (EXT && env_i + NT < n_p) ext_prefetch_async(args.env, gi + NT, stage_addr, NT, filt, so);
        }
        // ---------------- phase A: everything that only reads the snapshot ----------------------
        uint32_t cell = 0;
        float target = 0.0f;
        bool done = false, success = false;
        int code = 0;
        uint32_t ep_steps = 0;
        double ep_return = 0.0;
        Env e;
        if (valid) {
          env_unpack(cur_raw, e);
          const uint32_t sid = e.sid;
          // R9/R10: epsilon-greedy on the snapshot; both draws are always consumed (quirk Q4).  With eps = 0
          // (every working step > 0) no draw can change the outcome and the Philox call is skipped.
          int a = sh.greedy[sid];
          uint32_t noise_w0 = 0u, noise_w1 = 0u;      // words z, w of the step draw feed the observation noise (off by default)
          if (w == 0 || (GENERIC && kk.noise_enabled)) {
            const uint4 d = philox4x32_10(make_uint4((uint32_t)env_i, t, PURPOSE_STEP, pp.population_id), pp.seed_lo, pp.seed_hi);
            if (w == 0) {
              const uint32_t thr = __ldg(args.eps_threshold + min(e.episode, (uint32_t)(DQLB200_EPS_LUT - 1)));
              if ((d.x >> 8) < thr) a = (int)__umulhi(d.y, 3u);
            }
            if (GENERIC) { noise_w0 = d.z; noise_w1 = d.w; }
          }
          size_t trace_i = 0;
          if (TRACE) {
            trace_i = (size_t)k * (size_t)args.n_total + gi;
            if (args.trace.action_override) {
              const int o = args.trace.action_override[trace_i];
              if (o >= 0) a = o;
            }
          }
          cell = sid * 3u + (uint32_t)a;
          // R3: set-point (float64).  A fresh episode starts from 0 but keeps the old value for shaping.
          const double prev_sp = e.theta_sp;
          const double sp = apply_action(kk, e.fresh ? 0.0 : e.theta_sp, a);
          // R4
          dyn_advance(kk, pp, e.b, (float)sp, filt ? &kf : nullptr, so ? &ex : nullptr, kk.vz_train);
          const uint32_t step_count = e.step_count + 1u;
          Obs o = dyn_observe(kk, pp, e.b, (int)step_count, kk.dz_train, filt ? &kf : nullptr, so ? &ex : nullptr);
          if (GENERIC && kk.noise_enabled) add_observation_noise(kk, o, noise_w0, noise_w1);
          // R5
          const DState ds = discretise_cuts(sh.cuts, kk.angle_cut, o, w);
          const uint32_t sid2 = (uint32_t)ds.id();
          // R6 (sticky result: only ever set, quirk Q9)
          // The priority chain of PKG/mdp.py:359-425 as selects (the ladder of branches diverges inside a warp).
          const bool t_fx = !(o.rel_p >= kk.fz_lo) || (o.rel_p >= kk.fz_hi);
          const bool t_zmin = !(o.z >= kk.z_min_cut), t_zmax = o.z >= kk.z_max_cut;
          const bool t_time = (int)step_count >= kk.timeout_steps;
          const bool goal_bins = !(o.contact || t_fx || t_zmin || t_zmax || t_time) && ds.bp == 1 && ds.bv == 1;
          const bool at_level = sid >= (uint32_t)(w * DQLB200_STATES_PER_LEVEL) && ds.level == w;   // previous level == w (it never exceeds w)
          const uint32_t cc = goal_bins ? (at_level ? e.curriculum_check + 1u : 0u) : e.curriculum_check;
          code = e.sticky_success ? DQLB200_NON_TERMINAL_SUCCESS : DQLB200_NON_TERMINAL;
          if (goal_bins && at_level) code = ((int)cc >= kk.success_steps) ? DQLB200_TERMINAL_SUCCESS : DQLB200_NON_TERMINAL_SUCCESS;
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
          const double phi_p = shaping(kk.w_p, o.rel_p, kk.p_max, kk.rcp_p_max, kk.clip_p_f, GENERIC);
          const double phi_v = shaping(kk.w_v, o.rel_v, kk.v_max, kk.rcp_v_max, kk.clip_v_f, GENERIC);
          const double phi_t = __dmul_rn(kk.w_theta, fabs(div_f64_by_const(sp, kk.theta_max, kk.rcp_theta_max)));
          const double prev_p = shaping(kk.w_p, e.prev_rel_p, kk.p_max, kk.rcp_p_max, kk.clip_p_f, GENERIC);
          const double prev_v = shaping(kk.w_v, e.prev_rel_v, kk.v_max, kk.rcp_v_max, kk.clip_v_f, GENERIC);
          const double prev_t = __dmul_rn(kk.w_theta, fabs(div_f64_by_const(prev_sp, kk.theta_max, kk.rcp_theta_max)));
          const bool succ_reward = code == DQLB200_NON_TERMINAL_SUCCESS || code == DQLB200_TERMINAL_SUCCESS;
          const double r = reward_f64(kk, sh.reward[ds.level], phi_p, phi_v, phi_t, prev_p, prev_v, prev_t, succ_reward);
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
        // ---------------- hand the slot to the service warp ----------------------------------
        const uint32_t dmask = __ballot_sync(FULL, valid && done);
        const unsigned entry = ring_it % RING;
        const unsigned b_off = 8u * ((unsigned)warp * RING + entry);
        mbar_wait(bar_empty0 + b_off, ((ring_it / RING) & 1u) ^ 1u);      // the entry's previous record has been consumed
        rec[((unsigned)warp * RING + entry) * 32u + lane] =
            make_uint2(valid ? (cell | (done ? REC_DONE : 0u) | (success ? REC_SUCCESS : 0u)) : REC_INVALID, __float_as_uint(target));
        if (dmask) {
          // reductions over the finished episodes of this warp-slot: deterministic (fixed-tree) sum of their returns
          double ret = (valid && done) ? ep_return : 0.0;
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) ret = __dadd_rn(ret, __shfl_xor_sync(FULL, ret, off));
          const int last = 31 - __clz(dmask);
          const int last_steps = __shfl_sync(FULL, (int)ep_steps, last);
          const int last_code = __shfl_sync(FULL, code, last);
          const double last_cum = __shfl_sync(FULL, ep_return, last);
          if (lane == 0) sh.tail[warp][entry] = RecordTail{ret, last_cum, last_steps, last_code};
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full0 + b_off);
        ++ring_it;
        // order-independent episode counters
        if (dmask) {
          if (valid && done) {
            atomicAdd(&sh.step_hist[code], 1u);
            atomicAdd(&sh.step_ep_steps, ep_steps);
          }
          if (lane == 0) atomicAdd(&sh.step_episodes, (uint32_t)__popc(dmask));
        }
        if (valid) env_store_p(p_cur, args.env_stride, args.env_stride2, e);
        if (filt && valid) kf_store(args.env, gi, kf);
        if (so && valid) ext_store(args.env, gi, ex);
        // queue the finished envs of this warp for the batched reset
        if (dmask) {
          if (valid && done) reset_queue[n_queued + __popc(dmask & ((1u << lane) - 1u))] = (uint16_t)(slot * 32 + lane);
          n_queued += __popc(dmask);
          if (n_queued > RESET_QUEUE - 32) flush_resets();       // never overflows, whatever fraction of envs finishes at once
        }
      }
      flush_resets();
      // slot 0 of the next global step (this warp's envs are final: resets only touch the warp's own)
      if (k + 1 < args.k_steps) {
        issue_tile(0, true);
        if (EXT) ext_prefetch_async(args.env, env_base + (size_t)min(tid, n_p - 1), stage_addr, NT, filt, so);
      }
    } else {
      // =============================== service warp: ordered commit ===============================
      // Semantics S1: the updates of a global step are applied one after the other in env order (slot, worker warp, lane) on
      // the live table, each with the learning rate of the live pre-increment count.  Inside a warp-slot, lanes with the
      // same cell form a group (__match_any_sync) whose updates are applied in lane order by shuffles.
      for (int slot = 0; slot < n_slots; ++slot) {
        const unsigned entry = ring_it % RING, par = (ring_it / RING) & 1u;
        // Pass 1, all worker warps of the slot: pick the records up, form the same-cell groups and start the learning-rate
        // loads.  The LUT lives in global memory (an L2 round trip: the SM's L1 is carved out for shared memory); fetched
        // here for the count each update will PROBABLY see (live count now + rank in its group), the loads of all warps are in
        // flight together instead of one per commit in the serial chain.
        uint2 rc[WARPS];
        uint32_t peers_[WARPS], c_hint[WARPS];
        float a_hint[WARPS];
#pragma unroll
        for (int ww = 0; ww < WARPS; ++ww) {
          mbar_wait(bar_full0 + 8u * ((unsigned)ww * RING + entry), par);
          rc[ww] = rec[((unsigned)ww * RING + entry) * 32u + lane];
          const bool valid = (rc[ww].x & REC_INVALID) == 0u;
          const uint32_t cell = rc[ww].x & REC_CELL_MASK;
          peers_[ww] = __match_any_sync(FULL, valid ? cell : (0x8000u | (uint32_t)lane));
          c_hint[ww] = min(sh.cnt[cell] + (uint32_t)__popc(peers_[ww] & ((1u << lane) - 1u)), (uint32_t)(DQLB200_ALPHA_LUT - 1));
          a_hint[ww] = __ldg(alpha_lut + c_hint[ww]);
        }
        // Pass 2: the commits, one worker warp after the other
#pragma unroll
        for (int ww = 0; ww < WARPS; ++ww) {
          const bool valid = (rc[ww].x & REC_INVALID) == 0u;
          const uint32_t cell = rc[ww].x & REC_CELL_MASK;
          const float target = __uint_as_float(rc[ww].y);
          const uint32_t peers = peers_[ww];
          const int rank = __popc(peers & ((1u << lane) - 1u));
          float q = sh.qa[cell];
          const uint32_t c0 = sh.cnt[cell];
          const uint32_t c_pre = min(c0 + (uint32_t)rank, (uint32_t)(DQLB200_ALPHA_LUT - 1));      // R11: pre-increment count
          float alpha = a_hint[ww];
          if (c_pre != c_hint[ww]) alpha = __ldg(alpha_lut + c_pre);      // an earlier warp of this slot visited the cell too
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
          const uint32_t dmask = __ballot_sync(FULL, valid && (rc[ww].x & REC_DONE) != 0u);
          if (dmask) {
            const uint32_t smask = __ballot_sync(FULL, valid && (rc[ww].x & REC_SUCCESS) != 0u);
            if (lane == 0) {
              const RecordTail tl = sh.tail[ww][entry];
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
              ps.return_sum = __dadd_rn(ps.return_sum, tl.ret);
              ps.last_code = tl.last_code;
              ps.last_steps = tl.last_steps;
              ps.last_cumulative = tl.last_cum;
              sh.step_success += (uint32_t)__popc(smask);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_empty0 + 8u * ((unsigned)ww * RING + entry));
        }
        ++ring_it;
      }
    }
    __syncthreads();       // every commit and every reset of the step is done
    // ---------------- end of the global step: promotion / next curriculum step (R13, R14) -----
    steps_done += (uint64_t)n_p;
    if (tid == 0) {
      sh.ps.t = t + 1u;
      sh.do_advance = (sh.promote || sh.advance) ? 1 : 0;
      sh.n_episodes += sh.step_episodes; sh.n_success += sh.step_success; sh.ep_steps += sh.step_ep_steps;
      sh.step_episodes = sh.step_success = sh.step_ep_steps = 0u;
    }
    if (tid < 9) { sh.hist[tid] += sh.step_hist[tid]; sh.step_hist[tid] = 0u; }
    if (k + 1 < args.k_steps) build_snapshot(w);      // for the next step, unless the curriculum advances (rebuilt there)
    __syncthreads();
    if (sh.do_advance) advance_curriculum(w, t + 1u);
  }

  // ---- write back (live rows only) ----------------------------------------------------------------
  if (worker) {
    take_tile();          // a tile fetched for a step that did not run
    if (EXT) asm volatile("cp.async.wait_all;" ::: "memory");
  }
  __syncthreads();
  for (int i = tid; i < (sh.ps.working_step + 1) * DQLB200_CELLS_PER_LEVEL; i += NTT) {
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
    for (int i = tid; i < PS_WORDS; i += NTT) gps[i] = reinterpret_cast<const uint32_t*>(&sh.ps)[i];
  }
}


}  // namespace dql
