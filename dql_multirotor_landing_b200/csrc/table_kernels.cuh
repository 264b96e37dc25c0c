// table_kernels.cuh -- kernels on the Q/count tables: transfer, shared-table pack/apply, replica merge, atomic-roof micro-benchmark
// Part of libdqlb200 (see dqlb200.cu for the kernel inventory and the C-ABI).
#pragma once
#include "train_kernel.cuh"

namespace dql {

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

// Shared-table mode (one agent replicated on G devices).  Each rank trains on its own envs for a few steps; the ranks' tables
// are then merged exactly like the replicas of replica-merge mode, the ranks being the replicas:
//   dcount_g = count_g - count_snap,  Q <- Q_snap + sum_g (Q_g - Q_snap) * dcount_g / sum_g dcount_g  (RANK ORDER, float32),
//   count <- count_snap + sum_g dcount_g;  a cell ONE rank visited takes that rank's value bit for bit.
// The exchange is an all-reduce realised as all-gather + a reduction in rank order inside shared_apply_kernel: every rank
// contributes its raw table words (Q_a bits, 32-bit counts) and integer trainer counters, so counts and counters are exact
// for any number of visits, and the float32 sum has ONE defined order -- identical on every rank, from run to run and for
// every NCCL algorithm (an fp32 SUM all-reduce is none of these).
// One "agent" = a group of R = replicas_per_population consecutive populations whose tables are identical (after
// replica_merge_kernel; R = 1: a plain population).  packed: ONE entry of DQLB200_SHARED_WORDS 32-bit words per agent:
// [Q_a bits | count | successes in the windows, finished episodes of the curriculum step (lo, hi), alive].
constexpr int SHARED_WORDS = DQLB200_SHARED_WORDS;
__global__ void __launch_bounds__(256) shared_pack_kernel(const uint32_t* tables, uint32_t* packed, const dqlb200_population_state* ps,
                                                          int n_agents, int R) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long long)n_agents * CELLS) {
    const long long g = i / CELLS, c = i % CELLS;
    const size_t tb = (size_t)g * R * 3 * CELLS;
    uint32_t* d = packed + (size_t)g * SHARED_WORDS;
    d[c] = tables[tb + c];
    d[CELLS + c] = tables[tb + 2 * CELLS + c];
  }
  // The agent's pooled trainer counters: the block that holds cell 0 of an agent (at most one agent starts inside a block,
  // CELLS > blockDim) reduces over the agent's R replicas with all its threads -- one thread walking R population states took
  // longer than everything else in a sync (R = 888 per GPU in shared-table mode).
  const long long first = (long long)blockIdx.x * blockDim.x;
  const long long g0 = (first + CELLS - 1) / CELLS;
  if (g0 < n_agents && g0 * CELLS < first + blockDim.x) {      // block-uniform
    __shared__ unsigned long long s_succ, s_eps;
    __shared__ int s_dead;
    if (threadIdx.x == 0) { s_succ = s_eps = 0ull; s_dead = 0; }
    __syncthreads();
    unsigned long long successes = 0, episodes = 0;
    int dead = 0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
      const dqlb200_population_state& p = ps[g0 * R + r];
      successes += (unsigned long long)p.window_sum;
      episodes += (unsigned long long)p.episodes_in_step;
      dead |= (p.finished || p.pending_advance) ? 1 : 0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      successes += __shfl_xor_sync(FULL, successes, off);
      episodes += __shfl_xor_sync(FULL, episodes, off);
      dead |= __shfl_xor_sync(FULL, dead, off);
    }
    if ((threadIdx.x & 31) == 0) {
      if (successes) atomicAdd(&s_succ, successes);
      if (episodes) atomicAdd(&s_eps, episodes);
      if (dead) atomicOr(&s_dead, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t* d = packed + (size_t)g0 * SHARED_WORDS + 2 * CELLS;
      d[0] = (uint32_t)s_succ;
      d[1] = (uint32_t)s_eps;
      d[2] = (uint32_t)(s_eps >> 32);
      d[3] = s_dead ? 0u : 1u;
    }
  }
}
// gathered: [n_ranks][n_agents][SHARED_WORDS], rank-major (the layout all-gather produces)
__global__ void shared_apply_kernel(uint32_t* tables, uint32_t* snap, uint32_t* merge_snap, const uint32_t* __restrict__ gathered, int n_ranks,
                                    dqlb200_population_state* ps, int n_agents, int R, int pooled_promote, long long max_episodes) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_agents * CELLS) return;
  const long long g = i / CELLS, c = i % CELLS;
  const size_t tb = (size_t)g * R * 3 * CELLS, sb = (size_t)g * 3 * CELLS;
  const size_t rank_stride = (size_t)n_agents * SHARED_WORDS;
  const uint32_t* d = gathered + (size_t)g * SHARED_WORDS;
  const uint32_t c_snap = snap[sb + 2 * CELLS + c];
  const float q_snap = __uint_as_float(snap[sb + c]);
  float sum = 0.0f;
  unsigned long long total = 0ull;
  int visitors = 0;
  uint32_t single = 0u;
  for (int r = 0; r < n_ranks; ++r) {           // rank order: the one defined order of the float32 sum
    const uint32_t dc = d[r * rank_stride + CELLS + c] - c_snap;
    if (dc) {
      const uint32_t qb = d[r * rank_stride + c];
      sum = fadd(sum, fmul(fsub(__uint_as_float(qb), q_snap), __uint2float_rn(dc)));
      total += dc;
      single = qb;
      visitors += 1;
    }
  }
  uint32_t q_bits = snap[sb + c], cnt = c_snap;
  if (visitors == 1) q_bits = single;       // one rank visited: its value, bit for bit, on every rank
  else if (visitors > 1) q_bits = __float_as_uint(fadd(q_snap, __fdiv_rn(sum, __ull2float_rn(total))));
  if (visitors > 0) {
    const unsigned long long c_new = (unsigned long long)c_snap + total;
    cnt = c_new > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)c_new;        // the 32-bit count saturates (alpha is constant from 1002 on)
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
  if (c == 0 && pooled_promote > 0) {      // promotion pooled over every rank's windows (same decision on every rank), exact integers
    unsigned long long successes = 0ull, episodes = 0ull;
    bool alive = true;
    for (int r = 0; r < n_ranks; ++r) {
      const uint32_t* t = d + r * rank_stride + 2 * CELLS;
      successes += t[0];
      episodes += (unsigned long long)t[1] | ((unsigned long long)t[2] << 32);
      alive = alive && t[3] != 0u;
    }
    const int pending = !alive ? 0 : (successes >= (unsigned long long)pooled_promote ? 1 : ((long long)episodes >= max_episodes ? 2 : 0));
    if (pending)
      for (int r = 0; r < R; ++r) ps[g * R + r].pending_advance = pending;
  }
}

// Replica-merge mode: R consecutive populations are replicas of ONE agent.  One CTA (NW warps) per tile of 32 live cells:
//   load   : all warps stream the replicas' (count, Q_a) of the tile, lane <-> cell (128-byte coalesced rows), 16 replicas per
//            warp and round -- every load is independent of every other; NW = 32 takes 512 replicas in ONE round trip;
//   reduce : the only order-dependent quantity is the float32 sum of the visitors' terms in replica order (part of the
//            semantics: bit-exact vs oracle/loop.py): warp 0 carries that chain, ONE dependent fadd per replica.  A replica that
//            did not visit the cell contributes (q - q_snap) * 0 = +-0, which never changes the sum (it starts at +0 and
//            +0 + -0 = +0), so the chain needs no test.  Visit counts, the number of visitors and the single visitor's value
//            do not depend on the order: every warp accumulates them for the replicas it loads, combined once at the end;
//   write  : the merged value goes to every replica (all warps, coalesced) and to the snapshot.
// Only the live rows (levels 0..working step) can differ from the snapshot.  One extra CTA per group (blockIdx.x == number
// of tiles) pools the success windows and arms the promotion, beside the tiles instead of behind one of them.
constexpr int MERGE_PER_WARP = 16;
template <int NW>
__global__ void __launch_bounds__(NW * 32) replica_merge_kernel(uint32_t* tables, uint32_t* snap, dqlb200_population_state* ps,
                                                               int R, int pooled_promote, long long max_episodes) {
  constexpr int CHUNK = NW * MERGE_PER_WARP;
  extern __shared__ __align__(16) unsigned char merge_smem[];
  float (*s_term)[32] = reinterpret_cast<float (*)[32]>(merge_smem);                                   // [CHUNK][32]
  uint32_t (*s_ptot)[32] = reinterpret_cast<uint32_t (*)[32]>(merge_smem + (size_t)CHUNK * 32 * 4);      // [NW][32]
  uint32_t (*s_pvis)[32] = s_ptot + NW;
  uint32_t (*s_pq)[32] = s_pvis + NW;
  __shared__ uint32_t s_qnew[32], s_cnew[32], s_vis[32];
  const int g = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_tiles = (int)gridDim.x - 1;
  if ((int)blockIdx.x < n_tiles) {
    const int c = blockIdx.x * 32 + lane;
    uint32_t* sg = snap + (size_t)g * 3 * CELLS;
    const uint32_t* tg = tables + (size_t)g * R * 3 * CELLS;
    const int live = (ps[g * R].working_step + 1) * DQLB200_CELLS_PER_LEVEL;
    if ((int)blockIdx.x * 32 >= live) return;           // block-uniform
    const bool in = c < live;
    const float q_snap = in ? __uint_as_float(sg[c]) : 0.0f;
    const uint32_t cnt_snap = in ? sg[2 * CELLS + c] : 0u;
    float num = 0.0f, my_q = q_snap;
    uint32_t my_tot = 0, my_vis = 0;
    for (int r0 = 0; r0 < R; r0 += CHUNK) {
      const int n = min(CHUNK, R - r0);
      {   // MERGE_PER_WARP replicas per warp: all their loads are issued before the first one is consumed
        uint32_t cv[MERGE_PER_WARP], qv[MERGE_PER_WARP];
#pragma unroll
        for (int i = 0; i < MERGE_PER_WARP; ++i) {
          const int j = warp + NW * i;
          const uint32_t* tr = tg + (size_t)(r0 + min(j, n - 1)) * 3 * CELLS;
          cv[i] = in ? __ldcg(tr + 2 * CELLS + c) : cnt_snap;
          qv[i] = in ? __ldcg(tr + c) : 0u;
        }
#pragma unroll
        for (int i = 0; i < MERGE_PER_WARP; ++i) {
          const int j = warp + NW * i;
          if (j < n) {
            const uint32_t dc = cv[i] - cnt_snap;
            const float q_r = __uint_as_float(qv[i]);
            s_term[j][lane] = fmul(fsub(q_r, q_snap), __uint2float_rn(dc));
            my_tot += dc;
            my_vis += dc ? 1u : 0u;
            my_q = dc ? q_r : my_q;
          }
        }
      }
      __syncthreads();
      if (warp == 0) {
        // the chain is the critical path of the merge: the next eight terms are loaded before the eight dependent additions of
        // the current ones are issued (otherwise every batch exposes a shared-memory latency between 4-cycle additions)
        int j = 0;
        if (n >= 8) {
          float a0 = s_term[0][lane], a1 = s_term[1][lane], a2 = s_term[2][lane], a3 = s_term[3][lane], a4 = s_term[4][lane], a5 = s_term[5][lane],
                a6 = s_term[6][lane], a7 = s_term[7][lane];
          for (j = 8; j + 8 <= n; j += 8) {
            const float b0 = s_term[j][lane], b1 = s_term[j + 1][lane], b2 = s_term[j + 2][lane], b3 = s_term[j + 3][lane], b4 = s_term[j + 4][lane],
                        b5 = s_term[j + 5][lane], b6 = s_term[j + 6][lane], b7 = s_term[j + 7][lane];
            num = fadd(fadd(fadd(fadd(fadd(fadd(fadd(fadd(num, a0), a1), a2), a3), a4), a5), a6), a7);
            a0 = b0; a1 = b1; a2 = b2; a3 = b3; a4 = b4; a5 = b5; a6 = b6; a7 = b7;
          }
          num = fadd(fadd(fadd(fadd(fadd(fadd(fadd(fadd(num, a0), a1), a2), a3), a4), a5), a6), a7);
        }
        for (; j < n; ++j) num = fadd(num, s_term[j][lane]);
      }
      if (r0 + CHUNK < R) __syncthreads();      // the next round overwrites the terms
    }
    s_ptot[warp][lane] = my_tot;
    s_pvis[warp][lane] = my_vis;
    s_pq[warp][lane] = __float_as_uint(my_q);
    __syncthreads();
    if (warp == 0) {
      uint32_t tot = 0;
      int visitors = 0;
      float q_single = q_snap;
#pragma unroll
      for (int w8 = 0; w8 < NW; ++w8) {
        tot += s_ptot[w8][lane];
        visitors += (int)s_pvis[w8][lane];
        if (s_pvis[w8][lane]) q_single = __uint_as_float(s_pq[w8][lane]);      // used only when there is exactly one visitor
      }
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
      for (int r = warp; r < R; r += NW) {
        uint32_t* tr = tables + (size_t)(g * R + r) * 3 * CELLS;
        tr[c] = qn;
        tr[2 * CELLS + c] = cn;
      }
    }
  } else {          // pooled trainer counters of the group: block-wide reduction over the R replicas
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
    // warp totals first: a 64-bit shared-memory atomicAdd is a compare-and-swap loop, and R threads adding to ONE address
    // serialise (measured: 47 ns per replica, 24 of the 31 us of a merge of 512 replicas)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      successes += __shfl_xor_sync(FULL, successes, off);
      episodes += __shfl_xor_sync(FULL, episodes, off);
      dead |= __shfl_xor_sync(FULL, dead, off);
    }
    if (lane == 0) {
      if (successes) atomicAdd(&s_succ, successes);
      if (episodes) atomicAdd(&s_eps, episodes);
      if (dead) atomicOr(&s_dead, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0)
      s_pending = (s_dead || pooled_promote <= 0) ? 0 : ((long long)s_succ >= pooled_promote ? 1 : ((long long)s_eps >= max_episodes ? 2 : 0));
    __syncthreads();
    const int pending = s_pending;
    if (pending)
      for (int r = threadIdx.x; r < R; r += blockDim.x) ps[g * R + r].pending_advance = pending;
  }
}
inline size_t merge_smem_bytes(int nw) { return (size_t)nw * MERGE_PER_WARP * 32 * 4 + (size_t)3 * nw * 32 * 4; }

// -------------------------------------------------------------------------------------------------
// Un-fused agent entry points on the float32 device tables (the batched DoubleQLearningAgent.guess / update,
// PKG/double_q_learning.py:91-146, with the trainer's schedules PKG/trainer.py:88-126): same arithmetic, same draws and the
// same ordering ("S1") as the select / commit phases of train_kernel -- tests hold a loop of agent_select -> env_step ->
// agent_update bit-identical to the fused kernel.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) agent_select_kernel(const __grid_constant__ KC kc, EnvPtrs env, const uint32_t* __restrict__ tables,
                                                           const dqlb200_population_params* pop_params, const uint32_t* __restrict__ eps_threshold,
                                                           int w, uint32_t t, uint8_t* out_action, uint16_t* out_state) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_total = (long long)kc.n_populations * kc.envs_per_population;
  if (i >= n_total) return;
  const int pop = (int)(i / kc.envs_per_population);
  const uint32_t env_i = (uint32_t)(i % kc.envs_per_population);
  const dqlb200_population_params pp = pop_params[pop];
  const uint4 Cw = *reinterpret_cast<const uint4*>(env_addr(env, (size_t)i) + 1024);
  const uint32_t sid = Cw.x & ((1u << SID_BITS) - 1u), episode = Cw.y;
  const float* qa = reinterpret_cast<const float*>(tables + (size_t)pop * 3 * CELLS);
  const float* qb = qa + CELLS;
  const float p0 = fmul(fadd(qa[sid * 3 + 0], qb[sid * 3 + 0]), 0.5f);
  const float p1 = fmul(fadd(qa[sid * 3 + 1], qb[sid * 3 + 1]), 0.5f);
  const float p2 = fmul(fadd(qa[sid * 3 + 2], qb[sid * 3 + 2]), 0.5f);
  int a = 0;
  float best = p0;
  if (p1 > best) { best = p1; a = 1; }
  if (p2 > best) { a = 2; }
  if (w == 0) {        // exploration_rate is 0 for every later working step (PKG/trainer.py:112-126)
    const uint32_t thr = eps_threshold[min(episode, (uint32_t)(DQLB200_EPS_LUT - 1))];
    const uint4 d = philox4x32_10(make_uint4(env_i, t, PURPOSE_STEP, pp.population_id), pp.seed_lo, pp.seed_hi);
    if ((d.x >> 8) < thr) a = (int)__umulhi(d.y, 3u);
  }
  out_action[i] = (uint8_t)a;
  if (out_state) out_state[i] = (uint16_t)sid;
}

// One warp per population: the updates of its envs are applied in env-index order, 32 at a time (same-cell groups inside a
// chunk by __match_any_sync, lane order), against the bootstrap values of the tables as they were when the call started.
__global__ void __launch_bounds__(32) agent_update_kernel(const __grid_constant__ KC kc, uint32_t* tables, const dqlb200_population_params* pop_params,
                                                          const float* __restrict__ alpha_luts, const uint16_t* __restrict__ state,
                                                          const uint8_t* __restrict__ action, const uint16_t* __restrict__ next_state,
                                                          const double* __restrict__ reward) {
  __shared__ float qa[CELLS], qmax[STATES];
  __shared__ uint32_t cnt[CELLS];
  const int pop = blockIdx.x, lane = threadIdx.x, n_p = kc.envs_per_population;
  uint32_t* gt = tables + (size_t)pop * 3 * CELLS;
  const float* __restrict__ alpha_lut = alpha_luts + (size_t)pop_params[pop].alpha_lut * DQLB200_ALPHA_LUT;
  for (int i = lane; i < CELLS; i += 32) { qa[i] = __uint_as_float(gt[i]); cnt[i] = gt[2 * CELLS + i]; }
  __syncwarp();
  for (int st = lane; st < STATES; st += 32) qmax[st] = fmaxf(fmaxf(qa[st * 3], qa[st * 3 + 1]), qa[st * 3 + 2]);
  __syncwarp();
  const size_t base = (size_t)pop * n_p;
  for (int e0 = 0; e0 < n_p; e0 += 32) {
    const bool valid = e0 + lane < n_p;
    const size_t gi = base + (size_t)min(e0 + lane, n_p - 1);
    const uint32_t s = state[gi], s2 = next_state[gi], a = action[gi];
    const uint32_t cell = s * 3u + a;
    const float changed = (((s / 63u) % 3u) != ((s2 / 63u) % 3u)) ? 1.0f : 0.0f;
    const float target = fadd((float)reward[gi], fmul(fmul(kc.gamma, qmax[s2]), changed));
    const uint32_t peers = __match_any_sync(FULL, valid ? cell : (0x80000000u | (uint32_t)lane));
    const int rank = __popc(peers & ((1u << lane) - 1u));
    float q = valid ? qa[cell] : 0.0f;
    const uint32_t c0 = valid ? cnt[cell] : 0u;
    const float alpha = alpha_lut[min(c0 + (uint32_t)rank, (uint32_t)(DQLB200_ALPHA_LUT - 1))];      // R11: pre-increment count
    uint32_t rem = valid ? peers : 0u;
    while (__any_sync(FULL, rem != 0u)) {
      const int src = rem ? (__ffs(rem) - 1) : lane;
      const float a_j = __shfl_sync(FULL, alpha, src), t_j = __shfl_sync(FULL, target, src);
      if (rem) q = fadd(q, fmul(a_j, fsub(t_j, q)));
      rem &= rem - 1u;
    }
    if (valid && rank == 0) { qa[cell] = q; cnt[cell] = c0 + (uint32_t)__popc(peers); }
    __syncwarp();
  }
  for (int i = lane; i < CELLS; i += 32) { gt[i] = __float_as_uint(qa[i]); gt[2 * CELLS + i] = cnt[i]; }
}

// Measurement aid: what a launch of train_kernel's shape costs before it does anything (CTA launch with the same grid, block,
// dynamic shared memory and by-value parameter block).
__global__ void launch_floor_kernel(const __grid_constant__ KC kc, int* sink) {
  extern __shared__ __align__(16) unsigned char floor_smem[];
  if (kc.n_populations < 0) { floor_smem[threadIdx.x] = 1; sink[0] = floor_smem[0]; }     // never taken: keeps the operands alive
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
