// env_state.cuh -- the 48-byte environment state (tiles of 32 envs x three 16-byte vectors): pack/unpack, reset law, asynchronous prefetch
// Part of libdqlb200 (see dqlb200.cu for the kernel inventory and the C-ABI).
#pragma once
#include "dqlb200_device.cuh"

namespace dql {


constexpr int CELLS = DQLB200_MAX_CELLS;
constexpr uint32_t FULL = 0xFFFFFFFFu;

// 16-byte read through a shared-state-space address
__device__ __forceinline__ uint4 lds128(unsigned addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// env-state word C.x layout
constexpr uint32_t SID_BITS = 10, STEP_SHIFT = 10, STEP_BITS = 9, CC_SHIFT = 19, CC_BITS = 5;
constexpr uint32_t STICKY_BIT = 1u << 24, FRESH_BIT = 1u << 25, BP_SHIFT = 26;   // bits 26-27: position bin of `sid`

struct Env {
  Body b;
  uint32_t sp_idx;     // pitch set-point as an index into dqlb200_config.setpoint_value; NOT cleared by an episode reset while
                       // `fresh` (keeps the shaping potential, quirk Q11)
  float prev_rel_p, prev_rel_v;
  uint32_t sid, bp, step_count, curriculum_check;     // bp = position bin of sid ((sid / 63) % 3, kept to avoid the division)
  bool sticky_success, fresh;
  uint32_t episode;
  double cum_reward;
};

// Layout of the env state in HBM: [population][tile][3][32 envs][16 B] -- a tile is the 48 bytes of 32 consecutive envs of ONE
// population as three 512-byte runs (vectors A, B, C), 1536 contiguous bytes: exactly what one warp reads and writes per slot,
// with the vectors of an env at constant offsets (+512, +1024) from its A vector.  A population occupies tiles_per_pop =
// ceil(envs_per_population / 32) tiles (the last one padded), so a population's state -- and any range of populations -- is one
// contiguous block (host-buffer calls copy it with one transfer).
constexpr size_t ENV_TILE_BYTES = 3 * 32 * 16;
struct EnvPtrs {
  unsigned char* base;
  int n_p, tiles_per_pop;      // envs per population, tiles per population
  uint4* d;      // acceleration-estimator state (accel_mode != 0 only, else null): {x, P, v_ref, n}
  uint4* e;      // second-order model state (dynamics_model != 0 only, else null): [2][n] {omega, z, v_z, integral}, {e1, f1, f2, f3}
  size_t n;      // envs in the SoA (stride of `e`)
  // the set-point tables of the configuration (device copies inside dqlb200_config, see include/dqlb200.h)
  const double* sp_value;       // [n_setpoints]
  const uint2* sp_next;         // [n_setpoints][3] {next index, float32 bits of its value}
  const double* sp_rtheta;      // [2][DQLB200_MAX_SETPOINTS][3]
  int sp_zero, n_sp;
};
// address of vector A of env `env` of population `pop` (B at +512, C at +1024)
__device__ __forceinline__ unsigned char* env_addr(const EnvPtrs& p, int pop, int env) {
  return p.base + ((size_t)pop * p.tiles_per_pop + (size_t)(env >> 5)) * ENV_TILE_BYTES + (size_t)(env & 31) * 16;
}
// the same from the global env index i = pop * envs_per_population + env (the index of traces, action arrays and extension state)
__device__ __forceinline__ unsigned char* env_addr(const EnvPtrs& p, size_t i) {
  const int pop = (int)(i / (size_t)p.n_p);
  return env_addr(p, pop, (int)(i - (size_t)pop * p.n_p));
}
__device__ __forceinline__ Ext ext_load(const EnvPtrs& p, size_t i) {
  const uint4 u = p.e[i], v = p.e[p.n + i];
  return Ext{__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w),
             __uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
}
__device__ __forceinline__ void ext_store(const EnvPtrs& p, size_t i, const Ext& x) {
  p.e[i] = make_uint4(__float_as_uint(x.omega), __float_as_uint(x.z), __float_as_uint(x.v_z), __float_as_uint(x.integ));
  p.e[p.n + i] = make_uint4(__float_as_uint(x.e1), __float_as_uint(x.f1), __float_as_uint(x.f2), __float_as_uint(x.f3));
}
// a simulator that has been hovering: the integral holds the hover thrust (m g / Ki), the filter memory is empty
template <class KT>
__device__ __forceinline__ Ext ext_initial(const KT& kc) { return Ext{0.0f, kc.z_init, 0.0f, kc.pid_i0, 0.0f, 0.0f, 0.0f, 0.0f}; }
__device__ __forceinline__ Kf kf_load(const EnvPtrs& p, size_t i) {
  const uint4 v = p.d[i];
  return Kf{__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), v.w};
}
__device__ __forceinline__ void kf_store(const EnvPtrs& p, size_t i, const Kf& f) {
  p.d[i] = make_uint4(__float_as_uint(f.x), __float_as_uint(f.P), __float_as_uint(f.v_ref), f.n);
}
// The extension state of an env through three staging slots of its thread ([3][threads] x 16 B behind the env tiles; extended
// kernel variant, stage_addr = shared-space address of the thread's first slot there): issued next to the env
// prefetch, picked up one slot later -- a plain load at the head of a slot would expose an L2 round trip per slot.
__device__ __forceinline__ void ext_prefetch_async(const EnvPtrs& p, size_t i, unsigned stage_addr, int nt, bool filt, bool so) {
  if (filt) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage_addr), "l"(p.d + i) : "memory");
  if (so) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage_addr + 16u * nt), "l"(p.e + i) : "memory");
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(stage_addr + 32u * nt), "l"(p.e + p.n + i) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ Kf kf_take(unsigned stage_addr, int nt) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const uint4 v = lds128(stage_addr);
  return Kf{__uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), v.w};
}
__device__ __forceinline__ Ext ext_take(unsigned stage_addr, int nt) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const uint4 u = lds128(stage_addr + 16u * nt), v = lds128(stage_addr + 32u * nt);
  return Ext{__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w),
             __uint_as_float(v.x), __uint_as_float(v.y), __uint_as_float(v.z), __uint_as_float(v.w)};
}
__device__ __forceinline__ Kf kf_initial() { return Kf{0.0f, 1.0f, 0.0f, 0u}; }      // PKG/filters.py:15-16

struct EnvRaw {
  float4 A;
  uint4 B, C;
};
// The env state streams: it is read once and written once per global step and never re-read through L1, so every access is
// L2-only (ld/st.global.cg).  What stays in the small L1 next to 228 KB of shared memory are the look-up tables all CTAs of an SM
// share (epsilon thresholds, learning rates, set-point rewards): ncu showed a 31 % L1 hit rate for them with default stores.
__device__ __forceinline__ EnvRaw env_fetch(const unsigned char* pa) {
  EnvRaw r;
  r.A = __ldcg(reinterpret_cast<const float4*>(pa));
  r.B = __ldcg(reinterpret_cast<const uint4*>(pa + 512));
  r.C = __ldcg(reinterpret_cast<const uint4*>(pa + 1024));
  return r;
}
// Asynchronous prefetch of a warp's next tile -- the 48-byte states of 32 consecutive envs, ENV_TILE_BYTES contiguous bytes in
// HBM -- into the warp's staging tile in shared memory by ONE bulk copy of the copy engine (cp.async.bulk, SASS UBLKCP), issued
// by an elected lane and completed on the warp's mbarrier (complete_tx::bytes); the lanes pick their three vectors up one slot
// later.  It replaces three 16-byte cp.async per thread plus their address arithmetic and group bookkeeping, holds no registers
// while in flight and cannot be consumed early by the scheduler's copies (a register prefetch was).
//   mbar       shared-space address of the warp's mbarrier (8 bytes, initialised with count 1 by tile_mbar_init)
//   stage_tile shared-space address of the warp's staging tile ([3][32] x 16 B, the layout of a tile in HBM)
__device__ __forceinline__ void tile_mbar_init(unsigned mbar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the initialised barrier is seen by the copy engine
}
__device__ __forceinline__ void tile_prefetch_bulk(const unsigned char* tile, unsigned stage_tile, unsigned mbar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "n"((int)ENV_TILE_BYTES) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(stage_tile), "l"(tile),
               "n"((int)ENV_TILE_BYTES), "r"(mbar)
               : "memory");
}
// all lanes: wait for the phase `parity` of the warp's mbarrier (the bytes of the tile have landed)
__device__ __forceinline__ void tile_wait(unsigned mbar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TILE_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TILE_DONE;\n"
      "bra TILE_WAIT;\n"
      "TILE_DONE:\n"
      "}\n" ::"r"(mbar),
      "r"(parity)
      : "memory");
}
// this lane's env of the staged tile (stage_lane = stage_tile + lane * 16)
__device__ __forceinline__ EnvRaw tile_take(unsigned stage_lane) {
  EnvRaw r;
  const uint4 a = lds128(stage_lane);
  r.A = make_float4(__uint_as_float(a.x), __uint_as_float(a.y), __uint_as_float(a.z), __uint_as_float(a.w));
  r.B = lds128(stage_lane + 512u);
  r.C = lds128(stage_lane + 1024u);
  return r;
}

__device__ __forceinline__ void env_unpack(const EnvRaw& r, Env& e) {
  const float4 A = r.A;
  const uint4 B = r.B;
  const uint4 Cw = r.C;
  e.b.x_d = A.x; e.b.v_d = A.y; e.b.theta = A.z; e.b.phase = __float_as_uint(A.w); e.b.a_d = 0.0f;
  e.sp_idx = B.x;
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
__device__ __forceinline__ void env_store(unsigned char* pa, const Env& e) {
  __stcg(reinterpret_cast<float4*>(pa), make_float4(e.b.x_d, e.b.v_d, e.b.theta, __uint_as_float(e.b.phase)));
  __stcg(reinterpret_cast<uint4*>(pa + 512), make_uint4(e.sp_idx, 0u, __float_as_uint(e.prev_rel_p), __float_as_uint(e.prev_rel_v)));
  const uint32_t packed = e.sid | (e.step_count << STEP_SHIFT) | (e.curriculum_check << CC_SHIFT) |
                          (e.sticky_success ? STICKY_BIT : 0u) | (e.fresh ? FRESH_BIT : 0u) | (e.bp << BP_SHIFT);
  __stcg(reinterpret_cast<uint4*>(pa + 1024), make_uint4(packed, e.episode, (uint32_t)__double2loint(e.cum_reward),
                                                          (uint32_t)__double2hiint(e.cum_reward)));
}
// by global env index (kernels off the hot path)
__device__ __forceinline__ void env_load(const EnvPtrs& p, size_t i, Env& e) { env_unpack(env_fetch(env_addr(p, i)), e); }
__device__ __forceinline__ void env_store(const EnvPtrs& p, size_t i, const Env& e) { env_store(env_addr(p, i), e); }

// R1 + R8: new episode.  `fresh_mdp` additionally clears what only a NEW TrainingMdp clears
// (shaping potentials, PKG/trainer.py:176 + quirk Q11) and the per-step episode index.
template <class KT, class AC>
__device__ __forceinline__ void env_reset(const KT& kc, const dqlb200_population_params& pp,
                                          const dqlb200_cuts& cuts, const AC& angle_cut, Env& e,
                                          uint32_t env_index, uint32_t birth, int w, bool fresh_mdp, uint32_t sp_zero, Kf* kf = nullptr, Ext* ext = nullptr) {
  const uint4 d = philox4x32_10(make_uint4(env_index, birth, PURPOSE_RESET, pp.population_id), pp.seed_lo, pp.seed_hi);
  Obs o = dyn_reset(kc, pp, e.b, d, /*normal_init=*/w == 0, /*simulation=*/false, kc.dz_train, kf, ext);
  if (kc.noise_enabled) {
    const uint4 dn = philox4x32_10(make_uint4(env_index, birth, PURPOSE_RESET_NOISE, pp.population_id), pp.seed_lo, pp.seed_hi);
    add_observation_noise(kc, o, dn.x, dn.y);
  }
  const DState ds0 = discretise_cuts(cuts, angle_cut, o);
  e.sid = (uint32_t)ds0.id();
  e.bp = (uint32_t)ds0.bp;
  e.step_count = 0;
  e.curriculum_check = 0;
  e.sticky_success = false;
  e.fresh = true;
  e.cum_reward = 0.0;
  if (fresh_mdp) {
    e.sp_idx = sp_zero;
    e.prev_rel_p = 0.0f;
    e.prev_rel_v = 0.0f;
    e.episode = 0;
  }
}


}  // namespace dql
