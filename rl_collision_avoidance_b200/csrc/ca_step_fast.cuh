// ca_step_fast.cuh — the step kernels specialised on the number of agent slots per world (kA).
//
// Same algorithm, data layout and numerics as ca_world_kernel<true> (ca_kernels.cuh; the step body itself —
// step_take_action / step_reward_done — is literally shared), but with kA a compile-time constant every per-other loop
// is fully unrolled: the lane's int32 sort keys, p_orth and centre distances of its (kA-1) potential neighbours live
// in registers (fast_pair_pass / fast_write_obs_row), the neighbour order is a pairwise rank (each unordered pair of
// keys is compared once) and nothing but the finished observation rows goes through shared memory.
//   ca_step_kernel         one-shot: one CTA = 4 warps = 4 chunks of floor(32/kA) worlds, one TMA bulk store per warp
//   ca_step_stream_kernel  (ca_step_stream.cuh) persistent: warps pull chunks from an atomic counter, the next chunk's
//                          state block is in flight (TMA bulk load + mbarrier) while the current one is computed
// Used for closest_first / closest_last sorting; time_to_impact sorting, reset, and agent counts without an
// instantiation run on the generic kernel.
#pragma once
#include <limits.h>

#include "ca_kernels.cuh"

namespace ca {

template <int kA>
struct OthersLite {
  static constexpr int kN = kA > 1 ? kA - 1 : 1;
  int key[kN];      // rint(100 * dist_2_other); INT_MAX for an absent / unobserved other
  double po[kN];    // p_orth (float64: it breaks ties between equal keys)
  // observation-only float32 values, final already: p_parallel and the boundary distance (the other's radius and
  // velocity are fetched by float32 shuffles when the row is written)
  float pprl[kN], d2o[kN];
};

// Number of per-other iterations the lanes of this warp need: kA - 1, or for the larger specialisations (ragged
// batches: BASELINE configs[3]) the largest agent count among the warp's worlds minus one — warp-uniform, so the
// unrolled loops below skip whole iterations without divergence.
template <int kA>
__device__ __forceinline__ int others_bound(int n) {
  if (kA < 6) return kA - 1;
  return __reduce_max_sync(kFull, n) - 1;
}

// kGen selects the general variant (finite SENSING_HORIZON, neighbour-index output for parity tests, M < kA - 1
// clipping); the production instantiation (kGen = false) carries neither their instructions nor their registers.
// First loop of the sensor (OtherAgentsStatesSensor.sense :72-103) fused with _check_for_collisions (:370-409).
// kAll: every world of the warp has all kA agents (nm1 == kA - 1): the instantiation without the per-iteration bound
// checks (and without the register copies their control-flow joins cost) — TrainPhase1 / all-present batches.
template <int kA, bool kCollide, bool kGen, bool kAll>
__device__ __forceinline__ void fast_pair_pass(const Params& p, const Agent& a, const Ego& e, bool valid, int n, int nm1,
                                               int i, int base, OthersLite<kA>& o, bool& coll, double& nearest) {
  coll = false;
  nearest = INFINITY;
  const bool horizon = kGen && isfinite(p.sensing_horizon);
#pragma unroll
  for (int k = 0; k < kA - 1; ++k) {
    o.key[k] = INT_MAX;
    o.po[k] = 0.0;
    o.pprl[k] = 0.f; o.d2o[k] = 0.f;
    if (!kAll && k >= nm1) continue;   // warp-uniform
    const int j = k + (k >= i ? 1 : 0);
    const int src = (base + j) & 31;
    const double xj = shfl_d(a.px, src), yj = shfl_d(a.py, src), rj = shfl_d(a.rad, src);
    const bool live = valid && j < n;
    const double rx = xj - a.px, ry = yj - a.py;
    const double d = sqrt(rx * rx + ry * ry);  // l2norm (util.py:8-12)
    if (kCollide) {
      const double R = a.rad + rj;
      const double gap = d - R;
      coll = coll || (live && d <= R);
      if (live && j > i && gap < nearest) nearest = gap;  // only the lower index is updated (:393)
    }
    const bool seen = live && !(horizon && d > p.sensing_horizon);
    const double d2o = d - a.rad - rj;
    if (seen) o.key[k] = __double2int_rn(d2o * 100.0);
    o.po[k] = dot2(rx, ry, -e.pry, e.prx);
    o.pprl[k] = obs_pprl(rx, ry, e);
    o.d2o[k] = (float)d2o;
  }
}

__device__ __forceinline__ bool key_first(int q1, double p1, int q2, double p2) {
  return (q1 < q2) || (q1 == q2 && p1 <= p2);
}

// slot[k] += number of keys among the first kUse that sort before key k (stable: equal keys keep index order); each
// unordered pair is compared once.  b(k1, k2) = key_first(k1, k2) for k1 < k2; with L[k] = sum_{k1<k} b(k1, k) and
// W[k] = sum_{k2>k} b(k, k2) the rank is L[k] + (kUse - 1 - k) - W[k].  The comparison is written as three chained
// predicate instructions (setp.eq -> setp.le.and.f64 -> setp.lt.or) and two predicated adds per pair; the compiler's own
// rendering of the boolean expression plus two selects cost ~11 instructions per pair (20 % of the 10-agent kernel).
template <int kN, int kUse>
__device__ __forceinline__ void rank_pairs(const int* key, const double* po, int* slot) {
  int L[kN], Wn[kN];
#pragma unroll
  for (int k = 0; k < kN; ++k) { L[k] = 0; Wn[k] = 0; }
#pragma unroll
  for (int k1 = 0; k1 < kUse; ++k1)
#pragma unroll
    for (int k2 = k1 + 1; k2 < kUse; ++k2) {
      asm("{\n\t"
          ".reg .pred peq, ple, pb;\n\t"
          "setp.eq.s32 peq, %2, %3;\n\t"
          "setp.le.and.f64 ple, %4, %5, peq;\n\t"
          "setp.lt.or.s32 pb, %2, %3, ple;\n\t"
          "@pb add.s32 %0, %0, 1;\n\t"
          "@pb add.s32 %1, %1, 1;\n\t"
          "}"
          : "+r"(L[k2]), "+r"(Wn[k1])
          : "r"(key[k1]), "r"(key[k2]), "d"(po[k1]), "d"(po[k2]));
    }
#pragma unroll
  for (int k = 0; k < kUse; ++k) slot[k] += L[k] - Wn[k] + (kUse - 1 - k);
}
// the same with the (warp-uniform) number of keys in use known only at run time: one fully unrolled body per count
template <int kN, int kUse>
__device__ __forceinline__ void rank_dispatch(int nm1, const int* key, const double* po, int* slot) {
  if constexpr (kUse >= 2) {
    if (nm1 >= kUse) rank_pairs<kN, kUse>(key, po, slot);
    else rank_dispatch<kN, kUse - 1>(nm1, key, po, slot);
  }
}

// All lanes of the warp clear `nfloats` floats at t (the widest store the alignment of t and nfloats allows).
__device__ __forceinline__ void zero_warp_tile(float* t, int nfloats, int lane) {
  const uintptr_t addr = reinterpret_cast<uintptr_t>(t);
  if (((addr | (uintptr_t)(nfloats * 4)) & 15u) == 0) {
    float4* t4 = reinterpret_cast<float4*>(t);
    for (int q = lane; q < nfloats / 4; q += 32) t4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  } else if (((addr | (uintptr_t)(nfloats * 4)) & 7u) == 0) {
    float2* t2 = reinterpret_cast<float2*>(t);
    for (int q = lane; q < nfloats / 2; q += 32) t2[q] = make_float2(0.f, 0.f);
  } else {
    for (int q = lane; q < nfloats; q += 32) t[q] = 0.f;
  }
}

// Second loop of the sensor (:105-144) + the dense row of GCA/envs/wrappers.py:130-139: rank the others, (general
// variant: clip to M, closest_last order), write the lane's observation row into the tile.
template <int kA, bool kGen, bool kAll>
__device__ __forceinline__ void fast_write_obs_row(const Params& p, const Agent& a, const Ego& e, bool world_ok,
                                                   bool valid, int nm1, int i, int base, const OthersLite<kA>& o,
                                                   float* row, int32_t* sidx_row_in, float* wt, int wt_floats) {
  int32_t* const sidx_row = kGen ? sidx_row_in : nullptr;
  constexpr int kN = kA - 1;
  constexpr int kNN = kN > 0 ? kN : 1;
  const int M = kGen ? p.M : kN;
  int count = 0;
  int key[kNN];
#pragma unroll
  for (int k = 0; k < kN; ++k) {
    key[k] = o.key[k];
    count += (key[k] != INT_MAX) ? 1 : 0;
  }
  if (kGen && count > M) {   // first sort + clip to the M closest
    int rank1[kNN];
#pragma unroll
    for (int k = 0; k < kN; ++k) rank1[k] = 0;
    rank_pairs<kNN, kN>(o.key, o.po, rank1);
#pragma unroll
    for (int k = 0; k < kN; ++k)
      if (rank1[k] >= M) key[k] = INT_MAX;
    count = M;
  }
  if (kGen && p.sort_method == CA_SORT_CLOSEST_LAST) {
#pragma unroll
    for (int k = 0; k < kN; ++k)
      if (key[k] != INT_MAX) key[k] = -key[k];
  }
  int slot[kNN];
#pragma unroll
  for (int k = 0; k < kN; ++k) slot[k] = 0;
  if (kAll) rank_pairs<kNN, kN>(key, o.po, slot);
  else rank_dispatch<kNN, kN>(nm1, key, o.po, slot);
  // Rows of absent agents and the unused tail of short rows are zeros (wrappers.py:115-139).  A lane clearing its own
  // row serialises against the lanes that have values to write, so when any row of the warp needs zeros all 32 lanes
  // clear the warp's whole tile first with wide stores (ragged batches: ~20 instructions instead of ~300).
  const bool need_zeros = world_ok && (!valid || count < M);
  if (__any_sync(kFull, need_zeros)) {
    zero_warp_tile(wt, wt_floats, threadIdx.x & 31);
    __syncwarp();
  }
  if (world_ok) {
    if (valid) {
      row[0] = (a.policy == CA_POLICY_LEARNING_GA3C || a.policy == CA_POLICY_LEARNING) ? 1.f : 0.f;
      row[1] = (float)count;
      row[2] = (float)e.dist;
      row[3] = e.hego;
      row[4] = (float)a.ps;
      row[5] = (float)a.rad;
      if (sidx_row) for (int k = count; k < M; ++k) sidx_row[k] = -1;
    } else {
      if (sidx_row) for (int k = 0; k < M; ++k) sidx_row[k] = -1;
    }
  }
  // second loop of the sensor: the others' radius and velocity still have to be fetched, as float32
  const float prxf = (float)e.prx, pryf = (float)e.pry, raf = (float)a.rad;
  const float vxf = (float)a.vx, vyf = (float)a.vy;
#pragma unroll
  for (int k = 0; k < kN; ++k) {
    if (!kAll && k >= nm1) continue;   // warp-uniform
    const int j = k + (k >= i ? 1 : 0);
    const int src = (base + j) & 31;
    const float vxj = __shfl_sync(kFull, vxf, src), vyj = __shfl_sync(kFull, vyf, src);
    const float rrj = __shfl_sync(kFull, raf, src);
    if (valid && key[k] != INT_MAX) {
      float* s = row + CA_OBS_HOST_LEN + CA_OBS_OTHER_LEN * slot[k];
      s[0] = o.pprl[k];
      s[1] = (float)o.po[k];
      obs_vel(vxj, vyj, prxf, pryf, s[2], s[3]);
      s[4] = rrj;
      s[5] = raf + rrj;
      s[6] = o.d2o[k];
      if (sidx_row) sidx_row[slot[k]] = j;
    }
  }
}

// Dispatch on "all agents present" (always true for the small specialisations, whose loops are not bounded): the
// launch-uniform host hint AND the warp's own agent counts, so a stale hint costs speed, never correctness.
template <int kA, bool kCollide, bool kGen>
__device__ __forceinline__ void fast_pair_pass_any(const Params& p, const Agent& a, const Ego& e, bool valid, int n, int nm1,
                                                   int i, int base, OthersLite<kA>& o, bool& coll, double& nearest) {
  if (kA < 6 || (p.all_present && nm1 == kA - 1)) fast_pair_pass<kA, kCollide, kGen, true>(p, a, e, valid, n, nm1, i, base, o, coll, nearest);
  else fast_pair_pass<kA, kCollide, kGen, false>(p, a, e, valid, n, nm1, i, base, o, coll, nearest);
}
template <int kA, bool kGen>
__device__ __forceinline__ void fast_write_obs_row_any(const Params& p, const Agent& a, const Ego& e, bool world_ok,
                                                       bool valid, int nm1, int i, int base, const OthersLite<kA>& o,
                                                       float* row, int32_t* sidx_row_in, float* wt, int wt_floats) {
  if (kA < 6 || (p.all_present && nm1 == kA - 1))
    fast_write_obs_row<kA, kGen, true>(p, a, e, world_ok, valid, nm1, i, base, o, row, sidx_row_in, wt, wt_floats);
  else
    fast_write_obs_row<kA, kGen, false>(p, a, e, world_ok, valid, nm1, i, base, o, row, sidx_row_in, wt, wt_floats);
}

// Warp-level store of the warp's observation rows (its wpw worlds are contiguous in global memory).  The rows were
// assembled at t0 = (16-byte aligned tile base) + shift floats, where shift is the 16-byte phase of dst, so source and
// destination of the 16-byte-aligned body agree mod 16 for ANY tile size (10 agents x 69 floats included): the body
// leaves as one TMA bulk store (SASS UBLKCP), at most 3 + 3 head / tail floats as scalar stores.  Returns true when a
// bulk store was committed: the caller must cp.async.bulk.wait_group.read before the tile is reused or the CTA exits.
__device__ __forceinline__ int tile_shift(const float* dst) { return (int)((reinterpret_cast<uintptr_t>(dst) >> 2) & 3u); }

__device__ __forceinline__ bool warp_tile_store(const Params& p, float* dst, const float* t0, int shift, int nf, int lane) {
  const int head = (4 - shift) & 3;
  const int body = nf >= head ? ((nf - head) & ~3) : 0;
  if (p.use_bulk_store == 1 && body > 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + head),
                   "r"(smem_u32(t0 + head)), "r"(body * 4)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (lane < head) dst[lane] = t0[lane];
    const int tail = nf - head - body;
    if (lane < tail) dst[head + body + lane] = t0[head + body + lane];
    return true;
  }
  __syncwarp();
  for (int q = lane; q < nf; q += 32) dst[q] = t0[q];
  __syncwarp();
  return false;
}

// Worlds in which no learning agent is still running after the move (goal reached, time budget spent, or done before
// this step) end in this step whatever the collision test says, and with auto-reset reload their snapshot block — a
// cold DRAM read in the middle of the chunk, ~8 % of the chunks of a training batch, and the warps that pay it are the
// tail of the launch.  Pull that block towards L2 now (one TMA bulk prefetch per warp, SASS UBLKPF): the all-pairs pass
// and the reward run while it is in flight.  Episodes ended by a collision in this very step are not caught here.
__device__ __forceinline__ void prefetch_snapshot_if_ending(const Params& p, const Agent& a, bool world_ok, bool valid,
                                                            unsigned gmask, const double* blk0, int lane) {
  if (!p.prefetch_snapshot) return;
  const bool learning = valid && (a.policy == CA_POLICY_LEARNING_GA3C || a.policy == CA_POLICY_LEARNING);
  const unsigned running = __ballot_sync(kFull, learning && !(a.flags & CA_F_DONE_MASK));
  if (__any_sync(kFull, world_ok && (running & gmask) == 0u) && lane == 0)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(blk0), "r"(kBlkReadBytes) : "memory");
}

// The grid runs in about two rounds of resident CTAs: pull the state block and the actions of the chunk that will run in
// this warp's slot one round later into L2 (TMA bulk prefetch), so the second round does not start with another DRAM
// round trip.
__device__ __forceinline__ void prefetch_next_round(const Params& p, int chunk, int wpw, int kA, int lane) {
  const int pc = chunk + p.prefetch_chunks;
  if (lane == 0 && pc * wpw < p.W) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(blk_ptr(p.s, pc)), "r"(kBlkReadBytes) : "memory");
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.actions + (size_t)pc * wpw * kA) : "memory");
  }
}

// bytes of shared memory per warp for the observation tile: rows + 16 bytes of phase room, 16-byte granular
__host__ __device__ __forceinline__ int warp_tile_region(int tile_floats) { return ((tile_floats * 4 + 15) / 16) * 16 + 16; }

template <int kA, int kMinBlocks, bool kGen>
__global__ void __launch_bounds__(kBlock, kMinBlocks) ca_step_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int wpw = (32 / kA) < 16 ? (32 / kA) : 16;
  // the warp index through a shuffle: the compiler then treats everything derived from it (chunk, tile and destination
  // addresses, the operands of the bulk store) as warp-uniform and keeps it on the uniform datapath
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(kFull, tid >> 5, 0);
  const int wl = lane / kA;
  const int i = lane - wl * kA;
  const int base = wl * kA;
  // 32-bit indexing: ca_create refuses W * A >= 2^31
  const int chunk = blockIdx.x * (int)(blockDim.x >> 5) + warp;   // 1, 2 or 4 warps per CTA (ca_step.cu: fast_warps)
  const int first_world_warp = chunk * wpw;
  const int w = first_world_warp + wl;
  const bool world_ok = wl < wpw && w < p.W;
  const unsigned g = world_ok ? (unsigned)w * kA + i : 0u;
  double* const blk = blk_ptr(p.s, chunk);
  const unsigned gmask = (kA >= 32 ? kFull : ((1u << kA) - 1u)) << (base & 31);

  // per-warp tile: rows of the warp's wpw worlds, assembled at the 16-byte phase of their destination
  const int tile_floats = wpw * kA * p.L;
  float* const dst = p.obs + (size_t)first_world_warp * kA * p.L;
  const int shift = tile_shift(dst);
  float* wtile = reinterpret_cast<float*>(smem_raw + warp * warp_tile_region(tile_floats));
  float* row = wtile + shift + (wl * kA + i) * p.L;
  int32_t* sidx_row = (kGen && p.sidx && world_ok) ? p.sidx + (size_t)g * p.M : nullptr;

  CA_STAMP(p, chunk, 0, lane, 0);
  pdl_wait();                // nothing produced by the previous kernel is read above this line
  pdl_launch_dependents();
  CA_STAMP(p, chunk, 1, lane, 0);
  // The state loads do not wait for the agent count: every slot of an existing world is loaded (absent slots hold
  // zeros / stale values and are discarded below), so only ONE DRAM round trip is exposed instead of two.
  Agent a;
  int n = 0, act = 0;
  if (world_ok) {
    n = blk_nag(blk)[wl];
    load_agent<false>(blk, lane, a);
    act = p.actions[g];
  }
  bool valid = world_ok && i < n;
  if (!valid) { zero_agent(a); act = 0; }

  step_take_action<kGen>(p, a, act, g, valid);
  CA_STAMP(p, chunk, 2, lane, a.flags);
  if (!kGen) prefetch_snapshot_if_ending(p, a, world_ok, valid, gmask, blk_ptr(p.s0, chunk), lane);

  Ego e = ego_frame(a.px, a.py, a.gx, a.gy, a.hd);
  OthersLite<kA> o;
  bool coll;
  double nearest;
  const int nm1 = others_bound<kA>(n);
  fast_pair_pass_any<kA, true, kGen>(p, a, e, valid, n, nm1, i, base, o, coll, nearest);

  // by now this round's own load burst has drained and DRAM is idle while the warps compute
  if (p.prefetch_chunks > 0) prefetch_next_round(p, chunk, wpw, kA, lane);

  bool dn, over;
  const float r = step_reward_done<kGen>(p, a, valid, i, coll, nearest, gmask, dn, over);
  CA_STAMP(p, chunk, 3, lane, __float_as_int(r));
  if (world_ok) {
    p.reward[g] = r;
    p.done[g] = dn ? 1 : 0;
    if (i == 0) p.over[w] = over ? 1 : 0;
  }
  const bool do_reset = world_ok && over && p.auto_reset;

  // The state is written back BEFORE the observation rows are assembled: heading, time budget, goal and flags die
  // here instead of occupying registers through the ranking / row code.
  if (!__any_sync(kFull, do_reset)) {
    if (valid) store_agent(blk, lane, a, a.policy == CA_POLICY_STATIC, false);
    fast_write_obs_row_any<kA, kGen>(p, a, e, world_ok, valid, nm1, i, base, o, row, sidx_row, wtile, (tile_floats + 7) & ~3);
  } else {
    // DummyVecEnv semantics: worlds that finished reload their injected initial state and observe again;
    // the other worlds of the warp observe their post-step state.
    if (do_reset) {
      double* const blk0 = blk_ptr(p.s0, chunk);
      n = blk_nag(blk0)[wl];
      valid = i < n;
      if (i == 0) { blk_nag(blk)[wl] = n; p.consumed[w] = 1; }
      if (valid) load_agent_reset(blk0, lane, a); else zero_agent(a);  // a snapshot is at rest
      e = ego_frame(a.px, a.py, a.gx, a.gy, a.hd);
    }
    if (valid || do_reset)  // a reset rewrites every slot of the world (the new scenario may have fewer agents)
      store_agent(blk, lane, a, a.policy == CA_POLICY_STATIC, do_reset);
    bool c_unused;
    double n_unused;
    const int nm1r = others_bound<kA>(n);
    fast_pair_pass_any<kA, false, kGen>(p, a, e, valid, n, nm1r, i, base, o, c_unused, n_unused);
    fast_write_obs_row_any<kA, kGen>(p, a, e, world_ok, valid, nm1r, i, base, o, row, sidx_row, wtile, (tile_floats + 7) & ~3);
  }

  CA_STAMP(p, chunk, 4, lane, 0);
  const int worlds_left = p.W - first_world_warp;
  if (worlds_left > 0) {
    const int nf = (worlds_left < wpw ? worlds_left : wpw) * kA * p.L;
    if (warp_tile_store(p, dst, wtile + shift, shift, nf, lane) && lane == 0)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile must outlive the store's reads
  }
  CA_STAMP(p, chunk, 5, lane, 0);
#ifdef CA_TRACE
  if (lane == 0) { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); p.trace[(size_t)chunk * 8 + 6] = sm_; p.trace[(size_t)chunk * 8 + 7] = 0; }
#endif
}

}  // namespace ca
