// ca_step_fast.cuh — the one-shot step kernel specialised on the number of agent slots per world (kA): the default.
//
// Same algorithm, data layout and numerics as ca_world_kernel<true> (ca_kernels.cuh), but with kA a
// compile-time constant every per-other loop is fully unrolled: the lane's int32 sort keys, p_orth and centre
// distances of its (kA-1) potential neighbours live in registers (shared with ca_step_pipe.cuh: pipe_pair_pass /
// pipe_write_obs_row), the neighbour order is a pairwise rank (each unordered pair of keys is compared once) and
// nothing but the finished observation rows goes through shared memory; each warp hands its rows to one TMA bulk
// store.  One CTA = 4 warps = 4 chunks of floor(32/kA) worlds; the launch-bounds occupancy target (kMinBlocks) is
// tuned per kA so that the grid finishes in as few rounds of resident warps as possible.  Used for closest_first /
// closest_last sorting; time_to_impact sorting, reset, and agent counts without an instantiation run on the
// generic kernel.
#pragma once
#include "ca_kernels.cuh"
#include "ca_step_pipe.cuh"

namespace ca {

// Warp-level store of the warp's observation rows (its wpw worlds are contiguous in global memory).
template <int kA>
__device__ __forceinline__ void fast_store_warp_tile(const Params& p, const float* wtile, long first_world_warp,
                                                     int lane) {
  constexpr int wpw = (32 / kA) < 16 ? (32 / kA) : 16;
  const long worlds_left = (long)p.W - first_world_warp;
  if (worlds_left <= 0) return;
  const int nw = worlds_left < wpw ? (int)worlds_left : wpw;
  const int nfloats = nw * kA * p.L;
  float* dst = p.obs + (size_t)first_world_warp * kA * p.L;
  if (p.use_bulk_store == 1 && nw == wpw) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(wtile)),
                   "r"(nfloats * 4)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  __syncwarp();
  if (p.use_bulk_store == 2 && (nfloats & 3) == 0) {  // 128-bit coalesced copy-out, no wait on the async proxy
    const float4* s4 = reinterpret_cast<const float4*>(wtile);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int q = lane; q < nfloats / 4; q += 32) d4[q] = s4[q];
    return;
  }
  for (int q = lane; q < nfloats; q += 32) dst[q] = wtile[q];
}

template <int kA, int kMinBlocks, bool kDbg>
__global__ void __launch_bounds__(kBlock, kMinBlocks) ca_step_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int wpw = (32 / kA) < 16 ? (32 / kA) : 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wl = lane / kA;
  const int i = lane - wl * kA;
  const int base = wl * kA;
  const long first_world_warp = ((long)blockIdx.x * kWarps + warp) * wpw;
  const long w = first_world_warp + wl;
  const bool world_ok = wl < wpw && w < p.W;
  const size_t g = world_ok ? (size_t)w * kA + i : 0;
  const long chunk = (long)blockIdx.x * kWarps + warp;
  double* const blk = blk_ptr(p.s, chunk);
  pdl_wait();                // nothing produced by the previous kernel is read above this line
  pdl_launch_dependents();
  int n = world_ok ? blk_nag(blk)[wl] : 0;
  bool valid = world_ok && i < n;
  const unsigned gmask = (kA >= 32 ? kFull : ((1u << kA) - 1u)) << (base & 31);

  // per-warp tile: rows of the warp's wpw worlds; warp tiles are laid out back to back (no padding) so the
  // CTA tile is also contiguous
  float* wtile = reinterpret_cast<float*>(smem_raw) + (size_t)warp * wpw * kA * p.L;
  float* row = wtile + ((size_t)wl * kA + i) * p.L;
  int32_t* sidx_row = (p.sidx && world_ok) ? p.sidx + g * p.M : nullptr;

  // The state loads do not wait for the agent count: every slot of an existing world is loaded (absent slots hold
  // zeros / stale values and are discarded below), so only ONE DRAM round trip is exposed instead of two.
  Agent a;
  int act = 0;
  if (world_ok) {
    load_agent<false>(blk, lane, a);
    act = p.actions[g];
  }
  if (!valid) { zero_agent(a); act = 0; }

  // ---- _take_action (:217-252)
  const bool was_done = (a.flags & CA_F_DONE_MASK) != 0;
  float cmd_speed = 0.f, cmd_dh = 0.f;
  if (valid && !was_done) {
    if (a.policy == CA_POLICY_NONCOOP) {
      Ego e0;
      ego_axes(a.px, a.py, a.gx, a.gy, e0);
      cmd_speed = (float)a.ps;
      cmd_dh = (float)(-heading_ego_exact(e0, a.hd));
    } else if (a.policy == CA_POLICY_LEARNING_GA3C) {
      const int k = act < 0 ? 0 : (act > 10 ? 10 : act);
      cmd_speed = (float)(a.ps * kActSpeed[k]);
      cmd_dh = (float)kActDhead[k];
    } else if (a.policy == CA_POLICY_LEARNING) {
      double e0 = 0.0, e1 = 0.5;
      if (p.cont) { e0 = p.cont[2 * g]; e1 = p.cont[2 * g + 1]; }
      cmd_speed = (float)(a.ps * e0);
      cmd_dh = (float)(p.max_heading_change * (2. * e1 - 1.));
    } else if (a.policy == CA_POLICY_STATIC) {
      a.gx = a.px;
      a.gy = a.py;
    }
  }
  // ---- Agent.take_action (agent.py:190-238)
  if (valid) {
    if (was_done) {
      if (a.flags & CA_F_AT_GOAL) a.flags |= CA_F_WAS_AT_GOAL;
      if (a.flags & CA_F_IN_COLLISION) a.flags |= CA_F_WAS_IN_COLLISION;
      a.vx = 0.0;
      a.vy = 0.0;
    } else {
      const double speed = (double)cmd_speed;
      const double h = wrap_angle((double)cmd_dh + a.hd);
      double sh, ch;
      sincos(h, &sh, &ch);
      a.px += speed * ch * p.dt;
      a.py += speed * sh * p.dt;
      a.vx = speed * ch;
      a.vy = speed * sh;
      a.hd = h;
      const double ex = a.px - a.gx, ey = a.py - a.gy;
      if (ex * ex + ey * ey <= p.thr_sq) a.flags |= CA_F_AT_GOAL; else a.flags &= ~CA_F_AT_GOAL;
      a.tr -= p.dt;
      if (a.tr <= 0.0) a.flags |= CA_F_RAN_OUT_OF_TIME;
    }
  }

  Ego e = ego_frame(a.px, a.py, a.gx, a.gy, a.hd);
  OthersLite<kA> o;
  bool coll;
  double nearest;
  pipe_pair_pass<kA, true, kDbg>(p, a, e, valid, n, i, base, o, coll, nearest);

  if (p.prefetch_chunks > 0 && lane == 0) {
    // The grid runs in about two rounds of resident CTAs.  By now this round's own load burst has drained and DRAM is
    // idle while the warps compute: pull the state block and the actions of the chunk that will run in this slot one
    // round later into L2 (TMA bulk prefetch), so the second round does not start with another DRAM burst.
    const long pc = chunk + p.prefetch_chunks;
    if (pc * wpw < p.W) {
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(blk_ptr(p.s, pc)), "r"(kBlkReadBytes) : "memory");
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.actions + (size_t)pc * wpw * kA) : "memory");
    }
  }

  // ---- _compute_rewards (:319-368)
  double r = p.r_step;
  if (valid) {
    if (a.flags & CA_F_AT_GOAL) {
      if (!(a.flags & CA_F_WAS_AT_GOAL)) r = p.r_goal;
    } else if (!(a.flags & CA_F_WAS_IN_COLLISION)) {
      if (coll) {
        r = p.r_coll;
        a.flags |= CA_F_IN_COLLISION;
      } else if (nearest <= p.close_range) {
        r = -0.1 - nearest / 2.;
      }
    }
    r = fmin(fmax(r, p.r_min), p.r_max);
    if (p.over_mode == CA_OVER_FIRST_AGENT_DONE && i > 0) r = 0.0;
  } else {
    r = 0.0;
  }
  // ---- _check_which_agents_done (:411-439)
  const bool dn = valid ? (a.flags & CA_F_DONE_MASK) != 0 : true;
  const bool learning = valid && (a.policy == CA_POLICY_LEARNING_GA3C || a.policy == CA_POLICY_LEARNING);
  bool blocks_over;
  if (p.over_mode == CA_OVER_ALL_DONE) blocks_over = valid && !dn;
  else if (p.over_mode == CA_OVER_FIRST_AGENT_DONE) blocks_over = valid && i == 0 && !dn;
  else blocks_over = learning && !dn;
  const unsigned alive = __ballot_sync(kFull, blocks_over) & gmask;
  const bool over = alive == 0u;
  if (world_ok) {
    p.reward[g] = (float)r;
    p.done[g] = dn ? 1 : 0;
    if (i == 0) p.over[w] = over ? 1 : 0;
  }
  const bool do_reset = world_ok && over && p.auto_reset;

  // The state is written back BEFORE the observation rows are assembled: heading, time budget, goal and flags die
  // here instead of occupying registers through the ranking / row code.
  if (!__any_sync(kFull, do_reset)) {
    if (valid) store_agent(blk, lane, a, a.policy == CA_POLICY_STATIC, false);
    pipe_write_obs_row<kA, kDbg>(p, a, e, world_ok, valid, i, base, o, row, sidx_row);
  } else {
    // DummyVecEnv semantics: worlds that finished reload their injected initial state and observe again;
    // the other worlds of the warp observe their post-step state.
    if (do_reset) {
      double* const blk0 = blk_ptr(p.s0, chunk);
      n = blk_nag(blk0)[wl];
      valid = i < n;
      if (i == 0) { blk_nag(blk)[wl] = n; p.consumed[w] = 1; }
      if (valid) load_agent(blk0, lane, a); else zero_agent(a);
      e = ego_frame(a.px, a.py, a.gx, a.gy, a.hd);
    }
    if (valid || do_reset)  // a reset rewrites every slot of the world (the new scenario may have fewer agents)
      store_agent(blk, lane, a, a.policy == CA_POLICY_STATIC, do_reset);
    bool c_unused;
    double n_unused;
    pipe_pair_pass<kA, false, kDbg>(p, a, e, valid, n, i, base, o, c_unused, n_unused);
    pipe_write_obs_row<kA, kDbg>(p, a, e, world_ok, valid, i, base, o, row, sidx_row);
  }

  if (p.warp_store) fast_store_warp_tile<kA>(p, wtile, first_world_warp, lane);
  else store_tile(p, reinterpret_cast<float*>(smem_raw), (long)blockIdx.x * kWarps * wpw, tid);
}

}  // namespace ca
