// ca_step_pipe.cuh — persistent, software-pipelined step kernel (CA_STEP_KERNEL=pipe; the one-shot kernel of
// ca_step_fast.cuh is the default and shares this file's pair pass and row assembly).
//
// Same arithmetic as ca_world_kernel<true> / ca_step_kernel<kA> (bitwise-identical outputs, see
// tests/test_gpu_parity.py::test_specialised_and_generic_kernels_agree_bitwise); what changes is how the data
// moves:
//   * the grid is sized to the number of CTAs that fit on the chip; every WARP owns a strided sequence of
//     32-lane chunks (= floor(32/kA) worlds) and loops over them, no CTA-wide barrier anywhere;
//   * while a warp computes chunk c, the ten float64 state arrays of its next chunk are already in flight:
//     one lane arms an mbarrier with the expected byte count and issues ten TMA bulk copies
//     (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, SASS UBLKCP) into the warp's
//     shared-memory stage; flags / policy / action / agent-count of the next chunk are prefetched into registers;
//   * the finished observation rows leave through one TMA bulk store per chunk; the wait for that store
//     (cp.async.bulk.wait_group.read) is deferred to just before the tile is overwritten one chunk later.
// Sort keys are int32 (rint(100*dist) is an integer), so ranking runs on the integer pipe.
#pragma once
#include <limits.h>

#include "ca_kernels.cuh"

namespace ca {

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, int bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int kStageBytes = kBlkReadBytes;            // one warp's state stage = the part of a chunk block a step reads (2176 B)

template <int kA>
struct OthersLite {
  static constexpr int kN = kA > 1 ? kA - 1 : 1;
  int key[kN];      // rint(100 * dist_2_other); INT_MAX for an absent / unobserved other
  double po[kN];    // p_orth (float64: it breaks ties between equal keys)
  // observation-only copies (float32 outputs with 1e-5 tolerance, no decision reads them): relative position of the
  // other, its radius and the boundary distance, so that the row assembly needs no second round of float64 shuffles
  float rx[kN], ry[kN], rr[kN], d2o[kN];
};

// kDbg selects the rarely used features (finite SENSING_HORIZON, neighbour-index output for parity tests); the
// production instantiation (kDbg = false) carries neither their instructions nor their registers.
template <int kA, bool kCollide, bool kDbg>
__device__ __forceinline__ void pipe_pair_pass(const Params& p, const Agent& a, const Ego& e, bool valid, int n, int i,
                                               int base, OthersLite<kA>& o, bool& coll, double& nearest) {
  coll = false;
  nearest = INFINITY;
  const bool horizon = kDbg && isfinite(p.sensing_horizon);
#pragma unroll
  for (int k = 0; k < kA - 1; ++k) {
    const int j = k + (k >= i ? 1 : 0);
    const int src = (base + j) & 31;
    const double xj = shfl_d(a.px, src), yj = shfl_d(a.py, src), rj = shfl_d(a.rad, src);
    const bool live = valid && j < n;
    const double rx = xj - a.px, ry = yj - a.py;
    const double d = sqrt(rx * rx + ry * ry);
    if (kCollide && live) {
      const double R = a.rad + rj;
      if (d <= R) coll = true;
      if (j > i) nearest = fmin(nearest, d - R);
    }
    const bool seen = live && !(horizon && d > p.sensing_horizon);
    const double d2o = d - a.rad - rj;
    o.key[k] = seen ? __double2int_rn(d2o * 100.0) : INT_MAX;
    o.po[k] = dot2(rx, ry, -e.pry, e.prx);
    o.rx[k] = (float)rx; o.ry[k] = (float)ry; o.rr[k] = (float)rj; o.d2o[k] = (float)d2o;
  }
}

__device__ __forceinline__ bool key_first(int q1, double p1, int q2, double p2) {
  return (q1 < q2) || (q1 == q2 && p1 <= p2);
}

template <int kA, bool kDbg>
__device__ __forceinline__ void pipe_write_obs_row(const Params& p, const Agent& a, const Ego& e, bool world_ok,
                                                   bool valid, int i, int base, const OthersLite<kA>& o, float* row,
                                                   int32_t* sidx_row_in) {
  int32_t* const sidx_row = kDbg ? sidx_row_in : nullptr;
  constexpr int kN = kA - 1;
  constexpr int kNN = kN > 0 ? kN : 1;
  const int M = p.M;
  int count = 0;
  int key[kNN];
#pragma unroll
  for (int k = 0; k < kN; ++k) {
    key[k] = o.key[k];
    count += (key[k] != INT_MAX) ? 1 : 0;
  }
  if (count > M) {
    int rank1[kNN];
#pragma unroll
    for (int k = 0; k < kN; ++k) rank1[k] = 0;
#pragma unroll
    for (int k1 = 0; k1 < kN; ++k1)
#pragma unroll
      for (int k2 = k1 + 1; k2 < kN; ++k2) {
        const bool b = key_first(o.key[k1], o.po[k1], o.key[k2], o.po[k2]);
        rank1[k1] += b ? 0 : 1;
        rank1[k2] += b ? 1 : 0;
      }
#pragma unroll
    for (int k = 0; k < kN; ++k)
      if (rank1[k] >= M) key[k] = INT_MAX;
    count = M;
  }
  if (p.sort_method == CA_SORT_CLOSEST_LAST) {
#pragma unroll
    for (int k = 0; k < kN; ++k)
      if (key[k] != INT_MAX) key[k] = -key[k];
  }
  int slot[kNN];
#pragma unroll
  for (int k = 0; k < kN; ++k) slot[k] = 0;
#pragma unroll
  for (int k1 = 0; k1 < kN; ++k1)
#pragma unroll
    for (int k2 = k1 + 1; k2 < kN; ++k2) {
      const bool b = key_first(key[k1], o.po[k1], key[k2], o.po[k2]);
      slot[k1] += b ? 0 : 1;
      slot[k2] += b ? 1 : 0;
    }
  if (world_ok) {
    if (valid) {
      row[0] = (a.policy == CA_POLICY_LEARNING_GA3C || a.policy == CA_POLICY_LEARNING) ? 1.f : 0.f;
      row[1] = (float)count;
      row[2] = (float)e.dist;
      row[3] = e.hego;
      row[4] = (float)a.ps;
      row[5] = (float)a.rad;
      for (int q = CA_OBS_HOST_LEN + CA_OBS_OTHER_LEN * count; q < p.L; ++q) row[q] = 0.f;
      if (sidx_row) for (int k = count; k < M; ++k) sidx_row[k] = -1;
    } else {
      for (int q = 0; q < p.L; ++q) row[q] = 0.f;
      if (sidx_row) for (int k = 0; k < M; ++k) sidx_row[k] = -1;
    }
  }
  // second loop of the sensor (:105-144): only the others' velocities still have to be fetched, as float32
  const float prxf = (float)e.prx, pryf = (float)e.pry, raf = (float)a.rad;
  const float vxf = (float)a.vx, vyf = (float)a.vy;
#pragma unroll
  for (int k = 0; k < kN; ++k) {
    const int j = k + (k >= i ? 1 : 0);
    const int src = (base + j) & 31;
    const float vxj = __shfl_sync(kFull, vxf, src), vyj = __shfl_sync(kFull, vyf, src);
    if (valid && key[k] != INT_MAX) {
      float* s = row + CA_OBS_HOST_LEN + CA_OBS_OTHER_LEN * slot[k];
      s[0] = fmaf(o.ry[k], pryf, o.rx[k] * prxf);
      s[1] = (float)o.po[k];
      s[2] = fmaf(vyj, pryf, vxj * prxf);
      s[3] = fmaf(vyj, prxf, -(vxj * pryf));
      s[4] = o.rr[k];
      s[5] = raf + o.rr[k];
      s[6] = o.d2o[k];
      if (sidx_row) sidx_row[slot[k]] = j;
    }
  }
}

template <int kA, int kMinBlocks, bool kDbg>
__global__ void __launch_bounds__(kBlock, kMinBlocks) ca_step_pipe_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int wpw = (32 / kA) < 16 ? (32 / kA) : 16;
  constexpr int kLanesUsed = wpw * kA;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wl = lane / kA;
  const int i = lane - wl * kA;
  const int base = wl * kA;
  const unsigned gmask = (kA >= 32 ? kFull : ((1u << kA) - 1u)) << (base & 31);
  const bool lane_used = lane < kLanesUsed;

  // per-warp shared memory: [state stage 2560 B][observation tile][mbarrier]
  const int tile_floats = wpw * kA * p.L;
  const int tile_bytes16 = ((tile_floats * 4 + 15) / 16) * 16;
  const int region = kStageBytes + tile_bytes16 + 16;
  unsigned char* wbase = smem_raw + (size_t)warp * region;
  double* stage = reinterpret_cast<double*>(wbase);
  float* wtile = reinterpret_cast<float*>(wbase + kStageBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(wbase + kStageBytes + tile_bytes16);
  float* row = wtile + ((size_t)wl * kA + i) * p.L;

  const long n_chunks = ((long)p.W + wpw - 1) / wpw;
  const long gw = (long)blockIdx.x * kWarps + warp;
  const long GW = (long)gridDim.x * kWarps;
  const bool tile_bulk = p.use_bulk_store && (tile_floats % 4) == 0;

  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  // issue the loads of chunk c: ONE TMA bulk copy brings the chunk's whole state block (10 float64 fields, flags,
  // policies, agent counts) into the warp's stage; only the caller's action array is prefetched into a register
  int pf_act = 0;
  auto prefetch = [&](long c) {
    const long w = c * wpw + wl;
    const bool ok = lane_used && w < p.W;
    pf_act = ok ? p.actions[(size_t)w * kA + i] : 0;
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, kBlkReadBytes);
      tma_load_1d(stage, blk_ptr(p.s, c), kBlkReadBytes, bar);
    }
  };

  unsigned parity = 0;
  bool store_pending = false;
  pdl_wait();                // nothing produced by the previous kernel is read above this line
  pdl_launch_dependents();
  if (gw < n_chunks) prefetch(gw);

  for (long c = gw; c < n_chunks; c += GW) {
    const long first_world = c * wpw;
    const long w = first_world + wl;
    const bool world_ok = lane_used && w < p.W;
    const size_t g = world_ok ? (size_t)w * kA + i : 0;
    const bool full = (c + 1) * wpw <= p.W;
    double* const blk = blk_ptr(p.s, c);
    int32_t* sidx_row = (p.sidx && world_ok) ? p.sidx + g * p.M : nullptr;

    // the chunk's block has landed in the stage (padded blocks make partial last chunks loadable too)
    mbar_wait(bar, parity);
    parity ^= 1u;
    int n = world_ok ? blk_nag(stage)[wl] : 0;
    bool valid = world_ok && i < n;
    Agent a;
    if (valid) load_agent<false>(stage, lane, a); else zero_agent(a);
    const int act = pf_act;
    __syncwarp();  // every lane has copied its state out of the stage: it may be refilled
    if (c + GW < n_chunks) prefetch(c + GW);

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

    if (__any_sync(kFull, do_reset)) {
      if (do_reset) {
        double* const blk0 = blk_ptr(p.s0, c);
        n = blk_nag(blk0)[wl];
        valid = i < n;
        if (i == 0) { blk_nag(blk)[wl] = n; p.consumed[w] = 1; }
        if (valid) load_agent(blk0, lane, a); else zero_agent(a);
        e = ego_frame(a.px, a.py, a.gx, a.gy, a.hd);
      }
      bool c_unused;
      double n_unused;
      pipe_pair_pass<kA, false, kDbg>(p, a, e, valid, n, i, base, o, c_unused, n_unused);
    }

    // the tile still feeds the previous chunk's bulk store until that store has read it
    if (store_pending) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      store_pending = false;
    }
    pipe_write_obs_row<kA, kDbg>(p, a, e, world_ok, valid, i, base, o, row, sidx_row);

    // ---- state write-back (coalesced)
    if (valid || do_reset) store_agent(blk, lane, a, a.policy == CA_POLICY_STATIC, do_reset);

    // ---- observation tile -> global
    float* dst = p.obs + (size_t)first_world * kA * p.L;
    if (tile_bulk && full) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(wtile)),
                     "r"(tile_floats * 4)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      store_pending = true;
    } else {
      __syncwarp();
      const long worlds_left = (long)p.W - first_world;
      const int nw = worlds_left < wpw ? (int)worlds_left : wpw;
      const int nf = nw * kA * p.L;
      for (int q = lane; q < nf; q += 32) dst[q] = wtile[q];
      __syncwarp();
    }
  }
  if (store_pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

}  // namespace ca
