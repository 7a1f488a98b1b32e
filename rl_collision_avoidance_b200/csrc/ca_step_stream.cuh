// ca_step_stream.cuh — persistent, dynamically scheduled, TMA-pipelined step kernel (the default for the agent counts
// with a specialisation; CA_STEP_KERNEL=oneshot / generic select the other two).
//
// Same step body and arithmetic as ca_world_kernel<true> / ca_step_kernel<kA> (bitwise-identical outputs, see
// tests/test_gpu_parity.py::test_specialised_and_generic_kernels_agree_bitwise); what changes is how the work and the
// data move:
//   * the grid is the number of CTAs that are resident at once; every WARP is an independent worker (no CTA-wide
//     barrier anywhere).  Warp gw starts with chunk gw and then pulls further chunks from an atomic counter in global
//     memory, so a warp that finishes early takes more work instead of waiting for a "round" to end and the warps of an
//     SM drift out of phase: one warp's load / store meets another's arithmetic.  The counter resets itself: exactly
//     n_chunks tickets are drawn per launch and the warp that draws the last one writes 0 back.
//   * while a warp computes chunk c, the state block of its next chunk is already in flight: one lane arms an mbarrier
//     with the byte count and issues ONE TMA bulk copy of the 2176 bytes a step reads
//     (cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes, SASS UBLKCP) into the warp's stage; the
//     action of the next chunk is prefetched into a register and the ticket after that is in flight too.
//   * the finished observation rows leave through one TMA bulk store per chunk.  The rows are assembled at the same
//     16-byte phase in shared memory as their destination has in global memory, so the bulk store covers the
//     16-byte-aligned body of ANY tile size (10 agents x 69 floats included) and at most 3 + 3 floats go out as scalar
//     head / tail stores; the wait for the store engine to have read the tile
//     (cp.async.bulk.wait_group.read) is deferred to just before the tile is overwritten one chunk later.
#pragma once
#include "ca_kernels.cuh"
#include "ca_step_fast.cuh"

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

constexpr int kStageBytes = kBlkReadBytes;  // one warp's state stage = the part of a chunk block a step reads (2176 B)

// bytes of shared memory one warp of the streaming kernel needs: [state stage][tile + 16 B phase room][mbarrier]
__host__ __device__ __forceinline__ int stream_warp_region(int tile_floats) {
  return kStageBytes + warp_tile_region(tile_floats) + 16;
}

template <int kA, int kMinBlocks, bool kGen>
__global__ void __launch_bounds__(kBlock, kMinBlocks) ca_step_stream_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int wpw = (32 / kA) < 16 ? (32 / kA) : 16;
  constexpr int kLanesUsed = wpw * kA;
  const int lane = threadIdx.x & 31, warp = __shfl_sync(kFull, threadIdx.x >> 5, 0);   // warp-uniform for the compiler
  const int wl = lane / kA;
  const int i = lane - wl * kA;
  const int base = wl * kA;
  const unsigned gmask = (kA >= 32 ? kFull : ((1u << kA) - 1u)) << (base & 31);
  const bool lane_used = lane < kLanesUsed;

  const int tile_floats = wpw * kA * p.L;
  unsigned char* wbase = smem_raw + (size_t)warp * stream_warp_region(tile_floats);
  double* stage = reinterpret_cast<double*>(wbase);
  float* wtile = reinterpret_cast<float*>(wbase + kStageBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(wbase + stream_warp_region(tile_floats) - 16);

  const int n_chunks = (int)(((long)p.W + wpw - 1) / wpw);
  const int gw = blockIdx.x * kWarps + warp;
  const int GW = gridDim.x * kWarps;
  const int first_ticket_chunk = GW < n_chunks ? GW : n_chunks;  // chunk of ticket 0

  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  pdl_wait();                // nothing produced by the previous kernel is read above this line
  pdl_launch_dependents();
  if (gw >= n_chunks) return;

  // loads of chunk c: the whole state block by ONE bulk copy into the stage; the caller's actions are only pulled into
  // L2 here (a register holding them across the whole iteration was spilled at 72 registers, which made the warp wait
  // for the load at once) and are loaded at the end of the iteration, just before the tile store
  auto prefetch = [&](int c) {
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, kBlkReadBytes);
      tma_load_1d(stage, blk_ptr(p.s, c), kBlkReadBytes, bar);
    }
    if (lane < 2) {  // the chunk's wpw * kA actions: at most two 128-byte lines
      const int32_t* a0 = p.actions + (size_t)c * kLanesUsed;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(lane == 0 ? a0 : a0 + kLanesUsed - 1) : "memory");
    }
  };
  auto load_action = [&](int c) {
    const int w = c * wpw + wl;
    return (lane_used && w < p.W) ? p.actions[(unsigned)w * kA + i] : 0;
  };
  // draw the next ticket (lane 0; the result is consumed one chunk later, so its latency is hidden)
  int ticket = 0;
  auto draw = [&]() {
    if (lane == 0) {
      ticket = p.dynamic_sched ? (int)atomicAdd(p.ticket, 1u) : 0;
    }
  };

  unsigned parity = 0;
  bool store_pending = false;
  int c = gw;
  int static_next = gw + GW;  // schedule without the counter (p.dynamic_sched = 0): strided
  prefetch(c);
  int act_next = load_action(c);
  draw();

  while (true) {
    const int first_world = c * wpw;   // 32-bit indexing: ca_create refuses W * A >= 2^31
    const int w = first_world + wl;
    const bool world_ok = lane_used && w < p.W;
    const unsigned g = world_ok ? (unsigned)w * kA + i : 0u;
    double* const blk = blk_ptr(p.s, c);
    int32_t* sidx_row = (kGen && p.sidx && world_ok) ? p.sidx + (size_t)g * p.M : nullptr;

    // the chunk's block has landed in the stage (padded blocks make partial last chunks loadable too)
    CA_STAMP(p, c, 0, lane, 0);
    mbar_wait(bar, parity);
    parity ^= 1u;
    CA_STAMP(p, c, 1, lane, 0);
    int n = world_ok ? blk_nag(stage)[wl] : 0;
    bool valid = world_ok && i < n;
    Agent a;
    if (valid) load_agent<false>(stage, lane, a); else zero_agent(a);
    const int act = valid ? act_next : 0;
    __syncwarp();  // every lane has copied its state out of the stage: it may be refilled

    // next chunk of this warp: the ticket drawn one iteration ago
    int nxt;
    if (p.dynamic_sched) {
      const int t = __shfl_sync(kFull, ticket, 0);
      nxt = first_ticket_chunk + t;
      if (t == n_chunks - 1 && lane == 0) *p.ticket = 0u;  // the last ticket of the launch: reset for the next launch
    } else {
      nxt = static_next;
      static_next += GW;
    }
    const bool more = nxt < n_chunks;
    if (more) {
      prefetch(nxt);
      draw();
    }

    step_take_action<kGen>(p, a, act, g, valid);
    CA_STAMP(p, c, 2, lane, a.flags);
    if (!kGen) prefetch_snapshot_if_ending(p, a, world_ok, valid, gmask, blk_ptr(p.s0, c), lane);

    Ego e = ego_frame(a.px, a.py, a.gx, a.gy, a.hd);
    OthersLite<kA> o;
    bool coll;
    double nearest;
    const int nm1 = others_bound<kA>(n);
    fast_pair_pass_any<kA, true, kGen>(p, a, e, valid, n, nm1, i, base, o, coll, nearest);

    bool dn, over;
    const float r = step_reward_done<kGen>(p, a, valid, i, coll, nearest, gmask, dn, over);
    CA_STAMP(p, c, 3, lane, __float_as_int(r));
    if (world_ok) {
      p.reward[g] = r;
      p.done[g] = dn ? 1 : 0;
      if (i == 0) p.over[w] = over ? 1 : 0;
    }
    const bool do_reset = world_ok && over && p.auto_reset;

    // observation rows are assembled at the destination's 16-byte phase
    float* const dst = p.obs + (size_t)first_world * kA * p.L;
    const int shift = tile_shift(dst);
    float* const row = wtile + shift + ((size_t)wl * kA + i) * p.L;

    // the tile still feeds the previous chunk's bulk store until that store has read it
    if (store_pending) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      store_pending = false;
    }

    if (!__any_sync(kFull, do_reset)) {
      // state write-back first: heading, time budget, goal and flags die here instead of living through the row code
      if (valid) store_agent(blk, lane, a, a.policy == CA_POLICY_STATIC, false);
      fast_write_obs_row_any<kA, kGen>(p, a, e, world_ok, valid, nm1, i, base, o, row, sidx_row, wtile, (tile_floats + 7) & ~3);
    } else {
      // DummyVecEnv semantics: worlds that finished reload their snapshot and observe again
      if (do_reset) {
        double* const blk0 = blk_ptr(p.s0, c);
        n = blk_nag(blk0)[wl];
        valid = i < n;
        if (i == 0) { blk_nag(blk)[wl] = n; p.consumed[w] = 1; }
        if (valid) load_agent_reset(blk0, lane, a); else zero_agent(a);  // a snapshot is at rest
        e = ego_frame(a.px, a.py, a.gx, a.gy, a.hd);
      }
      if (valid || do_reset) store_agent(blk, lane, a, a.policy == CA_POLICY_STATIC, do_reset);
      bool c_unused;
      double n_unused;
      const int nm1r = others_bound<kA>(n);
      fast_pair_pass_any<kA, false, kGen>(p, a, e, valid, n, nm1r, i, base, o, c_unused, n_unused);
      fast_write_obs_row_any<kA, kGen>(p, a, e, world_ok, valid, nm1r, i, base, o, row, sidx_row, wtile, (tile_floats + 7) & ~3);
    }

    if (more) act_next = load_action(nxt);   // an L2 hit by now; consumed after the next chunk's barrier wait
    // ---- observation tile -> global (the wait for the store is deferred to the next chunk)
    CA_STAMP(p, c, 4, lane, 0);
    const int worlds_left = p.W - first_world;
    const int nf = (worlds_left < wpw ? worlds_left : wpw) * kA * p.L;
    store_pending = warp_tile_store(p, dst, wtile + shift, shift, nf, lane);
    CA_STAMP(p, c, 5, lane, 0);
#ifdef CA_TRACE
    if (lane == 0) { unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_)); p.trace[(size_t)c * 8 + 6] = sm_; p.trace[(size_t)c * 8 + 7] = (unsigned)gw; }
#endif
    if (!more) break;
    c = nxt;
  }
  if (store_pending && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

}  // namespace ca
