// ca_kernels.cuh — the fused env.step() kernel for sm_100a.
//
// One launch advances every world one time step (reference: CollisionAvoidanceEnv.step,
// GCA/envs/collision_avoidance_env.py:131-194).  Mapping: lane = agent; a warp carries
// floor(32 / A) whole worlds in consecutive lane groups, so the flat (world, agent) index of a lane is
// first_world_of_warp * A + lane and every SoA state array is read/written as one contiguous,
// fully coalesced run per warp.  Other agents' states are fetched with warp shuffles inside the lane
// group (all-pairs distance), the neighbour order is a stable rank-by-counting over per-lane keys
// held in shared memory, and the observation rows are assembled in a shared-memory tile that the CTA
// writes out as one contiguous block (TMA bulk store cp.async.bulk.global.shared::cta when the tile is
// 16-byte aligned, else coalesced scalar stores).
//
// Numerics: state and every decision (collision d <= R, goal test, time-out,
// sort keys) are IEEE float64 computed with the same operation order as the reference's Python
// float64 arithmetic; the file is compiled with -fmad=false so nvcc never contracts a*b+c, and the
// places where NumPy itself fuses (np.dot on 2-vectors) use __fma_rn explicitly.  Only the commanded
// [speed, dheading] is rounded to float32 (collision_avoidance_env.py:238) and outputs are float32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ca_step.h"
#include "ca_math.cuh"

namespace ca {

constexpr int kBlock = 128;           // threads per CTA
constexpr int kWarps = kBlock / 32;   // warps per CTA
constexpr unsigned kFull = 0xffffffffu;

// ---- state layout in HBM: chunk-major AoSoA --------------------------------------------------------------------------
// A "chunk" is what one warp processes: wpw = min(32 / A, 16) whole worlds = wpw * A (<= 32) agent slots, one per lane.
// All state of a chunk lives in ONE contiguous, 128-byte aligned block of kBlkBytes = 2560 bytes:
//     double field[7][32]   px py heading | gx gy radius pref_speed                                   (read by a step)
//     uint32 meta[32]       flags | policy << 8 | countdown << 16, per lane                           (read by a step)
//     int32  num_agents[16] per world of the chunk                                                    (read by a step)
//     float  speed[32]      the float32 speed command the agent last executed (0 once it is done)
//     double tr0[32]        Agent.time_remaining_to_reach_goal at the last reset
//     uint32 n0[32]         the countdown at the last reset
// so a warp reads/writes each field as one coalesced run at a constant offset from a single base pointer, and the part
// of a block that a step has to fetch is ONE TMA bulk copy of kBlkReadBytes = 1984 bytes (ca_step_stream.cuh).  The
// reset snapshot uses the same layout in a second buffer.
// The velocity is not stored: a step writes it as speed * (cos h, sin h) with h the heading it also writes
// (UnicycleDynamics.step, dynamics/UnicycleDynamics.py:33-34), so the 4-byte command reproduces both float64 components
// bit for bit (ca_get_state, first observation after a partial reset).
// The time budget is not stored per step either.  The reference subtracts dt from time_remaining_to_reach_goal every
// step the agent is still running and flags ran_out_of_time when the result is <= 0 (agent.py:232-236); the result of
// that repeated rounded subtraction has no closed form, but the NUMBER of subtractions until it is <= 0 is fixed once
// the budget and dt are known.  Whoever writes a budget (unpack_init_kernel, the scenario generator, ca_set_dt) runs the
// subtractions once (countdown_steps) and a step only decrements a 16-bit countdown that shares a word with the flags:
// 4 bytes read + written per agent-step instead of 17, the same step at which the flag rises, and ca_get_state rebuilds
// the float64 value by repeating the (n0 - countdown) subtractions from tr0.
constexpr int kFields = 7;
constexpr int O_PX = 0, O_PY = 32, O_HD = 64, O_GX = 96, O_GY = 128, O_RAD = 160, O_PS = 192;
constexpr int O_META = 224;                     // uint32[32]
constexpr int O_NAG = 240;                      // int32[16]
constexpr int O_SPD = 248;                      // float[32]; the read window of a step ends here
constexpr int O_TR0 = 264;                      // double[32]
constexpr int O_N0 = 296;                       // uint32[32]
constexpr int kBlkDoubles = 320;
constexpr int kBlkBytes = kBlkDoubles * 8;      // 2560 bytes
constexpr int kBlkReadBytes = O_SPD * 8;        // 1984 bytes
constexpr unsigned kNever = 0xFFFFu;            // countdown value of a budget that does not run out within 65 534 steps

struct StateBlocks {
  double* base;  // [n_chunks][kBlkDoubles]
};

__host__ __device__ __forceinline__ int worlds_per_chunk(int A) { return (32 / A) < 16 ? (32 / A) : 16; }

__device__ __forceinline__ double* blk_ptr(const StateBlocks& s, long chunk) { return s.base + chunk * kBlkDoubles; }
__device__ __forceinline__ uint32_t* blk_meta(double* blk) { return reinterpret_cast<uint32_t*>(blk + O_META); }
__device__ __forceinline__ int32_t* blk_nag(double* blk) { return reinterpret_cast<int32_t*>(blk + O_NAG); }
__device__ __forceinline__ float* blk_spd(double* blk) { return reinterpret_cast<float*>(blk + O_SPD); }
__device__ __forceinline__ uint32_t* blk_n0(double* blk) { return reinterpret_cast<uint32_t*>(blk + O_N0); }
__host__ __device__ __forceinline__ uint32_t pack_meta(unsigned flags, int policy, unsigned cd) {
  return (flags & 0xffu) | (((unsigned)policy & 0xffu) << 8) | (cd << 16);
}

// Number of `-= dt` steps after which a time budget tr0 is <= 0 (agent.py:232-236), saturating at kNever.
__host__ __device__ __forceinline__ unsigned countdown_steps(double tr0, double dt) {
  double t = tr0;
  unsigned n = 0;
  do { t -= dt; ++n; } while (t > 0.0 && n < kNever);
  return n;
}
// The budget after k of those steps.
__host__ __device__ __forceinline__ double budget_after(double tr0, double dt, unsigned k) {
  double t = tr0;
  for (unsigned q = 0; q < k; ++q) t -= dt;
  return t;
}

// (world, agent) -> (chunk, lane) for kernels that are not organised warp-per-chunk
__device__ __forceinline__ void slot_of(int w, int i, int A, long& chunk, int& lane, int& wl) {
  const int wpw = worlds_per_chunk(A);
  chunk = w / wpw;
  wl = w - (int)chunk * wpw;
  lane = wl * A + i;
}

struct Params {
  int W, A, M, L, wpw;  // wpw = worlds per warp (chunk) = min(32 / A, 16)
  int sort_method, over_mode, auto_reset;
  int tile_floats;      // floats in the CTA's obs tile = kWarps * wpw * A * L
  int use_bulk_store;   // 1: TMA bulk store of full tiles
  int prefetch_chunks;  // one-shot kernel: L2-prefetch the state block of chunk + prefetch_chunks (0 = off)
  int all_present;      // host-side hint: every world has all A agents (ca_step.cu); the kernels still check per warp
  int prefetch_snapshot;  // specialised kernels: L2-prefetch the snapshot block of a chunk with an ending world (auto-reset)
  int dynamic_sched;    // streaming kernel: 1 = warps pull chunks from *ticket, 0 = strided static schedule
  unsigned* ticket;     // streaming kernel: self-resetting work counter (one per env handle)
  double dt, thr_sq, close_range, r_goal, r_coll, r_step, r_min, r_max, max_heading_change, sensing_horizon;
  float r_goal_f, r_coll_f, r_step_f, r_min_f, r_max_f;  // the reward constants as the float32 that leaves the kernel
  StateBlocks s;        // live state (agent counts inside the blocks)
  StateBlocks s0;       // snapshot injected by ca_set_world_state / ca_set_reset_state / the generator (for reset)
  uint8_t* consumed;    // [W] set to 1 when a world takes its snapshot (the scenario generator refills those)
  // I/O (device)
  const int32_t* actions;  // [W*A]
  const double* cont;      // [W*A*2] or null
  const uint8_t* mask;     // reset kernel: [W] or null (= all)
  float* obs;              // [W*A*L]
  float* reward;           // [W*A]
  uint8_t* done;           // [W*A]
  uint8_t* over;           // [W]
  int32_t* sidx;           // [W*A*M] or null
#ifdef CA_TRACE
  unsigned long long* trace;  // experiment builds only (scripts/step_timeline.py): [n_chunks][8] globaltimer stamps
#endif
};

// Timeline stamps of the experiment build (-DCA_TRACE, scripts/step_timeline.py); nothing in the shipped library.
#ifdef CA_TRACE
#define CA_STAMP(p, chunk, k, lane, dep)                                                   \
  do {                                                                                     \
    if ((lane) == 0) {                                                                     \
      unsigned long long t_;                                                               \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_) : "r"((int)(dep)) : "memory");  \
      (p).trace[(size_t)(chunk) * 8 + (k)] = t_;                                           \
    }                                                                                      \
  } while (0)
#else
#define CA_STAMP(p, chunk, k, lane, dep) do { } while (0)
#endif

// Actions table, GCA/envs/policies/GA3C_CADRL/network.py:13-16 (values of the reference's np.mgrid expression)
__constant__ double kActSpeed[11] = {1.0, 1.0, 1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 0.0, 0.0, 0.0};
__constant__ double kActDhead[11] = {-0x1.0c152382d7365p-1, -0x1.0c152382d7365p-2, 0.0, 0x1.0c152382d7366p-2,
                                     0x1.0c152382d7365p-1,  -0x1.0c152382d7365p-1, 0.0, 0x1.0c152382d7365p-1,
                                     -0x1.0c152382d7365p-1, 0.0,                   0x1.0c152382d7365p-1};

// np.dot on 2-vectors as NumPy/OpenBLAS evaluates it: fma(a1, b1, a0*b0) (see oracle/ca_oracle.c np_dot2)
__device__ __forceinline__ double dot2(double a0, double a1, double b0, double b1) {
  return __fma_rn(a1, b1, __dmul_rn(a0, b0));
}

struct Ego {
  double dist, prx, pry;  // orth = (-pry, prx)
  float hego;             // heading_ego_frame as it goes into the float32 observation
};

// Agent.get_ref (GCA/envs/agent.py:326-346): distance to goal and the ego-frame axes, float64.  The two divisions by
// the same distance share one refined reciprocal (ca_math.cuh: same correctly rounded quotients as two IEEE divisions).
__device__ __forceinline__ void ego_axes(double px, double py, double gx, double gy, Ego& e) {
  const double dx = gx - px, dy = gy - py;
  e.dist = sqrt(dx * dx + dy * dy);
  if (e.dist > 1e-8) {
    const double r = recip_refined(e.dist);
    e.prx = div_by(dx, e.dist, r);
    e.pry = div_by(dy, e.dist, r);
  } else {
    e.prx = dx;
    e.pry = dy;
  }
}

// heading_ego_frame = wrap(heading - atan2(ref_prll)) (GCA/envs/dynamics/Dynamics.py:31-35) in float64: this is the
// value the non-cooperative policy turns into its command, i.e. it feeds the dynamics and must be exact.
__device__ __forceinline__ double heading_ego_exact(const Ego& e, double hd) {
  return wrap_angle(hd - atan2(e.pry, e.prx));
}

// The same quantity for the observation vector.  Observations are float32 (tolerance 1e-5); nothing downstream of
// the env reads it back, so it is evaluated with a branch-free float32 atan2 (|error| < 1e-6 rad) which is several
// times cheaper than the float64 one and sits at the end of the kernel's longest dependency chain.
__device__ __forceinline__ float heading_ego_obs(const Ego& e, double hd) {
  const float pi = 3.14159265358979323846f;
  float a = (float)hd - atan2f_obs((float)e.pry, (float)e.prx);
  if (a >= pi) a -= 2.f * pi;
  if (a < -pi) a += 2.f * pi;
  return a;
}

// Observation-only projections (float32 outputs, 1e-5 tolerance; no decision reads them).  One definition for every
// kernel so that they agree bit for bit.  p_parallel as the reference evaluates it (np.dot in float64), then rounded.
__device__ __forceinline__ float obs_pprl(double rx, double ry, const Ego& e) { return (float)dot2(rx, ry, e.prx, e.pry); }
// velocity of the other agent in the ego frame, from its float32 velocity
__device__ __forceinline__ void obs_vel(float vxj, float vyj, float prxf, float pryf, float& vprl, float& vorth) {
  vprl = fmaf(vyj, pryf, vxj * prxf);
  vorth = fmaf(vyj, prxf, -(vxj * pryf));
}

__device__ __forceinline__ Ego ego_frame(double px, double py, double gx, double gy, double hd) {
  Ego e;
  ego_axes(px, py, gx, gy, e);
  e.hego = heading_ego_obs(e, hd);
  return e;
}

// Programmatic dependent launch (PDL): when the kernel is launched with the programmatic-stream-serialization
// attribute its CTAs may become resident while the previous kernel in the stream is still draining; pdl_wait()
// blocks until that kernel has completed and its writes are visible, pdl_launch_dependents() lets the NEXT kernel's
// CTAs start filling SMs that this grid no longer needs.  Both are no-ops for a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }

// compute_time_to_impact, GCA/envs/util.py:14-104 (host pos/vel, other pos/vel, combined radius)
__device__ double time_to_impact(double hx, double hy, double hvx, double hvy, double ox, double oy, double ovx,
                                 double ovy, double r) {
  const double v0 = hvx - ovx, v1 = hvy - ovy;
  const double ex = hx - ox, ey = hy - oy;
  const double sqd = ex * ex + ey * ey - r * r;
  if (sqd < 0) return 0.0;
  const double sqrt_term = sqrt(sqd);
  const double xnum1 = r * r * ex, xnum2 = r * ey * sqrt_term;
  const double ynum1 = r * r * ey, ynum2 = r * ex * sqrt_term;
  const double den = ex * ex + ey * ey;
  const double c1x = ((xnum1 + xnum2) / den + ox) - hx, c1y = ((ynum1 - ynum2) / den + oy) - hy;
  const double c2x = ((xnum1 - xnum2) / den + ox) - hx, c2y = ((ynum1 + ynum2) / den + oy) - hy;
  const double c1v = c1x * v1 - c1y * v0, c12 = c1x * c2y - c1y * c2x;
  const double c2v = c2x * v1 - c2y * v0, c21 = c2x * c1y - c2y * c1x;
  if (!(c1v * c12 >= 0 && c2v * c21 >= 0)) return INFINITY;
  if (fabs(v0) < 1e-5 && fabs(v1) < 1e-5) return INFINITY;
  double x1, x2, y1, y2;
  if (fabs(v0) < 1e-5) {
    x1 = x2 = hx;
    const double B = -2 * oy, C = oy * oy + (hx - ox) * (hx - ox) - r * r;
    const double disc = sqrt(B * B - 4 * C);
    y1 = (-B + disc) / 2;
    y2 = (-B - disc) / 2;
  } else {
    const double m = v1 / v0;
    const double Aq = 1 + m * m;
    const double B = -2 * ox + 2 * m * (hy - oy - m * hx);
    const double t = m * hx - (hy - oy);
    const double C = ox * ox - r * r + t * t;
    const double disc = sqrt(B * B - 4 * Aq * C);
    x1 = (-B + disc) / (2 * Aq);
    x2 = (-B - disc) / (2 * Aq);
    y1 = m * (x1 - hx) + hy;
    y2 = m * (x2 - hx) + hy;
  }
  const double d1 = sqrt(dot2(x1 - hx, y1 - hy, x1 - hx, y1 - hy));
  const double d2 = sqrt(dot2(x2 - hx, y2 - hy, x2 - hx, y2 - hy));
  const double d = (d2 < d1) ? d2 : d1;
  return d / sqrt(dot2(v0, v1, v0, v1));
}

// Strict "sorts before" on the sensor's keys (OtherAgentsStatesSensor.get_clipped_sorted_inds,
// GCA/envs/sensors/OtherAgentsStatesSensor.py:20-55) with the list position (= agent index) as the
// final tiebreak, which is what Python's stable sorted() amounts to.  q = rint(100*dist_2_other) orders
// exactly like round(dist_2_other, 2) (division by 100 is monotone and injective on these integers).
//   mode 0: (q, p_orth)    mode 1: (-q, p_orth)    mode 2: (-tti, -q, p_orth)
__device__ __forceinline__ bool key_before(int mode, double qa, double pa, double ta, int ia, double qb, double pb,
                                           double tb, int ib) {
  const bool tail = (pa < pb) || (pa == pb && ia < ib);
  if (mode == 0) return (qa < qb) || (qa == qb && tail);
  const bool qtail = (qa > qb) || (qa == qb && tail);
  if (mode == 1) return qtail;
  return (ta > tb) || (ta == tb && qtail);
}

// The per-lane agent record kept in registers.
struct Agent {
  double px, py, hd, vx, vy, gx, gy, rad, ps;
  double tr0;   // time budget at the last reset: only loaded / stored where a world is (re)set
  float spd;    // float32 speed command last executed (what the block stores instead of the velocity)
  unsigned flags;
  unsigned cd;  // steps left until ran_out_of_time (kNever: never)
  int policy;
};

// kVel = true: the velocity is rebuilt from the stored float32 speed command and the heading exactly as the step
// that produced them wrote it (speed * cos h, speed * sin h with the same sincos).  kVel = false: the caller
// overwrites vx, vy before using them (every step kernel does: a moving agent gets its new velocity, a finished one
// zero), so neither the command nor the trigonometry is needed.
template <bool kVel = true>
__device__ __forceinline__ void load_agent(const double* blk, int lane, Agent& a) {
  a.px = blk[O_PX + lane]; a.py = blk[O_PY + lane]; a.hd = blk[O_HD + lane];
  a.gx = blk[O_GX + lane]; a.gy = blk[O_GY + lane]; a.rad = blk[O_RAD + lane];
  a.ps = blk[O_PS + lane];
  const uint32_t m = blk_meta(const_cast<double*>(blk))[lane];
  a.flags = m & 0xffu;
  a.policy = (int)((m >> 8) & 0xffu);
  a.cd = m >> 16;
  a.tr0 = 0.0;
  if (kVel) {
    a.spd = blk_spd(const_cast<double*>(blk))[lane];
    a.vx = 0.0; a.vy = 0.0;
    if (a.spd != 0.f) {  // the agent has moved: its heading is a wrapped angle and this is the step's own expression
      double sh, ch;
      sincos_wrapped(a.hd, sh, ch);
      a.vx = (double)a.spd * ch; a.vy = (double)a.spd * sh;
    }
  } else {
    a.spd = 0.f; a.vx = 0.0; a.vy = 0.0;
  }
}

// a snapshot (or any block a world is (re)set from): the step fields plus the time budget the countdown stands for
__device__ __forceinline__ void load_agent_reset(const double* blk, int lane, Agent& a) {
  load_agent<false>(blk, lane, a);
  a.tr0 = blk[O_TR0 + lane];
}

// write-back of one lane: the dynamic fields always, goal / static fields only when they changed; all = a world is
// (re)set: every field, and the countdown as it stands becomes the reference point n0 of the new episode
__device__ __forceinline__ void store_agent(double* blk, int lane, const Agent& a, bool goal_too, bool all) {
  blk[O_PX + lane] = a.px; blk[O_PY + lane] = a.py; blk[O_HD + lane] = a.hd;
  blk_spd(blk)[lane] = a.spd;
  blk_meta(blk)[lane] = pack_meta(a.flags, a.policy, a.cd);
  if (goal_too || all) { blk[O_GX + lane] = a.gx; blk[O_GY + lane] = a.gy; }
  if (all) {
    blk[O_RAD + lane] = a.rad; blk[O_PS + lane] = a.ps;
    blk[O_TR0 + lane] = a.tr0; blk_n0(blk)[lane] = a.cd;
  }
}

__device__ __forceinline__ void zero_agent(Agent& a) {
  a.px = a.py = a.hd = a.vx = a.vy = a.gx = a.gy = a.rad = a.ps = a.tr0 = 0.0;
  a.spd = 0.f; a.flags = 0; a.cd = 0; a.policy = 0;
}

// ---- the step body shared by every step kernel (generic, one-shot, streaming) ------------------------------------------
// _take_action (collision_avoidance_env.py:217-252): every agent picks its float32 command from the pre-step state
// (LearningPolicyGA3C.py:13-27 + action table, LearningPolicy.py:13-33, NonCooperativePolicy.py:9-22,
// StaticPolicy.py:9-23), then all agents move (Agent.take_action agent.py:190-238, UnicycleDynamics.step
// dynamics/UnicycleDynamics.py:14-47, _check_if_at_goal agent.py:148-151, time budget agent.py:232-236).
// `act` is the lane's discrete action, `g` its flat (world, agent) index (continuous actions are read from p.cont).
// kGen = false (the production instantiations of the specialised kernels): no continuous-action array, game_over =
// all learning agents done — the launcher picks the general instantiation for anything else.
template <bool kGen = true>
__device__ __forceinline__ void step_take_action(const Params& p, Agent& a, int act, size_t g, bool valid) {
  const bool was_done = (a.flags & CA_F_DONE_MASK) != 0;
  float cmd_speed = 0.f, cmd_dh = 0.f;  // all_actions is float32 (:238)
  if (valid && !was_done) {
    if (a.policy == CA_POLICY_LEARNING_GA3C) {
      const int k = act < 0 ? 0 : (act > 10 ? 10 : act);
      cmd_speed = (float)(a.ps * kActSpeed[k]);
      cmd_dh = (float)kActDhead[k];
    } else if (a.policy == CA_POLICY_NONCOOP) {
      // reads the ego heading of the pre-step state; it feeds the dynamics, so float64 atan2
      Ego e0;
      ego_axes(a.px, a.py, a.gx, a.gy, e0);
      cmd_speed = (float)a.ps;
      cmd_dh = (float)(-heading_ego_exact(e0, a.hd));
    } else if (a.policy == CA_POLICY_LEARNING) {
      double e0 = 0.0, e1 = 0.5;
      if (kGen && p.cont) { e0 = p.cont[2 * g]; e1 = p.cont[2 * g + 1]; }
      cmd_speed = (float)(a.ps * e0);
      cmd_dh = (float)(p.max_heading_change * (2. * e1 - 1.));
    } else if (a.policy == CA_POLICY_STATIC) {  // goal := pos
      a.gx = a.px;
      a.gy = a.py;
    }
  }
  if (!valid) return;
  if (was_done) {
    if (a.flags & CA_F_AT_GOAL) a.flags |= CA_F_WAS_AT_GOAL;
    if (a.flags & CA_F_IN_COLLISION) a.flags |= CA_F_WAS_IN_COLLISION;
    a.vx = 0.0;
    a.vy = 0.0;
    a.spd = 0.f;
  } else {
    const double speed = (double)cmd_speed;
    const double h = wrap_angle((double)cmd_dh + a.hd);
    double sh, ch;
    sincos_wrapped(h, sh, ch);
    a.px += speed * ch * p.dt;
    a.py += speed * sh * p.dt;
    a.vx = speed * ch;
    a.vy = speed * sh;
    a.spd = cmd_speed;
    a.hd = h;
    const double ex = a.px - a.gx, ey = a.py - a.gy;
    if (ex * ex + ey * ey <= p.thr_sq) a.flags |= CA_F_AT_GOAL; else a.flags &= ~CA_F_AT_GOAL;
    if (a.cd != kNever) a.cd -= 1;   // time_remaining_to_reach_goal -= dt ... (agent.py:232-236, see the layout notes)
    if (a.cd == 0) a.flags |= CA_F_RAN_OUT_OF_TIME;
  }
}

// _compute_rewards (:319-368; sets in_collision) and _check_which_agents_done (:411-439) for one lane, given the result
// of the all-pairs pass.  gmask = the lanes of this lane's world.  Returns the reward; dn = agent done, over = game_over.
template <bool kGen = true>
__device__ __forceinline__ float step_reward_done(const Params& p, Agent& a, bool valid, int i, bool coll, double nearest,
                                                  unsigned gmask, bool& dn, bool& over) {
  // float32 is what leaves the kernel; rounding is monotone, so clipping after the rounding gives the same float as
  // clipping the float64 value first (np.clip at :364) and rounding then
  float r = p.r_step_f;
  if (valid) {
    if (a.flags & CA_F_AT_GOAL) {
      if (!(a.flags & CA_F_WAS_AT_GOAL)) r = p.r_goal_f;  // goal beats collision; in_collision is not set
    } else if (!(a.flags & CA_F_WAS_IN_COLLISION)) {
      if (coll) {
        r = p.r_coll_f;
        a.flags |= CA_F_IN_COLLISION;
      } else if (nearest <= p.close_range) {
        r = (float)(-0.1 - nearest / 2.);
      }
    }
    r = fminf(fmaxf(r, p.r_min_f), p.r_max_f);
    if (kGen && p.over_mode == CA_OVER_FIRST_AGENT_DONE && i > 0) r = 0.f;  // rewards = rewards[0] (:365-366)
  } else {
    r = 0.f;
  }
  dn = valid ? (a.flags & CA_F_DONE_MASK) != 0 : true;
  const bool learning = valid && (a.policy == CA_POLICY_LEARNING_GA3C || a.policy == CA_POLICY_LEARNING);
  bool blocks_over;  // this agent keeps the episode alive
  if (kGen && p.over_mode == CA_OVER_ALL_DONE) blocks_over = valid && !dn;
  else if (kGen && p.over_mode == CA_OVER_FIRST_AGENT_DONE) blocks_over = valid && i == 0 && !dn;
  else blocks_over = learning && !dn;
  const unsigned alive = __ballot_sync(kFull, blocks_over) & gmask;
  over = alive == 0u;
  return r;
}

// Shared-memory carve-up of one CTA.
struct Smem {
  float* tile;   // [kWarps*wpw][A][L] observation rows of this CTA's worlds
  double* kq;    // [A][kBlock] rint(100 * dist_2_other) of (lane, other j)
  double* kp;    // [A][kBlock] p_orth
  double* kd;    // [A][kBlock] centre distance
  double* kt;    // [A][kBlock] time to impact (only with CA_SORT_TIME_TO_IMPACT)
};

// All-pairs pass for the lane's agent i against every j of its world:
//   kCollide: collision flag and nearest gap (CollisionAvoidanceEnv._check_for_collisions, :370-409)
//   always:   sensor keys into shared memory (OtherAgentsStatesSensor.sense first loop, :72-103)
template <bool kCollide>
__device__ __forceinline__ void pair_pass(const Params& p, const Smem& sm, const Agent& a, const Ego& e, bool valid,
                                          int n, int i, int base, int tid, bool& coll, double& nearest) {
  coll = false;
  nearest = INFINITY;
  for (int j = 0; j < p.A; ++j) {
    const int src = (base + j) & 31;
    const double xj = shfl_d(a.px, src), yj = shfl_d(a.py, src), rj = shfl_d(a.rad, src);
    double vxj = 0, vyj = 0;
    if (p.sort_method == CA_SORT_TIME_TO_IMPACT) { vxj = shfl_d(a.vx, src); vyj = shfl_d(a.vy, src); }
    if (valid && j < n && j != i) {
      const double rx = xj - a.px, ry = yj - a.py;
      const double d = sqrt(rx * rx + ry * ry);  // l2norm / vec2_l2_norm (util.py:8-12,106-112)
      if (kCollide) {
        const double R = a.rad + rj;
        if (d <= R) coll = true;
        if (j > i) nearest = fmin(nearest, d - R);  // only the lower index is updated (:393)
      }
      const double d2o = d - a.rad - rj;
      const int at = j * kBlock + tid;
      sm.kq[at] = rint(d2o * 100.0);
      sm.kp[at] = dot2(rx, ry, -e.pry, e.prx);
      sm.kd[at] = d;
      if (p.sort_method == CA_SORT_TIME_TO_IMPACT)
        sm.kt[at] = time_to_impact(a.px, a.py, a.vx, a.vy, xj, yj, vxj, vyj, a.rad + rj);
    }
  }
}

// Rank the others, clip to M, order per sort method and write the lane's observation row into the tile
// (OtherAgentsStatesSensor.sense second loop :105-144; dense row of GCA/envs/wrappers.py:130-139).
__device__ __forceinline__ void write_obs_row(const Params& p, const Smem& sm, const Agent& a, const Ego& e,
                                              bool world_ok, bool valid, int n, int i, int base, int tid,
                                              float* row, int32_t* sidx_row) {
  const int M = p.M;
  const bool tti = p.sort_method == CA_SORT_TIME_TO_IMPACT;
  const int mode1 = tti ? 2 : 0;                                   // key of the first (clipping) sort
  const int mode2 = p.sort_method == CA_SORT_CLOSEST_LAST ? 1 : mode1;  // key of the final order
  const bool horizon = isfinite(p.sensing_horizon);
  // candidates: others inside the sensing horizon (config.py:76, sensor :92-94)
  unsigned cand = 0;
  if (valid)
    for (int j = 0; j < n; ++j)
      if (j != i && !(horizon && sm.kd[j * kBlock + tid] > p.sensing_horizon)) cand |= 1u << j;
  int count = __popc(cand);
  unsigned sel = cand;
  const bool clip = count > M;
  if (clip) {  // first sort + clip to the M closest
    sel = 0;
    for (int j = 0; j < n; ++j) {
      if (!((cand >> j) & 1u)) continue;
      const int aj = j * kBlock + tid;
      const double qj = sm.kq[aj], pj = sm.kp[aj], tj = tti ? sm.kt[aj] : 0.0;
      int rank = 0;
      for (int k = 0; k < n; ++k) {
        if (k == j || !((cand >> k) & 1u)) continue;
        const int ak = k * kBlock + tid;
        rank += key_before(mode1, sm.kq[ak], sm.kp[ak], tti ? sm.kt[ak] : 0.0, k, qj, pj, tj, j) ? 1 : 0;
      }
      if (rank < M) sel |= 1u << j;
    }
    count = M;
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
  // second loop: every lane takes part in the shuffles, selected others land in their slot
  for (int j = 0; j < p.A; ++j) {
    const int src = (base + j) & 31;
    const double xj = shfl_d(a.px, src), yj = shfl_d(a.py, src), rj = shfl_d(a.rad, src);
    const double vxj = shfl_d(a.vx, src), vyj = shfl_d(a.vy, src);
    if (!((sel >> j) & 1u)) continue;
    const int aj = j * kBlock + tid;
    const double qj = sm.kq[aj], pj = sm.kp[aj], tj = tti ? sm.kt[aj] : 0.0;
    int slot = 0;
    for (int k = 0; k < n; ++k) {
      if (k == j || !((sel >> k) & 1u)) continue;
      const int ak = k * kBlock + tid;
      slot += key_before(mode2, sm.kq[ak], sm.kp[ak], tti ? sm.kt[ak] : 0.0, k, qj, pj, tj, j) ? 1 : 0;
    }
    // observation-only projections: shared definitions (obs_pprl / obs_vel), so that all step kernels agree bit for
    // bit; p_orth is the float64 tie-break key
    const float prxf = (float)e.prx, pryf = (float)e.pry;
    const float rjf = (float)rj, raf = (float)a.rad;
    float* s = row + CA_OBS_HOST_LEN + CA_OBS_OTHER_LEN * slot;
    s[0] = obs_pprl(xj - a.px, yj - a.py, e);
    s[1] = (float)pj;
    obs_vel((float)vxj, (float)vyj, prxf, pryf, s[2], s[3]);
    s[4] = rjf;
    s[5] = raf + rjf;
    s[6] = (float)(sm.kd[aj] - a.rad - rj);
    if (sidx_row) sidx_row[slot] = j;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* ptr) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(ptr));
}

// CTA-wide copy of the observation tile to global memory.
__device__ __forceinline__ void store_tile(const Params& p, const float* tile, long first_world, int tid) {
  const long worlds_left = (long)p.W - first_world;
  if (worlds_left <= 0) return;
  const int tile_worlds = kWarps * p.wpw;
  const int nw = worlds_left < tile_worlds ? (int)worlds_left : tile_worlds;
  const int nfloats = nw * p.A * p.L;
  float* dst = p.obs + (size_t)first_world * p.A * p.L;
  if (p.use_bulk_store && nw == tile_worlds) {
    // TMA bulk store (UBLKCP): generic-proxy writes to smem must be made visible to the async proxy first
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(tile)),
                   "r"(nfloats * 4)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  __syncthreads();
  for (int q = tid; q < nfloats; q += kBlock) dst[q] = tile[q];
}

template <bool kStep>
__global__ void __launch_bounds__(kBlock) ca_world_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  Smem sm;
  sm.tile = reinterpret_cast<float*>(smem_raw);
  const int tile_bytes = ((p.tile_floats * 4 + 127) / 128) * 128;
  sm.kq = reinterpret_cast<double*>(smem_raw + tile_bytes);
  sm.kp = sm.kq + p.A * kBlock;
  sm.kd = sm.kp + p.A * kBlock;
  sm.kt = sm.kd + p.A * kBlock;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int A = p.A;
  const int wl = lane / A;      // world slot inside the warp
  const int i = lane - wl * A;  // agent index inside the world
  const int base = wl * A;      // first lane of the world's group
  const long first_world_cta = (long)blockIdx.x * kWarps * p.wpw;
  const long w = first_world_cta + (long)warp * p.wpw + wl;
  const bool world_ok = wl < p.wpw && w < p.W;
  const long chunk = (long)blockIdx.x * kWarps + warp;
  double* const blk = blk_ptr(p.s, chunk);
  double* const blk0 = blk_ptr(p.s0, chunk);
  pdl_wait();
  pdl_launch_dependents();
  int n = world_ok ? blk_nag(blk)[wl] : 0;
  bool valid = world_ok && i < n;
  const size_t g = world_ok ? (size_t)w * A + i : 0;
  const unsigned gmask = (A >= 32 ? kFull : ((1u << A) - 1u)) << (base & 31);

  float* row = sm.tile + ((size_t)(warp * p.wpw + wl) * A + i) * p.L;
  int32_t* sidx_row = (p.sidx && world_ok) ? p.sidx + g * p.M : nullptr;

  Agent a;
  if (valid) load_agent<!kStep>(blk, lane, a); else zero_agent(a);  // a step overwrites the velocity before using it

  bool do_reset;  // world reloads its injected initial state
  if (kStep) step_take_action(p, a, (valid && a.policy == CA_POLICY_LEARNING_GA3C) ? p.actions[g] : 0, g, valid);

  Ego e = ego_frame(a.px, a.py, a.gx, a.gy, a.hd);

  if (kStep) {
    bool coll;
    double nearest;
    pair_pass<true>(p, sm, a, e, valid, n, i, base, tid, coll, nearest);
    bool dn, over;
    const float r = step_reward_done(p, a, valid, i, coll, nearest, gmask, dn, over);
    if (world_ok) {
      p.reward[g] = r;
      p.done[g] = dn ? 1 : 0;
      if (i == 0) p.over[w] = over ? 1 : 0;
    }
    do_reset = world_ok && over && p.auto_reset;
    write_obs_row(p, sm, a, e, world_ok, valid, n, i, base, tid, row, sidx_row);
  } else {
    do_reset = world_ok && (p.mask == nullptr || p.mask[w] != 0);
  }

  // ---- reset path (CollisionAvoidanceEnv.reset :196-215; DummyVecEnv auto-reset when called from step).
  // Lanes whose world does not reset still take part in the shuffles but compute and write nothing.
  if (!kStep || __any_sync(kFull, do_reset)) {
    const bool again = kStep ? do_reset : true;  // this lane (re)computes its observation
    if (do_reset) {
      n = blk_nag(blk0)[wl];
      valid = i < n;
      if (i == 0) { blk_nag(blk)[wl] = n; p.consumed[w] = 1; }
      if (valid) load_agent_reset(blk0, lane, a); else zero_agent(a);  // a snapshot is at rest
      e = ego_frame(a.px, a.py, a.gx, a.gy, a.hd);
    }
    bool c_unused;
    double n_unused;
    pair_pass<false>(p, sm, a, e, valid && again, n, i, base, tid, c_unused, n_unused);
    write_obs_row(p, sm, a, e, world_ok && again, valid && again, n, i, base, tid, row, sidx_row);
  }

  // ---- state write-back (coalesced; goal only changes for static agents / on reset)
  if ((valid && kStep) || do_reset)  // a reset rewrites every slot of the world (the scenario may have changed)
    store_agent(blk, lane, a, a.policy == CA_POLICY_STATIC, do_reset);

  store_tile(p, sm.tile, first_world_cta, tid);
}

// n-step discounted return, ProcessAgent._accumulate_rewards recursion (GA3C/ProcessAgent.py:71-76)
__global__ void nstep_returns_kernel(const float* __restrict__ reward, const float* __restrict__ bootstrap,
                                     float* __restrict__ out, int T, int N, float gamma) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= N) return;
  float R = bootstrap[k];
  for (int t = T - 1; t >= 0; --t) {
    R = gamma * R + reward[(size_t)t * N + k];
    out[(size_t)t * N + k] = R;
  }
}

}  // namespace ca
