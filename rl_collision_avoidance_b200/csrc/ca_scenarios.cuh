// ca_scenarios.cuh — on-device scenario generator: fills the reset snapshot of worlds with fresh random test
// cases so that an auto-reset starts a NEW scenario without a host round trip.
//
// Reference (distributional parity; NumPy's global MT19937 order cannot be reproduced in parallel):
//   get_testcase_random            GCA/envs/test_cases.py:95-118   (num_agents ~ U{2..A}, side length by agent count)
//   generate_rand_test_case_multi  GCA/envs/policies/CADRL/scripts/multi/gen_rand_testcases.py:104-135
//                                  (15 % swap, 15 % circle, 70 % random)
//   generate_rand_case             :137-226 (rejection sampling incl. "a straight line must NOT already be a
//                                  solution": a candidate is rejected when it has no conflict with any earlier agent)
//   generate_swap_case / generate_circle_case   :313-416
//   if_permitStraightLineSoln / find_dist_between_segs / distPointToSegment   :47-102, :418-437
//   cadrl_test_case_to_agents      GCA/envs/test_cases.py:263-326 (policy mix with an ensured learner, heading U(-pi,pi))
//   Agent.reset                    GCA/envs/agent.py:98-103 (time budget)
// One WARP per world.  The reference's rejection loops are sequential ("draw a candidate, test it against the agents
// placed so far, retry"); here the 32 lanes draw and test 32 consecutive candidates of that sequence at once (lane l
// uses the side length / circle radius the l-th retry would have seen) and the lowest accepted lane wins — the same
// distribution as the sequential loop, one round instead of ~10 dependent iterations of float64 sqrt/div chains.  The
// agents placed so far live in shared memory (broadcast reads); policy, heading and time budget of agent i are drawn
// and written by lane i.  Philox4x32-10 stream per (seed, world, lane).
#pragma once
#include <curand_kernel.h>

#include "ca_kernels.cuh"

namespace ca {

struct ScenarioParams {
  ca_scenario_config c;
  StateBlocks s0;
  uint8_t* consumed;  // [W] set by the step kernels when a world resets; nullptr = regenerate every world
  int W, A;
  int only_consumed;
  double dt, thr, max_time_ratio;
  double step_dt;     // dt of the steps (ca_set_dt): what the time budget is counted down in
  unsigned long long seed, offset;
};

struct Rng {
  curandStatePhilox4_32_10_t st;
  __device__ double u() { return 1.0 - curand_uniform_double(&st); }  // [0, 1) like np.random.rand()
};

__device__ inline double norm2d(double x, double y) { return sqrt(x * x + y * y); }

// The "a straight line must not already be a solution" filter (if_permitStraightLineSoln and helpers) only shapes the
// DISTRIBUTION of accepted test cases — nothing downstream depends on its last bits — so it is evaluated in float32
// (hardware sqrt / division instead of ~25-instruction float64 sequences; the filter is the generator's critical path).
// The hard spacing constraints (collides) stay in float64.
__device__ __forceinline__ float norm2f(float x, float y) { return sqrtf(x * x + y * y); }

// distPointToSegment(p1, p2, p3), :84-102
__device__ __forceinline__ float dist_point_segment(float p1x, float p1y, float p2x, float p2y, float p3x, float p3y) {
  const float dx = p2x - p1x, dy = p2y - p1y;
  const float nd = norm2f(dx, dy);
  float u = 0.f;
  if (!(nd < 1e-5f)) u = (dx * (p3x - p1x) + dy * (p3y - p1y)) / (nd * nd);
  u = fmaxf(0.f, fminf(u, 1.f));
  return norm2f(p3x - (p1x + u * dx), p3y - (p1y + u * dy));
}

// find_dist_between_segs(x1, x2, y1, y2) for one end point, :47-80
__device__ __forceinline__ float dist_between_segs(float x1x, float x1y, float x2x, float x2y, float y1x, float y1y,
                                                   float y2x, float y2y) {
  const float end_dist = norm2f(x2x - y2x, x2y - y2y);
  float critical = end_dist;
  const float zx = (x2x - x1x) - (y2x - y1x), zy = (x2y - x1y) - (y2y - y1y);
  if (norm2f(zx, zy) > 0.f) {
    const float t = -((x1x - y1x) * zx + (x1y - y1y) * zy) / (zx * zx + zy * zy);
    if (t > 0.f && t < 1.f)
      critical = norm2f(x1x + (x2x - x1x) * t - y1x - (y2x - y1x) * t, x1y + (x2y - x1y) * t - y1y - (y2y - y1y) * t);
  }
  return fminf(end_dist, critical);
}

// if_permitStraightLineSoln(x1, x2, s1, y1, y2, s2, radius), :418-437
__device__ __forceinline__ bool permits_straight_line(float x1x, float x1y, float x2x, float x2y, float s1, float y1x,
                                                      float y1y, float y2x, float y2y, float s2, float radius) {
  const float t1 = norm2f(x2x - x1x, x2y - x1y) / s1;
  const float t2 = norm2f(y2x - y1x, y2y - y1y) / s2;
  float xcx, xcy, ycx, ycy;
  if (t1 < t2) {
    xcx = x2x; xcy = x2y;
    ycx = y1x + t1 * (y2x - y1x) / t2; ycy = y1y + t1 * (y2y - y1y) / t2;
    if (dist_point_segment(ycx, ycy, y2x, y2y, xcx, xcy) < radius) return false;
  } else {
    xcx = x1x + t2 * (x2x - x1x) / t1; xcy = x1y + t2 * (x2y - x1y) / t1;
    ycx = y2x; ycy = y2y;
    if (dist_point_segment(xcx, xcy, x2x, x2y, ycx, ycy) < radius) return false;
  }
  const float start_dist = norm2f(x1x - y1x, x1y - y1y);
  const float end_dist = norm2f(xcx - ycx, xcy - ycy);
  const float mid_dist = dist_between_segs(x1x, x1y, xcx, xcy, y1x, y1y, ycx, ycy);
  return !(fminf(start_dist, fminf(end_dist, mid_dist)) < radius);
}

constexpr int kGenWarps = 4;

struct GenAgents {  // test case rows [px, py, gx, gy, pref_speed, radius] of one world, in shared memory
  double px[CA_MAX_AGENTS], py[CA_MAX_AGENTS], gx[CA_MAX_AGENTS], gy[CA_MAX_AGENTS], sp[CA_MAX_AGENTS], rd[CA_MAX_AGENTS];
};

__device__ __forceinline__ double bcast_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// 1.01^e by repeated multiplication (e is small: the reference multiplies once per rejected draw)
__device__ __forceinline__ double grow_1p01(int e) {
  double g = 1.0;
  for (int k = 0; k < e; ++k) g *= 1.01;
  return g;
}

__global__ void __launch_bounds__(kGenWarps * 32) generate_scenarios_kernel(const ScenarioParams p) {
  __shared__ GenAgents sh[kGenWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w = blockIdx.x * kGenWarps + warp;
  if (w >= p.W) return;                                                 // warp-uniform
  if (p.only_consumed && p.consumed && !p.consumed[w]) return;          // warp-uniform
  __syncwarp();
  if (p.consumed && lane == 0) p.consumed[w] = 0;
  const ca_scenario_config& c = p.c;
  const int A = p.A;
  GenAgents& ag = sh[warp];
  Rng rng;
  curand_init(p.seed, (unsigned long long)w * 32ull + lane, p.offset, &rng.st);
  const double close_range = 0.2;  // gen_rand_testcases GETTING_CLOSE_RANGE (global_var.py:8)

  // world-level draws: lane 0 draws, everybody gets a copy
  int n = c.min_agents + (int)(bcast_d(rng.u(), 0) * (c.max_agents - c.min_agents + 1));  // np.random.randint(2, A+1)
  n = n < 1 ? 1 : (n > A ? A : n);
  const double u_side = bcast_d(rng.u(), 0);
  double side = n < c.side_split_agents ? c.side_small_lo + u_side * (c.side_small_hi - c.side_small_lo)
                                        : c.side_large_lo + u_side * (c.side_large_hi - c.side_large_lo);
  const double kind = bcast_d(rng.u(), 0);

  // radius and preferred speed of agent i: drawn by lane i (generate_rand_case :145-150, max of two speed draws)
  {
    const double r0 = rng.u(), s1 = rng.u(), s2 = rng.u();
    if (lane < n) {
      ag.rd[lane] = (c.radius_hi - c.radius_lo) * r0 + c.radius_lo;
      ag.sp[lane] = fmax((c.speed_hi - c.speed_lo) * s1 + c.speed_lo, (c.speed_hi - c.speed_lo) * s2 + c.speed_lo);
    }
  }
  __syncwarp();

  auto collides = [&](int i, double sx, double sy, double ex, double ey) {
    for (int j = 0; j < i; ++j) {   // |start_i - start_j| < r_i + r_j + 0.2 (same for the goals), compared squared
      const double lim = ag.rd[j] + ag.rd[i] + close_range, lim2 = lim * lim;
      const double ax = sx - ag.px[j], ay = sy - ag.py[j], bx = ex - ag.gx[j], by = ey - ag.gy[j];
      if (ax * ax + ay * ay < lim2 || bx * bx + by * by < lim2) return true;
    }
    return false;
  };
  // the lowest accepted lane's candidate becomes agent i; returns that lane or -1
  auto commit = [&](int i, bool ok, double sx, double sy, double ex, double ey) {
    const unsigned acc = __ballot_sync(0xffffffffu, ok);
    if (acc == 0u) return -1;
    const int k = __ffs(acc) - 1;
    if (lane == k) { ag.px[i] = sx; ag.py[i] = sy; ag.gx[i] = ex; ag.gy[i] = ey; }
    __syncwarp();
    return k;
  };
  const int kMaxRounds = 100000 / 32;

  if (kind < c.p_swap || kind < c.p_swap + c.p_circle) {  // generate_swap_case :313-366 / generate_circle_case :369-416
    const bool swap = kind < c.p_swap;
    const double r_min = n / 2.0;
    double r = bcast_d(rng.u(), 0) * 2.0 + r_min;
    double r_swap = 0.0, off_y = 0.0;
    if (swap) {
      r_swap = 1.5 + bcast_d(rng.u(), 0) * 2.0;
      off_y = 1.0 + r_min + bcast_d(rng.u(), 0) * 2.0;
      if (bcast_d(rng.u(), 0) > 0.5) off_y = -off_y;
    }
    for (int i = 0; i < n; ++i) {
      if (swap && i < 2) {
        if (lane == 0) {
          ag.px[i] = i == 0 ? -r_swap : r_swap; ag.py[i] = 0; ag.gx[i] = i == 0 ? r_swap : -r_swap; ag.gy[i] = 0;
        }
        __syncwarp();
        continue;
      }
      // the reference grows r by 1 % after every 11 rejected draws: candidate m of this agent sees r * 1.01^(m / 11)
      for (int round = 0; round < kMaxRounds; ++round) {
        const int m = round * 32 + lane;
        const double rr = r * grow_1p01(m / 11);
        const double a0 = rng.u() * 2 * kPi - kPi;   // goal angle = pi + a0: the antipodal point (to ~1e-16)
        double sa, ca_;
        sincos(a0, &sa, &ca_);
        const double sx = rr * ca_, sy = rr * sa + off_y, ex = -(rr * ca_), ey = -(rr * sa) + off_y;
        const int k = commit(i, !collides(i, sx, sy, ex, ey), sx, sy, ex, ey);
        if (k >= 0) { r *= grow_1p01((round * 32 + k) / 11); break; }
      }
    }
  } else {  // generate_rand_case :137-226
    const double g_lane = grow_1p01(lane + 1), g_round = grow_1p01(32);
    for (int i = 0; i < n; ++i) {
      for (int round = 0; round < kMaxRounds; ++round) {
        const double my_side = side * g_lane;   // side *= 1.01 before every draw
        const double sx = my_side * 2 * rng.u() - my_side, sy = my_side * 2 * rng.u() - my_side;
        const double ex = my_side * 2 * rng.u() - my_side, ey = my_side * 2 * rng.u() - my_side;
        bool ok = !collides(i, sx, sy, ex, ey);
        if (ok && i >= 1) {  // reject if every earlier agent permits a straight-line solution (no interaction), :197-214
          bool all_permit = true;
          for (int j = 0; j < i; ++j)
            if (!permits_straight_line((float)ag.px[j], (float)ag.py[j], (float)ag.gx[j], (float)ag.gy[j], (float)ag.sp[j],
                                       (float)sx, (float)sy, (float)ex, (float)ey, (float)ag.sp[i],
                                       (float)(ag.rd[j] + ag.rd[i] + close_range))) { all_permit = false; break; }
          ok = !all_permit;
        }
        ok = ok && norm2d(sx - ex, sy - ey) > my_side * 0.5;
        const int k = commit(i, ok, sx, sy, ex, ey);
        if (k >= 0) { side *= grow_1p01(k + 1); break; }
        side *= g_round;
      }
    }
  }

  // policy mix with an ensured learner, cadrl_test_case_to_agents (test_cases.py:275-293): lane i decides agent i
  const double u_pol = rng.u();
  int pol = u_pol < c.p_noncoop ? CA_POLICY_NONCOOP : (u_pol < c.p_noncoop + c.p_learning ? CA_POLICY_LEARNING_GA3C : CA_POLICY_STATIC);
  const bool has_learner = (__ballot_sync(0xffffffffu, lane < n && pol == CA_POLICY_LEARNING_GA3C)) != 0u;
  const int pick = (int)(bcast_d(rng.u(), 0) * n) % n;
  if (c.ensure_learner && !has_learner && lane == pick) pol = CA_POLICY_LEARNING_GA3C;
  const double heading = rng.u() * 2 * kPi - kPi;  // np.random.uniform(-pi, pi), test_cases.py:315

  long chunk;
  int lane0, wl;
  slot_of(w, 0, A, chunk, lane0, wl);
  double* blk = blk_ptr(p.s0, chunk);
  if (lane == 0) blk_nag(blk)[wl] = n;
  if (lane < A) {  // lane i writes agent slot i
    const int i = lane, dst = lane0 + i;
    const bool live = i < n;
    double t0 = 0.0;
    if (live) {
      t0 = p.max_time_ratio * ((norm2d(ag.px[i] - ag.gx[i], ag.py[i] - ag.gy[i]) - p.thr) / ag.sp[i]);
      if (!(t0 > p.dt)) t0 = p.dt;
    }
    blk[O_PX + dst] = live ? ag.px[i] : 0.0; blk[O_PY + dst] = live ? ag.py[i] : 0.0;
    blk[O_GX + dst] = live ? ag.gx[i] : 0.0; blk[O_GY + dst] = live ? ag.gy[i] : 0.0;
    blk[O_HD + dst] = live ? heading : 0.0;
    const unsigned cd = live ? countdown_steps(t0, p.step_dt) : 0u;   // the budget as a number of steps (ca_kernels.cuh)
    blk_spd(blk)[dst] = 0.f; blk[O_TR0 + dst] = t0; blk_n0(blk)[dst] = cd;
    blk[O_RAD + dst] = live ? ag.rd[i] : 0.0; blk[O_PS + dst] = live ? ag.sp[i] : 0.0;
    blk_meta(blk)[dst] = pack_meta(0u, live ? pol : 0, cd);
  }
}

}  // namespace ca
