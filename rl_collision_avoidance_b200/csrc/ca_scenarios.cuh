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
// One thread per world, Philox4x32-10 stream (seed, world), scalar rejection loops like the reference.
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
  unsigned long long seed, offset;
};

struct Rng {
  curandStatePhilox4_32_10_t st;
  __device__ double u() { return 1.0 - curand_uniform_double(&st); }  // [0, 1) like np.random.rand()
};

__device__ inline double norm2d(double x, double y) { return sqrt(x * x + y * y); }

// distPointToSegment(p1, p2, p3), :84-102
__device__ inline double dist_point_segment(double p1x, double p1y, double p2x, double p2y, double p3x, double p3y) {
  const double dx = p2x - p1x, dy = p2y - p1y;
  const double nd = norm2d(dx, dy);
  double u = 0.0;
  if (!(nd < 1e-5)) u = (dx * (p3x - p1x) + dy * (p3y - p1y)) / (nd * nd);
  u = fmax(0.0, fmin(u, 1.0));
  return norm2d(p3x - (p1x + u * dx), p3y - (p1y + u * dy));
}

// find_dist_between_segs(x1, x2, y1, y2) for one end point, :47-80
__device__ inline double dist_between_segs(double x1x, double x1y, double x2x, double x2y, double y1x, double y1y,
                                           double y2x, double y2y) {
  const double end_dist = norm2d(x2x - y2x, x2y - y2y);
  double critical = end_dist;
  const double zx = (x2x - x1x) - (y2x - y1x), zy = (x2y - x1y) - (y2y - y1y);
  if (norm2d(zx, zy) > 0) {
    const double t = -((x1x - y1x) * zx + (x1y - y1y) * zy) / (zx * zx + zy * zy);
    if (t > 0 && t < 1.0)
      critical = norm2d(x1x + (x2x - x1x) * t - y1x - (y2x - y1x) * t, x1y + (x2y - x1y) * t - y1y - (y2y - y1y) * t);
  }
  return fmin(end_dist, critical);
}

// if_permitStraightLineSoln(x1, x2, s1, y1, y2, s2, radius), :418-437
__device__ inline bool permits_straight_line(double x1x, double x1y, double x2x, double x2y, double s1, double y1x,
                                             double y1y, double y2x, double y2y, double s2, double radius) {
  const double t1 = norm2d(x2x - x1x, x2y - x1y) / s1;
  const double t2 = norm2d(y2x - y1x, y2y - y1y) / s2;
  double xcx, xcy, ycx, ycy;
  if (t1 < t2) {
    xcx = x2x; xcy = x2y;
    ycx = y1x + t1 * (y2x - y1x) / t2; ycy = y1y + t1 * (y2y - y1y) / t2;
    if (dist_point_segment(ycx, ycy, y2x, y2y, xcx, xcy) < radius) return false;
  } else {
    xcx = x1x + t2 * (x2x - x1x) / t1; xcy = x1y + t2 * (x2y - x1y) / t1;
    ycx = y2x; ycy = y2y;
    if (dist_point_segment(xcx, xcy, x2x, x2y, ycx, ycy) < radius) return false;
  }
  const double start_dist = norm2d(x1x - y1x, x1y - y1y);
  const double end_dist = norm2d(xcx - ycx, xcy - ycy);
  const double mid_dist = dist_between_segs(x1x, x1y, xcx, xcy, y1x, y1y, ycx, ycy);
  return !(fmin(start_dist, fmin(end_dist, mid_dist)) < radius);
}

__global__ void __launch_bounds__(128) generate_scenarios_kernel(const ScenarioParams p) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= p.W) return;
  if (p.only_consumed && p.consumed && !p.consumed[w]) return;
  if (p.consumed) p.consumed[w] = 0;
  const ca_scenario_config& c = p.c;
  const int A = p.A;
  Rng rng;
  curand_init(p.seed, (unsigned long long)w, p.offset, &rng.st);

  // test case rows [px, py, gx, gy, pref_speed, radius]
  double px[CA_MAX_AGENTS], py[CA_MAX_AGENTS], gx[CA_MAX_AGENTS], gy[CA_MAX_AGENTS], sp[CA_MAX_AGENTS], rd[CA_MAX_AGENTS];

  int n = c.min_agents + (int)(rng.u() * (c.max_agents - c.min_agents + 1));  // np.random.randint(2, A+1)
  n = n < 1 ? 1 : (n > A ? A : n);
  double side = n < c.side_split_agents ? c.side_small_lo + rng.u() * (c.side_small_hi - c.side_small_lo)
                                        : c.side_large_lo + rng.u() * (c.side_large_hi - c.side_large_lo);
  const double kind = rng.u();
  const double close_range = 0.2;  // gen_rand_testcases GETTING_CLOSE_RANGE (global_var.py:8)

  auto draw_size_speed = [&](int i) {
    rd[i] = (c.radius_hi - c.radius_lo) * rng.u() + c.radius_lo;
    const double s1 = (c.speed_hi - c.speed_lo) * rng.u() + c.speed_lo;
    const double s2 = (c.speed_hi - c.speed_lo) * rng.u() + c.speed_lo;
    sp[i] = fmax(s1, s2);
  };
  auto collides = [&](int i, double sx, double sy, double ex, double ey) {
    for (int j = 0; j < i; ++j) {
      const double lim = rd[j] + rd[i] + close_range;
      if (norm2d(sx - px[j], sy - py[j]) < lim) return true;
      if (norm2d(ex - gx[j], ey - gy[j]) < lim) return true;
    }
    return false;
  };

  if (kind < c.p_swap) {  // generate_swap_case :313-366
    const double r_min = n / 2.0;
    double r = rng.u() * 2.0 + r_min;
    const double r_swap = 1.5 + rng.u() * 2.0;
    double off_y = 1.0 + r_min + rng.u() * 2.0;
    if (rng.u() > 0.5) off_y = -off_y;
    for (int i = 0; i < n; ++i) {
      draw_size_speed(i);
      if (i == 0) { px[i] = -r_swap; py[i] = 0; gx[i] = r_swap; gy[i] = 0; continue; }
      if (i == 1) { px[i] = r_swap; py[i] = 0; gx[i] = -r_swap; gy[i] = 0; continue; }
      int counter = 0;
      for (int it = 0; it < 100000; ++it) {
        if (counter > 10) { r *= 1.01; counter = 0; }
        const double a0 = rng.u() * 2 * kPi - kPi, a1 = kPi + a0;
        const double sx = r * cos(a0), sy = r * sin(a0) + off_y, ex = r * cos(a1), ey = r * sin(a1) + off_y;
        px[i] = sx; py[i] = sy; gx[i] = ex; gy[i] = ey;
        if (!collides(i, sx, sy, ex, ey)) break;
        ++counter;
      }
    }
  } else if (kind < c.p_swap + c.p_circle) {  // generate_circle_case :369-416
    const double r_min = n / 2.0;
    double r = rng.u() * 2.0 + r_min;
    for (int i = 0; i < n; ++i) {
      draw_size_speed(i);
      int counter = 0;
      for (int it = 0; it < 100000; ++it) {
        if (counter > 10) { r *= 1.01; counter = 0; }
        const double a0 = rng.u() * 2 * kPi - kPi, a1 = kPi + a0;
        const double sx = r * cos(a0), sy = r * sin(a0), ex = r * cos(a1), ey = r * sin(a1);
        px[i] = sx; py[i] = sy; gx[i] = ex; gy[i] = ey;
        if (!collides(i, sx, sy, ex, ey)) break;
        ++counter;
      }
    }
  } else {  // generate_rand_case :137-226
    for (int i = 0; i < n; ++i) {
      draw_size_speed(i);
      for (int it = 0; it < 100000; ++it) {
        side *= 1.01;
        const double sx = side * 2 * rng.u() - side, sy = side * 2 * rng.u() - side;
        const double ex = side * 2 * rng.u() - side, ey = side * 2 * rng.u() - side;
        px[i] = sx; py[i] = sy; gx[i] = ex; gy[i] = ey;
        if (collides(i, sx, sy, ex, ey)) continue;
        if (i >= 1) {  // reject if every earlier agent permits a straight-line solution (no interaction)
          bool all_permit = true;
          for (int j = 0; j < i; ++j)
            if (!permits_straight_line(px[j], py[j], gx[j], gy[j], sp[j], sx, sy, ex, ey, sp[i],
                                       rd[j] + rd[i] + close_range)) { all_permit = false; break; }
          if (all_permit) continue;
        }
        if (norm2d(sx - ex, sy - ey) > side * 0.5) break;
      }
    }
  }

  // policy mix with an ensured learner, cadrl_test_case_to_agents (test_cases.py:275-293)
  int pol[CA_MAX_AGENTS];
  bool has_learner = false;
  for (int i = 0; i < n; ++i) {
    const double u = rng.u();
    pol[i] = u < c.p_noncoop ? CA_POLICY_NONCOOP : (u < c.p_noncoop + c.p_learning ? CA_POLICY_LEARNING_GA3C : CA_POLICY_STATIC);
    has_learner |= pol[i] == CA_POLICY_LEARNING_GA3C;
  }
  if (c.ensure_learner && !has_learner) pol[(int)(rng.u() * n) % n] = CA_POLICY_LEARNING_GA3C;

  long chunk;
  int lane0, wl;
  slot_of(w, 0, A, chunk, lane0, wl);
  double* blk = blk_ptr(p.s0, chunk);
  blk_nag(blk)[wl] = n;
  for (int i = 0; i < A; ++i) {
    const int lane = lane0 + i;
    const bool live = i < n;
    double t0 = 0.0;
    if (live) {
      t0 = p.max_time_ratio * ((norm2d(px[i] - gx[i], py[i] - gy[i]) - p.thr) / sp[i]);
      if (!(t0 > p.dt)) t0 = p.dt;
    }
    blk[O_PX + lane] = live ? px[i] : 0.0; blk[O_PY + lane] = live ? py[i] : 0.0;
    blk[O_GX + lane] = live ? gx[i] : 0.0; blk[O_GY + lane] = live ? gy[i] : 0.0;
    blk[O_HD + lane] = live ? (rng.u() * 2 * kPi - kPi) : 0.0;  // np.random.uniform(-pi, pi), test_cases.py:315
    blk[O_VX + lane] = 0.0; blk[O_VY + lane] = 0.0; blk[O_TR + lane] = t0;
    blk[O_RAD + lane] = live ? rd[i] : 0.0; blk[O_PS + lane] = live ? sp[i] : 0.0;
    blk_flags(blk)[lane] = 0; blk_policy(blk)[lane] = live ? (uint8_t)pol[i] : 0;
  }
}

}  // namespace ca
