/*
 * ca_oracle.c — TEST INFRASTRUCTURE.  CPU restatement (plain C, float64) of the reference's
 * CollisionAvoidanceEnv.step() path, used ONLY as the checker in tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  The product (libcastep.so) never links,
 * loads or calls anything in this file.
 *
 * Parity pinning: tests/test_oracle_golden.py checks this restatement against golden vectors
 * recorded from the UNMODIFIED reference (oracle/gen_golden.py, tests/golden/ *.npz): flags, done,
 * game_over and neighbour indices bit-exact, float64 state to 1e-9.
 *
 * It follows the reference's "Mode B" numerics: the commanded [speed, dheading]
 * is rounded to float32 (collision_avoidance_env.py:238) and everything else is IEEE double.
 * Citations: GCA = /root/reference/gym-collision-avoidance/gym_collision_avoidance.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no compiler-introduced FMA; the two
 * places where NumPy itself uses FMA are written with fma() explicitly, see np_dot2/np_norm2).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ca_step.h"

#define PI_D 3.141592653589793 /* np.pi */

typedef struct {
  /* dynamic (GCA/envs/agent.py:66-136) */
  double px, py, heading, vx, vy, t_rem, gx, gy;
  /* static */
  double radius, pref_speed;
  uint32_t flags;
  int32_t policy;
  /* ego frame, recomputed by update_ego_frame (GCA/envs/dynamics/Dynamics.py:24-41) */
  double dist_to_goal, heading_ego, prll_x, prll_y, orth_x, orth_y;
} agent_t;

typedef struct ca_oracle {
  ca_config cfg;
  int W, A, M, L;
  agent_t* agents;      /* [W][A] */
  agent_t* init_agents; /* [W][A] snapshot for reset */
  int32_t* n;           /* [W] live agent count */
  int32_t* n0;          /* [W] agent count of the reset snapshot */
  int initialised;
} ca_oracle;

/* x**2 on an np.float64 scalar goes through C pow() (numpy scalarmath); glibc pow(x,2) is not
 * always the correctly rounded x*x (measured: 0.08 % of random inputs differ by 1 ulp). */
static inline double sq(double x) { return pow(x, 2.0); }

/* np.dot on two length-2 float64 vectors -> OpenBLAS ddot; on the build container's x86-64 kernel
 * this equals fma(a1, b1, a0*b0) for 50000/50000 random pairs (probe recorded in DESIGN.md). */
static inline double np_dot2(double a0, double a1, double b0, double b1) { return fma(a1, b1, a0 * b0); }
/* np.linalg.norm of a length-2 vector = sqrt(dot(x, x)) with the same FMA accumulation. */
static inline double np_norm2(double x0, double x1) { return sqrt(fma(x1, x1, x0 * x0)); }

/* GCA/envs/util.py:132-137 */
static double wrap(double a) {
  while (a >= PI_D) a -= 2 * PI_D;
  while (a < -PI_D) a += 2 * PI_D;
  return a;
}

/* Actions table, GCA/envs/policies/GA3C_CADRL/network.py:13-16 (np.mgrid arithmetic; the value of row 3
 * is the mgrid result 0x1.0c152382d7366p-2, one ulp above pi/12 — irrelevant after the float32 cast). */
static const double ACT_SPEED[11] = {1.0, 1.0, 1.0, 1.0, 1.0, 0.5, 0.5, 0.5, 0.0, 0.0, 0.0};
static const double ACT_DHEAD[11] = {-0x1.0c152382d7365p-1, -0x1.0c152382d7365p-2, 0.0, 0x1.0c152382d7366p-2,
                                     0x1.0c152382d7365p-1,  -0x1.0c152382d7365p-1, 0.0, 0x1.0c152382d7365p-1,
                                     -0x1.0c152382d7365p-1, 0.0,                   0x1.0c152382d7365p-1};

/* Agent.get_ref (agent.py:326-346) + Dynamics.update_ego_frame (dynamics/Dynamics.py:24-41) */
static void update_ego_frame(agent_t* a) {
  double dx = a->gx - a->px, dy = a->gy - a->py;
  a->dist_to_goal = sqrt(sq(dx) + sq(dy)); /* math.sqrt(goal_direction[0]**2 + goal_direction[1]**2) */
  if (a->dist_to_goal > 1e-8) {
    a->prll_x = dx / a->dist_to_goal;
    a->prll_y = dy / a->dist_to_goal;
  } else {
    a->prll_x = dx;
    a->prll_y = dy;
  }
  a->orth_x = -a->prll_y;
  a->orth_y = a->prll_x;
  a->heading_ego = wrap(a->heading - atan2(a->prll_y, a->prll_x));
}

static int is_done(const agent_t* a) { return (a->flags & CA_F_DONE_MASK) != 0; }

/* tangent_vecs_from_external_pt, GCA/envs/util.py:76-104; returns 0 when "None, None" */
static int tangent_vecs(double xp, double yp, double a, double b, double r, double v1[2], double v2[2]) {
  double sqd = sq(xp - a) + sq(yp - b) - sq(r);
  if (sqd < 0) return 0;
  double sqrt_term = sqrt(sq(xp - a) + sq(yp - b) - sq(r));
  double xnum1 = sq(r) * (xp - a);
  double xnum2 = r * (yp - b) * sqrt_term;
  double ynum1 = sq(r) * (yp - b);
  double ynum2 = r * (xp - a) * sqrt_term;
  double den = sq(xp - a) + sq(yp - b);
  double p1x = (xnum1 + xnum2) / den + a, p1y = (ynum1 - ynum2) / den + b;
  double p2x = (xnum1 - xnum2) / den + a, p2y = (ynum1 + ynum2) / den + b;
  v1[0] = p1x - xp; v1[1] = p1y - yp;
  v2[0] = p2x - xp; v2[1] = p2y - yp;
  return 1;
}

static inline double cross2(const double a[2], const double b[2]) { return a[0] * b[1] - a[1] * b[0]; }

/* compute_time_to_impact, GCA/envs/util.py:14-74 */
static double time_to_impact(const agent_t* h, const agent_t* o, double combined_radius) {
  double v_rel[2] = {h->vx - o->vx, h->vy - o->vy};
  double c1[2], c2[2];
  if (!tangent_vecs(h->px, h->py, o->px, o->py, combined_radius, c1, c2)) return 0.0;
  if (cross2(c1, v_rel) * cross2(c1, c2) >= 0 && cross2(c2, v_rel) * cross2(c2, c1) >= 0) {
    double v0 = v_rel[0], v1 = v_rel[1];
    if (fabs(v0) < 1e-5 && fabs(v1) < 1e-5) return INFINITY;
    double px = h->px, py = h->py, a = o->px, b = o->py, r = combined_radius;
    double x1, x2, y1, y2;
    if (fabs(v0) < 1e-5) {
      x1 = x2 = px;
      double A = 1, B = -2 * b, C = sq(b) + sq(px - a) - sq(r);
      y1 = (-B + sqrt(sq(B) - 4 * A * C)) / (2 * A);
      y2 = (-B - sqrt(sq(B) - 4 * A * C)) / (2 * A);
    } else {
      double m = v1 / v0;
      double A = 1 + sq(m);
      double B = -2 * a + 2 * m * (py - b - m * px);
      double C = sq(a) - sq(r) + sq(m * px - (py - b));
      x1 = (-B + sqrt(sq(B) - 4 * A * C)) / (2 * A);
      x2 = (-B - sqrt(sq(B) - 4 * A * C)) / (2 * A);
      y1 = m * (x1 - px) + py;
      y2 = m * (x2 - px) + py;
    }
    double d1 = np_norm2(x1 - px, y1 - py);
    double d2 = np_norm2(x2 - px, y2 - py);
    double d = (d2 < d1) ? d2 : d1; /* Python min(d1, d2) */
    double spd = np_norm2(v_rel[0], v_rel[1]);
    return d / spd;
  }
  return INFINITY;
}

typedef struct { int idx; double rd, porth, tti; } crit_t;

/* Python tuple "<" on (k0,k1,k2) followed by insertion index (stable sort) */
static int key_less3(double a0, double a1, double a2, int ai, double b0, double b1, double b2, int bi) {
  if (a0 != b0) return a0 < b0;
  if (a1 != b1) return a1 < b1;
  if (a2 != b2) return a2 < b2;
  return ai < bi;
}

static void stable_sort(crit_t* c, int cnt, int method_key /*0:(rd,porth) 1:(-rd,porth) 2:(-tti,-rd,porth)*/) {
  /* insertion sort; `pos` = position in the incoming list provides Python's stability */
  for (int i = 1; i < cnt; ++i) {
    crit_t x = c[i];
    int j = i - 1;
    for (; j >= 0; --j) {
      int less;
      const crit_t* y = &c[j];
      if (method_key == 0) less = key_less3(x.rd, x.porth, 0, 1, y->rd, y->porth, 0, 0);
      else if (method_key == 1) less = key_less3(-x.rd, x.porth, 0, 1, -y->rd, y->porth, 0, 0);
      else less = key_less3(-x.tti, -x.rd, x.porth, 1, -y->tti, -y->rd, y->porth, 0);
      /* x came after y: it moves before y only if strictly smaller on the key */
      if (!less) break;
      c[j + 1] = c[j];
    }
    c[j + 1] = x;
  }
}

/* OtherAgentsStatesSensor.sense + get_clipped_sorted_inds, GCA/envs/sensors/OtherAgentsStatesSensor.py:20-144,
 * then the dense row of MultiagentDictToMultiagentArrayWrapper.observation (GCA/envs/wrappers.py:130-139). */
static void sense_and_fill(const ca_oracle* o, const agent_t* ag, int n, int i, double* row, int32_t* sidx) {
  const int M = o->M;
  const agent_t* host = &ag[i];
  crit_t crit[CA_MAX_AGENTS];
  int cnt = 0;
  for (int j = 0; j < n; ++j) {
    if (j == i) continue;
    const agent_t* oth = &ag[j];
    double rx = oth->px - host->px, ry = oth->py - host->py;
    double p_orth = np_dot2(rx, ry, host->orth_x, host->orth_y);
    double dcen = sqrt(sq(rx) + sq(ry)); /* vec2_l2_norm, util.py:106-112 */
    double d2o = dcen - host->radius - oth->radius;
    double combined = host->radius + oth->radius;
    if (dcen > o->cfg.sensing_horizon) continue;
    double tti = 0.0;
    if (o->cfg.sort_method == CA_SORT_TIME_TO_IMPACT) tti = time_to_impact(host, oth, combined);
    crit[cnt].idx = j;
    crit[cnt].rd = rint(d2o * 100.0) / 100.0; /* round(np.float64, 2) == np.round: rint(x*100)/100 */
    crit[cnt].porth = p_orth;
    crit[cnt].tti = tti;
    ++cnt;
  }
  int first_key = (o->cfg.sort_method == CA_SORT_TIME_TO_IMPACT) ? 2 : 0;
  stable_sort(crit, cnt, first_key);
  if (cnt > M) cnt = M;
  int second_key = o->cfg.sort_method == CA_SORT_CLOSEST_LAST ? 1 : (o->cfg.sort_method == CA_SORT_TIME_TO_IMPACT ? 2 : 0);
  stable_sort(crit, cnt, second_key);

  row[0] = (host->policy == CA_POLICY_LEARNING_GA3C || host->policy == CA_POLICY_LEARNING) ? 1.0 : 0.0;
  row[1] = (double)cnt;
  row[2] = host->dist_to_goal;
  row[3] = host->heading_ego;
  row[4] = host->pref_speed;
  row[5] = host->radius;
  for (int k = 0; k < M; ++k) {
    double* s = row + CA_OBS_HOST_LEN + CA_OBS_OTHER_LEN * k;
    if (k < cnt) {
      const agent_t* oth = &ag[crit[k].idx];
      double rx = oth->px - host->px, ry = oth->py - host->py;
      s[0] = np_dot2(rx, ry, host->prll_x, host->prll_y);
      s[1] = np_dot2(rx, ry, host->orth_x, host->orth_y);
      s[2] = np_dot2(oth->vx, oth->vy, host->prll_x, host->prll_y);
      s[3] = np_dot2(oth->vx, oth->vy, host->orth_x, host->orth_y);
      s[4] = oth->radius;
      s[5] = host->radius + oth->radius;
      s[6] = np_norm2(rx, ry) - host->radius - oth->radius;
      if (sidx) sidx[k] = crit[k].idx;
    } else {
      for (int q = 0; q < CA_OBS_OTHER_LEN; ++q) s[q] = 0.0;
      if (sidx) sidx[k] = -1;
    }
  }
}

static void observe_world(const ca_oracle* o, int w, double* obs, int32_t* sorted_idx) {
  const int A = o->A, L = o->L, M = o->M;
  const agent_t* ag = o->agents + (size_t)w * A;
  int n = o->n[w];
  for (int i = 0; i < A; ++i) {
    double* row = obs + ((size_t)w * A + i) * L;
    int32_t* sidx = sorted_idx ? sorted_idx + ((size_t)w * A + i) * M : NULL;
    if (i < n) {
      sense_and_fill(o, ag, n, i, row, sidx);
    } else {
      memset(row, 0, sizeof(double) * L);
      if (sidx) for (int k = 0; k < M; ++k) sidx[k] = -1;
    }
  }
}

static void reset_world(ca_oracle* o, int w) {
  o->n[w] = o->n0[w];
  memcpy(o->agents + (size_t)w * o->A, o->init_agents + (size_t)w * o->A, sizeof(agent_t) * o->A);
}

/* CollisionAvoidanceEnv.step for one world, GCA/envs/collision_avoidance_env.py:131-194 */
static void step_world(ca_oracle* o, int w, const int32_t* actions, const double* cont_actions, double* obs,
                       double* reward, uint8_t* done, uint8_t* game_over, int32_t* sorted_idx) {
  const int A = o->A;
  const ca_config* c = &o->cfg;
  agent_t* ag = o->agents + (size_t)w * A;
  const int n = o->n[w];
  float act[CA_MAX_AGENTS][2]; /* all_actions is float32, collision_avoidance_env.py:238 */

  /* _take_action, :217-252 — every agent chooses first, then everybody moves */
  for (int i = 0; i < n; ++i) {
    agent_t* a = &ag[i];
    act[i][0] = 0.f; act[i][1] = 0.f;
    if (is_done(a)) continue; /* agent.is_done, :242 */
    switch (a->policy) {
      case CA_POLICY_LEARNING_GA3C: { /* LearningPolicyGA3C.py:13-27 */
        int k = actions[(size_t)w * A + i];
        act[i][0] = (float)(a->pref_speed * ACT_SPEED[k]);
        act[i][1] = (float)ACT_DHEAD[k];
      } break;
      case CA_POLICY_LEARNING: { /* LearningPolicy.py:13-33 */
        static const double NO_COMMAND[2] = {0.0, 0.5}; /* cont_actions == NULL: stand still, keep heading */
        const double* e = cont_actions ? cont_actions + ((size_t)w * A + i) * 2 : NO_COMMAND;
        double heading_change = c->max_heading_change * (2. * e[1] - 1.);
        double speed = a->pref_speed * e[0];
        act[i][0] = (float)speed;
        act[i][1] = (float)heading_change;
      } break;
      case CA_POLICY_NONCOOP: /* NonCooperativePolicy.py:9-22 */
        act[i][0] = (float)a->pref_speed;
        act[i][1] = (float)(-a->heading_ego);
        break;
      case CA_POLICY_STATIC: /* StaticPolicy.py:9-23: goal := pos */
        a->gx = a->px; a->gy = a->py;
        break;
    }
  }
  for (int i = 0; i < n; ++i) { /* Agent.take_action, agent.py:190-238 */
    agent_t* a = &ag[i];
    if (a->flags & CA_F_DONE_MASK) {
      if (a->flags & CA_F_AT_GOAL) a->flags |= CA_F_WAS_AT_GOAL;
      if (a->flags & CA_F_IN_COLLISION) a->flags |= CA_F_WAS_IN_COLLISION;
      a->vx = 0.0; a->vy = 0.0;
      continue;
    }
    /* UnicycleDynamics.step, dynamics/UnicycleDynamics.py:14-47 (float32 command promoted to double) */
    double speed = (double)act[i][0];
    double sel_heading = wrap((double)act[i][1] + a->heading);
    double ch = cos(sel_heading), sh = sin(sel_heading);
    double dx = speed * ch * c->dt;
    double dy = speed * sh * c->dt;
    a->px += dx; a->py += dy;
    a->vx = speed * ch; a->vy = speed * sh;
    a->heading = sel_heading;
    update_ego_frame(a);
    /* _check_if_at_goal, agent.py:148-151 */
    if (sq(a->px - a->gx) + sq(a->py - a->gy) <= sq(c->near_goal_threshold)) a->flags |= CA_F_AT_GOAL;
    else a->flags &= ~CA_F_AT_GOAL;
    a->t_rem -= c->dt; /* agent.py:232-236 */
    if (a->t_rem <= 0.0) a->flags |= CA_F_RAN_OUT_OF_TIME;
  }

  /* _check_for_collisions, :370-409 */
  int coll[CA_MAX_AGENTS];
  double nearest[CA_MAX_AGENTS];
  for (int i = 0; i < n; ++i) { coll[i] = 0; nearest[i] = INFINITY; }
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      double d = sqrt(sq(ag[i].px - ag[j].px) + sq(ag[i].py - ag[j].py)); /* l2norm, util.py:8-12 */
      double R = ag[i].radius + ag[j].radius;
      double gap = d - R;
      if (gap < nearest[i]) nearest[i] = gap; /* min(nearest[i], gap): only agent i */
      if (d <= R) { coll[i] = 1; coll[j] = 1; }
    }
  /* _compute_rewards, :319-368 */
  for (int i = 0; i < n; ++i) {
    agent_t* a = &ag[i];
    double r = c->reward_time_step;
    if (a->flags & CA_F_AT_GOAL) {
      if (!(a->flags & CA_F_WAS_AT_GOAL)) r = c->reward_at_goal;
    } else if (!(a->flags & CA_F_WAS_IN_COLLISION)) {
      if (coll[i]) {
        r = c->reward_collision_with_agent;
        a->flags |= CA_F_IN_COLLISION;
      } else if (nearest[i] <= c->getting_close_range) {
        r = -0.1 - nearest[i] / 2.;
      }
    }
    if (r < c->min_possible_reward) r = c->min_possible_reward; /* np.clip, :364 */
    if (r > c->max_possible_reward) r = c->max_possible_reward;
    reward[(size_t)w * A + i] = r;
  }
  for (int i = n; i < A; ++i) reward[(size_t)w * A + i] = 0.0;
  if (c->game_over_mode == CA_OVER_FIRST_AGENT_DONE) /* rewards = rewards[0], :365-366 */
    for (int i = 1; i < n; ++i) reward[(size_t)w * A + i] = 0.0;

  /* _get_obs, :441-461 */
  observe_world(o, w, obs, sorted_idx);

  /* _check_which_agents_done, :411-439 */
  int all_done = 1, all_learning_done = 1;
  for (int i = 0; i < A; ++i) {
    int d = i < n ? is_done(&ag[i]) : 1;
    done[(size_t)w * A + i] = (uint8_t)d;
    if (i < n) {
      if (!d) all_done = 0;
      int learning = ag[i].policy == CA_POLICY_LEARNING_GA3C || ag[i].policy == CA_POLICY_LEARNING;
      if (learning && !d) all_learning_done = 0;
    }
  }
  int over = c->game_over_mode == CA_OVER_ALL_DONE ? all_done
           : c->game_over_mode == CA_OVER_FIRST_AGENT_DONE ? is_done(&ag[0]) : all_learning_done;
  game_over[w] = (uint8_t)over;

  /* DummyVecEnv.step_wait (baselines@ea25b9e): done env is reset at once, new obs returned */
  if (over && c->auto_reset) {
    reset_world(o, w);
    observe_world(o, w, obs, sorted_idx);
  }
}

/* ------------------------------------------------------------------ public oracle API */

int ca_oracle_create(const ca_config* cfg, ca_oracle** out) {
  if (!cfg || !out || cfg->num_worlds < 1 || cfg->max_agents < 1 || cfg->max_agents > CA_MAX_AGENTS ||
      cfg->max_others_observed < 1)
    return CA_ERR_INVALID_ARG;
  ca_oracle* o = (ca_oracle*)calloc(1, sizeof(ca_oracle));
  if (!o) return CA_ERR_ALLOC;
  o->cfg = *cfg;
  o->W = cfg->num_worlds; o->A = cfg->max_agents; o->M = cfg->max_others_observed;
  o->L = CA_OBS_LEN(o->M);
  o->agents = (agent_t*)calloc((size_t)o->W * o->A, sizeof(agent_t));
  o->init_agents = (agent_t*)calloc((size_t)o->W * o->A, sizeof(agent_t));
  o->n = (int32_t*)calloc((size_t)o->W, sizeof(int32_t));
  o->n0 = (int32_t*)calloc((size_t)o->W, sizeof(int32_t));
  if (!o->agents || !o->init_agents || !o->n || !o->n0) return CA_ERR_ALLOC;
  *out = o;
  return CA_OK;
}

void ca_oracle_destroy(ca_oracle* o) {
  if (!o) return;
  free(o->agents); free(o->init_agents); free(o->n); free(o->n0); free(o);
}

/* Agent.__init__/reset, GCA/envs/agent.py:29-136 */
static int load_snapshot(ca_oracle* o, const double* init, const int32_t* num_agents) {
  for (int w = 0; w < o->W; ++w) {
    int n = num_agents[w];
    if (n < 1 || n > o->A) return CA_ERR_INVALID_ARG;
    o->n0[w] = n;
    for (int i = 0; i < o->A; ++i) {
      agent_t* a = &o->init_agents[(size_t)w * o->A + i];
      memset(a, 0, sizeof(*a));
      if (i >= n) continue;
      const double* r = init + ((size_t)w * o->A + i) * CA_INIT_STRIDE;
      a->px = r[CA_I_PX]; a->py = r[CA_I_PY]; a->gx = r[CA_I_GX]; a->gy = r[CA_I_GY];
      a->pref_speed = r[CA_I_PREF_SPEED]; a->radius = r[CA_I_RADIUS]; a->heading = r[CA_I_HEADING];
      a->policy = (int32_t)r[CA_I_POLICY];
      double t0 = r[CA_I_TIME_REMAINING];
      if (isnan(t0)) { /* agent.py:98-103 */
        t0 = o->cfg.max_time_ratio * ((np_norm2(a->px - a->gx, a->py - a->gy) - o->cfg.near_goal_threshold) / a->pref_speed);
        if (!(t0 > o->cfg.dt)) t0 = o->cfg.dt; /* max(t0, dt) */
      }
      a->t_rem = t0;
      update_ego_frame(a);
    }
  }
  return CA_OK;
}

int ca_oracle_set_world_state(ca_oracle* o, const double* init, const int32_t* num_agents) {
  int rc = load_snapshot(o, init, num_agents);
  if (rc != CA_OK) return rc;
  memcpy(o->agents, o->init_agents, sizeof(agent_t) * (size_t)o->W * o->A);
  memcpy(o->n, o->n0, sizeof(int32_t) * (size_t)o->W);
  o->initialised = 1;
  return CA_OK;
}

/* only the reset snapshot changes; live worlds pick it up at their next reset (≙ test_case_fn on env.reset()) */
int ca_oracle_set_reset_state(ca_oracle* o, const double* init, const int32_t* num_agents) {
  if (!o->initialised) return CA_ERR_NOT_INITIALISED;
  return load_snapshot(o, init, num_agents);
}

int ca_oracle_reset(ca_oracle* o, const uint8_t* world_mask, double* obs, int32_t* sorted_idx) {
  if (!o->initialised) return CA_ERR_NOT_INITIALISED;
  for (int w = 0; w < o->W; ++w) {
    if (!world_mask || world_mask[w]) reset_world(o, w);
    if (obs) observe_world(o, w, obs, sorted_idx);
  }
  return CA_OK;
}

typedef struct {
  ca_oracle* o; int w0, w1;
  const int32_t* actions; const double* cont; double* obs; double* reward; uint8_t* done; uint8_t* over; int32_t* sidx;
} job_t;

static void* job_main(void* p) {
  job_t* j = (job_t*)p;
  for (int w = j->w0; w < j->w1; ++w) step_world(j->o, w, j->actions, j->cont, j->obs, j->reward, j->done, j->over, j->sidx);
  return NULL;
}

int ca_oracle_step(ca_oracle* o, const int32_t* actions, const double* cont_actions, double* obs, double* reward,
                   uint8_t* done, uint8_t* game_over, int32_t* sorted_idx, int nthreads) {
  if (!o->initialised) return CA_ERR_NOT_INITIALISED;
  if (nthreads <= 1) {
    job_t j = {o, 0, o->W, actions, cont_actions, obs, reward, done, game_over, sorted_idx};
    job_main(&j);
    return CA_OK;
  }
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256]; job_t jobs[256];
  int per = (o->W + nthreads - 1) / nthreads;
  int started = 0;
  for (int t = 0; t < nthreads; ++t) {
    int w0 = t * per, w1 = w0 + per > o->W ? o->W : w0 + per;
    if (w0 >= w1) break;
    jobs[t] = (job_t){o, w0, w1, actions, cont_actions, obs, reward, done, game_over, sorted_idx};
    pthread_create(&th[t], NULL, job_main, &jobs[t]);
    ++started;
  }
  for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
  return CA_OK;
}

int ca_oracle_get_state(const ca_oracle* o, double* out) {
  for (size_t k = 0; k < (size_t)o->W * o->A; ++k) {
    const agent_t* a = &o->agents[k];
    double* r = out + k * CA_STATE_STRIDE;
    r[CA_S_PX] = a->px; r[CA_S_PY] = a->py; r[CA_S_HEADING] = a->heading; r[CA_S_VX] = a->vx; r[CA_S_VY] = a->vy;
    r[CA_S_TIME_REMAINING] = a->t_rem; r[CA_S_GX] = a->gx; r[CA_S_GY] = a->gy; r[CA_S_RADIUS] = a->radius;
    r[CA_S_PREF_SPEED] = a->pref_speed; r[CA_S_FLAGS] = (double)a->flags; r[CA_S_POLICY] = (double)a->policy;
  }
  return CA_OK;
}

/* ProcessAgent._accumulate_rewards inner recursion, GA3C/ProcessAgent.py:71-76 (float64 like the reference) */
int ca_oracle_nstep_returns(const double* reward, const double* bootstrap, double* out, int T, int N, double gamma) {
  for (int k = 0; k < N; ++k) {
    double R = bootstrap[k];
    for (int t = T - 1; t >= 0; --t) {
      R = gamma * R + reward[(size_t)t * N + k];
      out[(size_t)t * N + k] = R;
    }
  }
  return CA_OK;
}
