// ca_step.cu — host side of libcastep.so: the C-ABI of include/ca_step.h over the kernels in ca_kernels.cuh.
// No torch, no CPU fallback: every entry point needs a CUDA device.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "ca_kernels.cuh"
#include "ca_step_fast.cuh"
#include "ca_step_stream.cuh"
#include "ca_ga3c.cuh"
#include "ca_scenarios.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CA_CUDA(expr)                                                                                      \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      return fail(CA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Makes env->device current for the duration of a call and restores the caller's device afterwards.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    target = dev;
  }
  ~DeviceGuard() {
    if (ok && prev >= 0 && prev != target) cudaSetDevice(prev);
  }
  int target = -1;
};

}  // namespace

// used by the other translation units of the library (ca_predict.cu) to record a message for ca_last_error()
int ca_fail_external(int code, const char* msg) { return fail(code, "%s", msg); }

constexpr int kWarpsDefault = 4;

struct ca_env {
  ca_config cfg;
  double step_dt = 0.0;        // dt of the next steps (ca_set_dt); cfg.dt = Config.DT is what reset's time budget uses
  int W = 0, A = 0, M = 0, L = 0, wpw = 0;
  int grid = 0;
  size_t smem_bytes = 0;
  int tile_floats = 0;
  bool bulk_ok = true;
  bool force_generic = false;
  int kernel_choice = 1;       // 1 = one-shot specialised kernel, 0 = streaming persistent kernel (ca_step_stream.cuh)
  bool dynamic_sched = true;   // streaming kernel: chunks pulled from the ticket counter (CA_STREAM_STATIC=1: strided)
  unsigned* ticket = nullptr;  // streaming kernel: self-resetting work counter
  bool use_pdl = true;         // programmatic dependent launch for step kernels (CA_DISABLE_PDL=1 turns it off)
  int fast_warps = kWarpsDefault;  // warps per CTA of the one-shot kernel (CA_ONESHOT_WARPS = 1 | 2 | 4)
  int fast_grid = 0;
  // host-side knowledge "every world has all A agents, live state and snapshot alike" (host-provided agent counts, or a
  // generator configured with min_agents == max_agents == A): the step launch then takes the instantiation of the
  // per-other loops without bound checks; anything unknown (device-resident counts) or ragged takes the bounded one.
  // Launch-uniform on purpose: the two renderings of a 10-agent step do not fit the instruction cache together.
  bool all_present_live = false, all_present_snapshot = false;
  bool snapshot_prefetch = true;  // early L2 prefetch of the snapshot of ending worlds (CA_DISABLE_SNAPSHOT_PREFETCH=1: off)
  int pipe_min_blocks = 0;     // 0 = default instantiation
  int pipe_grid = 0;
  size_t smem_pipe = 0;
  int prefetch_chunks = 0;     // resident warps of the one-shot kernel (= distance of its L2 prefetch), 0 = off
  size_t smem_fast = 0;        // dynamic shared memory of the specialised step kernel (observation tile only)
  double* slab = nullptr;      // [2][n_chunks][kBlkDoubles]: live state blocks, then the reset snapshot (ca_kernels.cuh)
  long n_chunks = 0;
  uint8_t* consumed = nullptr; // [W]: world took its snapshot since the last ca_generate_scenarios
  uint64_t gen_calls = 0;
  ca::StateBlocks s{}, s0{};
  bool initialised = false;
  int64_t launches = 0;
  // staging for the *_host entry points (lazily allocated)
  cudaStream_t hstream = nullptr;
  int32_t* d_actions = nullptr;
  double* d_cont = nullptr;
  float* d_obs = nullptr;
  float* d_reward = nullptr;
  uint8_t* d_done = nullptr;
  uint8_t* d_over = nullptr;
  uint8_t* d_mask = nullptr;
  int32_t* d_sidx = nullptr;
  // staging for the host-side ca_set_world_state / ca_set_reset_state / ca_get_state (lazily allocated, owned by the handle)
  double* d_boundary = nullptr;   // max(W*A*CA_INIT_STRIDE, W*A*CA_STATE_STRIDE) doubles
  int32_t* d_boundary_nag = nullptr;
#ifdef CA_TRACE
  unsigned long long* trace = nullptr;  // experiment build: [n_chunks][8] stamps of the last step launch
#endif
};

namespace {

using ca::kBlock;
using ca::kWarps;

// init[W][A][CA_INIT_STRIDE] (AoS, float64) -> chunk blocks of the snapshot (+ live state).  Agent.__init__/reset,
// agent.py:29-136.
__global__ void unpack_init_kernel(const double* __restrict__ init, const int32_t* __restrict__ nag_in, ca::StateBlocks s,
                                   ca::StateBlocks s0, int W, int A, double max_time_ratio, double thr, double dt,
                                   double step_dt, bool snapshot_only, unsigned* ragged_flag) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)W * A) return;
  const int w = (int)(g / A), i = (int)(g - (long)w * A);
  int n = nag_in[w];
  n = n < 1 ? 1 : (n > A ? A : n);
  long chunk;
  int lane, wl;
  ca::slot_of(w, i, A, chunk, lane, wl);
  double* b0 = ca::blk_ptr(s0, chunk);
  double* b = ca::blk_ptr(s, chunk);
  if (i == 0) {
    ca::blk_nag(b0)[wl] = n;
    if (!snapshot_only) ca::blk_nag(b)[wl] = n;
    if (n != A) *ragged_flag = 1u;   // some world has fewer than A agents (read back by set_state_impl)
  }
  double v[CA_INIT_STRIDE];
#pragma unroll
  for (int c = 0; c < CA_INIT_STRIDE; ++c) v[c] = i < n ? init[g * CA_INIT_STRIDE + c] : 0.0;
  double t0 = v[CA_I_TIME_REMAINING];
  if (i < n && isnan(t0)) {  // agent.py:98-103 (np.linalg.norm accumulates with FMA, see oracle np_norm2)
    const double ex = v[CA_I_PX] - v[CA_I_GX], ey = v[CA_I_PY] - v[CA_I_GY];
    const double nrm = sqrt(__fma_rn(ey, ey, __dmul_rn(ex, ex)));
    t0 = max_time_ratio * ((nrm - thr) / v[CA_I_PREF_SPEED]);
    if (!(t0 > dt)) t0 = dt;
  }
  ca::Agent a;
  a.px = v[CA_I_PX]; a.py = v[CA_I_PY]; a.hd = v[CA_I_HEADING]; a.vx = 0.0; a.vy = 0.0; a.spd = 0.f;
  a.tr0 = i < n ? t0 : 0.0;
  a.cd = i < n ? ca::countdown_steps(t0, step_dt) : 0u;   // the steps after which the budget is spent (ca_kernels.cuh)
  a.gx = v[CA_I_GX]; a.gy = v[CA_I_GY]; a.rad = v[CA_I_RADIUS]; a.ps = v[CA_I_PREF_SPEED];
  a.flags = 0; a.policy = (int)v[CA_I_POLICY];
  ca::store_agent(b0, lane, a, true, true);
  if (!snapshot_only) ca::store_agent(b, lane, a, true, true);
}

__global__ void pack_state_kernel(ca::StateBlocks s, double* __restrict__ out, int W, int A, double step_dt) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)W * A) return;
  const int w = (int)(g / A), i = (int)(g - (long)w * A);
  long chunk;
  int lane, wl;
  ca::slot_of(w, i, A, chunk, lane, wl);
  ca::Agent a;
  double* const blk = ca::blk_ptr(s, chunk);
  ca::load_agent(blk, lane, a);
  // time_remaining_to_reach_goal: the subtractions the agent's steps stand for, repeated from the budget of its last reset
  const unsigned n0 = ca::blk_n0(blk)[lane];
  const double tr = ca::budget_after(blk[ca::O_TR0 + lane], step_dt, (a.cd == ca::kNever || n0 < a.cd) ? 0u : n0 - a.cd);
  double* r = out + g * CA_STATE_STRIDE;
  r[CA_S_PX] = a.px; r[CA_S_PY] = a.py; r[CA_S_HEADING] = a.hd; r[CA_S_VX] = a.vx; r[CA_S_VY] = a.vy;
  r[CA_S_TIME_REMAINING] = tr; r[CA_S_GX] = a.gx; r[CA_S_GY] = a.gy; r[CA_S_RADIUS] = a.rad;
  r[CA_S_PREF_SPEED] = a.ps; r[CA_S_FLAGS] = (double)a.flags; r[CA_S_POLICY] = (double)a.policy;
}

// ca_set_dt with a new dt: the countdowns were counted in steps of the old one.  Live agents: the budget they have left
// (old-dt subtractions from tr0) becomes the new reference point; snapshots: the same budget, recounted.
__global__ void rebase_countdown_kernel(ca::StateBlocks s, ca::StateBlocks s0, int W, int A, double dt_old, double dt_new) {
  const long g = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (long)W * A) return;
  const int w = (int)(g / A), i = (int)(g - (long)w * A);
  long chunk;
  int lane, wl;
  ca::slot_of(w, i, A, chunk, lane, wl);
  double* b = ca::blk_ptr(s, chunk);
  double* b0 = ca::blk_ptr(s0, chunk);
  {
    const uint32_t m = ca::blk_meta(b)[lane];
    const unsigned cd = m >> 16, n0 = ca::blk_n0(b)[lane];
    const double left = ca::budget_after(b[ca::O_TR0 + lane], dt_old, (cd == ca::kNever || n0 < cd) ? 0u : n0 - cd);
    // an agent whose budget is spent (cd == 0) stays there; an absent slot (n0 == 0, cd == 0) too
    const unsigned ncd = cd == 0 ? 0u : ca::countdown_steps(left, dt_new);
    b[ca::O_TR0 + lane] = left;
    ca::blk_n0(b)[lane] = ncd;
    ca::blk_meta(b)[lane] = (m & 0xffffu) | (ncd << 16);
  }
  {
    const uint32_t m = ca::blk_meta(b0)[lane];
    const unsigned ncd = (m >> 16) == 0 ? 0u : ca::countdown_steps(b0[ca::O_TR0 + lane], dt_new);
    ca::blk_n0(b0)[lane] = ncd;
    ca::blk_meta(b0)[lane] = (m & 0xffffu) | (ncd << 16);
  }
}

void carve(ca_env* e) {
  e->s.base = e->slab;
  e->s0.base = e->slab + (size_t)e->n_chunks * ca::kBlkDoubles;
}

ca::Params make_params(const ca_env* e) {
  ca::Params p;
  memset(&p, 0, sizeof(p));
  const ca_config& c = e->cfg;
  p.W = e->W; p.A = e->A; p.M = e->M; p.L = e->L; p.wpw = e->wpw;
  p.sort_method = c.sort_method; p.over_mode = c.game_over_mode; p.auto_reset = c.auto_reset;
  p.tile_floats = e->tile_floats;
  p.dt = e->step_dt;
  p.thr_sq = std::pow(c.near_goal_threshold, 2.0);  // near_goal_threshold**2 as Python evaluates it (agent.py:150)
  p.close_range = c.getting_close_range;
  p.r_goal = c.reward_at_goal; p.r_coll = c.reward_collision_with_agent; p.r_step = c.reward_time_step;
  p.r_min = c.min_possible_reward; p.r_max = c.max_possible_reward;
  p.r_goal_f = (float)p.r_goal; p.r_coll_f = (float)p.r_coll; p.r_step_f = (float)p.r_step;
  p.r_min_f = (float)p.r_min; p.r_max_f = (float)p.r_max;
  p.max_heading_change = c.max_heading_change;
  p.sensing_horizon = c.sensing_horizon;
  p.s = e->s; p.s0 = e->s0; p.consumed = e->consumed;
  p.ticket = e->ticket; p.dynamic_sched = e->dynamic_sched ? 1 : 0;
  p.all_present = (e->all_present_live && e->all_present_snapshot) ? 1 : 0;
  p.prefetch_snapshot = (c.auto_reset && c.game_over_mode == CA_OVER_ALL_LEARNING_DONE && e->snapshot_prefetch) ? 1 : 0;
#ifdef CA_TRACE
  p.trace = e->trace;
#endif
  return p;
}

bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15u) == 0; }

// The specialised step kernels (ca_step_fast.cuh) exist for these agent-slot counts.
#define CA_FAST_SIZES(X) X(2) X(3) X(4) X(5) X(6) X(8) X(10)

bool has_fast_kernel(const ca_env* e) {
  if (e->force_generic || e->cfg.sort_method == CA_SORT_TIME_TO_IMPACT) return false;
  switch (e->A) {
#define X(n) case n:
    CA_FAST_SIZES(X)
#undef X
    return true;
    default: return false;
  }
}

cudaError_t set_fast_smem_attr(int A, int bytes);

// Instantiations of the streaming kernel: (agent slots, min CTAs per SM used for register allocation).
#define CA_PIPE_VARIANTS(X) X(2, 8) X(3, 8) X(4, 6) X(4, 7) X(4, 8) X(5, 6) X(6, 6) X(8, 5) X(8, 6) X(10, 4) X(10, 5)

int default_pipe_min_blocks(int A) {
  switch (A) {
    case 2: case 3: return 8;
    case 4: return 7;
    case 5: case 6: return 6;
    case 8: return 6;
    default: return 5;
  }
}

const void* pipe_kernel_ptr(int A, int mb, bool dbg = false) {
#define X(a, b)          \
  if (A == a && mb == b) \
    return dbg ? (const void*)ca::ca_step_stream_kernel<a, b, true> : (const void*)ca::ca_step_stream_kernel<a, b, false>;
  CA_PIPE_VARIANTS(X)
#undef X
  return nullptr;
}

// One-shot specialised kernels: (agent slots, min CTAs/SM the register allocation targets).  The second number was
// picked from -Xptxas -v (largest occupancy without heavy spilling) and, for A = 4, measured on B200:
// 7 CTAs/SM lets 65 536 x 4 worlds (8192 warp chunks) finish in 2 rounds of 4144 resident warps instead of 3.
#define CA_ONESHOT_VARIANTS(X) X(2, 8) X(3, 8) X(4, 6) X(4, 7) X(4, 8) X(4, 9) X(5, 6) X(6, 6) X(8, 4) X(8, 5) X(8, 6) X(10, 3) X(10, 4) X(10, 5) X(10, 6)

int default_oneshot_min_blocks(int A) {
  switch (A) {
    case 2: case 3: return 8;
    case 4: return 7;          // measured: 6 -> 18.0, 7 -> 17.6, 8 -> 17.8, 9 -> 19.1 us at 4 x 65536
    case 5: case 6: return 6;
    case 8: return 6;          // measured: 5 -> 31.1 (before the register diet), 6 -> 26.3 us at 8 x 32768
    default: return 5;         // A = 10, measured: 4 -> 25.8, 5 -> 24.1 us at 10 x 16384
  }
}

const void* fast_kernel_ptr(int A, bool dbg) {
  int mb = default_oneshot_min_blocks(A);
  const char* env = getenv("CA_ONESHOT_MINBLOCKS");
  const int v = env ? atoi(env) : 0;
#define X(a, b) if (A == a && v == b) mb = v;
  CA_ONESHOT_VARIANTS(X)
#undef X
#define X(a, b) \
  if (A == a && mb == b) return dbg ? (const void*)ca::ca_step_kernel<a, b, true> : (const void*)ca::ca_step_kernel<a, b, false>;
  CA_ONESHOT_VARIANTS(X)
#undef X
  return nullptr;
}

// Launch with the programmatic-stream-serialization attribute (PDL): back-to-back steps overlap the launch latency
// and ramp-up of step t+1 with the tail of step t; the kernels call griddepcontrol.wait before touching global data.
cudaError_t set_fast_smem_attr(int A, int bytes) {
  for (int dbg = 0; dbg < 2; ++dbg) {
    const void* fn = fast_kernel_ptr(A, dbg != 0);
    if (!fn) continue;
    cudaError_t ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (ce != cudaSuccess) return ce;
  }
  return cudaSuccess;
}

int launch_pdl(const void* fn, int grid, size_t smem, cudaStream_t st, ca::Params& p, bool pdl, int threads = kBlock) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  void* args[] = {&p};
  CA_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  return CA_OK;
}

int launch_world_kernel(ca_env* e, bool step, ca::Params& p, cudaStream_t st) {
  p.use_bulk_store = (e->bulk_ok && aligned16(p.obs) && (e->tile_floats % 4) == 0) ? 1 : 0;
  int rc;
  // the instantiations with neighbour-index output / finite sensing horizon are only used when asked for
  // production instantiation: everything the GA3C training configs use; anything else takes the general one
  const bool dbg = p.sidx != nullptr || std::isfinite(e->cfg.sensing_horizon) || e->M != e->A - 1 || p.cont != nullptr ||
                   e->cfg.game_over_mode != CA_OVER_ALL_LEARNING_DONE || e->cfg.sort_method != CA_SORT_CLOSEST_FIRST;
  if (step && e->cfg.auto_reset && !e->all_present_snapshot) e->all_present_live = false;   // worlds may adopt ragged snapshots
  if (!step) e->all_present_live = p.mask ? (e->all_present_live && e->all_present_snapshot) : e->all_present_snapshot;
  if (step && has_fast_kernel(e) && e->kernel_choice == 0) {
    p.use_bulk_store = e->bulk_ok ? 1 : 0;  // the kernel aligns the tile to the destination's 16-byte phase itself
    rc = launch_pdl(pipe_kernel_ptr(e->A, e->pipe_min_blocks, dbg), e->pipe_grid, e->smem_pipe, st, p, e->use_pdl);
  } else if (step && has_fast_kernel(e)) {
    p.use_bulk_store = e->bulk_ok ? 1 : 0;  // the kernel aligns the tile to the destination's 16-byte phase itself
    p.prefetch_chunks = e->prefetch_chunks;
    rc = launch_pdl(fast_kernel_ptr(e->A, dbg), e->fast_grid, e->smem_fast, st, p, e->use_pdl, e->fast_warps * 32);
  } else if (step) {
    rc = launch_pdl((const void*)ca::ca_world_kernel<true>, e->grid, e->smem_bytes, st, p, e->use_pdl);
  } else {
    rc = launch_pdl((const void*)ca::ca_world_kernel<false>, e->grid, e->smem_bytes, st, p, false);
  }
  if (rc != CA_OK) return rc;
  CA_CUDA(cudaPeekAtLastError());
  e->launches += 1;
  return CA_OK;
}

int ensure_staging(ca_env* e) {
  if (e->hstream) return CA_OK;
  const size_t n = (size_t)e->W * e->A;
  cudaStream_t st = nullptr;
  CA_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  // all or nothing: hstream (the "staging is ready" marker) is only set once every buffer exists
  const bool ok = cudaMalloc(&e->d_actions, n * sizeof(int32_t)) == cudaSuccess &&
                  cudaMalloc(&e->d_cont, n * 2 * sizeof(double)) == cudaSuccess &&
                  cudaMalloc(&e->d_obs, n * e->L * sizeof(float)) == cudaSuccess &&
                  cudaMalloc(&e->d_reward, n * sizeof(float)) == cudaSuccess &&
                  cudaMalloc(&e->d_done, n) == cudaSuccess &&
                  cudaMalloc(&e->d_over, (size_t)e->W) == cudaSuccess &&
                  cudaMalloc(&e->d_mask, (size_t)e->W) == cudaSuccess &&
                  cudaMalloc(&e->d_sidx, n * e->M * sizeof(int32_t)) == cudaSuccess &&
                  cudaMemsetAsync(e->d_actions, 0, n * sizeof(int32_t), st) == cudaSuccess;
  if (!ok) {
    cudaFree(e->d_actions); cudaFree(e->d_cont); cudaFree(e->d_obs); cudaFree(e->d_reward);
    cudaFree(e->d_done); cudaFree(e->d_over); cudaFree(e->d_mask); cudaFree(e->d_sidx);
    e->d_actions = nullptr; e->d_cont = nullptr; e->d_obs = nullptr; e->d_reward = nullptr;
    e->d_done = nullptr; e->d_over = nullptr; e->d_mask = nullptr; e->d_sidx = nullptr;
    cudaStreamDestroy(st);
    cudaGetLastError();
    return fail(CA_ERR_ALLOC, "cudaMalloc of the host-path staging buffers failed");
  }
  e->hstream = st;
  return CA_OK;
}

int ensure_boundary(ca_env* e) {
  if (e->d_boundary) return CA_OK;
  const size_t n = (size_t)e->W * e->A;
  const size_t stride = CA_INIT_STRIDE > CA_STATE_STRIDE ? CA_INIT_STRIDE : CA_STATE_STRIDE;
  CA_CUDA(cudaMalloc(&e->d_boundary, n * stride * sizeof(double)));
  if (cudaMalloc(&e->d_boundary_nag, (size_t)e->W * sizeof(int32_t)) != cudaSuccess) {
    cudaFree(e->d_boundary);
    e->d_boundary = nullptr;
    cudaGetLastError();
    return fail(CA_ERR_ALLOC, "cudaMalloc of the boundary staging failed");
  }
  return CA_OK;
}

}  // namespace

extern "C" {

int ca_abi_version(void) { return CA_ABI_VERSION; }

const char* ca_last_error(void) { return g_last_error.c_str(); }

const char* ca_strerror(int code) {
  switch (code) {
    case CA_OK: return "ok";
    case CA_ERR_INVALID_ARG: return "invalid argument";
    case CA_ERR_CUDA: return "CUDA error";
    case CA_ERR_NOT_INITIALISED: return "world state not set (call ca_set_world_state first)";
    case CA_ERR_UNSUPPORTED: return "unsupported configuration";
    case CA_ERR_ALLOC: return "allocation failed";
    default: return "unknown error";
  }
}

int ca_default_config(ca_config* cfg, int32_t num_worlds, int32_t max_agents) {
  if (!cfg) return fail(CA_ERR_INVALID_ARG, "cfg is NULL");
  memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = CA_ABI_VERSION;
  cfg->num_worlds = num_worlds;
  cfg->max_agents = max_agents;
  cfg->max_others_observed = max_agents > 1 ? max_agents - 1 : 1;
  cfg->sort_method = CA_SORT_CLOSEST_FIRST;
  cfg->game_over_mode = CA_OVER_ALL_LEARNING_DONE;
  cfg->auto_reset = 0;
  cfg->device = 0;
  cfg->dt = 0.2;                            // GCA/envs/config.py:45
  cfg->near_goal_threshold = 0.2;           // :46
  cfg->getting_close_range = 0.2;           // :39
  cfg->reward_at_goal = 1.0;                // :30
  cfg->reward_collision_with_agent = -0.25; // :31
  cfg->reward_time_step = 0.0;              // :35
  cfg->min_possible_reward = -0.25;         // collision_avoidance_env.py:475-483
  cfg->max_possible_reward = 1.0;
  cfg->max_time_ratio = 2.0;                // :47
  cfg->max_heading_change = 3.141592653589793 / 3;  // collision_avoidance_env.py:76
  cfg->sensing_horizon = INFINITY;          // :76
  return CA_OK;
}

int ca_create(const ca_config* cfg, ca_env** out) {
  if (!cfg || !out) return fail(CA_ERR_INVALID_ARG, "cfg/out is NULL");
  *out = nullptr;
  if (cfg->abi_version != CA_ABI_VERSION)
    return fail(CA_ERR_INVALID_ARG, "abi_version %d != %d", cfg->abi_version, CA_ABI_VERSION);
  if (cfg->num_worlds < 1) return fail(CA_ERR_INVALID_ARG, "num_worlds must be >= 1");
  if (cfg->max_agents < 1 || cfg->max_agents > CA_MAX_AGENTS)
    return fail(CA_ERR_INVALID_ARG, "max_agents must be in 1..%d", CA_MAX_AGENTS);
  if (cfg->max_others_observed < 1 || cfg->max_others_observed > CA_MAX_AGENTS - 1)
    return fail(CA_ERR_INVALID_ARG, "max_others_observed must be in 1..%d", CA_MAX_AGENTS - 1);
  if (cfg->sort_method < 0 || cfg->sort_method > CA_SORT_TIME_TO_IMPACT)
    return fail(CA_ERR_INVALID_ARG, "bad sort_method %d", cfg->sort_method);
  if (cfg->game_over_mode < 0 || cfg->game_over_mode > CA_OVER_FIRST_AGENT_DONE)
    return fail(CA_ERR_INVALID_ARG, "bad game_over_mode %d", cfg->game_over_mode);
  if (!(cfg->dt > 0)) return fail(CA_ERR_INVALID_ARG, "dt must be > 0");
  if ((int64_t)cfg->num_worlds * cfg->max_agents * CA_OBS_LEN(cfg->max_others_observed) > (int64_t)1 << 40 ||
      (int64_t)cfg->num_worlds * cfg->max_agents >= (int64_t)1 << 31)
    return fail(CA_ERR_INVALID_ARG, "problem too large");
  int ndev = 0;
  CA_CUDA(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev)
    return fail(CA_ERR_CUDA, "device %d not available (%d CUDA devices visible)", cfg->device, ndev);
  DeviceGuard guard(cfg->device);
  if (!guard.ok) return fail(CA_ERR_CUDA, "cudaSetDevice(%d) failed", cfg->device);

  {  // constant tables of the step kernels (per device; idempotent)
    static const double sincos_table[18] = CA_SINCOS_TABLE;
    CA_CUDA(cudaMemcpyToSymbol(ca::kSC, sincos_table, sizeof(sincos_table)));
  }
  ca_env* e = new (std::nothrow) ca_env();
  if (!e) return fail(CA_ERR_ALLOC, "out of host memory");
  e->cfg = *cfg;
  e->step_dt = cfg->dt;
  e->W = cfg->num_worlds; e->A = cfg->max_agents; e->M = cfg->max_others_observed;
  e->L = CA_OBS_LEN(e->M);
  e->wpw = ca::worlds_per_chunk(e->A);
  e->n_chunks = ((long)e->W + e->wpw - 1) / e->wpw;
  const int worlds_per_cta = kWarps * e->wpw;
  e->grid = (e->W + worlds_per_cta - 1) / worlds_per_cta;
  e->tile_floats = worlds_per_cta * e->A * e->L;
  const size_t tile_bytes = (((size_t)e->tile_floats * 4 + 127) / 128) * 128;
  const int nkeys = cfg->sort_method == CA_SORT_TIME_TO_IMPACT ? 4 : 3;
  e->smem_bytes = tile_bytes + (size_t)nkeys * e->A * kBlock * sizeof(double);
  {  // the warps of the one-shot kernel are independent workers: fewer warps per CTA = finer-grained slot reuse
    const char* fw = getenv("CA_ONESHOT_WARPS");
    const int v = fw ? atoi(fw) : 0;
    if (v == 1 || v == 2 || v == 4) e->fast_warps = v;
  }
  e->fast_grid = (int)((e->n_chunks + e->fast_warps - 1) / e->fast_warps);
  e->smem_fast = (size_t)e->fast_warps * ca::warp_tile_region(e->tile_floats / kWarps);
  const char* nb = getenv("CA_DISABLE_BULK_STORE");
  e->bulk_ok = !(nb && nb[0] == '1');
  const char* fg = getenv("CA_FORCE_GENERIC");
  e->force_generic = fg && fg[0] == '1';
  int max_optin = 0;
  cudaError_t ce = cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
  if (ce != cudaSuccess) { delete e; return fail(CA_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(ce)); }
  if (e->smem_bytes > (size_t)max_optin) {
    delete e;
    return fail(CA_ERR_UNSUPPORTED, "configuration needs %zu B shared memory per CTA, device allows %d", e->smem_bytes, max_optin);
  }
  ce = cudaFuncSetAttribute(ca::ca_world_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes);
  if (ce == cudaSuccess)
    ce = cudaFuncSetAttribute(ca::ca_world_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_bytes);
  if (ce == cudaSuccess) ce = set_fast_smem_attr(e->A, (int)e->smem_fast);
  if (ce == cudaSuccess && has_fast_kernel(e)) {  // how many CTAs of the one-shot kernel are resident at once
    int per_sm = 0, sms = 0;
    ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fast_kernel_ptr(e->A, false), e->fast_warps * 32, e->smem_fast);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    const char* pf = getenv("CA_DISABLE_L2_PREFETCH");
    if (ce == cudaSuccess && !(pf && pf[0] == '1') && (long)per_sm * sms < (long)e->fast_grid)
      e->prefetch_chunks = per_sm * sms * e->fast_warps;
    if (getenv("CA_VERBOSE")) fprintf(stderr, "[ca] one-shot kernel: %d warps/CTA, %d CTAs/SM resident, grid %d\n", e->fast_warps, per_sm, e->fast_grid);
  }
  const char* kc = getenv("CA_STEP_KERNEL");  // "oneshot" (default) | "pipe" | "generic"
  if (kc && strcmp(kc, "oneshot") == 0) e->kernel_choice = 1;
  if (kc && (strcmp(kc, "stream") == 0 || strcmp(kc, "pipe") == 0)) e->kernel_choice = 0;
  const char* ss = getenv("CA_STREAM_STATIC");
  e->dynamic_sched = !(ss && ss[0] == '1');
  const char* np = getenv("CA_DISABLE_PDL");
  e->use_pdl = !(np && np[0] == '1');
  const char* sp = getenv("CA_DISABLE_SNAPSHOT_PREFETCH");
  e->snapshot_prefetch = !(sp && sp[0] == '1');
  if (kc && strcmp(kc, "generic") == 0) e->force_generic = true;
  if (ce == cudaSuccess && has_fast_kernel(e)) {
    e->pipe_min_blocks = default_pipe_min_blocks(e->A);
    const char* mb = getenv("CA_PIPE_MINBLOCKS");
    if (mb && pipe_kernel_ptr(e->A, atoi(mb))) e->pipe_min_blocks = atoi(mb);
    const void* fn = pipe_kernel_ptr(e->A, e->pipe_min_blocks);
    e->smem_pipe = (size_t)kWarps * ca::stream_warp_region(e->tile_floats / kWarps);
    ce = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_pipe);
    if (ce == cudaSuccess)
      ce = cudaFuncSetAttribute(pipe_kernel_ptr(e->A, e->pipe_min_blocks, true), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)e->smem_pipe);
    int per_sm = 0, sms = 0;
    if (ce == cudaSuccess) ce = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kBlock, e->smem_pipe);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
    if (ce == cudaSuccess) {
      const char* pg = getenv("CA_PIPE_CTAS_PER_SM");
      if (pg && atoi(pg) > 0 && atoi(pg) < per_sm) per_sm = atoi(pg);
      const int resident = sms * (per_sm > 0 ? per_sm : 1);
      e->pipe_grid = e->grid < resident ? e->grid : resident;
      const char* sg = getenv("CA_STREAM_GRID");  // tests: a tiny grid makes every warp loop over many chunks
      if (sg && atoi(sg) > 0 && atoi(sg) < e->pipe_grid) e->pipe_grid = atoi(sg);
    }
  }
  if (ce != cudaSuccess) {
    delete e;
    return fail(CA_ERR_CUDA, "kernel image not usable on this device (built for sm_100a): %s", cudaGetErrorString(ce));
  }
  const size_t slab_bytes = (size_t)e->n_chunks * ca::kBlkBytes * 2;
  if (cudaMalloc(&e->slab, slab_bytes) != cudaSuccess || cudaMalloc(&e->consumed, (size_t)e->W) != cudaSuccess ||
      cudaMalloc(&e->ticket, 128) != cudaSuccess) {
    cudaFree(e->slab); cudaFree(e->consumed); cudaFree(e->ticket);
    delete e;
    cudaGetLastError();
    return fail(CA_ERR_ALLOC, "cudaMalloc of %zu state bytes failed", slab_bytes);
  }
  cudaMemset(e->slab, 0, slab_bytes);
  carve(e);
  cudaMemset(e->consumed, 0, (size_t)e->W);
  cudaMemset(e->ticket, 0, 128);
#ifdef CA_TRACE
  cudaMalloc(&e->trace, (size_t)e->n_chunks * 64);
  cudaMemset(e->trace, 0, (size_t)e->n_chunks * 64);
#endif
  *out = e;
  return CA_OK;
}

#ifdef CA_TRACE
// experiment build only (scripts/step_timeline.py): the stamps of the handle's last step launch -> host
int ca_trace_read(ca_env* e, unsigned long long* out) {
  DeviceGuard guard(e->cfg.device);
  CA_CUDA(cudaDeviceSynchronize());
  CA_CUDA(cudaMemcpy(out, e->trace, (size_t)e->n_chunks * 64, cudaMemcpyDeviceToHost));
  return CA_OK;
}
#endif

int ca_destroy(ca_env* e) {
  if (!e) return CA_OK;
  DeviceGuard guard(e->cfg.device);
  cudaDeviceSynchronize();
  cudaFree(e->slab); cudaFree(e->consumed); cudaFree(e->ticket);
  cudaFree(e->d_actions); cudaFree(e->d_cont); cudaFree(e->d_obs); cudaFree(e->d_reward);
  cudaFree(e->d_done); cudaFree(e->d_over); cudaFree(e->d_mask); cudaFree(e->d_sidx);
  cudaFree(e->d_boundary); cudaFree(e->d_boundary_nag);
#ifdef CA_TRACE
  cudaFree(e->trace);
#endif
  if (e->hstream) cudaStreamDestroy(e->hstream);
  delete e;
  return CA_OK;
}

static int set_state_impl(ca_env* e, const double* init, const int32_t* num_agents, int on_device, void* stream,
                          bool snapshot_only) {
  if (!e || !init || !num_agents) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  DeviceGuard guard(e->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)e->W * e->A;
  const double* d_init = init;
  const int32_t* d_nag = num_agents;
  bool all_present = !on_device;   // device-resident counts: answered by the unpack kernel below
  if (!on_device) {
    for (int w = 0; w < e->W; ++w) {
      if (num_agents[w] < 1 || num_agents[w] > e->A)
        return fail(CA_ERR_INVALID_ARG, "num_agents[%d] = %d outside 1..%d", w, num_agents[w], e->A);
      all_present = all_present && num_agents[w] == e->A;
    }
    const int rc = ensure_boundary(e);   // handle-owned staging: nothing to free on the error paths below
    if (rc != CA_OK) return rc;
    CA_CUDA(cudaMemcpyAsync(e->d_boundary, init, n * CA_INIT_STRIDE * sizeof(double), cudaMemcpyHostToDevice, st));
    CA_CUDA(cudaMemcpyAsync(e->d_boundary_nag, num_agents, (size_t)e->W * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    d_init = e->d_boundary;
    d_nag = e->d_boundary_nag;
  }
  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads);
  unsigned* const ragged_flag = e->ticket + 16;   // a spare word of the handle's 128-byte counter block
  if (on_device) CA_CUDA(cudaMemsetAsync(ragged_flag, 0, sizeof(unsigned), st));
  unpack_init_kernel<<<blocks, threads, 0, st>>>(d_init, d_nag, e->s, e->s0, e->W, e->A, e->cfg.max_time_ratio,
                                                 e->cfg.near_goal_threshold, e->cfg.dt, e->step_dt, snapshot_only, ragged_flag);
  CA_CUDA(cudaPeekAtLastError());
  e->launches += 1;
  if (on_device) {   // device-resident counts: the kernel reports whether every world has all A agents
    unsigned ragged = 1u;
    CA_CUDA(cudaMemcpyAsync(&ragged, ragged_flag, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
    all_present = ragged == 0u;
  } else {
    CA_CUDA(cudaStreamSynchronize(st));   // the staging may be reused by the next call
  }
  e->all_present_snapshot = all_present;
  if (!snapshot_only) { e->initialised = true; e->all_present_live = all_present; }
  return CA_OK;
}

int ca_set_world_state(ca_env* e, const double* init, const int32_t* num_agents, int on_device, void* stream) {
  return set_state_impl(e, init, num_agents, on_device, stream, false);
}

int ca_set_reset_state(ca_env* e, const double* init, const int32_t* num_agents, int on_device, void* stream) {
  if (e && !e->initialised) return fail(CA_ERR_NOT_INITIALISED, "ca_set_reset_state before ca_set_world_state");
  return set_state_impl(e, init, num_agents, on_device, stream, true);
}

int ca_reset(ca_env* e, const uint8_t* world_mask, float* obs, int32_t* sorted_idx, void* stream) {
  if (!e || !obs) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  if (!e->initialised) return fail(CA_ERR_NOT_INITIALISED, "ca_reset before ca_set_world_state");
  DeviceGuard guard(e->cfg.device);
  ca::Params p = make_params(e);
  p.mask = world_mask; p.obs = obs; p.sidx = sorted_idx;
  return launch_world_kernel(e, false, p, (cudaStream_t)stream);
}

int ca_step(ca_env* e, const int32_t* actions, const double* cont_actions, float* obs, float* reward, uint8_t* done,
            uint8_t* game_over, int32_t* sorted_idx, void* stream) {
  if (!e || !actions || !obs || !reward || !done || !game_over) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  if (!e->initialised) return fail(CA_ERR_NOT_INITIALISED, "ca_step before ca_set_world_state");
  DeviceGuard guard(e->cfg.device);
  ca::Params p = make_params(e);
  p.actions = actions; p.cont = cont_actions; p.obs = obs; p.reward = reward; p.done = done; p.over = game_over;
  p.sidx = sorted_idx;
  return launch_world_kernel(e, true, p, (cudaStream_t)stream);
}

// VecEnv.step_async (openai/baselines vec_env.py, the interface MultiagentDummyVecEnv implements,
// GCA/envs/wrappers.py:104-109): enqueue H2D actions -> step kernel -> D2H results on the handle's own stream and
// return.  The host buffers must stay valid (and, to overlap, be page-locked) until ca_step_host_wait.
int ca_step_host_async(ca_env* e, const int32_t* actions, const double* cont_actions, float* obs, float* reward,
                       uint8_t* done, uint8_t* game_over, int32_t* sorted_idx) {
  if (!e || !actions || !obs || !reward || !done || !game_over) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  if (!e->initialised) return fail(CA_ERR_NOT_INITIALISED, "ca_step_host before ca_set_world_state");
  DeviceGuard guard(e->cfg.device);
  int rc = ensure_staging(e);
  if (rc != CA_OK) return rc;
  const size_t n = (size_t)e->W * e->A;
  cudaStream_t st = e->hstream;
  CA_CUDA(cudaMemcpyAsync(e->d_actions, actions, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (cont_actions) CA_CUDA(cudaMemcpyAsync(e->d_cont, cont_actions, n * 2 * sizeof(double), cudaMemcpyHostToDevice, st));
  ca::Params p = make_params(e);
  p.actions = e->d_actions; p.cont = cont_actions ? e->d_cont : nullptr; p.obs = e->d_obs; p.reward = e->d_reward;
  p.done = e->d_done; p.over = e->d_over; p.sidx = sorted_idx ? e->d_sidx : nullptr;
  rc = launch_world_kernel(e, true, p, st);
  if (rc != CA_OK) return rc;
  CA_CUDA(cudaMemcpyAsync(obs, e->d_obs, n * e->L * sizeof(float), cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(reward, e->d_reward, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(done, e->d_done, n, cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaMemcpyAsync(game_over, e->d_over, (size_t)e->W, cudaMemcpyDeviceToHost, st));
  if (sorted_idx) CA_CUDA(cudaMemcpyAsync(sorted_idx, e->d_sidx, n * e->M * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  return CA_OK;
}

// VecEnv.step_wait: block until the results of the last ca_step_host_async are in the caller's buffers.
int ca_step_host_wait(ca_env* e) {
  if (!e) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  if (!e->hstream) return CA_OK;  // nothing was ever enqueued
  DeviceGuard guard(e->cfg.device);
  CA_CUDA(cudaStreamSynchronize(e->hstream));
  return CA_OK;
}

// VecEnv.step = step_async + step_wait
int ca_step_host(ca_env* e, const int32_t* actions, const double* cont_actions, float* obs, float* reward,
                 uint8_t* done, uint8_t* game_over, int32_t* sorted_idx) {
  const int rc = ca_step_host_async(e, actions, cont_actions, obs, reward, done, game_over, sorted_idx);
  if (rc != CA_OK) return rc;
  return ca_step_host_wait(e);
}

int ca_reset_host(ca_env* e, const uint8_t* world_mask, float* obs, int32_t* sorted_idx) {
  if (!e || !obs) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  if (!e->initialised) return fail(CA_ERR_NOT_INITIALISED, "ca_reset_host before ca_set_world_state");
  DeviceGuard guard(e->cfg.device);
  int rc = ensure_staging(e);
  if (rc != CA_OK) return rc;
  const size_t n = (size_t)e->W * e->A;
  cudaStream_t st = e->hstream;
  if (world_mask) CA_CUDA(cudaMemcpyAsync(e->d_mask, world_mask, (size_t)e->W, cudaMemcpyHostToDevice, st));
  ca::Params p = make_params(e);
  p.mask = world_mask ? e->d_mask : nullptr; p.obs = e->d_obs; p.sidx = sorted_idx ? e->d_sidx : nullptr;
  rc = launch_world_kernel(e, false, p, st);
  if (rc != CA_OK) return rc;
  CA_CUDA(cudaMemcpyAsync(obs, e->d_obs, n * e->L * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (sorted_idx) CA_CUDA(cudaMemcpyAsync(sorted_idx, e->d_sidx, n * e->M * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CA_CUDA(cudaStreamSynchronize(st));
  return CA_OK;
}

int ca_get_state(ca_env* e, double* out, int on_device, void* stream) {
  if (!e || !out) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  if (!e->initialised) return fail(CA_ERR_NOT_INITIALISED, "ca_get_state before ca_set_world_state");
  DeviceGuard guard(e->cfg.device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)e->W * e->A;
  double* d_out = out;
  if (!on_device) {
    const int rc = ensure_boundary(e);
    if (rc != CA_OK) return rc;
    d_out = e->d_boundary;
  }
  const int threads = 256;
  pack_state_kernel<<<(int)((n + threads - 1) / threads), threads, 0, st>>>(e->s, d_out, e->W, e->A, e->step_dt);
  CA_CUDA(cudaPeekAtLastError());
  e->launches += 1;
  if (!on_device) {
    CA_CUDA(cudaMemcpyAsync(out, d_out, n * CA_STATE_STRIDE * sizeof(double), cudaMemcpyDeviceToHost, st));
    CA_CUDA(cudaStreamSynchronize(st));
  }
  return CA_OK;
}

int ca_set_dt(ca_env* e, double dt) {
  if (!e) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  if (!(dt > 0) || !std::isfinite(dt)) return fail(CA_ERR_INVALID_ARG, "dt must be > 0");
  if (dt != e->step_dt && e->initialised) {
    // the time budgets are kept as step countdowns (ca_kernels.cuh): recount them for the new step length.  No stream
    // argument in this entry point: the device is drained on both sides of the recount.
    DeviceGuard guard(e->cfg.device);
    CA_CUDA(cudaDeviceSynchronize());
    const long n = (long)e->W * e->A;
    rebase_countdown_kernel<<<(int)((n + 255) / 256), 256>>>(e->s, e->s0, e->W, e->A, e->step_dt, dt);
    CA_CUDA(cudaPeekAtLastError());
    CA_CUDA(cudaDeviceSynchronize());
    e->launches += 1;
  }
  e->step_dt = dt;   // make_params() copies it into the next launch's parameters; Config.DT (reset time budget) stays
  return CA_OK;
}

int ca_launch_count(const ca_env* e, int64_t* out) {
  if (!e || !out) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  *out = e->launches;
  return CA_OK;
}

int ca_host_alloc(void** out, uint64_t bytes) {
  if (!out) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  CA_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
  return CA_OK;
}

int ca_host_free(void* ptr) {
  if (ptr) CA_CUDA(cudaFreeHost(ptr));
  return CA_OK;
}

int ca_nstep_returns(const float* reward, const float* bootstrap, float* out, int32_t T, int32_t N, float gamma,
                     int device, void* stream) {
  if (!reward || !bootstrap || !out || T < 0 || N < 1) return fail(CA_ERR_INVALID_ARG, "bad argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(CA_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  const int threads = 256;
  ca::nstep_returns_kernel<<<(N + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(reward, bootstrap, out, T,
                                                                                              N, gamma);
  CA_CUDA(cudaPeekAtLastError());
  return CA_OK;
}

int ca_default_scenario_config(ca_scenario_config* c, int32_t max_agents) {
  if (!c || max_agents < 1 || max_agents > CA_MAX_AGENTS) return fail(CA_ERR_INVALID_ARG, "bad argument");
  memset(c, 0, sizeof(*c));
  c->min_agents = max_agents >= 2 ? 2 : 1;  // np.random.randint(2, MAX_NUM_AGENTS_IN_ENVIRONMENT + 1), test_cases.py:97
  c->max_agents = max_agents;
  c->side_split_agents = 5;                 // config.py:57-60
  c->ensure_learner = 1;                    // config.py:51
  c->side_small_lo = 4; c->side_small_hi = 5; c->side_large_lo = 6; c->side_large_hi = 8;
  c->p_swap = 0.15; c->p_circle = 0.15;     // gen_rand_testcases.py:123-133
  c->speed_lo = 0.5; c->speed_hi = 2.0; c->radius_lo = 0.2; c->radius_hi = 0.8;  // config.py:55-56
  c->p_noncoop = 0.05; c->p_learning = 0.9; // config.py:52-54
  return CA_OK;
}

int ca_generate_scenarios(ca_env* e, const ca_scenario_config* c, uint64_t seed, int only_consumed, void* stream) {
  if (!e || !c) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  if (c->min_agents < 1 || c->max_agents > e->A || c->min_agents > c->max_agents)
    return fail(CA_ERR_INVALID_ARG, "agent count range %d..%d outside 1..%d", c->min_agents, c->max_agents, e->A);
  if (!(c->speed_lo > 0) || c->speed_hi < c->speed_lo || !(c->radius_lo > 0) || c->radius_hi < c->radius_lo ||
      !(c->side_small_lo > 0) || !(c->side_large_lo > 0) || c->p_swap < 0 || c->p_circle < 0 || c->p_swap + c->p_circle > 1 ||
      c->p_noncoop < 0 || c->p_learning < 0 || c->p_noncoop + c->p_learning > 1.0 + 1e-12)
    return fail(CA_ERR_INVALID_ARG, "bad scenario distribution parameters");
  DeviceGuard guard(e->cfg.device);
  ca::ScenarioParams p;
  memset(&p, 0, sizeof(p));
  p.c = *c; p.s0 = e->s0; p.consumed = e->consumed; p.W = e->W; p.A = e->A;
  p.only_consumed = only_consumed; p.dt = e->cfg.dt; p.thr = e->cfg.near_goal_threshold;
  p.step_dt = e->step_dt;
  p.max_time_ratio = e->cfg.max_time_ratio; p.seed = seed; p.offset = e->gen_calls * 4096ull;
  e->gen_calls += 1;
  ca::generate_scenarios_kernel<<<(e->W + ca::kGenWarps - 1) / ca::kGenWarps, ca::kGenWarps * 32, 0, (cudaStream_t)stream>>>(p);
  CA_CUDA(cudaPeekAtLastError());
  e->launches += 1;
  {
    const bool fixed = c->min_agents == e->A && c->max_agents == e->A;
    e->all_present_snapshot = only_consumed ? (e->all_present_snapshot && fixed) : fixed;
  }
  if (!e->initialised && !only_consumed) {
    // first use without ca_set_world_state: the generated snapshot defines the worlds; the caller must ca_reset (all
    // worlds) before stepping.  Live worlds of an initialised handle are never touched by the generator.
    // (live blocks are all zero from ca_create: agent count 0 everywhere until the reset adopts the snapshot)
    e->initialised = true;
  }
  return CA_OK;
}

int ca_ga3c_record(const ca_ga3c_buffers* b, int64_t t, int32_t ring_slots, int32_t num_slots, int32_t agents_per_world,
                   int32_t obs_len, int32_t time_max, float gamma, const int32_t* actions, const float* values,
                   const float* reward, const uint8_t* done, const uint8_t* game_over, int device, void* stream) {
  if (!b || !actions || !values || !reward || !done || !game_over) return fail(CA_ERR_INVALID_ARG, "NULL argument");
  if (!b->obs_ring || !b->act_ring || !b->rew_ring || !b->length || !b->tcount || !b->done_trained || !b->out_x ||
      !b->out_r || !b->out_a || !b->out_count || !b->out_src || !b->gathered)
    return fail(CA_ERR_INVALID_ARG, "NULL buffer in ca_ga3c_buffers");
  if (t < 0 || num_slots < 1 || agents_per_world < 1 || num_slots % agents_per_world != 0 || obs_len < 2 || time_max < 1 ||
      ring_slots < time_max + 2 || b->capacity < 1)
    return fail(CA_ERR_INVALID_ARG, "bad sizes (need ring_slots >= time_max + 2, num_slots %% agents_per_world == 0)");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(CA_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  ca::Ga3cParams p;
  p.b = *b; p.t = t; p.R = ring_slots; p.N = num_slots; p.A = agents_per_world; p.L = obs_len; p.time_max = time_max;
  p.gamma = gamma; p.actions = actions; p.values = values; p.reward = reward; p.done = done; p.over = game_over;
  const int threads = 128;
  ca::ga3c_record_kernel<<<(num_slots + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(p);
  CA_CUDA(cudaPeekAtLastError());
  // copy the observation rows of the experiences emitted at this step (ring slots are recycled by the next env step)
  ca::GatherParams gp;
  gp.b = *b; gp.L = obs_len; gp.gathered = b->gathered;
  const int rows_per_block = (256 / 32) * ca::kGatherRows;
  long blocks = ((long)num_slots * 2 + rows_per_block - 1) / rows_per_block;   // ~2 rows per slot in a normal step
  if (blocks > 4 * 148) blocks = 4 * 148;
  if (blocks < 1) blocks = 1;
  ca::ga3c_gather_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(gp);
  CA_CUDA(cudaPeekAtLastError());
  return CA_OK;
}

int ca_ga3c_episode_stats(const float* obs_now, const float* reward, const uint8_t* game_over, float* ep_reward,
                          int32_t* ep_steps, double* stats, int32_t num_worlds, int32_t agents_per_world,
                          int32_t obs_len, int device, void* stream) {
  if (!obs_now || !reward || !game_over || !ep_reward || !ep_steps || !stats || num_worlds < 1 || agents_per_world < 1)
    return fail(CA_ERR_INVALID_ARG, "bad argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(CA_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  const int threads = 128;
  ca::ga3c_episode_stats_kernel<<<(num_worlds + threads - 1) / threads, threads, 0, (cudaStream_t)stream>>>(
      obs_now, reward, game_over, ep_reward, ep_steps, stats, num_worlds, agents_per_world, obs_len);
  CA_CUDA(cudaPeekAtLastError());
  return CA_OK;
}

int ca_lstm_cell_forward(const float* z, const float* c_prev, const float* h_prev, const float* seq_len, int32_t seq_stride,
                         int32_t t, float* gates, float* c, float* h, int32_t batch, int device, void* stream) {
  if (!z || !c_prev || !h_prev || !seq_len || !gates || !c || !h || batch < 1 || t < 0 || seq_stride < 1)
    return fail(CA_ERR_INVALID_ARG, "bad argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(CA_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  const long n = (long)batch * 64;
  ca::lstm_cell_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z, c_prev, h_prev, seq_len, seq_stride,
                                                                                          t, gates, c, h, batch);
  CA_CUDA(cudaPeekAtLastError());
  return CA_OK;
}

int ca_lstm_cell_backward(const float* gates, const float* c_prev, const float* c_new, const float* seq_len,
                          int32_t seq_stride, int32_t t, const float* dc, const float* dh, float* dz, float* dc_prev,
                          float* dh_pass, int32_t batch, int device, void* stream) {
  if (!gates || !c_prev || !c_new || !seq_len || !dz || !dc_prev || !dh_pass || batch < 1 || t < 0 || seq_stride < 1)
    return fail(CA_ERR_INVALID_ARG, "bad argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(CA_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  const long n = (long)batch * 64;
  ca::lstm_cell_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gates, c_prev, c_new, seq_len,
                                                                                          seq_stride, t, dc, dh, dz, dc_prev,
                                                                                          dh_pass, batch);
  CA_CUDA(cudaPeekAtLastError());
  return CA_OK;
}

int ca_lstm_step(const float* obs, int32_t obs_stride, const float* zh, const float* Kx, const float* bias,
                 const float* avg7, const float* std7, float* c, float* h, int32_t batch, int32_t t, int device,
                 void* stream) {
  if (!obs || !Kx || !bias || !avg7 || !std7 || !c || !h || batch < 1 || t < 0 || obs_stride < 6 + 7 * (t + 1))
    return fail(CA_ERR_INVALID_ARG, "bad argument");
  DeviceGuard guard(device);
  if (!guard.ok) return fail(CA_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  ca::lstm_step_kernel<<<(batch + 3) / 4, 256, 0, (cudaStream_t)stream>>>(obs, obs_stride, zh, Kx, bias, avg7, std7, c, h,
                                                                          batch, t);
  CA_CUDA(cudaPeekAtLastError());
  return CA_OK;
}

}  // extern "C"
