// ca_ga3c.cuh — device-side experience bookkeeping of the GA3C actor (vectorised ProcessAgent.run_episode).
//
// Reference: GA3C/ProcessAgent.py:105-211 (run_episode) and :54-79 (_accumulate_rewards).  There, every
// learning agent of one env keeps a Python list of Experience objects that grows by one per env step and is
// flushed to the trainer when the agent is done or after TIME_MAX steps (SURVEY.md §8 N10 lists the quirks:
// the last experience is held back as the bootstrap unless the agent is done, a done agent keeps re-flushing
// 2-element lists every step until the world ends, a terminal flush of exactly TIME_MAX+1 experiences emits the
// terminal one separately with its raw reward, rewards are overwritten in place by the discounted return).
//
// Here the "list" of agent slot g is the window of the last `length[g]` time steps of three rings indexed by
// (t mod R): the observation ring the env kernel writes into, an action ring and a (mutable) reward ring.
// One thread per agent slot applies the list logic; emitted training rows (x = pre-step observation without
// the is_learning column, discounted return, action index) are appended to compact output arrays, the slot
// range being claimed with one atomicAdd per warp and the 26/68-float rows copied by the whole warp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ca_step.h"

namespace ca {

struct Ga3cParams {
  ca_ga3c_buffers b;
  long long t;     // global env-step counter; ring slot = t mod R
  int R, N, A, L, time_max;
  float gamma;
  const int32_t* actions;  // [N] action taken at step t
  const float* values;     // [N] V(s_t) predicted at step t
  const float* reward;     // [N] reward of step t
  const uint8_t* done;     // [N] which_agents_done after step t
  const uint8_t* over;     // [N / A] game_over after step t
};

// ring slot of list element k of a list whose first element was recorded first_t_off steps ago (0 <= k <= first_t_off < R)
__device__ __forceinline__ int ring_slot(int slot_now, int first_t_off, int k, int R) {
  const int s = slot_now - first_t_off + k;
  return s < 0 ? s + R : s;
}

__global__ void __launch_bounds__(128) ga3c_record_kernel(const __grid_constant__ Ga3cParams p) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int R = p.R, L = p.L, XL = p.L - 1;
  const size_t N = (size_t)p.N;
  const int slot_now = (int)(p.t % R);
  const bool in_range = g < p.N;
  // is_learning is column 0 of the observation the prediction was made from (ProcessAgent.py:128-133)
  const bool learning = in_range && p.b.obs_ring[((size_t)slot_now * N + g) * L] != 0.f;

  int n_emit = 0;       // rows this agent emits: list elements [0, n_main) plus optionally the held-back last one
  int n_main = 0, first_t_off = 0;
  bool emit_last = false;
  if (learning) {
    const bool dn = p.done[g] != 0;
    p.b.act_ring[(size_t)slot_now * N + g] = p.actions[g];
    p.b.rew_ring[(size_t)slot_now * N + g] = p.reward[g];
    int len = p.b.length[g] + 1;
    int tc = p.b.tcount[g];
    bool trained = p.b.done_trained[g] != 0;
    if (dn || (tc == p.time_max && !trained)) {  // ProcessAgent.py:186 (operator precedence as written there)
      float Rv = dn ? 0.f : p.values[g];
      if (dn) trained = true;
      first_t_off = len - 1;  // list element k lives at time t - first_t_off + k
      if (len == 1) {
        n_main = 1;           // returned unchanged, raw reward (:62-63)
      } else {
        emit_last = dn && len == p.time_max + 1;             // leftover_term_exp (:65-66)
        n_main = (dn && len != p.time_max + 1) ? len : len - 1;
        for (int k = n_main - 1; k >= 0; --k) {              // :71-76, rewards overwritten in place
          float* r = &p.b.rew_ring[(size_t)ring_slot(slot_now, first_t_off, k, R) * N + g];
          Rv = p.gamma * Rv + *r;
          *r = Rv;
        }
      }
      n_emit = n_main + (emit_last ? 1 : 0);
      tc = 0;
      len = 1;  // experiences[i] = [experiences[i][-1]] (:208)
    }
    tc += 1;
    if (p.over[g / p.A]) { len = 0; tc = 0; trained = false; }  // the next run_episode() starts from scratch
    p.b.length[g] = len;
    p.b.tcount[g] = tc;
    p.b.done_trained[g] = trained ? 1 : 0;
  }

  // claim output rows: exclusive prefix sum inside the warp, one atomicAdd per warp
  int incl = n_emit;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += v;
  }
  const int warp_total = __shfl_sync(0xffffffffu, incl, 31);
  if (warp_total == 0) return;
  int warp_base = 0;
  if (lane == 31) warp_base = atomicAdd(p.b.out_count, warp_total);
  warp_base = __shfl_sync(0xffffffffu, warp_base, 31);
  __syncwarp();  // the discounted returns written above are read back below by other lanes of the warp

  // Warp-cooperative row copy.  The warp's rows are numbered f = 0 .. warp_total-1 in lane order; row f belongs to the
  // first lane whose inclusive prefix exceeds f.  kU rows are handled per round: all their loads are issued before the
  // first store, so a warp keeps kU * ceil(XL / 32) independent 128-byte requests in flight instead of one (a
  // synchronised flush emits 21 rows for every agent of the warp at once).
  constexpr int kU = 4;
  const float* __restrict__ ring = p.b.obs_ring;
  float* __restrict__ out_x = p.b.out_x;
  for (int f0 = 0; f0 < warp_total; f0 += kU) {
    const float* x[kU];
    float* dst[kU];
    size_t ra[kU];   // index into rew_ring / act_ring
    int row[kU];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int f = f0 + u;
      const int src = __popc(__ballot_sync(0xffffffffu, incl <= f)) & 31;   // owner lane (any lane when f is out of range: nothing is accessed)
      const int s_incl = __shfl_sync(0xffffffffu, incl, src);
      const int s_cnt = __shfl_sync(0xffffffffu, n_emit, src);
      const int s_g = __shfl_sync(0xffffffffu, g, src);
      const int s_off = __shfl_sync(0xffffffffu, first_t_off, src);
      const int s_nm = __shfl_sync(0xffffffffu, n_main, src);
      const int e = f - (s_incl - s_cnt);
      row[u] = (f < warp_total) ? warp_base + f : p.b.capacity;  // rows past the capacity are dropped (out_count tells)
      const int k = e < s_nm ? e : s_off;                         // the optional extra row is the last list element
      const int sl = ring_slot(slot_now, s_off, (f < warp_total) ? k : 0, R);
      ra[u] = (size_t)sl * N + s_g;
      x[u] = ring + ra[u] * L + 1;                                // drop the is_learning column
      dst[u] = out_x + (size_t)row[u] * XL;
    }
    for (int q0 = 0; q0 < XL; q0 += 32) {
      const int q = q0 + lane;
      float v[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) v[u] = (q < XL && row[u] < p.b.capacity) ? x[u][q] : 0.f;
#pragma unroll
      for (int u = 0; u < kU; ++u)
        if (q < XL && row[u] < p.b.capacity) dst[u][q] = v[u];
    }
    int rr = p.b.capacity;   // lane u < kU writes r_ and a_ of row u
    size_t ia = 0;
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if (lane == u) { rr = row[u]; ia = ra[u]; }
    if (rr < p.b.capacity) {
      p.b.out_r[rr] = p.b.rew_ring[ia];
      p.b.out_a[rr] = p.b.act_ring[ia];
    }
  }
}

// Per-world episode statistics (ProcessAgent.run :220-243, ProcessStats.run :62-117): score = sum over learning
// agents of their rewards / number of learning agents; length = frames pushed to the trainer.
// stats[0] += finished episodes, stats[1] += sum of scores, stats[2] += learning-agent steps of finished episodes.
__global__ void ga3c_episode_stats_kernel(const float* __restrict__ obs_now, const float* __restrict__ reward,
                                          const uint8_t* __restrict__ over, float* __restrict__ ep_reward,
                                          int32_t* __restrict__ ep_steps, double* __restrict__ stats, int W, int A,
                                          int L) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  float sum = 0.f;
  int learners = 0;
  for (int i = 0; i < A; ++i) {
    const size_t g = (size_t)w * A + i;
    if (obs_now[g * L] != 0.f) { sum += reward[g]; ++learners; }
  }
  float acc = ep_reward[w] + sum;
  int steps = ep_steps[w] + learners;
  if (over[w]) {
    atomicAdd(&stats[0], 1.0);
    atomicAdd(&stats[1], learners > 0 ? (double)acc / learners : 0.0);
    atomicAdd(&stats[2], (double)steps);
    acc = 0.f;
    steps = 0;
  }
  ep_reward[w] = acc;
  ep_steps[w] = steps;
}

}  // namespace ca

namespace ca {

// One LSTM time step of the predictor, fused: input projection (7 -> 256, weights in shared memory), recurrent
// pre-activation add, gates, state update and the dynamic_rnn sequence-length mask in a single pass.
// TF-1.15 LSTMCell semantics (GA3C/NetworkVP_rnn.py:63-66): gate order i, j, f, o; c' = sigmoid(f + 1) c + sigmoid(i) tanh(j);
// h' = sigmoid(o) tanh(c'); rows with t >= num_other_agents keep their state.  Reads the raw observation rows
// (stride obs_stride floats; num_other_agents at column 1, the t-th other agent at columns 6 + 7t .. 6 + 7t + 6) and
// normalises on the fly with (x - avg) / std, so no normalised copy of the observations is ever materialised.
__global__ void __launch_bounds__(256) lstm_step_kernel(const float* __restrict__ obs, int obs_stride,
                                                        const float* __restrict__ zh, const float* __restrict__ Kx,
                                                        const float* __restrict__ bias, const float* __restrict__ avg7,
                                                        const float* __restrict__ std7, float* __restrict__ c,
                                                        float* __restrict__ h, int B, int t) {
  __shared__ float sK[7 * 256];
  __shared__ float sb[256];
  __shared__ float sa[7], ss[7];
  for (int q = threadIdx.x; q < 7 * 256; q += 256) sK[q] = Kx[q];
  sb[threadIdx.x] = bias[threadIdx.x];
  if (threadIdx.x < 7) { sa[threadIdx.x] = avg7[threadIdx.x]; ss[threadIdx.x] = std7[threadIdx.x]; }
  __syncthreads();
  const int u = threadIdx.x & 63;
  const long row = (long)blockIdx.x * 4 + (threadIdx.x >> 6);
  if (row >= B) return;
  const float* o = obs + row * obs_stride;
  if (!(o[1] > (float)t)) return;  // sequence_length = raw num_other_agents
  float x[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) x[k] = (o[6 + 7 * t + k] - sa[k]) / ss[k];
  float g[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int col = q * 64 + u;
    float z = sb[col] + (zh ? zh[row * 256 + col] : 0.f);
#pragma unroll
    for (int k = 0; k < 7; ++k) z += x[k] * sK[k * 256 + col];
    g[q] = z;
  }
  const float cp = c[row * 64 + u];
  const float sig_i = 1.f / (1.f + expf(-g[0])), sig_f = 1.f / (1.f + expf(-(g[2] + 1.f))), sig_o = 1.f / (1.f + expf(-g[3]));
  const float cn = sig_f * cp + sig_i * tanhf(g[1]);
  c[row * 64 + u] = cn;
  h[row * 64 + u] = sig_o * tanhf(cn);
}

}  // namespace ca
