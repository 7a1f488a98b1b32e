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
// range being claimed with one atomicAdd per warp; a second kernel (ga3c_gather_kernel) copies the 26/68-float rows.
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

  // The thread lists its rows for the gather kernel: out_src[row] = (ring slot, agent slot) of the experience.  The
  // 4 (L - 1)-byte observation rows themselves are copied by ga3c_gather_kernel, which runs over the flat row range with
  // many independent loads in flight; doing the copy here (one warp walking its agents' rows one after the other) left
  // the kernel latency-bound: 66 us for 77 MB in a normal step, 608 us for 1.2 GB in a synchronised flush (profiles/).
  const int my_base = warp_base + incl - n_emit;
  for (int e = 0; e < n_emit; ++e) {
    const int row = my_base + e;
    if (row >= p.b.capacity) break;                      // overflow is reported through out_count > capacity
    const int k = e < n_main ? e : first_t_off;          // the optional extra row is the last list element
    p.b.out_src[row] = ring_slot(slot_now, first_t_off, k, R) * p.N + g;
  }
}

// Copies the training rows listed by ga3c_record_kernel: rows [*done_in, *out_count) of out_src -> out_x / out_r / out_a.
// One warp per kGatherRows rows and round; all loads of a round are issued before the first store.
constexpr int kGatherRows = 8;
struct GatherParams {
  ca_ga3c_buffers b;
  int L;
  int32_t* gathered;        // [0] rows already copied by earlier launches since the last take(), [1] block ticket
};

__global__ void __launch_bounds__(256) ga3c_gather_kernel(const __grid_constant__ GatherParams p) {
  const int lane = threadIdx.x & 31;
  const int XL = p.L - 1;
  const int begin = p.gathered[0];
  int end = *p.b.out_count;
  if (end > p.b.capacity) end = p.b.capacity;
  const int warps = gridDim.x * (blockDim.x >> 5);
  const int gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const float* __restrict__ ring = p.b.obs_ring;
  float* __restrict__ out_x = p.b.out_x;
  for (int r0 = begin + gw * kGatherRows; r0 < end; r0 += warps * kGatherRows) {
    const int mine = r0 + lane;
    int src_l = 0;
    if (lane < kGatherRows && mine < end) {
      src_l = p.b.out_src[mine];
      p.b.out_r[mine] = p.b.rew_ring[src_l];
      p.b.out_a[mine] = p.b.act_ring[src_l];
    }
    const float* x[kGatherRows];
    float* dst[kGatherRows];
    bool ok[kGatherRows];
#pragma unroll
    for (int u = 0; u < kGatherRows; ++u) {
      const int src = __shfl_sync(0xffffffffu, src_l, u);
      ok[u] = r0 + u < end;
      x[u] = ring + (size_t)src * p.L + 1;              // drop the is_learning column
      dst[u] = out_x + (size_t)(r0 + u) * XL;
    }
    for (int q0 = 0; q0 < XL; q0 += 32) {
      const int q = q0 + lane;
      float v[kGatherRows];
#pragma unroll
      for (int u = 0; u < kGatherRows; ++u) v[u] = (q < XL && ok[u]) ? x[u][q] : 0.f;
#pragma unroll
      for (int u = 0; u < kGatherRows; ++u)
        if (q < XL && ok[u]) dst[u][q] = v[u];
    }
  }
  // the last block to finish publishes the new watermark (every block has read the old one by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&p.gathered[1], 1) == (int)gridDim.x - 1) {
      p.gathered[0] = end;
      p.gathered[1] = 0;
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

namespace ca {

// ---- trainer: the LSTM cell of NetworkVP_rnn as one forward and one backward kernel ---------------------------------------
// GA3C/NetworkVP_rnn.py:63-66 builds tf.nn.rnn_cell.LSTMCell(64) under dynamic_rnn(sequence_length); TF differentiates it op
// by op.  A framework re-expression does the same with ~12 elementwise kernels forward and ~25 backward per time step,
// each streaming [B, 64..256] tensors through HBM.  Here the whole cell is one pass: thread = (row, unit).
//   forward : z [B][256] = x_t Kx + h Kh + b (gate order i, j, f, o) -> gate activations (saved for backward), c', h';
//             rows with t >= sequence_length keep c, h (dynamic_rnn copies the state through).
//   backward: (dc', dh') -> dz [B][256], dc, and the part of dh that bypasses the cell for masked rows.
// float32 with the accurate expf / tanhf (the trainer must match the framework's fp32 gradients, tests/test_gpu_ga3c.py).
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(const float* __restrict__ z, const float* __restrict__ c_prev,
                                                            const float* __restrict__ h_prev,
                                                            const float* __restrict__ seq_len, int seq_stride, int t,
                                                            float* __restrict__ gates, float* __restrict__ c,
                                                            float* __restrict__ h, int B) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long row = idx >> 6;
  const int u = (int)(idx & 63);
  if (row >= B) return;
  const float* zr = z + row * 256;
  const float si = sigmoid_acc(zr[u]), tj = tanhf(zr[64 + u]), sf = sigmoid_acc(zr[128 + u] + 1.0f), so = sigmoid_acc(zr[192 + u]);
  float* gr = gates + row * 256;
  gr[u] = si; gr[64 + u] = tj; gr[128 + u] = sf; gr[192 + u] = so;
  const float cp = c_prev[row * 64 + u];
  const bool live = seq_len[row * (long)seq_stride] > (float)t;
  const float cn = sf * cp + si * tj;
  c[row * 64 + u] = live ? cn : cp;
  h[row * 64 + u] = live ? so * tanhf(cn) : h_prev[row * 64 + u];
}

__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                                            const float* __restrict__ c_new,
                                                            const float* __restrict__ seq_len, int seq_stride, int t,
                                                            const float* __restrict__ dc, const float* __restrict__ dh,
                                                            float* __restrict__ dz, float* __restrict__ dc_prev,
                                                            float* __restrict__ dh_pass, int B) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long row = idx >> 6;
  const int u = (int)(idx & 63);
  if (row >= B) return;
  const long k = row * 64 + u;
  const float gdc = dc ? dc[k] : 0.f, gdh = dh ? dh[k] : 0.f;
  float* dzr = dz + row * 256;
  const bool live = seq_len[row * (long)seq_stride] > (float)t;
  if (!live) {
    dzr[u] = 0.f; dzr[64 + u] = 0.f; dzr[128 + u] = 0.f; dzr[192 + u] = 0.f;
    dc_prev[k] = gdc;
    dh_pass[k] = gdh;
    return;
  }
  const float* gr = gates + row * 256;
  const float si = gr[u], tj = gr[64 + u], sf = gr[128 + u], so = gr[192 + u];
  const float tc = tanhf(c_new[k]);
  const float dco = gdc + gdh * so * (1.f - tc * tc);
  dzr[u] = dco * tj * si * (1.f - si);
  dzr[64 + u] = dco * si * (1.f - tj * tj);
  dzr[128 + u] = dco * c_prev[k] * sf * (1.f - sf);
  dzr[192 + u] = gdh * tc * so * (1.f - so);
  dc_prev[k] = dco * sf;
  dh_pass[k] = 0.f;
}

}  // namespace ca

