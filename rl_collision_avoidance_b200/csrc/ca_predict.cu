// ca_predict.cu — the GA3C predictor (ThreadPredictor.run + NetworkVP_rnn forward) as ONE fused sm_100a kernel.
//
// Reference: GA3C/ThreadPredictor.py:40-75 (batch -> predict_p_and_v), GA3C/NetworkVP_rnn.py:39-108 (input
// normalisation, LSTMCell(64) over the other-agent rows with sequence_length = num_other_agents, concat(host, h) ->
// Dense256 x3 (ReLU)), GA3C/NetworkVPCore.py:64-77 (softmax policy head with MIN_POLICY, value head),
// GA3C/ProcessAgent.py:98-103 (select_action: np.random.choice(p) or argmax).
//
// Design (B200): a tile is 128 observation rows (= UMMA M).  Every matrix product of the network runs on the 5th-gen
// tensor cores as tcgen05.mma.cta_group::1.kind::f16 (fp16 operands, fp32 accumulation) with
//   * A = the tile's activations, written by the CTA's own epilogue threads into shared memory in the canonical
//     K-major no-swizzle core-matrix layout (8 rows x 16 bytes per core matrix; k-group-major: [k/8][row][8 halfs]),
//   * B = the layer's weights, pre-packed ONCE per weight update into exactly that shared-memory image
//     (ca_predictor_pack) so that a chunk of weights is one contiguous TMA bulk copy (cp.async.bulk + mbarrier),
//   * D = a 128 x 256 fp32 accumulator in tensor memory (256 TMEM columns; two CTAs per SM share the 512),
// and everything between two products (LSTM gates and state update with the dynamic_rnn sequence mask, bias, ReLU,
// fp16 repack, softmax, action sampling) happens in the epilogue of the product that feeds it: two threads own a row
// (TMEM lane r; one column half each, 8 warps per tile), read their part of the accumulator row with tcgen05.ld — one
// block ahead of the math — and write the next A operand.  Two CTAs share an SM, so one tile's products run while the
// other tile's epilogue keeps the special-function unit busy.  The LSTM biases ride in the product (a weight row that
// meets a constant-1 input), the 1/2 of sigmoid(x) = 0.5 tanh(x/2) + 0.5 is folded into the packed weights and h is kept
// doubled, which leaves 5 tanh.approx + 5 arithmetic instructions per unit and step.  Per observation row the kernel
// reads L floats from HBM and writes 11 + 1 floats (+ 1 int32 action); no intermediate ever leaves the SM.
// fp16 operands carry an 11-bit significand (the same as TF32) with fp32 accumulation; tests/test_gpu_predictor.py
// compares against the fp32 PyTorch network with the tolerance stated there.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/ca_step.h"

namespace cap {

constexpr int kRows = 128;      // rows per tile = UMMA M = TMEM lanes
constexpr int kThreads = 256;   // two threads per row: warp w reads TMEM lanes 32 (w & 3) .. +31 (the hardware's lane
                                // window of a warp) and owns the column half (w >> 2) of every epilogue
constexpr int kHid = 64;        // LSTM units
constexpr int kN = 256;         // gate pre-activations (4 x 64) and dense width
constexpr int kOutN = 16;       // policy logits (11) + value (1), padded to the smallest UMMA N
constexpr int kKgA = kRows * 16;   // bytes of one k-group (8 halfs) of an A operand: 128 rows x 16 B
constexpr int kKgB = kN * 16;      // ... of a 256-row B operand
constexpr int kKgOut = kOutN * 16; // ... of the 16-row head weights

// ---- packed parameter blob (device memory, written by pack_kernel) -----------------------------------------------------
constexpr int kOffWLstm = 0;                          // [10][256][8] halfs: k 0..63 = h rows (x 1/2), 64..70 = x rows, 71 = bias, rest 0
constexpr int kOffWL1 = kOffWLstm + 10 * kKgB;        // [10][256][8]: k 0..63 = h rows (x 1/2), 64..67 = host rows, rest 0
constexpr int kOffWL2 = kOffWL1 + 10 * kKgB;          // [32][256][8]
constexpr int kOffWFc1 = kOffWL2 + 32 * kKgB;         // [32][256][8]
constexpr int kOffWOut = kOffWFc1 + 32 * kKgB;        // [32][16][8]: n 0..10 logits_p, 11 logits_v, rest 0
constexpr int kOffF32 = kOffWOut + 32 * kKgOut;       // float section
constexpr int kFbLstm = 0, kFbL1 = 256, kFbL2 = 512, kFbFc1 = 768, kFbOut = 1024;  // biases (LSTM: forget bias and the sigmoid 1/2 folded in)
constexpr int kFAvgO = 1040, kFIstdO = 1048, kFAvgH = 1056, kFIstdH = 1060;        // input normalisation
constexpr int kFwV = 1064;                            // logits_v kernel in float32 (the value head runs on the CUDA cores)
constexpr int kF32Count = 1064 + 256;
constexpr int kBlobBytes = kOffF32 + kF32Count * 4;   // 357 536
constexpr int kFVpart = kF32Count;                    // shared memory only: the two column halves' partial value sums [2][128]
constexpr int kF32Smem = kF32Count + 2 * 128;
static_assert(kBlobBytes == CA_PREDICTOR_BLOB_BYTES, "include/ca_step.h disagrees with the blob layout");

// ---- shared memory carve-up of one CTA ----------------------------------------------------------------------------------
constexpr int kSmAct = 0;                    // 64 KB: dense activations [32 kg][128][8]; during the LSTM phase:
                                             //   kg 0..7 = h, kg 8..8+M-1 = x_t, kg 8+M = host, kg 9+M = zeros
constexpr int kSmW = 32 * kKgA;              // 40 KB weight stage (LSTM / layer1 image, or 2 x 16 KB chunks, or head)
constexpr int kSmF32 = kSmW + 10 * kKgB;     // float section copy
constexpr int kSmBar = kSmF32 + ((kF32Smem * 4 + 127) / 128) * 128;
constexpr int kSmTotal = kSmBar + 128;       // 12 mbarriers, TMEM base, max sequence length of the tile
// The 128 KB of a dense layer's weights stream in chunks of kUmmaPerChunk K = 16 products through the 40 KB weight region.
// The copy path favours large bulk copies: two 16 KB stages (0.208 / 0.184 ms at M = 9 / 3) beat five 8 KB stages
// (0.211 / 0.196 ms) although fewer copies are in flight.
constexpr int kUmmaPerChunk = 2;
constexpr int kChunkBytes = kUmmaPerChunk * 2 * kKgB;   // 16 KB
constexpr int kStages = 40960 / kChunkBytes;            // 2
constexpr int kChunks = 16 / kUmmaPerChunk;             // 8 per layer
constexpr int kMaxOthers = 22;               // (8 + M + 2) k-groups must fit the activation buffer

struct Params {
  const float* obs;      // [B][stride] raw observation rows (column 0 is_learning, 1 num_other_agents, 2..5 host, 6.. others)
  int stride, B, M;
  const unsigned char* blob;
  float* p;              // [B][11] or null
  float* v;              // [B] or null
  int32_t* actions;      // [B] or null
  int greedy;            // argmax instead of sampling
  float min_policy;
  unsigned long long seed, offset;
  int* error;            // device int, set when a barrier wait times out
  // optional row plan (ca_predict_plan): the kernel processes rows row_index[0 .. *n_rows) instead of 0 .. B
  const int32_t* row_index;
  const int32_t* n_rows;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a descriptor / protocol bug must not hang the GPU.  ~2 s at 2 GHz, then flag + trap.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) {
      if (err) atomicExch(err, 1);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load(uint32_t dst, const void* src, int bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 at [0,14), leading
// byte offset >> 4 at [16,30) = distance between the two core matrices of one K = 16 step, stride byte offset >> 4 at
// [32,46) = distance between 8-row groups, version 1 at [46,48), layout type 0 (no swizzle) at [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D fp32 (1 at [4,6)), A/B fp16 (0), both K-major, N >> 3 at
// [17,23), M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t instr_desc(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// tcgen05.ld 32x32b: thread i of the warp reads lane (warp's base lane + i), 16 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -60000.f), 60000.f), fminf(fmaxf(b, -60000.f), 60000.f));
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_h2_raw(float a, float b) {  // values known to be inside the fp16 range
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Philox-free counter hash (splitmix64 finaliser) -> uniform in [0, 1): one draw per (seed, call, row).
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long offset, long row) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(row + 1) + 0xD1B54A32D192ED03ull * (offset + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// Wait for the accumulator of the product in flight.  Only the issuing thread polls the mbarrier; the other warps sleep
// in the CTA barrier, so the waiting tile does not take issue slots from the CTA that shares the SM (with every thread
// polling, the try_wait loops were 16 % of all executed instructions, profiles/).
__device__ __forceinline__ void acc_wait(uint32_t bar, uint32_t& phase, int* err, int tid) {
  if (tid == 0) mbar_wait(bar, phase, err);
  phase ^= 1;
  __syncthreads();
  tc_fence_after();
}

// bias + ReLU + fp16 repack of the thread's accumulator row -> activation buffer (the next product's A operand).
// kValue (fullyconnected1 only): the value head is evaluated right here in float32 — the ReLU outputs BEFORE their fp16
// rounding times the float32 logits_v kernel — and the thread's partial sum over its 128 columns is returned.  The
// value is what the actors bootstrap their n-step returns from (ProcessAgent.py:71-76) while the trainer evaluates V in
// float32; the fp16 rounding of the 256 head inputs was the largest single contribution to |dv| (oracle/
// network_oracle.py, tests/test_network.py): 1.4e-2 -> 1.9e-3 with the trained IROS18 weights.
template <bool kValue>
__device__ __forceinline__ float dense_epilogue(uint32_t tmem_row, const float* bias, uint32_t act_row, int half,
                                                const float* wv = nullptr) {
  float vacc = 0.f;
  constexpr int kIters = kN / 2 / 32;   // this thread's 128 columns in blocks of 32, loads one block ahead of the math
  float buf[2][32];
  const int cbase = half * (kN / 2);
  tmem_ld16(tmem_row + cbase, buf[0]);
  tmem_ld16(tmem_row + cbase + 16, buf[0] + 16);
#pragma unroll
  for (int it = 0; it < kIters; ++it) {
    const int c0 = cbase + it * 32;
    const float* a = buf[it & 1];
    tmem_ld_wait();
    if (it + 1 < kIters) {
      tmem_ld16(tmem_row + c0 + 32, buf[(it + 1) & 1]);
      tmem_ld16(tmem_row + c0 + 48, buf[(it + 1) & 1] + 16);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int c = q * 8 + e * 2;
        const float2 b = *reinterpret_cast<const float2*>(bias + c0 + c);
        const float y0 = fminf(fmaxf(a[c] + b.x, 0.f), 60000.f), y1 = fminf(fmaxf(a[c + 1] + b.y, 0.f), 60000.f);
        if (kValue) {
          const float2 wvv = *reinterpret_cast<const float2*>(wv + c0 + c);
          vacc = __fmaf_rn(y0, wvv.x, vacc);
          vacc = __fmaf_rn(y1, wvv.y, vacc);
        }
        w[e] = pack_h2_raw(y0, y1);
      }
      st_shared_v4(act_row + (uint32_t)(c0 / 8 + q) * kKgA, w[0], w[1], w[2], w[3]);
    }
  }
  return vacc;
}

__global__ void __launch_bounds__(kThreads, 2) predict_kernel(const Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int half = warp >> 2;                        // column half of the epilogues this thread owns
  const int r = ((warp & 3) << 5) | (tid & 31);      // tile row = TMEM lane
  const uint32_t s_act = smem_u32(smem + kSmAct), s_w = smem_u32(smem + kSmW);
  float* fsec = reinterpret_cast<float*>(smem + kSmF32);
  const uint32_t bar0 = smem_u32(smem + kSmBar);
  // barriers: [0] whole-image loads (LSTM / layer1 / head weights), [1] accumulator, [2..6] ring stage filled,
  // [7..11] ring stage drained
  const uint32_t bar_img = bar0, bar_acc = bar0 + 8, bar_w0 = bar0 + 16, bar_e0 = bar0 + 16 + 8 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kSmBar + 104);
  int* smax = reinterpret_cast<int*>(smem + kSmBar + 112);
  const int M = p.M;
  const int kg_host = 8 + M, kg_zero = 9 + M;
  const long n_rows = p.n_rows ? (long)*p.n_rows : (long)p.B;
  const long n_tiles = (n_rows + kRows - 1) / kRows;

  // ---- one-time setup: parameters, barriers, tensor memory
  {
    const float* src = reinterpret_cast<const float*>(p.blob + kOffF32);
    for (int q = tid; q < kF32Count; q += kThreads) fsec[q] = src[q];
  }
  if (tid == 0) {
    for (int b = 0; b < 2 + 2 * kStages; ++b) mbar_init(bar0 + 8 * b, 1);
    *smax = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_row = tmem + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 lanes
  const uint32_t act_row = s_act + (uint32_t)r * 16;                      // this row's 16 bytes inside every k-group

  uint32_t ph_img = 0, ph_acc = 0, ph_w = 0, ph_e = 0;   // ph_w / ph_e: one phase bit per ring stage
  constexpr uint32_t kIdesc256 = instr_desc(kRows, kN), kIdesc16 = instr_desc(kRows, kOutN);

  if (tid == 0 && (long)blockIdx.x < n_tiles) {  // LSTM weights of the first tile
    mbar_expect_tx(bar_img, 10 * kKgB);
    tma_load(s_w, p.blob + kOffWLstm, 10 * kKgB, bar_img);
  }

  for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long pos = tile * kRows + r;
    const bool ok = pos < n_rows;
    const long row = ok ? (p.row_index ? (long)p.row_index[pos] : pos) : 0;
    const float* o = p.obs + row * (long)p.stride;
    // ---- stage the tile: sequence length, host features, the M other-agent rows (normalised, fp16), h = 0
    int seq = 0;
    {
      float nf = ok ? o[1] : 0.f;
      nf = fminf(fmaxf(nf, 0.f), (float)M);
      seq = (int)ceilf(nf);  // dynamic_rnn runs step t for rows with t < sequence_length
      const int wmax = __reduce_max_sync(0xffffffffu, seq);
      if ((tid & 31) == 0 && wmax > 0) atomicMax(smax, wmax);
      // the two threads of a row share the staging: h = 0 for their own unit half, other agents t = half, half + 2, ...
      for (int kg = half * 4; kg < half * 4 + 4; ++kg) st_shared_v4(act_row + kg * kKgA, 0u, 0u, 0u, 0u);
      if (half == 0) {
        st_shared_v4(act_row + kg_zero * kKgA, 0u, 0u, 0u, 0u);
        float hf[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) hf[k] = ok ? (o[2 + k] - fsec[kFAvgH + k]) * fsec[kFIstdH + k] : 0.f;
        st_shared_v4(act_row + kg_host * kKgA, pack_h2(hf[0], hf[1]), pack_h2(hf[2], hf[3]), 0u, 0u);
      }
      for (int t = half; t < M; t += 2) {
        float x[8];
#pragma unroll
        for (int k = 0; k < 7; ++k) x[k] = ok ? (o[6 + 7 * t + k] - fsec[kFAvgO + k]) * fsec[kFIstdO + k] : 0.f;
        x[7] = ok ? 1.f : 0.f;   // meets the bias row of the packed LSTM kernel: the bias add happens inside the product
        st_shared_v4(act_row + (8 + t) * kKgA, pack_h2(x[0], x[1]), pack_h2(x[2], x[3]), pack_h2(x[4], x[5]),
                     pack_h2(x[6], x[7]));
      }
    }
    float c[kHid / 2];   // cell state of this thread's 32 units
#pragma unroll
    for (int u = 0; u < kHid / 2; ++u) c[u] = 0.f;
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    const int steps = *smax;  // LSTM steps any row of this tile needs (CTA-uniform)

    // ---- LSTM over the other agents (GA3C/NetworkVP_rnn.py:58-66)
    for (int t = 0; t < steps; ++t) {
      if (tid == 0) {
        if (t == 0) { mbar_wait(bar_img, ph_img, p.error); ph_img ^= 1; }
        tc_fence_after();
        // x_t part: one K = 16 product (k-group 8+t of A and the k-group after it, which meets zero weights)
        umma(tmem, smem_desc(s_act + (8 + t) * kKgA, kKgA, 128), smem_desc(s_w + 8 * kKgB, kKgB, 128), kIdesc256, 0);
        if (t > 0) {  // h part (h = 0 at t = 0)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma(tmem, smem_desc(s_act + ks * 2 * kKgA, kKgA, 128), smem_desc(s_w + ks * 2 * kKgB, kKgB, 128),
                 kIdesc256, 1);
        }
        umma_commit(bar_acc);
      }
      acc_wait(bar_acc, ph_acc, p.error, tid);
      if (tid == 0 && t == steps - 1) {  // the LSTM image is no longer needed: fetch layer1's behind the gate math
        mbar_expect_tx(bar_img, 10 * kKgB);
        tma_load(s_w, p.blob + kOffWL1, 10 * kKgB, bar_img);
      }
      const bool live = t < seq;
      // Tensor memory reads are slow (64 B per clock and SM: the 128 KB accumulator of a tile takes ~2000 cycles per step,
      // as long as its gate math), so the loads of block b + 1 are in flight while block b is evaluated: two register
      // buffers, tcgen05.wait::ld right before the next block's loads are issued.
      constexpr int kBlocks = kHid / 2 / 8;
      float g[2][32];
      auto load_block = [&](int b, float* dst) {
        const int u0 = half * (kHid / 2) + b * 8;
        tmem_ld8(tmem_row + u0, dst);
        tmem_ld8(tmem_row + 64 + u0, dst + 8);
        tmem_ld8(tmem_row + 128 + u0, dst + 16);
        tmem_ld8(tmem_row + 192 + u0, dst + 24);
      };
      load_block(0, g[0]);
#pragma unroll
      for (int b = 0; b < kBlocks; ++b) {
        const int u0 = half * (kHid / 2) + b * 8;   // first of 8 units = one k-group of h
        const float* gi = g[b & 1];
        const float *gj = gi + 8, *gf = gi + 16, *go = gi + 24;
        tmem_ld_wait();
        if (b + 1 < kBlocks) load_block(b + 1, g[(b + 1) & 1]);
        float hn[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          // sigmoid(x) = 0.5 tanh(x / 2) + 0.5.  The packed kernel carries the 1/2 of the i, f, o columns, the biases (as
          // the weight row that meets the constant 1 of x_t, forget bias 1.0 included) and a factor 1/2 on the rows that
          // multiply h, because what is stored as h is 2h:
          //   c' = sig(f) c + sig(i) tanh(j) = 0.5 [(tf c + c) + (ti tj + tj)],   2 h' = 2 sig(o) tanh(c') = to tc + tc
          // i.e. 5 tanh.approx + 3 FMA + 1 add + 1 multiply per unit.
          const float ti = tanh_fast(gi[q]), tj = tanh_fast(gj[q]), tf = tanh_fast(gf[q]), to = tanh_fast(go[q]);
          const float cn = 0.5f * (__fmaf_rn(tf, c[b * 8 + q], c[b * 8 + q]) + __fmaf_rn(ti, tj, tj));
          const float tc = tanh_fast(cn);
          hn[q] = __fmaf_rn(to, tc, tc);
          if (live) c[b * 8 + q] = cn;
        }
        if (live)  // rows whose sequence ended keep c and h
          st_shared_v4(act_row + (u0 / 8) * kKgA, pack_h2_raw(hn[0], hn[1]), pack_h2_raw(hn[2], hn[3]),
                       pack_h2_raw(hn[4], hn[5]), pack_h2_raw(hn[6], hn[7]));
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
    }

    // ---- layer1: relu(concat(host, h) @ W1 + b1)
    if (tid == 0) {
      *smax = 0;
      if (steps == 0) {  // no row had another agent: the stage still holds (or is receiving) the LSTM image
        mbar_wait(bar_img, ph_img, p.error); ph_img ^= 1;
        mbar_expect_tx(bar_img, 10 * kKgB);
        tma_load(s_w, p.blob + kOffWL1, 10 * kKgB, bar_img);
      }
      mbar_wait(bar_img, ph_img, p.error); ph_img ^= 1;
      tc_fence_after();
      umma(tmem, smem_desc(s_act + kg_host * kKgA, kKgA, 128), smem_desc(s_w + 8 * kKgB, kKgB, 128), kIdesc256, 0);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma(tmem, smem_desc(s_act + ks * 2 * kKgA, kKgA, 128), smem_desc(s_w + ks * 2 * kKgB, kKgB, 128), kIdesc256, 1);
      umma_commit(bar_acc);
    }
    acc_wait(bar_acc, ph_acc, p.error, tid);

    // ---- layer2 and fullyconnected1: the 128 KB of weights of a layer stream through the ring of kStages stages
    // (kUmmaPerChunk K = 16 products per chunk).  A stage is refilled one chunk behind the product that reads it, so the
    // issuing thread never waits for the product it has just issued.
#pragma unroll 1
    for (int layer = 0; layer < 2; ++layer) {
      const unsigned char* wsrc = p.blob + (layer == 0 ? kOffWL2 : kOffWFc1);
      if (tid == 0) {  // the previous product is complete: its weights may be overwritten while its epilogue runs
        for (int st = 0; st < kStages; ++st) {
          mbar_expect_tx(bar_w0 + 8 * st, kChunkBytes);
          tma_load(s_w + st * kChunkBytes, wsrc + st * kChunkBytes, kChunkBytes, bar_w0 + 8 * st);
        }
      }
      dense_epilogue<false>(tmem_row, fsec + (layer == 0 ? kFbL1 : kFbL2), act_row, half);
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        int st = 0, st_prev = kStages - 1;
#pragma unroll 1
        for (int ck = 0; ck < kChunks; ++ck) {
          mbar_wait(bar_w0 + 8 * st, (ph_w >> st) & 1u, p.error); ph_w ^= 1u << st;
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < kUmmaPerChunk; ++ks)
            umma(tmem, smem_desc(s_act + (ck * kUmmaPerChunk + ks) * 2 * kKgA, kKgA, 128),
                 smem_desc(s_w + st * kChunkBytes + ks * 2 * kKgB, kKgB, 128), kIdesc256, (ck | ks) != 0);
          umma_commit(bar_e0 + 8 * st);
          if (ck >= 1) {   // the stage of chunk ck - 1: wait for its product (the drained barrier is consumed every time,
                           // so its phase stays in step) and refill it with chunk ck - 1 + kStages if there is one
            mbar_wait(bar_e0 + 8 * st_prev, (ph_e >> st_prev) & 1u, p.error); ph_e ^= 1u << st_prev;
            if (ck - 1 + kStages < kChunks) {
              mbar_expect_tx(bar_w0 + 8 * st_prev, kChunkBytes);
              tma_load(s_w + st_prev * kChunkBytes, wsrc + (ck - 1 + kStages) * kChunkBytes, kChunkBytes, bar_w0 + 8 * st_prev);
            }
          }
          st_prev = st;
          st = st + 1 == kStages ? 0 : st + 1;
        }
        mbar_wait(bar_e0 + 8 * st_prev, (ph_e >> st_prev) & 1u, p.error); ph_e ^= 1u << st_prev;   // last chunk: all products done
        umma_commit(bar_acc);
      }
      acc_wait(bar_acc, ph_acc, p.error, tid);
    }

    // ---- heads: logits_p (11) and logits_v (1) as one N = 16 product
    if (tid == 0) {
      mbar_expect_tx(bar_img, 32 * kKgOut);
      tma_load(s_w, p.blob + kOffWOut, 32 * kKgOut, bar_img);
    }
    fsec[kFVpart + half * kRows + r] = dense_epilogue<true>(tmem_row, fsec + kFbFc1, act_row, half, fsec + kFwV);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(bar_img, ph_img, p.error); ph_img ^= 1;
      tc_fence_after();
#pragma unroll 4
      for (int ks = 0; ks < 16; ++ks)
        umma(tmem, smem_desc(s_act + ks * 2 * kKgA, kKgA, 128), smem_desc(s_w + ks * 2 * kKgOut, kKgOut, 128), kIdesc16,
             ks != 0);
      umma_commit(bar_acc);
    }
    acc_wait(bar_acc, ph_acc, p.error, tid);
    if (tid == 0 && tile + gridDim.x < n_tiles) {  // next tile's LSTM image, behind the softmax
      mbar_expect_tx(bar_img, 10 * kKgB);
      tma_load(s_w, p.blob + kOffWLstm, 10 * kKgB, bar_img);
    }
    if (half == 0) {
      float z[16];
      tmem_ld16(tmem_row, z);
      tmem_ld_wait();
      const float* bo = fsec + kFbOut;
      float mx = -INFINITY;
#pragma unroll
      for (int a = 0; a < 11; ++a) { z[a] += bo[a]; mx = fmaxf(mx, z[a]); }
      float sum = 0.f;
#pragma unroll
      for (int a = 0; a < 11; ++a) { z[a] = __expf(z[a] - mx); sum += z[a]; }
      const float inv = 1.f / sum, mp = p.min_policy, den = 1.f / (1.f + mp * 11.f);
      float best = -1.f;
      int arg = 0;
#pragma unroll
      for (int a = 0; a < 11; ++a) {
        z[a] = (z[a] * inv + mp) * den;  // NetworkVPCore.py:74-75
        if (z[a] > best) { best = z[a]; arg = a; }
      }
      if (ok) {
        if (p.p) {
          float* dst = p.p + row * 11;
#pragma unroll
          for (int a = 0; a < 11; ++a) dst[a] = z[a];
        }
        if (p.v) p.v[row] = (fsec[kFVpart + r] + fsec[kFVpart + kRows + r]) + bo[11];   // float32 value head
        if (p.actions) {
          int act = arg;
          if (!p.greedy) {  // np.random.choice(actions, p=prediction): inverse CDF on one uniform draw
            const float uu = uniform01(p.seed, p.offset, row);
            float cum = 0.f;
            bool found = false;
            act = 10;
#pragma unroll
            for (int a = 0; a < 11; ++a) {  // smallest a with u < p_0 + ... + p_a
              cum += z[a];
              if (!found && uu < cum) { act = a; found = true; }
            }
          }
          p.actions[row] = act;
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // every thread has read its head outputs: the accumulator and the activation buffer are free
  }

  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
  }
}

// Pack the TF-layout fp32 parameters ([in][out] kernels) into the blob (fp16 shared-memory images + float section).
struct PackParams {
  const float *k_lstm, *b_lstm, *k_l1, *b_l1, *k_l2, *b_l2, *k_fc1, *b_fc1, *k_p, *b_p, *k_v, *b_v, *avg, *std;
  unsigned char* blob;
};

__device__ __forceinline__ __half to_h(float x) { return __float2half_rn(fminf(fmaxf(x, -60000.f), 60000.f)); }

__global__ void pack_kernel(const PackParams q) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  __half* wl = reinterpret_cast<__half*>(q.blob + kOffWLstm);
  __half* w1 = reinterpret_cast<__half*>(q.blob + kOffWL1);
  __half* w2 = reinterpret_cast<__half*>(q.blob + kOffWL2);
  __half* w3 = reinterpret_cast<__half*>(q.blob + kOffWFc1);
  __half* wo = reinterpret_cast<__half*>(q.blob + kOffWOut);
  float* f = reinterpret_cast<float*>(q.blob + kOffF32);
  if (e < 10 * kN * 8) {  // LSTM kernel [(7 + 64)][256]: TF rows 0..6 = x, 7..70 = h; and layer1 [(4 + 64)][256]
    const int k = (e / (kN * 8)) * 8 + (e & 7), n = (e >> 3) % kN;
    float a = 0.f, b = 0.f;
    if (k < 64) { a = 0.5f * q.k_lstm[(7 + k) * kN + n]; b = 0.5f * q.k_l1[(4 + k) * kN + n]; }   // rows that meet 2h
    else if (k < 71) { a = q.k_lstm[(k - 64) * kN + n]; if (k < 68) b = q.k_l1[(k - 64) * kN + n]; }
    else if (k == 71) a = q.b_lstm[n] + ((n >= 128 && n < 192) ? 1.0f : 0.f);   // bias row (forget_bias = 1.0, TF1 LSTMCell)
    if (n < 64 || n >= 128) a *= 0.5f;  // i, f, o gates: sigmoid(x) = 0.5 tanh(x / 2) + 0.5 (gate j = columns 64..127)
    wl[e] = to_h(a);
    w1[e] = to_h(b);
  }
  if (e < 32 * kN * 8) {
    const int k = (e / (kN * 8)) * 8 + (e & 7), n = (e >> 3) % kN;
    w2[e] = to_h(q.k_l2[k * kN + n]);
    w3[e] = to_h(q.k_fc1[k * kN + n]);
  }
  if (e < 32 * kOutN * 8) {
    const int k = (e / (kOutN * 8)) * 8 + (e & 7), n = (e >> 3) % kOutN;
    float a = 0.f;
    if (n < 11) a = q.k_p[k * 11 + n]; else if (n == 11) a = q.k_v[k];
    wo[e] = to_h(a);
  }
  if (e < 256) {
    const float bl = q.b_lstm[e] + ((e >= 128 && e < 192) ? 1.0f : 0.f);  // forget_bias = 1.0 (TF1 LSTMCell default)
    f[kFbLstm + e] = (e < 64 || e >= 128) ? 0.5f * bl : bl;
    f[kFbL1 + e] = q.b_l1[e];
    f[kFbL2 + e] = q.b_l2[e];
    f[kFbFc1 + e] = q.b_fc1[e];
  }
  if (e < 16) f[kFbOut + e] = e < 11 ? q.b_p[e] : (e == 11 ? q.b_v[0] : 0.f);
  if (e < 256) f[kFwV + e] = q.k_v[e];
  if (e < 8) {  // NN input index = observation column - 1: host = 1..4, first other agent = 5..11
    f[kFAvgO + e] = e < 7 ? q.avg[5 + e] : 0.f;
    f[kFIstdO + e] = e < 7 ? 1.f / q.std[5 + e] : 0.f;
  }
  if (e < 4) {
    f[kFAvgH + e] = q.avg[1 + e];
    f[kFIstdH + e] = 1.f / q.std[1 + e];
  }
}

// ---- row plan: which rows need a prediction, grouped by LSTM sequence length ---------------------------------------------
// Only learning agents act on a prediction (GA3C/ProcessAgent.py:128-133 asks the predictor for agents whose is_learning
// observation is set; absent slots and non-learning agents never reach ThreadPredictor), and dynamic_rnn runs
// num_other_agents steps per row.  The plan lists the learning rows sorted by descending sequence length, so that a tile
// of 128 consecutive plan entries shares (almost always) one sequence length and its LSTM loop runs exactly that many
// steps instead of the maximum.  counters: [0] = number of planned rows, [1 + b] = rows with sequence length b,
// [32 + b] = scatter cursor of bucket b (all zeroed by the caller of plan_count_kernel).
constexpr int kPlanBuckets = kMaxOthers + 1;

__device__ __forceinline__ int plan_bucket(const float* o, int M) {  // -1: no prediction needed
  if (o[0] == 0.f) return -1;
  const float nf = fminf(fmaxf(o[1], 0.f), (float)M);
  return (int)ceilf(nf);
}

__global__ void __launch_bounds__(256) plan_count_kernel(const float* __restrict__ obs, int stride, int B, int M,
                                                          int32_t* __restrict__ counters) {
  __shared__ int hist[kPlanBuckets];
  if (threadIdx.x < kPlanBuckets) hist[threadIdx.x] = 0;
  __syncthreads();
  const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row < B) {
    const int b = plan_bucket(obs + row * (long)stride, M);
    if (b >= 0) atomicAdd(&hist[b], 1);
  }
  __syncthreads();
  if (threadIdx.x < kPlanBuckets && hist[threadIdx.x] > 0) {
    atomicAdd(&counters[1 + threadIdx.x], hist[threadIdx.x]);
    atomicAdd(&counters[0], hist[threadIdx.x]);
  }
}

__global__ void __launch_bounds__(256) plan_scatter_kernel(const float* __restrict__ obs, int stride, int B, int M,
                                                            int32_t* __restrict__ counters, int32_t* __restrict__ row_index,
                                                            float* __restrict__ v, int32_t* __restrict__ actions) {
  __shared__ int start[kPlanBuckets];   // first plan position of bucket b (descending sequence length)
  __shared__ int hist[kPlanBuckets], base[kPlanBuckets];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = M; b >= 0; --b) { start[b] = acc; acc += counters[1 + b]; }
  }
  if (threadIdx.x < kPlanBuckets) hist[threadIdx.x] = 0;
  __syncthreads();
  const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
  int b = -1, local = 0;
  if (row < B) {
    b = plan_bucket(obs + row * (long)stride, M);
    if (b >= 0) local = atomicAdd(&hist[b], 1);
    else {  // rows without a prediction get defined outputs
      if (v) v[row] = 0.f;
      if (actions) actions[row] = 0;
    }
  }
  __syncthreads();
  if (threadIdx.x < kPlanBuckets && hist[threadIdx.x] > 0)
    base[threadIdx.x] = atomicAdd(&counters[32 + threadIdx.x], hist[threadIdx.x]);
  __syncthreads();
  if (b >= 0) row_index[start[b] + base[b] + local] = (int32_t)row;
}

}  // namespace cap

// ---- C-ABI ------------------------------------------------------------------------------------------------------------
int ca_fail_external(int code, const char* msg);  // ca_step.cu: records the message for ca_last_error

// Makes `device` current for the duration of a call and restores the caller's device afterwards (same contract as the
// entry points of ca_step.cu).
namespace {
struct PredDeviceGuard {
  int prev = -1, target = -1;
  bool ok = true;
  explicit PredDeviceGuard(int dev) : target(dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~PredDeviceGuard() {
    if (ok && prev >= 0 && prev != target) cudaSetDevice(prev);
  }
};
}  // namespace

extern "C" {

int ca_predictor_pack(const ca_predictor_params* w, void* blob, int device, void* stream) {
  if (!w || !blob) return ca_fail_external(CA_ERR_INVALID_ARG, "ca_predictor_pack: NULL argument");
  const float* const* ptrs = reinterpret_cast<const float* const*>(w);
  for (int i = 0; i < 14; ++i)
    if (!ptrs[i]) return ca_fail_external(CA_ERR_INVALID_ARG, "ca_predictor_pack: NULL parameter pointer");
  PredDeviceGuard guard(device);
  if (!guard.ok) return ca_fail_external(CA_ERR_CUDA, "cudaSetDevice failed");
  cap::PackParams q;
  q.k_lstm = w->lstm_kernel; q.b_lstm = w->lstm_bias; q.k_l1 = w->layer1_kernel; q.b_l1 = w->layer1_bias;
  q.k_l2 = w->layer2_kernel; q.b_l2 = w->layer2_bias; q.k_fc1 = w->fc1_kernel; q.b_fc1 = w->fc1_bias;
  q.k_p = w->logits_p_kernel; q.b_p = w->logits_p_bias; q.k_v = w->logits_v_kernel; q.b_v = w->logits_v_bias;
  q.avg = w->input_avg; q.std = w->input_std;
  q.blob = static_cast<unsigned char*>(blob);
  const int n = 32 * cap::kN * 8;
  cap::pack_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(q);
  if (cudaPeekAtLastError() != cudaSuccess) return ca_fail_external(CA_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
  return CA_OK;
}

static int predict_launch(const float* obs, int32_t obs_stride, int32_t batch, int32_t num_others, const void* blob, float* p,
                          float* v, int32_t* actions, int32_t greedy, float min_policy, uint64_t seed, uint64_t offset,
                          int32_t* error_flag, const int32_t* row_index, const int32_t* n_rows, int device, void* stream) {
  if (!obs || !blob || batch < 1 || num_others < 1 || obs_stride < 6 + 7 * num_others)
    return ca_fail_external(CA_ERR_INVALID_ARG, "ca_predict: bad argument");
  if ((row_index == nullptr) != (n_rows == nullptr))
    return ca_fail_external(CA_ERR_INVALID_ARG, "ca_predict_rows: row_index and n_rows go together");
  if (num_others > cap::kMaxOthers)
    return ca_fail_external(CA_ERR_UNSUPPORTED, "ca_predict: more than 22 observed other agents");
  PredDeviceGuard guard(device);
  if (!guard.ok) return ca_fail_external(CA_ERR_CUDA, "cudaSetDevice failed");
  // function attributes are per device and idempotent: set on every call (two cheap driver calls) instead of keeping a
  // process-wide table that concurrent callers would race on
  if (cudaFuncSetAttribute(cap::predict_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap::kSmTotal) != cudaSuccess)
    return ca_fail_external(CA_ERR_CUDA, "ca_predict: kernel image not usable on this device (built for sm_100a)");
  cudaFuncSetAttribute(cap::predict_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);  // two CTAs per SM
  int n_sm = 0;
  if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || n_sm < 1) n_sm = 148;
  cap::Params q;
  q.obs = obs; q.stride = obs_stride; q.B = batch; q.M = num_others;
  q.blob = static_cast<const unsigned char*>(blob);
  q.p = p; q.v = v; q.actions = actions; q.greedy = greedy; q.min_policy = min_policy;
  q.seed = seed; q.offset = offset; q.error = error_flag;
  q.row_index = row_index; q.n_rows = n_rows;
  const long n_tiles = ((long)batch + cap::kRows - 1) / cap::kRows;   // upper bound when a plan is given
  const long resident = 2l * n_sm;
  const int grid = (int)(n_tiles < resident ? n_tiles : resident);
  cap::predict_kernel<<<grid, cap::kThreads, cap::kSmTotal, (cudaStream_t)stream>>>(q);
  if (cudaPeekAtLastError() != cudaSuccess) return ca_fail_external(CA_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
  return CA_OK;
}

int ca_predict(const float* obs, int32_t obs_stride, int32_t batch, int32_t num_others, const void* blob, float* p, float* v,
               int32_t* actions, int32_t greedy, float min_policy, uint64_t seed, uint64_t offset, int32_t* error_flag,
               int device, void* stream) {
  return predict_launch(obs, obs_stride, batch, num_others, blob, p, v, actions, greedy, min_policy, seed, offset, error_flag,
                        nullptr, nullptr, device, stream);
}

int ca_predict_plan(const float* obs, int32_t obs_stride, int32_t batch, int32_t num_others, int32_t* row_index,
                    int32_t* counters, float* v, int32_t* actions, int device, void* stream) {
  if (!obs || !row_index || !counters || batch < 1 || num_others < 1 || obs_stride < 6 + 7 * num_others)
    return ca_fail_external(CA_ERR_INVALID_ARG, "ca_predict_plan: bad argument");
  if (num_others > cap::kMaxOthers)
    return ca_fail_external(CA_ERR_UNSUPPORTED, "ca_predict_plan: more than 22 observed other agents");
  PredDeviceGuard guard(device);
  if (!guard.ok) return ca_fail_external(CA_ERR_CUDA, "cudaSetDevice failed");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(counters, 0, CA_PREDICT_PLAN_COUNTERS * sizeof(int32_t), st) != cudaSuccess)
    return ca_fail_external(CA_ERR_CUDA, "ca_predict_plan: memset failed");
  const int blocks = (batch + 255) / 256;
  cap::plan_count_kernel<<<blocks, 256, 0, st>>>(obs, obs_stride, batch, num_others, counters);
  cap::plan_scatter_kernel<<<blocks, 256, 0, st>>>(obs, obs_stride, batch, num_others, counters, row_index, v, actions);
  if (cudaPeekAtLastError() != cudaSuccess) return ca_fail_external(CA_ERR_CUDA, cudaGetErrorString(cudaGetLastError()));
  return CA_OK;
}

int ca_predict_rows(const float* obs, int32_t obs_stride, int32_t batch, int32_t num_others, const void* blob,
                    const int32_t* row_index, const int32_t* n_rows, float* p, float* v, int32_t* actions, int32_t greedy,
                    float min_policy, uint64_t seed, uint64_t offset, int32_t* error_flag, int device, void* stream) {
  if (!row_index || !n_rows) return ca_fail_external(CA_ERR_INVALID_ARG, "ca_predict_rows: NULL plan");
  return predict_launch(obs, obs_stride, batch, num_others, blob, p, v, actions, greedy, min_policy, seed, offset, error_flag,
                        row_index, n_rows, device, stream);
}

}  // extern "C"
