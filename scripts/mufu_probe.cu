// Throughput probes on sm_100a: the special-function unit (tanh.approx.f32 vs ex2.approx, per SM and clock) and tensor-memory
// reads (tcgen05.ld).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mufu_probe scripts/mufu_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(float* out, int iters) {
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// tcgen05.ld throughput: every warp of the CTA reads its 32-lane window of a 128 x 512 fp32 TMEM allocation, `iters` times
// kLoads back-to-back loads of shape 32x32b.x{8,16,32} (1 / 2 / 4 KB per warp and instruction) followed by one wait::ld.
template <int X>
__device__ __forceinline__ float ld_x(unsigned addr);
template <>
__device__ __forceinline__ float ld_x<8>(unsigned addr) {
  unsigned r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(addr) : "memory");
  return __uint_as_float(r[0] ^ r[7]);
}
template <>
__device__ __forceinline__ float ld_x<16>(unsigned addr) {
  unsigned r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr) : "memory");
  return __uint_as_float(r[0] ^ r[15]);
}
template <>
__device__ __forceinline__ float ld_x<32>(unsigned addr) {
  unsigned r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(addr) : "memory");
  return __uint_as_float(r[0] ^ r[31]);
}

template <int X, int kLoads>
__global__ void tmem_read(float* out, int iters) {
  __shared__ unsigned slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned base = slot + ((unsigned)((warp & 3) * 32) << 16);
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float v[kLoads];
#pragma unroll
    for (int k = 0; k < kLoads; ++k) v[k] = ld_x<X>(base + (((it * kLoads + k) * X) & 511));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int k = 0; k < kLoads; ++k) acc += v[k];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}

template <int X, int kLoads>
void run_tmem(int warps, int sms, float mhz) {
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  const int iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  tmem_read<X, kLoads><<<sms, warps * 32>>>(out, 64);
  cudaEventRecord(e0);
  tmem_read<X, kLoads><<<sms, warps * 32>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = (double)sms * warps * iters * kLoads * X * 128.0;
  printf("tcgen05.ld 32x32b.x%d, %d loads per wait, %2d warps per SM: %.3f ms = %.1f bytes per clock per SM at %.0f MHz (%s)\n", X,
         kLoads, warps, ms, bytes / (ms * 1e-3) / (mhz * 1e6) / sms, mhz, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

template <int OP>
void run(const char* name, int sms, float mhz) {
  float* out;
  cudaMalloc(&out, sizeof(float) * sms * 8 * 1024);
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<OP><<<sms * 2, 1024>>>(out, 16);
  cudaEventRecord(e0);
  k<OP><<<sms * 2, 1024>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)sms * 2 * 1024 * iters * 8;
  printf("%-14s %.3f ms  %.1f Gop/s  = %.2f lane-ops per clock per SM at %.0f MHz\n", name, ms, ops / ms / 1e6,
         ops / (ms * 1e-3) / (mhz * 1e6) / sms, mhz);
  cudaFree(out);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const float mhz = khz / 1000.f;
  printf("%s, %d SMs, max clock %.0f MHz\n", p.name, p.multiProcessorCount, mhz);
  run<0>("tanh.approx", p.multiProcessorCount, mhz);
  run<1>("ex2.approx", p.multiProcessorCount, mhz);
  run<2>("rcp.approx", p.multiProcessorCount, mhz);
  run<3>("fma.rn.f32", p.multiProcessorCount, mhz);
  const int n = p.multiProcessorCount;
  run_tmem<32, 1>(4, n, mhz); run_tmem<32, 1>(8, n, mhz); run_tmem<32, 1>(16, n, mhz);
  run_tmem<32, 4>(4, n, mhz); run_tmem<32, 4>(8, n, mhz); run_tmem<32, 4>(16, n, mhz);
  run_tmem<16, 4>(8, n, mhz); run_tmem<8, 4>(8, n, mhz); run_tmem<8, 4>(16, n, mhz);
  return 0;
}
