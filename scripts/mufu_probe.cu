// Throughput of the special-function unit on sm_100a: tanh.approx.f32 vs ex2.approx / rcp.approx (per SM and clock).
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
  return 0;
}
