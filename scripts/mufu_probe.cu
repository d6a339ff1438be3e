// Micro-benchmark: MUFU throughput of tanh.approx.f32 vs tanh.approx.f16x2 vs ex2.approx (f32, f16x2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/mufu_probe.cu -o scripts/_bin/mufu_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(uint32_t* out, int iters) {
  uint32_t x0 = threadIdx.x * 0x00010001u + 0x30003000u, x1 = x0 + 0x01010101u, x2 = x0 ^ 0x00110011u, x3 = x0 + 77u;
  uint32_t x4 = x0 + 3u, x5 = x1 + 5u, x6 = x2 + 7u, x7 = x3 + 9u;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#define STEP(x)                                                                                   \
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+r"(x));                                 \
  else if (OP == 1) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(x));                          \
  else if (OP == 2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(x));                         \
  else if (OP == 3) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x));                           \
  else if (OP == 4) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(x));
#pragma unroll
    for (int r = 0; r < 8; ++r) { STEP(x0) STEP(x1) STEP(x2) STEP(x3) STEP(x4) STEP(x5) STEP(x6) STEP(x7) }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

template <int OP>
void run(const char* name, uint32_t* d) {
  const int iters = 4096, blocks = 148 * 4, threads = 512;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<blocks, threads>>>(d, 16);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(d, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double ops = double(blocks) * threads * iters * 64;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-22s %8.3f ms  %7.2f Ginstr/s  = %.2f thread-instr/clk/SM at %d MHz nominal\n", name, ms, ops / ms * 1e-6,
         ops / (ms * 1e-3) / 148 / (clk * 1e3), clk / 1000);
}

int main() {
  uint32_t* d; cudaMalloc(&d, 148 * 4 * 512 * 4);
  run<0>("tanh.approx.f32", d);
  run<1>("tanh.approx.f16x2", d);
  run<4>("tanh.approx.bf16x2", d);
  run<2>("ex2.approx.ftz.f32", d);
  run<3>("ex2.approx.f16x2", d);
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
