// Grid-barrier latency probe: 148 CTAs x 256 threads, N back-to-back barriers of several designs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=true scripts/barrier_probe.cu -o scripts/_bin/barrier_probe
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void red_release(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

template <int MODE>
__global__ void probe(unsigned* bar, int n, long long* out) {
  cg::grid_group grid = cg::this_grid();
  unsigned epoch = 0;
  const unsigned G = gridDim.x;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (MODE == 0) {
      grid.sync();
    } else if (MODE == 1) {            // flat: one counter, atomicAdd + acquire spin, fences
      __syncthreads();
      if (threadIdx.x == 0) {
        ++epoch; __threadfence(); atomicAdd(bar, 1u);
        { unsigned sp = 0; while ((int)(ld_acquire(bar) - epoch * G) < 0) { if (++sp > 50000000u) __trap(); } }
        __threadfence();
      }
      __syncthreads();
    } else if (MODE == 2) {            // hierarchical 16 groups
      __syncthreads();
      if (threadIdx.x == 0) {
        ++epoch;
        const unsigned ng = 16, grp = blockIdx.x % ng, gs = (G - grp + ng - 1) / ng;
        __threadfence();
        unsigned old = atomicAdd(bar + 32 * (1 + grp), 1u);
        if (old + 1 == epoch * gs) atomicAdd(bar, 1u);
        { unsigned sp = 0; while ((int)(ld_acquire(bar) - epoch * ng) < 0) { if (++sp > 50000000u) __trap(); } }
        __threadfence();
      }
      __syncthreads();
    } else if (MODE == 3) {            // flat, red.release (no return value) + relaxed polling + one fence
      __syncthreads();
      if (threadIdx.x == 0) {
        ++epoch; red_release(bar);
        { unsigned sp = 0; while ((int)(ld_relaxed(bar) - epoch * G) < 0) { if (++sp > 50000000u) __trap(); } }
        __threadfence();
      }
      __syncthreads();
    } else if (MODE == 4) {            // per-CTA flags: every CTA writes its own slot, CTA 0's warp gathers, then releases
      __syncthreads();
      ++epoch;
      if (blockIdx.x == 0) {
        if (threadIdx.x >= 1 && threadIdx.x < G) { unsigned sp = 0; while ((int)(ld_acquire(bar + 32 + threadIdx.x) - epoch) < 0) { if (++sp > 50000000u) __trap(); } }
        __syncthreads();
        if (threadIdx.x == 0) { __threadfence(); asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar), "r"(epoch) : "memory"); }
      } else if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 32 + blockIdx.x), "r"(epoch) : "memory");
        { unsigned sp = 0; while ((int)(ld_acquire(bar) - epoch) < 0) { if (++sp > 50000000u) __trap(); } }
      }
      __syncthreads();
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int MODE>
void run(const char* name, unsigned* bar, long long* out, int n) {
  cudaMemset(bar, 0, 64 * 1024);
  void* args[] = {&bar, &n, &out};
  int dev = 0, sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(bar, 0, 64 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)probe<MODE>, dim3(sms), dim3(256), args, 0, 0);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long cyc = 0; cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
    if (rep == 1) { printf("%-40s %s  %.3f us/barrier  (%lld cycles/barrier)\n", name, cudaGetErrorString(err), ms * 1e3 / n, cyc / n); fflush(stdout); }
  }
}

int main() {
  unsigned* bar; long long* out;
  cudaMalloc(&bar, 64 * 1024); cudaMalloc(&out, 64);
  const int n = 2000;
  run<0>("cooperative_groups grid.sync", bar, out, n);
  run<1>("flat atomicAdd + acquire spin", bar, out, n);
  run<2>("hierarchical 16 groups", bar, out, n);
  run<3>("flat red.release + relaxed poll", bar, out, n);
  run<4>("per-CTA flags gathered by CTA 0", bar, out, n);
  return 0;
}
