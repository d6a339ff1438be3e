// How many thread-block clusters of size 2 / 4 / 8 with ~200 KB of shared memory per CTA (one CTA per SM) can be
// co-resident on this GPU?  (Decides whether wider clusters -- TMA multicast of shared operands -- are usable without
// leaving SMs idle.)   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/cluster_probe.cu -o scripts/_bin/cluster_probe
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((sms / cs) * cs); cfg.blockDim = dim3(320); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d of %d SMs busy (%s)\n", cs, n, n * cs, sms, cudaGetErrorString(e));
  }
  return 0;
}
