// Greedy-decode joint step in fp32: tokens[n] = argmax_v ( W[v,:] . tanh(a_n + p_n) + bias[v] ).
//
// Replaces joint.single_forward + argmax at rnnt/model.py:66-69 / :110-113 for a whole batch of utterances
// per step.  fp32 FFMA throughout (accurate tanhf, no tensor cores) so the token stream can be bit-exact with
// the fp32 reference; ties resolve to the lowest index like torch.argmax.  Three tiny launches per step:
// hidden rows, logits (each warp keeps one W row in registers and sweeps the N rows), per-row argmax + top-2 gap.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace rb {
namespace {

__global__ void decode_hidden_kernel(const float* __restrict__ a, long long a_stride, const float* __restrict__ q,
                                     long long q_stride, int N, int H, float* __restrict__ hbuf) {
  const long long n = static_cast<long long>(N) * H;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / H), k = static_cast<int>(i % H);
    hbuf[i] = tanhf(a[r * a_stride + k] + q[r * q_stride + k]);
  }
}

constexpr int kMaxKPerLane = 64;   // H <= 2048

__global__ void decode_logits_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                     const float* __restrict__ hbuf, int N, int H, int V,
                                     float* __restrict__ logits) {
  const int lane = threadIdx.x & 31;
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= V) return;
  float w[kMaxKPerLane];
  const int per = (H + 31) / 32;
#pragma unroll
  for (int i = 0; i < kMaxKPerLane; ++i) {
    const int k = i * 32 + lane;
    w[i] = (i < per && k < H) ? __ldg(W + static_cast<long long>(v) * H + k) : 0.f;
  }
  const float bv = __ldg(bias + v);
  for (int r = 0; r < N; ++r) {
    const float* h = hbuf + static_cast<long long>(r) * H;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxKPerLane; ++i) {
      const int k = i * 32 + lane;
      if (i < per && k < H) acc = fmaf(w[i], h[k], acc);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) logits[static_cast<long long>(r) * V + v] = acc + bv;
  }
}

__global__ void decode_argmax_kernel(const float* __restrict__ logits, int N, int V, int* __restrict__ tokens,
                                     float* __restrict__ top2) {
  const int r = blockIdx.x;
  if (r >= N) return;
  const float* row = logits + static_cast<long long>(r) * V;
  float best = -INFINITY, second = -INFINITY;
  int idx = 0x7fffffff;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float x = row[v];
    if (x > best) { second = best; best = x; idx = v; }
    else if (x > second) second = x;
  }
  __shared__ float sb[256], ss[256];
  __shared__ int si[256];
  sb[threadIdx.x] = best; ss[threadIdx.x] = second; si[threadIdx.x] = idx;
  __syncthreads();
  for (int o = blockDim.x >> 1; o; o >>= 1) {
    if (threadIdx.x < o) {
      const float b2 = sb[threadIdx.x + o], s2 = ss[threadIdx.x + o];
      const int i2 = si[threadIdx.x + o];
      float b1 = sb[threadIdx.x], s1 = ss[threadIdx.x];
      int i1 = si[threadIdx.x];
      if (b2 > b1 || (b2 == b1 && i2 < i1)) {
        s1 = fmaxf(b1, s2);
        b1 = b2; i1 = i2;
      } else {
        s1 = fmaxf(s1, b2);
      }
      sb[threadIdx.x] = b1; ss[threadIdx.x] = s1; si[threadIdx.x] = i1;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tokens[r] = static_cast<unsigned>(si[0]) < static_cast<unsigned>(V) ? si[0] : 0;   // all-NaN / -inf row -> 0
    if (top2) top2[r] = sb[0] - ss[0];
  }
}

}  // namespace

size_t joint_argmax_scratch_bytes(int N, int V) {
  return static_cast<size_t>(N) * 2048 * sizeof(float) + static_cast<size_t>(N) * V * sizeof(float);
}

int launch_joint_argmax(const float* enc_rows, long long enc_stride, const float* pred_rows, long long pred_stride,
                        const float* W, const float* bias, int N, int H, int V, int* tokens, float* top2,
                        float* scratch, cudaStream_t stream) {
  ProfScope prof_(kProfOther, stream);
  RB_REQUIRE(H <= 32 * kMaxKPerLane, -6, "decode kernel supports hidden_features <= %d (got %d)", 32 * kMaxKPerLane, H);
  if (N <= 0) return 0;
  float* hbuf = scratch;
  float* logits = scratch + static_cast<size_t>(N) * 2048;
  const long long n = static_cast<long long>(N) * H;
  decode_hidden_kernel<<<static_cast<int>(std::min<long long>((n + 255) / 256, 1184)), 256, 0, stream>>>(
      enc_rows, enc_stride, pred_rows, pred_stride, N, H, hbuf);
  decode_logits_kernel<<<(V + 7) / 8, 256, 0, stream>>>(W, bias, hbuf, N, H, V, logits);
  decode_argmax_kernel<<<N, 256, 0, stream>>>(logits, N, V, tokens, top2);
  RB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace rb
