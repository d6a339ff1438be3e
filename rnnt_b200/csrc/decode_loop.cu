// Whole batched greedy decode as ONE persistent cooperative kernel (fp32 throughout).
//
// Replaces the host-driven loop of rnnt/model.py:90-128 (`_greedy_decode_conv`): per step
//   joint.single_forward(audio[:, t], pred[:, -1])  (joint.py:44-55)  -> argmax -> blank / 10-emit rule -> append
//   -> ConvPredictor re-run (predictor.py:209-229, model.py:119-123)
// for a whole batch, with every piece of loop state on the device and no host round trip.  The predictor is
// evaluated incrementally: it is causal with a 7-token receptive field, so per utterance the last 2 layer-normed
// embeddings (conv1, k=3) and the last 4 conv1 outputs (conv2, k=5) are kept and one new position is computed per
// emitted token; zero-initialised state is the left zero padding of rnnt/causalconv.py:29.
//
// One CTA per SM, 256 threads, phases separated by grid-wide barriers:
//   P1  feats = LayerNorm(lin) for utterances that just emitted; h = tanh(enc[b, t_b] + feats[b])
//   P2  logits[b, v] = W_j[v,:] . h[b,:] + b_j[v]        (CTA per 8 classes, K split over the threads, see batched_gemv)
//   P3  argmax (lowest index on ties) + top-2 margin, blank / max-per-frame rule, token append, x = LN(emb[tok])
//   P4  y = gelu(conv1 tap-GEMV)   P5  z = gelu(conv2 tap-GEMV), shift conv1 state   P6  lin = linear(z), shift conv2 state
// P4-P6 only run in steps where some utterance emitted.  Every dot product is accumulated in a fixed order (thread-
// strided partial sums, butterfly, then warps 0..7), so results are deterministic.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace rb {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < kThreads / 32) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
    t = warp_sum(t);
    if (l == 0) red[0] = t;
  }
  __syncthreads();
  return red[0];
}

// dst[i] = LayerNorm(src)[i] * w[i] + b[i] over n elements, by the whole CTA (torch: biased variance, eps 1e-5)
__device__ void block_layer_norm(const float* __restrict__ src, const float* __restrict__ w,
                                 const float* __restrict__ b, float* __restrict__ dst, int n, float* red) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += kThreads) s += src[i];
  const float mean = block_sum(s, red) / n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n; i += kThreads) { const float d = src[i] - mean; q += d * d; }
  const float var = block_sum(q, red) / n;
  const float rstd = rsqrtf(var + 1e-5f);
  for (int i = threadIdx.x; i < n; i += kThreads) dst[i] = (src[i] - mean) * rstd * w[i] + b[i];
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// out[r, o] = act(bias[o] + W[o, :] . in_r) for the rows r listed in `rows` (nrows of them), o in [0, nout).
// in_r is the concatenation of up to two row segments; W is (nout, klen) row-major, klen a multiple of 4.
//
// A CTA task = OUT consecutive outputs.  The 256 threads split K (thread t takes float4 t, t+256, ...), so every
// activation element is read from shared memory ONCE per CTA (not once per warp); each thread keeps OUT x kRows
// partial sums, a warp transpose-reduce leaves lane l with the warp total of partial l, and one pass over
// partial[warp][l] in shared memory finishes the sum.  Activation rows are staged kRows at a time with cp.async into
// a double buffer so the next chunk streams in from L2 while the current one is being multiplied.
struct Seg { const float* base; long long row_stride; int len; };

constexpr int kRows = 8;
constexpr int kMaxIt = 3;   // float4 K-steps per thread: supports K <= 4 * 3 * 256 = 3072

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_pending(int n) {   // at most n most-recent groups still in flight
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}

// v[32] per lane -> lane l returns the sum over the warp's lanes of v[l] (31 shuffles instead of 32 x 5).
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int j = 0; j < step; ++j) {
      const float send = up ? v[j] : v[j + step];
      const float keep = up ? v[j + step] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return v[0];
}

template <bool GELU, int OUT>
__device__ void batched_gemv(const float* __restrict__ W, const float* __restrict__ bias, int nout, int klen,
                             const Seg* segs, int nsegs, const int* __restrict__ rows, int nrows,
                             float* __restrict__ out, long long out_stride, float* smem, int smem_floats,
                             float* partial, int rsplit = 1) {
  static_assert(OUT * kRows == 32 || OUT * kRows == 64, "partial sums per thread must be 32 or 64");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k4 = klen >> 2;
  // a task = OUT outputs x one of `rsplit` ranges of row chunks (more tasks than SMs would cost a second wave, fewer
  // leave SMs idle: the callers pick OUT and rsplit so that ntasks is just under the grid size)
  const int ntasks = ((nout + OUT - 1) / OUT) * rsplit;
  const int nchunks_all = (nrows + kRows - 1) / kRows;
  const int chunks_per_part = (nchunks_all + rsplit - 1) / rsplit;
  const int depth = max(2, min(4, smem_floats / (kRows * klen)));   // staging ring: chunks c .. c+depth-1 in flight

  auto stage = [&](int chunk, float* dst) {
    const int r0 = chunk * kRows, nr = min(kRows, nrows - r0);
    for (int rr = 0; rr < nr; ++rr) {
      const int r = rows[r0 + rr];
      int off = 0;
      for (int sg = 0; sg < nsegs; ++sg) {
        const float* src = segs[sg].base + r * segs[sg].row_stride;
        for (int i = tid; i < (segs[sg].len >> 2); i += kThreads) cp_async16(dst + rr * klen + off + 4 * i, src + 4 * i);
        off += segs[sg].len;
      }
    }
    cp_async_commit();
  };

  for (int task = blockIdx.x; task < ntasks; task += gridDim.x) {
    const int o0 = (task / rsplit) * OUT;
    const int c_lo = (task % rsplit) * chunks_per_part;
    const int nchunks = min(nchunks_all, c_lo + chunks_per_part) - c_lo;   // chunks c_lo .. c_lo + nchunks - 1
    if (nchunks <= 0) continue;         // uniform over the CTA
    const float4* w4[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) w4[o] = reinterpret_cast<const float4*>(W + static_cast<long long>(min(o0 + o, nout - 1)) * klen);
    __syncthreads();                    // previous users of the staging buffers are done
    for (int c = 0; c < depth - 1; ++c) {
      if (c < nchunks) stage(c_lo + c, smem + (c % depth) * kRows * klen); else cp_async_commit();
    }
    // this thread's slice of the OUT weight rows stays in registers for the whole task (K <= 3072)
    float4 wreg[kMaxIt][OUT];
#pragma unroll
    for (int it = 0; it < kMaxIt; ++it) {
      const int i = tid + it * kThreads;
#pragma unroll
      for (int o = 0; o < OUT; ++o) wreg[it][o] = (i < k4) ? __ldg(w4[o] + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int c = 0; c < nchunks; ++c) {
      const int r0 = (c_lo + c) * kRows, nr = min(kRows, nrows - r0);
      const int ahead = c + depth - 1;    // one commit per iteration keeps the group count uniform
      if (ahead < nchunks) stage(c_lo + ahead, smem + (ahead % depth) * kRows * klen); else cp_async_commit();
      cp_async_wait_pending(depth - 1);
      __syncthreads();                  // chunk c has landed for every thread
      const float4* x4 = reinterpret_cast<const float4*>(smem + (c % depth) * kRows * klen);
      float acc[OUT * kRows];
#pragma unroll
      for (int q = 0; q < OUT * kRows; ++q) acc[q] = 0.f;
#pragma unroll
      for (int it = 0; it < kMaxIt; ++it) {
        const int i = tid + it * kThreads;
        if (i < k4) {
#pragma unroll
          for (int rr = 0; rr < kRows; ++rr) {
            if (rr < nr) {
              const float4 x = x4[rr * k4 + i];
#pragma unroll
              for (int o = 0; o < OUT; ++o) {
                const float4 w = wreg[it][o];
                float a = acc[o * kRows + rr];
                a = fmaf(w.x, x.x, a); a = fmaf(w.y, x.y, a); a = fmaf(w.z, x.z, a); a = fmaf(w.w, x.w, a);
                acc[o * kRows + rr] = a;
              }
            }
          }
        }
      }
      // warp totals: lane l <- total of partial l (and l + 32 when there are 64)
      float t0, t1 = 0.f;
      if (OUT * kRows == 64) {
        float lo[32], hi[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) { lo[q] = acc[q]; hi[q] = acc[q + 32]; }
        t0 = warp_transpose_reduce(lo);
        t1 = warp_transpose_reduce(hi);
      } else {
        float lo[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) lo[q] = acc[q];
        t0 = warp_transpose_reduce(lo);
      }
      partial[warp * 64 + lane] = t0;
      if (OUT * kRows == 64) partial[warp * 64 + 32 + lane] = t1;
      __syncthreads();
      if (tid < OUT * kRows) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += partial[w * 64 + tid];
        const int o = tid / kRows, rr = tid % kRows;
        if (rr < nr && o0 + o < nout) {
          v += bias ? __ldg(bias + o0 + o) : 0.f;
          if (GELU) v = gelu_erf(v);
          out[rows[r0 + rr] * out_stride + o0 + o] = v;
        }
      }
      // the next iteration's first __syncthreads (after its cp.async wait) orders these reads before partial is rewritten
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) greedy_decode_kernel(DecodeArgs p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float smem[];
  __shared__ float red[kThreads / 32];
  __shared__ float partial[(kThreads / 32) * 64];
  __shared__ float s_best[kThreads];
  __shared__ float s_second[kThreads];
  __shared__ int s_idx[kThreads];
  const int B = p.B, H = p.H, V = p.V, E = p.E;
  const int tid = threadIdx.x;
  int* rows_act = p.rows;          // compacted list of active utterances (P2) ...
  int* rows_emit = p.rows + B;     // ... and of utterances that emitted in this step (P4-P6)
  int* counts = p.flags;           // [0] n_active, [1] n_emit, [2] n_active_next
  const bool timer = (blockIdx.x == 0 && tid == 0 && p.prof != nullptr);
  long long tmark = timer ? clock64() : 0;
  auto lap = [&](int slot) { if (timer) { const long long now = clock64(); p.prof[slot] += now - tmark; tmark = now; } };
  if (timer) for (int i = 0; i < 8; ++i) p.prof[i] = 0;

  // ---- init: state, seed token = blank for every utterance, everyone "emits" the seed so the predictor runs once
  for (int i = blockIdx.x * kThreads + tid; i < B * 2 * E; i += gridDim.x * kThreads) p.xs[i] = 0.f;
  for (int i = blockIdx.x * kThreads + tid; i < B * 4 * E; i += gridDim.x * kThreads) p.ys[i] = 0.f;
  if (blockIdx.x == 0) {
    for (int b = tid; b < B; b += kThreads) {
      p.t_idx[b] = 0; p.per[b] = 0; p.ntok[b] = 1; p.emit[b] = 1; rows_emit[b] = b;
    }
    if (tid == 0) { counts[0] = 0; counts[1] = B; counts[2] = 0; }
  }
  for (int b = blockIdx.x; b < B; b += gridDim.x)
    block_layer_norm(p.emb + static_cast<long long>(p.blank) * E, p.ln1_w, p.ln1_b, p.xnew + b * E, E, red);
  grid.sync();

  for (int step = 0;; ++step) {
    const int n_emit = counts[1];
    if (n_emit > 0) {
      // ---- P4: conv1 at the new position: [xs(b,0), xs(b,1), xnew(b)] . w1^T
      {
        Seg segs[2] = {{p.xs, 2LL * E, 2 * E}, {p.xnew, E, E}};
        batched_gemv<true, 8>(p.w1, p.b1, E, 3 * E, segs, 2, rows_emit, n_emit, p.ynew, E, smem, p.smem_floats, partial, p.rsplit_e);
      }
      lap(3);
      grid.sync();
      lap(7);
      // ---- P5: conv2 at the new position: [ys(b,0..3), ynew(b)] . w2^T ; conv1 state shifts (xs is no longer read)
      {
        Seg segs[2] = {{p.ys, 4LL * E, 4 * E}, {p.ynew, E, E}};
        batched_gemv<true, 8>(p.w2, p.b2, E, 5 * E, segs, 2, rows_emit, n_emit, p.z, E, smem, p.smem_floats, partial, p.rsplit_e);
      }
      for (int j = blockIdx.x; j < n_emit; j += gridDim.x) {
        const int b = rows_emit[j];
        for (int i = tid; i < E; i += kThreads) {
          p.xs[(b * 2 + 0) * E + i] = p.xs[(b * 2 + 1) * E + i];
          p.xs[(b * 2 + 1) * E + i] = p.xnew[b * E + i];
        }
      }
      lap(4);
      grid.sync();
      lap(7);
      // ---- P6: linear ; conv2 state shifts (ys is no longer read)
      {
        Seg segs[1] = {{p.z, E, E}};
        batched_gemv<false, 8>(p.wl, p.bl, H, E, segs, 1, rows_emit, n_emit, p.lin, H, smem, p.smem_floats, partial);
      }
      for (int j = blockIdx.x; j < n_emit; j += gridDim.x) {
        const int b = rows_emit[j];
        for (int i = tid; i < E; i += kThreads) {
          const float y0 = p.ys[(b * 4 + 1) * E + i], y1 = p.ys[(b * 4 + 2) * E + i], y2 = p.ys[(b * 4 + 3) * E + i];
          p.ys[(b * 4 + 0) * E + i] = y0;
          p.ys[(b * 4 + 1) * E + i] = y1;
          p.ys[(b * 4 + 2) * E + i] = y2;
          p.ys[(b * 4 + 3) * E + i] = p.ynew[b * E + i];
        }
      }
      lap(5);
      grid.sync();
      lap(7);
    }
    if (step >= p.max_steps) break;

    // ---- P1: refresh predictor features where a token was emitted; joint hidden rows for active utterances
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
      if (p.emit[b]) block_layer_norm(p.lin + static_cast<long long>(b) * H, p.ln2_w, p.ln2_b, p.feats + static_cast<long long>(b) * H, H, red);
      __syncthreads();
      const int t = p.t_idx[b];
      const bool active = t < p.T_len[b] && p.ntok[b] < p.max_len;
      if (active) {
        const float* e = p.enc + b * p.enc_sb + static_cast<long long>(t) * p.enc_st;
        for (int i = tid; i < H; i += kThreads) p.hbuf[static_cast<long long>(b) * H + i] = tanhf(e[i] + p.feats[static_cast<long long>(b) * H + i]);
      }
    }
    if (blockIdx.x == 0) {
      // compact the active list (single warp, order preserved)
      if (tid < 32) {
        int n = 0;
        for (int b0 = 0; b0 < B; b0 += 32) {
          const int b = b0 + tid;
          const bool a = b < B && p.t_idx[b] < p.T_len[b] && p.ntok[b] < p.max_len;
          const unsigned m = __ballot_sync(0xffffffffu, a);
          if (a) rows_act[n + __popc(m & ((1u << tid) - 1))] = b;
          n += __popc(m);
        }
        if (tid == 0) { counts[0] = n; counts[1] = 0; counts[2] = 0; }
      }
    }
    lap(0);
    grid.sync();
    lap(7);
    const int n_act = counts[0];
    if (n_act == 0) break;

    // ---- P2: joint logits for the active rows
    {
      Seg segs[1] = {{p.hbuf, H, H}};
      batched_gemv<false, 8>(p.Wj, p.bj, V, H, segs, 1, rows_act, n_act, p.logits, V, smem, p.smem_floats, partial);
    }
    lap(1);
    grid.sync();
    lap(7);

    // ---- P3: argmax + decode rule + embedding layer norm, one CTA per active utterance
    for (int j = blockIdx.x; j < n_act; j += gridDim.x) {
      const int b = rows_act[j];
      const float* row = p.logits + static_cast<long long>(b) * V;
      float best = -INFINITY, second = -INFINITY;
      int idx = 0x7fffffff;
      for (int v = tid; v < V; v += kThreads) {
        const float x = row[v];
        if (x > best) { second = best; best = x; idx = v; }
        else if (x > second) second = x;
      }
      __syncthreads();
      s_best[tid] = best; s_second[tid] = second; s_idx[tid] = idx;
      __syncthreads();
      for (int o = kThreads >> 1; o; o >>= 1) {
        if (tid < o) {
          const float b2 = s_best[tid + o], s2 = s_second[tid + o];
          const int i2 = s_idx[tid + o];
          float b1 = s_best[tid], s1 = s_second[tid];
          int i1 = s_idx[tid];
          if (b2 > b1 || (b2 == b1 && i2 < i1)) { s1 = fmaxf(b1, s2); b1 = b2; i1 = i2; }
          else s1 = fmaxf(s1, b2);
          s_best[tid] = b1; s_second[tid] = s1; s_idx[tid] = i1;
        }
        __syncthreads();
      }
      // torch.argmax always returns an in-range index; a row without any finite maximum (all NaN / -inf) maps to 0
      const int tok = static_cast<unsigned>(s_idx[0]) < static_cast<unsigned>(V) ? s_idx[0] : 0;
      const bool advance = (tok == p.blank) || (p.per[b] >= p.max_per_frame);
      __syncthreads();
      if (tid == 0) {
        if (p.margins) p.margins[static_cast<long long>(step) * B + b] = s_best[0] - s_second[0];
        if (advance) {
          p.t_idx[b] += 1; p.per[b] = 0; p.emit[b] = 0;
        } else {
          p.tokens[static_cast<long long>(b) * p.max_len + p.ntok[b] - 1] = tok;
          p.ntok[b] += 1; p.per[b] += 1; p.emit[b] = 1;
          rows_emit[atomicAdd(&counts[1], 1)] = b;
        }
      }
      if (!advance) block_layer_norm(p.emb + static_cast<long long>(tok) * E, p.ln1_w, p.ln1_b, p.xnew + b * E, E, red);
      __syncthreads();
    }
    // (a finished utterance may keep emit = 1; P1 then recomputes the same feats from an unchanged lin -- harmless)
    lap(2);
    grid.sync();
    lap(7);
  }
}

}  // namespace

// Scratch layout: [8 x int64 phase timers | float regions, each padded to a multiple of 4 floats so that every region
// starts 16-byte aligned whatever B, V, E are (cp.async / float4 need it) | int regions].
namespace {
inline size_t pad4(size_t n) { return (n + 3) & ~static_cast<size_t>(3); }
constexpr size_t kTimerBytes = 64;
}  // namespace

size_t greedy_decode_scratch_bytes(int B, int H, int V, int E) {
  const size_t b = static_cast<size_t>(B);
  const size_t floats = 3 * pad4(b * H) /*feats, hbuf, lin*/ + pad4(b * V) /*logits*/ + pad4(b * 2 * E) + pad4(b * 4 * E) +
                        3 * pad4(b * E) /*xnew, ynew, z*/;
  const size_t ints = b * 5 + 16;
  return kTimerBytes + floats * 4 + ints * 4 + 256;
}

int launch_greedy_decode(DecodeArgs a, float* scratch, cudaStream_t stream) {
  ProfScope prof_(kProfOther, stream);
  RB_REQUIRE(a.B > 0 && a.H > 0 && a.V > 1 && a.E > 0 && a.max_len >= 1, -1, "invalid decode shape");
  const int klen_max = std::max(std::max(a.H, 5 * a.E), 3 * a.E);
  const size_t smem_min = 2 * static_cast<size_t>(kRows) * klen_max * sizeof(float);   // at least double-buffered staging
  const size_t smem = std::max<size_t>(smem_min, 192 * 1024);   // deeper ring (up to 4 chunks) for the narrower phases
  a.smem_floats = static_cast<int>(smem / sizeof(float));
  // conv phases: E/8 output blocks x rsplit row ranges should just fill the grid (E=512: 64 x 2 = 128 tasks on 148 SMs)
  a.rsplit_e = std::max(1, std::min(4, device_sm_count() / std::max(1, (a.E + 7) / 8)));
  RB_REQUIRE(smem <= 200 * 1024 && klen_max <= 4 * kMaxIt * kThreads, -6,
             "decode kernel supports hidden_features <= 3072 and embedding dim <= 614");
  // carve the scratch
  RB_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, -3, "decode scratch must be 16-byte aligned");
  a.prof = reinterpret_cast<long long*>(scratch);          // 8 x int64 at the (aligned) start
  float* f = scratch + kTimerBytes / sizeof(float);
  const int B = a.B, H = a.H, V = a.V, E = a.E;
  const size_t b = static_cast<size_t>(B);
  a.feats = f; f += pad4(b * H);
  a.hbuf = f; f += pad4(b * H);
  a.logits = f; f += pad4(b * V);
  a.xs = f; f += pad4(b * 2 * E);
  a.ys = f; f += pad4(b * 4 * E);
  a.xnew = f; f += pad4(b * E);
  a.ynew = f; f += pad4(b * E);
  a.z = f; f += pad4(b * E);
  a.lin = f; f += pad4(b * H);
  int* ip = reinterpret_cast<int*>(f);
  a.t_idx = ip; ip += B;
  a.per = ip; ip += B;
  a.emit = ip; ip += B;
  a.rows = ip; ip += 2 * B;
  a.flags = ip; ip += 16;
  RB_CUDA_CHECK(cudaFuncSetAttribute(greedy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  RB_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_decode_kernel, kThreads, smem));
  RB_REQUIRE(per_sm >= 1, -6, "decode kernel does not fit on an SM");
  const int grid = device_sm_count();
  void* args[] = {&a};
  RB_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(greedy_decode_kernel), dim3(grid), dim3(kThreads),
                                            args, smem, stream));
  return 0;
}

}  // namespace rb
