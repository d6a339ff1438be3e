// Backward contraction #2:  dh = g . W  (K = V), fused with the tanh backward and the broadcast reductions.
//
// Replaces autograd through rnnt/joint.py:32-39 for the activations:
//   dz[c,k]       = (sum_v g[c,v] W[v,k]) * (1 - h[c,k]^2)
//   d_enc[b,t,k]  = sum_u dz[b,t,u,k]      d_pred[b,u,k] = sum_t dz[b,t,u,k]
// A operand: gradient ring g [ring_rows, Vp] bf16 (K-major, TMA).  B operand: W^T read as an MN-major
// view of the same bf16 W[Vp, Hp] buffer the forward uses (no transposed copy).
// Tile = 128 lattice cells (16 t x 8 u) x 512 hidden columns (two 256-column TMEM accumulators).
// Epilogue: TMEM -> dz -> padded smem transpose -> per-(t,k) / per-(u,k) partial sums -> coalesced fp32 atomics.
#include "common.cuh"
#include "kernels.h"

namespace rb {
namespace {

constexpr int kStagesA = 3;
constexpr int kStagesB = 4;
constexpr int kBytesA = kTileM * kBK * 2;   // 16 KB (128 cells x 64 v)
constexpr int kBytesB = kBK * kBN * 2;      // 32 KB (64 v x 256 k), four 64x64 MN-major boxes
constexpr int kNumThreads = 192;
constexpr int kTmemCols = 512;
constexpr int kTbStride = 33;

struct SmemLayout {
  static constexpr int b_ring = 0;
  static constexpr int a_ring = b_ring + kStagesB * kBytesB;
  static constexpr int tbuf = a_ring + kStagesA * kBytesA;
  static constexpr int bars = tbuf + 2 * kTileM * kTbStride * 4;
  static constexpr int total = bars + 256;
};

__global__ void __launch_bounds__(kNumThreads, 1)
dh_gemm_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmWmn, DhArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t b_ring = smem_base + SmemLayout::b_ring;
  const uint32_t a_ring = smem_base + SmemLayout::a_ring;
  float* tbuf = reinterpret_cast<float*>(smem_gen + SmemLayout::tbuf);
  const uint32_t bars = smem_base + SmemLayout::bars;
  const uint32_t b_full = bars, b_empty = bars + 8 * kStagesB;
  const uint32_t a_full = bars + 16 * kStagesB, a_empty = a_full + 8 * kStagesA;
  const uint32_t tmem_full = a_empty + 8 * kStagesA, tmem_empty = tmem_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + SmemLayout::bars + 200);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = __ldg(p.tile_off + p.B);
  const int tile_end = min(total_tiles, p.tile_begin + p.tile_cap);
  const int nk = p.Vp / kBK;
  const int nblk_total = (p.Hp + kBN - 1) / kBN;
  const int npass = (nblk_total + 1) / 2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmG);
    tma_prefetch_desc(&tmWmn);
    for (int s = 0; s < kStagesB; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
    for (int s = 0; s < kStagesA; ++s) { mbar_init(a_full + 8 * s, 1); mbar_init(a_empty + 8 * s, 1); }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, 4);
    mbar_fence_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(tmem_slot), kTmemCols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ita = 0, itb = 0;
      for (int tile = p.tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
        const int ring_row0 = (tile - p.tile_begin) * kTileM;
        for (int pass = 0; pass < npass; ++pass) {
          const int nblk = min(2, nblk_total - pass * 2);
          for (int kc = 0; kc < nk; ++kc, ++ita) {
            const uint32_t sa = ita % kStagesA, pha = (ita / kStagesA) & 1;
            mbar_wait(a_empty + 8 * sa, pha ^ 1);
            mbar_expect_tx(a_full + 8 * sa, kBytesA);
            tma_load_2d(a_ring + sa * kBytesA, &tmG, a_full + 8 * sa, kc * kBK, ring_row0);
            for (int blk = 0; blk < nblk; ++blk, ++itb) {
              const uint32_t sb = itb % kStagesB, phb = (itb / kStagesB) & 1;
              mbar_wait(b_empty + 8 * sb, phb ^ 1);
              mbar_expect_tx(b_full + 8 * sb, kBytesB);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                tma_load_2d(b_ring + sb * kBytesB + j * 8192, &tmWmn, b_full + 8 * sb,
                            (pass * 2 + blk) * kBN + j * 64, kc * kBK);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_bf16(kTileM, kBN, 0, 1);
    uint32_t ita = 0, itb = 0, pc = 0;
    for (int tile = p.tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
      for (int pass = 0; pass < npass; ++pass, ++pc) {
        const int nblk = min(2, nblk_total - pass * 2);
        mbar_wait(tmem_empty, (pc & 1) ^ 1);
        tc_fence_after();
        for (int kc = 0; kc < nk; ++kc, ++ita) {
          const uint32_t sa = ita % kStagesA, pha = (ita / kStagesA) & 1;
          mbar_wait(a_full + 8 * sa, pha);
          for (int blk = 0; blk < nblk; ++blk, ++itb) {
            const uint32_t sb = itb % kStagesB, phb = (itb / kStagesB) & 1;
            mbar_wait(b_full + 8 * sb, phb);
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_addr = a_ring + sa * kBytesA, b_addr = b_ring + sb * kBytesB;
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                const uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024);        // K-major
                const uint64_t bd = make_smem_desc(b_addr + k * 2048, 8192, 1024);    // MN-major
                umma_f16(tmem_base + blk * kBN, ad, bd, idesc, (kc | k) != 0);
              }
              umma_commit(b_empty + 8 * sb);
            }
            __syncwarp();
          }
          if (lane == 0) umma_commit(a_empty + 8 * sa);
          __syncwarp();
        }
        if (lane == 0) umma_commit(tmem_full);
        __syncwarp();
      }
    }
  } else {
    const int lane_grp = warp & 3;
    const int row = lane_grp * 32 + lane;
    const int epi_tid = (warp - 2) * 32 + lane;
    const int rcol = epi_tid & 31, part = epi_tid >> 5;
    uint32_t pc = 0, cc = 0;
    for (int tile = p.tile_begin + blockIdx.x; tile < tile_end; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p.tile_off, p.T_len, p.U_len, p.B, tile);
      const int ring_row = (tile - p.tile_begin) * kTileM + row;
      const __nv_bfloat16* hrow = p.h_ring + static_cast<long long>(ring_row) * p.Hp;
      for (int pass = 0; pass < npass; ++pass, ++pc) {
        const int nblk = min(2, nblk_total - pass * 2);
        mbar_wait(tmem_full, pc & 1);
        tc_fence_after();
        for (int c32 = 0; c32 < nblk * (kBN / 32); ++c32) {
          const int col0 = pass * 2 * kBN + c32 * 32;
          if (col0 >= p.Hp) break;   // uniform: remaining columns are zero padding
          float v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + c32 * 32, v);
          tmem_ld_wait();
          if (p.dbg_dh != nullptr) {
            float* d = p.dbg_dh + static_cast<long long>(ring_row) * p.Hp + col0;
#pragma unroll
            for (int j = 0; j < 32; ++j) d[j] = v[j];
          }
          const uint4* h4 = reinterpret_cast<const uint4*>(hrow + col0);
          float* tb = tbuf + (cc & 1) * (kTileM * kTbStride);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 hh = __ldg(h4 + q);
            const uint32_t w[4] = {hh.x, hh.y, hh.z, hh.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float h0 = bf16lo_to_f32(w[e]), h1 = bf16hi_to_f32(w[e]);
              const int j = q * 8 + e * 2;
              tb[row * kTbStride + j] = v[j] * (1.f - h0 * h0);
              tb[row * kTbStride + j + 1] = v[j + 1] * (1.f - h1 * h1);
            }
          }
          named_bar_sync(1, 128);
          ++cc;
          const int col = col0 + rcol;
          if (col < p.H) {
            float su[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) su[j] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float st = 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float x = tb[(part * 32 + i * 8 + j) * kTbStride + rcol];
                st += x;
                su[j] += x;
              }
              const int t = tc.t0 + part * 4 + i;
              if (t < tc.Tb) atomicAdd(p.d_enc + (static_cast<long long>(tc.b) * p.T + t) * p.H + col, st);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int u = tc.u0 + j;
              if (u <= tc.Ub) atomicAdd(p.d_pred + (static_cast<long long>(tc.b) * p.U1 + u) * p.H + col, su[j]);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

}  // namespace

int launch_dh_gemm(const CUtensorMap& tmG, const CUtensorMap& tmWmn, const DhArgs& args, int grid,
                   cudaStream_t stream) {
  ProfScope prof_(kProfDh, stream);
  const size_t smem = SmemLayout::total + 1024;
  RB_CUDA_CHECK(cudaFuncSetAttribute(dh_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dh_gemm_kernel<<<grid, kNumThreads, smem, stream>>>(tmG, tmWmn, args);
  RB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace rb
