// Fused joint GEMM kernel (forward "F" and backward-recompute "G" modes).
//
// Replaces rnnt/joint.py:32-39 (broadcast add + tanh + joint_ln) fused with the first half of
// torchaudio's rnnt_loss (ReduceMax2D / ReduceLogSumExpGivenMax2D / ComputeLogProbs, reference call
// site rnnt/model.py:35-41) in F mode, and with ComputeGradients in G mode.  The (B,T,U1,V) logit
// tensor only ever exists as fp32 accumulators in TMEM.
//
//   logits[c, v] = sum_k f16(tanh(enc[b,t,k] + pred[b,u,k])) * f16(W[v,k])  (+ bias[v] in the epilogue)
// (fp16 operands: tanh output lies in [-1,1] and joint weights are O(1), so fp16's 11-bit mantissa gives 8x
// finer rounding than bf16 at the same tensor-core rate; accumulation is fp32 in TMEM.)
//
// Warp roles (480 threads, 1 CTA / SM, persistent over 16(t)x8(u) lattice tiles):
//   warp 0      TMA producer: W tiles (256 rows x 64 k, 128B-swizzled) into a 4-stage ring
//   warp 1      tcgen05.mma issuer (M=128, N=256, K=16; two 256-column accumulators per N pass)
//   warps 2-5   epilogue: tcgen05.ld -> online log-sum-exp + blank/label gather (F) or
//               softmax*gamma - one-hots -> fp16 (scaled by S) -> smem -> TMA store into the gradient ring (G)
//   warps 6-13  A-operand producers: tanh(enc+pred) -> fp16 -> swizzled smem (4-stage ring)
//   warp 14     (G mode) TMA-stores each finished A stage into the hidden-activation ring
#include "common.cuh"
#include "kernels.h"

#ifndef RNNT_G_STAGES_A
#define RNNT_G_STAGES_A 2
#define RNNT_G_STAGES_B 4
#endif

namespace rb {

namespace {

// Shared-memory budget (227 KB): the W-tile (B operand) ring is the latency-critical one (TMA round trip ~1 us
// vs 0.26 us of MMA work per 32 KB stage), the computed A operand only needs double buffering.
template <int MODE> struct StagesA { static constexpr int value = MODE == 0 ? 4 : RNNT_G_STAGES_A; };
template <int MODE> struct StagesB { static constexpr int value = MODE == 0 ? 4 : RNNT_G_STAGES_B; };
constexpr int kBytesA = kTileM * kBK * 2;  // 16 KB
constexpr int kBytesB = kBN * kBK * 2;     // 32 KB
constexpr int kBytesG = kTileM * 64 * 2;   // 16 KB staging box for the gradient ring store
constexpr int kNumThreads = 608;
constexpr int kFirstEpiWarp = 2;      // warps 2-5: epilogue set 0 (accumulator 0), warps 6-9: set 1 (accumulator 1)
constexpr int kNumEpiWarps = 8;
constexpr int kFirstProdWarp = 10;
constexpr int kNumProdWarps = 8;
constexpr int kStoreWarp = 18;
constexpr int kTmemCols = 512;

template <int MODE>
struct SmemLayout {
  // offsets from the 1024-aligned base
  static constexpr int b_ring = 0;
  static constexpr int a_ring = b_ring + StagesB<MODE>::value * kBytesB;
  static constexpr int g_stage = a_ring + StagesA<MODE>::value * kBytesA;
  static constexpr int xchg = g_stage + (MODE == 1 ? 4 * kBytesG : 0);     // F: set-1 -> set-0 softmax partials
  static constexpr int bars = xchg + (MODE == 0 ? 2 * kTileM * 16 : 0);
  static constexpr int total = bars + 256;
};

}  // namespace

size_t joint_gemm_smem_bytes() { return SmemLayout<0>::total + 1024; }

template <int MODE>  // 0 = forward (lse + gather), 1 = backward recompute (gradient ring)
__global__ void __launch_bounds__(kNumThreads, 1)
joint_gemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmG,
                  const __grid_constant__ CUtensorMap tmHr, JointArgs p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  using SL = SmemLayout<MODE>;
  constexpr int kStagesB = StagesB<MODE>::value;
  constexpr int kStagesA = StagesA<MODE>::value;

  const uint32_t b_ring = smem_base + SL::b_ring;
  const uint32_t a_ring = smem_base + SL::a_ring;
  const uint32_t g_stage = smem_base + SL::g_stage;
  const uint32_t bars = smem_base + SL::bars;
  // barrier map (8 bytes each)
  const uint32_t b_full = bars, b_empty = bars + 8 * kStagesB;
  const uint32_t a_full = bars + 16 * kStagesB, a_empty = a_full + 8 * kStagesA;
  const uint32_t tmem_full = a_empty + 8 * kStagesA, tmem_empty = tmem_full + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + SL::bars + 200);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int total_tiles = p.n_active ? __ldg(p.n_active) : __ldg(p.tile_off + p.B);
  const int tile_end = min(total_tiles, p.tile_begin + p.tile_cap);   // work-list slots [tile_begin, tile_end)
  const int nk = p.Hp / kBK;
  const int nblk_total = p.Vp / kBN;
  const int npass = (nblk_total + 1) / 2;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmW);
    if (MODE == 1) {
      tma_prefetch_desc(&tmG);
      tma_prefetch_desc(&tmHr);
    }
    for (int s = 0; s < kStagesB; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    for (int s = 0; s < kStagesA; ++s) {
      mbar_init(a_full + 8 * s, kNumProdWarps);
      mbar_init(a_empty + 8 * s, MODE == 1 ? 2 : 1);
    }
    mbar_init(tmem_full, 1);
    mbar_init(tmem_empty, kNumEpiWarps);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer (W tiles)
    if (lane == 0) {
      uint32_t it = 0;
      for (int slot = p.tile_begin + blockIdx.x; slot < tile_end; slot += gridDim.x) {
        for (int pass = 0; pass < npass; ++pass) {
          const int nblk = min(2, nblk_total - pass * 2);
          for (int kc = 0; kc < nk; ++kc) {
            for (int blk = 0; blk < nblk; ++blk, ++it) {
              const uint32_t s = it % kStagesB, ph = (it / kStagesB) & 1;
              mbar_wait(b_empty + 8 * s, ph ^ 1);
              mbar_expect_tx(b_full + 8 * s, kBytesB);
              tma_load_2d(b_ring + s * kBytesB, &tmW, b_full + 8 * s, kc * kBK, (pass * 2 + blk) * kBN);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    constexpr uint32_t idesc = make_idesc(kTileM, kBN, 0, 0, kFmtF16, kFmtF16);
    uint32_t ita = 0, itb = 0, pc = 0;
    for (int slot = p.tile_begin + blockIdx.x; slot < tile_end; slot += gridDim.x) {
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
                const uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024);
                const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024);
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
  } else if (warp >= kFirstEpiWarp && warp < kFirstEpiWarp + kNumEpiWarps) {
    // ===================================================================== epilogue
    // Two epilogue sets work in parallel: set e drains accumulator e (256 columns) of every pass.
    const int lane_grp = warp & 3;  // TMEM lane quarter this warp may access
    const int eset = (warp - kFirstEpiWarp) >> 2;
    const int row = lane_grp * 32 + lane;
    const int ti = row >> 3, ui = row & 7;
    const int set_tid = ((warp - kFirstEpiWarp) & 3) * 32 + lane;
    float4* xchg = reinterpret_cast<float4*>(smem_gen + SL::xchg);
    uint32_t pc = 0, box_count = 0, tcount = 0;
    for (int slot = p.tile_begin + blockIdx.x; slot < tile_end; slot += gridDim.x, ++tcount) {
      const int tile = p.tile_list ? __ldg(p.tile_list + slot) : slot;
      const TileCoord tc = decode_tile(p.tile_off, p.T_len, p.U_len, p.B, tile);
      const int t = tc.t0 + ti, u = tc.u0 + ui;
      const bool valid = (t < tc.Tb) && (u <= tc.Ub);
      const long long cell = (static_cast<long long>(tc.b) * p.T + min(t, p.T - 1)) * p.U1 + min(u, p.U1 - 1);
      int tgt = -1;
      if (valid && u < tc.Ub) tgt = __ldg(p.targets + static_cast<long long>(tc.b) * p.tgt_ld + u);
      // F state (log2 units)
      float m = -INFINITY, ssum = 0.f, x_tgt = -INFINITY, x_blank = -INFINITY;
      // G state
      float gam = 0.f, eB = 0.f, eE = 0.f, lse2 = 1e30f, cbound = 0.f;
      if (MODE == 1 && valid) {
        const float4 c4 = __ldg(p.coef + cell);
        gam = c4.x; eB = c4.y; eE = c4.z; lse2 = c4.w * kLog2e;
        if (p.clamp > 0.f) cbound = p.clamp * fabsf(p.dcost ? __ldg(p.dcost + tc.b) : 1.f) * __ldg(p.gscale);
      }
      const int ring_row0 = (slot - p.tile_begin) * kTileM;

      for (int pass = 0; pass < npass; ++pass, ++pc) {
        const int nblk = min(2, nblk_total - pass * 2);
        mbar_wait(tmem_full, pc & 1);
        tc_fence_after();
        const int nchunk = ((p.dbg & 1) || eset >= nblk) ? 0 : (kBN / 32);
        for (int c32 = 0; c32 < nchunk; ++c32) {
          const int col0 = (pass * 2 + eset) * kBN + c32 * 32;  // global column of v[0]
          float v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + eset * kBN + c32 * 32, v);
          tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(p.bias2 + col0);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 bb = __ldg(b4 + q);
            v[4 * q + 0] = fmaf(v[4 * q + 0], kLog2e, bb.x);
            v[4 * q + 1] = fmaf(v[4 * q + 1], kLog2e, bb.y);
            v[4 * q + 2] = fmaf(v[4 * q + 2], kLog2e, bb.z);
            v[4 * q + 3] = fmaf(v[4 * q + 3], kLog2e, bb.w);
          }
          if (MODE == 0) {
            float cmax = v[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, v[j]);
            const float m_new = fmaxf(m, cmax);
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              acc += ex2_approx(v[j] - m_new);
              if (col0 + j == tgt) x_tgt = v[j];
            }
            ssum = ssum * ex2_approx(m - m_new) + acc;
            m = m_new;
            if (p.blank >= col0 && p.blank < col0 + 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j == p.blank) x_blank = v[j];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float g = ex2_approx(v[j] - lse2) * gam;
              if (col0 + j == tgt) g -= eE;
              v[j] = g;
            }
            if (p.blank >= col0 && p.blank < col0 + 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j == p.blank) v[j] -= eB;
            }
            if (p.clamp > 0.f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], -cbound), cbound);
            }
            // stage fp16 into this set's 128 x 64 swizzled box (two 32-column halves per box)
            const uint32_t buf = g_stage + (eset * 2 + (box_count & 1)) * kBytesG;
            if ((c32 & 1) == 0) {
              // the TMA store that last used this buffer (two boxes ago) must have finished reading it
              if (set_tid == 0) tma_store_wait_read<1>();
              named_bar_sync(1 + eset, 128);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t w0 = pack_f16x2(v[8 * q + 0], v[8 * q + 1]);
              const uint32_t w1 = pack_f16x2(v[8 * q + 2], v[8 * q + 3]);
              const uint32_t w2 = pack_f16x2(v[8 * q + 4], v[8 * q + 5]);
              const uint32_t w3 = pack_f16x2(v[8 * q + 6], v[8 * q + 7]);
              const uint32_t chunk = (c32 & 1) * 4 + q;
              const uint32_t addr = buf + row * 128 + ((chunk ^ (row & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2),
                           "r"(w3)
                           : "memory");
            }
            if ((c32 & 1) == 1) {
              fence_proxy_async();
              named_bar_sync(1 + eset, 128);
              if (set_tid == 0) {
                tma_store_2d(&tmG, buf, (pass * 2 + eset) * kBN + (c32 >> 1) * 64, ring_row0);
                tma_store_commit();
              }
              ++box_count;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty);
      }
      if (MODE == 0) {
        // combine the two sets' online-softmax partials: set 1 -> smem -> set 0 writes lp / lse
        float4* slot = xchg + (tcount & 1) * kTileM + row;
        if (eset == 1) *slot = make_float4(m, ssum, x_tgt, x_blank);
        named_bar_sync(3, 256);
        if (eset == 0 && valid && !(p.dbg & 1)) {
          const float4 o = *slot;
          const float mm = fmaxf(m, o.x);
          const float stot = ssum * ex2_approx(m - mm) + o.y * ex2_approx(o.x - mm);
          const float l2 = mm + lg2_approx(stot);
          p.lse[cell] = l2 * kLn2;
          float2 out;
          out.x = (fmaxf(x_blank, o.w) - l2) * kLn2;
          out.y = (tgt >= 0) ? (fmaxf(x_tgt, o.z) - l2) * kLn2 : 0.f;
          reinterpret_cast<float2*>(p.lp)[cell] = out;
        }
      }
    }
    if (MODE == 1 && set_tid == 0) tma_store_wait_all<0>();
  } else if (warp >= kFirstProdWarp && warp < kFirstProdWarp + kNumProdWarps) {
    // ===================================================================== A producers
    // Warp rg produces rows rg*16 .. rg*16+15 of the tile (t-rows 2rg, 2rg+1; all 8 u).  Lane l owns the 16-byte
    // column chunk c = l & 7 (8 hidden units) of the four rows rs, rs+4, rs+8, rs+12 (rs = l >> 3), i.e. the pairs
    // (t-row 0|1) x (u = rs | rs+4): 4 x 8 inputs, 32 tanh, four 16-byte shared stores.  Few wide stores matter:
    // fence.proxy.async is a MEMBAR.ALL.CTA whose latency grows with the number of shared stores in flight.
    const int rg = warp - kFirstProdWarp;
    const int c = lane & 7, rs = lane >> 3;
    uint32_t it = 0;
    for (int slot = p.tile_begin + blockIdx.x; slot < tile_end; slot += gridDim.x) {
      const int tile = p.tile_list ? __ldg(p.tile_list + slot) : slot;
      const TileCoord tc = decode_tile(p.tile_off, p.T_len, p.U_len, p.B, tile);
      const float* e_ptr[2];
      const float* p_ptr[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        e_ptr[i] = p.enc + tc.b * p.enc_sb + static_cast<long long>(min(tc.t0 + 2 * rg + i, p.T - 1)) * p.enc_st + 8 * c;
        p_ptr[i] = p.pred + tc.b * p.pred_sb + static_cast<long long>(min(tc.u0 + rs + 4 * i, p.U1 - 1)) * p.pred_su + 8 * c;
      }
      float4 e_cur[2][2], p_cur[2][2];
      auto load_chunk = [&](int kc, float4 (&e)[2][2], float4 (&q)[2][2]) {
        const int col = kc * kBK + 8 * c;
        if (col < p.H) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            e[i][0] = __ldg(reinterpret_cast<const float4*>(e_ptr[i] + kc * kBK));
            e[i][1] = __ldg(reinterpret_cast<const float4*>(e_ptr[i] + kc * kBK) + 1);
            q[i][0] = __ldg(reinterpret_cast<const float4*>(p_ptr[i] + kc * kBK));
            q[i][1] = __ldg(reinterpret_cast<const float4*>(p_ptr[i] + kc * kBK) + 1);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            e[i][0] = e[i][1] = q[i][0] = q[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      load_chunk(0, e_cur, p_cur);
      for (int pass = 0; pass < npass; ++pass) {
        for (int kc = 0; kc < nk; ++kc, ++it) {
          const uint32_t s = it % kStagesA, ph = (it / kStagesA) & 1;
          mbar_wait(a_empty + 8 * s, ph ^ 1);
          const uint32_t stage = a_ring + s * kBytesA;
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {        // t-row
#pragma unroll
              for (int j = 0; j < 2; ++j) {      // u = rs + 4j
                const float4 a0 = e_cur[i][0], a1 = e_cur[i][1], b0 = p_cur[j][0], b1 = p_cur[j][1];
                const uint32_t w0 = pack_f16x2(tanh_approx(a0.x + b0.x), tanh_approx(a0.y + b0.y));
                const uint32_t w1 = pack_f16x2(tanh_approx(a0.z + b0.z), tanh_approx(a0.w + b0.w));
                const uint32_t w2 = pack_f16x2(tanh_approx(a1.x + b1.x), tanh_approx(a1.y + b1.y));
                const uint32_t w3 = pack_f16x2(tanh_approx(a1.z + b1.z), tanh_approx(a1.w + b1.w));
                const int ui = rs + 4 * j;                       // = row & 7
                const int r = rg * 16 + i * 8 + ui;
                const uint32_t addr = stage + r * 128 + ((static_cast<uint32_t>(c) ^ static_cast<uint32_t>(ui)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3)
                             : "memory");
              }
            }
          }
          // the inputs of the next chunk are requested now (their registers are free again); the fence, the barrier
          // hand-off and the other producer warp of this scheduler cover the load latency
          if (kc + 1 < nk || pass + 1 < npass) load_chunk((kc + 1 < nk) ? kc + 1 : 0, e_cur, p_cur);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(a_full + 8 * s);
        }
      }
    }
  } else if (warp == kStoreWarp) {
    // ===================================================================== hidden-ring store (G mode)
    if (MODE == 1 && lane == 0) {
      uint32_t it = 0;
      for (int slot = p.tile_begin + blockIdx.x; slot < tile_end; slot += gridDim.x) {
        const int ring_row0 = (slot - p.tile_begin) * kTileM;
        for (int pass = 0; pass < npass; ++pass) {
          for (int kc = 0; kc < nk; ++kc, ++it) {
            const uint32_t s = it % kStagesA, ph = (it / kStagesA) & 1;
            mbar_wait(a_full + 8 * s, ph);
            if (pass == 0) {
              tma_store_2d(&tmHr, a_ring + s * kBytesA, kc * kBK, ring_row0);
              tma_store_commit();
              tma_store_wait_read<0>();
            }
            mbar_arrive(a_empty + 8 * s);
          }
        }
      }
      tma_store_wait_all<0>();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

int launch_joint_gemm(int mode, const CUtensorMap& tmW, const CUtensorMap& tmG, const CUtensorMap& tmHr,
                      const JointArgs& args, int grid, cudaStream_t stream) {
  ProfScope prof_(mode == 0 ? kProfJointF : kProfJointG, stream);
  const size_t smem = (mode == 0 ? SmemLayout<0>::total : SmemLayout<1>::total) + 1024;
  if (mode == 0) {
    RB_CUDA_CHECK(cudaFuncSetAttribute(joint_gemm_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    joint_gemm_kernel<0><<<grid, kNumThreads, smem, stream>>>(tmW, tmG, tmHr, args);
  } else {
    RB_CUDA_CHECK(cudaFuncSetAttribute(joint_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    joint_gemm_kernel<1><<<grid, kNumThreads, smem, stream>>>(tmW, tmG, tmHr, args);
  }
  RB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace rb
