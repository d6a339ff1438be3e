// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM, descriptors.
// Hand-written PTX wrappers; no CUTLASS/CuTe dependency.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

namespace rb {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ---------------------------------------------------------------------------------------------
// Tile geometry shared by every kernel of the path (see DESIGN.md "Tile map").
// The lattice of every utterance is cut into HALF-TILES of 16(t) x 4(u) cells = 64 GEMM rows (row r <-> (t0 + (r>>2),
// u0 + (r&3))), numbered utterance by utterance, t-block by t-block, u-block by u-block; tile_off[b] is the number of
// half-tiles before utterance b.  A GEMM unit (128 rows, one CTA) is two half-tiles: consecutive ids in the forward,
// any two entries of the active list in the backward.  The activation buffer, the gradient ring and the work list are
// all organised in half-tiles.
constexpr int kTileT = 16;
constexpr int kTileU = 8;
constexpr int kTileM = 128;
constexpr int kHalfU = 4;
constexpr int kHalfRows = 64;
constexpr int kBK = 64;     // K chunk: 64 bf16 = one 128-byte swizzle row
constexpr int kBN = 256;    // N per MMA instruction / accumulator / B stage

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// ---------------------------------------------------------------------------------------------
// mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// (A suspend-time hint on try_wait was measured and rejected: wake-up latency rose so much that the forward kernel
// went from 2.9 ms to 5.4 ms.  So were explicit .acquire.cluster / .release.cluster qualifiers on the barriers the
// MMA issue loop touches: 1.6x slower; the default-scope forms below are what CUTLASS uses for CTA pairs, too.)
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("rnnt_b200: mbarrier timeout block %d thread %d bar 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}

// (Also measured and rejected: letting only lane 0 poll and parking the other lanes at __syncwarp() made every
// pipeline hand-off slower -- the forward went from 2.9 ms to 5.8 ms -- so all lanes of a waiting warp poll.
// A nanosleep back-off on the off-critical-path waits was neutral (3.27 vs 3.22 ms sustained) and was dropped.)

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2D tiled tensor maps
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives when all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = TMEM lane = tile row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2): one M=256 MMA spans both SMs, each CTA stages its own 128 rows of
// the A operand and HALF of the B operand, so the L2 -> shared-memory operand traffic per SM drops by a third to a half.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `saddr` (a shared::cta address of this CTA) in the CTA of rank `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on an mbarrier of either CTA of the pair (`bar_cluster_addr`)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const void* tmap, uint32_t bar_cluster_addr, int c0,
                                                 int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both CTAs, 128 rows each] * B[smem of both CTAs, N/2 rows each]; leader CTA only.
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the mbarrier at this shared-memory offset in every CTA of `cta_mask` once all MMAs issued so far are done
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout per the sm_100 matrix/instruction descriptor format)
//
// Shared-memory matrix descriptor, 128-byte swizzle, descriptor version 1:
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 | [61,64) layout type (2 = SWIZZLE_128B)
// K-major operand tile (rows = M or N, 64 bf16 of K per 128-byte row, 8-row 1024-byte swizzle atoms):
//   SBO = 1024 (next 8-row group), LBO unused (1).  Advance K by 16 elements = +32 bytes on the start address.
// MN-major operand tile (rows = K, 64 bf16 of M/N per 128-byte row):
//   SBO = 1024 (next 8 K-rows), LBO = byte stride between 64-element M/N blocks.
//   Advance K by 16 rows = +2048 bytes on the start address.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 operands and fp32 accumulation:
//   [4,6) D format (1 = f32) | [7,10) A format (0 = f16, 1 = bf16) | [10,13) B format (0 = f16, 1 = bf16)
//   [15] A major (0 = K, 1 = MN) | [16] B major | [17,23) N >> 3 | [24,29) M >> 4
// Operand formats: 0 = fp16, 1 = bf16 (A and B are encoded independently, so they may differ).
constexpr int kFmtF16 = 0;
constexpr int kFmtBF16 = 1;
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major, int a_fmt,
                                                  int b_fmt) {
  return (1u << 4) | (static_cast<uint32_t>(a_fmt) << 7) | (static_cast<uint32_t>(b_fmt) << 10) |
         (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// math
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Packed fp32 pairs (FFMA2 / FADD2 on sm_100: one issue slot for two lanes of work) and the 3-input max (FMNMX3).  The
// operands are adjacent registers of the tcgen05.ld result, so the b64 moves cost nothing.
__device__ __forceinline__ void fma2(float& x0, float& x1, float a, float b0, float b1) {   // x = x * a + b
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %2};\n\tmov.b64 rc, {%3, %4};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
      : "+f"(x0), "+f"(x1)
      : "f"(a), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void add2(float& x0, float& x1, float b0, float b1) {            // x = x + b
  asm("{.reg .b64 ra, rc, rd;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rc, {%2, %3};\n\t"
      "add.rn.f32x2 rd, ra, rc;\n\tmov.b64 {%0, %1}, rd;}"
      : "+f"(x0), "+f"(x1)
      : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void mul2(float& x0, float& x1, float a) {                        // x = x * a
  asm("{.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %2};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;}"
      : "+f"(x0), "+f"(x1)
      : "f"(a));
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// Gradient scale.  The logit-gradients g (|g| <= |dcost_b|: softmax x occupancy minus occupancy-weighted one-hots) are
// stored in fp16 as g * S with S = 2^kGradShift / 2^ceil(log2 max_b |dcost_b|), i.e. max |g S| in [2^11, 2^12]: typical
// entries (occupancy ~1e-2 x softmax ~1e-3 = 1e-5 of the maximum) then sit well inside fp16's NORMAL range (>= 6.1e-5)
// instead of its subnormals, which is what bounded the gradient error at the full bench shape (1.1e-3 -> 3e-4 relative).
constexpr int kGradShift = 12;
// Occupancy sparsity threshold, in units of max_b |dcost_b|: half-tiles whose every logit-gradient is below 2^-25 of the
// largest possible gradient entry are dropped from the backward (the fp32 reference itself resolves 2^-24 of it).
constexpr float kSkipBelow = 2.98023224e-8f * static_cast<float>(1 << kGradShift);   // 2^-25 in S-scaled units

// Deterministic mode of the backward: sums that cross CTAs are accumulated as 64-bit fixed point (2^-26 units; integer
// addition is associative, so the result does not depend on the order in which the atomics land).  The operands are
// the S-scaled gradients: entries are bounded by 2^13 max|W| U1 (d_enc), 2^13 B (T+U) (dW) < 2^36, typical ones are O(1),
// so 2^-26 resolution is ~1e-8 relative and the 64-bit sums cannot overflow.
constexpr float kFxScale = 67108864.f;                // 2^26
constexpr double kFxInv = 1.0 / 67108864.0;
__device__ __forceinline__ void fx_add(long long* p, float x) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(__float2ll_rn(x * kFxScale)));
}
__device__ __forceinline__ void red_add_v4(float* dst, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// tile -> utterance lookup: largest b with tile_off[b] <= tile (tile_off has B+1 entries).  Plain loads: the table
// may be the kernel's shared-memory copy (stage_len_tables) or the global one.
__device__ __forceinline__ int find_utt(const int* tile_off, int B, int tile) {
  int lo = 0, hi = B;  // invariant: tile_off[lo] <= tile < tile_off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (tile_off[mid] <= tile) lo = mid; else hi = mid;
  }
  return lo;
}

struct TileCoord {
  int b, t0, u0, Tb, Ub;
};
// half-tile id -> utterance, first frame t0, first label position u0, and the utterance's lengths
__device__ __forceinline__ TileCoord decode_half(const int* tile_off, const int* T_len, const int* U_len, int B, int T,
                                                 int U1, int half_id) {
  TileCoord c;
  c.b = find_utt(tile_off, B, half_id);
  // same clamp as tile_table_kernel / lattice_kernel / coef_kernel: an out-of-range length (reported through the status
  // word, which poisons the costs with NaN) must not desynchronise the tile enumeration or index past `targets`
  c.Tb = max(1, min(T_len[c.b], T));
  c.Ub = max(0, min(U_len[c.b], U1 - 1));
  const int nu = (c.Ub + 1 + kHalfU - 1) / kHalfU;
  const int local = half_id - tile_off[c.b];
  c.t0 = (local / nu) * kTileT;
  c.u0 = (local % nu) * kHalfU;
  return c;
}

// The three per-utterance tables every tile decode walks (binary search over tile_off, then the lengths), copied to
// shared memory once per CTA: the GEMM kernels leave the L1 a few KB next to > 200 KB of shared memory, so each step of
// that dependent chain was an L2 round trip in front of an accumulator drain.  Falls back to the global tables for
// B > kLenTabMaxB.  The caller synchronises the CTA before the first use.
constexpr int kLenTabMaxB = 340;
constexpr int kLenTabBytes = (3 * kLenTabMaxB + 1) * 4 + 12;       // 4096
struct LenTables {
  const int* tile_off;
  const int* T_len;
  const int* U_len;
};
__device__ __forceinline__ LenTables stage_len_tables(int* tab, const int* tile_off, const int* T_len, const int* U_len,
                                                      int B) {
  if (B > kLenTabMaxB) return LenTables{tile_off, T_len, U_len};
  for (int i = threadIdx.x; i <= B; i += blockDim.x) tab[i] = __ldg(tile_off + i);
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    tab[B + 1 + i] = __ldg(T_len + i);
    tab[2 * B + 1 + i] = __ldg(U_len + i);
  }
  return LenTables{tab, tab + B + 1, tab + 2 * B + 1};
}

}  // namespace rb
