// Block-scaled FP8 GEMM (OCP MXFP8: e4m3 elements, one UE8M0 scale per 32 K-elements) on tcgen05:
//   tcgen05.mma.kind::mxf8f6f4.block_scale, operands through 128B-swizzled TMA tiles (K-block = 128 fp8 = 128 B),
//   scale factors staged global -> smem (cp.async.bulk) -> TMEM (tcgen05.cp 32x128b.warpx4), fp32 accumulators
//   double-buffered in TMEM, the SAME fused epilogues as the bf16 GEMM (bias/GELU/gated residual/QKV+RoPE/...).
//
// Scale-factor storage ("chunk" layout, identical for A and B, produced by quantize_mxfp8_rows below):
//   for every 128 rows x 128 K-elements one 512-byte chunk; inside it the scale of row r = m0 + 32*m1
//   (m0 < 32, m1 < 4) and K-block kb < 4 sits at byte m0*16 + m1*4 + kb.  That is exactly the
//   32-lane x 128-bit image tcgen05.cp expects: lane m0, 32-bit column m1, byte kb (selected per MMA through
//   the instruction descriptor's sf_id fields).
#include <cstdlib>

#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include "../common/host.h"
#include "gemm_tcgen05.cuh"

namespace pa {

// Tile shapes.  A 128x128x32 SS-MMA streams 8 KB of operands per 64 tensor cycles = the whole 128 B/clk
// shared-memory bandwidth of an SM, so N = 128 tiles run at half of the fp8 peak; N >= 224 is needed.  TMEM has
// 512 columns: double-buffered accumulators (2 x N) + scale factors (4 + 4*ceil(N/128)) fit for N <= 240
// -> <224, 2> for the generic epilogues; the fused QKV epilogue needs whole 128-wide heads per tile
// -> <256, 1> (single accumulator, epilogue not overlapped with the next tile's main loop).
template <int BN, int ACC>
struct Mx8Cfg {
  static constexpr int BM = 128, BK = 128;                       // BK in elements == bytes
  static constexpr int NCHUNK = (BN + 127) / 128;                // 128-row scale-factor chunks per B tile
  static constexpr int STAGES = 4;
  static constexpr uint32_t A_BYTES = BM * BK, B_BYTES = BN * BK;
  static constexpr uint32_t SFA_BYTES = 512, SFB_BYTES = 512 * NCHUNK;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES + SFA_BYTES + SFB_BYTES;
  static constexpr uint32_t STAGE_STRIDE = (STAGE_BYTES + 1023) / 1024 * 1024;
  static constexpr uint32_t ACC_COLS = ACC * BN;
  static constexpr uint32_t SF_COL = ACC_COLS;                   // SFA: 4 columns, SFB: 4 * NCHUNK columns
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_STRIDE + 256 + 1024;
  static_assert(ACC_COLS + 4 + 4 * NCHUNK <= 512, "TMEM budget");
};

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_dst),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

// smem descriptor of a 32-row x 16-byte scale-factor image (no swizzle; 8-row atoms 128 B apart)
__device__ __forceinline__ uint64_t make_sf_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(128 >> 4) << 32;   // SBO
  d |= static_cast<uint64_t>(1) << 46;          // descriptor version (sm_100)
  return d;
}

template <int BN, int ACC>
__global__ void __launch_bounds__(384, 1)
gemm_mxfp8_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const uint8_t* __restrict__ sfa, const uint8_t* __restrict__ sfb, const GemmParams p) {
  using Cfg = Mx8Cfg<BN, ACC>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES, NCHUNK = Cfg::NCHUNK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_STRIDE);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);      // provably warp-uniform role index

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 8);          // eight epilogue warps
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 2) ptx::tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  static_assert(ACC == 1 || ACC == 2, "one or two accumulator stages");
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_per_batch = (p.rows + BM - 1) / BM;
  const int num_m = m_per_batch * p.batch;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = p.K / BK;                      // K % 128 == 0 (checked on the host)
  const int sfa_mt = p.sfa_mtiles > 0 ? p.sfa_mtiles : m_per_batch;      // scale chunks per batch entry of A's buffer
  constexpr int GROUP_M = 8;
  auto decode = [&](int t, int& mt, int& nt) {
    const int per_group = GROUP_M * num_n;
    const int g = t / per_group;
    const int first = g * GROUP_M;
    const int gsz = min(num_m - first, GROUP_M);
    const int r = t - g * per_group;
    mt = first + r % gsz;
    nt = r / gsz;
  };

  // Producer and MMA roles run as whole warps on warp-uniform values with one elected lane issuing (inside a
  // `lane == 0` branch ptxas wraps every UTMALDG / UTCQMMA in an ELECT + R2UR.BROADCAST waterfall loop).
  // Single accumulator (ACC == 1): the epilogue warps keep a whole 128-column accumulator row in registers (drain-first
  // epilogue below), so registers move from the producer / MMA warpgroup to the two epilogue warpgroups.  The
  // setmaxnreg sits INSIDE each role's branch so that ptxas allocates the branch with that budget.
  if (warp < 4) {
   if constexpr (ACC == 1) ptx::setmaxnreg_dec<72>();   // 128*72 + 256*216 == 384*168: the pool is what WG0 frees
   if (warp_u == 0) {
    const bool leader = ptx::elect_one();
    const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      int mt, nt;
      decode(t, mt, nt);
      const int b = mt / m_per_batch, mrow = (mt - b * m_per_batch) * BM;
      const uint8_t* sfa_t = sfa + (static_cast<long long>(b) * sfa_mt + (mt - b * m_per_batch)) * num_k * 512;
      const uint8_t* sfb_t = sfb + static_cast<long long>(nt) * NCHUNK * num_k * 512;
      for (int kb = 0; kb < num_k; ++kb) {
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        if (leader) {
          ptx::mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          const uint32_t sa = smem_u + stage * Cfg::STAGE_STRIDE;
          ptx::tma_load_3d_s(sa, &tmA, &full[stage], kb * BK, mrow, b);
          ptx::tma_load_2d_s(sa + Cfg::A_BYTES, &tmB, &full[stage], kb * BK, nt * BN);
          bulk_load_1d(sa + Cfg::A_BYTES + Cfg::B_BYTES, sfa_t + static_cast<long long>(kb) * 512, 512, &full[stage]);
#pragma unroll
          for (int j = 0; j < NCHUNK; ++j)
            bulk_load_1d(sa + Cfg::A_BYTES + Cfg::B_BYTES + 512 + j * 512,
                         sfb_t + (static_cast<long long>(j) * num_k + kb) * 512, 512, &full[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp_u == 1) {
    const bool leader = ptx::elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    const uint32_t sfa_t = tmem_u + Cfg::SF_COL, sfb_t = tmem_u + Cfg::SF_COL + 4;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const int acc = ACC == 2 ? (it & 1) : 0;
      const uint32_t acc_phase = ACC == 2 ? ((it >> 1) & 1) : (it & 1);
      ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_u + acc * BN;
      for (int kb = 0; kb < num_k; ++kb) {
        ptx::mbar_wait(&full[stage], phase);
        ptx::tc_fence_after();
        if (leader) {
          const uint32_t sa = smem_u + stage * Cfg::STAGE_STRIDE;
          // scale factors of this K-block: smem -> TMEM (ordered with the MMAs in the tensor-core pipe)
          ptx::tmem_cp_32x128b_warpx4(sfa_t, make_sf_desc(sa + Cfg::A_BYTES + Cfg::B_BYTES));
#pragma unroll
          for (int j = 0; j < NCHUNK; ++j)
            ptx::tmem_cp_32x128b_warpx4(sfb_t + 4 * j, make_sf_desc(sa + Cfg::A_BYTES + Cfg::B_BYTES + 512 + j * 512));
          const uint64_t adesc = ptx::make_desc_kmajor_sw128(sa);
          const uint64_t bdesc = ptx::make_desc_kmajor_sw128(sa + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {        // 4 x (K = 32 fp8 = 32 bytes); sf_id = K-block inside the chunk
            const uint32_t idesc = ptx::make_idesc_mxf8(BM, BN, 0, 0, k, k);
            ptx::mma_mxf8_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, sfa_t, sfb_t, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::tc_commit(&empty[stage]);
          if (kb == num_k - 1) ptx::tc_commit(&tfull[acc]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
   }
  } else {
    if constexpr (ACC == 1) ptx::setmaxnreg_inc<216>();
    // eight epilogue warps, two per TMEM lane quadrant: the fp8 main loop is twice as fast as the bf16 one, so the
    // epilogue (same cost per element) would otherwise set the tile period (see gemm_2cta.cuh)
    constexpr int H0 = BN >= 256 ? 128 : (BN > 128 ? 128 : 64);      // columns of the first half: 128 | 128 | 64
    constexpr int H1 = BN - H0;                                      //                second half: 128 |  96 | 64
    const int q4 = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r_in_tile = q4 * 32 + lane;
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      int mt, nt;
      decode(t, mt, nt);
      const int b = mt / m_per_batch;
      const int row = (mt - b * m_per_batch) * BM + r_in_tile;
      const int acc = ACC == 2 ? (it & 1) : 0;
      ptx::mbar_wait(&tfull[acc], ACC == 2 ? ((it >> 1) & 1) : (it & 1));
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + acc * BN;
      if constexpr (ACC == 1 && BN == 256) {
        if (p.mode == EPI_QKV_ROPE) {
          // drain-first: this warp's 32 rows x 128 columns (one attention head / 128 MLP columns) -> registers, hand
          // the single accumulator back to the MMA warp, THEN do the RMSNorm / RoPE / GELU math and the stores while
          // the next tile's main loop already runs (measured before: 1.53 PFLOP/s with the tensor pipe idle for the
          // whole epilogue vs 2.5 PFLOP/s for the double-buffered 224-wide tiles).
          const int ng = nt * BN + half * 128;
          const bool live = ng < p.N;
          uint32_t areg[128];
          if (live) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              ptx::tmem_ld_32x32b_x32(taddr + half * 128 + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&areg[c * 32]));
            ptx::tmem_ld_wait();
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
          if (live) {
            if (p.bias != nullptr && p.rope != nullptr)
              epilogue_qkv_from_regs_fast(p, areg, b, row, row < p.rows, ng);
            else
              epilogue_qkv_from_regs(p, areg, b, row, row < p.rows, ng);
          }
          continue;
        }
      }
      if (half == 0)
        epilogue_tile<H0>(p, taddr, b, row, row < p.rows, nt * BN);
      else if (nt * BN + H0 < p.N)
        epilogue_tile<H1>(p, taddr + H0, b, row, row < p.rows, nt * BN + H0);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ CTA-pair variant
// A cluster of two CTAs computes one 256 x BN tile (tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale, M = 256).
// Per 128-byte K block a CTA fills 16 KB of A + BN/2 rows of B (16 KB) instead of 16 + 32 KB: the one-CTA kernel is
// bound by the L2 -> SM operand stream (ncu: 10.7 TB/s at 56 % tensor pipe), not by the tensor pipe.
// Scale factors: each CTA keeps the SFA of ITS 128 rows and the SFB of ALL BN columns in its own tensor memory (the
// tensor core of a CTA scales its 128 x BN accumulator half).  They travel as 512-byte rows of a plain (unswizzled)
// tensor map so that the peer's loads can complete on the leader's barrier like its A / B tiles do, and one
// tcgen05.cp.cta_group::2 issued by the leader copies smem -> TMEM in both CTAs.
template <int BN, int ACC>
struct Mx8PairCfg {
  static constexpr int BM = 128, BK = 128;
  static constexpr int NCHUNK = (BN + 127) / 128;
  static constexpr int STAGES = 6;
  static constexpr uint32_t A_BYTES = BM * BK, B_BYTES = (BN / 2) * BK;
  static constexpr uint32_t SFA_BYTES = 512, SFB_BYTES = 512 * NCHUNK;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES + SFA_BYTES + SFB_BYTES;
  static constexpr uint32_t STAGE_STRIDE = (STAGE_BYTES + 1023) / 1024 * 1024;
  static constexpr uint32_t ACC_COLS = ACC * BN;
  static constexpr uint32_t SF_COL = ACC_COLS;
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_STRIDE + 256 + 1024;
  static_assert(ACC_COLS + 4 + 4 * NCHUNK <= 512, "TMEM budget");
  static_assert(B_BYTES % 1024 == 0, "B half must be whole 8-row swizzle atoms");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int BN, int ACC>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_mxfp8_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ CUtensorMap tmSFA, const __grid_constant__ CUtensorMap tmSFB,
                       const GemmParams p) {
  using Cfg = Mx8PairCfg<BN, ACC>;
  constexpr int BM = Cfg::BM, BK = Cfg::BK, STAGES = Cfg::STAGES, NCHUNK = Cfg::NCHUNK;
  constexpr uint32_t SF_OFF = Cfg::A_BYTES + Cfg::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_STRIDE);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t rank = ptx::cluster_ctarank();              // 0: leader (issues copies + MMAs), 1: peer
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmSFA);
    ptx::prefetch_tmap(&tmSFB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 16);         // eight epilogue warps of each CTA of the pair
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 2) ptx::tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m128_per_batch = (p.rows + BM - 1) / BM;                 // scale-factor chunks are per 128 rows
  const int m_per_batch = (p.rows + 2 * BM - 1) / (2 * BM);          // 256-row pair tiles per batch entry
  const int num_m = m_per_batch * p.batch;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = p.K / BK;
  constexpr int GROUP_M = 4;
  auto decode = [&](int t, int& mt, int& nt) {
    const int per_group = GROUP_M * num_n;
    const int g = t / per_group;
    const int first = g * GROUP_M;
    const int gsz = min(num_m - first, GROUP_M);
    const int r = t - g * per_group;
    mt = first + r % gsz;
    nt = r / gsz;
  };

  if (warp < 4) {
    if constexpr (ACC == 1) ptx::setmaxnreg_dec<72>();
    if (warp_u == 0) {
      // ===================== TMA producer (both CTAs; all bytes land on the leader's `full`) =====================
      const bool leader_lane = ptx::elect_one();
      const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        int mt, nt;
        decode(t, mt, nt);
        const int b = mt / m_per_batch, mp = mt - b * m_per_batch;
        const int mrow = mp * 2 * BM + static_cast<int>(rank) * BM;
        const int nrow = nt * BN + static_cast<int>(rank) * (BN / 2);
        // a pair whose second half lies beyond `rows` still loads (zero-filled) A rows; its scale chunk is clamped
        // to the last real one (those accumulator rows are never stored)
        const int m128 = min(mp * 2 + static_cast<int>(rank), m128_per_batch - 1);
        const int sfa_row = (b * (p.sfa_mtiles > 0 ? p.sfa_mtiles : m128_per_batch) + m128) * num_k;
        const int sfb_row = nt * NCHUNK * num_k;
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          if (leader_lane) {
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t sa = smem_u + stage * Cfg::STAGE_STRIDE;
            ptx::tma_load_3d_2cta(sa, &tmA, &full[stage], kb * BK, mrow, b);
            ptx::tma_load_2d_2cta(sa + Cfg::A_BYTES, &tmB, &full[stage], kb * BK, nrow);
            ptx::tma_load_2d_2cta(sa + SF_OFF, &tmSFA, &full[stage], 0, sfa_row + kb);
#pragma unroll
            for (int j = 0; j < NCHUNK; ++j)
              ptx::tma_load_2d_2cta(sa + SF_OFF + 512 + j * 512, &tmSFB, &full[stage], 0, sfb_row + j * num_k + kb);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    } else if (warp_u == 1 && rank == 0) {
      // ===================== scale-factor copies + MMAs (leader CTA only) =====================
      const bool leader_lane = ptx::elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t sfa_t = tmem_u + Cfg::SF_COL, sfb_t = tmem_u + Cfg::SF_COL + 4;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        const int acc = ACC == 2 ? (it & 1) : 0;
        const uint32_t acc_phase = ACC == 2 ? ((it >> 1) & 1) : (it & 1);
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_u + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          if (leader_lane) {
            const uint32_t sa = smem_u + stage * Cfg::STAGE_STRIDE;
            ptx::tmem_cp_32x128b_warpx4_2cta(sfa_t, make_sf_desc(sa + SF_OFF));
#pragma unroll
            for (int j = 0; j < NCHUNK; ++j)
              ptx::tmem_cp_32x128b_warpx4_2cta(sfb_t + 4 * j, make_sf_desc(sa + SF_OFF + 512 + j * 512));
            const uint64_t adesc = ptx::make_desc_kmajor_sw128(sa);
            const uint64_t bdesc = ptx::make_desc_kmajor_sw128(sa + Cfg::A_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t idesc = ptx::make_idesc_mxf8(2 * BM, BN, 0, 0, k, k);
              ptx::mma_mxf8_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, sfa_t, sfb_t, (kb | k) != 0 ? 1u : 0u);
            }
            ptx::tc_commit_2cta(&empty[stage], 3);
            if (kb == num_k - 1) ptx::tc_commit_2cta(&tfull[acc], 3);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    }
  } else {
    if constexpr (ACC == 1) ptx::setmaxnreg_inc<216>();
    // eight epilogue warps per CTA: this CTA's 128 rows of the pair's tile, two warps per TMEM lane quadrant
    constexpr int H0 = BN > 128 ? 128 : 64;
    constexpr int H1 = BN - H0;
    const int q4 = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r_in_tile = q4 * 32 + lane;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      int mt, nt;
      decode(t, mt, nt);
      const int b = mt / m_per_batch;
      const int row = (mt - b * m_per_batch) * 2 * BM + static_cast<int>(rank) * BM + r_in_tile;
      const int acc = ACC == 2 ? (it & 1) : 0;
      ptx::mbar_wait(&tfull[acc], ACC == 2 ? ((it >> 1) & 1) : (it & 1));
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + acc * BN;
      if constexpr (ACC == 1 && BN == 256) {
        if (p.mode == EPI_QKV_ROPE) {
          // drain-first (see the one-CTA kernel): registers first, accumulator back to the leader's MMA warp, then math
          const int ng = nt * BN + half * 128;
          const bool live = ng < p.N;
          uint32_t areg[128];
          if (live) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              ptx::tmem_ld_32x32b_x32(taddr + half * 128 + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&areg[c * 32]));
            ptx::tmem_ld_wait();
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(&tempty[acc], 0);
          if (live) {
            if (p.bias != nullptr && p.rope != nullptr)
              epilogue_qkv_from_regs_fast(p, areg, b, row, row < p.rows, ng);
            else
              epilogue_qkv_from_regs(p, areg, b, row, row < p.rows, ng);
          }
          continue;
        }
      }
      if (half == 0)
        epilogue_tile<H0>(p, taddr, b, row, row < p.rows, nt * BN);
      else if (nt * BN + H0 < p.N)
        epilogue_tile<H1>(p, taddr + H0, b, row, row < p.rows, nt * BN + H0);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&tempty[acc], 0);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                        // nobody leaves (or frees tensor memory) while the pair is still working
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ CTA pairs, split-N accumulators (256-wide tiles)
// The 256 x 256 pair tile above has ONE accumulator (2 x 256 columns + scale factors do not fit the 512 TMEM columns), and
// tensor memory drains at 64 B/clk: 128 KB per CTA = 2 048 cycles per tile during which the tensor pipe idles (ncu: 72 %
// active vs 86 % for the double-buffered 224-wide tiles).  Here the tile is two 128-column halves in THREE rotating
// 128-column accumulators (384 + 6 x 12 scale-factor columns).  Per K block the leader issues the half-0 MMAs (N = 128) and
// the half-1 MMAs separately; half 0 runs LAG K blocks ahead at the start of a tile and finishes LAG K blocks early at
// its end:
//     h0[0..LAG) | wait buffer of h1 | h1[0..LAG) | (h0[k] h1[k]) ... | h0[last LAG] -> commit | h1[last LAG] -> commit
// The next tile's half 0 starts at once in the spare buffer; its half 1 needs the buffer that this tile's half 0 handed
// to the epilogue 2 x LAG x 256 tensor cycles earlier - more than the ~1 000 cycles its drain takes.  Scale factors are
// kept per pipeline stage in tensor memory (a stage is touched by half 0 first and by half 1 up to LAG K blocks later).
struct Mx8SplitCfg {
  static constexpr int BM = 128, BN = 256, BK = 128, HN = 128;
  static constexpr int STAGES = 6, LAG = 3, NACC = 3;
  static constexpr uint32_t A_BYTES = BM * BK, BH_BYTES = (HN / 2) * BK;          // this CTA's 64 rows of one B half
  static constexpr uint32_t B_BYTES = 2 * BH_BYTES;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES + 512 + 1024;
  static constexpr uint32_t STAGE_STRIDE = (STAGE_BYTES + 1023) / 1024 * 1024;
  static constexpr uint32_t SF_COL = NACC * HN;                                   // 12 columns per stage from here
  static constexpr uint32_t TMEM_COLS = 512;
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_STRIDE + 256 + 1024;
  static_assert(SF_COL + 12 * STAGES <= 512, "TMEM budget");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_mxfp8_2cta_split_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                             const __grid_constant__ CUtensorMap tmSFA, const __grid_constant__ CUtensorMap tmSFB,
                             const GemmParams p) {
  using Cfg = Mx8SplitCfg;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, HN = Cfg::HN, STAGES = Cfg::STAGES, LAG = Cfg::LAG;
  constexpr uint32_t SF_OFF = Cfg::A_BYTES + Cfg::B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_STRIDE);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;          // 3
  uint64_t* tempty = tfull + 3;              // 3
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmSFA);
    ptx::prefetch_tmap(&tmSFB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 3; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 8);          // four epilogue warps (one per lane quadrant) of each CTA drain a half
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 2) ptx::tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m128_per_batch = (p.rows + BM - 1) / BM;
  const int m_per_batch = (p.rows + 2 * BM - 1) / (2 * BM);
  const int num_m = m_per_batch * p.batch;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = p.K / BK;
  constexpr int GROUP_M = 4;
  auto decode = [&](int t, int& mt, int& nt) {
    const int per_group = GROUP_M * num_n;
    const int g = t / per_group;
    const int first = g * GROUP_M;
    const int gsz = min(num_m - first, GROUP_M);
    const int r = t - g * per_group;
    mt = first + r % gsz;
    nt = r / gsz;
  };

  if (warp < 4) {
    ptx::setmaxnreg_dec<72>();
    if (warp_u == 0) {
      // ===================== TMA producer (both CTAs) =====================
      const bool leader_lane = ptx::elect_one();
      const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < num_tiles; t += num_pairs) {
        int mt, nt;
        decode(t, mt, nt);
        const int b = mt / m_per_batch, mp = mt - b * m_per_batch;
        const int mrow = mp * 2 * BM + static_cast<int>(rank) * BM;
        const int nrow = nt * BN + static_cast<int>(rank) * (HN / 2);       // + h * 128 for half h
        const int m128 = min(mp * 2 + static_cast<int>(rank), m128_per_batch - 1);
        const int sfa_row = (b * (p.sfa_mtiles > 0 ? p.sfa_mtiles : m128_per_batch) + m128) * num_k;
        const int sfb_row = nt * 2 * num_k;
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          if (leader_lane) {
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
            const uint32_t sa = smem_u + stage * Cfg::STAGE_STRIDE;
            ptx::tma_load_3d_2cta(sa, &tmA, &full[stage], kb * BK, mrow, b);
            ptx::tma_load_2d_2cta(sa + Cfg::A_BYTES, &tmB, &full[stage], kb * BK, nrow);
            ptx::tma_load_2d_2cta(sa + Cfg::A_BYTES + Cfg::BH_BYTES, &tmB, &full[stage], kb * BK, nrow + HN);
            ptx::tma_load_2d_2cta(sa + SF_OFF, &tmSFA, &full[stage], 0, sfa_row + kb);
            ptx::tma_load_2d_2cta(sa + SF_OFF + 512, &tmSFB, &full[stage], 0, sfb_row + kb);
            ptx::tma_load_2d_2cta(sa + SF_OFF + 1024, &tmSFB, &full[stage], 0, sfb_row + num_k + kb);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    } else if (warp_u == 1 && rank == 0) {
      // ===================== scale-factor copies + MMAs (leader CTA only) =====================
      const bool leader_lane = ptx::elect_one();
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
      int s0 = 0, s1 = 0;                     // pipeline stage cursors of half 0 (waits for the loads) and half 1 (releases)
      uint32_t ph0 = 0;
      const bool lagged = num_k >= 2 * LAG;
      // (macros, not lambdas: by-reference captures put the stage cursors into local memory)
      // half 0 of K block kb: first touch of the stage -> wait for its loads, copy its scale factors to TMEM
#define PA_MX8_H0(kb, d_tmem)                                                                                              \
  do {                                                                                                                     \
    ptx::mbar_wait(&full[s0], ph0);                                                                                        \
    ptx::tc_fence_after();                                                                                                 \
    if (leader_lane) {                                                                                                     \
      const uint32_t sa = smem_u + s0 * Cfg::STAGE_STRIDE;                                                                 \
      const uint32_t sf_t = tmem_u + Cfg::SF_COL + 12 * s0;                                                                \
      ptx::tmem_cp_32x128b_warpx4_2cta(sf_t, make_sf_desc(sa + SF_OFF));                                                   \
      ptx::tmem_cp_32x128b_warpx4_2cta(sf_t + 4, make_sf_desc(sa + SF_OFF + 512));                                         \
      ptx::tmem_cp_32x128b_warpx4_2cta(sf_t + 8, make_sf_desc(sa + SF_OFF + 1024));                                        \
      const uint64_t adesc = ptx::make_desc_kmajor_sw128(sa);                                                              \
      const uint64_t bdesc = ptx::make_desc_kmajor_sw128(sa + Cfg::A_BYTES);                                               \
      _Pragma("unroll") for (int k = 0; k < 4; ++k)                                                                        \
          ptx::mma_mxf8_ss_2cta((d_tmem), adesc + 2 * k, bdesc + 2 * k, ptx::make_idesc_mxf8(2 * BM, HN, 0, 0, k, k), sf_t, \
                                sf_t + 4, ((kb) | k) != 0 ? 1u : 0u);                                                      \
    }                                                                                                                      \
    if (++s0 == STAGES) {                                                                                                  \
      s0 = 0;                                                                                                              \
      ph0 ^= 1;                                                                                                            \
    }                                                                                                                      \
  } while (0)
      // half 1 of K block kb: the stage is resident and its scale factors sit in TMEM; release the stage afterwards
#define PA_MX8_H1(kb, d_tmem)                                                                                              \
  do {                                                                                                                     \
    if (leader_lane) {                                                                                                     \
      const uint32_t sa = smem_u + s1 * Cfg::STAGE_STRIDE;                                                                 \
      const uint32_t sf_t = tmem_u + Cfg::SF_COL + 12 * s1;                                                                \
      const uint64_t adesc = ptx::make_desc_kmajor_sw128(sa);                                                              \
      const uint64_t bdesc = ptx::make_desc_kmajor_sw128(sa + Cfg::A_BYTES + Cfg::BH_BYTES);                               \
      _Pragma("unroll") for (int k = 0; k < 4; ++k)                                                                        \
          ptx::mma_mxf8_ss_2cta((d_tmem), adesc + 2 * k, bdesc + 2 * k, ptx::make_idesc_mxf8(2 * BM, HN, 0, 0, k, k), sf_t, \
                                sf_t + 8, ((kb) | k) != 0 ? 1u : 0u);                                                      \
      ptx::tc_commit_2cta(&empty[s1], 3);                                                                                  \
    }                                                                                                                      \
    if (++s1 == STAGES) s1 = 0;                                                                                            \
  } while (0)
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        const int u0 = 2 * it, u1 = 2 * it + 1;                 // running use index of the accumulator buffers
        const int b0 = u0 % 3, b1 = u1 % 3;
        const uint32_t d0 = tmem_u + b0 * HN, d1 = tmem_u + b1 * HN;
        ptx::mbar_wait(&tempty[b0], ((u0 / 3) & 1) ^ 1);
        ptx::tc_fence_after();
        if (lagged) {
          for (int kb = 0; kb < LAG; ++kb) PA_MX8_H0(kb, d0);
          ptx::mbar_wait(&tempty[b1], ((u1 / 3) & 1) ^ 1);
          ptx::tc_fence_after();
          for (int kb = 0; kb < LAG; ++kb) PA_MX8_H1(kb, d1);
          for (int kb = LAG; kb < num_k - LAG; ++kb) {
            PA_MX8_H0(kb, d0);
            PA_MX8_H1(kb, d1);
          }
          for (int kb = num_k - LAG; kb < num_k; ++kb) PA_MX8_H0(kb, d0);
          if (leader_lane) ptx::tc_commit_2cta(&tfull[b0], 3);
          for (int kb = num_k - LAG; kb < num_k; ++kb) PA_MX8_H1(kb, d1);
          if (leader_lane) ptx::tc_commit_2cta(&tfull[b1], 3);
        } else {
          ptx::mbar_wait(&tempty[b1], ((u1 / 3) & 1) ^ 1);
          ptx::tc_fence_after();
          for (int kb = 0; kb < num_k; ++kb) {
            PA_MX8_H0(kb, d0);
            PA_MX8_H1(kb, d1);
          }
          if (leader_lane) {
            ptx::tc_commit_2cta(&tfull[b0], 3);
            ptx::tc_commit_2cta(&tfull[b1], 3);
          }
        }
      }
#undef PA_MX8_H0
#undef PA_MX8_H1
      __syncwarp();
    }
  } else {
    ptx::setmaxnreg_inc<216>();
    // eight epilogue warps per CTA: warps 4-7 drain half 0, warps 8-11 half 1 (one warp per TMEM lane quadrant each)
    const int q4 = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r_in_tile = q4 * 32 + lane;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      int mt, nt;
      decode(t, mt, nt);
      const int b = mt / m_per_batch;
      const int row = (mt - b * m_per_batch) * 2 * BM + static_cast<int>(rank) * BM + r_in_tile;
      const int u = 2 * it + half;
      const int buf = u % 3;
      ptx::mbar_wait(&tfull[buf], (u / 3) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + buf * HN;
      const int ng = nt * BN + half * HN;
      const bool live = ng < p.N;
      if (p.mode == EPI_QKV_ROPE) {
        uint32_t areg[128];
        if (live) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
            ptx::tmem_ld_32x32b_x32(taddr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&areg[c * 32]));
          ptx::tmem_ld_wait();
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(&tempty[buf], 0);
        if (live) {
          if (p.bias != nullptr && p.rope != nullptr)
            epilogue_qkv_from_regs_fast(p, areg, b, row, row < p.rows, ng);
          else
            epilogue_qkv_from_regs(p, areg, b, row, row < p.rows, ng);
        }
        continue;
      }
      if (live) epilogue_tile<HN>(p, taddr, b, row, row < p.rows, ng);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&tempty[buf], 0);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------ quantiser
// bf16 [batch, rows, K] (strided) -> e4m3 [batch, rows, K] + UE8M0 scales in the chunk layout.
// One thread per 32-element block.  scale = 2^ceil(log2(amax / 448)); rows >= `rows` of the last 128-row
// tile keep scale byte 0 (the caller zero-initialises the SF buffer) and are zero-filled by TMA.
// `tile_rows`: rows per GEMM tile of this operand (128 for A; the B tile width for weights).  Scale chunks
// are grouped per tile: chunk(tile t, j) covers rows [t*tile_rows + 128 j, ... + 128) of that tile.
__global__ void quantize_mxfp8_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long x_bs,
                                      uint8_t* __restrict__ q, uint8_t* __restrict__ sf, int batch, int rows, int K,
                                      int tile_rows) {
  const int kblocks = K >> 5;
  const long long total = static_cast<long long>(batch) * rows * kblocks;
  const int cpt = (tile_rows + 127) >> 7;                          // chunks per tile
  const int m_tiles = ((rows + tile_rows - 1) / tile_rows) * cpt, kchunks = K >> 7;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int kb = static_cast<int>(i % kblocks);
    const long long rr = i / kblocks;
    const int r = static_cast<int>(rr % rows);
    const int b = static_cast<int>(rr / rows);
    const uint4* src = reinterpret_cast<const uint4*>(x + b * x_bs + static_cast<long long>(r) * ldx + kb * 32);
    float v[32];
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 u = __ldg(src + j);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(w[e]);
        v[j * 8 + 2 * e] = f.x;
        v[j * 8 + 2 * e + 1] = f.y;
        amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
      }
    }
    int e = -127;
    if (amax > 0.f) {
      e = static_cast<int>(ceilf(log2f(amax * (1.0f / 448.0f))));
      e = max(-127, min(127, e));
    }
    const float inv = exp2f(static_cast<float>(-e));
    uint32_t packed[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j] * inv, v[4 * j + 1] * inv),
                                                               __NV_SATFINITE, __NV_E4M3);
      const __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j + 2] * inv, v[4 * j + 3] * inv),
                                                               __NV_SATFINITE, __NV_E4M3);
      packed[j] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
    }
    uint4* dst = reinterpret_cast<uint4*>(q + (static_cast<long long>(b) * rows + r) * K + kb * 32);
    dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    const int tile = r / tile_rows, rt = r - tile * tile_rows;
    const int mt = tile * cpt + (rt >> 7), r128 = rt & 127;
    const long long chunk = (static_cast<long long>(b) * m_tiles + mt) * kchunks + (kb >> 2);
    sf[chunk * 512 + (r128 & 31) * 16 + (r128 >> 5) * 4 + (kb & 3)] = static_cast<uint8_t>(e + 127);
  }
}

int quantize_mxfp8_rows(const void* x, long long ldx, long long x_bs, void* q, void* sf, int batch, int rows, int K,
                        int tile_rows, cudaStream_t st) {
  if (K % 128 || ldx % 8 || x_bs % 8 || tile_rows < 128 || tile_rows > 256) return -1;
  const long long total = static_cast<long long>(batch) * rows * (K / 32);
  const int blocks = static_cast<int>(total / 256 + 1 < 148 * 16 ? total / 256 + 1 : 148 * 16);
  quantize_mxfp8_kernel<<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ldx, x_bs,
                                                static_cast<uint8_t*>(q), static_cast<uint8_t*>(sf), batch, rows, K,
                                                tile_rows);
  return (int)cudaGetLastError();
}

template <int BN, int ACC>
static int launch_mx8(const CUtensorMap& ta, const CUtensorMap& tb, const void* sfa, const void* sfb,
                      const GemmParams& p, int tiles, cudaStream_t st) {
  using Cfg = Mx8Cfg<BN, ACC>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mxfp8_kernel<BN, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  const int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_mxfp8_kernel<BN, ACC><<<grid, 384, Cfg::SMEM_BYTES, st>>>(ta, tb, static_cast<const uint8_t*>(sfa),
                                                                 static_cast<const uint8_t*>(sfb), p);
  return (int)cudaGetLastError();
}

template <int BN, int ACC>
static int launch_mx8_pair(const void* A, const void* sfa, const void* W, const void* sfb, const GemmParams& p,
                           long long a_bstride, cudaStream_t st) {
  using Cfg = Mx8PairCfg<BN, ACC>;
  CUtensorMap ta, tb, tsa, tsb;
  const int num_k = p.K / 128, m128 = (p.rows + 127) / 128, num_n = (p.N + BN - 1) / BN;
  {
    uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)p.rows, (uint64_t)p.batch};
    uint64_t str[3] = {1, (uint64_t)p.K, a_bstride > 0 ? (uint64_t)a_bstride : (uint64_t)p.rows * p.K};
    uint32_t box[3] = {128, 128, 1};
    if (make_tmap(&ta, A, 3, dims, str, box, 1, nullptr)) return -20;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[2] = {1, (uint64_t)p.K};
    uint32_t box[2] = {128, (uint32_t)(BN / 2)};
    if (make_tmap(&tb, W, 2, dims, str, box, 1, nullptr)) return -21;
  }
  {   // scale chunks as rows of 128 x u32
    uint64_t dims[2] = {128, (uint64_t)p.batch * (p.sfa_mtiles > 0 ? p.sfa_mtiles : m128) * num_k};
    uint64_t str[2] = {4, 512};
    uint32_t box[2] = {128, 1};
    if (make_tmap(&tsa, sfa, 2, dims, str, box, 4, nullptr, false)) return -22;
    dims[1] = (uint64_t)num_n * Cfg::NCHUNK * num_k;
    if (make_tmap(&tsb, sfb, 2, dims, str, box, 4, nullptr, false)) return -23;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mxfp8_2cta_kernel<BN, ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  const int tiles = ((p.rows + 255) / 256) * p.batch * num_n;
  const int max_pairs = num_sms() / 2;
  const int grid = 2 * (tiles < max_pairs ? tiles : max_pairs);
  gemm_mxfp8_2cta_kernel<BN, ACC><<<grid, 384, Cfg::SMEM_BYTES, st>>>(ta, tb, tsa, tsb, p);
  return (int)cudaGetLastError();
}

static int launch_mx8_split(const void* A, const void* sfa, const void* W, const void* sfb, const GemmParams& p,
                            long long a_bstride, cudaStream_t st) {
  using Cfg = Mx8SplitCfg;
  CUtensorMap ta, tb, tsa, tsb;
  const int num_k = p.K / 128, m128 = (p.rows + 127) / 128, num_n = (p.N + 255) / 256;
  {
    uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)p.rows, (uint64_t)p.batch};
    uint64_t str[3] = {1, (uint64_t)p.K, a_bstride > 0 ? (uint64_t)a_bstride : (uint64_t)p.rows * p.K};
    uint32_t box[3] = {128, 128, 1};
    if (make_tmap(&ta, A, 3, dims, str, box, 1, nullptr)) return -20;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[2] = {1, (uint64_t)p.K};
    uint32_t box[2] = {128, 64};
    if (make_tmap(&tb, W, 2, dims, str, box, 1, nullptr)) return -21;
  }
  {
    uint64_t dims[2] = {128, (uint64_t)p.batch * (p.sfa_mtiles > 0 ? p.sfa_mtiles : m128) * num_k};
    uint64_t str[2] = {4, 512};
    uint32_t box[2] = {128, 1};
    if (make_tmap(&tsa, sfa, 2, dims, str, box, 4, nullptr, false)) return -22;
    dims[1] = (uint64_t)num_n * 2 * num_k;
    if (make_tmap(&tsb, sfb, 2, dims, str, box, 4, nullptr, false)) return -23;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_mxfp8_2cta_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  const int tiles = ((p.rows + 255) / 256) * p.batch * num_n;
  const int max_pairs = num_sms() / 2;
  const int grid = 2 * (tiles < max_pairs ? tiles : max_pairs);
  gemm_mxfp8_2cta_split_kernel<<<grid, 384, Cfg::SMEM_BYTES, st>>>(ta, tb, tsa, tsb, p);
  return (int)cudaGetLastError();
}

static bool mx8_split_default() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("PA_MXFP8_SPLITN");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static bool mx8_pair_default() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("PA_MXFP8_2CTA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// A: e4m3 [batch, rows, K] contiguous (scales: 128-row tiles), W: e4m3 [N, K] contiguous with scales grouped for
// `w_tile` = 224 (generic epilogues, double-buffered accumulators), 256 (fused QKV epilogue) or 128.
// 224- and 256-wide tiles run on CTA pairs (256 x w_tile per pair) when every batch entry has at least 256 rows.
int gemm_mxfp8(const void* A, const void* sfa, const void* W, const void* sfb, GemmParams p, int w_tile,
               cudaStream_t st, int pair, long long a_bstride) {
  if (p.K % 128 || p.N % 32) return -10;
  if (w_tile != 128 && w_tile != 224 && w_tile != 256) return -11;
  if (p.mode == EPI_QKV_ROPE && w_tile == 224) return -12;
  const bool can_pair = w_tile != 128 && p.rows >= 256;
  if (pair >= 1 && !can_pair) return -13;
  if (pair == 2 && w_tile != 256) return -14;
  if (can_pair && (pair >= 1 || (pair < 0 && mx8_pair_default()))) {
    if (w_tile == 224) return launch_mx8_pair<224, 2>(A, sfa, W, sfb, p, a_bstride, st);
    // 256-wide tiles: split-N accumulators (three rotating 128-column buffers) unless the classic single accumulator
    // is asked for (pair == 1 / PA_MXFP8_SPLITN=0)
    if (pair == 2 || (pair < 0 && mx8_split_default())) return launch_mx8_split(A, sfa, W, sfb, p, a_bstride, st);
    return launch_mx8_pair<256, 1>(A, sfa, W, sfb, p, a_bstride, st);
  }
  CUtensorMap ta, tb;
  {
    uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)p.rows, (uint64_t)p.batch};
    uint64_t str[3] = {1, (uint64_t)p.K, a_bstride > 0 ? (uint64_t)a_bstride : (uint64_t)p.rows * p.K};
    uint32_t box[3] = {128, 128, 1};
    if (make_tmap(&ta, A, 3, dims, str, box, 1, nullptr)) return -20;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[2] = {1, (uint64_t)p.K};
    uint32_t box[2] = {128, (uint32_t)w_tile};
    if (make_tmap(&tb, W, 2, dims, str, box, 1, nullptr)) return -21;
  }
  const int tiles = ((p.rows + 127) / 128) * p.batch * ((p.N + w_tile - 1) / w_tile);
  if (w_tile == 224) return launch_mx8<224, 2>(ta, tb, sfa, sfb, p, tiles, st);
  if (w_tile == 256) return launch_mx8<256, 1>(ta, tb, sfa, sfb, p, tiles, st);
  return launch_mx8<128, 2>(ta, tb, sfa, sfb, p, tiles, st);
}

}  // namespace pa
