// CTA-pair (cta_group::2) variant of the bf16 GEMM: a cluster of two CTAs on neighbouring SMs computes one
// 256 x 256 output tile per step.  Each CTA loads ITS 128 rows of A and ITS 128 of the 256 B rows (W rows) per
// K-block - 32 KB instead of the 48 KB a lone 128x256 CTA needs - and the even CTA issues one
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) that reads both shared memories and writes 128 accumulator
// rows into each CTA's tensor memory.  Per output element that is 1/3 less shared-memory fill + L2 traffic and half
// the B-operand reads per SM: the step is power-capped (sw_power_cap), so data movement saved is clock gained.
//
//   warp 0   TMA producer (both CTAs; transaction bytes of both land on the LEADER's `full` barrier)
//   warp 1   MMA issuer (leader CTA only); commits multicast to `empty` / `tfull` of both CTAs
//   warp 2   tensor-memory allocator (cta_group::2 form, same warp id in both CTAs)
//   warps 4-11 epilogue (each CTA drains its own 128 x 256 accumulator half, two warps per TMEM lane quadrant);
//              `tempty` lives in the leader and counts the 16 epilogue warps of the pair (peer arrives remotely)
#pragma once

#include "gemm_tcgen05.cuh"

namespace pa {

struct Gemm2CtaCfg {
  static constexpr int BM = 128, BN = 256, BK = 64;          // per CTA: 128 rows of the pair's 256 x 256 tile
  static constexpr int STAGES = 6;
  static constexpr uint32_t A_BYTES = BM * BK * 2;           // 16 KB
  static constexpr uint32_t B_BYTES = (BN / 2) * BK * 2;     // this CTA's half of the B tile: 16 KB
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr uint32_t TMEM_COLS = 512;                 // two 256-column accumulator stages
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
  static constexpr int THREADS = 384;                        // 4 control warps + 8 epilogue warps
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const GemmParams p) {
  using Cfg = Gemm2CtaCfg;
  constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, STAGES = Cfg::STAGES;
  constexpr uint32_t IDESC = ptx::make_idesc_f16(2 * BM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();              // 0: leader (issues the MMAs), 1: peer
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

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
      ptx::mbar_init(&tempty[a], 16);         // eight epilogue warps of each CTA of the pair
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 2) ptx::tmem_alloc_2cta<Cfg::TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();                        // both CTAs' barriers are initialised before any remote arrive / TMA
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int m_per_batch = (p.rows + 2 * BM - 1) / (2 * BM);          // 256-row tiles per batch entry
  const int num_m = m_per_batch * p.batch;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int num_k = (p.K + BK - 1) / BK;
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

  if (warp_u == 0) {
    // ===================== TMA producer (both CTAs) =====================
    const bool leader_lane = ptx::elect_one();
    const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < num_tiles; t += num_pairs) {
      int mt, nt;
      decode(t, mt, nt);
      const int b = mt / m_per_batch;
      const int mrow = (mt - b * m_per_batch) * 2 * BM + static_cast<int>(rank) * BM;
      const int nrow = nt * BN + static_cast<int>(rank) * (BN / 2);
      for (int kb = 0; kb < num_k; ++kb) {
        ptx::mbar_wait(&empty[stage], phase ^ 1);
        if (leader_lane) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);   // bytes of BOTH CTAs
          const uint32_t sa = smem_u + stage * Cfg::STAGE_BYTES;
          ptx::tma_load_3d_2cta(sa, &tmA, &full[stage], kb * BK, mrow, b);
          ptx::tma_load_2d_2cta(sa + Cfg::A_BYTES, &tmB, &full[stage], kb * BK, nrow);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp_u == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
      const bool leader_lane = ptx::elect_one();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ptx::mbar_wait(&tempty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_u + acc * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          if (leader_lane) {
            const uint32_t sa = smem_u + stage * Cfg::STAGE_BYTES;
            const uint64_t adesc = ptx::make_desc_kmajor_sw128(sa);
            const uint64_t bdesc = ptx::make_desc_kmajor_sw128(sa + Cfg::A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              ptx::mma_f16_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0 ? 1u : 0u);
            ptx::tc_commit_2cta(&empty[stage], 3);                     // frees this stage in both CTAs
            if (kb == num_k - 1) ptx::tc_commit_2cta(&tfull[acc], 3);  // accumulator halves ready in both CTAs
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== epilogue: this CTA's 128 rows of the pair's tile =====================
    // Eight warps: two per TMEM lane quadrant, each draining 128 of the 256 accumulator columns.  The fused QKV
    // epilogue (RMSNorm + RoPE per 128-wide head) and the erf-GELU epilogue cost ~15 instructions per element; with
    // four warps they took about as long as the 48-K-block main loop of the next tile (85 % of peak vs 97 % for
    // the plain epilogues), with eight they hide under it.
    const int q4 = warp & 3;
    const int half = (warp - 4) >> 2;
    const int r_in_tile = q4 * 32 + lane;
    int it = 0;
    for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
      int mt, nt;
      decode(t, mt, nt);
      const int b = mt / m_per_batch;
      const int row = (mt - b * m_per_batch) * 2 * BM + static_cast<int>(rank) * BM + r_in_tile;
      const int acc = it & 1;
      ptx::mbar_wait(&tfull[acc], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + acc * BN + half * (BN / 2);
      if (nt * BN + half * (BN / 2) < p.N) epilogue_tile<BN / 2>(p, taddr, b, row, row < p.rows, nt * BN + half * (BN / 2));
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&tempty[acc], 0);        // the leader's MMA warp waits for all 8 warps
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

}  // namespace pa
