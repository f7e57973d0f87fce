// Attention, CTA-pair ping-pong variant (head dim 128): a cluster of two CTAs shares every K/V tile.
//
// attention2.cu left two losses on the table: Q.K^T ran at ~85 instead of 64 cycles per MMA because
// its two shared-memory operands (4 KB + 4 KB per 128x128x16 MMA) saturate the 128 B/clk shared-memory port that the
// TMA fills also use, and every CTA pulled the full K/V stream from L2.  Here the pair issues M = 256 MMAs
// (tcgen05.mma.cta_group::2): each CTA contributes its own 128 query rows and HALF of the B operand -
//   S_X[256 x 128 keys] = [Q_X(cta0); Q_X(cta1)] . K_j^T      each CTA holds 64 of the 128 keys  (4 + 2 KB per MMA)
//   O_X[256 x 128 d]   += [P_X(cta0); P_X(cta1)] . V_j        each CTA holds 64 of the 128 d-columns
// so the shared-memory traffic per MMA drops to 96 B/clk, the K/V fill per CTA halves, and four 128-row query tiles
// (two ping-pong tiles A / B per CTA) reuse each K/V tile.  The halved K/V stages also leave room for P in SHARED
// memory (bf16, swizzled like Q): S_X in tensor memory is free again as soon as the softmax warps hold it in
// registers, so Q.K^T of tile j+1 runs under the softmax of tile j and the per-tile dependency loop
// P.V(j), Q.K^T(j+1) -> softmax(j+1) that bounded attention2.cu disappears.
//
//   warp 0    TMA producer (both CTAs; bytes of both land on the LEADER's full barriers)
//   warp 1    MMA issuer (even CTA only); commits multicast to s_full / o_full / k_empty / v_empty of both CTAs
//   warp 2    TMEM allocator (cta_group::2)              warps 4-7 / 8-11  softmax of tile A / B (local TMEM)
//   p_full / p_hi / s_free live in the leader and count the 4 softmax warps of BOTH CTAs (the peer arrives remotely).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "../common/host.h"
#include "../common/ptx.cuh"
#include "softmax_math.cuh"

namespace pa {
namespace a3 {

using namespace smx;

// timeline capture (TRACE = true instantiation only): [role 0..6][kv tile 0..63][slot 0..7] clock64 stamps of the
// cluster (blockIdx.x = 2, 3; blockIdx.y = 1).  roles: 0/1 MMA thread (tile A / B), 2/3 leader softmax A / B,
// 4 leader producer, 5/6 peer softmax A / B.  Clocks of the two SMs are not synchronised: compare within a CTA.
__device__ long long g_trace[7 * 64 * 8];
#define PA_TR3(role, j, slot)                                                                     \
  do {                                                                                            \
    if (TRACE && (blockIdx.x >> 1) == 1 && blockIdx.y == 1 && (j) < 64)                           \
      g_trace[((role) * 64 + (j)) * 8 + (slot)] = clock64();                                      \
  } while (0)

constexpr int D = 128, BN = 128, KV_STAGES = 2;
constexpr uint32_t Q_TILE = 128 * D * 2;            // 32 KB: one 128-row query tile (two 64-column slices)
constexpr uint32_t Q_SLICE = 128 * 64 * 2;
constexpr uint32_t K_HALF = 64 * D * 2;             // 16 KB: this CTA's 64 keys x 128 d (two 64-column slices of 8 KB)
constexpr uint32_t K_SLICE = 64 * 64 * 2;
constexpr uint32_t V_HALF = BN * 64 * 2;            // 16 KB: 128 keys x this CTA's 64 d-columns
constexpr uint32_t OFF_Q = 0;
constexpr uint32_t OFF_K = 2 * Q_TILE;
constexpr uint32_t OFF_V = OFF_K + KV_STAGES * K_HALF;
constexpr uint32_t OFF_P = OFF_V + KV_STAGES * V_HALF;    // P_A, P_B: bf16 [128 q x 128 keys], K-major SW128 (as Q)
constexpr uint32_t OFF_BAR = OFF_P + 2 * Q_TILE;
constexpr uint32_t SMEM = OFF_BAR + 512 + 1024;
constexpr int POLY_MASK = 0x22;                     // 2 of every 8 exp2 pairs on the FMA pipe (see attention2.cu)

template <bool TRACE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
attention3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, long long ldo,
                  long long o_bstride, int H, int Lq, int Lk, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars;                     // 1   (leader)
  uint64_t* k_full = bars + 1;                 // KV_STAGES (leader)
  uint64_t* k_empty = k_full + KV_STAGES;      // KV_STAGES (each CTA, multicast commit)
  uint64_t* v_full = k_empty + KV_STAGES;
  uint64_t* v_empty = v_full + KV_STAGES;
  uint64_t* s_full = v_empty + KV_STAGES;      // 2 (tile A, B; each CTA)
  uint64_t* o_full = s_full + 2;               // 2 (each CTA)
  uint64_t* p_full = o_full + 2;               // 2 (leader, 8 arrivals): keys 0..63 of P_X stored in both CTAs
  uint64_t* p_hi = p_full + 2;                 // 2 (leader, 8 arrivals): keys 64..127
  uint64_t* s_free = p_hi + 2;                 // 2 (leader, 8 arrivals): S_X is in registers in both CTAs
  uint64_t* p_empty = s_free + 2;              // 2 (each CTA): P_X.V has consumed the P_X buffer
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t rank = ptx::cluster_ctarank();
  const int q0 = blockIdx.x * 256;
  const int bh = blockIdx.y;
  const int n_kv = (Lk + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&o_full[i], 1);
      ptx::mbar_init(&p_full[i], 8);
      ptx::mbar_init(&p_hi[i], 8);
      ptx::mbar_init(&s_free[i], 8);
      ptx::mbar_init(&p_empty[i], 1);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 2) ptx::tmem_alloc_2cta<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    ptx::setmaxnreg_dec<56>();
    if (warp_u == 0) {
      // ===================== TMA producer (both CTAs) =====================
      const bool leader = ptx::elect_one();
      const uint32_t sb = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
      const int hb = bh / H, hh = bh - hb * H;
      if (leader) {
        if (rank == 0) ptx::mbar_arrive_expect_tx(q_full, 4 * Q_TILE);          // two query tiles from each CTA
#pragma unroll
        for (int t = 0; t < 2; ++t)
#pragma unroll
          for (int sl = 0; sl < 2; ++sl)
            ptx::tma_load_4d_2cta(sb + OFF_Q + t * Q_TILE + sl * Q_SLICE, &tmQ, q_full, sl * 64, q0 + t * 128, hh, hb);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % KV_STAGES;
        const uint32_t ph = (j / KV_STAGES) & 1;
        ptx::mbar_wait(&k_empty[s], ph ^ 1);
        if (leader) {
          if (rank == 0) PA_TR3(4, j, 0);
          if (rank == 0) ptx::mbar_arrive_expect_tx(&k_full[s], 2 * K_HALF);
#pragma unroll
          for (int sl = 0; sl < 2; ++sl)                                        // this CTA's 64 keys of tile j
            ptx::tma_load_4d_2cta(sb + OFF_K + s * K_HALF + sl * K_SLICE, &tmK, &k_full[s], sl * 64,
                                  j * BN + static_cast<int>(rank) * 64, hh, hb);
        }
        ptx::mbar_wait(&v_empty[s], ph ^ 1);
        if (leader) {
          if (rank == 0) PA_TR3(4, j, 1);
          if (rank == 0) ptx::mbar_arrive_expect_tx(&v_full[s], 2 * V_HALF);
          ptx::tma_load_4d_2cta(sb + OFF_V + s * V_HALF, &tmV, &v_full[s], static_cast<int>(rank) * 64, j * BN, hh,
                                hb);                                            // all 128 keys, this CTA's 64 d-columns
        }
      }
      __syncwarp();
    } else if (warp_u == 1 && rank == 0) {
      // ===================== MMA issuer (leader CTA) =====================
      constexpr uint32_t IDESC_QK = ptx::make_idesc_f16(256, 128, 1, 0, 0);
      constexpr uint32_t IDESC_PV = ptx::make_idesc_f16(256, 128, 1, 0, 1);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t sb = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
      const bool leader = ptx::elect_one();
      auto qk = [&](int x, int i) {            // S_X = Q_X K_i^T over the pair
        if (leader) {
          const uint64_t qd = ptx::make_desc_kmajor_sw128(sb + OFF_Q + x * Q_TILE);
          const uint64_t kd = ptx::make_desc_kmajor_sw128(sb + OFF_K + (i % KV_STAGES) * K_HALF);
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) {
            const uint32_t qoff = ((kk >> 2) * Q_SLICE + (kk & 3) * 32) >> 4;
            const uint32_t koff = ((kk >> 2) * K_SLICE + (kk & 3) * 32) >> 4;
            ptx::mma_f16_ss_2cta(tmem_u + x * 128, qd + qoff, kd + koff, IDESC_QK, kk != 0);
          }
          ptx::tc_commit_2cta(&s_full[x], 3);
        }
      };
      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      qk(0, 0);
      qk(1, 0);
      if (leader) ptx::tc_commit_2cta(&k_empty[0], 3);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % KV_STAGES;
        // S_X(j+1) as soon as S_X(j) sits in registers: P lives in shared memory, so nothing else keeps S_X busy and
        // the softmax of tile j+1 never waits for P.V(j) - the per-tile dependency loop of attention2.cu is gone.
        if (j + 1 < n_kv) {
          ptx::mbar_wait(&k_full[(j + 1) % KV_STAGES], ((j + 1) / KV_STAGES) & 1);
          if (leader) PA_TR3(0, j, 0);
#pragma unroll 1
          for (int x = 0; x < 2; ++x) {
            ptx::mbar_wait(&s_free[x], j & 1);
            if (leader) PA_TR3(x, j, 1);
            ptx::tc_fence_after();
            qk(x, j + 1);
            if (leader) PA_TR3(x, j, 2);
          }
          if (leader) ptx::tc_commit_2cta(&k_empty[(j + 1) % KV_STAGES], 3);
        }
        ptx::mbar_wait(&v_full[st], (j / KV_STAGES) & 1);
        if (leader) PA_TR3(1, j, 0);
        // V half of this CTA: 128 keys x 64 d, MN-major; one 64-wide MN group per CTA, 8-key groups 1024 B apart
        const uint64_t vd = ptx::make_desc_mnmajor_sw128(sb + OFF_V + st * V_HALF, V_HALF, 1024);
#pragma unroll 1
        for (int x = 0; x < 2; ++x) {
          const uint64_t pd = ptx::make_desc_kmajor_sw128(sb + OFF_P + x * Q_TILE);
          ptx::mbar_wait(&p_full[x], j & 1);
          ptx::tc_fence_after();
          if (leader) {
            PA_TR3(x, j, 3);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              ptx::mma_f16_ss_2cta(tmem_u + 256 + x * 128, pd + ((kk * 32) >> 4), vd + kk * 128, IDESC_PV, (j | kk) != 0);
          }
          ptx::mbar_wait(&p_hi[x], j & 1);
          ptx::tc_fence_after();
          if (leader) {
            PA_TR3(x, j, 4);
#pragma unroll
            for (int kk = 4; kk < 8; ++kk)
              ptx::mma_f16_ss_2cta(tmem_u + 256 + x * 128, pd + ((Q_SLICE + (kk & 3) * 32) >> 4), vd + kk * 128, IDESC_PV,
                                   1u);
            ptx::tc_commit_2cta(&o_full[x], 3);
            ptx::tc_commit_2cta(&p_empty[x], 3);
            if (x == 1) ptx::tc_commit_2cta(&v_empty[st], 3);
            PA_TR3(x, j, 5);
          }
        }
      }
      __syncwarp();
    }
  } else {
    ptx::setmaxnreg_inc<208>();
    // ===================== softmax warpgroups (this CTA's rows; identical to attention2.cu, NS = 1) =====================
    const int x = (warp - 4) >> 2;                      // query tile of this warpgroup
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
    const uint32_t s_addr = lane_addr + x * 128;
    const uint32_t o_addr = lane_addr + 256 + x * 128;
    float m_used = -INFINITY, l = 0.f;
    const unsigned long long sl2 = pack2(scale_log2, scale_log2);
    uint8_t* p_row = smem + OFF_P + x * Q_TILE + r * 128;
    const bool tracer = (q4 == 0 && lane == 0);
    const int trole = (rank == 0 ? 2 : 5) + x;

    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait(&s_full[x], j & 1);
      if (tracer) PA_TR3(trole, j, 0);
      ptx::tc_fence_after();
      uint32_t sv[128];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        ptx::tmem_ld_32x32b_x32(s_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&s_free[x], 0);      // S_X may be overwritten by Q.K^T of tile j+1
      if (tracer) PA_TR3(trole, j, 1);
      const int kv_left = Lk - j * BN;
      if (kv_left < BN) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= kv_left) sv[i] = 0xff800000u;
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 8) {
        mx0 = fmax3(mx0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
        mx1 = fmax3(mx1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
        mx2 = fmax3(mx2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
        mx3 = fmax3(mx3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
      }
      const float m_new = fmaxf(fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)), m_used);
      const bool need = (m_new - m_used) * scale_log2 > 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2f((m_used - m_new) * scale_log2) : 1.0f;
        if (need) m_used = m_new;
        l *= alpha;
        if (j > 0) {
          ptx::mbar_wait(&o_full[x], (j - 1) & 1);
          ptx::tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < D / 32; ++c) {
            uint32_t t[32];
            ptx::tmem_ld_32x32b_x32(o_addr + c * 32, t);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            ptx::tmem_st_32x32b_x32(o_addr + c * 32, t);
          }
          ptx::tmem_st_wait();
        }
      }
      const float mneg_f = -m_used * scale_log2;
      const unsigned long long mneg = pack2(mneg_f, mneg_f);
      unsigned long long sum2 = pack2(0.f, 0.f);
      if (tracer) PA_TR3(trole, j, 2);
      ptx::mbar_wait(&p_empty[x], (j & 1) ^ 1);                    // P_X.V of tile j-1 has read the P_X buffer
      if (tracer) PA_TR3(trole, j, 3);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float a0, a1;
          const unsigned long long x2 =
              fma2(pack2(__uint_as_float(sv[c * 32 + i]), __uint_as_float(sv[c * 32 + i + 1])), sl2, mneg);
          if ((POLY_MASK >> ((i >> 1) & 7)) & 1) {
            exp2_poly2(x2, a0, a1);
          } else {
            unpack2(x2, a0, a1);
            a0 = ex2f(a0);
            a1 = ex2f(a1);
          }
          sum2 = add2(sum2, pack2(a0, a1));
          __nv_bfloat162 hv = __floats2bfloat162_rn(a0, a1);
          pk[i >> 1] = *reinterpret_cast<uint32_t*>(&hv);
        }
        // P_X row r, keys 32c..32c+31 -> K-major SW128 operand tile: slice c/2, 16-byte chunks (c&1)*4 .. +3 of the row
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int ci = (c & 1) * 4 + q;
          *reinterpret_cast<uint4*>(p_row + (c >> 1) * Q_SLICE + ((ci ^ (r & 7)) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        if (c == 1) {                                     // keys 0..63 complete: the first four P.V MMAs may start
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(&p_full[x], 0);
          if (tracer) PA_TR3(trole, j, 4);
        }
      }
      float s0, s1;
      unpack2(sum2, s0, s1);
      l += s0 + s1;
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&p_hi[x], 0);
      if (tracer) PA_TR3(trole, j, 5);
    }

    ptx::mbar_wait(&o_full[x], (n_kv - 1) & 1);
    ptx::tc_fence_after();
    const int q_row = q0 + x * 128 + r;
    const float inv = 1.0f / l;
    const int b = bh / H, h = bh - b * H;
    __nv_bfloat16* dst = out + b * o_bstride + static_cast<long long>(q_row) * ldo + h * D;
#pragma unroll 1
    for (int c = 0; c < D / 32; ++c) {
      uint32_t t[32];
      ptx::tmem_ld_32x32b_x32(o_addr + c * 32, t);
      ptx::tmem_ld_wait();
      if (q_row < Lq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 u;
          __nv_bfloat162 a0 = __floats2bfloat162_rn(__uint_as_float(t[i]) * inv, __uint_as_float(t[i + 1]) * inv);
          __nv_bfloat162 a1 = __floats2bfloat162_rn(__uint_as_float(t[i + 2]) * inv, __uint_as_float(t[i + 3]) * inv);
          __nv_bfloat162 a2 = __floats2bfloat162_rn(__uint_as_float(t[i + 4]) * inv, __uint_as_float(t[i + 5]) * inv);
          __nv_bfloat162 a3 = __floats2bfloat162_rn(__uint_as_float(t[i + 6]) * inv, __uint_as_float(t[i + 7]) * inv);
          u.x = *reinterpret_cast<uint32_t*>(&a0);
          u.y = *reinterpret_cast<uint32_t*>(&a1);
          u.z = *reinterpret_cast<uint32_t*>(&a2);
          u.w = *reinterpret_cast<uint32_t*>(&a3);
          *reinterpret_cast<uint4*>(dst + c * 32 + i) = u;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2cta<512>(tmem);
  }
}

}  // namespace a3

// same contract as attention2_bf16; head dim 128 only, the query length is covered by pairs of 256-row CTAs
int attention3_trace_read(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, a3::g_trace, sizeof(long long) * 7 * 64 * 8);
}

static bool a3_trace_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("PA_ATTN3_TRACE");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int attention3_bf16(const void* q, const void* k, const void* v, void* out, long long ldo, long long o_bstride, int B,
                    int H, int Lq, int Lk, int D, const long long* qs, const long long* ks, const long long* vs,
                    float scale, cudaStream_t st) {
  if (D != 128) return -11;
  for (int i = 0; i < 3; ++i)
    if (qs[i] % 8 || ks[i] % 8 || vs[i] % 8) return -10;
  CUtensorMap tq, tk, tv;
  auto mk = [&](CUtensorMap* m, const void* p, int L, const long long* s3, uint32_t b0, uint32_t b1) {
    uint64_t dims[4] = {(uint64_t)D, (uint64_t)L, (uint64_t)H, (uint64_t)B};
    uint64_t str[4] = {2, (uint64_t)s3[2] * 2, (uint64_t)s3[1] * 2, (uint64_t)s3[0] * 2};
    const uint32_t box[4] = {b0, b1, 1, 1};
    return make_tmap(m, p, 4, dims, str, box, 2, nullptr);
  };
  if (mk(&tq, q, Lq, qs, 64, 128)) return -20;
  if (mk(&tk, k, Lk, ks, 64, 64)) return -21;          // 64 keys per CTA
  if (mk(&tv, v, Lk, vs, 64, 128)) return -22;         // 128 keys x 64 d-columns per CTA
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(a3::attention3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a3::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(a3::attention3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a3::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  const int blocks_x = ((Lq + 255) / 256 + 1) & ~1;    // whole pairs; a padding CTA only computes masked rows
  dim3 grid(blocks_x, B * H);
  if (a3_trace_enabled())
    a3::attention3_kernel<true><<<grid, 384, a3::SMEM, st>>>(tq, tk, tv, static_cast<__nv_bfloat16*>(out), ldo, o_bstride, H,
                                                           Lq, Lk, scale * 1.4426950408889634f);
  else
    a3::attention3_kernel<false><<<grid, 384, a3::SMEM, st>>>(tq, tk, tv, static_cast<__nv_bfloat16*>(out), ldo, o_bstride,
                                                            H, Lq, Lk, scale * 1.4426950408889634f);
  return (int)cudaGetLastError();
}

}  // namespace pa
