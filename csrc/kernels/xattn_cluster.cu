// Cross-attention with a SMALL conditioning sequence (SDXL: 77 text tokens, head_dim 64) - the north star's
// "thread-block clusters where conditioning tiles are shared".
//
// The general kernel (attention2.cu) re-loads the K/V tile of a (batch, head) for every 256 query rows.  With Lk <= 128
// the whole conditioning fits ONE tile, so here a CLUSTER of two CTAs serves one (batch, head): each CTA fetches half of
// the K tile and half of the V tile with a multicast TMA load (`.multicast::cluster`, mask 0b11) - the tile crosses
// L2 -> SM once and lands in BOTH CTAs' shared memory - and then each CTA streams its share of the query tiles through a
// two-deep pipeline:  S = Q K^T (tcgen05, S in TMEM)  ->  softmax over the valid keys in registers (one pass, no
// running max: there is only one KV tile)  ->  P back to TMEM (aliasing S)  ->  O = P V  ->  O / l  ->  bf16 rows.
//
//   warp 0  TMA: K/V halves (multicast, once) + Q tile ring      warp 2  TMEM allocator
//   warp 1  tcgen05.mma issuer                                     warps 4-7 / 8-11  softmax + epilogue of even / odd tiles
//   TMEM    S_0 @0, S_1 @128 (fp32, 128 cols; P_x aliases the first 64 columns of S_x), O_0 @256, O_1 @320 (64 cols)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../common/host.h"
#include "../common/ptx.cuh"
#include "softmax_math.cuh"

namespace pa {
namespace xa {

using namespace smx;

constexpr int D = 64;
constexpr uint32_t TILE = 128 * D * 2;                 // 16 KB: 128 rows x 64 bf16, 128-byte swizzled rows
constexpr uint32_t OFF_Q = 0;                          // 2 stages
constexpr uint32_t OFF_K = 2 * TILE;
constexpr uint32_t OFF_V = 3 * TILE;
constexpr uint32_t OFF_BAR = 4 * TILE;
constexpr uint32_t SMEM = OFF_BAR + 256 + 1024;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
xattn_cluster_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, long long ldo,
                     long long o_bstride, int H, int Lq, int Lk, int tiles_per_cluster, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = bars;          // 1
  uint64_t* q_full = bars + 1;       // 2
  uint64_t* q_empty = bars + 3;      // 2
  uint64_t* s_full = bars + 5;       // 2
  uint64_t* p_full = bars + 7;       // 2 (4 warp arrivals)
  uint64_t* o_full = bars + 9;       // 2
  uint64_t* o_empty = bars + 11;     // 2 (4 warp arrivals): S_x / O_x may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const uint32_t rank = ptx::cluster_ctarank();
  const int cluster = blockIdx.x >> 1;
  const int bh = blockIdx.y;
  const int n_q = (Lq + 127) / 128;
  const int t_begin = cluster * tiles_per_cluster;
  const int t_end = min(n_q, t_begin + tiles_per_cluster);
  // query tiles of this CTA: t_begin + rank, + 2, ...
  const int n_local = t_end > t_begin + static_cast<int>(rank) ? (t_end - t_begin - static_cast<int>(rank) + 1) / 2 : 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmQ);
    ptx::prefetch_tmap(&tmK);
    ptx::prefetch_tmap(&tmV);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1);
      ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], 4);
      ptx::mbar_init(&o_full[i], 1);
      ptx::mbar_init(&o_empty[i], 4);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();               // the peer's kv_full barrier exists before any multicast load can signal it
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int hb = bh / H, hh = bh - hb * H;

  if (warp < 4) {
    ptx::setmaxnreg_dec<56>();
    if (warp_u == 0) {
      // ===================== TMA producer =====================
      const bool leader = ptx::elect_one();
      const uint32_t smem_base = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
      if (leader) {
        // conditioning tile shared by the cluster: this CTA fetches key rows [64 rank, 64 rank + 64) of K and V and
        // multicasts them into both CTAs; every CTA's barrier therefore sees the full 2 x 16 KB
        ptx::mbar_arrive_expect_tx(kv_full, 2 * TILE);
        ptx::tma_load_4d_mcast(smem_base + OFF_K + rank * (TILE / 2), &tmK, kv_full, 0, static_cast<int>(rank) * 64, hh, hb, 0x3);
        ptx::tma_load_4d_mcast(smem_base + OFF_V + rank * (TILE / 2), &tmV, kv_full, 0, static_cast<int>(rank) * 64, hh, hb, 0x3);
      }
      for (int i = 0; i < n_local; ++i) {
        const int s = i & 1;
        ptx::mbar_wait(&q_empty[s], ((i >> 1) & 1) ^ 1);
        if (leader) {
          const int t = t_begin + static_cast<int>(rank) + 2 * i;
          ptx::mbar_arrive_expect_tx(&q_full[s], TILE);
          ptx::tma_load_4d_s(smem_base + OFF_Q + s * TILE, &tmQ, &q_full[s], 0, t * 128, hh, hb);
        }
      }
      __syncwarp();
    } else if (warp_u == 1) {
      // ===================== MMA issuer (warp-uniform code, one elected lane issues) =====================
      constexpr uint32_t IDESC_QK = ptx::make_idesc_f16(128, 128, 1, 0, 0);
      constexpr uint32_t IDESC_PV = ptx::make_idesc_f16(128, D, 1, 0, 1);
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t smem_base = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
      const bool leader = ptx::elect_one();
      ptx::mbar_wait(kv_full, 0);
      ptx::tc_fence_after();
      const uint64_t kd = ptx::make_desc_kmajor_sw128(smem_base + OFF_K);
      const uint64_t vd = ptx::make_desc_mnmajor_sw128(smem_base + OFF_V, TILE, 1024);
      auto qk = [&](int i) {                      // S_x = Q_i K^T
        const int x = i & 1;
        ptx::mbar_wait(&q_full[x], (i >> 1) & 1);
        ptx::mbar_wait(&o_empty[x], ((i >> 1) & 1) ^ 1);       // the softmax warps are done with S_x / O_x of tile i - 2
        ptx::tc_fence_after();
        if (leader) {
          const uint64_t qd = ptx::make_desc_kmajor_sw128(smem_base + OFF_Q + x * TILE);
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk)
            ptx::mma_f16_ss(tmem_u + x * 128, qd + 2 * kk, kd + 2 * kk, IDESC_QK, kk != 0);
          ptx::tc_commit(&s_full[x]);
        }
      };
      if (n_local > 0) qk(0);
      for (int i = 0; i < n_local; ++i) {
        const int x = i & 1;
        if (i + 1 < n_local) qk(i + 1);           // the tensor pipe works on the next tile while tile i is in softmax
        ptx::mbar_wait(&p_full[x], (i >> 1) & 1);
        ptx::tc_fence_after();
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            ptx::mma_f16_ts(tmem_u + 256 + x * 64, tmem_u + x * 128 + kk * 8, vd + kk * 128, IDESC_PV, kk != 0);
          ptx::tc_commit(&o_full[x]);
          ptx::tc_commit(&q_empty[x]);
        }
      }
      __syncwarp();
    }
  } else {
    ptx::setmaxnreg_inc<208>();
    // ===================== softmax + epilogue warpgroups (x = 0: even local tiles, 1: odd) =====================
    const int x = (warp - 4) >> 2;
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
    const uint32_t s_addr = lane_addr + x * 128, o_addr = lane_addr + 256 + x * 64;
    const int b = hb, h = hh;
    for (int i = x; i < n_local; i += 2) {
      const uint32_t ph = (i >> 1) & 1;
      ptx::mbar_wait(&s_full[x], ph);
      ptx::tc_fence_after();
      uint32_t sv[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld_32x32b_x32(s_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 128; ++j)
        if (j >= Lk) sv[j] = 0xff800000u;          // keys past the conditioning length (the K rows there are TMA zero fill)
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 128; j += 4) {
        mx0 = fmax3(mx0, __uint_as_float(sv[j]), __uint_as_float(sv[j + 1]));
        mx1 = fmax3(mx1, __uint_as_float(sv[j + 2]), __uint_as_float(sv[j + 3]));
      }
      const float m = fmaxf(mx0, mx1);
      const float mneg = -m * scale_log2;
      float l = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
        if (c * 32 >= Lk) {                                        // chunk entirely past the conditioning length: P = 0
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = 0u;
          ptx::tmem_st_32x32b_x16(s_addr + c * 16, pk);
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float a0 = ex2f(fmaf(__uint_as_float(sv[c * 32 + j]), scale_log2, mneg));
          const float a1 = ex2f(fmaf(__uint_as_float(sv[c * 32 + j + 1]), scale_log2, mneg));
          l += a0 + a1;
          __nv_bfloat162 hv = __floats2bfloat162_rn(a0, a1);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hv);
        }
        ptx::tmem_st_32x32b_x16(s_addr + c * 16, pk);          // P aliases the first 64 columns of S_x
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[x]);
      // ---- O / l -> bf16 rows
      ptx::mbar_wait(&o_full[x], ph);
      ptx::tc_fence_after();
      const int t = t_begin + static_cast<int>(rank) + 2 * i;
      const int q_row = t * 128 + r;
      const float inv = 1.0f / l;
      __nv_bfloat16* dst = out + b * o_bstride + static_cast<long long>(q_row) * ldo + h * D;
#pragma unroll
      for (int c = 0; c < D / 32; ++c) {
        uint32_t tv[32];
        ptx::tmem_ld_32x32b_x32(o_addr + c * 32, tv);
        ptx::tmem_ld_wait();
        if (q_row < Lq) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 u;
            __nv_bfloat162 a0 = __floats2bfloat162_rn(__uint_as_float(tv[j]) * inv, __uint_as_float(tv[j + 1]) * inv);
            __nv_bfloat162 a1 = __floats2bfloat162_rn(__uint_as_float(tv[j + 2]) * inv, __uint_as_float(tv[j + 3]) * inv);
            __nv_bfloat162 a2 = __floats2bfloat162_rn(__uint_as_float(tv[j + 4]) * inv, __uint_as_float(tv[j + 5]) * inv);
            __nv_bfloat162 a3 = __floats2bfloat162_rn(__uint_as_float(tv[j + 6]) * inv, __uint_as_float(tv[j + 7]) * inv);
            u.x = *reinterpret_cast<uint32_t*>(&a0);
            u.y = *reinterpret_cast<uint32_t*>(&a1);
            u.z = *reinterpret_cast<uint32_t*>(&a2);
            u.w = *reinterpret_cast<uint32_t*>(&a3);
            *reinterpret_cast<uint4*>(dst + c * 32 + j) = u;
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&o_empty[x]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync();               // neither CTA's shared memory goes away while the peer may still multicast into it
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

}  // namespace xa

// q [B, H, Lq, 64], k / v [B, H, Lk <= 128, 64] strided views (element strides of dims 0..2) -> out[b, l, h*64 + d]
int xattn_cluster_bf16(const void* q, const void* k, const void* v, void* out, long long ldo, long long o_bstride, int B,
                       int H, int Lq, int Lk, int D, const long long* qs, const long long* ks, const long long* vs, float scale,
                       cudaStream_t st) {
  using namespace xa;
  if (D != xa::D || Lk < 1 || Lk > 128) return -11;
  for (int i = 0; i < 3; ++i)
    if (qs[i] % 8 || ks[i] % 8 || vs[i] % 8) return -10;
  CUtensorMap tq, tk, tv;
  auto mk = [&](CUtensorMap* m, const void* p, int L, const long long* s3, uint32_t rows) {
    uint64_t dims[4] = {(uint64_t)xa::D, (uint64_t)L, (uint64_t)H, (uint64_t)B};
    uint64_t str[4] = {2, (uint64_t)s3[2] * 2, (uint64_t)s3[1] * 2, (uint64_t)s3[0] * 2};
    const uint32_t box[4] = {64, rows, 1, 1};
    return make_tmap(m, p, 4, dims, str, box, 2, nullptr);
  };
  if (mk(&tq, q, Lq, qs, 128)) return -20;
  if (mk(&tk, k, Lk, ks, 64)) return -21;            // half tiles: each CTA of the pair multicasts 64 key rows
  if (mk(&tv, v, Lk, vs, 64)) return -22;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(xattn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  const int n_q = (Lq + 127) / 128;
  // clusters per (batch, head): enough CTAs to fill the machine, at least ~4 query tiles per CTA to amortise the K/V fetch
  int clusters = 1;
  while (clusters * 2 <= (n_q + 7) / 8 && static_cast<long long>(B) * H * clusters * 2 < 2LL * num_sms()) clusters *= 2;
  const int tiles_per_cluster = (n_q + clusters - 1) / clusters;
  dim3 grid(2 * clusters, B * H);
  xattn_cluster_kernel<<<grid, 384, SMEM, st>>>(tq, tk, tv, static_cast<__nv_bfloat16*>(out), ldo, o_bstride, H, Lq, Lk,
                                                 tiles_per_cluster, scale * 1.4426950408889634f);
  return (int)cudaGetLastError();
}

}  // namespace pa
