// Fused SCATTER + first convolution + timestep embedding for latent-diffusion UNets (SURVEY K1/K2,
// north-star fused path #1: "scatter fused with the timestep-embedding add and first conv").
//
// The reference moves a replica's latent shard with a blocking ``.to(dev)`` from a Python thread
// (/root/reference/any_device_parallel.py:1372-1378) and the model then runs ``conv_in`` on it.  Here ONE
// kernel does both: every CTA owns 128 consecutive pixels of one sample of this rank's shard; its four
// producer warps read the 3x3xC neighbourhood of those pixels straight out of the LEAD GPU's NCHW buffer
// (ld.global on an NVLink peer mapping, 2-byte loads coalesced along W, halo = zero padding), write the
// im2col rows as the 128B-swizzled K-major A operand (K = 9*C <= 64), and a tcgen05 GEMM against the
// tap-major packed conv weight produces the NHWC activation rows
//
//     h0[b, y*W + x, :] = sum_{tap,c} x_peer[off + b, c, y+dy, x+dx] * Wc[:, tap*C + c] + bias      (N = 320)
//
// The centre tap doubles as the local NCHW copy of the shard that the Euler/gather epilogue needs later, and
// a spare warp of CTA 0 emits the sinusoidal timestep embeddings of the shard from the lead's timesteps.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../common/host.h"
#include "../common/ptx.cuh"
#include "scatter_params.h"

namespace pa {

namespace sc {
constexpr int BM = 128, BN = 256, BK = 64, WSTAGES = 2;
constexpr uint32_t A_BYTES = BM * BK * 2, W_BYTES = BN * BK * 2;
constexpr uint32_t OFF_W = A_BYTES;
constexpr uint32_t OFF_STAGE = OFF_W + WSTAGES * W_BYTES;      // 4 slabs of [128 rows x 64 cols] bf16, 128B-swizzled
constexpr uint32_t STAGE_BYTES = BM * BN * 2;
constexpr uint32_t OFF_BAR = OFF_STAGE + STAGE_BYTES;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 256 + 1024;
}  // namespace sc

template <int C>
__global__ void __launch_bounds__(256, 1)
scatter_conv_in_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
                       const ScatterConvParams p) {
  using namespace sc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* a_full = bars;              // 1 (128 arrivals)
  uint64_t* w_full = bars + 1;          // WSTAGES
  uint64_t* w_empty = w_full + WSTAGES; // WSTAGES
  uint64_t* tfull = w_empty + WSTAGES;  // 2
  uint64_t* tempty = tfull + 2;         // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
  const int HW = p.H * p.W;
  const int m_per = (HW + BM - 1) / BM;
  const int b = blockIdx.x / m_per;
  const int pix0 = (blockIdx.x - b * m_per) * BM;
  const int num_n = (p.N + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmO);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(a_full, 128);
    for (int s = 0; s < WSTAGES; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&w_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tfull[a], 1);
      ptx::mbar_init(&tempty[a], 4);
    }
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp_u == 0) {
    // packed conv-weight tiles [BN, 64] through a TMA ring (one elected lane issues from warp-uniform code)
    const bool leader = ptx::elect_one();
    const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
    for (int nt = 0; nt < num_n; ++nt) {
      const int s = nt % WSTAGES;
      const uint32_t ph = (nt / WSTAGES) & 1;
      ptx::mbar_wait(&w_empty[s], ph ^ 1);
      if (leader) {
        ptx::mbar_arrive_expect_tx(&w_full[s], W_BYTES);
        ptx::tma_load_2d_s(smem_u + OFF_W + s * W_BYTES, &tmW, &w_full[s], 0, nt * BN);
      }
    }
    __syncwarp();
  } else if (warp_u == 1) {
    constexpr uint32_t IDESC = ptx::make_idesc_f16(BM, BN);
    const bool leader = ptx::elect_one();
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
    const uint32_t smem_u = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
    ptx::mbar_wait(a_full, 0);
    ptx::tc_fence_after();
    const uint64_t adesc = ptx::make_desc_kmajor_sw128(smem_u);
    for (int nt = 0; nt < num_n; ++nt) {
      const int s = nt % WSTAGES, acc = nt & 1;
      ptx::mbar_wait(&w_full[s], (nt / WSTAGES) & 1);
      ptx::mbar_wait(&tempty[acc], ((nt >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      if (leader) {
        const uint64_t bdesc = ptx::make_desc_kmajor_sw128(smem_u + OFF_W + s * W_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::mma_f16_ss(tmem_u + acc * BN, adesc + 2 * k, bdesc + 2 * k, IDESC, k != 0);
        ptx::tc_commit(&w_empty[s]);
        ptx::tc_commit(&tfull[acc]);
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // sinusoidal timestep embeddings of the shard (timesteps read from the lead GPU): [cos | sin]
    if (blockIdx.x == 0) {
      const int half = p.temb_dim >> 1;
      const float lp = -logf(p.max_period) / static_cast<float>(half);
      for (int i = lane; i < p.n * half; i += 32) {
        const int s = i / half, f = i - s * half;
        const float freq = expf(lp * static_cast<float>(f));
        float sn, cs;
        sincosf(__bfloat162float(p.t_src[s]) * p.time_factor * freq, &sn, &cs);
        p.t_emb[s * p.temb_dim + f] = __float2bfloat16(cs);
        p.t_emb[s * p.temb_dim + half + f] = __float2bfloat16(sn);
      }
    }
  } else if (warp >= 4) {
    // ---- A-operand producer: peer loads of the 3x3xC neighbourhood (im2col) -> swizzled smem row
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int pix = pix0 + r;
    uint8_t* arow = smem + r * 128;
    const int sw = r & 7;
    uint32_t kreg[32];                        // 64 bf16 of this pixel's im2col row, k = tap*C + c
#pragma unroll
    for (int i = 0; i < 32; ++i) kreg[i] = 0;
    if (pix < HW) {
      const int y = pix / p.W, x = pix - y * p.W;
      const long long sample = static_cast<long long>(b) * C * HW;
      const unsigned short* xs = reinterpret_cast<const unsigned short*>(p.x_src) + sample;
      unsigned short* xc = p.x_copy ? reinterpret_cast<unsigned short*>(p.x_copy) + sample : nullptr;
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
        const bool in = yy >= 0 && yy < p.H && xx >= 0 && xx < p.W;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int k = tap * C + c;                // compile-time after unrolling: kreg stays in registers
          unsigned short v = 0;
          if (in) v = xs[static_cast<long long>(c) * HW + yy * p.W + xx];
          if (tap == 4 && xc) xc[static_cast<long long>(c) * HW + pix] = v;
          kreg[k >> 1] |= static_cast<uint32_t>(v) << ((k & 1) * 16);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<uint4*>(arow + ((j ^ sw) * 16)) = make_uint4(kreg[4 * j], kreg[4 * j + 1], kreg[4 * j + 2], kreg[4 * j + 3]);
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive(a_full);
    // ---- epilogue: bias -> bf16 -> 128B-swizzled shared-memory staging -> TMA store.  A thread owns one accumulator
    // ROW, so storing straight to global memory makes every warp-level store touch 32 different 128-byte lines (16
    // useful bytes each); staged through shared memory the tile leaves the SM as whole 128-byte rows of a bulk tensor
    // store (UTMASTG), rows past the end of the sample are clipped by the tensor map.
    uint8_t* stage = smem + OFF_STAGE;
    const uint32_t stage_u = ptx::smem_u32(stage);
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(q4 * 32) << 16);
    for (int nt = 0; nt < num_n; ++nt) {
      const int acc = nt & 1;
      ptx::mbar_wait(&tfull[acc], (nt >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n = nt * BN + c * 32;
        if (n >= p.N) break;
        uint32_t t[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + acc * BN + c * 32, t);
        ptx::tmem_ld_wait();
        uint8_t* srow = stage + (c >> 1) * (BM * 128) + r * 128;          // slab of 64 columns, this thread's row
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 bu = __ldg(reinterpret_cast<const uint4*>(p.bias + n + g * 8));
          const uint32_t bw[4] = {bu.x, bu.y, bu.z, bu.w};
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 bb = *reinterpret_cast<const __nv_bfloat162*>(&bw[e]);
            const float2 bf = __bfloat1622float2(bb);
            __nv_bfloat162 ov = __floats2bfloat162_rn(__uint_as_float(t[g * 8 + 2 * e]) + bf.x,
                                                      __uint_as_float(t[g * 8 + 2 * e + 1]) + bf.y);
            o[e] = *reinterpret_cast<uint32_t*>(&ov);
          }
          const int chunk = ((c & 1) * 4 + g) ^ sw;                        // 16-byte chunk inside the 128-byte row
          *reinterpret_cast<uint4*>(srow + chunk * 16) = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      // accumulator drained: hand it back to the MMA warp before the (asynchronous) store
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty[acc]);
      ptx::fence_proxy_async_smem();
      ptx::named_bar_sync(1, 128);                                         // all 128 rows of the tile are staged
      if (warp == 4 && ptx::elect_one()) {
#pragma unroll
        for (int sl = 0; sl < BN / 64; ++sl)
          if (nt * BN + sl * 64 < p.N) ptx::tma_store_3d(&tmO, stage_u + sl * (BM * 128), nt * BN + sl * 64, pix0, b);
        ptx::tma_store_commit();
        ptx::tma_store_wait_read<0>();                                     // staging may be overwritten afterwards
      }
      ptx::named_bar_sync(1, 128);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

int scatter_conv_in(const void* W, long long ldw, const ScatterConvParams& p, cudaStream_t st) {
  using namespace sc;
  if (p.C != 4 || p.N % 32 || p.temb_dim % 2) return -1;      // specialised for 4 latent channels (SD / SDXL)
  CUtensorMap tw;
  uint64_t dims[2] = {64, (uint64_t)p.N};
  uint64_t str[2] = {2, (uint64_t)ldw * 2};
  uint32_t box[2] = {64, (uint32_t)BN};
  if (make_tmap(&tw, W, 2, dims, str, box, 2)) return -20;
  CUtensorMap to;                                   // NHWC output rows for the epilogue's TMA stores
  {
    uint64_t od[3] = {(uint64_t)p.N, (uint64_t)p.H * p.W, (uint64_t)p.n};
    uint64_t os[3] = {2, (uint64_t)p.N * 2, (uint64_t)p.H * p.W * p.N * 2};
    uint32_t ob[3] = {64, (uint32_t)BM, 1};
    if ((p.N % 8) || (reinterpret_cast<uintptr_t>(p.out) & 15)) return -2;
    if (make_tmap(&to, p.out, 3, od, os, ob, 2)) return -21;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(scatter_conv_in_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  const int grid = p.n * ((p.H * p.W + BM - 1) / BM);
  scatter_conv_in_kernel<4><<<grid, 256, SMEM_BYTES, st>>>(tw, to, p);
  return (int)cudaGetLastError();
}

}  // namespace pa
