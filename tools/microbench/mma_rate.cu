// tcgen05.mma issue/execute rate for the operand forms the attention kernel uses (one CTA per SM, one elected
// lane issues REP x 8 MMAs back to back, one commit, wait).  Data is whatever shared/tensor memory holds.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../csrc -o mma_rate.bin mma_rate.cu && ./mma_rate.bin
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda.h>
#include "common/ptx.cuh"

using namespace pa;

// FORM 0: SS 128x128x16, A and B K-major (Q.K^T)     1: TS 128x128x16, A in TMEM, B MN-major (P.V)
//      2: SS 128x256x16 (GEMM tile)                  3: SS 128x128x16, B MN-major
//      4: SS 128x64x16
template <int FORM>
__global__ void __launch_bounds__(128, 1) k(long long* cyc, int rep) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
    ptx::fence_proxy_async_smem();
  }
  if (warp == 0) ptx::tmem_alloc<512>(&tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 1) {
    const bool leader = ptx::elect_one();
    const uint32_t sb = __shfl_sync(0xffffffffu, ptx::smem_u32(smem), 0);
    constexpr int N = FORM == 2 ? 256 : (FORM == 4 ? 64 : 128);
    constexpr uint32_t IDESC = ptx::make_idesc_f16(128, N, 1, 0, (FORM == 1 || FORM == 3) ? 1 : 0);
    const uint64_t ad = ptx::make_desc_kmajor_sw128(sb);
    const uint64_t bd = (FORM == 1 || FORM == 3) ? ptx::make_desc_mnmajor_sw128(sb + 65536, 16384, 1024)
                                                 : ptx::make_desc_kmajor_sw128(sb + 65536);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (leader) {
      t0 = clock64();
      for (int r = 0; r < rep; ++r) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = ((kk >> 2) * 16384 + (kk & 3) * 32) >> 4;
          if (FORM == 1)
            ptx::mma_f16_ts(tmem + 256, tmem + kk * 8, bd + kk * 128, IDESC, 1u);
          else if (FORM == 3)
            ptx::mma_f16_ss(tmem + 256, ad + off, bd + kk * 128, IDESC, 1u);
          else
            ptx::mma_f16_ss(tmem + (r & 1) * 256, ad + off, bd + off, IDESC, 1u);
        }
      }
      t1 = clock64();
      ptx::tc_commit(&bar);
    }
    ptx::mbar_wait(&bar, 0);
    if (leader) {
      t2 = clock64();
      cyc[blockIdx.x * 2] = t1 - t0;
      cyc[blockIdx.x * 2 + 1] = t2 - t0;
    }
    __syncwarp();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem);
  }
}

template <int FORM>
void run(const char* name) {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 16);
  cudaFuncSetAttribute(k<FORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int rep = 64;
  for (int grid : {1, 148}) {
    k<FORM><<<grid, 128, 200 * 1024>>>(cyc, rep);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[296];
    cudaMemcpy(h, cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    double a = 0, b = 0;
    for (int i = 0; i < grid; ++i) { a += h[2 * i]; b += h[2 * i + 1]; }
    printf("%-40s grid %3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA  %s\n", name, grid, a / grid / (rep * 8),
           b / grid / (rep * 8), cudaGetErrorString(e));
  }
  cudaFree(cyc);
}

int main() {
  run<0>("SS 128x128x16 K-major/K-major (QK^T)");
  run<1>("TS 128x128x16 A=TMEM, B MN-major (PV)");
  run<3>("SS 128x128x16 B MN-major");
  run<2>("SS 128x256x16 (GEMM)");
  run<4>("SS 128x64x16");
  return 0;
}
