// NVSwitch multicast (NVLS) broadcast kernel: the lead GPU reads a tensor out of its own HBM ONCE and stores it with
// ``multimem.st`` through a multicast mapping - the switch replicates every 16-byte store into the memory that each
// GPU of the team bound to the multicast object, so rank 0's NVLink egress is paid once instead of once per peer
// (replaces the reference's ``source.cpu()`` + per-key H2D clone loop,
// /root/reference/any_device_parallel.py:600-663).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../common/host.h"

namespace pa {

__global__ void __launch_bounds__(256) multimem_bcast_kernel(const uint4* __restrict__ src, uint4* mc_dst, long long n16) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  // 4 independent 16-byte loads in flight per thread before the (posted) multicast stores
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const uint4 a = __ldg(src + i), b = __ldg(src + i + stride), c = __ldg(src + i + 2 * stride),
                d = __ldg(src + i + 3 * stride);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i),
                 "f"(__uint_as_float(a.x)), "f"(__uint_as_float(a.y)), "f"(__uint_as_float(a.z)), "f"(__uint_as_float(a.w)) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i + stride),
                 "f"(__uint_as_float(b.x)), "f"(__uint_as_float(b.y)), "f"(__uint_as_float(b.z)), "f"(__uint_as_float(b.w)) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i + 2 * stride),
                 "f"(__uint_as_float(c.x)), "f"(__uint_as_float(c.y)), "f"(__uint_as_float(c.z)), "f"(__uint_as_float(c.w)) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i + 3 * stride),
                 "f"(__uint_as_float(d.x)), "f"(__uint_as_float(d.y)), "f"(__uint_as_float(d.z)), "f"(__uint_as_float(d.w)) : "memory");
  }
  for (; i < n16; i += stride) {
    const uint4 a = __ldg(src + i);
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_dst + i),
                 "f"(__uint_as_float(a.x)), "f"(__uint_as_float(a.y)), "f"(__uint_as_float(a.z)), "f"(__uint_as_float(a.w)) : "memory");
  }
}

// src: local device memory (16-byte aligned), mc_dst: multicast virtual address, bytes % 16 == 0.
int multimem_bcast(const void* src, void* mc_dst, long long bytes, cudaStream_t st) {
  if (bytes <= 0) return 0;
  if ((bytes & 15) || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(mc_dst) & 15)) return -1;
  const long long n16 = bytes >> 4;
  long long blocks = (n16 + 256 * 4 - 1) / (256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  multimem_bcast_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(static_cast<const uint4*>(src),
                                                                   static_cast<uint4*>(mc_dst), n16);
  return (int)cudaGetLastError();
}

}  // namespace pa
