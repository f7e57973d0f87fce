// GroupNorm (+ optional fused SiLU) on channels-last bf16 activations [B, HW, C] — the layout the conv
// implicit-GEMM consumes.  Two kernels: (1) per-(batch, group) sum / sum-of-squares with a split
// reduction over HW (fp32 atomics into a zeroed workspace), (2) normalise + affine + SiLU, 16-byte I/O.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../common/host.h"

namespace pa {

static inline long long pa_min_ll(long long a, long long b) { return a < b ? a : b; }

namespace {
__device__ __forceinline__ float2 bf2f(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ uint32_t f2bf(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// grid: (splits, B).  Each block walks rows [hw0, hw1) of one sample; thread t owns channel pair(s).
__global__ void gn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ ws, int HW, int C, int G,
                                int rows_per_split) {
  extern __shared__ float sacc[];                 // [G][2]
  const int b = blockIdx.y;
  const int hw0 = blockIdx.x * rows_per_split;
  const int hw1 = min(HW, hw0 + rows_per_split);
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cg = C / G;
  const int pairs = C / 2;
  const uint32_t* xb = reinterpret_cast<const uint32_t*>(x + static_cast<long long>(b) * HW * C);
  for (int pr = threadIdx.x; pr < pairs; pr += blockDim.x) {
    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
    for (int r = hw0; r < hw1; ++r) {
      const float2 v = bf2f(__ldg(xb + static_cast<long long>(r) * pairs + pr));
      s0 += v.x; q0 += v.x * v.x;
      s1 += v.y; q1 += v.y * v.y;
    }
    const int g0 = (2 * pr) / cg, g1 = (2 * pr + 1) / cg;
    atomicAdd(&sacc[2 * g0], s0);
    atomicAdd(&sacc[2 * g0 + 1], q0);
    atomicAdd(&sacc[2 * g1], s1);
    atomicAdd(&sacc[2 * g1 + 1], q1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) atomicAdd(&ws[static_cast<long long>(b) * 2 * G + i], sacc[i]);
}

__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                const float* __restrict__ ws, const __nv_bfloat16* __restrict__ gamma,
                                const __nv_bfloat16* __restrict__ beta, int B, int HW, int C, int G, float eps,
                                int apply_silu) {
  const int cg = C / G;
  const float inv_n = 1.0f / (static_cast<float>(cg) * HW);
  const long long pairs_total = static_cast<long long>(B) * HW * C / 2;
  const uint32_t* xi = reinterpret_cast<const uint32_t*>(x);
  uint32_t* oo = reinterpret_cast<uint32_t*>(out);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < pairs_total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>((i * 2) % C);
    const int b = static_cast<int>((i * 2) / (static_cast<long long>(HW) * C));
    float2 v = bf2f(xi[i]);
    float r[2] = {v.x, v.y};
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ch = c + e;
      const int g = ch / cg;
      const float s = ws[(static_cast<long long>(b) * G + g) * 2], q = ws[(static_cast<long long>(b) * G + g) * 2 + 1];
      const float mean = s * inv_n;
      const float var = fmaxf(q * inv_n - mean * mean, 0.f);
      float y = (r[e] - mean) * rsqrtf(var + eps);
      y = y * __bfloat162float(gamma[ch]) + __bfloat162float(beta[ch]);
      if (apply_silu) y = y / (1.0f + __expf(-y));
      r[e] = y;
    }
    oo[i] = f2bf(r[0], r[1]);
  }
}
}  // namespace

// `workspace`: zeroed float[B * groups * 2]
int groupnorm_silu_nhwc_ws(const void* x, void* out, const void* gamma, const void* beta, float* workspace, int B,
                           int HW, int C, int groups, float eps, int apply_silu, cudaStream_t st) {
  if (C % groups || C % 2) return -1;
  const int rows_per_split = max(1, (HW + 63) / 64);
  const int splits = (HW + rows_per_split - 1) / rows_per_split;
  gn_stats_kernel<<<dim3(splits, B), 256, 2 * groups * sizeof(float), st>>>(static_cast<const __nv_bfloat16*>(x),
                                                                           workspace, HW, C, groups, rows_per_split);
  const long long pairs = static_cast<long long>(B) * HW * C / 2;
  const int blocks = static_cast<int>(pa_min_ll((pairs + 255) / 256, 148 * 16));
  gn_apply_kernel<<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out),
                                          workspace, static_cast<const __nv_bfloat16*>(gamma),
                                          static_cast<const __nv_bfloat16*>(beta), B, HW, C, groups, eps, apply_silu);
  return (int)cudaGetLastError();
}

}  // namespace pa
