// GroupNorm (+ optional fused SiLU) on channels-last bf16 activations [B, HW, C] — the layout the conv
// implicit-GEMM consumes.
//
// Default: ONE kernel, one thread-block CLUSTER of 8 CTAs per sample.  Every CTA owns a slab of HW/8 rows: pass 1
// accumulates per-channel sum / sum-of-squares over its slab with 16-byte loads (8 channels per thread, fixed per
// thread), folds them to per-group partials in shared memory and the 8 CTAs exchange the partials through
// DISTRIBUTED SHARED MEMORY (cluster.map_shared_rank) - no global atomics, no zeroed workspace, no second launch;
// pass 2 re-reads the slab (it was just streamed through L2), applies y = x * a[c] + b[c] (+ SiLU) and stores 16 bytes
// per thread.  Profile that motivated it (profiles/r2/profile_sdxl_b16_run2.txt): the two-kernel version below ran at
// ~1 TB/s (4-byte loads, an integer division and a rsqrt per element) and was 11 % of the SDXL step.
// Fallback (channel counts that are not a multiple of 8, or more than 512 16-byte vectors per row): the old pair of
// kernels - (1) per-(batch, group) sums with fp32 atomics into a zeroed workspace, (2) normalise + affine + SiLU.
#include <cooperative_groups.h>
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

// ------------------------------------------------------------------ cluster kernel
namespace {
constexpr int GN_CS = 8;          // CTAs per cluster (= per sample); portable cluster size

__global__ void __cluster_dims__(GN_CS, 1, 1) __launch_bounds__(512)
gn_cluster_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                  const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta, int HW, int C, int G,
                  int slabs, float eps, int apply_silu) {
  // One cluster = one (sample, channel slab): a slab is a run of whole groups whose width is a multiple of 8 channels,
  // so small batches still fill the machine (B * slabs clusters) while every thread keeps 16-byte accesses.
  namespace cgx = cooperative_groups;
  cgx::cluster_group cluster = cgx::this_cluster();
  extern __shared__ float sm[];
  const int Cs = C / slabs, Gs = G / slabs;                 // channels / groups of this slab
  float* chs = sm;                 // [Cs] per-channel sum over this CTA's rows
  float* chq = sm + Cs;            // [Cs] per-channel sum of squares
  float* gpart = sm + 2 * Cs;      // [2Gs] this CTA's per-group partials (read by the other CTAs through DSMEM)
  float* gtot = gpart + 2 * Gs;    // [2Gs] cluster totals
  float* ca = gtot + 2 * Gs;       // [Cs]  y = x * ca + cb
  float* cb = ca + Cs;
  const int cl = blockIdx.x / GN_CS;
  const int b = cl / slabs, slab = cl - b * slabs;
  const int c0 = slab * Cs;
  const int rank = static_cast<int>(cluster.block_rank());
  const int rows_per = (HW + GN_CS - 1) / GN_CS;
  const int r0 = rank * rows_per, r1 = min(HW, r0 + rows_per);
  const int vpr = Cs >> 3;                                  // 16-byte vectors of this slab per row
  const int row_vecs = C >> 3;                              // row stride in vectors
  const int lane = threadIdx.x % vpr, rsub = threadIdx.x / vpr, rp = blockDim.x / vpr;
  const int cg = C / G;
  const uint4* xb = reinterpret_cast<const uint4*>(x + static_cast<long long>(b) * HW * C + c0);
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  int r = r0 + rsub;
  for (; r + 3 * rp < r1; r += 4 * rp) {                    // four independent 16-byte loads in flight
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(xb + static_cast<long long>(r + u * rp) * row_vecs + lane);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t w0[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 a = bf2f(w0[e]);
        s[2 * e] += a.x;
        q[2 * e] = fmaf(a.x, a.x, q[2 * e]);
        s[2 * e + 1] += a.y;
        q[2 * e + 1] = fmaf(a.y, a.y, q[2 * e + 1]);
      }
    }
  }
  for (; r < r1; r += rp) {
    const uint4 v0 = __ldg(xb + static_cast<long long>(r) * row_vecs + lane);
    const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 a = bf2f(w0[e]);
      s[2 * e] += a.x;
      q[2 * e] = fmaf(a.x, a.x, q[2 * e]);
      s[2 * e + 1] += a.y;
      q[2 * e + 1] = fmaf(a.y, a.y, q[2 * e + 1]);
    }
  }
  // deterministic fold over the rp row-parallel threads of every channel: partials to shared memory, fixed-order sum
  // (no atomics anywhere in this kernel -> the same input always gives bit-identical output, unlike the fallback)
  float* part = cb + Cs;           // [2][rp][Cs]
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    part[rsub * Cs + lane * 8 + e] = s[e];
    part[(rp + rsub) * Cs + lane * 8 + e] = q[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
    float ts = 0.f, tq = 0.f;
    for (int k = 0; k < rp; ++k) {
      ts += part[k * Cs + c];
      tq += part[(rp + k) * Cs + c];
    }
    chs[c] = ts;
    chq[c] = tq;
  }
  __syncthreads();
  for (int g = threadIdx.x; g < Gs; g += blockDim.x) {
    float gs = 0.f, gq = 0.f;
    for (int c = g * cg; c < (g + 1) * cg; ++c) {
      gs += chs[c];
      gq += chq[c];
    }
    gpart[2 * g] = gs;
    gpart[2 * g + 1] = gq;
  }
  cluster.sync();                                           // every CTA's partials are in its shared memory
  for (int i = threadIdx.x; i < 2 * Gs; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < GN_CS; ++k) t += *cluster.map_shared_rank(gpart + i, k);      // DSMEM read of CTA k's partial
    gtot[i] = t;
  }
  cluster.sync();                                           // remote reads done before any CTA may exit; gtot visible
  const float inv_n = 1.0f / (static_cast<float>(cg) * HW);
  for (int c = threadIdx.x; c < Cs; c += blockDim.x) {
    const int g = c / cg;
    const float mean = gtot[2 * g] * inv_n;
    const float var = fmaxf(gtot[2 * g + 1] * inv_n - mean * mean, 0.f);
    const float a = rsqrtf(var + eps) * __bfloat162float(gamma[c0 + c]);
    ca[c] = a;
    cb[c] = __bfloat162float(beta[c0 + c]) - mean * a;
  }
  __syncthreads();
  float a8[8], b8[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    a8[e] = ca[lane * 8 + e];
    b8[e] = cb[lane * 8 + e];
  }
  uint4* ob = reinterpret_cast<uint4*>(out + static_cast<long long>(b) * HW * C + c0);
  for (r = r0 + rsub; r < r1; r += rp) {
    const uint4 v0 = xb[static_cast<long long>(r) * row_vecs + lane];     // L2: this slab was streamed in pass 1
    const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 t = bf2f(w0[e]);
      float y0 = fmaf(t.x, a8[2 * e], b8[2 * e]), y1 = fmaf(t.y, a8[2 * e + 1], b8[2 * e + 1]);
      if (apply_silu) {
        y0 = y0 / (1.0f + __expf(-y0));
        y1 = y1 / (1.0f + __expf(-y1));
      }
      o[e] = f2bf(y0, y1);
    }
    ob[static_cast<long long>(r) * row_vecs + lane] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}
}  // namespace

// One launch, no workspace.  Returns -100 when the shape needs the two-kernel fallback.
int groupnorm_silu_nhwc_cluster(const void* x, void* out, const void* gamma, const void* beta, int B, int HW, int C,
                                int groups, float eps, int apply_silu, cudaStream_t st) {
  if (C % groups) return -1;
  if (C % 8) return -100;
  const int cg = C / groups;
  // channel slabs: whole groups, a multiple of 8 channels wide; as many as needed to put >= ~1 cluster CTA on every SM
  int unit = cg;                                            // smallest legal slab width = lcm(cg, 8)
  while (unit % 8) unit += cg;
  if (C % unit) return -100;
  int slabs = 1;
  const int max_slabs = C / unit;
  while (B * slabs * GN_CS < 148 && slabs * 2 <= max_slabs && (C / (slabs * 2)) % unit == 0) slabs *= 2;
  if (groups % slabs) return -100;
  const int vpr = (C / slabs) / 8;
  if (vpr > 512 || vpr < 1) return -100;
  const int threads = vpr * (512 / vpr);
  const int rp = threads / vpr;
  const size_t smem = (4 * static_cast<size_t>(C / slabs) + 4 * static_cast<size_t>(groups / slabs) +
                       2 * static_cast<size_t>(rp) * (C / slabs)) * sizeof(float);
  if (smem > 48 * 1024) return -100;
  gn_cluster_kernel<<<B * slabs * GN_CS, threads, smem, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                               static_cast<__nv_bfloat16*>(out),
                                                               static_cast<const __nv_bfloat16*>(gamma),
                                                               static_cast<const __nv_bfloat16*>(beta), HW, C, groups,
                                                               slabs, eps, apply_silu);
  return (int)cudaGetLastError();
}

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
