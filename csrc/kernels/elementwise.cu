// Memory-bound glue kernels (bf16 I/O, fp32 math, 16-byte vector accesses):
// LayerNorm + AdaLN modulate, sinusoidal timestep embedding, patchify (NCHW latent -> tokens, source may be a
// peer GPU), SiLU/add, GroupNorm(+SiLU) NHWC, CFG + Euler update with (peer) store, cross-GPU flag words.
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include "../common/host.h"
#include "../common/ptx.cuh"

namespace pa {

static inline long long pa_min_ll(long long a, long long b) { return a < b ? a : b; }

__device__ __forceinline__ float2 bf2f(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ uint32_t f2bf(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&o)[8]) {
  float2 a = bf2f(u.x), b = bf2f(u.y), c = bf2f(u.z), d = bf2f(u.w);
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y; o[4] = c.x; o[5] = c.y; o[6] = d.x; o[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  uint4 u;
  u.x = f2bf(v[0], v[1]); u.y = f2bf(v[2], v[3]); u.z = f2bf(v[4], v[5]); u.w = f2bf(v[6], v[7]);
  return u;
}
// MUFU.TANH (abs error ~2^-11): plenty for a bf16 gate vector
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------ LayerNorm + modulate
// One warp per row; the row (D <= 8192) is held in registers between the statistics and the write pass.
template <int MAX_VEC>   // MAX_VEC uint4 per lane: D <= 32*8*MAX_VEC
__global__ void __launch_bounds__(256) ln_mod_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                     long long x_bs, __nv_bfloat16* __restrict__ out, long long ldo,
                                                     long long o_bs, const __nv_bfloat16* __restrict__ scale,
                                                     const __nv_bfloat16* __restrict__ shift, long long mod_bs,
                                                     const __nv_bfloat16* __restrict__ gamma,
                                                     const __nv_bfloat16* __restrict__ beta, int batch, int rows, int D,
                                                     float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch * rows) return;
  const int b = warp / rows, r = warp - b * rows;
  const uint4* xr = reinterpret_cast<const uint4*>(x + b * x_bs + static_cast<long long>(r) * ldx);
  const int nvec = D >> 3;
  uint4 buf[MAX_VEC];
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      buf[i] = xr[idx];
      float v[8];
      unpack8(buf[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) { s += v[e]; ss += v[e] * v[e]; }
    }
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const float mean = s / D;
  const float var = fmaxf(ss / D - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  uint4* orow = reinterpret_cast<uint4*>(out + b * o_bs + static_cast<long long>(r) * ldo);
  const uint4* sc = scale ? reinterpret_cast<const uint4*>(scale + b * mod_bs) : nullptr;
  const uint4* sh = shift ? reinterpret_cast<const uint4*>(shift + b * mod_bs) : nullptr;
  const uint4* ga = gamma ? reinterpret_cast<const uint4*>(gamma) : nullptr;
  const uint4* be = beta ? reinterpret_cast<const uint4*>(beta) : nullptr;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      float v[8];
      unpack8(buf[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (v[e] - mean) * rstd;
      if (ga) {
        float g[8];
        unpack8(__ldg(ga + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= g[e];
      }
      if (be) {
        float g[8];
        unpack8(__ldg(be + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += g[e];
      }
      if (sc) {
        float g[8];
        unpack8(__ldg(sc + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= (1.0f + g[e]);
      }
      if (sh) {
        float g[8];
        unpack8(__ldg(sh + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += g[e];
      }
      orow[idx] = pack8(v);
    }
  }
}

// Specialised variants: which of gamma / beta / scale / shift exist is a compile-time flag set (bit 0 gamma, 1 beta,
// 2 scale, 3 shift), so no predicated-off instructions are issued, and the affine part is folded into ONE fma per
// element:  y = x * m + c  with  m = rstd * gamma * (1 + scale),  c = (-mean * rstd * gamma + beta) * (1 + scale) + shift.
// The generic kernel above issued ~25 instructions per element (ncu: 43 % issue-slot utilisation, 2.2 TB/s); this is ~10.
// F8: the output is the next GEMM's MXFP8 A operand (e4m3 [batch, rows, D] contiguous + UE8M0 scale chunks in the
// gemm_mxfp8.cu layout) instead of bf16: four neighbouring lanes hold one 32-element MX block (8 elements each), its
// amax is two shuffles away; `out` then points at the e4m3 bytes and `sf8` at the scale chunks.
template <int MAX_VEC, int FLAGS, bool F8 = false>
__global__ void __launch_bounds__(256) ln_mod_fast_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                          long long x_bs, __nv_bfloat16* __restrict__ out, long long ldo,
                                                          long long o_bs, const __nv_bfloat16* __restrict__ scale,
                                                          const __nv_bfloat16* __restrict__ shift, long long mod_bs,
                                                          const __nv_bfloat16* __restrict__ gamma,
                                                          const __nv_bfloat16* __restrict__ beta, int batch, int rows,
                                                          int D, float eps, uint8_t* __restrict__ sf8 = nullptr) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch * rows) return;
  const int b = warp / rows, r = warp - b * rows;
  const uint4* xr = reinterpret_cast<const uint4*>(x + b * x_bs + static_cast<long long>(r) * ldx);
  const int nvec = D >> 3;
  uint4 buf[MAX_VEC];
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) buf[i] = xr[idx];
  }
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    if (lane + i * 32 < nvec) {
      float v[8];
      unpack8(buf[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s += v[e];
        ss = fmaf(v[e], v[e], ss);
      }
    }
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const float mean = s / D;
  const float rstd = rsqrtf(fmaxf(ss / D - mean * mean, 0.f) + eps);
  const float nmr = -mean * rstd;
  uint4* orow = reinterpret_cast<uint4*>(out + b * o_bs + static_cast<long long>(r) * ldo);
  const uint4* sc = reinterpret_cast<const uint4*>(scale + ((FLAGS & 4) ? b * mod_bs : 0));
  const uint4* sh = reinterpret_cast<const uint4*>(shift + ((FLAGS & 8) ? b * mod_bs : 0));
  const uint4* ga = reinterpret_cast<const uint4*>(gamma);
  const uint4* be = reinterpret_cast<const uint4*>(beta);
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      float v[8], m[8], c[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { m[e] = rstd; c[e] = nmr; }
      if (FLAGS & 1) {
        float g[8];
        unpack8(__ldg(ga + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) { m[e] *= g[e]; c[e] *= g[e]; }
      }
      if (FLAGS & 2) {
        float g[8];
        unpack8(__ldg(be + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) c[e] += g[e];
      }
      if (FLAGS & 4) {
        float g[8];
        unpack8(__ldg(sc + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float t = 1.0f + g[e];
          m[e] *= t;
          c[e] *= t;
        }
      }
      if (FLAGS & 8) {
        float g[8];
        unpack8(__ldg(sh + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) c[e] += g[e];
      }
      unpack8(buf[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], m[e], c[e]);
      if (F8) {
        // D % 32 == 0 (checked on the host), so the four lanes of an MX block are all inside `idx < nvec` together
        float amax = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) amax = fmaxf(amax, fabsf(v[e]));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        int ex = -127;
        if (amax > 0.f) {
          ex = static_cast<int>(ceilf(log2f(amax * (1.0f / 448.0f))));
          ex = max(-127, min(127, ex));
        }
        const float inv = exp2f(static_cast<float>(-ex));
        uint32_t pk[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const __nv_fp8x2_storage_t lo =
              __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j] * inv, v[4 * j + 1] * inv), __NV_SATFINITE, __NV_E4M3);
          const __nv_fp8x2_storage_t hi =
              __nv_cvt_float2_to_fp8x2(make_float2(v[4 * j + 2] * inv, v[4 * j + 3] * inv), __NV_SATFINITE, __NV_E4M3);
          pk[j] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
        }
        uint8_t* q8 = reinterpret_cast<uint8_t*>(out) + (static_cast<long long>(b) * rows + r) * D;
        *reinterpret_cast<uint2*>(q8 + idx * 8) = make_uint2(pk[0], pk[1]);
        if ((lane & 3) == 0) {
          const int kb = idx >> 2;                                          // 32-element block index inside the row
          const long long chunk = (static_cast<long long>(b) * ((rows + 127) >> 7) + (r >> 7)) * (D >> 7) + (kb >> 2);
          sf8[chunk * 512 + (r & 31) * 16 + ((r >> 5) & 3) * 4 + (kb & 3)] = static_cast<uint8_t>(ex + 127);
        }
      } else {
        orow[idx] = pack8(v);
      }
    }
  }
}

// LayerNorm + modulate with MXFP8 output (scale + shift present, the FLUX case): q8 [batch, rows, D] e4m3 contiguous,
// sf8 scale chunks (zero-initialised by the caller).
int layernorm_modulate_fp8(const void* x, long long ldx, long long x_bs, void* q8, void* sf8, const void* scale,
                           const void* shift, long long mod_bs, int batch, int rows, int D, float eps, cudaStream_t st) {
  if (D % 256 || ldx % 8 || x_bs % 8 || mod_bs % 8 || !scale || !shift) return -1;   // whole warps per vector step
  const long long warps = static_cast<long long>(batch) * rows;
  const int threads = 256;
  const int blocks = static_cast<int>((warps * 32 + threads - 1) / threads);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto SC = static_cast<const __nv_bfloat16*>(scale);
  auto SH = static_cast<const __nv_bfloat16*>(shift);
  auto O = static_cast<__nv_bfloat16*>(q8);
  auto S8 = static_cast<uint8_t*>(sf8);
  if (D <= 32 * 8 * 4)
    ln_mod_fast_kernel<4, 12, true><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, 0, 0, SC, SH, mod_bs, nullptr, nullptr, batch, rows, D, eps, S8);
  else if (D <= 32 * 8 * 12)
    ln_mod_fast_kernel<12, 12, true><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, 0, 0, SC, SH, mod_bs, nullptr, nullptr, batch, rows, D, eps, S8);
  else if (D <= 32 * 8 * 20)
    ln_mod_fast_kernel<20, 12, true><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, 0, 0, SC, SH, mod_bs, nullptr, nullptr, batch, rows, D, eps, S8);
  else
    return -2;
  return (int)cudaGetLastError();
}

template <int FLAGS>
static bool launch_ln_fast(int blocks, int threads, cudaStream_t st, const __nv_bfloat16* X, long long ldx, long long x_bs,
                           __nv_bfloat16* O, long long ldo, long long o_bs, const __nv_bfloat16* SC,
                           const __nv_bfloat16* SH, long long mod_bs, const __nv_bfloat16* G, const __nv_bfloat16* Bt,
                           int batch, int rows, int D, float eps) {
  if (D <= 32 * 8 * 4)
    ln_mod_fast_kernel<4, FLAGS><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, ldo, o_bs, SC, SH, mod_bs, G, Bt, batch, rows, D, eps);
  else if (D <= 32 * 8 * 12)
    ln_mod_fast_kernel<12, FLAGS><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, ldo, o_bs, SC, SH, mod_bs, G, Bt, batch, rows, D, eps);
  else if (D <= 32 * 8 * 20)
    ln_mod_fast_kernel<20, FLAGS><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, ldo, o_bs, SC, SH, mod_bs, G, Bt, batch, rows, D, eps);
  else
    return false;
  return true;
}

int layernorm_modulate(const void* x, long long ldx, long long x_bs, void* out, long long ldo, long long o_bs,
                       const void* scale, const void* shift, long long mod_bs, const void* gamma, const void* beta,
                       int batch, int rows, int D, float eps, cudaStream_t st) {
  if (D % 8 || ldx % 8 || ldo % 8 || x_bs % 8 || o_bs % 8 || mod_bs % 8) return -1;
  const long long warps = static_cast<long long>(batch) * rows;
  const int threads = 256;
  const int blocks = static_cast<int>((warps * 32 + threads - 1) / threads);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto O = static_cast<__nv_bfloat16*>(out);
  auto SC = static_cast<const __nv_bfloat16*>(scale);
  auto SH = static_cast<const __nv_bfloat16*>(shift);
  auto G = static_cast<const __nv_bfloat16*>(gamma);
  auto Bt = static_cast<const __nv_bfloat16*>(beta);
  const int flags = (G ? 1 : 0) | (Bt ? 2 : 0) | (SC ? 4 : 0) | (SH ? 8 : 0);
#define PA_LN_FAST(F)                                                                                              \
  if (flags == F && launch_ln_fast<F>(blocks, threads, st, X, ldx, x_bs, O, ldo, o_bs, SC, SH, mod_bs, G, Bt, batch, rows, \
                                      D, eps))                                                                    \
    return (int)cudaGetLastError();
  PA_LN_FAST(12) PA_LN_FAST(4) PA_LN_FAST(3) PA_LN_FAST(0)
#undef PA_LN_FAST
  if (D <= 32 * 8 * 4)
    ln_mod_kernel<4><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, ldo, o_bs, SC, SH, mod_bs, G, Bt, batch, rows, D, eps);
  else if (D <= 32 * 8 * 12)
    ln_mod_kernel<12><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, ldo, o_bs, SC, SH, mod_bs, G, Bt, batch, rows, D, eps);
  else if (D <= 32 * 8 * 20)
    ln_mod_kernel<20><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, ldo, o_bs, SC, SH, mod_bs, G, Bt, batch, rows, D, eps);
  else
    return -2;
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ RMSNorm + modulate / gated residual
// NextDiT-style (Z-Image, Lumina) blocks:   out = rms(x) * w * (1 + scale)            (pre-norm + AdaLN scale)
//                                           out = residual + tanh(gate) * rms(x) * w  (post-norm "sandwich" + gate)
// One warp per row, row in registers; scale / gate are per-sample vectors (stride mod_bs), either may be null.
template <int MAX_VEC, int FLAGS>     // FLAGS: bit 0 weight, 1 scale, 2 gate, 3 residual present; -1 = decide at run time
__global__ void __launch_bounds__(256) rms_mod_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, long long x_bs,
                                                      __nv_bfloat16* __restrict__ out, long long ldo, long long o_bs,
                                                      const __nv_bfloat16* __restrict__ weight,
                                                      const __nv_bfloat16* __restrict__ scale,
                                                      const __nv_bfloat16* __restrict__ gate, long long mod_bs,
                                                      const __nv_bfloat16* residual, long long ldr, long long r_bs,
                                                      int batch, int rows, int D, float eps, int tanh_gate) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch * rows) return;
  const int b = warp / rows, r = warp - b * rows;
  const uint4* xr = reinterpret_cast<const uint4*>(x + b * x_bs + static_cast<long long>(r) * ldx);
  const int nvec = D >> 3;
  uint4 buf[MAX_VEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      buf[i] = xr[idx];
      float v[8];
      unpack8(buf[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / D + eps);
  uint4* orow = reinterpret_cast<uint4*>(out + b * o_bs + static_cast<long long>(r) * ldo);
  const uint4* rrow =
      residual ? reinterpret_cast<const uint4*>(residual + b * r_bs + static_cast<long long>(r) * ldr) : nullptr;
  const bool has_w = FLAGS < 0 ? weight != nullptr : (FLAGS & 1) != 0;
  const bool has_sc = FLAGS < 0 ? scale != nullptr : (FLAGS & 2) != 0;
  const bool has_gt = FLAGS < 0 ? gate != nullptr : (FLAGS & 4) != 0;
  const bool has_res = FLAGS < 0 ? residual != nullptr : (FLAGS & 8) != 0;
  const uint4* wv = reinterpret_cast<const uint4*>(weight);
  const uint4* sc = reinterpret_cast<const uint4*>(scale + (has_sc ? b * mod_bs : 0));
  const uint4* gt = reinterpret_cast<const uint4*>(gate + (has_gt ? b * mod_bs : 0));
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      float v[8];
      unpack8(buf[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] *= rstd;
      if (has_w) {
        float g[8];
        unpack8(__ldg(wv + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= g[e];
      }
      if (has_sc) {
        float g[8];
        unpack8(__ldg(sc + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= (1.0f + g[e]);
      }
      if (has_gt) {
        float g[8];
        unpack8(__ldg(gt + idx), g);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] *= tanh_gate ? tanh_fast(g[e]) : g[e];
      }
      if (has_res) {
        float g[8];
        unpack8(rrow[idx], g);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] += g[e];
      }
      orow[idx] = pack8(v);
    }
  }
}

int rmsnorm_mod(const void* x, long long ldx, long long x_bs, void* out, long long ldo, long long o_bs,
                const void* weight, const void* scale, const void* gate, long long mod_bs, const void* residual,
                long long ldr, long long r_bs, int batch, int rows, int D, float eps, int tanh_gate, cudaStream_t st) {
  if (D % 8 || ldx % 8 || ldo % 8 || x_bs % 8 || o_bs % 8 || mod_bs % 8 || ldr % 8 || r_bs % 8) return -1;
  const long long warps = static_cast<long long>(batch) * rows;
  const int threads = 256;
  const int blocks = static_cast<int>((warps * 32 + threads - 1) / threads);
  auto X = static_cast<const __nv_bfloat16*>(x);
  auto O = static_cast<__nv_bfloat16*>(out);
  auto Wt = static_cast<const __nv_bfloat16*>(weight);
  auto SC = static_cast<const __nv_bfloat16*>(scale);
  auto G = static_cast<const __nv_bfloat16*>(gate);
  auto R = static_cast<const __nv_bfloat16*>(residual);
  // the combinations the NextDiT blocks use are compiled without run-time null checks (no predicated-off instructions)
  const int flags = (Wt ? 1 : 0) | (SC ? 2 : 0) | (G ? 4 : 0) | (R ? 8 : 0);
#define PA_RMS_MOD(V, F)                                                                                                  \
  rms_mod_kernel<V, F><<<blocks, threads, 0, st>>>(X, ldx, x_bs, O, ldo, o_bs, Wt, SC, G, mod_bs, R, ldr, r_bs, batch, rows, \
                                                  D, eps, tanh_gate)
#define PA_RMS_MOD_D(F)                     \
  {                                         \
    if (D <= 32 * 8 * 4) PA_RMS_MOD(4, F);  \
    else if (D <= 32 * 8 * 12) PA_RMS_MOD(12, F); \
    else if (D <= 32 * 8 * 20) PA_RMS_MOD(20, F); \
    else return -2;                         \
  }
  if (flags == 3) PA_RMS_MOD_D(3)
  else if (flags == 13) PA_RMS_MOD_D(13)
  else if (flags == 9) PA_RMS_MOD_D(9)
  else if (flags == 1) PA_RMS_MOD_D(1)
  else PA_RMS_MOD_D(-1)
#undef PA_RMS_MOD_D
#undef PA_RMS_MOD
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ full-width RMSNorm (+ RoPE) in place
// WAN-style q/k norm: RMS over the whole hidden dim (all heads), learned weight, then per-head RoPE on
// adjacent pairs with a [L, 64] (cos, sin) table.  One warp per row, row held in registers.
template <int MAX_VEC>
__global__ void __launch_bounds__(256) rms_rope_kernel(__nv_bfloat16* __restrict__ x, long long ldx, long long x_bs,
                                                       const __nv_bfloat16* __restrict__ w,
                                                       const float2* __restrict__ rope, int batch, int rows, int D,
                                                       float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch * rows) return;
  const int b = warp / rows, r = warp - b * rows;
  uint4* xr = reinterpret_cast<uint4*>(x + b * x_bs + static_cast<long long>(r) * ldx);
  const int nvec = D >> 3;
  uint4 buf[MAX_VEC];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      buf[i] = xr[idx];
      float v[8];
      unpack8(buf[i], v);
#pragma unroll
      for (int e = 0; e < 8; ++e) ss += v[e] * v[e];
    }
  }
  ss = warp_sum(ss);
  const float rr = rsqrtf(ss / D + eps);
  const uint4* wv = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int i = 0; i < MAX_VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx < nvec) {
      float v[8], g[8];
      unpack8(buf[i], v);
      unpack8(__ldg(wv + idx), g);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = v[e] * rr * g[e];
      if (rope != nullptr) {
        const int d0 = (idx * 8) & 127;                      // offset inside the 128-wide head
        const float4* rp = reinterpret_cast<const float4*>(rope + static_cast<long long>(r) * 64 + (d0 >> 1));
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const float4 cs = __ldg(rp + j);
          const float x0 = v[4 * j], x1 = v[4 * j + 1], x2 = v[4 * j + 2], x3 = v[4 * j + 3];
          v[4 * j] = cs.x * x0 - cs.y * x1;
          v[4 * j + 1] = cs.y * x0 + cs.x * x1;
          v[4 * j + 2] = cs.z * x2 - cs.w * x3;
          v[4 * j + 3] = cs.w * x2 + cs.z * x3;
        }
      }
      xr[idx] = pack8(v);
    }
  }
}

int rms_rope_inplace(void* x, long long ldx, long long x_bs, const void* w, const void* rope, int batch, int rows,
                     int D, float eps, cudaStream_t st) {
  if (D % 128 || ldx % 8 || x_bs % 8) return -1;
  const long long warps = static_cast<long long>(batch) * rows;
  const int blocks = static_cast<int>((warps * 32 + 255) / 256);
  auto X = static_cast<__nv_bfloat16*>(x);
  auto Wt = static_cast<const __nv_bfloat16*>(w);
  auto R = static_cast<const float2*>(rope);
  if (D <= 32 * 8 * 4) rms_rope_kernel<4><<<blocks, 256, 0, st>>>(X, ldx, x_bs, Wt, R, batch, rows, D, eps);
  else if (D <= 32 * 8 * 12) rms_rope_kernel<12><<<blocks, 256, 0, st>>>(X, ldx, x_bs, Wt, R, batch, rows, D, eps);
  else if (D <= 32 * 8 * 20) rms_rope_kernel<20><<<blocks, 256, 0, st>>>(X, ldx, x_bs, Wt, R, batch, rows, D, eps);
  else return -2;
  return (int)cudaGetLastError();
}

// out[b, i, :] = a[b, :] + m[i, :]     (per-block modulation tables = learned offsets + time projection)
__global__ void bcast_add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ m, uint4* __restrict__ out,
                                 int B, int nblk, int nv) {
  const long long total = static_cast<long long>(B) * nblk * nv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % nv);
    const long long r = i / nv;
    const int blk = static_cast<int>(r % nblk);
    const int b = static_cast<int>(r / nblk);
    float x[8], y[8];
    unpack8(__ldg(a + static_cast<long long>(b) * nv + c), x);
    unpack8(__ldg(m + static_cast<long long>(blk) * nv + c), y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
    out[i] = pack8(x);
  }
}

int bcast_add(const void* a, const void* m, void* out, int B, int nblk, int n, cudaStream_t st) {
  if (n % 8) return -1;
  const long long total = static_cast<long long>(B) * nblk * (n / 8);
  bcast_add_kernel<<<static_cast<int>(pa_min_ll((total + 255) / 256, 148 * 8)), 256, 0, st>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(m), static_cast<uint4*>(out), B, nblk, n / 8);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ row softmax (in place, bf16)
// One CTA per row; used by the VAE's single-head 512-wide attention, which is evaluated as
// GEMM -> softmax -> GEMM on the tensor cores (it runs once per image, not per denoise step).
__global__ void __launch_bounds__(256) softmax_rows_kernel(__nv_bfloat16* __restrict__ x, long long ld, int n,
                                                           float scale_log2) {
  __shared__ float red[8];
  uint4* row = reinterpret_cast<uint4*>(x + static_cast<long long>(blockIdx.x) * ld);
  const int nv = n >> 3;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float v[8];
    unpack8(row[i], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) mx = fmaxf(mx, v[e]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float v[8];
    unpack8(row[i], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) sum += exp2f((v[e] - mx) * scale_log2);
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float v[8];
    unpack8(row[i], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = exp2f((v[e] - mx) * scale_log2) * inv;
    row[i] = pack8(v);
  }
}

int softmax_rows(void* x, long long ld, int rows, int n, float scale, cudaStream_t st) {
  if (n % 8 || ld % 8) return -1;
  softmax_rows_kernel<<<rows, 256, 0, st>>>(static_cast<__nv_bfloat16*>(x), ld, n, scale * 1.4426950408889634f);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ timestep embedding
__global__ void temb_kernel(const void* t, __nv_bfloat16* out, long long ldo, int B, int dim, float time_factor,
                            float max_period, int t_is_bf16) {
  const int b = blockIdx.x;
  const int half = dim / 2;
  float tv = t_is_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(t)[b]) : static_cast<const float*>(t)[b];
  tv *= time_factor;
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float f = expf(-logf(max_period) * static_cast<float>(i) / static_cast<float>(half));
    float sn, cs;
    sincosf(tv * f, &sn, &cs);
    out[b * ldo + i] = __float2bfloat16(cs);
    out[b * ldo + half + i] = __float2bfloat16(sn);
  }
}

int timestep_embedding(const void* t, void* out, long long ldo, int B, int dim, float time_factor, float max_period,
                       int t_is_bf16, cudaStream_t st) {
  temb_kernel<<<B, 128, 0, st>>>(t, static_cast<__nv_bfloat16*>(out), ldo, B, dim, time_factor, max_period, t_is_bf16);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ patchify (the unfused scatter path)
// out[b, hh*Wp + ww, c*ps*ps + ph*ps + pw] = x[b, c, hh*ps+ph, ww*ps+pw];  x may be a peer mapping.
__global__ void patchify_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, long long ldo,
                                long long o_bs, int B, int C, int H, int W, int ps) {
  const int Wp = W / ps, Hp = H / ps;
  const long long total = static_cast<long long>(B) * C * H * (W / 2);   // pairs along W
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int w2 = static_cast<int>(i % (W / 2));
    long long r = i / (W / 2);
    const int h = static_cast<int>(r % H);
    r /= H;
    const int c = static_cast<int>(r % C);
    const int b = static_cast<int>(r / C);
    const uint32_t v = *reinterpret_cast<const uint32_t*>(x + ((static_cast<long long>(b) * C + c) * H + h) * W + w2 * 2);
    const int hh = h / ps, ph = h % ps;
    if (ps == 2) {
      const int ww = w2;
      __nv_bfloat16* o = out + b * o_bs + static_cast<long long>(hh * Wp + ww) * ldo + c * 4 + ph * 2;
      *reinterpret_cast<uint32_t*>(o) = v;
    } else {
      const __nv_bfloat16* pv = reinterpret_cast<const __nv_bfloat16*>(&v);
      for (int e = 0; e < 2; ++e) {
        const int w = w2 * 2 + e, ww = w / ps, pw = w % ps;
        out[b * o_bs + static_cast<long long>(hh * Wp + ww) * ldo + (c * ps + ph) * ps + pw] = pv[e];
      }
    }
  }
  (void)Hp;
}

int patchify(const void* x, void* out, long long ldo, long long o_bs, int B, int C, int H, int W, int ps,
             cudaStream_t st) {
  if (W % 2) return -1;
  const long long total = static_cast<long long>(B) * C * H * (W / 2);
  const int blocks = static_cast<int>(pa_min_ll((total + 255) / 256, 148 * 8));
  patchify_kernel<<<blocks, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), ldo,
                                          o_bs, B, C, H, W, ps);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ small elementwise
__global__ void silu_kernel(const uint4* x, uint4* out, long long nvec) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float v[8];
    unpack8(x[i], v);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = v[e] / (1.0f + __expf(-v[e]));
    out[i] = pack8(v);
  }
}
int silu_bf16(const void* x, void* out, long long n, cudaStream_t st) {
  if (n % 8) return -1;
  const long long nv = n / 8;
  silu_kernel<<<static_cast<int>(pa_min_ll((nv + 255) / 256, 148 * 8)), 256, 0, st>>>(
      static_cast<const uint4*>(x), static_cast<uint4*>(out), nv);
  return (int)cudaGetLastError();
}

__global__ void add_kernel(const uint4* a, const uint4* b, uint4* out, long long nvec) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float x[8], y[8];
    unpack8(a[i], x);
    unpack8(b[i], y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
    out[i] = pack8(x);
  }
}
int add_bf16(const void* a, const void* b, void* out, long long n, cudaStream_t st) {
  if (n % 8) return -1;
  const long long nv = n / 8;
  add_kernel<<<static_cast<int>(pa_min_ll((nv + 255) / 256, 148 * 8)), 256, 0, st>>>(
      static_cast<const uint4*>(a), static_cast<const uint4*>(b), static_cast<uint4*>(out), nv);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ CFG + Euler + (peer) store
// mode 0: x_out = eps (passthrough)        mode 1: x_out = x + (s' - s) * d,  d = u + cfg*(c - u)
// (flow / v-prediction "CONST" parameterisation);  mode 2: eps-prediction Euler: d = eps,
// x_out = x + (s' - s) * d as well (k-diffusion's to_d(x, sigma, x - sigma*eps) == eps).
__global__ void cfg_euler_kernel(const uint4* __restrict__ x, const uint4* __restrict__ ec, const uint4* __restrict__ eu,
                                 uint4* __restrict__ xo, const float* __restrict__ sigmas, float cfg,
                                 long long nvec_per_sample, int batch, long long out_vec_off, int mode) {
  const long long total = nvec_per_sample * batch;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / nvec_per_sample);
    float d[8];
    unpack8(ec[i], d);
    if (eu != nullptr) {
      float u[8];
      unpack8(eu[i], u);
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] = u[e] + cfg * (d[e] - u[e]);
    }
    if (mode != 0) {
      const float dt = sigmas[2 * b + 1] - sigmas[2 * b];
      float xv[8];
      unpack8(x[i], xv);
#pragma unroll
      for (int e = 0; e < 8; ++e) d[e] = xv[e] + dt * d[e];
    }
    xo[out_vec_off + i] = pack8(d);
  }
}

int cfg_euler_store(const void* x, const void* ec, const void* eu, void* x_out, const void* sigmas, float cfg,
                    long long n_per_sample, int batch, long long out_sample_off, int mode, cudaStream_t st) {
  if (n_per_sample % 8) return -1;
  const long long nv = n_per_sample / 8;
  const long long total = nv * batch;
  cfg_euler_kernel<<<static_cast<int>(pa_min_ll((total + 255) / 256, 148 * 8)), 256, 0, st>>>(
      static_cast<const uint4*>(x), static_cast<const uint4*>(ec), static_cast<const uint4*>(eu),
      static_cast<uint4*>(x_out), static_cast<const float*>(sigmas), cfg, nv, batch, out_sample_off * nv, mode);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------ cross-GPU flags
// signal: write `value` into slot `slot` of every peer's flag array (release at system scope, after all
// prior writes of this stream are visible).  wait: spin (acquire) until flags[first..first+n) >= value,
// bounded by a cycle budget -> sets *error_word instead of hanging.
__global__ void signal_kernel(uint32_t* const* peers, int n_peers, int slot, uint32_t value) {
  __threadfence_system();
  const int i = threadIdx.x;
  if (i < n_peers) ptx::st_release_sys_u32(peers[i] + slot, value);
}

__global__ void wait_kernel(const uint32_t* flags, int first, int n, uint32_t value, long long timeout,
                            uint32_t* err) {
  const int i = threadIdx.x;
  if (i < n) {
    const long long t0 = clock64();
    while (static_cast<int32_t>(ptx::ld_acquire_sys_u32(flags + first + i) - value) < 0) {
      if (clock64() - t0 > timeout) {
        if (err) atomicExch(err, 0xDEAD0000u | static_cast<uint32_t>(first + i));
        break;
      }
      __nanosleep(64);
    }
  }
  __threadfence_system();
}

int signal_flags(uint32_t* const* peer_flag_ptrs, int n_peers, int slot, uint32_t value, cudaStream_t st) {
  signal_kernel<<<1, 32, 0, st>>>(peer_flag_ptrs, n_peers, slot, value);
  return (int)cudaGetLastError();
}

int wait_flags(const uint32_t* flags, int first, int n, uint32_t value, long long timeout, uint32_t* err,
               cudaStream_t st) {
  wait_kernel<<<1, 32, 0, st>>>(flags, first, n, value, timeout, err);
  return (int)cudaGetLastError();
}

}  // namespace pa
