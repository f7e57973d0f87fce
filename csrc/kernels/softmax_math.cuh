// Packed-fp32 softmax math shared by the attention kernels: 3-input max, ex2, fma/add.rn.f32x2 and the FMA-pipe exp2
// (Cody-Waite range reduction + degree-3 minimax polynomial) that takes a share of the exponentials off the 16-lane MUFU.
#pragma once

#include <cuda_runtime.h>

namespace pa {
namespace smx {

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// 2^x for two packed floats on the FMA pipe.  x = n + f with n = round(x) (magic-number add: the low mantissa bits of
// x + 1.5 * 2^23 hold n), 2^f by a degree-3 minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5; P is rounded
// to bf16, 2^-9, right after), the exponent patched in with an integer add.  x must be <= ~100; very negative inputs are
// clamped (the result underflows to ~2^-126).
__device__ __forceinline__ void exp2_poly2(unsigned long long x2, float& r0, float& r1) {
  float x0, x1;
  unpack2(x2, x0, x1);
  x0 = fmaxf(x0, -126.0f);
  x1 = fmaxf(x1, -126.0f);
  const unsigned long long x = pack2(x0, x1);
  const unsigned long long xr = add2(x, pack2(12582912.0f, 12582912.0f));
  const unsigned long long nf = add2(xr, pack2(-12582912.0f, -12582912.0f));
  float n0, n1;
  unpack2(nf, n0, n1);
  const unsigned long long f = add2(x, pack2(-n0, -n1));
  unsigned long long p = pack2(0.05517167f, 0.05517167f);
  p = fma2(p, f, pack2(0.24261113f, 0.24261113f));
  p = fma2(p, f, pack2(0.69326097f, 0.69326097f));
  p = fma2(p, f, pack2(0.99992806f, 0.99992806f));
  float p0, p1, xr0, xr1;
  unpack2(p, p0, p1);
  unpack2(xr, xr0, xr1);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(xr0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(xr1) << 23));
}

}  // namespace smx
}  // namespace pa
