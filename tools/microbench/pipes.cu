// Per-SM throughput of the non-tensor pipes that the attention softmax uses (sm_100a):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes.bin pipes.cu && ./pipes.bin
// Every test runs 148 CTAs x 512 threads (4 warps per scheduler), 8 independent dependency chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048

template <int OP>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, float seed) {
  float a[8];
  unsigned long long p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = seed + i + threadIdx.x * 1e-3f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[i]), "f"(a[i] + 1.f));
  }
  const float c = seed * 0.5f, d = seed * 0.25f;
  unsigned long long c2, d2;
  asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(c));
  asm("mov.b64 %0, {%1, %1};" : "=l"(d2) : "f"(d));
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(d2));
      if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
      if (OP == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 4) {   // 4 FFMA : 1 MUFU
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
        if ((i & 3) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[(i + 4) & 7]));
      }
      if (OP == 5) {   // 2 FFMA2 : 1 MUFU
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(d2));
        if ((i & 1) == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      }
      if (OP == 6) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
      if (OP == 7) {
        uint32_t h;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a[i]), "f"(c));
        a[i] = __uint_as_float(h);
      }
      if (OP == 8) {
        int v = __float_as_int(a[i]);
        asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v) : "r"(8388608), "r"(__float_as_int(c)));
        a[i] = __int_as_float(v);
      }
      if (OP == 9) {   // 1 FFMA : 1 FFMA2
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(c), "f"(d));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(d2));
      }
      if (OP == 10) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
      if (OP == 11) {  // FMNMX + FFMA
        asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(c2), "l"(d2));
      }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
    s += a[i] + lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, double ops_per_iter_per_thread) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  k<OP><<<148, 512>>>(out, cyc, 1.0f);
  k<OP><<<148, 512>>>(out, cyc, 1.0f);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("%-28s %10.0f cycles  %7.2f thread-instr/clk/SM  (%.2f cycles per warp-instr per scheduler)\n", name, avg,
         ops_per_iter_per_thread * ITER * 512 / avg, avg / (ops_per_iter_per_thread * ITER * 4));
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run<0>("FFMA", 8);
  run<10>("FADD", 8);
  run<1>("FFMA2 (fma.rn.f32x2)", 8);
  run<2>("FADD2 (add.rn.f32x2)", 8);
  run<3>("MUFU.EX2", 8);
  run<6>("FMNMX3", 8);
  run<7>("F2FP.BF16.PACK", 8);
  run<8>("IMAD", 8);
  run<4>("4 FFMA : 1 MUFU", 10);
  run<5>("2 FFMA2 : 1 MUFU", 12);
  run<9>("1 FFMA : 1 FFMA2", 16);
  run<11>("1 FMNMX : 1 FFMA2", 16);
  return 0;
}
