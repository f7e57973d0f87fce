// Layout / glue kernels for the convolutional executors (channels-last bf16 everywhere):
//   nchw_to_nhwc_pad      NCHW latent (possibly on a peer GPU) -> NHWC with channels padded to a multiple of 8
//   upsample_nearest2x    NHWC, 16-byte vectors
//   concat_channels       [B, HW, C1] ++ [B, HW, C2] (UNet skip connections)
//   unet_out_gather       the fused GATHER for UNet-family models: NHWC eps (padded channels) -> optional CFG
//                         (cond/uncond halves of the local batch) -> Euler update -> NCHW store into the lead
//                         GPU's buffer (peer mapping)
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../common/host.h"

namespace pa {

static inline long long min_ll(long long a, long long b) { return a < b ? a : b; }
static inline int grid_for(long long n) { return (int)min_ll((n + 255) / 256, 148 * 16); }

__global__ void nchw_to_nhwc_pad_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int B,
                                        int C, int HW, int Cpad) {
  const long long total = static_cast<long long>(B) * HW * Cpad;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    const long long r = i / Cpad;
    const int p = static_cast<int>(r % HW);
    const int b = static_cast<int>(r / HW);
    out[i] = c < C ? x[(static_cast<long long>(b) * C + c) * HW + p] : __float2bfloat16(0.f);
  }
}

int nchw_to_nhwc_pad(const void* x, void* out, int B, int C, int HW, int Cpad, cudaStream_t st) {
  const long long total = static_cast<long long>(B) * HW * Cpad;
  nchw_to_nhwc_pad_kernel<<<grid_for(total), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x),
                                                           static_cast<__nv_bfloat16*>(out), B, C, HW, Cpad);
  return (int)cudaGetLastError();
}

__global__ void upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int B, int H, int W, int Cv) {
  const int Ho = 2 * H, Wo = 2 * W;
  const long long total = static_cast<long long>(B) * Ho * Wo * Cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cv);
    long long r = i / Cv;
    const int wo = static_cast<int>(r % Wo);
    r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int b = static_cast<int>(r / Ho);
    out[i] = __ldg(x + ((static_cast<long long>(b) * H + (ho >> 1)) * W + (wo >> 1)) * Cv + c);
  }
}

int upsample_nearest2x_nhwc(const void* x, void* out, int B, int H, int W, int C, cudaStream_t st) {
  if (C % 8) return -1;
  const long long total = static_cast<long long>(B) * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(total), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(out), B, H, W,
                                                     C / 8);
  return (int)cudaGetLastError();
}

__global__ void concat_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                              long long rows, int c1v, int c2v) {
  const int cv = c1v + c2v;
  const long long total = rows * cv;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % cv);
    const long long r = i / cv;
    out[i] = c < c1v ? __ldg(a + r * c1v + c) : __ldg(b + r * c2v + (c - c1v));
  }
}

int concat_channels(const void* a, const void* b, void* out, long long rows, int C1, int C2, cudaStream_t st) {
  if (C1 % 8 || C2 % 8) return -1;
  const long long total = rows * ((C1 + C2) / 8);
  concat_kernel<<<grid_for(total), 256, 0, st>>>(static_cast<const uint4*>(a), static_cast<const uint4*>(b),
                                                 static_cast<uint4*>(out), rows, C1 / 8, C2 / 8);
  return (int)cudaGetLastError();
}

// eps: [Bl, HW, Cpad] NHWC model output of the LOCAL batch.  If `cfg_pairs` != 0 the local batch is laid out
// as [cond (n) | uncond (n)] (pairs kept on one rank, SURVEY §2.3 "CFG parallel") and n = Bl / 2 samples are
// produced, else n = Bl.  mode 0: store eps/denoise direction d;  mode 1: x_out = x + (s' - s) * d (Euler,
// eps-prediction: k-diffusion's to_d(x, sigma, x - sigma*eps) == eps).
__global__ void unet_out_gather_kernel(const __nv_bfloat16* __restrict__ eps, const __nv_bfloat16* __restrict__ x,
                                       __nv_bfloat16* __restrict__ x_out, const float* __restrict__ sigmas, int n,
                                       int C, int HW, int Cpad, int cfg_pairs, float cfg, int mode,
                                       long long out_sample_off) {
  const long long total = static_cast<long long>(n) * C * HW;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i % HW);
    const long long r = i / HW;
    const int c = static_cast<int>(r % C);
    const int b = static_cast<int>(r / C);
    float d = __bfloat162float(eps[(static_cast<long long>(b) * HW + p) * Cpad + c]);
    if (cfg_pairs) {
      const float u = __bfloat162float(eps[(static_cast<long long>(b + n) * HW + p) * Cpad + c]);
      d = u + cfg * (d - u);
    }
    if (mode == 1) d = __bfloat162float(x[i]) + (sigmas[2 * b + 1] - sigmas[2 * b]) * d;
    x_out[out_sample_off * C * HW + i] = __float2bfloat16(d);
  }
}

int unet_out_gather(const void* eps, const void* x, void* x_out, const void* sigmas, int n, int C, int HW, int Cpad,
                    int cfg_pairs, float cfg, int mode, long long out_sample_off, cudaStream_t st) {
  const long long total = static_cast<long long>(n) * C * HW;
  unet_out_gather_kernel<<<grid_for(total), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(eps), static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(x_out),
      static_cast<const float*>(sigmas), n, C, HW, Cpad, cfg_pairs, cfg, mode, out_sample_off);
  return (int)cudaGetLastError();
}

}  // namespace pa
