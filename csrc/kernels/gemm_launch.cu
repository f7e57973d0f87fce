// Host side of the tcgen05 GEMM: TMA descriptor construction + launch.
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <mutex>

#include "../common/host.h"
#include "gemm_tcgen05.cuh"
#include "gemm_2cta.cuh"

#include <cstdlib>

namespace pa {

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

// rank-N bf16/u8 tensor map with 128-byte swizzle.  dims/strides innermost first; strides[0] is implicit.
int make_tmap(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box, int elem_bytes, const uint32_t* elem_strides, bool swizzle128) {
  auto fn = encode_fn();
  if (!fn) return -1;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUtensorMapDataType dt = elem_bytes == 2   ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                           : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32
                                             : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = fn(out, dt, rank, const_cast<void*>(ptr), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[pa] cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu box %u %u)\n", (int)r, rank,
            (unsigned long long)dims[0], (unsigned long long)dims[1], box[0], box[1]);
    return -2;
  }
  return 0;
}

// SM budget of the calling thread (0 = whole device).  Two independent kernel chains (FLUX's txt and img streams
// between two attention calls) run CONCURRENTLY on disjoint SM sets when each persistent kernel is launched with
// only its share of the SMs: 8 CTA pairs for the 512-row txt GEMMs + 66 pairs for the 4096-row img GEMMs finish in 3
// waves where the same two GEMMs back to back take 3 + 1 (profiles/README.md "txt/img SM partition").
static thread_local int g_sm_limit = 0;

void set_sm_limit(int n) { g_sm_limit = n > 0 ? n : 0; }

static int real_sms() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (n[dev] == 0) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev];
}

int num_sms() {
  const int real = real_sms();
  if (g_sm_limit > 0 && g_sm_limit < real) return g_sm_limit & ~1;      // whole CTA pairs (TPCs)
  return real;
}

// Widest tile that (a) divides N without a ragged last tile where possible and (b) still yields at least one
// tile per SM.  160 exists for the UNet channel counts (320 / 960 / 1920) that 256 and 128 tile badly.
static int pick_bn(int N, int m_tiles) {
  const int sms = num_sms();
  const int cands[4] = {256, 160, 128, 64};
  for (int c : cands) {
    if (N % c) continue;
    if (m_tiles * (N / c) >= sms) return c;
  }
  for (int c : cands)
    if (N % c == 0) return c == 256 ? 128 : c;        // small problems: prefer more, smaller tiles
  return 64;
}

template <int BN>
static int launch(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int tiles, cudaStream_t st) {
  using Cfg = GemmCfg<BN>;
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  int grid = tiles < num_sms() ? tiles : num_sms();
  gemm_bf16_tcgen05_kernel<BN><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, p);
  return (int)cudaGetLastError();
}

// CTA-pair kernel (gemm_2cta.cuh): 256 x 256 tiles, one cluster of two CTAs per tile.
static int launch_2cta(const void* A, long long lda, long long a_bstride, const void* W, long long ldw, const GemmParams& p,
                       cudaStream_t st) {
  using Cfg = Gemm2CtaCfg;
  CUtensorMap ta, tb;
  {
    uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)p.rows, (uint64_t)p.batch};
    uint64_t str[3] = {2, (uint64_t)lda * 2, (uint64_t)(p.batch > 1 ? a_bstride : (long long)p.rows * lda) * 2};
    uint32_t box[3] = {64, 128, 1};
    if (make_tmap(&ta, A, 3, dims, str, box, 2, nullptr)) return -20;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[2] = {2, (uint64_t)ldw * 2};
    uint32_t box[2] = {64, 128};                       // each CTA of the pair loads half of the 256 B rows
    if (make_tmap(&tb, W, 2, dims, str, box, 2, nullptr)) return -21;
  }
  static bool attr_set[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_set[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    attr_set[dev] = true;
  }
  const int tiles = ((p.rows + 255) / 256) * p.batch * ((p.N + 255) / 256);
  int grid = 2 * tiles < num_sms() ? 2 * tiles : (num_sms() & ~1);
  gemm_bf16_2cta_kernel<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, p);
  return (int)cudaGetLastError();
}

// Large GEMMs whose N is a multiple of 256 use the CTA-pair kernel (1 437 vs 1 348 TFLOP/s at 18432 x 9216 x 3072,
// profiles/selfcheck_gemm_2cta_run34.txt); PA_GEMM_2CTA=0 turns that off, force_bn == 512 / 256 selects a kernel
// explicitly (numerics checks, A/B timing).
static bool use_2cta_default() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("PA_GEMM_2CTA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// A: [batch, rows, K] with row stride lda and batch stride a_bstride (elements); W: [N, K] row stride ldw.
int gemm_bf16(const void* A, long long lda, long long a_bstride, const void* W, long long ldw, GemmParams p,
              int force_bn, cudaStream_t st) {
  if (p.K % 8 || lda % 8 || ldw % 8 || a_bstride % 8) return -10;   // TMA: 16-byte strides
  if (p.N % 32) return -11;
  const int m_tiles = ((p.rows + 127) / 128) * p.batch;
  if (p.conv_taps == 0 &&
      (force_bn == 512 || (force_bn == 0 && use_2cta_default() && p.N % 256 == 0 && p.rows >= 256 &&
                           (long long)m_tiles * (p.N / 256) >= 2LL * num_sms())))
    return launch_2cta(A, lda, a_bstride, W, ldw, p, st);
  int bn = force_bn;
  if (bn == 0) {
    if (p.mode == EPI_QKV_ROPE) {
      bn = 256;
    } else if (p.N <= 64) {
      bn = 64;
    } else {
      bn = pick_bn(p.N, m_tiles);
      if ((p.mode == EPI_GEGLU || p.mode == EPI_SWIGLU) && bn == 160) bn = 128;   // the GEGLU epilogue walks 64-column [a|g] groups
    }
  }
  if (p.mode == EPI_QKV_ROPE && bn < 128) return -12;
  CUtensorMap ta, tb;
  if (p.conv_taps > 0) return -30;   // use conv_bf16()
  {
    uint64_t dims[3] = {(uint64_t)p.K, (uint64_t)p.rows, (uint64_t)p.batch};
    uint64_t str[3] = {2, (uint64_t)lda * 2, (uint64_t)(p.batch > 1 ? a_bstride : (long long)p.rows * lda) * 2};
    uint32_t box[3] = {64, 128, 1};
    if (make_tmap(&ta, A, 3, dims, str, box, 2, nullptr)) return -20;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[2] = {2, (uint64_t)ldw * 2};
    uint32_t box[2] = {64, (uint32_t)bn};
    if (make_tmap(&tb, W, 2, dims, str, box, 2, nullptr)) return -21;
  }
  const int tiles = m_tiles * ((p.N + bn - 1) / bn);
  switch (bn) {
    case 256: return launch<256>(ta, tb, p, tiles, st);
    case 160: return launch<160>(ta, tb, p, tiles, st);
    case 128: return launch<128>(ta, tb, p, tiles, st);
    case 64: return launch<64>(ta, tb, p, tiles, st);
  }
  return -13;
}

// Implicit-GEMM convolution on NHWC bf16:  x [N, H, W, Cin] (Cin % 8 == 0), w [Cout, taps, Cin_pad] with
// Cin_pad = 64 * ceil(Cin / 64) (zero padded), out [N, Ho*Wo, Cout] through the usual epilogues.
int conv_bf16(const void* x, int N, int H, int W, int Cin, const void* w, int taps, int stride, GemmParams p,
              cudaStream_t st) {
  if (Cin % 8 || p.N % 32 || (taps != 1 && taps != 9)) return -10;
  const int pad = taps == 9 ? 1 : 0;
  const int Ho = (H + 2 * pad - (taps == 9 ? 3 : 1)) / stride + 1;
  const int Wo = (W + 2 * pad - (taps == 9 ? 3 : 1)) / stride + 1;
  const int tw = Wo >= 16 ? 16 : (Wo >= 8 ? 8 : 4);
  const int th = 128 / tw;
  p.conv_taps = taps;
  p.conv_cblocks = (Cin + 63) / 64;
  p.conv_tw = tw; p.conv_th = th; p.conv_wo = Wo; p.conv_ho = Ho;
  p.conv_stride = stride; p.conv_pad = pad;
  p.conv_tiles_w = (Wo + tw - 1) / tw;
  p.conv_tiles_h = (Ho + th - 1) / th;
  p.rows = Ho * Wo;
  p.batch = N;
  p.K = taps * p.conv_cblocks * 64;
  const int m_tiles = p.conv_tiles_w * p.conv_tiles_h * N;
  const int bn = pick_bn(p.N, m_tiles);
  CUtensorMap ta, tb;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[4] = {2, (uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {64, (uint32_t)(tw * stride), (uint32_t)(th * stride), 1};
    uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
    if (stride > 1) {            // the box spans (t-1)*stride + 1 input elements
      box[1] = (uint32_t)((tw - 1) * stride + 1);
      box[2] = (uint32_t)((th - 1) * stride + 1);
    }
    if (make_tmap(&ta, x, 4, dims, str, box, 2, es)) return -20;
  }
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.N};
    uint64_t str[2] = {2, (uint64_t)p.K * 2};
    uint32_t box[2] = {64, (uint32_t)bn};
    if (make_tmap(&tb, w, 2, dims, str, box, 2, nullptr)) return -21;
  }
  const int tiles = m_tiles * ((p.N + bn - 1) / bn);
  switch (bn) {
    case 256: return launch<256>(ta, tb, p, tiles, st);
    case 160: return launch<160>(ta, tb, p, tiles, st);
    case 128: return launch<128>(ta, tb, p, tiles, st);
    case 64: return launch<64>(ta, tb, p, tiles, st);
  }
  return -13;
}

}  // namespace pa
