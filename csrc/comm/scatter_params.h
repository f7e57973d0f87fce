// Parameter block of the fused scatter + patch-embed kernel (shared with the host binding).
#pragma once
#include <cuda_bf16.h>

namespace pa {

struct ScatterEmbedParams {
  const __nv_bfloat16* x_src;     // [n, C, Hl, Wl] at the lead (peer mapping or local)
  const __nv_bfloat16* t_src;     // [n]
  const __nv_bfloat16* g_src;     // [n] or nullptr
  __nv_bfloat16* t_emb;           // [n, 256]
  __nv_bfloat16* g_emb;           // [n, 256]
  __nv_bfloat16* x_copy;          // optional local copy of the latent shard (for the Euler epilogue)
  __nv_bfloat16* out;             // X[:, Lt:, :]
  long long ldo, out_bstride;
  const __nv_bfloat16* bias;      // [N]
  int n, C, Hl, Wl, N, Li;
  float time_factor;
};

}  // namespace pa
