// Shared between the sequence-parallel all-to-all kernels and the host binding.
#pragma once
#include <stdint.h>

namespace pa {

constexpr int SP_MAX_RANKS = 8;      // flag words per exchange slot
constexpr int SP_MAX_SLOTS = 256;    // exchanges per step (FLUX: 57 blocks x 2 + 2)

// one 2-D copy: `rows` rows of `row_bytes` (multiple of 16) from a peer mapping to local memory
struct SpPullDesc {
  const uint8_t* src;
  uint8_t* dst;
  long long src_pitch, dst_pitch;
  int rows, row_bytes;
};

}  // namespace pa
