// Sequence-parallel (Ulysses) all-to-all for batch == 1 over NVLink peer memory (SURVEY §2.3 "SP / Ulysses", §5
// "long-context"): every GPU owns L/N tokens for the linear layers and H/N heads for attention.  Between the two
// layouts the data moves with ONE pull kernel per GPU: after a device-side flag handshake it reads, straight out of the
// peers' HBM (ld.global on peer mappings, 16 bytes per lane), the slabs listed in a descriptor table and writes them
// into its own buffers.  No NCCL, no host synchronisation: epochs live in device memory, so the whole step (57 blocks x
// 2 exchanges) replays as one CUDA graph per GPU.
//
//   sp_signal  : "my data for exchange `slot` of this step is written"  -> st.release.sys of the step epoch into
//                every peer's flag word [slot][me]
//   sp_pull    : wait (ld.acquire.sys, bounded) until all peers signalled `slot` for this epoch, then run the 2-D copy
//                descriptors (peer -> local)
//   sp_epoch_inc: last node of the step graph: epoch += 1
#include <cuda_runtime.h>
#include <stdint.h>

#include "../common/host.h"
#include "../common/ptx.cuh"
#include "sp_params.h"

namespace pa {

__global__ void sp_signal_kernel(uint32_t* const* peer_flags, int n_peers, int slot, int me, const uint32_t* epoch) {
  __threadfence_system();
  const int i = threadIdx.x;
  if (i < n_peers) ptx::st_release_sys_u32(peer_flags[i] + slot * SP_MAX_RANKS + me, *epoch);
}

__global__ void sp_epoch_inc_kernel(uint32_t* epoch) { *epoch += 1; }

__global__ void __launch_bounds__(256) sp_pull_kernel(const SpPullDesc* __restrict__ descs, int n_desc,
                                                       const uint32_t* flags, int slot, int n_peers,
                                                       const uint32_t* epoch, long long timeout, uint32_t* err) {
  __shared__ int ok;
  if (threadIdx.x < 32) {
    const uint32_t e = *epoch;
    bool good = true;
    if (static_cast<int>(threadIdx.x) < n_peers) {
      const long long t0 = clock64();
      while (static_cast<int32_t>(ptx::ld_acquire_sys_u32(flags + slot * SP_MAX_RANKS + threadIdx.x) - e) < 0) {
        if (clock64() - t0 > timeout) {
          if (err) atomicExch(err, 0xDEAD1000u | static_cast<uint32_t>(slot & 0xFFF));
          good = false;
          break;
        }
        __nanosleep(40);
      }
    }
    good = __all_sync(0xffffffffu, good);
    if (threadIdx.x == 0) ok = good ? 1 : 0;
  }
  __syncthreads();
  if (!ok) return;                              // a dead peer: leave the buffers alone, the host sees the error word
  __threadfence_system();
  const int d = blockIdx.y;
  if (d >= n_desc) return;
  const SpPullDesc ds = descs[d];
  const int vpr = ds.row_bytes >> 4;                              // 16-byte vectors per row
  const int total = ds.rows * vpr;
  const int stride = gridDim.x * blockDim.x;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  // eight independent 16-byte peer loads in flight per thread before the first store: a peer load costs ~2 us of NVLink
  // round trip, so the copy is latency-bound unless every thread keeps several outstanding (first version: 1 -> 384 us
  // per exchange for 21 MB; the NVLink time of that payload is ~30 us)
  for (; i + 7 * stride < total; i += 8 * stride) {
    int4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = i + u * stride;
      const int r = j / vpr, c = j - r * vpr;
      v[u] = *reinterpret_cast<const int4*>(ds.src + r * ds.src_pitch + c * 16);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = i + u * stride;
      const int r = j / vpr, c = j - r * vpr;
      *reinterpret_cast<int4*>(ds.dst + r * ds.dst_pitch + c * 16) = v[u];
    }
  }
  for (; i < total; i += stride) {
    const int r = i / vpr, c = i - r * vpr;
    *reinterpret_cast<int4*>(ds.dst + r * ds.dst_pitch + c * 16) =
        *reinterpret_cast<const int4*>(ds.src + r * ds.src_pitch + c * 16);
  }
}

int sp_signal(uint32_t* const* peer_flags, int n_peers, int slot, int me, const uint32_t* epoch, cudaStream_t st) {
  if (n_peers > SP_MAX_RANKS || slot < 0 || slot >= SP_MAX_SLOTS) return -1;
  sp_signal_kernel<<<1, 32, 0, st>>>(peer_flags, n_peers, slot, me, epoch);
  return (int)cudaGetLastError();
}

int sp_pull(const SpPullDesc* descs, int n_desc, int blocks_per_desc, const uint32_t* flags, int slot, int n_peers,
            const uint32_t* epoch, long long timeout, uint32_t* err, cudaStream_t st) {
  if (n_peers > SP_MAX_RANKS || slot < 0 || slot >= SP_MAX_SLOTS || n_desc < 1) return -1;
  dim3 grid(blocks_per_desc, n_desc);
  sp_pull_kernel<<<grid, 256, 0, st>>>(descs, n_desc, flags, slot, n_peers, epoch, timeout, err);
  return (int)cudaGetLastError();
}

int sp_epoch_inc(uint32_t* epoch, cudaStream_t st) {
  sp_epoch_inc_kernel<<<1, 1, 0, st>>>(epoch);
  return (int)cudaGetLastError();
}

}  // namespace pa
