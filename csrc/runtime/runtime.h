// Native device runtime (host side, no Python in the hot path):
//   * peer access + CUDA-IPC symmetric buffers (one cudaMalloc per rank, mapped into every peer so
//     kernels can ld/st remote HBM over NVLink),
//   * HostExecutor: one OS thread per GPU that replays captured CUDA graphs / launches without the GIL,
//   * pinned staging buffers and event timers for the end-to-end (H2D -> step -> D2H) path.
#pragma once
#include <pybind11/pybind11.h>

namespace pa {
namespace rt {
void bind(pybind11::module_& m);
}
}  // namespace pa
