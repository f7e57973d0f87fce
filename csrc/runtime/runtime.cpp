#include "runtime.h"

#include "host_executor.h"

#include <ATen/cuda/CUDAContext.h>
#include <cuda_runtime.h>
#include <pybind11/stl.h>
#include <torch/extension.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <queue>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace py = pybind11;

namespace pa {
namespace rt {

static void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) throw std::runtime_error(std::string("[pa.rt] ") + what + ": " + cudaGetErrorString(e));
}

// ------------------------------------------------------------------ peer access
static bool can_access_peer(int dev, int peer) {
  int ok = 0;
  ck(cudaDeviceCanAccessPeer(&ok, dev, peer), "cudaDeviceCanAccessPeer");
  return ok != 0;
}

static bool enable_peer_access(int dev, int peer) {
  if (dev == peer) return true;
  if (!can_access_peer(dev, peer)) return false;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(dev);
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  cudaSetDevice(cur);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return true;
  }
  ck(e, "cudaDeviceEnablePeerAccess");
  return true;
}

// ------------------------------------------------------------------ raw device memory + IPC
static uintptr_t dev_malloc(int dev, size_t bytes, bool zero) {
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(dev);
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaSuccess && zero) e = cudaMemset(p, 0, bytes);
  cudaSetDevice(cur);
  ck(e, "cudaMalloc");
  return reinterpret_cast<uintptr_t>(p);
}

static void dev_free(uintptr_t p) { cudaFree(reinterpret_cast<void*>(p)); }

static py::bytes ipc_get_handle(uintptr_t p) {
  cudaIpcMemHandle_t h;
  ck(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(p)), "cudaIpcGetMemHandle");
  return py::bytes(reinterpret_cast<const char*>(&h), sizeof(h));
}

static uintptr_t ipc_open_handle(int dev, const std::string& bytes) {
  if (bytes.size() != sizeof(cudaIpcMemHandle_t)) throw std::runtime_error("[pa.rt] bad IPC handle size");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, bytes.data(), sizeof(h));
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(dev);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  cudaSetDevice(cur);
  ck(e, "cudaIpcOpenMemHandle");
  return reinterpret_cast<uintptr_t>(p);
}

static void ipc_close_handle(uintptr_t p) { cudaIpcCloseMemHandle(reinterpret_cast<void*>(p)); }

// Non-owning uint8 tensor over raw device memory (symmetric buffers live outside the caching allocator
// so that their base pointer is IPC-exportable).
static at::Tensor tensor_from_ptr(uintptr_t p, int64_t nbytes, int dev) {
  auto opts = at::TensorOptions().dtype(at::kByte).device(at::kCUDA, dev);
  return at::from_blob(reinterpret_cast<void*>(p), {nbytes}, [](void*) {}, opts);
}

static uintptr_t host_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  ck(cudaHostAlloc(&p, bytes, cudaHostAllocPortable), "cudaHostAlloc");
  return reinterpret_cast<uintptr_t>(p);
}
static void host_free_pinned(uintptr_t p) { cudaFreeHost(reinterpret_cast<void*>(p)); }

static void memcpy_async(uintptr_t dst, uintptr_t src, size_t bytes, int kind, uintptr_t stream) {
  ck(cudaMemcpyAsync(reinterpret_cast<void*>(dst), reinterpret_cast<const void*>(src), bytes,
                     static_cast<cudaMemcpyKind>(kind), reinterpret_cast<cudaStream_t>(stream)),
     "cudaMemcpyAsync");
}

// ------------------------------------------------------------------ HostExecutor
// One worker thread per device (host_executor.h).  Work items are (graph_exec, stream) launches or stream waits; the
// Python caller enqueues and returns immediately, `sync()` joins all queues with the GIL released.  This is the launch
// path of the in-process engine once every replica holds a captured step graph (engine._data_parallel_fused /
// _forward_ulysses): no Python, no GIL in the per-GPU path.
class HostExecutor {
 public:
  explicit HostExecutor(std::vector<int> devices)
      : devices_(devices), core_(static_cast<int>(devices.size()), [devices](int slot) { cudaSetDevice(devices[slot]); }) {}

  void launch_graph(int slot, uintptr_t graph_exec, uintptr_t stream) {
    core_.submit(slot, [graph_exec, stream] {
      ck(cudaGraphLaunch(reinterpret_cast<cudaGraphExec_t>(graph_exec), reinterpret_cast<cudaStream_t>(stream)),
         "cudaGraphLaunch");
    });
  }
  void stream_sync(int slot, uintptr_t stream) {
    core_.submit(slot, [stream] { ck(cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)), "cudaStreamSynchronize"); });
  }
  void sync() {
    py::gil_scoped_release nogil;
    core_.sync();
  }
  void shutdown() { core_.shutdown(); }
  int size() const { return static_cast<int>(devices_.size()); }

 private:
  std::vector<int> devices_;
  HostExecutorCore core_;
};

void bind_multicast(py::module_& m);

void bind(py::module_& m) {
  bind_multicast(m);
  m.def("can_access_peer", &can_access_peer);
  m.def("enable_peer_access", &enable_peer_access);
  m.def("dev_malloc", &dev_malloc, py::arg("device"), py::arg("bytes"), py::arg("zero") = true);
  m.def("dev_free", &dev_free);
  m.def("ipc_get_handle", &ipc_get_handle);
  m.def("ipc_open_handle", &ipc_open_handle);
  m.def("ipc_close_handle", &ipc_close_handle);
  m.def("tensor_from_ptr", &tensor_from_ptr);
  m.def("host_alloc_pinned", &host_alloc_pinned);
  m.def("host_free_pinned", &host_free_pinned);
  m.def("memcpy_async", &memcpy_async);
  py::class_<HostExecutor>(m, "HostExecutor")
      .def(py::init<std::vector<int>>())
      .def("launch_graph", &HostExecutor::launch_graph)
      .def("stream_sync", &HostExecutor::stream_sync)
      .def("sync", &HostExecutor::sync)
      .def("shutdown", &HostExecutor::shutdown)
      .def("size", &HostExecutor::size);
}

}  // namespace rt
}  // namespace pa
