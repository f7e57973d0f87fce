// NVSwitch multicast (NVLS) objects for the device runtime: one multicast object spanning the GPUs of a team, every
// GPU binds a same-size slab of physical memory to it, and a ``multimem.st`` through the multicast mapping lands in
// ALL slabs at once (the switch replicates).  Single-process teams (the in-process / ComfyUI engine) are created here
// directly; multi-process teams (one rank per GPU) exchange the object through POSIX file descriptors
// (``export_fd`` / ``from_fd``; python passes the fd over an AF_UNIX socket with SCM_RIGHTS).
//
// Driver-API entry points are resolved through cudaGetDriverEntryPoint, so the library keeps linking against
// libcudart only and simply reports "unsupported" on a driver / fabric without multicast.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../common/host.h"

namespace py = pybind11;

namespace pa {
namespace rt {

namespace {

struct Drv {
  PFN_cuMulticastCreate_v12010 mcCreate = nullptr;
  PFN_cuMulticastAddDevice_v12010 mcAddDevice = nullptr;
  PFN_cuMulticastBindMem_v12010 mcBindMem = nullptr;
  PFN_cuMulticastGetGranularity_v12010 mcGran = nullptr;
  PFN_cuMulticastUnbind_v12010 mcUnbind = nullptr;
  PFN_cuMemCreate_v10020 memCreate = nullptr;
  PFN_cuMemRelease_v10020 memRelease = nullptr;
  PFN_cuMemAddressReserve_v10020 vaReserve = nullptr;
  PFN_cuMemAddressFree_v10020 vaFree = nullptr;
  PFN_cuMemMap_v10020 memMap = nullptr;
  PFN_cuMemUnmap_v10020 memUnmap = nullptr;
  PFN_cuMemSetAccess_v10020 memSetAccess = nullptr;
  PFN_cuMemGetAllocationGranularity_v10020 memGran = nullptr;
  PFN_cuMemExportToShareableHandle_v10020 memExport = nullptr;
  PFN_cuMemImportFromShareableHandle_v10020 memImport = nullptr;
  PFN_cuDeviceGetAttribute_v2000 devAttr = nullptr;
  PFN_cuDeviceGet_v2000 devGet = nullptr;
  PFN_cuGetErrorString_v6000 errStr = nullptr;
  bool ok = false;
};

template <typename T>
bool load(const char* name, T* out) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p)
    return false;
  *out = reinterpret_cast<T>(p);
  return true;
}

Drv& drv() {
  static Drv d;
  static bool tried = false;
  if (!tried) {
    tried = true;
    cudaFree(nullptr);          // make sure a context / the driver is initialised
    d.ok = load("cuMulticastCreate", &d.mcCreate) && load("cuMulticastAddDevice", &d.mcAddDevice) &&
           load("cuMulticastBindMem", &d.mcBindMem) && load("cuMulticastGetGranularity", &d.mcGran) &&
           load("cuMulticastUnbind", &d.mcUnbind) && load("cuMemCreate", &d.memCreate) &&
           load("cuMemRelease", &d.memRelease) && load("cuMemAddressReserve", &d.vaReserve) &&
           load("cuMemAddressFree", &d.vaFree) && load("cuMemMap", &d.memMap) && load("cuMemUnmap", &d.memUnmap) &&
           load("cuMemSetAccess", &d.memSetAccess) && load("cuMemGetAllocationGranularity", &d.memGran) &&
           load("cuMemExportToShareableHandle", &d.memExport) &&
           load("cuMemImportFromShareableHandle", &d.memImport) && load("cuDeviceGetAttribute", &d.devAttr) &&
           load("cuDeviceGet", &d.devGet) && load("cuGetErrorString", &d.errStr);
  }
  return d;
}

void ck(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* s = nullptr;
  if (drv().errStr) drv().errStr(r, &s);
  throw std::runtime_error(std::string("[pa.rt.multicast] ") + what + ": " + (s ? s : "unknown driver error") + " (" +
                           std::to_string((int)r) + ")");
}

struct DevGuard {
  int prev = 0;
  explicit DevGuard(int dev) {
    cudaGetDevice(&prev);
    cudaSetDevice(dev);
    cudaFree(nullptr);          // primary context current on this thread
  }
  ~DevGuard() { cudaSetDevice(prev); }
};

CUmemAccessDesc rw(int dev) {
  CUmemAccessDesc a{};
  a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  a.location.id = dev;
  a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  return a;
}

}  // namespace

static bool multicast_supported(int dev) {
  Drv& d = drv();
  if (!d.ok) return false;
  CUdevice cd;
  if (d.devGet(&cd, dev) != CUDA_SUCCESS) return false;
  int v = 0;
  if (d.devAttr(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cd) != CUDA_SUCCESS) return false;
  return v != 0;
}

// One member (= one GPU of this process) of a multicast team.
struct McMember {
  int dev = -1;
  CUmemGenericAllocationHandle mem = 0;
  CUdeviceptr uc = 0, mc = 0;       // unicast view of the own slab / multicast view of the team
  bool bound = false;
};

class MulticastTeam {
 public:
  // Single-process team over ``devices``; ``shareable`` requests POSIX-fd exportable handles (multi-process teams).
  MulticastTeam(std::vector<int> devices, size_t bytes, int team_size, bool shareable)
      : team_size_(team_size > 0 ? team_size : (int)devices.size()), shareable_(shareable) {
    Drv& d = drv();
    if (!d.ok) throw std::runtime_error("[pa.rt.multicast] driver has no multicast entry points");
    for (int dev : devices)
      if (!multicast_supported(dev)) throw std::runtime_error("[pa.rt.multicast] device does not support multicast");
    CUmulticastObjectProp prop{};
    prop.numDevices = (unsigned)team_size_;
    prop.handleTypes = shareable ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : 0;
    prop.flags = 0;
    prop.size = bytes;
    size_t gran = 0;
    ck(d.mcGran(&gran, &prop, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
    gran_ = gran;
    size_ = (bytes + gran - 1) / gran * gran;
    prop.size = size_;
    ck(d.mcCreate(&mc_, &prop), "cuMulticastCreate");
    owner_ = true;
    for (int dev : devices) add_member(dev);
  }

  // Multi-process: import the team's object from a POSIX fd exported by the creating rank.
  static MulticastTeam* from_fd(int fd, int dev, size_t size, size_t gran) {
    Drv& d = drv();
    if (!d.ok) throw std::runtime_error("[pa.rt.multicast] driver has no multicast entry points");
    auto* t = new MulticastTeam();
    t->size_ = size;
    t->gran_ = gran;
    t->shareable_ = true;
    DevGuard g(dev);
    ck(d.memImport(&t->mc_, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
       "cuMemImportFromShareableHandle(multicast)");
    t->add_member(dev);
    return t;
  }

  int export_fd() {
    int fd = -1;
    ck(drv().memExport(&fd, mc_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
    return fd;
  }

  // All team members (in every process) must have been ADDED before any bind returns; python barriers in between.
  void bind_all() {
    Drv& d = drv();
    for (auto& m : members_) {
      if (m.bound) continue;
      DevGuard g(m.dev);
      CUmemAllocationProp ap{};
      ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
      ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
      ap.location.id = m.dev;
      ap.requestedHandleTypes = shareable_ ? CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR : CU_MEM_HANDLE_TYPE_NONE;
      size_t mg = 0;
      ck(d.memGran(&mg, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
      if (size_ % mg) throw std::runtime_error("[pa.rt.multicast] slab size is not a multiple of the allocation granularity");
      ck(d.memCreate(&m.mem, size_, &ap, 0), "cuMemCreate");
      ck(d.mcBindMem(mc_, 0, m.mem, 0, size_, 0), "cuMulticastBindMem");
      m.bound = true;
      CUmemAccessDesc acc = rw(m.dev);
      ck(d.vaReserve(&m.uc, size_, gran_, 0, 0), "cuMemAddressReserve(uc)");
      ck(d.memMap(m.uc, size_, 0, m.mem, 0), "cuMemMap(uc)");
      ck(d.memSetAccess(m.uc, size_, &acc, 1), "cuMemSetAccess(uc)");
      ck(d.vaReserve(&m.mc, size_, gran_, 0, 0), "cuMemAddressReserve(mc)");
      ck(d.memMap(m.mc, size_, 0, mc_, 0), "cuMemMap(mc)");
      ck(d.memSetAccess(m.mc, size_, &acc, 1), "cuMemSetAccess(mc)");
    }
  }

  uintptr_t mc_ptr(int slot) const { return (uintptr_t)members_.at(slot).mc; }
  uintptr_t uc_ptr(int slot) const { return (uintptr_t)members_.at(slot).uc; }
  size_t size() const { return size_; }
  size_t granularity() const { return gran_; }
  int members() const { return (int)members_.size(); }

  // lead's kernel: local ``src`` -> every slab at ``offset`` (multimem.st); stream-ordered
  void bcast(int slot, uintptr_t src, size_t offset, size_t bytes, uintptr_t stream) {
    if (offset + bytes > size_) throw std::runtime_error("[pa.rt.multicast] bcast beyond the slab");
    DevGuard g(members_.at(slot).dev);
    int r = multimem_bcast(reinterpret_cast<const void*>(src), reinterpret_cast<void*>(members_.at(slot).mc + offset),
                           (long long)bytes, reinterpret_cast<cudaStream_t>(stream));
    if (r != 0) throw std::runtime_error("[pa.rt.multicast] multimem_bcast launch failed: " + std::to_string(r));
  }

  void close() {
    Drv& d = drv();
    for (auto& m : members_) {
      DevGuard g(m.dev);
      cudaDeviceSynchronize();
      if (m.mc) { d.memUnmap(m.mc, size_); d.vaFree(m.mc, size_); m.mc = 0; }
      if (m.uc) { d.memUnmap(m.uc, size_); d.vaFree(m.uc, size_); m.uc = 0; }
      if (m.bound) {
        CUdevice cd;
        if (d.devGet(&cd, m.dev) == CUDA_SUCCESS) d.mcUnbind(mc_, cd, 0, size_);
        m.bound = false;
      }
      if (m.mem) { d.memRelease(m.mem); m.mem = 0; }
    }
    members_.clear();
    if (mc_) { d.memRelease(mc_); mc_ = 0; }
  }
  ~MulticastTeam() {
    try { close(); } catch (...) {}
  }

 private:
  MulticastTeam() = default;
  void add_member(int dev) {
    DevGuard g(dev);
    CUdevice cd;
    ck(drv().devGet(&cd, dev), "cuDeviceGet");
    ck(drv().mcAddDevice(mc_, cd), "cuMulticastAddDevice");
    McMember m;
    m.dev = dev;
    members_.push_back(m);
  }
  CUmemGenericAllocationHandle mc_ = 0;
  std::vector<McMember> members_;
  size_t size_ = 0, gran_ = 0;
  int team_size_ = 0;
  bool shareable_ = false, owner_ = false;
};

void bind_multicast(py::module_& m) {
  m.def("multicast_supported", &multicast_supported);
  py::class_<MulticastTeam>(m, "MulticastTeam")
      .def(py::init<std::vector<int>, size_t, int, bool>(), py::arg("devices"), py::arg("bytes"), py::arg("team_size") = 0,
           py::arg("shareable") = false)
      .def_static("from_fd", &MulticastTeam::from_fd, py::return_value_policy::take_ownership)
      .def("export_fd", &MulticastTeam::export_fd)
      .def("bind_all", &MulticastTeam::bind_all)
      .def("mc_ptr", &MulticastTeam::mc_ptr)
      .def("uc_ptr", &MulticastTeam::uc_ptr)
      .def("size", &MulticastTeam::size)
      .def("granularity", &MulticastTeam::granularity)
      .def("members", &MulticastTeam::members)
      .def("bcast", &MulticastTeam::bcast)
      .def("close", &MulticastTeam::close);
}

}  // namespace rt
}  // namespace pa
