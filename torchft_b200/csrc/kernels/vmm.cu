// CUDA VMM + NVLS multicast plumbing for symmetric memory (host side).
//
// Why: legacy cudaIpc handles (a) cannot back an NVLS multicast object and (b) leave
// the fate of an importer's mapping to the exporter's lifetime. VMM allocations
// (cuMemCreate) are refcounted physical memory: a peer that imported the handle keeps
// it alive even if the exporter is SIGKILLed, which is exactly what a fault-tolerant
// data plane wants, and they can be bound to a multicast object so kernels can issue
// multimem.ld_reduce / multimem.st (in-switch reduction / broadcast over NVSwitch).
//
// Handles travel between processes as POSIX file descriptors over SCM_RIGHTS on an
// abstract unix socket served by a tiny thread in every process (no ptrace/pidfd needed).
// All driver entry points are resolved at runtime (cudaGetDriverEntryPoint), so the
// extension still imports on a machine without libcuda.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <sys/socket.h>
#include <sys/un.h>
#include <unistd.h>

#include <atomic>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>

namespace py = pybind11;

namespace {

#define DRV(name) reinterpret_cast<decltype(&name)>(entry(#name))

void* entry(const char* name) {
  static std::mutex mu;
  static std::map<std::string, void*> cache;
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(name);
  if (it != cache.end()) return it->second;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
  if (e != cudaSuccess || fn == nullptr)
    throw std::runtime_error(std::string("driver entry point unavailable: ") + name + " (" + cudaGetErrorString(e) + ")");
  cache[name] = fn;
  return fn;
}

void ck(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) {
    const char* s = nullptr;
    auto f = DRV(cuGetErrorString);
    f(r, &s);
    throw std::runtime_error(std::string(what) + " failed: " + (s ? s : "?") + " (" + std::to_string((int)r) + ")");
  }
}

int current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) throw std::runtime_error("cudaGetDevice failed");
  cudaFree(0);  // make sure the primary context exists
  return d;
}

CUmemAllocationProp alloc_prop(int dev) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = dev;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

void map_rw(CUdeviceptr va, size_t size, CUmemGenericAllocationHandle h, int dev) {
  ck(DRV(cuMemMap)(va, size, 0, h, 0), "cuMemMap");
  CUmemAccessDesc a{};
  a.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  a.location.id = dev;
  a.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  ck(DRV(cuMemSetAccess)(va, size, &a, 1), "cuMemSetAccess");
}

// ------------------------------------------------------------------ fd passing
class FdServer {
 public:
  FdServer() {
    static std::atomic<int> seq{0};
    name_ = "tft_b200_fd_" + std::to_string(getpid()) + "_" + std::to_string(seq++);
    fd_ = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (fd_ < 0) throw std::runtime_error("fd server: socket() failed");
    sockaddr_un sa{};
    sa.sun_family = AF_UNIX;
    std::memcpy(sa.sun_path + 1, name_.data(), name_.size());  // abstract namespace
    if (::bind(fd_, (sockaddr*)&sa, offsetof(sockaddr_un, sun_path) + 1 + name_.size()) != 0 || ::listen(fd_, 64) != 0)
      throw std::runtime_error("fd server: bind/listen failed");
    th_ = std::thread([this] { loop(); });
  }
  ~FdServer() {
    stop_ = true;
    ::shutdown(fd_, SHUT_RDWR);
    ::close(fd_);
    if (th_.joinable()) th_.join();
  }
  const std::string& name() const { return name_; }
  void publish(const std::string& key, int fd) {
    std::lock_guard<std::mutex> g(mu_);
    fds_[key] = fd;
  }
  void unpublish(const std::string& key) {
    std::lock_guard<std::mutex> g(mu_);
    fds_.erase(key);
  }

 private:
  void loop() {
    while (!stop_) {
      int c = ::accept4(fd_, nullptr, nullptr, SOCK_CLOEXEC);
      if (c < 0) {
        if (stop_) return;
        continue;
      }
      char key[256] = {0};
      ssize_t n = ::recv(c, key, sizeof(key) - 1, 0);
      int fd = -1;
      if (n > 0) {
        std::lock_guard<std::mutex> g(mu_);
        auto it = fds_.find(std::string(key, (size_t)n));
        if (it != fds_.end()) fd = it->second;
      }
      char ok = fd >= 0 ? 1 : 0;
      iovec iov{&ok, 1};
      msghdr msg{};
      msg.msg_iov = &iov;
      msg.msg_iovlen = 1;
      char ctrl[CMSG_SPACE(sizeof(int))] = {0};
      if (fd >= 0) {
        msg.msg_control = ctrl;
        msg.msg_controllen = sizeof(ctrl);
        cmsghdr* cm = CMSG_FIRSTHDR(&msg);
        cm->cmsg_level = SOL_SOCKET;
        cm->cmsg_type = SCM_RIGHTS;
        cm->cmsg_len = CMSG_LEN(sizeof(int));
        std::memcpy(CMSG_DATA(cm), &fd, sizeof(int));
      }
      ::sendmsg(c, &msg, MSG_NOSIGNAL);
      ::close(c);
    }
  }
  std::string name_;
  int fd_ = -1;
  std::thread th_;
  std::atomic<bool> stop_{false};
  std::mutex mu_;
  std::map<std::string, int> fds_;
};

int fetch_fd(const std::string& server, const std::string& key) {
  int c = ::socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
  if (c < 0) throw std::runtime_error("fetch_fd: socket() failed");
  sockaddr_un sa{};
  sa.sun_family = AF_UNIX;
  std::memcpy(sa.sun_path + 1, server.data(), server.size());
  if (::connect(c, (sockaddr*)&sa, offsetof(sockaddr_un, sun_path) + 1 + server.size()) != 0) {
    ::close(c);
    throw std::runtime_error("fetch_fd: cannot connect to fd server " + server + " (peer gone?)");
  }
  ::send(c, key.data(), key.size(), MSG_NOSIGNAL);
  char ok = 0;
  iovec iov{&ok, 1};
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  char ctrl[CMSG_SPACE(sizeof(int))] = {0};
  msg.msg_control = ctrl;
  msg.msg_controllen = sizeof(ctrl);
  ssize_t n = ::recvmsg(c, &msg, 0);
  ::close(c);
  if (n <= 0 || !ok) throw std::runtime_error("fetch_fd: peer has no handle named " + key);
  cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  if (!cm || cm->cmsg_type != SCM_RIGHTS) throw std::runtime_error("fetch_fd: no descriptor in reply");
  int fd = -1;
  std::memcpy(&fd, CMSG_DATA(cm), sizeof(int));
  return fd;
}

// ------------------------------------------------------------------ VMM segments
struct VmmAlloc {
  CUmemGenericAllocationHandle handle = 0;
  CUdeviceptr va = 0;
  size_t size = 0;
  int fd = -1;  // exported descriptor (owner only)
};

size_t vmm_granularity() {
  CUmemAllocationProp p = alloc_prop(current_device());
  size_t g = 0;
  ck(DRV(cuMemGetAllocationGranularity)(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
  return g;
}

// returns (va, size, handle, fd)
py::tuple vmm_alloc(size_t nbytes) {
  const int dev = current_device();
  CUmemAllocationProp p = alloc_prop(dev);
  const size_t g = vmm_granularity();
  const size_t size = (nbytes + g - 1) / g * g;
  VmmAlloc a;
  a.size = size;
  ck(DRV(cuMemCreate)(&a.handle, size, &p, 0), "cuMemCreate");
  ck(DRV(cuMemAddressReserve)(&a.va, size, g, 0, 0), "cuMemAddressReserve");
  map_rw(a.va, size, a.handle, dev);
  if (cudaMemset((void*)a.va, 0, size) != cudaSuccess) throw std::runtime_error("cudaMemset on VMM segment failed");
  ck(DRV(cuMemExportToShareableHandle)(&a.fd, a.handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
     "cuMemExportToShareableHandle");
  return py::make_tuple((uintptr_t)a.va, size, (uint64_t)a.handle, a.fd);
}

// map a peer allocation received as fd; returns (va, handle)
py::tuple vmm_import(int fd, size_t size) {
  const int dev = current_device();
  CUmemGenericAllocationHandle h = 0;
  ck(DRV(cuMemImportFromShareableHandle)(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
     "cuMemImportFromShareableHandle");
  ::close(fd);
  CUdeviceptr va = 0;
  ck(DRV(cuMemAddressReserve)(&va, size, vmm_granularity(), 0, 0), "cuMemAddressReserve");
  map_rw(va, size, h, dev);
  return py::make_tuple((uintptr_t)va, (uint64_t)h);
}

void vmm_unmap(uintptr_t va, size_t size, uint64_t handle) {
  DRV(cuMemUnmap)((CUdeviceptr)va, size);
  DRV(cuMemAddressFree)((CUdeviceptr)va, size);
  DRV(cuMemRelease)((CUmemGenericAllocationHandle)handle);
}

// ------------------------------------------------------------------ multicast
bool multicast_supported() {
  try {
    int dev = current_device();
    CUdevice d;
    ck(DRV(cuDeviceGet)(&d, dev), "cuDeviceGet");
    int v = 0;
    ck(DRV(cuDeviceGetAttribute)(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d), "cuDeviceGetAttribute");
    return v != 0;
  } catch (...) {
    return false;
  }
}

CUmulticastObjectProp mc_prop(int ndev, size_t size) {
  CUmulticastObjectProp p{};
  p.numDevices = (unsigned)ndev;
  p.size = size;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  p.flags = 0;
  return p;
}

size_t mc_granularity(int ndev, size_t size) {
  CUmulticastObjectProp p = mc_prop(ndev, size);
  size_t g = 0;
  ck(DRV(cuMulticastGetGranularity)(&g, &p, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
  return g;
}

// creator: returns (mc_handle, fd)
py::tuple mc_create(int ndev, size_t size) {
  CUmulticastObjectProp p = mc_prop(ndev, size);
  CUmemGenericAllocationHandle mc = 0;
  ck(DRV(cuMulticastCreate)(&mc, &p), "cuMulticastCreate");
  int fd = -1;
  ck(DRV(cuMemExportToShareableHandle)(&fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
     "cuMemExportToShareableHandle(multicast)");
  return py::make_tuple((uint64_t)mc, fd);
}

uint64_t mc_import(int fd) {
  CUmemGenericAllocationHandle mc = 0;
  ck(DRV(cuMemImportFromShareableHandle)(&mc, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
     "cuMemImportFromShareableHandle(multicast)");
  ::close(fd);
  return (uint64_t)mc;
}

void mc_add_device(uint64_t mc) {
  CUdevice d;
  ck(DRV(cuDeviceGet)(&d, current_device()), "cuDeviceGet");
  ck(DRV(cuMulticastAddDevice)((CUmemGenericAllocationHandle)mc, d), "cuMulticastAddDevice");
}

// bind local physical memory and map the multicast object; returns the multicast VA
uintptr_t mc_bind_and_map(uint64_t mc, uint64_t mem_handle, size_t size) {
  const int dev = current_device();
  ck(DRV(cuMulticastBindMem)((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem_handle, 0, size, 0),
     "cuMulticastBindMem");
  CUdeviceptr va = 0;
  ck(DRV(cuMemAddressReserve)(&va, size, mc_granularity(2, size), 0, 0), "cuMemAddressReserve(multicast)");
  map_rw(va, size, (CUmemGenericAllocationHandle)mc, dev);
  return (uintptr_t)va;
}

void mc_release(uint64_t mc, uintptr_t va, size_t size) {
  if (va) {
    DRV(cuMemUnmap)((CUdeviceptr)va, size);
    DRV(cuMemAddressFree)((CUdeviceptr)va, size);
  }
  CUdevice d;
  if (DRV(cuDeviceGet)(&d, current_device()) == CUDA_SUCCESS) DRV(cuMulticastUnbind)((CUmemGenericAllocationHandle)mc, d, 0, size);
  DRV(cuMemRelease)((CUmemGenericAllocationHandle)mc);
}

}  // namespace

void bind_vmm(py::module_& m) {
  py::class_<FdServer>(m, "FdServer")
      .def(py::init<>())
      .def("name", &FdServer::name)
      .def("publish", &FdServer::publish)
      .def("unpublish", &FdServer::unpublish);
  m.def("fetch_fd", &fetch_fd, py::call_guard<py::gil_scoped_release>());
  m.def("vmm_granularity", &vmm_granularity);
  m.def("vmm_alloc", &vmm_alloc);
  m.def("vmm_import", &vmm_import);
  m.def("vmm_unmap", &vmm_unmap);
  m.def("multicast_supported", &multicast_supported);
  m.def("mc_granularity", &mc_granularity);
  m.def("mc_create", &mc_create);
  m.def("mc_import", &mc_import);
  m.def("mc_add_device", &mc_add_device);
  m.def("mc_bind_and_map", &mc_bind_and_map);
  m.def("mc_release", &mc_release);
  m.def("close_fd", [](int fd) { ::close(fd); });
}
