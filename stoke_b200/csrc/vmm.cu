// vmm.cu -- peer-visible memory on the CUDA virtual-memory-management API, and NVSwitch multicast objects over it.
//
// Why: cudaMalloc + cudaIpc* memory cannot be bound to a multicast object; NVLS (multimem.ld_reduce / multimem.st, the
// in-switch reduction NCCL's NVLS transport uses under the reference's DDP path, stoke/extensions.py:207-215) needs
// physical allocations made with cuMemCreate.  One process per GPU, so the allocation handles travel between processes as
// POSIX file descriptors: the owning context runs a tiny server thread on an abstract unix socket
// ("stk_b200.<pid>.<ctx serial>") that hands a descriptor to whoever asks for it by number (SCM_RIGHTS); the 64-byte blob
// the caller exchanges (same slot the cudaIpc handle uses in ipc mode) only carries {pid, serial, fd numbers, sizes}.
//
// The driver entry points are resolved at run time through cudaGetDriverEntryPoint, so the library has no link-time
// dependency on libcuda (it must load on a machine without a driver: the CPU test tier checks its exports).
#include <cuda.h>

#include <cstring>

#include "ctx.cuh"
#include "fdpass.h"

namespace {

struct Driver {
  bool tried = false, ok = false;
  decltype(&cuMemCreate) MemCreate = nullptr;
  decltype(&cuMemRelease) MemRelease = nullptr;
  decltype(&cuMemAddressReserve) MemAddressReserve = nullptr;
  decltype(&cuMemAddressFree) MemAddressFree = nullptr;
  decltype(&cuMemMap) MemMap = nullptr;
  decltype(&cuMemUnmap) MemUnmap = nullptr;
  decltype(&cuMemSetAccess) MemSetAccess = nullptr;
  decltype(&cuMemGetAllocationGranularity) MemGetAllocationGranularity = nullptr;
  decltype(&cuMemExportToShareableHandle) MemExportToShareableHandle = nullptr;
  decltype(&cuMemImportFromShareableHandle) MemImportFromShareableHandle = nullptr;
  decltype(&cuMulticastCreate) MulticastCreate = nullptr;
  decltype(&cuMulticastAddDevice) MulticastAddDevice = nullptr;
  decltype(&cuMulticastBindMem) MulticastBindMem = nullptr;
  decltype(&cuMulticastUnbind) MulticastUnbind = nullptr;
  decltype(&cuMulticastGetGranularity) MulticastGetGranularity = nullptr;
  decltype(&cuDeviceGetAttribute) DeviceGetAttribute = nullptr;
  decltype(&cuDeviceGet) DeviceGet = nullptr;
  decltype(&cuGetErrorString) GetErrorString = nullptr;
  bool mc_entry_points = false;
};
Driver g_drv;
std::mutex g_drv_mu;

template <typename F>
bool resolve(const char* name, F& fn) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  fn = reinterpret_cast<F>(p);
  return true;
}

Driver& drv() {
  std::lock_guard<std::mutex> lk(g_drv_mu);
  if (!g_drv.tried) {
    g_drv.tried = true;
    Driver& d = g_drv;
    bool ok = resolve("cuMemCreate", d.MemCreate) && resolve("cuMemRelease", d.MemRelease) &&
              resolve("cuMemAddressReserve", d.MemAddressReserve) && resolve("cuMemAddressFree", d.MemAddressFree) &&
              resolve("cuMemMap", d.MemMap) && resolve("cuMemUnmap", d.MemUnmap) && resolve("cuMemSetAccess", d.MemSetAccess) &&
              resolve("cuMemGetAllocationGranularity", d.MemGetAllocationGranularity) &&
              resolve("cuMemExportToShareableHandle", d.MemExportToShareableHandle) &&
              resolve("cuMemImportFromShareableHandle", d.MemImportFromShareableHandle) &&
              resolve("cuDeviceGetAttribute", d.DeviceGetAttribute) && resolve("cuDeviceGet", d.DeviceGet) &&
              resolve("cuGetErrorString", d.GetErrorString);
    d.ok = ok;
    d.mc_entry_points = ok && resolve("cuMulticastCreate", d.MulticastCreate) &&
                        resolve("cuMulticastAddDevice", d.MulticastAddDevice) &&
                        resolve("cuMulticastBindMem", d.MulticastBindMem) && resolve("cuMulticastUnbind", d.MulticastUnbind) &&
                        resolve("cuMulticastGetGranularity", d.MulticastGetGranularity);
  }
  return g_drv;
}

std::string cu_err(CUresult r) {
  const char* s = nullptr;
  if (g_drv.GetErrorString && g_drv.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "CUresult " + std::to_string((int)r);
}

#define STK_CU(ctx, call)                                                                         \
  do {                                                                                            \
    CUresult r__ = (call);                                                                        \
    if (r__ != CUDA_SUCCESS) return stk_fail(ctx, STK_ERR_CUDA, std::string(#call) + ": " + cu_err(r__)); \
  } while (0)

// ---- the 64-byte blob exchanged by the caller --------------------------------------------------------------------------
struct VmmBlob {
  uint32_t magic;     // 'STKV'
  int32_t pid;
  int32_t serial;
  int32_t mem_fd;     // descriptor number in the exporting process
  int32_t mc_fd;      // rank 0: descriptor of the multicast object, else -1
  uint32_t pad_;
  uint64_t bytes;     // mapped size
  uint64_t reserved[4];
};
static_assert(sizeof(VmmBlob) <= STK_IPC_HANDLE_BYTES, "blob must fit the handle slot");
constexpr uint32_t kMagic = 0x564b5453u;

size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace

struct FdServer : stk_fd::FdServer {};
using stk_fd::fetch_fd;

namespace {

CUmemAllocationProp mem_prop(int device) {
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

int map_handle(stk_ctx* c, CUmemGenericAllocationHandle h, size_t bytes, size_t align, void** out) {
  Driver& d = drv();
  CUdeviceptr va = 0;
  STK_CU(c, d.MemAddressReserve(&va, bytes, align, 0, 0));
  CUresult r = d.MemMap(va, bytes, 0, h, 0);
  if (r != CUDA_SUCCESS) {
    d.MemAddressFree(va, bytes);
    return stk_fail(c, STK_ERR_CUDA, "cuMemMap: " + cu_err(r));
  }
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = c->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.MemSetAccess(va, bytes, &acc, 1);
  if (r != CUDA_SUCCESS) {
    d.MemUnmap(va, bytes);
    d.MemAddressFree(va, bytes);
    return stk_fail(c, STK_ERR_CUDA, "cuMemSetAccess: " + cu_err(r));
  }
  *out = reinterpret_cast<void*>(va);
  return STK_OK;
}

void unmap(void* p, size_t bytes) {
  if (!p) return;
  Driver& d = drv();
  d.MemUnmap(reinterpret_cast<CUdeviceptr>(p), bytes);
  d.MemAddressFree(reinterpret_cast<CUdeviceptr>(p), bytes);
}

size_t g_gran_cache[64] = {};

size_t granularity(stk_ctx* c) {
  Driver& d = drv();
  if (c->device < 64 && g_gran_cache[c->device]) return g_gran_cache[c->device];
  CUmemAllocationProp prop = mem_prop(c->device);
  size_t g = 0;
  if (d.MemGetAllocationGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || g == 0) g = 2u << 20;
  if (c->multicast_ok) {
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)c->world;
    mp.size = g;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (d.MulticastGetGranularity(&mg, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) == CUDA_SUCCESS && mg > g) g = round_up(mg, g);
  }
  if (c->device < 64) g_gran_cache[c->device] = g;
  return g;
}

}  // namespace

bool stk_vmm_available(int device, bool* multicast) {
  Driver& d = drv();
  if (multicast) *multicast = false;
  if (!d.ok) return false;
  CUdevice dev;
  if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return false;
  int vmm = 0, fdh = 0, mc = 0;
  d.DeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev);
  d.DeviceGetAttribute(&fdh, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, dev);
  if (d.mc_entry_points) d.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev);
  if (multicast) *multicast = d.mc_entry_points && mc != 0;
  return vmm != 0 && fdh != 0;
}

int stk_vmm_alloc(stk_ctx* c, size_t bytes, void** local_ptr, unsigned char* handle_out) {
  Driver& d = drv();
  if (!d.ok) return stk_fail(c, STK_ERR_UNSUPPORTED, "CUDA driver VMM entry points are not available");
  if (!c->fd_server) {
    static std::atomic<int> g_serial{0};
    c->serial = ++g_serial;
    c->fd_server = new FdServer();
    if (!c->fd_server->start(c->serial)) {
      delete c->fd_server;
      c->fd_server = nullptr;
      return stk_fail(c, STK_ERR_CUDA, "could not start the descriptor server (unix socket)");
    }
  }
  const size_t gran = granularity(c);
  const size_t size = round_up(bytes, gran);
  CUmemAllocationProp prop = mem_prop(c->device);
  CUmemGenericAllocationHandle h = 0;
  STK_CU(c, d.MemCreate(&h, size, &prop, 0));
  void* p = nullptr;
  int rc = map_handle(c, h, size, gran, &p);
  if (rc != STK_OK) {
    d.MemRelease(h);
    return rc;
  }
  cudaError_t e = cudaMemset(p, 0, size);
  if (e != cudaSuccess) {
    unmap(p, size);
    d.MemRelease(h);
    return stk_fail(c, STK_ERR_CUDA, std::string("cudaMemset (vmm): ") + cudaGetErrorString(e));
  }
  cudaDeviceSynchronize();
  int fd = -1;
  CUresult r = d.MemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    unmap(p, size);
    d.MemRelease(h);
    return stk_fail(c, STK_ERR_CUDA, "cuMemExportToShareableHandle: " + cu_err(r));
  }
  stk_ctx::Shared s{};
  s.bytes = size;
  s.vmm = true;
  s.mem_handle = h;
  s.mem_fd = fd;
  s.peers[c->rank] = p;
  // rank 0 creates the multicast object of this buffer; a failure only means "no NVLS for this buffer"
  if (c->multicast_ok && c->rank == 0) {
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)c->world;
    mp.size = size;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    CUmemGenericAllocationHandle mc = 0;
    if (d.MulticastCreate(&mc, &mp) == CUDA_SUCCESS) {
      int mfd = -1;
      if (d.MemExportToShareableHandle(&mfd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0) == CUDA_SUCCESS) {
        s.mc_handle = mc;
        s.mc_fd = mfd;
      } else {
        d.MemRelease(mc);
      }
    }
  }
  {
    std::lock_guard<std::mutex> lk(c->fd_server->mu);
    c->fd_server->exported.insert(fd);
    if (s.mc_fd >= 0) c->fd_server->exported.insert(s.mc_fd);
  }
  VmmBlob b{};
  b.magic = kMagic;
  b.pid = (int32_t)getpid();
  b.serial = c->serial;
  b.mem_fd = fd;
  b.mc_fd = s.mc_fd;
  b.bytes = size;
  std::memset(handle_out, 0, STK_IPC_HANDLE_BYTES);
  std::memcpy(handle_out, &b, sizeof(b));
  c->shared[p] = s;
  *local_ptr = p;
  return STK_OK;
}

int stk_vmm_open(stk_ctx* c, stk_ctx::Shared& sh, const unsigned char* handles) {
  Driver& d = drv();
  const size_t gran = granularity(c);
  for (int r = 0; r < c->world; ++r) {
    VmmBlob b;
    std::memcpy(&b, handles + size_t(r) * STK_IPC_HANDLE_BYTES, sizeof(b));
    if (b.magic != kMagic) return stk_fail(c, STK_ERR_STATE, "stk_mem_open_peers: a peer's handle is not a vmm handle (ranks disagree on STK_MEM)");
    if (b.bytes != sh.bytes) return stk_fail(c, STK_ERR_INVALID, "stk_mem_open_peers: peers allocated different sizes");
    std::string why;
    if (r != c->rank) {
      int fd = fetch_fd(b.pid, b.serial, b.mem_fd, why);
      if (fd < 0) return stk_fail(c, STK_ERR_CUDA, "stk_mem_open_peers: " + why);
      CUmemGenericAllocationHandle h = 0;
      CUresult rr = d.MemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                                   CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
      close(fd);
      if (rr != CUDA_SUCCESS) return stk_fail(c, STK_ERR_CUDA, "cuMemImportFromShareableHandle: " + cu_err(rr));
      void* p = nullptr;
      int rc = map_handle(c, h, sh.bytes, gran, &p);
      if (rc != STK_OK) {
        d.MemRelease(h);
        return rc;
      }
      sh.peer_handles[r] = h;
      sh.peers[r] = p;
    }
    // multicast object: created by rank 0, imported by everybody else; every rank adds its device
    if (r == 0 && b.mc_fd >= 0 && c->multicast_ok) {
      CUmemGenericAllocationHandle mc = sh.mc_handle;
      if (c->rank != 0) {
        int fd = fetch_fd(b.pid, b.serial, b.mc_fd, why);
        if (fd >= 0) {
          if (d.MemImportFromShareableHandle(&mc, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)),
                                             CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR) != CUDA_SUCCESS)
            mc = 0;
          close(fd);
        }
      }
      if (mc) {
        CUdevice dev;
        if (d.DeviceGet(&dev, c->device) == CUDA_SUCCESS && d.MulticastAddDevice(mc, dev) == CUDA_SUCCESS) {
          sh.mc_handle = mc;
          sh.mc_added = true;
        } else {
          if (c->rank != 0) d.MemRelease(mc);
          if (c->rank != 0) sh.mc_handle = 0;
        }
      }
    }
  }
  return STK_OK;
}

int stk_vmm_mc_bind(stk_ctx* c, stk_ctx::Shared& sh) {
  Driver& d = drv();
  if (!sh.vmm || !sh.mc_handle || !sh.mc_added)
    return stk_fail(c, STK_ERR_UNSUPPORTED, "no multicast object for this buffer (device/driver without NVLS, or ipc memory mode)");
  if (sh.mc_ptr) return STK_OK;
  CUresult r = d.MulticastBindMem(sh.mc_handle, 0, sh.mem_handle, 0, sh.bytes, 0);
  if (r != CUDA_SUCCESS) return stk_fail(c, STK_ERR_UNSUPPORTED, "cuMulticastBindMem: " + cu_err(r));
  sh.mc_bound = true;
  void* p = nullptr;
  int rc = map_handle(c, sh.mc_handle, sh.bytes, granularity(c), &p);
  if (rc != STK_OK) return STK_ERR_UNSUPPORTED;
  sh.mc_ptr = p;
  return STK_OK;
}

int stk_vmm_mc_release(stk_ctx* c, stk_ctx::Shared& sh) {
  Driver& d = drv();
  if (sh.mc_ptr) {
    unmap(sh.mc_ptr, sh.bytes);
    sh.mc_ptr = nullptr;
  }
  if (sh.mc_bound) {
    CUdevice dev;
    if (d.DeviceGet(&dev, c->device) == CUDA_SUCCESS) d.MulticastUnbind(sh.mc_handle, dev, 0, sh.bytes);
    sh.mc_bound = false;
  }
  return STK_OK;
}

int stk_vmm_free(stk_ctx* c, void* local_ptr, stk_ctx::Shared& sh) {
  Driver& d = drv();
  stk_vmm_mc_release(c, sh);
  if (sh.mc_handle) {
    d.MemRelease(sh.mc_handle);
    sh.mc_handle = 0;
  }
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank || !sh.peers[r]) continue;
    unmap(sh.peers[r], sh.bytes);
    d.MemRelease(sh.peer_handles[r]);
    sh.peers[r] = nullptr;
  }
  unmap(local_ptr, sh.bytes);
  d.MemRelease(sh.mem_handle);
  if (c->fd_server) {
    std::lock_guard<std::mutex> lk(c->fd_server->mu);
    c->fd_server->exported.erase(sh.mem_fd);
    if (sh.mc_fd >= 0) c->fd_server->exported.erase(sh.mc_fd);
  }
  if (sh.mem_fd >= 0) close(sh.mem_fd);
  if (sh.mc_fd >= 0) close(sh.mc_fd);
  return STK_OK;
}

void stk_vmm_ctx_shutdown(stk_ctx* c) {
  if (c->fd_server) {
    c->fd_server->shutdown();
    delete c->fd_server;
    c->fd_server = nullptr;
  }
}

void* stk_mc_lookup(stk_ctx* c, const void* local) {
  auto it = c->shared.upper_bound(const_cast<void*>(local));
  if (it == c->shared.begin()) return nullptr;
  --it;
  const char* base = static_cast<const char*>(it->first);
  const char* q = static_cast<const char*>(local);
  if (q < base || q >= base + it->second.bytes || !it->second.mc_ptr) return nullptr;
  return static_cast<char*>(it->second.mc_ptr) + (q - base);
}
