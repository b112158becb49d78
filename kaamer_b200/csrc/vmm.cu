// vmm.cu — shareable device allocations for the key-range shards of mode P.
//
// A shard's table and postings are read by the search kernels of the OTHER GPUs of the NVSwitch
// domain through NVLink.  They are therefore allocated with the CUDA virtual-memory-management
// API (cuMemCreate, 2 MiB pages) and exported as POSIX file descriptors; an importing process maps
// them with cuMemImportFromShareableHandle + cuMemMap + cuMemSetAccess.  The legacy cudaIpc*
// mapping is NOT used: measured on B200 (profiles/r1_peer_gather.log) random 8-byte probes into a
// 7 GB cudaIpcOpenMemHandle mapping run at 0.08 G probes/s (the importer's mapping thrashes the
// GPU MMU) against 10.8 G probes/s into a large-page peer mapping of the same memory.
//
// The driver API is resolved at run time through cudaGetDriverEntryPoint, so libkaamer_gpu.so does
// not link against libcuda (the library must load on machines without a driver: tests/test_abi.py).
#include <cuda.h>
#include <unistd.h>

#include "internal.cuh"

namespace kaamer {

namespace {
struct Driver {
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
  decltype(&cuGetErrorString) GetErrorString = nullptr;
  bool ok = false;
};
Driver g_drv;
std::once_flag g_drv_once;

template <class F>
bool resolve(const char *name, F *out) {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult st;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
    cudaGetLastError();
    return false;
  }
  *out = reinterpret_cast<F>(p);
  return true;
}

bool driver() {
  std::call_once(g_drv_once, [] {
    Driver &d = g_drv;
    d.ok = resolve("cuMemCreate", &d.MemCreate) && resolve("cuMemRelease", &d.MemRelease) &&
           resolve("cuMemAddressReserve", &d.MemAddressReserve) && resolve("cuMemAddressFree", &d.MemAddressFree) &&
           resolve("cuMemMap", &d.MemMap) && resolve("cuMemUnmap", &d.MemUnmap) &&
           resolve("cuMemSetAccess", &d.MemSetAccess) &&
           resolve("cuMemGetAllocationGranularity", &d.MemGetAllocationGranularity) &&
           resolve("cuMemExportToShareableHandle", &d.MemExportToShareableHandle) &&
           resolve("cuMemImportFromShareableHandle", &d.MemImportFromShareableHandle) &&
           resolve("cuGetErrorString", &d.GetErrorString);
  });
  if (!g_drv.ok) set_error("CUDA driver entry points of the virtual memory management API are not available");
  return g_drv.ok;
}

const char *drv_err(CUresult r) {
  const char *s = nullptr;
  if (g_drv.GetErrorString && g_drv.GetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "unknown driver error";
}
#define KDRV(call)                                                          \
  do {                                                                      \
    CUresult _r = (call);                                                   \
    if (_r != CUDA_SUCCESS) {                                               \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, drv_err(_r)); \
      return KAAMER_ERR_CUDA;                                               \
    }                                                                       \
  } while (0)

CUmemAllocationProp props_for(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}

CUmemAccessDesc access_for(int device) {
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  return acc;
}

int map_with_access(VmmAlloc *a, int device) {
  Driver &d = g_drv;
  CUdeviceptr va = 0;
  KDRV(d.MemAddressReserve(&va, a->bytes, a->granularity, 0, 0));
  CUresult r = d.MemMap(va, a->bytes, 0, (CUmemGenericAllocationHandle)a->handle, 0);
  if (r == CUDA_SUCCESS) {
    CUmemAccessDesc acc = access_for(device);
    r = d.MemSetAccess(va, a->bytes, &acc, 1);
    if (r != CUDA_SUCCESS) d.MemUnmap(va, a->bytes);
  }
  if (r != CUDA_SUCCESS) {
    d.MemAddressFree(va, a->bytes);
    set_error("mapping %zu bytes for device %d: %s", a->bytes, device, drv_err(r));
    return KAAMER_ERR_CUDA;
  }
  a->ptr = reinterpret_cast<void *>(va);
  return KAAMER_OK;
}
}  // namespace

int vmm_alloc(int device, size_t bytes, VmmAlloc *out) {
  *out = VmmAlloc{};
  if (!driver()) return KAAMER_ERR_CUDA;
  Driver &d = g_drv;
  CUmemAllocationProp p = props_for(device);
  size_t gran = 0;
  KDRV(d.MemGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  if (gran == 0) gran = (size_t)2 << 20;
  out->granularity = gran;
  out->bytes = (bytes + gran - 1) / gran * gran;
  if (out->bytes == 0) out->bytes = gran;
  CUmemGenericAllocationHandle hnd = 0;
  CUresult r = d.MemCreate(&hnd, out->bytes, &p, 0);
  if (r != CUDA_SUCCESS) {
    set_error("cuMemCreate(%zu bytes, device %d): %s", out->bytes, device, drv_err(r));
    *out = VmmAlloc{};
    return r == CUDA_ERROR_OUT_OF_MEMORY ? KAAMER_ERR_NOMEM : KAAMER_ERR_CUDA;
  }
  out->handle = (unsigned long long)hnd;
  out->device = device;
  int rc = map_with_access(out, device);
  if (rc != KAAMER_OK) {
    d.MemRelease(hnd);
    *out = VmmAlloc{};
  }
  return rc;
}

int vmm_export_fd(const VmmAlloc &a, int *fd) {
  *fd = -1;
  if (!a.ptr) {
    set_error("allocation is not shareable");
    return KAAMER_ERR_ARG;
  }
  if (!driver()) return KAAMER_ERR_CUDA;
  int f = -1;
  KDRV(g_drv.MemExportToShareableHandle(&f, (CUmemGenericAllocationHandle)a.handle,
                                        CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  *fd = f;
  return KAAMER_OK;
}

int vmm_import_fd(int fd, size_t bytes, int device, VmmAlloc *out) {
  *out = VmmAlloc{};
  if (!driver()) return KAAMER_ERR_CUDA;
  Driver &d = g_drv;
  CUmemGenericAllocationHandle hnd = 0;
  KDRV(d.MemImportFromShareableHandle(&hnd, (void *)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  CUmemAllocationProp p = props_for(device);
  size_t gran = 0;
  if (d.MemGetAllocationGranularity(&gran, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || gran == 0)
    gran = (size_t)2 << 20;
  out->granularity = gran;
  out->bytes = bytes;
  out->handle = (unsigned long long)hnd;
  out->device = device;
  out->imported = true;
  int rc = map_with_access(out, device);
  if (rc != KAAMER_OK) {
    d.MemRelease(hnd);
    *out = VmmAlloc{};
  }
  return rc;
}

int vmm_grant(void *ptr, size_t bytes, int device) {
  if (!driver()) return KAAMER_ERR_CUDA;
  CUmemAccessDesc acc = access_for(device);
  KDRV(g_drv.MemSetAccess((CUdeviceptr)(uintptr_t)ptr, bytes, &acc, 1));
  return KAAMER_OK;
}

void vmm_free(VmmAlloc *a) {
  if (a->ptr && g_drv.ok) {
    Driver &d = g_drv;
    d.MemUnmap((CUdeviceptr)(uintptr_t)a->ptr, a->bytes);
    d.MemAddressFree((CUdeviceptr)(uintptr_t)a->ptr, a->bytes);
    d.MemRelease((CUmemGenericAllocationHandle)a->handle);
  }
  *a = VmmAlloc{};
}

}  // namespace kaamer
