// multicast.cu -- NVSwitch multicast windows (cuMulticast* / NVLS): a buffer that every rank of the node holds a
// copy of, plus a MULTICAST mapping through which one store lands in all copies.  The fused compute + all-gather
// (scimlop_kron_mul_mc) stores each finished output tile ONCE with multimem.st; the switch replicates it, so the
// NVLink egress per GPU is one slab instead of (ranks - 1) slabs.
//
// Set-up is collective and needs one barrier in the middle (every device must have joined the multicast object
// before memory can be bound), so it is two calls:
//   scimlop_mc_open : rank 0 creates the object and exports a POSIX file descriptor; the other ranks import it
//                     (the descriptor travels between the processes over a Unix socket, SCM_RIGHTS -- the caller's
//                     job); every rank adds its device.
//   --- barrier across ranks ---
//   scimlop_mc_bind : allocate this rank's physical copy, bind it, map it twice (unicast = the local array,
//                     multicast = the broadcast alias).
//   --- barrier across ranks --- before the first multicast store.
#include <unistd.h>

#include "context.h"

namespace {

struct McApi {
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*DeviceGet)(CUdevice*, int);
  bool ok = false;
};

template <typename F>
bool load(F& f, const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p || q != cudaDriverEntryPointSuccess) return false;
  f = reinterpret_cast<F>(p);
  return true;
}

McApi& api() {
  static McApi a;
  if (!a.ok) {
    a.ok = load(a.MulticastCreate, "cuMulticastCreate") && load(a.MulticastAddDevice, "cuMulticastAddDevice") &&
           load(a.MulticastBindMem, "cuMulticastBindMem") && load(a.MulticastGetGranularity, "cuMulticastGetGranularity") &&
           load(a.MulticastUnbind, "cuMulticastUnbind") && load(a.MemCreate, "cuMemCreate") &&
           load(a.MemRelease, "cuMemRelease") && load(a.MemAddressReserve, "cuMemAddressReserve") &&
           load(a.MemAddressFree, "cuMemAddressFree") && load(a.MemMap, "cuMemMap") && load(a.MemUnmap, "cuMemUnmap") &&
           load(a.MemSetAccess, "cuMemSetAccess") && load(a.MemExportToShareableHandle, "cuMemExportToShareableHandle") &&
           load(a.MemImportFromShareableHandle, "cuMemImportFromShareableHandle") && load(a.DeviceGet, "cuDeviceGet");
  }
  return a;
}

#define SCIMLOP_CU(c, call)                                                                                    \
  do {                                                                                                         \
    CUresult r__ = (call);                                                                                     \
    if (r__ != CUDA_SUCCESS)                                                                                   \
      return scimlop_fail((c), SCIMLOP_ERR_CUDA, "%s failed (CUresult %d) (%s:%d)", #call, (int)r__, __FILE__, __LINE__); \
  } while (0)

CUmulticastObjectProp mc_prop(int nranks, size_t bytes) {
  CUmulticastObjectProp p;
  memset(&p, 0, sizeof(p));
  p.numDevices = (unsigned)nranks;
  p.size = bytes;
  p.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

}  // namespace

// What a window consists of (opaque to the caller; freed by scimlop_mc_free).
struct scimlop_mc_window {
  CUmemGenericAllocationHandle mc, phys;
  CUdeviceptr uc_ptr, mc_ptr;
  size_t bytes;
  int device;
  bool bound;
};

extern "C" int scimlop_mc_padded_bytes(scimlop_handle_t h, size_t bytes, int nranks, size_t* padded) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!padded || nranks < 1) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "mc_padded_bytes: bad argument");
  McApi& a = api();
  if (!a.ok) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "multicast driver entry points are not available");
  scimlop_device_guard guard(h->device);
  cudaFree(0);
  CUmulticastObjectProp p = mc_prop(nranks, bytes);
  size_t gran = 0;
  SCIMLOP_CU(h, a.MulticastGetGranularity(&gran, &p, CU_MULTICAST_GRANULARITY_RECOMMENDED));
  *padded = (bytes + gran - 1) / gran * gran;
  return SCIMLOP_OK;
}

extern "C" int scimlop_mc_open(scimlop_handle_t h, size_t padded_bytes, int nranks, int rank, int fd_in, int* fd_out,
                               void** window) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  if (!window || !fd_out) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "mc_open: null argument");
  McApi& a = api();
  if (!a.ok) return scimlop_fail(h, SCIMLOP_ERR_UNSUPPORTED, "multicast driver entry points are not available");
  scimlop_device_guard guard(h->device);
  cudaFree(0);
  scimlop_mc_window* w = new scimlop_mc_window();
  memset(w, 0, sizeof(*w));
  w->bytes = padded_bytes;
  w->device = h->device;
  *fd_out = -1;
  if (rank == 0) {
    CUmulticastObjectProp p = mc_prop(nranks, padded_bytes);
    SCIMLOP_CU(h, a.MulticastCreate(&w->mc, &p));
    int fd = -1;
    SCIMLOP_CU(h, a.MemExportToShareableHandle(&fd, w->mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
    *fd_out = fd;
  } else {
    SCIMLOP_CU(h, a.MemImportFromShareableHandle(&w->mc, (void*)(uintptr_t)fd_in, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  }
  CUdevice dev;
  SCIMLOP_CU(h, a.DeviceGet(&dev, h->device));
  SCIMLOP_CU(h, a.MulticastAddDevice(w->mc, dev));
  *window = w;
  return SCIMLOP_OK;
}

extern "C" int scimlop_mc_bind(scimlop_handle_t h, void* window, void** local_ptr, void** mc_ptr) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  scimlop_mc_window* w = static_cast<scimlop_mc_window*>(window);
  if (!w || !local_ptr || !mc_ptr) return scimlop_fail(h, SCIMLOP_ERR_NULL_POINTER, "mc_bind: null argument");
  McApi& a = api();
  scimlop_device_guard guard(h->device);
  CUmemAllocationProp prop;
  memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = h->device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  SCIMLOP_CU(h, a.MemCreate(&w->phys, w->bytes, &prop, 0));
  SCIMLOP_CU(h, a.MulticastBindMem(w->mc, 0, w->phys, 0, w->bytes, 0));
  CUmemAccessDesc acc;
  memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = h->device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  SCIMLOP_CU(h, a.MemAddressReserve(&w->uc_ptr, w->bytes, 0, 0, 0));
  SCIMLOP_CU(h, a.MemMap(w->uc_ptr, w->bytes, 0, w->phys, 0));
  SCIMLOP_CU(h, a.MemSetAccess(w->uc_ptr, w->bytes, &acc, 1));
  SCIMLOP_CU(h, a.MemAddressReserve(&w->mc_ptr, w->bytes, 0, 0, 0));
  SCIMLOP_CU(h, a.MemMap(w->mc_ptr, w->bytes, 0, w->mc, 0));
  SCIMLOP_CU(h, a.MemSetAccess(w->mc_ptr, w->bytes, &acc, 1));
  w->bound = true;
  *local_ptr = reinterpret_cast<void*>(w->uc_ptr);
  *mc_ptr = reinterpret_cast<void*>(w->mc_ptr);
  return SCIMLOP_OK;
}

extern "C" int scimlop_mc_free(scimlop_handle_t h, void* window) {
  if (!scimlop_valid(h)) return SCIMLOP_ERR_BAD_HANDLE;
  scimlop_mc_window* w = static_cast<scimlop_mc_window*>(window);
  if (!w) return SCIMLOP_OK;
  McApi& a = api();
  scimlop_device_guard guard(h->device);
  cudaDeviceSynchronize();
  if (w->bound) {
    CUdevice dev;
    a.DeviceGet(&dev, w->device);
    a.MemUnmap(w->mc_ptr, w->bytes);
    a.MemAddressFree(w->mc_ptr, w->bytes);
    a.MulticastUnbind(w->mc, dev, 0, w->bytes);
    a.MemUnmap(w->uc_ptr, w->bytes);
    a.MemAddressFree(w->uc_ptr, w->bytes);
    a.MemRelease(w->phys);
  }
  if (w->mc) a.MemRelease(w->mc);
  delete w;
  return SCIMLOP_OK;
}
