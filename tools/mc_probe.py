"""Is NVSwitch multicast (cuMulticast*) usable from this container?  Prints device attributes and tries to create a
2-device multicast object (no kernels, no peer traffic)."""
from cuda.bindings import driver as drv

def ck(r):
    err = r[0]
    if err != drv.CUresult.CUDA_SUCCESS:
        raise RuntimeError(str(err))
    return r[1] if len(r) == 2 else r[1:]

ck(drv.cuInit(0))
n = ck(drv.cuDeviceGetCount())
print("devices", n)
for d in range(min(n, 2)):
    dev = ck(drv.cuDeviceGet(d))
    for a in ("CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED", "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED",
              "CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED"):
        print(d, a, ck(drv.cuDeviceGetAttribute(getattr(drv.CUdevice_attribute, a), dev)))
try:
    prop = drv.CUmulticastObjectProp()
    prop.numDevices = min(n, 2)
    prop.size = 512 << 20
    prop.handleTypes = drv.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    prop.flags = 0
    gran = ck(drv.cuMulticastGetGranularity(prop, drv.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED))
    print("granularity", gran)
    ctx = ck(drv.cuDevicePrimaryCtxRetain(ck(drv.cuDeviceGet(0))))
    ck(drv.cuCtxSetCurrent(ctx))
    h = ck(drv.cuMulticastCreate(prop))
    print("cuMulticastCreate OK", h)
except Exception as e:
    print("multicast unavailable:", e)
