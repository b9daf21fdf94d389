// vmm.hpp — replicated-factor memory for sharded fits: CUDA virtual-memory-management allocations that can be bound to
// an NVSwitch MULTICAST object, so that ONE store instruction (multimem.st) of the rank that solved a factor block
// lands in every rank's replica (NVLS), instead of N-1 unicast peer stores over the same NVLink egress.
//
// The driver API is reached through cudaGetDriverEntryPoint (no link-time dependency on libcuda: the library must
// still load — and export its symbols — on a box without a driver).
#pragma once

#include "common.cuh"

#include <cuda.h>

#include <mutex>

namespace b200 {

struct DriverApi {
    CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
    CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
    CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
    CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
    CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
    CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
    CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
    CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
    CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
    CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
    bool ok = false;

    static const DriverApi& get() {
        static DriverApi api;
        static std::once_flag once;
        std::call_once(once, [] {
            bool all = true;
            auto load = [&](const char* name, void** fn) {
                cudaDriverEntryPointQueryResult st{};
                if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !*fn) {
                    cudaGetLastError();
                    all = false;
                }
            };
            load("cuGetErrorString", reinterpret_cast<void**>(&api.GetErrorString));
            load("cuDeviceGet", reinterpret_cast<void**>(&api.DeviceGet));
            load("cuDeviceGetAttribute", reinterpret_cast<void**>(&api.DeviceGetAttribute));
            load("cuMemCreate", reinterpret_cast<void**>(&api.MemCreate));
            load("cuMemRelease", reinterpret_cast<void**>(&api.MemRelease));
            load("cuMemAddressReserve", reinterpret_cast<void**>(&api.MemAddressReserve));
            load("cuMemAddressFree", reinterpret_cast<void**>(&api.MemAddressFree));
            load("cuMemMap", reinterpret_cast<void**>(&api.MemMap));
            load("cuMemUnmap", reinterpret_cast<void**>(&api.MemUnmap));
            load("cuMemSetAccess", reinterpret_cast<void**>(&api.MemSetAccess));
            load("cuMemExportToShareableHandle", reinterpret_cast<void**>(&api.MemExportToShareableHandle));
            load("cuMemImportFromShareableHandle", reinterpret_cast<void**>(&api.MemImportFromShareableHandle));
            load("cuMulticastCreate", reinterpret_cast<void**>(&api.MulticastCreate));
            load("cuMulticastAddDevice", reinterpret_cast<void**>(&api.MulticastAddDevice));
            load("cuMulticastBindMem", reinterpret_cast<void**>(&api.MulticastBindMem));
            load("cuMulticastUnbind", reinterpret_cast<void**>(&api.MulticastUnbind));
            load("cuMulticastGetGranularity", reinterpret_cast<void**>(&api.MulticastGetGranularity));
            api.ok = all;
        });
        return api;
    }
};

inline void cu_check(CUresult r, const char* what, const char* file, int line) {
    if (r == CUDA_SUCCESS) return;
    const char* s = nullptr;
    const DriverApi& d = DriverApi::get();
    if (d.GetErrorString) d.GetErrorString(r, &s);
    throw CudaError(std::string(what) + " failed: " + (s ? s : "driver error") + " (" + std::to_string(static_cast<int>(r)) + ", " +
                    file + ":" + std::to_string(line) + ")");
}
#define B200_CU_CHECK(expr) ::b200::cu_check((expr), #expr, __FILE__, __LINE__)

// Does `device` support NVSwitch multicast (and the VMM API it needs)?
inline bool multicast_supported(int device) {
    const DriverApi& d = DriverApi::get();
    if (!d.ok) return false;
    CUdevice dev;
    int mc = 0, vmm = 0;
    if (d.DeviceGet(&dev, device) != CUDA_SUCCESS) return false;
    if (d.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, dev) != CUDA_SUCCESS) return false;
    if (d.DeviceGetAttribute(&vmm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, dev) != CUDA_SUCCESS) return false;
    return mc != 0 && vmm != 0;
}

// Binding / mapping granularity used for multicast-capable factors (the driver's RECOMMENDED multicast granularity:
// 512 MB on this pool's B200s — each factor costs at most that much padding of 180 GB).
inline size_t multicast_granularity(int num_devices) {
    const DriverApi& d = DriverApi::get();
    CUmulticastObjectProp mp{};
    mp.numDevices = static_cast<unsigned>(num_devices);
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mp.size = 1;
    size_t g = 0;
    B200_CU_CHECK(d.MulticastGetGranularity(&g, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED));
    return g;
}

// A replicated factor: fp32 [rows][KP]. Plain cudaMalloc by default (DeviceBuffer semantics: grow-only, contents not
// preserved); with want_vmm the next allocation is a VMM allocation (shareable as a POSIX file descriptor, rounded to
// `granularity`) that can be bound to a multicast object and mapped by peers.
struct FactorBuffer {
    float* ptr = nullptr;
    size_t count = 0;
    bool want_vmm = false;            // set before ensure(): allocate through the VMM API
    size_t granularity = 0;           // multicast granularity (want_vmm)
    int device = 0;
    // VMM state
    bool vmm = false;
    CUmemGenericAllocationHandle phys = 0;
    size_t phys_size = 0;

    FactorBuffer() = default;
    FactorBuffer(const FactorBuffer&) = delete;
    FactorBuffer& operator=(const FactorBuffer&) = delete;
    ~FactorBuffer() { release(); }
    size_t bytes() const { return count * sizeof(float); }

    void release() {
        if (ptr) {
            if (vmm) {
                const DriverApi& d = DriverApi::get();
                cudaSetDevice(device);
                cudaDeviceSynchronize();              // (cudaFree synchronises implicitly; unmapping does not)
                d.MemUnmap(reinterpret_cast<CUdeviceptr>(ptr), phys_size);
                d.MemAddressFree(reinterpret_cast<CUdeviceptr>(ptr), phys_size);
                d.MemRelease(phys);
            } else {
                cudaFree(ptr);
            }
        }
        ptr = nullptr;
        count = 0;
        vmm = false;
        phys = 0;
        phys_size = 0;
    }

    // true when the buffer was (re)allocated
    bool ensure(size_t n) {
        if (n == 0) n = 1;
        if (ptr && n <= count && vmm == want_vmm) return false;
        release();
        if (!want_vmm) {
            B200_CUDA_CHECK(cudaMalloc(&ptr, n * sizeof(float)));
            count = n;
            return true;
        }
        const DriverApi& d = DriverApi::get();
        B200_REQUIRE(d.ok && granularity > 0, "VMM allocation requested without driver support");
        const size_t size = (n * sizeof(float) + granularity - 1) / granularity * granularity;
        CUmemAllocationProp ap{};
        ap.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        ap.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        ap.location.id = device;
        ap.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
        CUmemGenericAllocationHandle h = 0;
        B200_CU_CHECK(d.MemCreate(&h, size, &ap, 0));
        CUdeviceptr va = 0;
        CUresult r = d.MemAddressReserve(&va, size, granularity, 0, 0);
        if (r != CUDA_SUCCESS) { d.MemRelease(h); B200_CU_CHECK(r); }
        r = d.MemMap(va, size, 0, h, 0);
        if (r != CUDA_SUCCESS) { d.MemAddressFree(va, size); d.MemRelease(h); B200_CU_CHECK(r); }
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = device;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = d.MemSetAccess(va, size, &acc, 1);
        if (r != CUDA_SUCCESS) { d.MemUnmap(va, size); d.MemAddressFree(va, size); d.MemRelease(h); B200_CU_CHECK(r); }
        ptr = reinterpret_cast<float*>(va);
        count = size / sizeof(float);
        phys = h;
        phys_size = size;
        vmm = true;
        return true;
    }
};

// A mapping of somebody else's physical allocation or of a multicast object into this device's address space.
struct MappedHandle {
    CUmemGenericAllocationHandle handle = 0;
    CUdeviceptr va = 0;
    size_t size = 0;
    bool bound = false;               // multicast object with this device's memory bound at offset 0
    int device = 0;
    float* fptr() const { return reinterpret_cast<float*>(va); }
    void map(CUmemGenericAllocationHandle h, size_t sz, size_t align, int dev) {
        const DriverApi& d = DriverApi::get();
        handle = h;
        size = sz;
        device = dev;
        va = 0;
        B200_CU_CHECK(d.MemAddressReserve(&va, size, align, 0, 0));
        CUresult r = d.MemMap(va, size, 0, handle, 0);
        if (r != CUDA_SUCCESS) { d.MemAddressFree(va, size); va = 0; B200_CU_CHECK(r); }
        CUmemAccessDesc acc{};
        acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc.location.id = dev;
        acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        r = d.MemSetAccess(va, size, &acc, 1);
        if (r != CUDA_SUCCESS) { d.MemUnmap(va, size); d.MemAddressFree(va, size); va = 0; B200_CU_CHECK(r); }
    }
    void release(bool release_handle = true) {
        const DriverApi& d = DriverApi::get();
        if (va) { d.MemUnmap(va, size); d.MemAddressFree(va, size); }
        if (handle) {
            if (bound) {
                CUdevice dev;
                if (d.DeviceGet(&dev, device) == CUDA_SUCCESS) d.MulticastUnbind(handle, dev, 0, size);
            }
            if (release_handle) d.MemRelease(handle);
        }
        handle = 0; va = 0; size = 0; bound = false;
    }
};

}  // namespace b200
