// vmm.h — one virtual address range whose physical pages are striped over several GPUs of a box (CUDA virtual memory
// management: cuMemAddressReserve / cuMemCreate / cuMemMap / cuMemSetAccess).  Used by the row-band mode for a single giant
// image: every workspace array is ONE pointer for all kernels, device k physically owns elements [k*S, (k+1)*S), and a
// kernel running on any GPU reaches the rest through NVLink peer loads / stores / atomics.  Driver entry points are fetched
// with cudaGetDriverEntryPoint, so the library has no link-time dependency on libcuda.
#pragma once
#include "rt.h"

namespace rt {

struct VmmArray {
    void *ptr = nullptr;
    size_t bytes = 0;          // reserved VA size
#ifndef VCB_HOSTSIM
    std::vector<CUmemGenericAllocationHandle> handles;
    std::vector<size_t> sizes;
#endif
};

#ifdef VCB_HOSTSIM
inline bool vmm_available() { return true; }
inline int vmm_alloc(VmmArray &a, size_t elems_per_dev, size_t elem_size, int ndev, const int *, size_t used_elems) {
    (void)used_elems;
    a.bytes = elems_per_dev * elem_size * (size_t)ndev;
    a.ptr = std::malloc(a.bytes ? a.bytes : 1);
    return a.ptr ? 0 : 1;
}
inline void vmm_free(VmmArray &a) { std::free(a.ptr); a.ptr = nullptr; a.bytes = 0; }
inline size_t vmm_granularity(int) { return 64; }
inline int enable_peer_access(int, const int *) { return 0; }
#else
struct VmmApi {
    CUresult (*getGranularity)(size_t *, const CUmemAllocationProp *, CUmemAllocationGranularity_flags) = nullptr;
    CUresult (*addressReserve)(CUdeviceptr *, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
    CUresult (*create)(CUmemGenericAllocationHandle *, size_t, const CUmemAllocationProp *, unsigned long long) = nullptr;
    CUresult (*map)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
    CUresult (*setAccess)(CUdeviceptr, size_t, const CUmemAccessDesc *, size_t) = nullptr;
    CUresult (*unmap)(CUdeviceptr, size_t) = nullptr;
    CUresult (*release)(CUmemGenericAllocationHandle) = nullptr;
    CUresult (*addressFree)(CUdeviceptr, size_t) = nullptr;
    bool ok = false;
};
inline VmmApi &vmm_api() {
    // initialised once, thread-safe (distinct contexts may be used from distinct threads)
    static VmmApi api = [] {
        VmmApi a;
        auto get = [](const char *name, void **fn) {
            cudaDriverEntryPointQueryResult q;
            return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess && *fn;
        };
        bool ok = true;
        ok &= get("cuMemGetAllocationGranularity", (void **)&a.getGranularity);
        ok &= get("cuMemAddressReserve", (void **)&a.addressReserve);
        ok &= get("cuMemCreate", (void **)&a.create);
        ok &= get("cuMemMap", (void **)&a.map);
        ok &= get("cuMemSetAccess", (void **)&a.setAccess);
        ok &= get("cuMemUnmap", (void **)&a.unmap);
        ok &= get("cuMemRelease", (void **)&a.release);
        ok &= get("cuMemAddressFree", (void **)&a.addressFree);
        if (!ok) cudaGetLastError();
        a.ok = ok;
        return a;
    }();
    return api;
}
inline bool vmm_available() { return vmm_api().ok; }
inline size_t vmm_granularity(int dev) {
    CUmemAllocationProp prop{};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    size_t g = 0;
    if (!vmm_api().ok || vmm_api().getGranularity(&g, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS || g == 0) g = 2u << 20;
    return g;
}
inline void vmm_free(VmmArray &a) {
    VmmApi &api = vmm_api();
    if (a.ptr) {
        size_t off = 0;
        for (size_t k = 0; k < a.handles.size(); ++k) {
            if (a.sizes[k]) { api.unmap((CUdeviceptr)a.ptr + off, a.sizes[k]); api.release(a.handles[k]); }
            off += a.sizes[k];
        }
        api.addressFree((CUdeviceptr)a.ptr, a.bytes);
    }
    a.ptr = nullptr; a.bytes = 0; a.handles.clear(); a.sizes.clear();
}
// elems_per_dev * elem_size must be a multiple of the granularity.  Chunks entirely beyond used_elems are still mapped
// (on the last device that owns data) so that the whole range is addressable.
inline int vmm_alloc(VmmArray &a, size_t elems_per_dev, size_t elem_size, int ndev, const int *devs, size_t used_elems) {
    VmmApi &api = vmm_api();
    if (!api.ok) return 1;
    const size_t chunk = elems_per_dev * elem_size;
    a.bytes = chunk * (size_t)ndev;
    CUdeviceptr base = 0;
    if (api.addressReserve(&base, a.bytes, 0, 0, 0) != CUDA_SUCCESS) return 2;
    a.ptr = (void *)base;
    int last_owner = devs[0];
    for (int k = 0; k < ndev; ++k) {
        int owner = devs[k];
        if ((size_t)k * elems_per_dev >= used_elems) owner = last_owner; else last_owner = owner;
        CUmemAllocationProp prop{};
        prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
        prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        prop.location.id = owner;
        CUmemGenericAllocationHandle h;
        if (api.create(&h, chunk, &prop, 0) != CUDA_SUCCESS) { a.handles.push_back(0); a.sizes.push_back(0); vmm_free(a); return 3; }
        if (api.map(base + (size_t)k * chunk, chunk, 0, h, 0) != CUDA_SUCCESS) { api.release(h); a.handles.push_back(0); a.sizes.push_back(0); vmm_free(a); return 4; }
        a.handles.push_back(h); a.sizes.push_back(chunk);
    }
    std::vector<CUmemAccessDesc> acc(ndev);
    for (int k = 0; k < ndev; ++k) {
        acc[k].location.type = CU_MEM_LOCATION_TYPE_DEVICE;
        acc[k].location.id = devs[k];
        acc[k].flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    }
    if (api.setAccess(base, a.bytes, acc.data(), (size_t)ndev) != CUDA_SUCCESS) { vmm_free(a); return 5; }
    return 0;
}
// plain cudaMalloc memory of `dev` becomes reachable from the kernels of the other devices
inline int enable_peer_access(int ndev, const int *devs) {
    for (int i = 0; i < ndev; ++i) {
        cudaSetDevice(devs[i]);
        for (int j = 0; j < ndev; ++j) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devs[i], devs[j]);
            if (!can) return 1;
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 2; }
            cudaGetLastError();
        }
        // stream-ordered (cudaMallocAsync) memory lives in the device's default pool, which has its own access list
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, devs[i]) == cudaSuccess) {
            std::vector<cudaMemAccessDesc> acc;
            for (int j = 0; j < ndev; ++j) {
                if (i == j) continue;
                cudaMemAccessDesc dsc{};
                dsc.location.type = cudaMemLocationTypeDevice;
                dsc.location.id = devs[j];
                dsc.flags = cudaMemAccessFlagsProtReadWrite;
                acc.push_back(dsc);
            }
            if (!acc.empty() && cudaMemPoolSetAccess(pool, acc.data(), acc.size()) != cudaSuccess) { cudaGetLastError(); return 3; }
        }
    }
    return 0;
}
#endif

}  // namespace rt
