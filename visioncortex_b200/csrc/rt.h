// rt.h — thin runtime layer under the kernels: CUDA runtime for the product build (nvcc, sm_100a); the test-only
// host simulator (tests/hostsim/cudasim.h, -DVCB_HOSTSIM) for debugging the same kernel sources without a GPU.
// The product library is only ever built WITHOUT VCB_HOSTSIM; there is no CPU execution path in it.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t i32;
typedef int64_t i64;

#ifdef VCB_HOSTSIM
#include "cudasim.h"
typedef int cudaError_t;
typedef struct SimStream *cudaStream_t;
#define cudaSuccess 0
#define VCB_SMEM(T, name) T *name = (T *)sim::smem_ptr()
#define VCB_GRID_SYNC() sim::tc.grid_bar->arrive_and_wait()
#define VCB_CLUSTER_SYNC() sim::tc.cluster_bar->arrive_and_wait()
template <class T> __device__ static inline T ldcg(const T *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
template <class T> __device__ static inline void stcg(T *p, T v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
__device__ static inline u8 ldcg8(const u8 *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
#else
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <execinfo.h>
#define VCB_SMEM(T, name)                                      \
    extern __shared__ __align__(128) unsigned char _vcb_smem[]; \
    T *name = reinterpret_cast<T *>(_vcb_smem)
#define VCB_GRID_SYNC() cooperative_groups::this_grid().sync()
#define VCB_CLUSTER_SYNC() cooperative_groups::this_cluster().sync()
template <class T> __device__ static __forceinline__ T ldcg(const T *p) { return __ldcg(p); }
template <class T> __device__ static __forceinline__ void stcg(T *p, T v) { __stcg(p, v); }
__device__ static __forceinline__ u8 ldcg8(const u8 *p) { return __ldcg(p); }
#endif

namespace rt {

enum LaunchMode { SEQ = 0, THR = 1, COOP = 2 };   // only meaningful to the simulator (COOP also selects a cooperative launch)

struct ProfEntry {
    const char *name;
#ifndef VCB_HOSTSIM
    cudaEvent_t e0, e1;
#endif
};

struct Stream {
    cudaStream_t s = nullptr;
    u64 launches = 0;
    bool profiling = false;
    std::vector<ProfEntry> prof;
    int last_error = 0;
    std::string last_error_text;
};

#ifdef VCB_HOSTSIM
inline int device_count() { return 1; }
inline int set_device(int) { return 0; }
inline int sm_count(int) { return 2; }
inline void *dev_malloc(size_t n) { return std::malloc(n ? n : 1); }
inline void dev_free(void *p) { std::free(p); }
inline void *host_malloc(size_t n) { return std::malloc(n ? n : 1); }
inline void host_free(void *p) { std::free(p); }
inline void *dev_malloc_async(Stream &, size_t n) { return std::malloc(n ? n : 1); }
inline void dev_free_async(Stream &, void *p) { std::free(p); }
inline int stream_create(Stream &) { return 0; }
inline void stream_destroy(Stream &) {}
struct Event { int dummy = 0; };
inline int event_create(Event &) { return 0; }
inline void event_destroy(Event &) {}
inline void event_record(Event &, Stream &) {}
inline void stream_wait(Stream &, Event &) {}
inline int stream_sync(Stream &) { return 0; }
inline int copy_h2d(Stream &, void *d, const void *h, size_t n) { std::memcpy(d, h, n); return 0; }
inline int copy_d2h(Stream &, void *h, const void *d, size_t n) { std::memcpy(h, d, n); return 0; }
inline int copy_d2d(Stream &, void *d, const void *s, size_t n) { std::memcpy(d, s, n); return 0; }
inline int fill(Stream &, void *d, int byte, size_t n) { std::memset(d, byte, n); return 0; }
inline const char *error_string(int) { return "sim"; }
inline int max_coop_blocks(const void *, int, size_t, int) { return 2; }

template <class... KArgs, class... Args>
inline void launch(Stream &st, const char *name, LaunchMode mode, void (*k)(KArgs...), dim3 grid, dim3 block,
                   size_t smem, Args... args) {
    (void)name;
    st.launches++;
    std::tuple<KArgs...> t(args...);
    static const bool force_thr = std::getenv("VCB_SIM_FORCE_THR") != nullptr;   // exercise real races in the simulator
    if (force_thr && mode == SEQ) mode = THR;
    sim::launch(grid, block, smem, (int)mode, [&] { std::apply(k, t); });
}
// thread-block cluster launch: `cluster` consecutive CTAs form a cluster (grid.x must be a multiple of it)
template <class... KArgs, class... Args>
inline void launch_cluster(Stream &st, const char *name, void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, unsigned cluster,
                           Args... args) {
    (void)name;
    st.launches++;
    std::tuple<KArgs...> t(args...);
    sim::launch(grid, block, smem, 2, [&] { std::apply(k, t); }, cluster);
}
#else
inline int device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
inline int set_device(int d) { return (int)cudaSetDevice(d); }
inline int sm_count(int d) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, d); return n; }
inline void *dev_malloc(size_t n) { void *p = nullptr; if (cudaMalloc(&p, n ? n : 1) != cudaSuccess) return nullptr; return p; }
inline int api_check(int e, const char *what, const void *p, size_t n);
inline void dev_free(void *p) { if (p) api_check((int)cudaFree(p), "dev_free", p, 0); }
inline void *host_malloc(size_t n) { void *p = nullptr; if (cudaMallocHost(&p, n ? n : 1) != cudaSuccess) return nullptr; return p; }
inline void host_free(void *p) { if (p) cudaFreeHost(p); }
// stream-ordered allocations from the device's default memory pool (kept resident: release threshold = max), so that
// per-call result buffers cost no cudaMalloc / cudaFree synchronisation after warm-up
inline void *dev_malloc_async(Stream &st, size_t n) {
    void *p = nullptr;
    if (cudaMallocAsync(&p, n ? n : 1, st.s) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
inline int api_check(int e, const char *what, const void *p, size_t n);
inline void dev_free_async(Stream &st, void *p) { if (p) api_check((int)cudaFreeAsync(p, st.s), "dev_free_async", p, 0); }
inline int stream_create(Stream &st) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return (int)cudaStreamCreateWithFlags(&st.s, cudaStreamNonBlocking);
}
inline void stream_destroy(Stream &st) { if (st.s) cudaStreamDestroy(st.s); st.s = nullptr; }
struct Event { cudaEvent_t e = nullptr; };
inline int event_create(Event &ev) { return (int)cudaEventCreateWithFlags(&ev.e, cudaEventDisableTiming); }
inline void event_destroy(Event &ev) { if (ev.e) cudaEventDestroy(ev.e); ev.e = nullptr; }
inline void event_record(Event &ev, Stream &st) { cudaEventRecord(ev.e, st.s); }
inline void stream_wait(Stream &st, Event &ev) { cudaStreamWaitEvent(st.s, ev.e, 0); }
inline int stream_sync(Stream &st) {
    cudaError_t e = cudaStreamSynchronize(st.s);
    if (e != cudaSuccess && !st.last_error) { st.last_error = (int)e; st.last_error_text = std::string("stream synchronize: ") + cudaGetErrorString(e); }
    return (int)e;
}
// VCB_CHECK_API=1 (debugging aid): report the runtime call that fails, with its arguments -- a failing copy / fill / free is
// otherwise only seen as the "last error" of the next kernel launch
inline int api_check(int e, const char *what, const void *p, size_t n) {
    static const bool on = std::getenv("VCB_CHECK_API") != nullptr;
    if (on && e != 0) {
        std::fprintf(stderr, "[vcb] %s(%p, %zu) failed: %s\n", what, p, n, cudaGetErrorString((cudaError_t)e));
        void *frames[16];
        backtrace_symbols_fd(frames, backtrace(frames, 16), 2);
    }
    return e;
}
inline int copy_h2d(Stream &st, void *d, const void *h, size_t n) { return api_check((int)cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, st.s), "copy_h2d", d, n); }
inline int copy_d2h(Stream &st, void *h, const void *d, size_t n) { return api_check((int)cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, st.s), "copy_d2h", d, n); }
inline int copy_d2d(Stream &st, void *d, const void *s, size_t n) { return api_check((int)cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st.s), "copy_d2d", d, n); }
inline int fill(Stream &st, void *d, int byte, size_t n) { return api_check((int)cudaMemsetAsync(d, byte, n, st.s), "fill", d, n); }
inline const char *error_string(int e) { return cudaGetErrorString((cudaError_t)e); }
inline int max_coop_blocks(const void *k, int threads, size_t smem, int dev) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem);
    return per_sm * sm_count(dev);
}

template <class... KArgs, class... Args>
inline void launch(Stream &st, const char *name, LaunchMode mode, void (*k)(KArgs...), dim3 grid, dim3 block,
                   size_t smem, Args... args) {
    st.launches++;
    ProfEntry pe{name, nullptr, nullptr};
    if (st.profiling) {
        cudaEventCreate(&pe.e0); cudaEventCreate(&pe.e1);
        cudaEventRecord(pe.e0, st.s);
    }
    if (mode == COOP) {
        std::tuple<KArgs...> t(args...);
        void *ptrs[sizeof...(KArgs) ? sizeof...(KArgs) : 1];
        size_t i = 0;
        std::apply([&](auto &...a) { ((ptrs[i++] = (void *)&a), ...); }, t);
        cudaError_t e = cudaLaunchCooperativeKernel((const void *)k, grid, block, ptrs, smem, st.s);
        if (e != cudaSuccess && !st.last_error) { st.last_error = (int)e; st.last_error_text = std::string(name) + ": " + cudaGetErrorString(e); }
    } else {
        if (smem > 48 * 1024) cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k<<<grid, block, smem, st.s>>>(KArgs(args)...);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && !st.last_error) { st.last_error = (int)e; st.last_error_text = std::string(name) + ": " + cudaGetErrorString(e); }
    }
    if (st.profiling) { cudaEventRecord(pe.e1, st.s); st.prof.push_back(pe); }
    static const bool sync_each = std::getenv("VCB_SYNC_EACH") != nullptr;   // debugging aid: name the faulting kernel
    if (sync_each) {
        cudaError_t e = cudaStreamSynchronize(st.s);
        if (e != cudaSuccess && !st.last_error) {
            st.last_error = (int)e; st.last_error_text = std::string(name) + ": " + cudaGetErrorString(e);
            std::fprintf(stderr, "[vcb] kernel %s failed: %s\n", name, cudaGetErrorString(e));
        }
    }
}
// thread-block cluster launch: `cluster` consecutive CTAs form a cluster (grid.x must be a multiple of it)
template <class... KArgs, class... Args>
inline void launch_cluster(Stream &st, const char *name, void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, unsigned cluster,
                           Args... args) {
    st.launches++;
    ProfEntry pe{name, nullptr, nullptr};
    if (st.profiling) { cudaEventCreate(&pe.e0); cudaEventCreate(&pe.e1); cudaEventRecord(pe.e0, st.s); }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st.s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (cluster > 8) cudaFuncSetAttribute((const void *)k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (smem > 48 * 1024) cudaFuncSetAttribute((const void *)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k, KArgs(args)...);
    if (e != cudaSuccess && !st.last_error) { st.last_error = (int)e; st.last_error_text = std::string(name) + ": " + cudaGetErrorString(e); }
    if (st.profiling) { cudaEventRecord(pe.e1, st.s); st.prof.push_back(pe); }
}
#endif

}  // namespace rt
