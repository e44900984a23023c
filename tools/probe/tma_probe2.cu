#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
typedef uint32_t u32;
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait0(u32 bar) {
    u32 done = 0;
    while (!done) asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
}
// mode 2: plain 1-D bulk copy (UBLKCP)
__global__ void k_bulk(const u32 *src, u32 *out, u32 bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *mbar = (unsigned long long *)(sm + 8192);
    const u32 bar = smem_addr(mbar);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(sm)), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
    wait0(bar);
    for (u32 i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = ((u32 *)sm)[i];
}
// mode 0/1: 2-D tensor copy, descriptor via grid_constant param (0) or global memory pointer (1)
__global__ void k_tma(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, int use_g, int cx, int cy, u32 *out, u32 bytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *mbar = (unsigned long long *)(sm + 8192);
    const u32 bar = smem_addr(mbar);
    if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long dp = use_g ? (unsigned long long)gmap : reinterpret_cast<unsigned long long>(&tmap);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_addr(sm)), "l"(dp), "r"(cx), "r"(cy), "r"(bar) : "memory");
    }
    wait0(bar);
    for (u32 i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = ((u32 *)sm)[i];
}
int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0; const int cx = argc > 2 ? atoi(argv[2]) : 0; const int cy = argc > 3 ? atoi(argv[3]) : 0; const int bw = argc > 4 ? atoi(argv[4]) : 64; const int ep = argc > 5 ? atoi(argv[5]) : 0;
    const u32 w = 1920, h = 1080;
    std::vector<u32> img((size_t)w * h);
    for (size_t i = 0; i < img.size(); ++i) img[i] = (u32)(i * 2654435761u) | 1u;
    u32 *d, *dout; cudaMalloc(&d, img.size() * 4); cudaMalloc(&dout, 8192);
    cudaMemcpy(d, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
    if (mode == 2) {
        k_bulk<<<1, 128, 8192 + 16>>>(d, dout, 4096);
        printf("bulk: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
        return 0;
    }
    CUtensorMap tm;
    cuuint64_t gdim[2] = {w, h}; cuuint64_t gstr[1] = {(cuuint64_t)w * 4};
    cuuint32_t box[2] = {(cuuint32_t)bw, 16}; cuuint32_t es[2] = {1, 1};
    typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    PFN enc = cuTensorMapEncodeTiled;
    if (ep) { void *fp = nullptr; cudaDriverEntryPointQueryResult q; cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q); enc = (PFN)fp; printf("entry point via runtime: %p vs linked %p\n", fp, (void*)cuTensorMapEncodeTiled); }
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode=%d\n", (int)r);
    CUtensorMap *gmap; cudaMalloc(&gmap, sizeof(tm)); cudaMemcpy(gmap, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    k_tma<<<1, 128, 8192 + 16>>>(tm, gmap, mode, cx, cy, dout, bw * 16 * 4);
    cudaError_t e = cudaDeviceSynchronize();
    printf("tma mode %d: %s\n", mode, cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<u32> out(bw * 16); cudaMemcpy(out.data(), dout, 4096, cudaMemcpyDeviceToHost);
        int bad = 0; for (int y = 0; y < 16; ++y) for (int x = 0; x < bw; ++x) { long gx = cx + x, gy = cy + y; u32 want = (gx >= 0 && gy >= 0 && gx < (long)w && gy < (long)h) ? img[(size_t)gy * w + gx] : 0; if (out[y * bw + x] != want) bad++; }
        printf("mismatches %d\n", bad);
    }
    return 0;
}
