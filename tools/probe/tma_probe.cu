// standalone probe of the TMA staging used by k_edges_tile_tma (debug tool)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
typedef uint32_t u32;
constexpr int TW = 64, TH = 32, PW = TW + 4;
constexpr size_t PIX_BYTES = (size_t)(TH + 1) * PW * 4;
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int cx, int cy, int cz, u32 *out, u32 bytes, int rank) {
    extern __shared__ __align__(128) unsigned char sm[];
    u32 *pix = (u32 *)sm;
    unsigned long long *mbar = (unsigned long long *)(sm + PIX_BYTES);
    const u32 bar = smem_addr(mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        if (rank == 3)
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_addr(pix)), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(cx), "r"(cy), "r"(cz), "r"(bar) : "memory");
        else
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_addr(pix)), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(cx), "r"(cy), "r"(bar) : "memory");
    }
    u32 done = 0;
    while (!done) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
    }
    for (int i = threadIdx.x; i < (TH + 1) * PW; i += blockDim.x) out[i] = pix[i];
}
typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    const int BOXW = argc > 1 ? atoi(argv[1]) : PW; const int RANK = argc > 2 ? atoi(argv[2]) : 3; const int L2P = argc > 3 ? atoi(argv[3]) : 2;
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    PFN enc = (PFN)fp;
    int shapes[][3] = {{1920, 1080, 2}, {20, 13, 1}, {64, 32, 1}, {4, 4, 1}, {128, 3, 3}};
    for (auto &sh : shapes) {
        const u32 w = sh[0], h = sh[1], nimg = sh[2];
        std::vector<u32> img((size_t)w * h * nimg);
        for (size_t i = 0; i < img.size(); ++i) img[i] = (u32)(i * 2654435761u) | 1u;
        u32 *d, *dout; cudaMalloc(&d, img.size() * 4); cudaMalloc(&dout, PIX_BYTES);
        cudaMemcpy(d, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
        CUtensorMap tm;
        cuuint64_t gdim[3] = {w, h, nimg}; cuuint64_t gstr[2] = {(cuuint64_t)w * 4, (cuuint64_t)w * h * 4};
        cuuint32_t box[3] = {(cuuint32_t)BOXW, TH + 1, 1}; cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, RANK, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         (CUtensorMapL2promotion)L2P, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("shape %ux%ux%u encode=%d\n", w, h, nimg, (int)r);
        if (r != CUDA_SUCCESS) continue;
        int coords[][3] = {{63, 31, (int)nimg - 1}, {0, 0, 0}};
        for (auto &c : coords) {
            probe<<<1, 256, PIX_BYTES + 16>>>(tm, c[0], c[1], c[2], dout, (u32)(BOXW * (TH + 1) * 4), RANK);
            cudaError_t e = cudaDeviceSynchronize();
            printf("  coords %d,%d,%d -> %s\n", c[0], c[1], c[2], cudaGetErrorString(e));
            if (e != cudaSuccess) return 1;
            if (RANK == 2) c[2] = 0;
            std::vector<u32> out((TH + 1) * PW);
            cudaMemcpy(out.data(), dout, PIX_BYTES, cudaMemcpyDeviceToHost);
            int bad = 0;
            for (int ly = 0; ly < TH + 1; ++ly) for (int lx = 0; lx < PW; ++lx) {
                long gx = c[0] + lx, gy = c[1] + ly; u32 want = 0;
                if (gx >= 0 && gy >= 0 && gx < (long)w && gy < (long)h) want = img[(size_t)c[2] * w * h + gy * w + gx];
                if (lx < BOXW && out[ly * BOXW + lx] != want) bad++;
            }
            printf("  mismatches %d\n", bad);
        }
        cudaFree(d); cudaFree(dout);
    }
    return 0;
}
