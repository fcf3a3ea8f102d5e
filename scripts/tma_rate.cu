// tma_rate.cu -- dev micro-benchmark: how fast can one SM pull 128-byte-row boxes (the operand tiles of the tcgen05
// contractions: [box_rows x 64 bf16], SWIZZLE_128B) through the TMA unit, as a function of the box height, the number of
// boxes in flight and the source (HBM stream vs L2-resident)?   nvcc -O2 -gencode arch=compute_100a,code=sm_100a
// scripts/tma_rate.cu -o gpurun_out/tma_rate && gpurun_out/tma_rate
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t tx) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(tx) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma2d(void* dst, const void* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// every issuing thread (lane 0 of warps 0..issuers-1) owns `stages` ring slots and streams its share of the tiles
__global__ void __launch_bounds__(128) rate_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                   int rows, int cols, int box_rows, int stages, int issuers, int reps, int two_maps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bars[4][16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int w = 0; w < 4; ++w)
            for (int s = 0; s < 16; ++s) mbar_init(&bars[w][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp >= issuers || lane != 0) return;
    const uint32_t box_bytes = (uint32_t)box_rows * 128u;
    unsigned char* base = smem + (size_t)warp * stages * box_bytes;
    const int row_tiles = rows / box_rows, col_tiles = cols / 64;
    const int64_t total = (int64_t)row_tiles * col_tiles;
    const int nworkers = gridDim.x * issuers, me = blockIdx.x * issuers + warp;
    int issued = 0, waited = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int64_t t = me; t < total; t += nworkers) {
            if (issued - waited == stages) {
                const int s = waited % stages;
                while (!mbar_try(&bars[warp][s], (waited / stages) & 1)) {}
                ++waited;
            }
            const int s = issued % stages;
            // tile order: consecutive workers read consecutive column blocks of the same row tile
            const int rt = (int)(t / col_tiles), ct = (int)(t % col_tiles);
            mbar_expect(&bars[warp][s], box_bytes);
            tma2d(base + (size_t)s * box_bytes, (two_maps && (issued & 1)) ? &map_b : &map_a, ct * 64, rt * box_rows, &bars[warp][s]);
            ++issued;
        }
    }
    while (waited < issued) {
        const int s = waited % stages;
        while (!mbar_try(&bars[warp][s], (waited / stages) & 1)) {}
        ++waited;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* base, int64_t rows, int64_t cols, int box_rows, CUtensorMapL2promotion promo) {
    CUtensorMap m;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return m;
}

int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const int cols = 1536;
    const int64_t big_rows = 245760;            // 755 MB per plane: an HBM stream
    const int64_t small_rows = 8192;            // 25 MB per plane: L2 resident when repeated
    unsigned short *a, *b;
    CK(cudaMalloc(&a, big_rows * cols * 2));
    CK(cudaMalloc(&b, big_rows * cols * 2));
    CK(cudaMemset(a, 0, big_rows * cols * 2));
    CK(cudaMemset(b, 0, big_rows * cols * 2));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    printf("source,box_rows,stages,issuers,two_maps,promo,GB/s,B/clk/SM(1.9GHz)\n");
    const CUtensorMapL2promotion promos[2] = {CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B};
    for (int src = 0; src < 2; ++src) {
        const int64_t rows = src == 0 ? big_rows : small_rows;
        const int reps = src == 0 ? 1 : 30;
        for (int pi = 0; pi < 2; ++pi)
        for (int box_rows : {32, 64, 128, 256})
            for (int issuers : {1, 2, 4})
                for (int stages : {2, 4, 8}) {
                    for (int two = 0; two < 2; ++two) {
                        if ((size_t)issuers * stages * box_rows * 128 > 200 * 1024) continue;
                        if (pi == 1 && !(box_rows == 128 && two == 1)) continue;
                        CUtensorMap ma = make_map(enc, a, rows, cols, box_rows, promos[pi]);
                        CUtensorMap mb = make_map(enc, b, rows, cols, box_rows, promos[pi]);
                        const size_t smem = (size_t)issuers * stages * box_rows * 128 + 1024;
                        rate_kernel<<<sms, 128, smem>>>(ma, mb, (int)rows, cols, box_rows, stages, issuers, reps, two);   // warm-up
                        CK(cudaEventRecord(e0));
                        rate_kernel<<<sms, 128, smem>>>(ma, mb, (int)rows, cols, box_rows, stages, issuers, reps, two);
                        CK(cudaEventRecord(e1));
                        CK(cudaEventSynchronize(e1));
                        float ms = 0.f;
                        CK(cudaEventElapsedTime(&ms, e0, e1));
                        const double bytes = (double)rows * cols * 2.0 * reps;
                        const double gbs = bytes / (ms * 1e-3) / 1e9;
                        printf("%s,%d,%d,%d,%d,%s,%.0f,%.1f\n", src == 0 ? "hbm" : "l2", box_rows, stages, issuers, two, pi ? "128B" : "256B", gbs,
                               gbs * 1e9 / sms / 1.9e9);
                    }
                }
    }
    CK(cudaGetLastError());
    return 0;
}
